// tcgen05 tensor-core convolution path (placeholder until the UMMA kernels land):
// reports "unsupported" for every layer so that conv() routes to the SIMT kernels.
#include <vector>
#include "common.h"

namespace orca {
int tc_pack_layer(ConvLayer&, const float*, std::vector<void*>&) { return ORCA_B200_OK; }
bool tc_supported(const ConvLayer&, const ConvCall&) { return false; }
int conv_tc(const ConvLayer&, const ConvCall&, cudaStream_t) {
  set_error("conv_tc: not built");
  return ORCA_B200_EUNSUPPORTED;
}
}  // namespace orca
