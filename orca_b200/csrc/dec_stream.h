// Decoder "stream" path (conv2d_stream.cu, dec_glue.cu): every dilated 3x3 convolution of one Decoder / Decoder_1m
// call (orca_modules.py:22-488, :499-800) in ONE persistent tcgen05 kernel, fed by tensor-map TMA
// (cp.async.bulk.tensor, SWIZZLE_128B) and synchronised by per-image-row completion flags instead of grid barriers.
//
// ---- map layout in HBM ("DMap") -------------------------------------------------------------------------
// A (nb, C, S, S) activation is ONE buffer [nb][S][S][2*C] bf16, channel-last: per pixel the C "hi" values then the
// C "lo" values (x = hi + lo, |x - hi - lo| <= 2^-17 |x|; three tensor-core products per algorithmic product,
// DESIGN.md section 3).  No padding pixels: image borders are the tensor map's out-of-bounds zero fill.
//   C = 32: one 128-byte row per pixel [hi 32 | lo 32]          (the bottleneck tensors)
//   C = 64: 256 bytes per pixel [hi 64 | lo 64]                  (the residual stream)
//   C = 128: 512 bytes per pixel [hi 128 | lo 128]               (the outer-sum lift; read as two 64-channel halves)
// A 64-channel hi (or lo) slice of one pixel is exactly one 128-byte K-major SWIZZLE_128B operand row, so a run of
// R = 128 + 2d pixels of an image row is one TMA box {64 ch, R px, 1 row} and the three dx taps of a dilated conv are
// the shared-memory descriptor start address + dx*d*128 B (tests/cuda/umma_sw128_probe.cu).
#pragma once
#include <stdint.h>
#include <vector>

#include "common.h"

namespace orca {

struct DMap {
  void* p = nullptr;
  int nb = 0, C = 0, S = 0;
};
inline size_t dmap_bytes(int nb, int C, int S) { return (size_t)nb * S * S * C * 4; }
inline DMap dmap_make(void* p, int nb, int C, int S) {
  DMap m;
  m.p = p; m.nb = nb; m.C = C; m.S = S;
  return m;
}

// weights of one 3x3 layer for the stream kernel: per 64-channel K half (c_in = 128 has two, else one) nine tap images
// [k-chunk][Bh rows | Bl rows][8] bf16 (K-major SWIZZLE_NONE, the [Bh;Bl] concat of DESIGN.md section 5)
bool ds_layer_eligible(const ConvLayer& L);
int ds_pack_layer(ConvLayer& L, const float* w_folded /*[tap][c_in][c_out]*/, std::vector<void*>& allocs);

class DecStream {
 public:
  DecStream();
  ~DecStream();
  DecStream(const DecStream&) = delete;
  DecStream& operator=(const DecStream&) = delete;
  // Append out = act(conv3x3_dil(in[:, k_half]) + bias) + res.  k_half: -1 = all input channels (c_in 32 / 64);
  // 0 / 1 = the lower / upper 64 input channels of a 128-channel `in` (the caller chains the two halves through
  // `res`; use_bias = 0 on the first).  in / res / out must be distinct buffers.
  int add(const ConvLayer& L, int k_half, int use_bias, const DMap& in, const DMap* res, DMap* out, int relu);
  // Upload tables into `scratch` (DEVICE, >= scratch_bytes(max layers, nb, S)) and launch on `s`.
  int run(void* scratch, size_t scratch_bytes, cudaStream_t s);
  int size() const;
  double flop() const;
  static size_t scratch_bytes(int max_layers, int nb, int S);

 private:
  struct Impl;
  Impl* impl;
};

// debugging (tests/cuda/dec_stream_test.cu): with the switch on, a wait that times out inside the kernel is recorded
// ({flag, source line, block, warp, ...}) and the kernel drains instead of trapping
int ds_debug_enable(int on);
int ds_debug_read(unsigned int out[184]);  // {abort flag, -, -, -, (source line << 8 | layer, count, a block) x 60}

// glue on DMaps (dec_glue.cu)
int ds_outer_sum(const float* xcl /*[nb][S][128]*/, DMap* out /*C = 128*/, cudaStream_t s);
int ds_extra_conv(const float* src, int64_t sB, int64_t sC, int64_t sH, int64_t sW, int n_extra, const float* w_extra,
                  DMap* out /*C = 64*/, int mode, cudaStream_t s);
int ds_final_head_tmp(const DMap& in, const ConvLayer& f0, const ConvLayer& f1, float* tmp /*[nb][O][S][S]*/, cudaStream_t s);
// debugging / tests: fp32 channel-last [nb][S][S][C] <-> DMap
int ds_from_f32(const float* x, DMap* out, cudaStream_t s);
int ds_to_f32(const DMap& in, float* x, cudaStream_t s);

}  // namespace orca
