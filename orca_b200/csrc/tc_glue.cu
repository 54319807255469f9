// Small HBM-bound kernels on 1D chunk planes used by the U-net encoders on the tcgen05 path:
// nearest x2 upsample (nn.Upsample(scale_factor=2), orca_modules.py:1079) and fp32 channel-last -> planes.
#include "common.h"
#include "tc.h"
#include "tc_device.cuh"

namespace orca {
namespace {
using namespace tcdev;

__global__ void upsample2_planes_kernel(const uint4* __restrict__ in_hi, const uint4* __restrict__ in_lo,
                                        uint4* __restrict__ out_hi, uint4* __restrict__ out_lo, int planes, int n_out,
                                        int npad_in, int npad_out) {
  const long long total = (long long)planes * npad_out;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pl = (int)(i / npad_out), r = (int)(i - (long long)pl * npad_out);
    const int l = r - 4;
    uint4 h = make_uint4(0, 0, 0, 0), lo = h;
    if (l >= 0 && l < n_out) {
      const size_t src = (size_t)pl * npad_in + (l >> 1) + 4;
      h = __ldg(in_hi + src);
      lo = __ldg(in_lo + src);
    }
    out_hi[i] = h;  // pad rows are written as zeros
    out_lo[i] = lo;
  }
}

__global__ void from_channel_last_kernel(const float* __restrict__ xcl, __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo, int nb, int C, int n, int npad) {
  const int C8 = C / 8;
  const long long total = (long long)nb * C8 * npad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % npad);
    const long long t = i / npad;
    const int ch = (int)(t % C8);
    const long long b = t / C8;
    const int l = r - 4;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (l >= 0 && l < n) {
      const float4* p = reinterpret_cast<const float4*>(xcl + ((size_t)b * n + l) * C + ch * 8);
      const float4 a = __ldg(p), c = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    }
    split_store8(v, hi + (size_t)i * 8, lo + (size_t)i * 8);
  }
}

unsigned grid_for(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)(g < 1 ? 1 : g);
}
}  // namespace

int tc_upsample2_planes(const TcAct& in, TcAct* out, cudaStream_t s) {
  if (out->n != in.n * 2 || out->C != in.C || out->nb != in.nb) { set_error("tc_upsample2_planes: bad geometry"); return ORCA_B200_EINVAL; }
  const int planes = in.nb * (in.C / 8);
  upsample2_planes_kernel<<<grid_for((long long)planes * out->npad), 256, 0, s>>>(
      static_cast<const uint4*>(in.hi), static_cast<const uint4*>(in.lo), static_cast<uint4*>(out->hi),
      static_cast<uint4*>(out->lo), planes, (int)out->n, (int)in.npad, (int)out->npad);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

int tc_from_channel_last(const float* xcl, TcAct* out, cudaStream_t s) {
  if (out->C % 8) { set_error("tc_from_channel_last: C %% 8"); return ORCA_B200_EINVAL; }
  from_channel_last_kernel<<<grid_for((long long)out->nb * (out->C / 8) * out->npad), 256, 0, s>>>(
      xcl, static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo), out->nb, out->C, (int)out->n,
      (int)out->npad);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
