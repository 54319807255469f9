// Glue around the decoder stream kernel, on channel-last hi/lo maps (dec_stream.h): the outer-sum lift, the
// extra-channel 3x3 conv (distance encoding / x2-upsampled coarse map), the 1x1 output head, fp32 <-> map conversion.
// All HBM-bound elementwise kernels: 16-byte accesses, consecutive threads on consecutive 16-byte chunks of a pixel.
#include "common.h"
#include "dec_stream.h"
#include "tc_device.cuh"

namespace orca {
namespace {
using namespace tcdev;

__device__ __forceinline__ __nv_bfloat16* px_ptr(void* base, long long pixel, int C) {
  return static_cast<__nv_bfloat16*>(base) + pixel * 2 * C;
}

// mat[b][i][j][c] = x[b][i][c] + x[b][j][c]   (orca_modules.py:462, :783); one thread per (pixel, 8-channel chunk)
__global__ void ds_outer_sum_kernel(const float* __restrict__ xcl /*[B][S][C]*/, void* __restrict__ out, int S, int C, long long total) {
  const int C8 = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C8);
    const long long pix = i / C8;
    const int x = (int)(pix % S);
    const long long t = pix / S;
    const int y = (int)(t % S);
    const long long b = t / S;
    const float4* pi = reinterpret_cast<const float4*>(xcl + (b * S + y) * C + ch * 8);
    const float4* pj = reinterpret_cast<const float4*>(xcl + (b * S + x) * C + ch * 8);
    const float4 a0 = __ldg(pi), a1 = __ldg(pi + 1), b0 = __ldg(pj), b1 = __ldg(pj + 1);
    const float v[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w, a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
    __nv_bfloat16* p = px_ptr(out, pix, C) + ch * 8;
    split_store8(v, p, p + C);
  }
}

__device__ __forceinline__ float extra_src(const float* sb, long long sH, long long sW, int S, int mode, int yy, int xx) {
  if (yy < 0 || yy >= S || xx < 0 || xx >= S) return 0.f;
  if (mode == 0) return __ldg(sb + yy * sH + xx * sW);
  if (mode == 1) return __ldg(sb + (yy >> 1) * sH + (xx >> 1) * sW);
  const int n = S >> 1;  // bilinear x2, align_corners=False (nn.Upsample default, orca_modules.py:430)
  const float fy = fmaxf((yy + 0.5f) * 0.5f - 0.5f, 0.f), fx = fmaxf((xx + 0.5f) * 0.5f - 0.5f, 0.f);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, n - 1), x1 = min(x0 + 1, n - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float v00 = __ldg(sb + y0 * sH + x0 * sW), v01 = __ldg(sb + y0 * sH + x1 * sW);
  const float v10 = __ldg(sb + y1 * sH + x0 * sW), v11 = __ldg(sb + y1 * sH + x1 * sW);
  return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// n_extra-channel 3x3 conv -> 64 channels: 8 threads per pixel, each owning one 8-channel chunk of the output
__global__ void __launch_bounds__(256) ds_extra_conv_kernel(const float* __restrict__ src, long long sB, long long sC, long long sH,
                                                            long long sW, int n_extra, const float* __restrict__ w /*[n_extra][9][64]*/,
                                                            void* __restrict__ out, int S, int mode, long long total) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < n_extra * 9 * 64; i += blockDim.x) sw[i] = __ldg(w + i);
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long n_iter = (total + stride - 1) / stride;
  for (long long it = 0; it < n_iter; ++it) {  // uniform trip count: the shuffles below need whole warps
    const long long i0 = it * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i0 < total;
    const long long i = ok ? i0 : total - 8 + (i0 & 7);  // idle threads recompute the last pixel (no store)
    const int ch = (int)(i & 7);
    const long long pix = i >> 3;
    const int x = (int)(pix % S);
    const long long t = pix / S;
    const int y = (int)(t % S);
    const long long b = t / S;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int e = 0; e < n_extra; ++e) {
      const float* sb = src + b * sB + e * sC;
      // the 9 (upsampled) source values of this pixel are evaluated ONCE by its 8 threads (lane c: tap c, lane 0 also tap 8)
      // and exchanged with shuffles, instead of 9 bilinear samples per thread
      const float mine = extra_src(sb, sH, sW, S, mode, y + ch / 3 - 1, x + ch % 3 - 1);
      const float last = ch == 0 ? extra_src(sb, sH, sW, S, mode, y + 1, x + 1) : 0.f;
      const int base = (threadIdx.x & 31) & ~7;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const float v = tp < 8 ? __shfl_sync(0xffffffffu, mine, base + tp) : __shfl_sync(0xffffffffu, last, base);
        const float4* wr = reinterpret_cast<const float4*>(sw + (e * 9 + tp) * 64 + ch * 8);
        const float4 w0 = wr[0], w1 = wr[1];
        acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
        acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
      }
    }
    if (ok) {
      __nv_bfloat16* p = px_ptr(out, pix, 64) + ch * 8;
      split_store8(acc, p, p + 64);
    }
  }
}

// output head: 1x1 64->H (+BN) ReLU, 1x1 H->O (orca_modules.py:423-428: H = 5, O = 1; orca_leukemia.py:923-926:
// O = num_2d, H = max(num_2d, 5)).  8 threads per pixel (one 8-channel chunk each), partial sums reduced by shuffles.
template <int H>
__global__ void __launch_bounds__(256) ds_final_head_kernel(const void* __restrict__ in, const float* __restrict__ w0 /*[64][H]*/,
                                                            const float* __restrict__ b0, const float* __restrict__ w1 /*[H][O]*/,
                                                            const float* __restrict__ b1, float* __restrict__ tmp, int S, long long total, int O) {
  __shared__ float sw[64 * H], sb0[H], sw1[H * 8], sb1[8];
  for (int i = threadIdx.x; i < 64 * H; i += blockDim.x) sw[i] = w0[i];
  if (threadIdx.x < H) sb0[threadIdx.x] = b0[threadIdx.x];
  if (threadIdx.x < H * O) sw1[threadIdx.x] = w1[threadIdx.x];
  if (threadIdx.x < O) sb1[threadIdx.x] = b1[threadIdx.x];
  __syncthreads();
  const long long img = (long long)S * S;
  const long long n_iter = (total + (long long)gridDim.x * blockDim.x - 1) / ((long long)gridDim.x * blockDim.x);
  for (long long it = 0; it < n_iter; ++it) {  // uniform trip count: the shuffles below need whole 8-lane groups
    const long long i = it * gridDim.x * blockDim.x + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = i < total;
    const int ch = (int)(i & 7);
    const long long pix = ok ? (i >> 3) : 0;
    float h[H];
#pragma unroll
    for (int k = 0; k < H; ++k) h[k] = 0.f;
    if (ok) {
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(in) + pix * 128 + ch * 8;
      add_hilo8(v, p, p + 64);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < H; ++k) h[k] = fmaf(v[j], sw[(ch * 8 + j) * H + k], h[k]);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
      h[k] += __shfl_xor_sync(0xffffffffu, h[k], 1);
      h[k] += __shfl_xor_sync(0xffffffffu, h[k], 2);
      h[k] += __shfl_xor_sync(0xffffffffu, h[k], 4);
    }
    if (ok && ch == 0) {
#pragma unroll
      for (int k = 0; k < H; ++k) h[k] = fmaxf(h[k] + sb0[k], 0.f);
      const long long b = pix / img, r = pix - b * img;
      for (int o = 0; o < O; ++o) {
        float acc = sb1[o];
#pragma unroll
        for (int k = 0; k < H; ++k) acc = fmaf(h[k], sw1[k * O + o], acc);
        tmp[(b * O + o) * img + r] = acc;
      }
    }
  }
}

__global__ void ds_from_f32_kernel(const float* __restrict__ x, void* __restrict__ out, int C, long long total) {
  const int C8 = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C8);
    const long long pix = i / C8;
    const float4* s = reinterpret_cast<const float4*>(x + pix * C + ch * 8);
    const float4 a = __ldg(s), b = __ldg(s + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __nv_bfloat16* p = px_ptr(out, pix, C) + ch * 8;
    split_store8(v, p, p + C);
  }
}
__global__ void ds_to_f32_kernel(const void* __restrict__ in, float* __restrict__ x, int C, long long total) {
  const int C8 = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C8);
    const long long pix = i / C8;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const __nv_bfloat16* p = static_cast<const __nv_bfloat16*>(in) + pix * 2 * C + ch * 8;
    add_hilo8(v, p, p + C);
    float4* d = reinterpret_cast<float4*>(x + pix * C + ch * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

unsigned grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace

int ds_outer_sum(const float* xcl, DMap* out, cudaStream_t s) {
  const long long total = (long long)out->nb * out->S * out->S * (out->C / 8);
  ds_outer_sum_kernel<<<grid_for(total, 256), 256, 0, s>>>(xcl, out->p, out->S, out->C, total);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

int ds_extra_conv(const float* src, int64_t sB, int64_t sC, int64_t sH, int64_t sW, int n_extra, const float* w_extra, DMap* out,
                  int mode, cudaStream_t s) {
  if (out->C != 64 || n_extra < 1 || n_extra > 8) { set_error("ds_extra_conv: C != 64 or bad extra-channel count"); return ORCA_B200_EINVAL; }
  const long long total = (long long)out->nb * out->S * out->S * 8;
  ds_extra_conv_kernel<<<grid_for(total, 256), 256, (size_t)n_extra * 9 * 64 * sizeof(float), s>>>(src, sB, sC, sH, sW, n_extra, w_extra,
                                                                                                     out->p, out->S, mode, total);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

int ds_final_head_tmp(const DMap& in, const ConvLayer& f0, const ConvLayer& f1, float* tmp, cudaStream_t s) {
  if (in.C != 64 || !final_head_ok(f0, f1)) { set_error("ds_final_head: bad layers (64->%d->%d)", f0.c_out, f1.c_out); return ORCA_B200_EINVAL; }
  const long long total = (long long)in.nb * in.S * in.S * 8;
  const unsigned g = grid_for(total, 256);
  const int O = f1.c_out;
  switch (f0.c_out) {
    case 5: ds_final_head_kernel<5><<<g, 256, 0, s>>>(in.p, f0.w, f0.b, f1.w, f1.b, tmp, in.S, total, O); break;
    case 6: ds_final_head_kernel<6><<<g, 256, 0, s>>>(in.p, f0.w, f0.b, f1.w, f1.b, tmp, in.S, total, O); break;
    case 7: ds_final_head_kernel<7><<<g, 256, 0, s>>>(in.p, f0.w, f0.b, f1.w, f1.b, tmp, in.S, total, O); break;
    default: ds_final_head_kernel<8><<<g, 256, 0, s>>>(in.p, f0.w, f0.b, f1.w, f1.b, tmp, in.S, total, O); break;
  }
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

int ds_from_f32(const float* x, DMap* out, cudaStream_t s) {
  const long long total = (long long)out->nb * out->S * out->S * (out->C / 8);
  ds_from_f32_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, out->p, out->C, total);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}
int ds_to_f32(const DMap& in, float* x, cudaStream_t s) {
  const long long total = (long long)in.nb * in.S * in.S * (in.C / 8);
  ds_to_f32_kernel<<<grid_for(total, 256), 256, 0, s>>>(in.p, x, in.C, total);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
