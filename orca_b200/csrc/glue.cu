// HBM-bound glue kernels of the Orca forward path (coalesced, vectorised; no tensor cores):
// residual-add + max-pool, nearest upsample, layout change, outer-sum lift, the odd
// "extra channel" 3x3 conv of the Decoder combiners (distance encoding / upsampled coarse
// prediction), the 1x1 output head + symmetrisation, and the background block-mean + log.
#include "common.h"

namespace orca {

static inline unsigned blocks_for(long long n, int per_block) {
  long long b = (n + per_block - 1) / per_block;
  return (unsigned)(b < 1 ? 1 : b);
}

// ---- add + maxpool (MaxPool1d(p) of `out_k + lout_k`, orca_modules.py:936-948) ------------
__global__ void add_maxpool1d_kernel(const float4* __restrict__ a, const float4* __restrict__ b,
                                     float4* __restrict__ out, long long n_out_vec, int C4, int p) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C4;  // global output row (b*L_out + l)
    const int c = (int)(i - row * C4);
    const long long base = row * p * C4 + c;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int k = 0; k < p; ++k) {
      float4 v = __ldg(a + base + (long long)k * C4);
      if (b) {
        const float4 w = __ldg(b + base + (long long)k * C4);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
    out[i] = m;
  }
}

int add_maxpool1d(const float* a, const float* bb, float* out, int B, int64_t L_in, int C, int p,
                  cudaStream_t s) {
  if (C % 4 || L_in % p) { set_error("add_maxpool1d: bad geometry L=%lld C=%d p=%d", (long long)L_in, C, p); return ORCA_B200_EINVAL; }
  const long long n = (long long)B * (L_in / p) * (C / 4);
  if (n == 0) return ORCA_B200_OK;
  unsigned grid = blocks_for(n, 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  add_maxpool1d_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(bb),
                                            reinterpret_cast<float4*>(out), n, C / 4, p);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ---- nearest x2 upsample along L (nn.Upsample(scale_factor=2), orca_modules.py:1079) ------
__global__ void upsample2_1d_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                    long long n_out_vec, int C4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / C4;  // b*2L + l ; 2L even so row/2 = b*L + l/2
    const int c = (int)(i - row * C4);
    out[i] = __ldg(in + (row >> 1) * C4 + c);
  }
}

int upsample2_1d(const float* in, float* out, int B, int64_t L_in, int C, cudaStream_t s) {
  if (C % 4) { set_error("upsample2_1d: C %% 4"); return ORCA_B200_EINVAL; }
  const long long n = (long long)B * L_in * 2 * (C / 4);
  if (n == 0) return ORCA_B200_OK;
  unsigned grid = blocks_for(n, 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  upsample2_1d_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n, C / 4);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ---- strided (B, C, L) -> channel-last [B][L][C] via a 32x32 shared-memory transpose -----------
__global__ void to_channel_last_kernel(const float* __restrict__ x, long long sB, long long sC,
                                       long long sL, float* __restrict__ out, int C, long long L) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long l0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* xb = x + (long long)b * sB;
  if (sC == 1) {  // already channel-last in memory: read rows of channels
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const long long l = l0 + r;
      const int c = c0 + threadIdx.x;
      if (l < L && c < C) out[((long long)b * L + l) * C + c] = __ldg(xb + l * sL + c);
    }
    return;
  }
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r: channel, x: position
    const int c = c0 + r;
    const long long l = l0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && l < L) ? __ldg(xb + c * sC + l * sL) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r: position, x: channel
    const long long l = l0 + r;
    const int c = c0 + threadIdx.x;
    if (l < L && c < C) out[((long long)b * L + l) * C + c] = tile[threadIdx.x][r];
  }
}

int to_channel_last(const float* x, int64_t sB, int64_t sC, int64_t sL, float* out, int B, int C,
                    int64_t L, cudaStream_t s) {
  if (B <= 0 || L <= 0) return ORCA_B200_OK;
  dim3 grid((unsigned)((L + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B), block(32, 8);
  to_channel_last_kernel<<<grid, block, 0, s>>>(x, sB, sC, sL, out, C, L);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ---- outer-sum lift: mat[b][i][j][c] = x[b][c][i] + x[b][c][j]  (orca_modules.py:462, :783) -----
__global__ void outer_sum_kernel(const float* __restrict__ xcl /*[B][S][C] channel-last*/,
                                 float4* __restrict__ mat, int S, int C4, long long n_vec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    const long long pix = i / C4;  // (b*S + ii)*S + jj
    const int jj = (int)(pix % S);
    const long long bi = pix / S;  // b*S + ii
    const long long b = bi / S;
    const float4 u = __ldg(reinterpret_cast<const float4*>(xcl) + bi * C4 + c);
    const float4 v = __ldg(reinterpret_cast<const float4*>(xcl) + (b * S + jj) * C4 + c);
    mat[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

int outer_sum(const float* xcl, float* mat, int B, int C, int S, cudaStream_t s) {
  if (C % 4) { set_error("outer_sum: C %% 4"); return ORCA_B200_EINVAL; }
  const long long n = (long long)B * S * S * (C / 4);
  if (n == 0) return ORCA_B200_OK;
  unsigned grid = blocks_for(n, 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  outer_sum_kernel<<<grid, 256, 0, s>>>(xcl, reinterpret_cast<float4*>(mat), S, C / 4, n);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ---- extra-channel 3x3 conv (dilation 1, zero padding) -------------------------------------------
// The Decoder's 129->64 and 65->64 convs (orca_modules.py:431-451) are evaluated as an aligned
// 128/64-channel implicit GEMM plus this single-channel term, which enters the GEMM epilogue as
// a residual.  Source value at full resolution (yy, xx):
//   mode 0: src[yy][xx]                                   (distenc, orca_modules.py:463)
//   mode 1: src[yy/2][xx/2]                               (nearest x2, nn.Upsample default)
//   mode 2: bilinear x2, align_corners=False              (orca_models.py:45; F.interpolate semantics:
//           s = max((d + 0.5)/2 - 0.5, 0), i0 = floor(s), i1 = min(i0+1, n-1), lambda = s - i0)
__device__ __forceinline__ float extra_src(const float* sb, long long sH, long long sW, int S, int mode,
                                           int yy, int xx) {
  if (yy < 0 || yy >= S || xx < 0 || xx >= S) return 0.f;
  if (mode == 0) return __ldg(sb + yy * sH + xx * sW);
  if (mode == 1) return __ldg(sb + (yy >> 1) * sH + (xx >> 1) * sW);
  const int n = S >> 1;
  float fy = fmaxf((yy + 0.5f) * 0.5f - 0.5f, 0.f), fx = fmaxf((xx + 0.5f) * 0.5f - 0.5f, 0.f);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, n - 1), x1 = min(x0 + 1, n - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float v00 = __ldg(sb + y0 * sH + x0 * sW), v01 = __ldg(sb + y0 * sH + x1 * sW);
  const float v10 = __ldg(sb + y1 * sH + x0 * sW), v11 = __ldg(sb + y1 * sH + x1 * sW);
  // same association as ATen's upsample_bilinear2d: (1-ly)*((1-lx)*v00 + lx*v01) + ly*((1-lx)*v10 + lx*v11)
  return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

__global__ void extra_channel_conv_kernel(const float* __restrict__ src, long long sB, long long sC, long long sH,
                                          long long sW, int n_extra, const float* __restrict__ w /*[n_extra][9][c_out]*/,
                                          float* __restrict__ out, int S, int c_out, int mode,
                                          long long n_pix) {
  // thread = (pixel, 4 output channels); the n_extra source channels (1 in orca_modules, num_2d in
  // orca_leukemia.py:931-951) are summed in channel order
  const int C4 = c_out >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix * C4;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    const long long pix = i / C4;
    const int xx = (int)(pix % S);
    const long long by = pix / S;
    const int yy = (int)(by % S);
    const long long b = by / S;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < n_extra; ++e) {
      const float* sb = src + b * sB + e * sC;
      const float* we = w + (long long)e * 9 * c_out;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float v = extra_src(sb, sH, sW, S, mode, yy + t / 3 - 1, xx + t % 3 - 1);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(we + t * c_out) + c);
        acc.x = fmaf(v, wv.x, acc.x); acc.y = fmaf(v, wv.y, acc.y);
        acc.z = fmaf(v, wv.z, acc.z); acc.w = fmaf(v, wv.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

int extra_channel_conv(const float* src, int64_t sB, int64_t sC, int64_t sH, int64_t sW, int n_extra,
                       const float* w_extra, float* out, int B, int S, int c_out, int mode, cudaStream_t s) {
  if (c_out % 4 || (mode != 0 && (S & 1)) || n_extra < 1) { set_error("extra_channel_conv: bad geometry"); return ORCA_B200_EINVAL; }
  const long long n_pix = (long long)B * S * S;
  if (n_pix == 0) return ORCA_B200_OK;
  unsigned grid = blocks_for(n_pix * (c_out / 4), 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  extra_channel_conv_kernel<<<grid, 256, 0, s>>>(src, sB, sC, sH, sW, n_extra, w_extra, out, S, c_out, mode, n_pix);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ---- output head: 1x1 64->H (+BN folded) + ReLU + 1x1 H->O, then symmetrise ------------------------
// (orca_modules.py:423-428 and :487-488: H = 5, O = 1; orca_leukemia.py:923-926: O = num_2d, H = max(num_2d, 5)).
// One warp per pixel: each lane holds 2 of the 64 channels, the H hidden units are reduced with warp shuffles.
// tmp is [B][O][S][S].
template <int H>
__global__ void final_head_kernel(const float* __restrict__ in, const float* __restrict__ w0 /*[64][H]*/,
                                  const float* __restrict__ b0, const float* __restrict__ w1 /*[H][O]*/,
                                  const float* __restrict__ b1, float* __restrict__ tmp, long long n_pix,
                                  long long img /*S*S*/, int O) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float wa[H], wb[H];
#pragma unroll
  for (int k = 0; k < H; ++k) {
    wa[k] = __ldg(w0 + (2 * lane) * H + k);
    wb[k] = __ldg(w0 + (2 * lane + 1) * H + k);
  }
  for (long long p = warp; p < n_pix; p += nwarps) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(in + p * 64) + lane);
    float h[H];
#pragma unroll
    for (int k = 0; k < H; ++k) h[k] = fmaf(v.x, wa[k], v.y * wb[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < H; ++k) h[k] += __shfl_xor_sync(0xffffffffu, h[k], o);
    if (lane < O) {  // lane o evaluates output channel o
      const long long b = p / img, q = p - b * img;
      float r = __ldg(b1 + lane);
#pragma unroll
      for (int k = 0; k < H; ++k) r = fmaf(fmaxf(h[k] + __ldg(b0 + k), 0.f), __ldg(w1 + k * O + lane), r);
      tmp[(b * O + lane) * img + q] = r;
    }
  }
}

__global__ void symmetrise_kernel(const float* __restrict__ tmp, float* __restrict__ out, int S) {
  __shared__ float t[32][33];
  const int b = blockIdx.z;
  const float* tb = tmp + (long long)b * S * S;
  float* ob = out + (long long)b * S * S;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // load the transposed block (j0.., i0..)
    const int jj = j0 + r, ii = i0 + threadIdx.x;
    t[r][threadIdx.x] = (jj < S && ii < S) ? tb[(long long)jj * S + ii] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int ii = i0 + r, jj = j0 + threadIdx.x;
    if (ii < S && jj < S) ob[(long long)ii * S + jj] = 0.5f * tb[(long long)ii * S + jj] + 0.5f * t[threadIdx.x][r];
  }
}

int symmetrise(const float* tmp, float* out, int B, int S, cudaStream_t s) {
  if (B <= 0 || S <= 0) return ORCA_B200_OK;
  dim3 g2((S + 31) / 32, (S + 31) / 32, B), b2(32, 8);
  symmetrise_kernel<<<g2, b2, 0, s>>>(tmp, out, S);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

bool final_head_ok(const ConvLayer& f0, const ConvLayer& f1) {
  return f0.c_in == 64 && f0.c_out >= 5 && f0.c_out <= 8 && f1.c_in == f0.c_out && f1.c_out >= 1 && f1.c_out <= 8 &&
         f0.kh == 1 && f0.kw == 1 && f1.kh == 1 && f1.kw == 1;
}

int final_head(const float* in, const ConvLayer& f0, const ConvLayer& f1, float* tmp, float* out, int B,
               int S, cudaStream_t s) {
  if (!final_head_ok(f0, f1)) { set_error("final_head: bad layers (64->%d->%d)", f0.c_out, f1.c_out); return ORCA_B200_EINVAL; }
  const long long n_pix = (long long)B * S * S, img = (long long)S * S;
  if (n_pix == 0) return ORCA_B200_OK;
  unsigned grid = blocks_for(n_pix, 8);
  if (grid > 148u * 16u) grid = 148u * 16u;
  const int O = f1.c_out;
  switch (f0.c_out) {
    case 5: final_head_kernel<5><<<grid, 256, 0, s>>>(in, f0.w, f0.b, f1.w, f1.b, tmp, n_pix, img, O); break;
    case 6: final_head_kernel<6><<<grid, 256, 0, s>>>(in, f0.w, f0.b, f1.w, f1.b, tmp, n_pix, img, O); break;
    case 7: final_head_kernel<7><<<grid, 256, 0, s>>>(in, f0.w, f0.b, f1.w, f1.b, tmp, n_pix, img, O); break;
    default: final_head_kernel<8><<<grid, 256, 0, s>>>(in, f0.w, f0.b, f1.w, f1.b, tmp, n_pix, img, O); break;
  }
  ORCA_LAUNCH_OK();
  return symmetrise(tmp, out, B * O, S, s);
}

// ---- Net.final_1d tail: Conv1d(128 -> num_1d, k=1) + Sigmoid (orca_modules.py:1824-1830) ---------
__global__ void head_1d_sigmoid_kernel(const float* __restrict__ in, const float* __restrict__ w /*[128][K]*/,
                                       const float* __restrict__ bias, float* __restrict__ out, int S, int K,
                                       long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i % S);
    const long long bk = i / S;
    const int k = (int)(bk % K);
    const long long b = bk / K;
    const float* v = in + (b * S + l) * 128;
    float acc = __ldg(bias + k);
    for (int c = 0; c < 128; ++c) acc = fmaf(__ldg(v + c), __ldg(w + c * K + k), acc);
    out[i] = 1.f / (1.f + expf(-acc));
  }
}

int head_1d_sigmoid(const float* in, const ConvLayer& L, float* out, int B, int S, cudaStream_t s) {
  const long long n = (long long)B * L.c_out * S;
  if (n == 0) return ORCA_B200_OK;
  head_1d_sigmoid_kernel<<<blocks_for(n, 128), 128, 0, s>>>(in, L.w, L.b, out, S, L.c_out, n);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

__global__ void copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

int copy_f32(const float* src, float* dst, int64_t n, cudaStream_t s) {
  if (n <= 0) return ORCA_B200_OK;
  unsigned grid = blocks_for(n, 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  copy_f32_kernel<<<grid, 256, 0, s>>>(src, dst, n);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ---- background level: block nanmean (inner axis first, then outer) + log -------------------------
// orca_predict.py:724-737 (np.nanmean(np.nanmean(reshape(...), axis=4), axis=2)) then :693-697 (log).
// One warp per output cell; lanes stride the f columns of each block row (coalesced fp64 reads),
// warp-shuffle reduction per row, then the row means are averaged.
__global__ void background_level_kernel(const double* __restrict__ nm, long long n, long long r0,
                                        long long f, int S, int flip, float* __restrict__ out,
                                        double* __restrict__ out_mean) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (long long)S * S) return;
  const int i = (int)(warp / S), j = (int)(warp % S);
  double rsum = 0.0;
  long long rcnt = 0;
  for (long long r = 0; r < f; ++r) {
    const double* row = nm + (r0 + (long long)i * f + r) * n + r0 + (long long)j * f;
    double cs = 0.0;
    int cc = 0;
    for (long long c = lane; c < f; c += 32) {
      const double v = row[c];
      if (!isnan(v)) { cs += v; ++cc; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cs += __shfl_xor_sync(0xffffffffu, cs, o);
      cc += __shfl_xor_sync(0xffffffffu, cc, o);
    }
    if (cc > 0) { rsum += cs / cc; ++rcnt; }  // nanmean over the inner axis; NaN rows skipped next
  }
  if (lane == 0) {
    const double m = rcnt > 0 ? rsum / (double)rcnt : nan("");
    const int oi = flip ? S - 1 - i : i, oj = flip ? S - 1 - j : j;
    out[(long long)oi * S + oj] = logf((float)m);
    if (out_mean) out_mean[(long long)i * S + j] = m;  // the un-flipped float64 block mean (output['normmats'], :729-737)
  }
}

int background_level(const double* normmat, int64_t n, int64_t r0, int64_t f, int64_t S, int flip,
                     float* out, double* out_mean, cudaStream_t s) {
  if (n <= 0 || f <= 0 || S <= 0 || r0 < 0 || r0 + S * f > n) {
    set_error("background_forward: window [%lld, %lld) outside the (%lld x %lld) matrix", (long long)r0,
              (long long)(r0 + S * f), (long long)n, (long long)n);
    return ORCA_B200_EINVAL;
  }
  const long long warps = (long long)S * S;
  background_level_kernel<<<blocks_for(warps, 8), 256, 0, s>>>(normmat, n, r0, f, (int)S, flip, out, out_mean);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ---- background matrix assembly for multi-region inputs (orca_predict.py:936-965, _retrieve_multi) -------
// out[i][j] = chrom[i] == chrom[j] ? cis[(long long)(|coord[i] - coord[j]| / binsize)] : trans
// coord[k] / chrom[k]: genomic coordinate and chromosome id of the bin that lands at row/column k (strand flips
// already applied by the host).  One thread per pair of columns (16-byte stores); the 1-D cis curve stays in L2.
__global__ void background_assemble_kernel(const double* __restrict__ coord, const int* __restrict__ chrom,
                                           const double* __restrict__ cis, double trans, double binsize,
                                           double* __restrict__ out, long long n) {
  const long long half = (n + 1) / 2;
  const long long total = n * half;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / half, j = (t - i * half) * 2;
    const double ci = __ldg(coord + i);
    const int ki = __ldg(chrom + i);
    double v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const long long jj = j + e < n ? j + e : n - 1;
      if (__ldg(chrom + jj) != ki) { v[e] = trans; continue; }
      const long long d = (long long)__ddiv_rn(fabs(__dsub_rn(ci, __ldg(coord + jj))), binsize);  // .astype(int) truncates
      v[e] = __ldg(cis + d);
    }
    double* o = out + i * n + j;
    if (j + 1 < n && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
      *reinterpret_cast<double2*>(o) = make_double2(v[0], v[1]);
    } else {
      o[0] = v[0];
      if (j + 1 < n) o[1] = v[1];
    }
  }
}

int background_assemble(const double* coord, const int* chrom, const double* cis, double trans, double binsize,
                        double* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return ORCA_B200_OK;
  const long long total = n * ((n + 1) / 2);
  unsigned grid = blocks_for(total, 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  background_assemble_kernel<<<grid, 256, 0, s>>>(coord, chrom, cis, trans, binsize, out, n);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
