// First encoder block on tensor cores: lconv1 = Conv1d(4,64,k9) BN Conv1d(64,64,k9) BN  (orca_modules.py:811-816)
// has no nonlinearity, so away from the sequence ends it IS one Conv1d(4 -> 64, k = 17):
//   y2[l] = bc + sum_{t<17} Wc[t] x[l + t - 8],  Wc[t] = sum_{t1+t2=t} W2f[t2] W1f[t1],  bc = b2f + sum_t2 W2f[t2] b1f
// (W*f / b*f = BatchNorm-folded weights; composed in double precision on the host).  K = 17*4 = 68 -> 80, i.e.
// 5 UMMA K-steps per 128-position tile instead of 36 for the 64->64 layer it replaces, and the 4->64
// CUDA-core layer disappears.  The reference zero-pads the INTERMEDIATE activation, so the first/last 4
// positions of the whole sequence differ from the composed conv; lconv1_edge_kernel recomputes those 8
// positions exactly with the two separate layers (fp32).
//
// A operand: K index k = 4*tap + channel, so the 8 K-elements of chunk j for row l are the 8 consecutive
// floats x[l + 2j - 8][0..3], x[l + 2j - 7][0..3] of the caller's channel-last input: the im2col tile
// [10 chunks][128 rows][16 B] (bf16 hi and lo) is built by the CTA's 128 threads straight from global memory.
#include <cstring>
#include <vector>

#include "common.h"
#include "tc.h"
#include "tc_device.cuh"
#include "seq_in.cuh"

namespace orca {
namespace {
using namespace tcdev;

constexpr int kTaps = 17, kChunks = 10;  // 68 K-elements padded to 80

// F16 = false: operands split into bf16 hi/lo, three products, output as bf16 hi/lo planes (or one fp16 plane).
// F16 = true (single-pass stage 1): the one-hot input is exact in fp16, the composed weights are rounded to fp16,
// ONE product per K step, output one fp16 plane; half the shared memory and TMEM columns, so twice the CTAs per SM
// for this latency-bound (stage, build, multiply, drain) kernel.
template <bool F16>
__global__ void __launch_bounds__(128) lconv1_tc_kernel(const SeqIn in, long long Ltot, long long l_begin, long long n,
                                                        int npad, const uint8_t* __restrict__ wimg /*[10][128 | 64][16 B]*/,
                                                        const float* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ out_hi,
                                                        __nv_bfloat16* __restrict__ out_lo, unsigned int* sat) {
  extern __shared__ __align__(128) uint8_t dsm[];  // operand images (dynamic: above the 48 KB static limit)
  constexpr int kBRows = F16 ? 64 : 128;          // weight rows per K chunk: [B] or [Bh | Bl]
  constexpr uint32_t kTmemCols = F16 ? 64 : 128;
  uint8_t* sAh = dsm;
  uint8_t* sAl = dsm + kChunks * 128 * 16;  // unused when F16
  uint8_t* sBw = dsm + (F16 ? 1 : 2) * kChunks * 128 * 16;
  __shared__ float sBias[64];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;

  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < kChunks * kBRows; i += 128) reinterpret_cast<uint4*>(sBw)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
  if (tid < 64) sBias[tid] = bias[tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  // each CTA walks tiles (weights, TMEM allocation and barrier set up once); 3 CTAs per SM overlap each other's phases
  const int n_tiles = (int)((n + 127) / 128);
  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
  const long long t0 = (long long)tile * 128;

  // stage the 148 input positions of the tile (l0-8 .. l0+139; zero outside [0, Ltot)) once, coalesced ...
  __shared__ __align__(16) float sX[148][4];
  for (int i = tid; i < 148; i += 128) {  // fp32 view (16 B/bp) or packed bases (1 B/bp): seq_in.cuh
    const long long p = l_begin + t0 - 8 + i;
    *reinterpret_cast<float4*>(&sX[i][0]) = (p >= 0 && p < Ltot) ? seq_load(in, b, p) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  // ... then the im2col row of this thread: chunk j = positions (tid + 2j, tid + 2j + 1) of the staged tile
#pragma unroll
  for (int j = 0; j < kChunks; ++j) {
    const float4 p0 = *reinterpret_cast<const float4*>(&sX[tid + 2 * j][0]);
    const float4 p1 = *reinterpret_cast<const float4*>(&sX[tid + 2 * j + 1][0]);
    const float v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    if (F16) store_h8(v, sAh + (j * 128 + tid) * 16);
    else split_store8(v, reinterpret_cast<__nv_bfloat16*>(sAh + (j * 128 + tid) * 16),
                      reinterpret_cast<__nv_bfloat16*>(sAl + (j * 128 + tid) * 16));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(64), idesc_cat = umma_idesc_bf16(128), idesc16 = umma_idesc_f16(64);
      const uint32_t ah = umma_desc_lo(smem_u32(sAh), 2048), al = umma_desc_lo(smem_u32(sAl), 2048);
      const uint32_t bw = umma_desc_lo(smem_u32(sBw), kBRows * 16);
#pragma unroll
      for (int ks = 0; ks < kChunks / 2; ++ks) {
        const uint32_t o = ks * ((2 * 2048) >> 4), ob = ks * ((2 * kBRows * 16) >> 4);
        if (F16) {
          umma_bf16(tmem, umma_desc64(ah + o), umma_desc64(bw + ob), idesc16, ks > 0 ? 1u : 0u);  // fp16 A * fp16 B
        } else {
          umma_bf16(tmem, umma_desc64(ah + o), umma_desc64(bw + ob), idesc_cat, ks > 0 ? 1u : 0u);  // [Ah*Bh | Ah*Bl]
          umma_bf16(tmem, umma_desc64(al + o), umma_desc64(bw + ob), idesc, 1u);                    // += Al*Bh
        }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar), it & 1);
  tc_fence_after();
  const long long row = t0 + warp * 32 + lane;
#pragma unroll
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r0[32], r1[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r0);
    if (!F16) tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 64 + c0, r1);
    if (row < n) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]) + (F16 ? 0.f : __uint_as_float(r1[j])) + sBias[c0 + j];
      if (!out_lo) guard_h<32>(v, sat);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const size_t off = (((size_t)b * 8 + (c0 >> 3) + ch) * npad + row + 4) * 8;
        if (out_lo) split_store8(v + 8 * ch, out_hi + off, out_lo + off);  // bf16 hi/lo planes
        else store_h8(v + 8 * ch, out_hi + off);                           // one fp16 plane (single-pass stages)
      }
    }
  }
  tc_fence_before();
  __syncthreads();  // every warp has drained its TMEM lanes and shared-memory reads before the next tile overwrites them
  tc_fence_after();
  }
  if (blockIdx.x == 0) {  // pad rows of this sample's planes
    const int tail0 = (int)n + 4, ntail = npad - tail0, per_plane = 4 + ntail;
    for (int i = tid; i < 8 * per_plane; i += 128) {
      const int p = i / per_plane, j = i - p * per_plane;
      const size_t r = ((size_t)b * 8 + p) * npad + (j < 4 ? j : tail0 + (j - 4));
      reinterpret_cast<uint4*>(out_hi)[r] = make_uint4(0, 0, 0, 0);
      if (out_lo) reinterpret_cast<uint4*>(out_lo)[r] = make_uint4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
}

// Exact two-layer evaluation of the 4 positions next to one end of the sequence (see file header).
// grid (2 ends, nb), 64 threads = output channels.  w1 [9][4][64], w2 [9][64][64] (folded, fp32).
__global__ void __launch_bounds__(64) lconv1_edge_kernel(const SeqIn in, long long Ltot, long long l_begin, long long n,
                                                         int npad, const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  __shared__ float y1[8][64];
  const int co = threadIdx.x, end = blockIdx.x, b = blockIdx.y;
  const long long p0 = end == 0 ? 0 : Ltot - 8;  // first of the 8 intermediate positions needed
  const long long l0 = end == 0 ? 0 : Ltot - 4;  // first of the 4 output positions
  if (l0 < l_begin || l0 + 4 > l_begin + n) return;  // this window does not own that end
  for (int i = 0; i < 8; ++i) {
    float acc = b1[co];
    for (int t = 0; t < 9; ++t) {
      const long long p = p0 + i + t - 4;
      if (p < 0 || p >= Ltot) continue;
      const float4 v4 = seq_load(in, b, p);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      for (int c = 0; c < 4; ++c) acc = fmaf(v[c], w1[(t * 4 + c) * 64 + co], acc);
    }
    y1[i][co] = acc;
  }
  __syncthreads();
  for (int i = 0; i < 4; ++i) {
    const long long l = l0 + i;
    float acc = b2[co];
    for (int t = 0; t < 9; ++t) {
      const long long p = l + t - 4;
      if (p < 0 || p >= Ltot) continue;  // zero padding of the INTERMEDIATE activation
      const int pi = (int)(p - p0);
      for (int m = 0; m < 64; ++m) acc = fmaf(y1[pi][m], w2[(t * 64 + m) * 64 + co], acc);
    }
    const size_t off = (((size_t)b * 8 + (co >> 3)) * npad + (size_t)(l - l_begin) + 4) * 8 + (co & 7);
    if (out_lo) {
      const __nv_bfloat16 h = __float2bfloat16_rn(acc);
      out_hi[off] = h;
      out_lo[off] = __float2bfloat16_rn(acc - __bfloat162float(h));
    } else {
      reinterpret_cast<__half*>(out_hi)[off] = __float2half_rn(acc);
    }
  }
}

uint16_t bits_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
float bits_f32(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
}  // namespace

// w1 [9][4][64], b1 [64], w2 [9][64][64], b2 [64]: BatchNorm-folded fp32 weights of lconv1[0], lconv1[1]
int tc_pack_lconv1(ConvLayer& L0, const float* w1, const float* b1, const float* w2, const float* b2,
                   std::vector<void*>& allocs) {
  std::vector<double> wc((size_t)kTaps * 4 * 64, 0.0), bc(64);
  for (int co = 0; co < 64; ++co) {
    double acc = b2[co];
    for (int t2 = 0; t2 < 9; ++t2)
      for (int m = 0; m < 64; ++m) acc += (double)w2[((size_t)t2 * 64 + m) * 64 + co] * (double)b1[m];
    bc[co] = acc;
  }
  for (int t2 = 0; t2 < 9; ++t2)
    for (int t1 = 0; t1 < 9; ++t1)
      for (int c = 0; c < 4; ++c)
        for (int m = 0; m < 64; ++m) {
          const double a = w1[((size_t)t1 * 4 + c) * 64 + m];
          const float* w2r = w2 + ((size_t)t2 * 64 + m) * 64;
          double* dst = &wc[((size_t)(t1 + t2) * 4 + c) * 64];
          for (int co = 0; co < 64; ++co) dst[co] += a * (double)w2r[co];
        }
  std::vector<uint16_t> img((size_t)kChunks * 128 * 8, 0);  // [chunk][Bh rows 0..63 | Bl rows 64..127][8]
  for (int j = 0; j < kChunks; ++j)
    for (int co = 0; co < 64; ++co)
      for (int e = 0; e < 8; ++e) {
        const int k = j * 8 + e, t = k / 4, c = k % 4;
        const float v = t < kTaps ? (float)wc[((size_t)t * 4 + c) * 64 + co] : 0.f;
        const uint16_t h = bits_rn(v);
        img[((size_t)j * 128 + co) * 8 + e] = h;
        img[((size_t)j * 128 + 64 + co) * 8 + e] = bits_rn(v - bits_f32(h));
      }
  std::vector<float> bf(64);
  for (int i = 0; i < 64; ++i) bf[i] = (float)bc[i];
  void* d = nullptr;
  ORCA_CUDA_OK(cudaMalloc(&d, img.size() * 2));
  allocs.push_back(d);
  ORCA_CUDA_OK(cudaMemcpy(d, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  void* db = nullptr;
  ORCA_CUDA_OK(cudaMalloc(&db, 64 * sizeof(float)));
  allocs.push_back(db);
  ORCA_CUDA_OK(cudaMemcpy(db, bf.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
  L0.tc_w = d;
  L0.tc_w_bytes = img.size() * 2;
  L0.tc_bias = static_cast<float*>(db);
  // fp16 image of the composed weights for the single-pass variant: [chunk][64 rows][8]
  std::vector<uint16_t> img16((size_t)kChunks * 64 * 8, 0);
  for (int j = 0; j < kChunks; ++j)
    for (int co = 0; co < 64; ++co)
      for (int e = 0; e < 8; ++e) {
        const int k = j * 8 + e, t = k / 4, c = k % 4;
        const __half hv = __float2half_rn(t < kTaps ? (float)wc[((size_t)t * 4 + c) * 64 + co] : 0.f);
        memcpy(&img16[((size_t)j * 64 + co) * 8 + e], &hv, 2);
      }
  void* d16 = nullptr;
  ORCA_CUDA_OK(cudaMalloc(&d16, img16.size() * 2));
  allocs.push_back(d16);
  ORCA_CUDA_OK(cudaMemcpy(d16, img16.data(), img16.size() * 2, cudaMemcpyHostToDevice));
  L0.tc_w16 = d16;
  return ORCA_B200_OK;
}

int tc_lconv1(const ConvLayer& L0, const ConvLayer& L1, const SeqIn& in, int nb, int64_t Ltot, int64_t l_begin, int64_t n,
              TcAct* out, cudaStream_t s) {
  if (!L0.tc_w || !L0.tc_bias || L0.c_in != 4 || L0.c_out != 64 || L1.c_in != 64 || L1.c_out != 64 || out->C != 64 ||
      out->n != n || out->nb != nb) {
    set_error("tc_lconv1: bad layers / geometry");
    return ORCA_B200_EINVAL;
  }
  const long long n_tiles = (n + 127) / 128;
  const bool f16 = out->fmt == 1 && L0.tc_w16;
  if (out->fmt == 1 && !f16) { set_error("tc_lconv1: no fp16 weight image"); return ORCA_B200_EUNSUPPORTED; }
  const int per_sm = f16 ? 6 : 3;
  dim3 grid((unsigned)(n_tiles < 148 * per_sm ? n_tiles : 148 * per_sm), (unsigned)nb);
  constexpr int kSmem = 3 * kChunks * 128 * 16, kSmem16 = kChunks * 128 * 16 + kChunks * 64 * 16;
  static bool configured_dev[32] = {};  // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 31];
  if (!configured) {
    ORCA_CUDA_OK(cudaFuncSetAttribute(lconv1_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    ORCA_CUDA_OK(cudaFuncSetAttribute(lconv1_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem16));
    configured = true;
  }
  if (f16)
    lconv1_tc_kernel<true><<<grid, 128, kSmem16, s>>>(in, Ltot, l_begin, n, (int)out->npad, static_cast<const uint8_t*>(L0.tc_w16),
                                                    L0.tc_bias, static_cast<__nv_bfloat16*>(out->hi), nullptr, out->sat);
  else
    lconv1_tc_kernel<false><<<grid, 128, kSmem, s>>>(in, Ltot, l_begin, n, (int)out->npad, static_cast<const uint8_t*>(L0.tc_w),
                                                     L0.tc_bias, static_cast<__nv_bfloat16*>(out->hi),
                                                     out->fmt ? nullptr : static_cast<__nv_bfloat16*>(out->lo), nullptr);
  ORCA_LAUNCH_OK();
  if (l_begin == 0 || l_begin + n == Ltot) {
    lconv1_edge_kernel<<<dim3(2, (unsigned)nb), 64, 0, s>>>(in, Ltot, l_begin, n, (int)out->npad, L0.w, L0.b, L1.w, L1.b,
                                                            static_cast<__nv_bfloat16*>(out->hi),
                                                            out->fmt ? nullptr : static_cast<__nv_bfloat16*>(out->lo));
    ORCA_LAUNCH_OK();
  }
  return ORCA_B200_OK;
}

}  // namespace orca
