// tcgen05 implicit-GEMM dilated 3x3 Conv2d for sm_100a -- the Decoder / Decoder_1m hot loop
// (orca_modules.py:22-488, :499-800), same arithmetic and kernel skeleton as conv_tc.cu.
//
// Layout: a (nb, C, S, S) map is hi/lo[nb][C/8][plane_rows][8] bf16 with pixel (y, x) at row
// y*Wp + 64 + x, Wp = S + 128: every image row carries 64 zero pixels on each side, which is the
// horizontal zero padding for every dilation d <= 64.  Vertical padding needs no memory: a tap row that
// falls outside the image contributes nothing, so its MMAs are skipped.
//
// Tile = 128 consecutive pixels of one image row (the last tile of a row is shifted left to end at S).
// For each dy in {-d, 0, +d} the producer bulk-copies ONE run of 128 + 2d pixels per 8-channel chunk;
// the three dx taps are descriptor start offsets (0, d, 2d rows) into that run.  Weights are staged per
// (dy, K-block) as three [Bh;Bl] tap images and stay resident in shared memory when they fit.
#include <cstring>
#include <vector>

#include "common.h"
#include "tc.h"
#include "tc_device.cuh"

namespace orca {

namespace {
using namespace tcdev;

constexpr int kPX = 64;  // zero pixels on each side of an image row
constexpr int kEpiWarps = 8;                   // two per TMEM lane quarter (the epilogue is latency bound per warp)
constexpr int kThreads2d = 64 + 32 * kEpiWarps;

struct Tc2dKArgs {
  const __nv_bfloat16* in_hi; const __nv_bfloat16* in_lo;
  const uint8_t* w;
  const float* bias;
  const __nv_bfloat16* res_hi; const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  long long plane_rows;
  int nb, S, Wp, d, relu, c_in;
  int tiles_per_row, total_tiles;
  int NA, NW, resident;            // ring sizes
  int a_slot_bytes, w_stage_bytes; // per-slot strides in shared memory
};

// All MMAs of one (dy, K-block) stage: 3 dx taps x KSTEPS K steps x {Ah x [Bh;Bl] (N = 2*C_OUT), Al x Bh}.
// FIRST0: the very first MMA of the tile overwrites the accumulator.
template <int C_OUT, int KSTEPS, bool FIRST>
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint32_t aLo, uint32_t bLo, uint32_t aStep, uint32_t aLoStep,
                                            uint32_t tapStep, uint32_t d, bool first_kb) {
  constexpr uint32_t idesc = umma_idesc_bf16(C_OUT), idesc_cat = umma_idesc_bf16(2 * C_OUT);
  constexpr uint32_t bStep = (2u * 2 * C_OUT * 16) >> 4;
#pragma unroll
  for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const uint32_t ao = aLo + ks * aStep + dxi * d;  // dx tap = dxi*d rows of 16 B into the run
      const uint32_t bo = bLo + dxi * tapStep + ks * bStep;
      const uint32_t accum = (FIRST && dxi == 0 && ks == 0) ? (first_kb ? 0u : 1u) : 1u;
      umma_bf16(d_tmem, umma_desc64(ao), umma_desc64(bo), idesc_cat, accum);
      umma_bf16(d_tmem, umma_desc64(ao + aLoStep), umma_desc64(bo), idesc, 1u);
    }
  }
}

template <int C_OUT, int KSTEPS>
__global__ void __launch_bounds__(kThreads2d, 1) conv2d_tc_kernel(const Tc2dKArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int NA = a.NA, NW = a.NW;
  uint8_t* sA = smem;
  uint8_t* sW = sA + (size_t)NA * a.a_slot_bytes;
  float* sBias = reinterpret_cast<float*>(sW + (size_t)NW * a.w_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + C_OUT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NA + 2 * NW + 4);
  const uint32_t bA_full = smem_u32(bars), bA_empty = bA_full + 8 * NA;
  const uint32_t bW_full = bA_empty + 8 * NA, bW_empty = bW_full + 8 * NW;
  const uint32_t bAcc_full = bW_empty + 8 * NW, bAcc_empty = bAcc_full + 16;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = (a.c_in + 63) / 64;
  const int R = 128 + 2 * a.d;                 // rows of one A run
  const int kc_all = (a.c_in < 64 ? a.c_in : 64) / 8;  // chunks per K-block (32 -> 4, 64/128 -> 8)
  const uint32_t aLoOff = (uint32_t)kc_all * R * 16;
  const uint32_t tapBytes = 2u * kc_all * C_OUT * 16;  // one tap image: [k-chunk][Bh rows | Bl rows][16 B]

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(bA_full + 8 * i, 1); mbar_init(bA_empty + 8 * i, 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(bW_full + 8 * i, 1); mbar_init(bW_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bAcc_full + 8 * i, 1); mbar_init(bAcc_empty + 8 * i, kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < C_OUT; i += kThreads2d) sBias[i] = a.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tiles_per_img = a.S * a.tiles_per_row;

  if (warp == 0) {
    // ================= producer =================
    // warp-uniform loop; one elected lane issues the bulk copies
    uint32_t a_it = 0, w_it = 0, loaded = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int y = rem / a.tiles_per_row, tx = rem - y * a.tiles_per_row;
      int x0 = tx * 128;
      if (x0 + 128 > a.S) x0 = a.S > 128 ? a.S - 128 : 0;
#pragma unroll 1
      for (int si = 0; si < 3; ++si) {
        const int dyi = si == 0 ? 1 : (si == 1 ? 0 : 2);  // centre row first (always inside the image)
        const int yy = y + (dyi - 1) * a.d;
        if (yy < 0 || yy >= a.S) continue;
        const long long row0 = (long long)yy * a.Wp + x0 + kPX - a.d;
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t slot = a_it % NA, ph = (a_it / NA) & 1;
          mbar_wait(bA_empty + 8 * slot, ph ^ 1);
          const int sid = dyi * nkb + kb;
          const bool need_w = a.resident ? !((loaded >> sid) & 1u) : true;
          const uint32_t ws = a.resident ? (uint32_t)sid : w_it % NW;
          if (!a.resident) mbar_wait(bW_empty + 8 * ws, ((w_it / NW) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bA_full + 8 * slot, 2u * kc_all * R * 16);
            if (need_w) {
              mbar_expect_tx(bW_full + 8 * ws, 3 * tapBytes);
              bulk_g2s(smem_u32(sW) + ws * a.w_stage_bytes, a.w + (size_t)sid * 3 * tapBytes, 3 * tapBytes, bW_full + 8 * ws);
            }
          }
          __syncwarp();
          if (lane < 2 * kc_all) {  // one bulk copy per lane (see conv2d_prog.cu: a single issuing lane paced the layer)
            const int c = lane >> 1, part = lane & 1;
            const uint32_t dst = smem_u32(sA) + slot * a.a_slot_bytes + (part ? aLoOff : 0u) + c * R * 16;
            const long long plane = (long long)b * (a.c_in / 8) + kb * 8 + c;
            const long long off = (plane * a.plane_rows + row0) * 8;
            bulk_g2s(dst, (part ? a.in_lo : a.in_hi) + off, R * 16, bA_full + 8 * slot);
          }
          __syncwarp();
          ++a_it;
          if (a.resident) loaded |= 1u << sid; else ++w_it;
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues =================
    uint32_t a_it = 0, w_it = 0, acc_it = 0, waited = 0;
    const uint32_t aStep = (uint32_t)(2 * R * 16) >> 4, aLoStep = aLoOff >> 4, tapStep = tapBytes >> 4;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int rem = tile % tiles_per_img;
      const int y = rem / a.tiles_per_row;
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      mbar_wait(bAcc_empty + 8 * as, aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + as * 128;
      const int last_dyi = (y + a.d < a.S) ? 2 : ((y - a.d >= 0) ? 0 : 1);
#pragma unroll 1
      for (int si = 0; si < 3; ++si) {
        const int dyi = si == 0 ? 1 : (si == 1 ? 0 : 2);  // same order as the producer
        const int yy = y + (dyi - 1) * a.d;
        if (yy < 0 || yy >= a.S) continue;
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t slot = a_it % NA;
          mbar_wait(bA_full + 8 * slot, (a_it / NA) & 1);
          const int sid = dyi * nkb + kb;
          uint32_t ws;
          if (a.resident) {
            ws = sid;
            if (!((waited >> sid) & 1u)) { waited |= 1u << sid; mbar_wait(bW_full + 8 * ws, 0); }
          } else {
            ws = w_it % NW;
            mbar_wait(bW_full + 8 * ws, (w_it / NW) & 1);
          }
          tc_fence_after();
          // warp-uniform descriptor base words
          const uint32_t aLo = __shfl_sync(0xffffffffu, umma_desc_lo(smem_u32(sA) + slot * a.a_slot_bytes, R * 16), 0);
          const uint32_t bLo = __shfl_sync(0xffffffffu, umma_desc_lo(smem_u32(sW) + ws * a.w_stage_bytes, 2 * C_OUT * 16), 0);
          if (elect_one()) {
            if (si == 0) issue_stage<C_OUT, KSTEPS, true>(d_tmem, aLo, bLo, aStep, aLoStep, tapStep, (uint32_t)a.d, kb == 0);
            else issue_stage<C_OUT, KSTEPS, false>(d_tmem, aLo, bLo, aStep, aLoStep, tapStep, (uint32_t)a.d, false);
            if (!a.resident) umma_commit(bW_empty + 8 * ws);
            umma_commit(bA_empty + 8 * slot);
            if (dyi == last_dyi && kb == nkb - 1) umma_commit(bAcc_full + 8 * as);
          }
          __syncwarp();
          if (!a.resident) ++w_it;
          ++a_it;
        }
      }
      ++acc_it;
    }
  } else {
    // ================= epilogue =================
    // warp (q, h): TMEM lane quarter q, 16-column units h, h+2, ... of the C_OUT output channels
    const int q = warp & 3, h = (warp - 2) >> 2;
    constexpr int UNITS = C_OUT / 16, MYU = UNITS / 2;  // units per warp (1 or 2)
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img, rem = tile - b * tiles_per_img;
      const int y = rem / a.tiles_per_row, tx = rem - y * a.tiles_per_row;
      int x0 = tx * 128;
      if (x0 + 128 > a.S) x0 = a.S > 128 ? a.S - 128 : 0;
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      const int x = x0 + q * 32 + lane;
      const bool valid = x < a.S;
      const long long r = (long long)y * a.Wp + kPX + x;
      // residual fetched BEFORE waiting for the accumulator: its L2 latency overlaps the tile's MMAs
      float res[MYU][16];
#pragma unroll
      for (int u = 0; u < MYU; ++u) {
#pragma unroll
        for (int j = 0; j < 16; ++j) res[u][j] = 0.f;
        if (a.res_hi && valid) {
          const int c0 = 16 * (h + 2 * u);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const long long off = (((long long)b * (C_OUT / 8) + (c0 >> 3) + ch) * a.plane_rows + r) * 8;
            add_hilo8(res[u] + 8 * ch, a.res_hi + off, a.res_lo + off);
          }
        }
      }
      mbar_wait(bAcc_full + 8 * as, aph);
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < MYU; ++u) {
        const int c0 = 16 * (h + 2 * u);
        uint32_t raw[16], raw2[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + as * 128 + c0, raw);
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + as * 128 + C_OUT + c0, raw2);
        if (valid) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float t = __uint_as_float(raw[j]) + __uint_as_float(raw2[j]) + sBias[c0 + j];
            v[j] = (a.relu ? fmaxf(t, 0.f) : t) + res[u][j];
          }
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const long long off = (((long long)b * (C_OUT / 8) + (c0 >> 3) + ch) * a.plane_rows + r) * 8;
            split_store8(v + 8 * ch, a.out_hi + off, a.out_lo + off);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bAcc_empty + 8 * as);  // one arrival per warp: 256 same-word atomics per tile serialise
      ++acc_it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

// ---- glue on map planes -----------------------------------------------------------------------------
// outer-sum lift straight into planes: mat[b][c][i][j] = x[b][i][c] + x[b][j][c]  (orca_modules.py:462, :783)
__global__ void outer_sum_planes_kernel(const float* __restrict__ xcl /*[B][S][C]*/, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo, int S, int C, int Wp, long long plane_rows,
                                        long long total) {
  const int C8 = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % S);
    long long t = i / S;
    const int y = (int)(t % S);
    t /= S;
    const int ch = (int)(t % C8);
    const long long b = t / C8;
    const float4* pi = reinterpret_cast<const float4*>(xcl + (b * S + y) * C + ch * 8);
    const float4* pj = reinterpret_cast<const float4*>(xcl + (b * S + x) * C + ch * 8);
    const float4 a0 = __ldg(pi), a1 = __ldg(pi + 1), b0 = __ldg(pj), b1 = __ldg(pj + 1);
    const float v[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w, a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
    const long long off = ((b * C8 + ch) * plane_rows + (long long)y * Wp + kPX + x) * 8;
    split_store8(v, hi + off, lo + off);
  }
}

// single-channel 3x3 conv (distance encoding / upsampled coarse map) -> 64-channel planes; see glue.cu
__device__ __forceinline__ float extra_src2(const float* sb, long long sH, long long sW, int S, int mode, int yy, int xx) {
  if (yy < 0 || yy >= S || xx < 0 || xx >= S) return 0.f;
  if (mode == 0) return __ldg(sb + yy * sH + xx * sW);
  if (mode == 1) return __ldg(sb + (yy >> 1) * sH + (xx >> 1) * sW);
  const int n = S >> 1;
  const float fy = fmaxf((yy + 0.5f) * 0.5f - 0.5f, 0.f), fx = fmaxf((xx + 0.5f) * 0.5f - 0.5f, 0.f);
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = min(y0 + 1, n - 1), x1 = min(x0 + 1, n - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float v00 = __ldg(sb + y0 * sH + x0 * sW), v01 = __ldg(sb + y0 * sH + x1 * sW);
  const float v10 = __ldg(sb + y1 * sH + x0 * sW), v11 = __ldg(sb + y1 * sH + x1 * sW);
  return (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
}

// one thread per pixel: the 9 (upsampled) source values of a channel are evaluated once and feed all 64 outputs
__global__ void __launch_bounds__(128) extra_conv_planes_kernel(const float* __restrict__ src, long long sB, long long sC, long long sH,
                                                                long long sW, int n_extra, const float* __restrict__ w /*[n_extra][9][64]*/,
                                                                __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int S, int Wp,
                                                                long long plane_rows, int mode, long long total) {
  extern __shared__ float sw[];  // the weights, [n_extra][9][64]
  for (int i = threadIdx.x; i < n_extra * 9 * 64; i += blockDim.x) sw[i] = __ldg(w + i);
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % S);
    const long long t = i / S;
    const int y = (int)(t % S);
    const long long b = t / S;
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    for (int e = 0; e < n_extra; ++e) {
      const float* sb = src + b * sB + e * sC;
      float v[9];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) v[tp] = extra_src2(sb, sH, sW, S, mode, y + tp / 3 - 1, x + tp % 3 - 1);
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const float4* wr = reinterpret_cast<const float4*>(sw + (e * 9 + tp) * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 w4 = wr[j];  // same address across the warp: shared-memory broadcast
          acc[4 * j] = fmaf(v[tp], w4.x, acc[4 * j]); acc[4 * j + 1] = fmaf(v[tp], w4.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v[tp], w4.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(v[tp], w4.w, acc[4 * j + 3]);
        }
      }
    }
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const long long off = ((b * 8 + ch) * plane_rows + (long long)y * Wp + kPX + x) * 8;
      split_store8(acc + 8 * ch, hi + off, lo + off);
    }
  }
}

// output head on planes: 1x1 64->H (+BN) ReLU, 1x1 H->O  (orca_modules.py:423-428: H = 5, O = 1;
// orca_leukemia.py:923-926: O = num_2d, H = max(num_2d, 5)); one thread per pixel, tmp is [nb][O][S][S]
template <int H>
__global__ void final_head_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                         const float* __restrict__ w0 /*[64][H]*/, const float* __restrict__ b0,
                                         const float* __restrict__ w1 /*[H][O]*/, const float* __restrict__ b1,
                                         float* __restrict__ tmp, int S, int Wp, long long plane_rows, long long total, int O) {
  __shared__ float sw[64 * H], sb0[H], sw1[H * 8], sb1[8];
  for (int i = threadIdx.x; i < 64 * H; i += blockDim.x) sw[i] = w0[i];
  if (threadIdx.x < H) sb0[threadIdx.x] = b0[threadIdx.x];
  if (threadIdx.x < H * O) sw1[threadIdx.x] = w1[threadIdx.x];
  if (threadIdx.x < O) sb1[threadIdx.x] = b1[threadIdx.x];
  __syncthreads();
  const long long img = (long long)S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % S);
    const long long t = i / S;
    const int y = (int)(t % S);
    const long long b = t / S;
    float h[H];
#pragma unroll
    for (int k = 0; k < H; ++k) h[k] = 0.f;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long off = ((b * 8 + ch) * plane_rows + (long long)y * Wp + kPX + x) * 8;
      add_hilo8(v, hi + off, lo + off);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int k = 0; k < H; ++k) h[k] = fmaf(v[j], sw[(ch * 8 + j) * H + k], h[k]);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) h[k] = fmaxf(h[k] + sb0[k], 0.f);
    for (int o = 0; o < O; ++o) {
      float r = sb1[o];
#pragma unroll
      for (int k = 0; k < H; ++k) r = fmaf(h[k], sw1[k * O + o], r);
      tmp[(b * O + o) * img + (long long)y * S + x] = r;
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
static inline uint16_t bf16_bits_rn2(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_to_f32_2(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

bool tc_layer2d_eligible(const ConvLayer& L) {
  return L.kh == 3 && L.kw == 3 && ((L.c_in == 32 && L.c_out == 64) || ((L.c_in == 64 || L.c_in == 128) && (L.c_out == 32 || L.c_out == 64))) &&
         L.dil >= 1 && L.dil <= 64;
}

// Stage images in consumption order: for dy, for K-block: three dx taps, each [k-chunk][Bh rows | Bl rows][8] bf16.
int tc_pack_layer2d(ConvLayer& L, const float* w /*[tap][c_in][c_out]*/, std::vector<void*>& allocs) {
  if (!tc_layer2d_eligible(L)) return ORCA_B200_OK;
  const int nkb = (L.c_in + 63) / 64, ks = L.c_in < 64 ? L.c_in : 64;
  std::vector<uint16_t> img;
  img.reserve((size_t)9 * L.c_in * L.c_out * 2);
  for (int dy = 0; dy < 3; ++dy)
    for (int kb = 0; kb < nkb; ++kb)
      for (int dx = 0; dx < 3; ++dx)
        for (int c = 0; c < ks / 8; ++c)
          for (int part = 0; part < 2; ++part)
            for (int n = 0; n < L.c_out; ++n)
              for (int j = 0; j < 8; ++j) {
                const int ci = kb * 64 + c * 8 + j, tap = dy * 3 + dx;
                const float v = w[((size_t)tap * L.c_in + ci) * L.c_out + n];
                const uint16_t h = bf16_bits_rn2(v);
                img.push_back(part == 0 ? h : bf16_bits_rn2(v - bf16_to_f32_2(h)));
              }
  void* d = nullptr;
  ORCA_CUDA_OK(cudaMalloc(&d, img.size() * 2));
  allocs.push_back(d);
  ORCA_CUDA_OK(cudaMemcpy(d, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  L.tc_w = d;
  L.tc_w_bytes = img.size() * 2;
  return ORCA_B200_OK;
}

static int sm_count2() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int tc_conv2d(const ConvLayer& L, const TcMap& in, const TcMap* res, TcMap* out, int relu, cudaStream_t s) {
  if (!L.tc_w || !tc_layer2d_eligible(L)) { set_error("tc_conv2d: layer %d->%d has no tensor-core weights", L.c_in, L.c_out); return ORCA_B200_EUNSUPPORTED; }
  if (in.C != L.c_in || out->C != L.c_out || out->S != in.S || out->nb != in.nb || (res && (res->C != L.c_out || res->S != in.S))) {
    set_error("tc_conv2d: geometry mismatch");
    return ORCA_B200_EINVAL;
  }
  Tc2dKArgs a;
  a.in_hi = static_cast<const __nv_bfloat16*>(in.hi); a.in_lo = static_cast<const __nv_bfloat16*>(in.lo);
  a.w = static_cast<const uint8_t*>(L.tc_w); a.bias = L.b;
  a.res_hi = res ? static_cast<const __nv_bfloat16*>(res->hi) : nullptr;
  a.res_lo = res ? static_cast<const __nv_bfloat16*>(res->lo) : nullptr;
  a.out_hi = static_cast<__nv_bfloat16*>(out->hi); a.out_lo = static_cast<__nv_bfloat16*>(out->lo);
  a.plane_rows = in.plane_rows; a.nb = in.nb; a.S = in.S; a.Wp = in.Wp; a.d = L.dil; a.relu = relu; a.c_in = L.c_in;
  a.tiles_per_row = (in.S + 127) / 128; a.total_tiles = in.nb * in.S * a.tiles_per_row;
  const int nkb = (L.c_in + 63) / 64, kc = (L.c_in < 64 ? L.c_in : 64) / 8, R = 128 + 2 * L.dil;
  a.a_slot_bytes = 2 * kc * R * 16;
  a.w_stage_bytes = 3 * 2 * kc * L.c_out * 16;
  const int n_stages = 3 * nkb;
  const int limit = 227 * 1024 - 2048;
  if (n_stages * a.w_stage_bytes + 2 * a.a_slot_bytes <= limit) {
    a.resident = 1; a.NW = n_stages;
  } else {
    a.resident = 0; a.NW = 2;
    if (2 * a.w_stage_bytes + a.a_slot_bytes > limit) a.NW = 1;
  }
  int na = (limit - a.NW * a.w_stage_bytes) / a.a_slot_bytes;
  if (na < 1) { set_error("tc_conv2d: shared memory budget exceeded"); return ORCA_B200_EUNSUPPORTED; }
  a.NA = na > 4 ? 4 : na;
  const int smem = a.NA * a.a_slot_bytes + a.NW * a.w_stage_bytes + L.c_out * 4 + (2 * a.NA + 2 * a.NW + 4) * 8 + 16 + 128;
  const int sms = sm_count2();
  const int grid = a.total_tiles < sms ? a.total_tiles : sms;
  if (a.total_tiles <= 0) return ORCA_B200_OK;
  static bool configured_dev[32] = {};  // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 31];
  if (!configured) {
    ORCA_CUDA_OK(cudaFuncSetAttribute(conv2d_tc_kernel<32, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ORCA_CUDA_OK(cudaFuncSetAttribute(conv2d_tc_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    ORCA_CUDA_OK(cudaFuncSetAttribute(conv2d_tc_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (L.c_out == 32 && kc == 8) conv2d_tc_kernel<32, 4><<<grid, kThreads2d, smem, s>>>(a);
  else if (L.c_out == 64 && kc == 8) conv2d_tc_kernel<64, 4><<<grid, kThreads2d, smem, s>>>(a);
  else if (L.c_out == 64 && kc == 4) conv2d_tc_kernel<64, 2><<<grid, kThreads2d, smem, s>>>(a);
  else { set_error("tc_conv2d: no kernel for %d->%d", L.c_in, L.c_out); return ORCA_B200_EUNSUPPORTED; }
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

static unsigned grid_for(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return (unsigned)(g < 1 ? 1 : g);
}

int tc_outer_sum(const float* xcl, TcMap* out, cudaStream_t s) {
  const long long total = (long long)out->nb * (out->C / 8) * out->S * out->S;
  outer_sum_planes_kernel<<<grid_for(total), 256, 0, s>>>(xcl, static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo),
                                                          out->S, out->C, out->Wp, out->plane_rows, total);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

int tc_extra_conv(const float* src, int64_t sB, int64_t sC, int64_t sH, int64_t sW, int n_extra, const float* w_extra,
                  TcMap* out, int mode, cudaStream_t s) {
  if (out->C != 64 || n_extra < 1 || n_extra > 8) { set_error("tc_extra_conv: C != 64 or bad extra-channel count"); return ORCA_B200_EINVAL; }
  const long long total = (long long)out->nb * out->S * out->S;
  long long grid = (total + 127) / 128;
  if (grid > 148 * 16) grid = 148 * 16;
  extra_conv_planes_kernel<<<(unsigned)grid, 128, (size_t)n_extra * 9 * 64 * sizeof(float), s>>>(
      src, sB, sC, sH, sW, n_extra, w_extra, static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo), out->S, out->Wp,
      out->plane_rows, mode, total);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

int tc_final_head_tmp(const TcMap& in, const ConvLayer& f0, const ConvLayer& f1, float* tmp, cudaStream_t s) {
  if (in.C != 64 || !final_head_ok(f0, f1)) { set_error("tc_final_head: bad layers (64->%d->%d)", f0.c_out, f1.c_out); return ORCA_B200_EINVAL; }
  const long long total = (long long)in.nb * in.S * in.S;
  const __nv_bfloat16* hi = static_cast<const __nv_bfloat16*>(in.hi);
  const __nv_bfloat16* lo = static_cast<const __nv_bfloat16*>(in.lo);
  const unsigned g = grid_for(total);
  const int O = f1.c_out;
  switch (f0.c_out) {
    case 5: final_head_planes_kernel<5><<<g, 256, 0, s>>>(hi, lo, f0.w, f0.b, f1.w, f1.b, tmp, in.S, in.Wp, in.plane_rows, total, O); break;
    case 6: final_head_planes_kernel<6><<<g, 256, 0, s>>>(hi, lo, f0.w, f0.b, f1.w, f1.b, tmp, in.S, in.Wp, in.plane_rows, total, O); break;
    case 7: final_head_planes_kernel<7><<<g, 256, 0, s>>>(hi, lo, f0.w, f0.b, f1.w, f1.b, tmp, in.S, in.Wp, in.plane_rows, total, O); break;
    default: final_head_planes_kernel<8><<<g, 256, 0, s>>>(hi, lo, f0.w, f0.b, f1.w, f1.b, tmp, in.S, in.Wp, in.plane_rows, total, O); break;
  }
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
