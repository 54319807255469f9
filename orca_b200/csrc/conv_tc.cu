// tcgen05 (5th-gen tensor core) implicit-GEMM Conv1d(k=9) for sm_100a  --  the Encoder hot loop
// (orca_modules.py:811-950) and the 128-channel U-nets.
//
// Arithmetic: fp32 parity needs more than one bf16 pass, so every fp32 operand
// is split x = hi + lo (two bf16) and D += Ah*Bh + Al*Bh + Ah*Bl accumulates in fp32 in TMEM
// (the dropped Al*Bl term is ~2^-16 relative).  Activations therefore live in HBM as two bf16 planes.
//
// Data layout ("chunk planes"): a (nb, C, n) activation is hi/lo[nb][C/8][npad][8] bf16, row r = l + 4
// (4 zero rows in front, zero rows behind up to npad = roundup(n,128) + 8).  One 16-byte element = 8
// channels of one position, so
//   * a tile's A operand is C/8 contiguous runs of 136 rows -> plain cp.async.bulk (UBLKCP), the zero
//     rows ARE the conv padding;
//   * in shared memory the runs form the K-major SWIZZLE_NONE UMMA layout [k-chunk][row][16 B]
//     (LBO = 136*16, SBO = 128): tap t of the k=9 stencil is the same buffer with the descriptor start
//     address advanced by t*16 bytes -- no im2col, no re-load (probe: tests/cuda/umma_probe.cu);
//   * the epilogue thread that owns TMEM lane (= position) l stores 16 contiguous bytes per chunk and
//     neighbouring lanes store neighbouring rows: fully coalesced.
//
// Kernel: persistent, one CTA per SM, warp-specialised: warp 0 = bulk-copy producer, warp 1 = MMA
// issuer (single thread) + TMEM owner, warps 2-5 = epilogue (TMEM -> registers -> bias/ReLU/residual/
// max-pool via warp shuffles -> bf16 hi/lo planes or fp32 channel-last).  mbarrier rings: A slots (one
// 64-channel K-block of one tile), weight stages (one (K-block, tap) [Bh;Bl] image), 2 TMEM accumulators.
#include <cuda_bf16.h>
#include <cstring>
#include <vector>

#include "common.h"
#include "tc.h"

namespace orca {

namespace {

constexpr int kRows = 136;                 // 128 output rows + 8 halo rows per A slot
constexpr int kASlotBytes = 2 * 8 * kRows * 16;  // hi + lo images of a 64-channel K-block
constexpr int kALoOff = 8 * kRows * 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start>>4 [0,14), LBO>>4 [16,30) = stride between the two 16-byte K chunks of one MMA,
// SBO>>4 [32,46) = stride between 8-row groups, version=1 [46,48), layout_type=0 [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split_store8(const float* v, __nv_bfloat16* hi_dst, __nv_bfloat16* lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * j]), h1 = __float2bfloat16_rn(v[2 * j + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * j] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * j + 1] - __bfloat162float(h1));
    h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void add_hilo8(float* v, const __nv_bfloat16* hi_src, const __nv_bfloat16* lo_src) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi_src));
  const uint4 l = __ldg(reinterpret_cast<const uint4*>(lo_src));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] += __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
    v[2 * j + 1] += __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
  }
}

struct TcKArgs {
  const __nv_bfloat16* in_hi; const __nv_bfloat16* in_lo;
  const uint8_t* w;
  const float* bias;
  const __nv_bfloat16* res_hi; const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  float* out_f32;
  int nb, n, npad_in, n_out, npad_out, pool, relu;
  int tiles_per_sample, total_tiles;
};

template <int C_IN, int C_OUT>
struct TcCfg {
  static constexpr int NKB = (C_IN + 63) / 64;
  static constexpr int STAGE_MAX = 2 * 8 * C_OUT * 16;  // [Bh;Bl] of a 64-channel K-block
  static constexpr int NW = C_OUT == 64 ? 9 : (C_OUT == 96 ? 4 : 3);
  static constexpr int NA = C_OUT == 64 ? 2 : 3;
  static constexpr bool RESIDENT = (9 * NKB <= NW);  // all weight stages fit: load once per CTA
  static constexpr int SMEM = NA * kASlotBytes + NW * STAGE_MAX + C_OUT * 4 + (2 * NA + 2 * NW + 4) * 8 + 16 + 128;
  __host__ __device__ static constexpr int kb_size(int kb) { return (kb == NKB - 1) ? C_IN - 64 * (NKB - 1) : 64; }
  __host__ __device__ static constexpr int stage_bytes(int kb) { return 2 * (kb_size(kb) / 8) * C_OUT * 16; }
  __host__ __device__ static constexpr int stage_offset(int kb, int tap) {
    int off = 0;
    for (int i = 0; i < kb; ++i) off += 9 * stage_bytes(i);
    return off + tap * stage_bytes(kb);
  }
};

template <int C_IN, int C_OUT>
__global__ void __launch_bounds__(192, 1) conv1d_tc_kernel(const TcKArgs a) {
  using Cfg = TcCfg<C_IN, C_OUT>;
  constexpr int NKB = Cfg::NKB, NW = Cfg::NW, NA = Cfg::NA;
  constexpr bool RESIDENT = Cfg::RESIDENT;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = sA + NA * kASlotBytes;
  float* sBias = reinterpret_cast<float*>(sW + NW * Cfg::STAGE_MAX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + C_OUT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NA + 2 * NW + 4);
  const uint32_t bA_full = smem_u32(bars), bA_empty = bA_full + 8 * NA;
  const uint32_t bW_full = bA_empty + 8 * NA, bW_empty = bW_full + 8 * NW;
  const uint32_t bAcc_full = bW_empty + 8 * NW, bAcc_empty = bAcc_full + 16;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(bA_full + 8 * i, 1); mbar_init(bA_empty + 8 * i, 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(bW_full + 8 * i, 1); mbar_init(bW_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bAcc_full + 8 * i, 1); mbar_init(bAcc_empty + 8 * i, 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < C_OUT; i += 192) sBias[i] = a.bias[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= producer: bulk copies (activation K-blocks, weight stages) =================
    if (lane == 0) {
      uint32_t a_it = 0, w_it = 0;
      int ti = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        const int b = tile / a.tiles_per_sample, t = tile - b * a.tiles_per_sample;
        const size_t row0 = (size_t)t * 128;  // padded row of the first halo row (l0 - 4 + 4)
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) {
          const int kc = Cfg::kb_size(kb) / 8;
          const uint32_t slot = a_it % NA, ph = (a_it / NA) & 1;
          mbar_wait(bA_empty + 8 * slot, ph ^ 1);
          mbar_expect_tx(bA_full + 8 * slot, 2u * kc * kRows * 16);
          const uint32_t dst = smem_u32(sA) + slot * kASlotBytes;
          for (int c = 0; c < kc; ++c) {
            const size_t plane = (size_t)b * (C_IN / 8) + kb * 8 + c;
            const size_t off = (plane * a.npad_in + row0) * 8;
            bulk_g2s(dst + c * kRows * 16, a.in_hi + off, kRows * 16, bA_full + 8 * slot);
            bulk_g2s(dst + kALoOff + c * kRows * 16, a.in_lo + off, kRows * 16, bA_full + 8 * slot);
          }
          ++a_it;
          if (RESIDENT && ti > 0) continue;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t ws = RESIDENT ? (uint32_t)(kb * 9 + tap) : w_it % NW, wph = (w_it / NW) & 1;
            if (!RESIDENT) mbar_wait(bW_empty + 8 * ws, wph ^ 1);
            mbar_expect_tx(bW_full + 8 * ws, Cfg::stage_bytes(kb));
            bulk_g2s(smem_u32(sW) + ws * Cfg::STAGE_MAX, a.w + Cfg::stage_offset(kb, tap), Cfg::stage_bytes(kb),
                     bW_full + 8 * ws);
            ++w_it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C_OUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t a_it = 0, w_it = 0, acc_it = 0;
      int ti = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
        const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
        mbar_wait(bAcc_empty + 8 * as, aph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem + as * 128;
        uint32_t accum = 0;
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) {
          const int ksteps = Cfg::kb_size(kb) / 16;
          const uint32_t slot = a_it % NA;
          mbar_wait(bA_full + 8 * slot, (a_it / NA) & 1);
          const uint32_t aBase = smem_u32(sA) + slot * kASlotBytes;
          const uint32_t bLoOff = (Cfg::kb_size(kb) / 8) * C_OUT * 16;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t ws = RESIDENT ? (uint32_t)(kb * 9 + tap) : w_it % NW;
            if (!RESIDENT || ti == 0) mbar_wait(bW_full + 8 * ws, RESIDENT ? 0u : ((w_it / NW) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bBase = smem_u32(sW) + ws * Cfg::STAGE_MAX;
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t aoff = (2 * ks) * kRows * 16 + tap * 16, boff = (2 * ks) * C_OUT * 16;
              const uint64_t ah = umma_desc(aBase + aoff, kRows * 16), al = umma_desc(aBase + kALoOff + aoff, kRows * 16);
              const uint64_t bh = umma_desc(bBase + boff, C_OUT * 16), bl = umma_desc(bBase + bLoOff + boff, C_OUT * 16);
              umma_bf16(d_tmem, ah, bh, idesc, accum);
              umma_bf16(d_tmem, al, bh, idesc, 1u);
              umma_bf16(d_tmem, ah, bl, idesc, 1u);
              accum = 1u;
            }
            if (!RESIDENT) umma_commit(bW_empty + 8 * ws);
            ++w_it;
          }
          umma_commit(bA_empty + 8 * slot);
          ++a_it;
        }
        umma_commit(bAcc_full + 8 * as);
        ++acc_it;
      }
    }
  } else {
    // ================= epilogue warps (TMEM lane quarter = warp % 4) =================
    const int q = warp & 3;
    if (blockIdx.x == 0 && a.out_hi) {  // zero the pad rows of the output planes (they are the next layer's padding)
      const int et = (warp - 2) * 32 + lane;
      const int planes = a.nb * (C_OUT / 8);
      const int tail0 = a.n_out + 4, ntail = a.npad_out - tail0;
      const int per_plane = 4 + ntail;
      for (int i = et; i < planes * per_plane; i += 128) {
        const int p = i / per_plane, j = i - p * per_plane;
        const size_t r = (size_t)p * a.npad_out + (j < 4 ? j : tail0 + (j - 4));
        reinterpret_cast<uint4*>(a.out_hi)[r] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(a.out_lo)[r] = make_uint4(0, 0, 0, 0);
      }
    }
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_sample, t = tile - b * a.tiles_per_sample;
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      mbar_wait(bAcc_full + 8 * as, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int l = t * 128 + q * 32 + lane;  // position within the sample
      const bool valid = l < a.n;
      const size_t r_in = (size_t)l + 4;
#pragma unroll 1
      for (int c0 = 0; c0 < C_OUT; c0 += 32) {
        uint32_t raw[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + as * 128 + c0, raw);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(raw[j]) + sBias[c0 + j];
          v[j] = a.relu ? fmaxf(x, 0.f) : x;
        }
        if (a.res_hi && valid) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const size_t off = (((size_t)b * (C_OUT / 8) + (c0 >> 3) + ch) * a.npad_in + r_in) * 8;
            add_hilo8(v + 8 * ch, a.res_hi + off, a.res_lo + off);
          }
        }
        if (a.pool > 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
            if (a.pool == 4) v[j] = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 2));
          }
        }
        if (valid && (a.pool == 1 || (l % a.pool) == 0)) {
          const size_t lo_row = (size_t)(l / a.pool);
          if (a.out_f32) {
            float4* dst = reinterpret_cast<float4*>(a.out_f32 + ((size_t)b * a.n_out + lo_row) * C_OUT + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const size_t off = (((size_t)b * (C_OUT / 8) + (c0 >> 3) + ch) * a.npad_out + lo_row + 4) * 8;
              split_store8(v + 8 * ch, a.out_hi + off, a.out_lo + off);
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bAcc_empty + 8 * as);
      ++acc_it;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

template <int C_IN, int C_OUT>
int launch_tc(const TcKArgs& a, int sms, cudaStream_t s) {
  using Cfg = TcCfg<C_IN, C_OUT>;
  static bool configured = false;  // per (C_IN, C_OUT) instantiation
  if (!configured) {
    ORCA_CUDA_OK(cudaFuncSetAttribute(conv1d_tc_kernel<C_IN, C_OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  const int grid = a.total_tiles < sms ? a.total_tiles : sms;
  conv1d_tc_kernel<C_IN, C_OUT><<<grid, 192, Cfg::SMEM, s>>>(a);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// ---- first layer (4 -> 64) writing chunk planes; pool-5 on planes ---------------------------------
__global__ void __launch_bounds__(256) conv_first_planes_kernel(const float* __restrict__ x, long long sB, long long sC,
                                                                long long sL, long long Ltot, long long l_begin,
                                                                long long n, int npad, const float* __restrict__ w,
                                                                const float* __restrict__ bias,
                                                                __nv_bfloat16* __restrict__ out_hi,
                                                                __nv_bfloat16* __restrict__ out_lo) {
  constexpr int TP = 128, HALO = 4, NX = TP + 2 * HALO;
  __shared__ __align__(16) float Xs[NX][4];
  __shared__ __align__(16) float Ws[9 * 4 * 64];
  __shared__ __align__(16) __nv_bfloat16 Th[8][TP][8];
  __shared__ __align__(16) __nv_bfloat16 Tl[8][TP][8];
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * TP;
  const float* xb = x + (long long)b * sB;
  for (int i = tid; i < 9 * 4 * 64; i += 256) Ws[i] = __ldg(w + i);
  if (sC == 1 || sC == -1) {
    for (int idx = tid; idx < NX * 4; idx += 256) {
      const int j = idx >> 2, c = idx & 3;
      const long long l = l_begin + t0 - HALO + j;
      Xs[j][c] = (l >= 0 && l < Ltot) ? __ldg(xb + l * sL + c * sC) : 0.f;
    }
  } else {
    for (int idx = tid; idx < NX * 4; idx += 256) {
      const int c = idx / NX, j = idx - c * NX;
      const long long l = l_begin + t0 - HALO + j;
      Xs[j][c] = (l >= 0 && l < Ltot) ? __ldg(xb + l * sL + c * sC) : 0.f;
    }
  }
  __syncthreads();
  const int cg = tid & 15, pgp = tid >> 4;
  float4 xw[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) xw[i] = *reinterpret_cast<const float4*>(&Xs[pgp * 8 + i][0]);
  float4 acc[8];
  const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + cg * 4));
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = bv;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const float4 wv = *reinterpret_cast<const float4*>(&Ws[(t * 4 + ci) * 64 + cg * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xv = reinterpret_cast<const float*>(&xw[i + t])[ci];
        acc[i].x = fmaf(xv, wv.x, acc[i].x); acc[i].y = fmaf(xv, wv.y, acc[i].y);
        acc[i].z = fmaf(xv, wv.z, acc[i].z); acc[i].w = fmaf(xv, wv.w, acc[i].w);
      }
    }
  // stage through shared memory so that the plane stores are coalesced 16-byte rows
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float vv[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
    __nv_bfloat16* th = &Th[cg >> 1][pgp * 8 + i][(cg & 1) * 4];
    __nv_bfloat16* tl = &Tl[cg >> 1][pgp * 8 + i][(cg & 1) * 4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 h = __float2bfloat16_rn(vv[j]);
      th[j] = h;
      tl[j] = __float2bfloat16_rn(vv[j] - __bfloat162float(h));
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 8 * TP; idx += 256) {
    const int ch = idx / TP, r = idx - ch * TP;
    const long long row = t0 + r;
    if (row < n) {
      const size_t off = (((size_t)b * 8 + ch) * npad + row + 4) * 8;
      *reinterpret_cast<uint4*>(out_hi + off) = *reinterpret_cast<const uint4*>(&Th[ch][r][0]);
      *reinterpret_cast<uint4*>(out_lo + off) = *reinterpret_cast<const uint4*>(&Tl[ch][r][0]);
    }
  }
  if (blockIdx.x == 0) {  // pad rows of this sample's planes
    const int tail0 = (int)n + 4, ntail = npad - tail0, per_plane = 4 + ntail;
    for (int i = tid; i < 8 * per_plane; i += 256) {
      const int p = i / per_plane, j = i - p * per_plane;
      const size_t r = ((size_t)b * 8 + p) * npad + (j < 4 ? j : tail0 + (j - 4));
      reinterpret_cast<uint4*>(out_hi)[r] = make_uint4(0, 0, 0, 0);
      reinterpret_cast<uint4*>(out_lo)[r] = make_uint4(0, 0, 0, 0);
    }
  }
}

// MaxPool1d(p) on chunk planes (used for p = 5, where a 128-row tile is not a whole number of groups)
__global__ void pool_planes_kernel(const __nv_bfloat16* __restrict__ in_hi, const __nv_bfloat16* __restrict__ in_lo,
                                   __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int planes,
                                   int n_out, int npad_in, int npad_out, int p) {
  const int per_plane = npad_out;  // also writes the zero pad rows
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)planes * per_plane;
       i += (long long)gridDim.x * blockDim.x) {
    const int pl = (int)(i / per_plane), r = (int)(i - (long long)pl * per_plane);
    const int lo_row = r - 4;
    float m[8];
    if (lo_row < 0 || lo_row >= n_out) {
      reinterpret_cast<uint4*>(out_hi)[i] = make_uint4(0, 0, 0, 0);
      reinterpret_cast<uint4*>(out_lo)[i] = make_uint4(0, 0, 0, 0);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int k = 0; k < p; ++k) {
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const size_t off = ((size_t)pl * npad_in + (size_t)lo_row * p + k + 4) * 8;
      add_hilo8(v, in_hi + off, in_lo + off);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
    split_store8(m, out_hi + (size_t)i * 8, out_lo + (size_t)i * 8);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host API (tc.h)
// ---------------------------------------------------------------------------------------------------
static inline uint16_t bf16_bits_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf16_to_f32(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

bool tc_layer_eligible(const ConvLayer& L) {
  return L.kh == 1 && L.kw == 9 && (L.c_in == 64 || L.c_in == 96 || L.c_in == 128) &&
         (L.c_out == 64 || L.c_out == 96 || L.c_out == 128) && L.c_out >= L.c_in;
}

// Stage images in consumption order: for K-block kb, for tap: [Bh][Bl], each [k-chunk][c_out][8] bf16.
int tc_pack_layer(ConvLayer& L, const float* w /*[tap][c_in][c_out]*/, std::vector<void*>& allocs) {
  if (!tc_layer_eligible(L)) return ORCA_B200_OK;
  const int nkb = (L.c_in + 63) / 64;
  std::vector<uint16_t> img;
  img.reserve((size_t)9 * L.c_in * L.c_out * 2);
  for (int kb = 0; kb < nkb; ++kb) {
    const int ks = (kb == nkb - 1) ? L.c_in - 64 * (nkb - 1) : 64;
    for (int tap = 0; tap < 9; ++tap)
      for (int part = 0; part < 2; ++part)
        for (int c = 0; c < ks / 8; ++c)
          for (int n = 0; n < L.c_out; ++n)
            for (int j = 0; j < 8; ++j) {
              const int ci = kb * 64 + c * 8 + j;
              const float v = w[((size_t)tap * L.c_in + ci) * L.c_out + n];
              const uint16_t h = bf16_bits_rn(v);
              img.push_back(part == 0 ? h : bf16_bits_rn(v - bf16_to_f32(h)));
            }
  }
  void* d = nullptr;
  ORCA_CUDA_OK(cudaMalloc(&d, img.size() * 2));
  allocs.push_back(d);
  ORCA_CUDA_OK(cudaMemcpy(d, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  L.tc_w = d;
  L.tc_w_bytes = img.size() * 2;
  return ORCA_B200_OK;
}

bool tc_supported(const ConvLayer&, const ConvCall&) { return false; }  // fp32 channel-last calls stay on SIMT
int conv_tc(const ConvLayer&, const ConvCall&, cudaStream_t) {
  set_error("conv_tc: the tcgen05 path takes chunk-plane activations (tc_conv1d)");
  return ORCA_B200_EUNSUPPORTED;
}

static int g_sms = 0;
static int sm_count() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

int tc_conv1d(const ConvLayer& L, const TcAct& in, const TcAct* res, TcAct* out_planes, float* out_f32, int pool,
              int relu, cudaStream_t s) {
  if (!L.tc_w) { set_error("tc_conv1d: layer %d->%d has no tensor-core weights", L.c_in, L.c_out); return ORCA_B200_EUNSUPPORTED; }
  if (in.C != L.c_in || (pool != 1 && pool != 2 && pool != 4) || in.n % pool != 0 || (out_planes == nullptr) == (out_f32 == nullptr)) {
    set_error("tc_conv1d: bad call (C=%d c_in=%d pool=%d n=%lld)", in.C, L.c_in, pool, (long long)in.n);
    return ORCA_B200_EINVAL;
  }
  TcKArgs a;
  a.in_hi = static_cast<const __nv_bfloat16*>(in.hi); a.in_lo = static_cast<const __nv_bfloat16*>(in.lo);
  a.w = static_cast<const uint8_t*>(L.tc_w); a.bias = L.b;
  a.res_hi = res ? static_cast<const __nv_bfloat16*>(res->hi) : nullptr;
  a.res_lo = res ? static_cast<const __nv_bfloat16*>(res->lo) : nullptr;
  a.out_hi = out_planes ? static_cast<__nv_bfloat16*>(out_planes->hi) : nullptr;
  a.out_lo = out_planes ? static_cast<__nv_bfloat16*>(out_planes->lo) : nullptr;
  a.out_f32 = out_f32;
  a.nb = in.nb; a.n = (int)in.n; a.npad_in = (int)in.npad; a.n_out = (int)(in.n / pool);
  a.npad_out = out_planes ? (int)out_planes->npad : 0;
  a.pool = pool; a.relu = relu;
  a.tiles_per_sample = (int)((in.n + 127) / 128); a.total_tiles = a.tiles_per_sample * in.nb;
  if (res && (res->C != L.c_out || res->n != in.n || res->npad != in.npad || res->nb != in.nb)) { set_error("tc_conv1d: residual geometry mismatch"); return ORCA_B200_EINVAL; }
  if (out_planes && (out_planes->C != L.c_out || out_planes->n != in.n / pool || out_planes->nb != in.nb)) { set_error("tc_conv1d: output geometry mismatch"); return ORCA_B200_EINVAL; }
  if (a.total_tiles <= 0) return ORCA_B200_OK;
  const int sms = sm_count();
  const int key = L.c_in * 1000 + L.c_out;
  switch (key) {
    case 64064: return launch_tc<64, 64>(a, sms, s);
    case 64096: return launch_tc<64, 96>(a, sms, s);
    case 96096: return launch_tc<96, 96>(a, sms, s);
    case 96128: return launch_tc<96, 128>(a, sms, s);
    case 128128: return launch_tc<128, 128>(a, sms, s);
    default: set_error("tc_conv1d: no kernel for %d->%d", L.c_in, L.c_out); return ORCA_B200_EUNSUPPORTED;
  }
}

int tc_conv_first(const ConvLayer& L, const float* x, int64_t sB, int64_t sC, int64_t sL, int nb, int64_t Ltot,
                  int64_t l_begin, int64_t n, TcAct* out, cudaStream_t s) {
  if (L.c_in != 4 || L.c_out != 64 || out->C != 64 || out->n != n || out->nb != nb) { set_error("tc_conv_first: bad geometry"); return ORCA_B200_EINVAL; }
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)nb), block(256);
  conv_first_planes_kernel<<<grid, block, 0, s>>>(x, sB, sC, sL, Ltot, l_begin, n, (int)out->npad, L.w, L.b,
                                                  static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo));
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

int tc_pool_planes(const TcAct& in, TcAct* out, int p, cudaStream_t s) {
  if (in.n % p || out->n != in.n / p || out->C != in.C || out->nb != in.nb) { set_error("tc_pool_planes: bad geometry"); return ORCA_B200_EINVAL; }
  const int planes = in.nb * (in.C / 8);
  const long long total = (long long)planes * out->npad;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  pool_planes_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(in.hi), static_cast<const __nv_bfloat16*>(in.lo),
                                          static_cast<__nv_bfloat16*>(out->hi), static_cast<__nv_bfloat16*>(out->lo), planes,
                                          (int)out->n, (int)in.npad, (int)out->npad, p);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
