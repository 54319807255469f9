// tcgen05 (5th-gen tensor core) implicit-GEMM Conv1d(k=9) for sm_100a  --  the Encoder hot loop
// (orca_modules.py:811-950) and the 128-channel U-nets (:991-1169, :1181-1276, :1286-1406).
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
// issuer + TMEM owner, warps 2-5 = epilogue (TMEM -> registers -> bias/ReLU/residual/max-pool via warp
// shuffles -> bf16 hi/lo planes and/or fp32 channel-last).  Producer and issuer run their loops
// warp-uniformly and elect one lane only for the issue instructions; UMMA descriptors are a base word
// plus compile-time offsets (loops over taps / K steps are fully unrolled).  mbarrier rings: A slots (one
// 64-channel K-block of one tile), weight stages (one (K-block, tap) image), 2 TMEM accumulators.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"
#include "tc.h"
#include "tc_device.cuh"
#include "seq_in.cuh"

namespace orca {

namespace {
using namespace tcdev;

constexpr int kRows = 136;                       // 128 output rows + 8 halo rows per A slot
constexpr int kALoOff = 8 * kRows * 16;          // FMT 0: the lo image of a 64-channel K-block follows the hi image
constexpr int kEpiWarps = 8;                    // epilogue warps per accumulator: two per TMEM lane quarter, each takes every other 32-column group
// The single-pass fp16 64 -> 64 kernel (stage 1: 32 M positions per strand) issues few MMAs per tile, so its epilogue
// (residual, ReLU, fused max-pool) is what paces the residual + pool layer: it runs TWO sets of epilogue warps, one per TMEM
// accumulator (set s drains the tiles that land in accumulator s), i.e. two tiles' epilogues in flight at once (measured:
// 11.4 -> 9.7 ms per step).  The 96-wide variants spill under the 576-thread register budget (96 registers) and need a
// smaller pool buffer next to their 166 KB of resident weights: measured slower (8.5 -> 9.1-9.5 ms), so they keep one set.
__host__ __device__ constexpr int epi_sets(int fmt, int c_out) { return (fmt && c_out == 64) ? 2 : 1; }
__host__ __device__ constexpr int threads_of(int fmt, int c_out) { return 64 + 32 * kEpiWarps * epi_sets(fmt, c_out); }  // warp 0 producer, warp 1 MMA issuer, then the epilogue warps

struct TcKArgs {
  const __nv_bfloat16* in_hi; const __nv_bfloat16* in_lo;
  const uint8_t* w;
  const float* bias;
  const __nv_bfloat16* res_hi; const __nv_bfloat16* res_lo;
  const __nv_bfloat16* res2_hi; const __nv_bfloat16* res2_lo;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  float* out_f32;
  unsigned int* sat;  // fp16 range guard word (FMT 1 outputs), or nullptr
  int nb, n, npad_in, n_out, npad_out, pool, relu;
  int tiles_per_sample, total_tiles;
};

// FMT 0: operands are bf16 hi/lo pairs, three products per conv (fp32-grade accuracy).
// FMT 1: operands are single fp16 images, ONE product per conv: a third of the tensor-core work and half the
//        activation bytes.  Used for the early encoder stages, whose rounding noise (2^-12 per element,
//        incoherent) is averaged away by the pooling + 1152-term sums of the later stages.
template <int C_IN, int C_OUT, int FMT>
struct TcCfg {
  static constexpr int NKB = (C_IN + 63) / 64;
  static constexpr int PARTS = FMT ? 1 : 2;               // operand images per K chunk
  static constexpr int A_SLOT = PARTS * 8 * kRows * 16;   // one 64-channel K-block of a tile
  static constexpr int STAGE_MAX = PARTS * 8 * C_OUT * 16;  // weights of one (64-channel K-block, tap)
  // single-pass 96 -> 96: all 18 (K-block, tap) stages stay resident (166 KB).  Re-streaming them for every tile made the
  // layer shared-memory-bandwidth bound (writes of 166 KB + operand reads of ~380 KB per tile through one 128 B/clk port).
  static constexpr bool BIG_RESIDENT = FMT && C_IN == 96 && C_OUT == 96;
  static constexpr int NW = BIG_RESIDENT ? 18 : FMT ? (C_OUT == 128 ? 6 : 9) : (C_OUT == 64 ? 9 : (C_OUT == 96 ? 4 : 3));
  // FMT 2 = FMT 1 with the shared memory re-balanced for the residual + max-pool layer of the resident 96 -> 96 kernel: that
  // layer is epilogue-bound (ncu: 2.05 ms vs 1.04 ms plain, tensor pipe 31 %), so it gives up one A slot for a pool buffer
  // that lets all 32 lanes of an epilogue warp reduce in one pass per 32-channel group instead of half of them in two.
  static constexpr int NA = BIG_RESIDENT ? (FMT == 2 ? 2 : 3) : FMT ? 4 : (C_OUT == 64 ? 2 : 3);
  static constexpr int POOL_CHUNKS = BIG_RESIDENT ? (FMT == 2 ? 4 : 2) : 4;  // 8-channel chunks per pass of the fused max-pool buffer
  static constexpr int ACC_STRIDE_CAT = C_OUT == 64 ? 128 : 256;  // TMEM columns per accumulator, concat mode
  static constexpr bool RESIDENT = (9 * NKB <= NW);  // all weight stages fit: load once per CTA
  static constexpr int POOL_SCRATCH = FMT ? epi_sets(FMT, C_OUT) * kEpiWarps * POOL_CHUNKS * 512 : 0;  // per-epilogue-warp transpose buffer of the fused max-pool
  // resident stages are packed back to back (a 32-channel K-block's stages are half size), streamed ones use a ring of
  // full-size slots
  static constexpr int W_BYTES = RESIDENT ? 9 * PARTS * (C_IN / 8) * C_OUT * 16 : NW * STAGE_MAX;
  static constexpr int SMEM = NA * A_SLOT + W_BYTES + POOL_SCRATCH + C_OUT * 4 + (2 * NA + 2 * NW + 4) * 8 + 16 + 128;
  __host__ __device__ static constexpr int kb_size(int kb) { return (kb == NKB - 1) ? C_IN - 64 * (NKB - 1) : 64; }
  __host__ __device__ static constexpr int stage_bytes(int kb) { return PARTS * (kb_size(kb) / 8) * C_OUT * 16; }
  __host__ __device__ static constexpr int stage_offset(int kb, int tap) {
    int off = 0;
    for (int i = 0; i < kb; ++i) off += 9 * stage_bytes(i);
    return off + tap * stage_bytes(kb);
  }
};

// CONCAT: the two products that share A = Ah run as ONE MMA against B = [Bh;Bl] (N = 2*C_OUT, two
// accumulator column blocks summed in the epilogue) -- 2 MMAs and 2 A-operand reads per K step instead of 3.
template <int C_IN, int C_OUT, bool CONCAT, int FMT>
__global__ void __launch_bounds__(threads_of(FMT, C_OUT), 1) conv1d_tc_kernel(const TcKArgs a) {
  using Cfg = TcCfg<C_IN, C_OUT, FMT>;
  constexpr int kThreads = threads_of(FMT, C_OUT);
  static_assert(!(FMT && CONCAT), "the single-pass format has nothing to concatenate");
  constexpr int NKB = Cfg::NKB, NW = Cfg::NW, NA = Cfg::NA;
  constexpr int kASlotBytes = Cfg::A_SLOT;
  constexpr bool RESIDENT = Cfg::RESIDENT;
  constexpr uint32_t ACC_STRIDE = FMT ? (C_OUT == 64 ? 64 : 128) : (CONCAT ? Cfg::ACC_STRIDE_CAT : 128);
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = sA + NA * kASlotBytes;
  uint8_t* sPool = sW + Cfg::W_BYTES;
  float* sBias = reinterpret_cast<float*>(sPool + Cfg::POOL_SCRATCH);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + C_OUT);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NA + 2 * NW + 4);
  const uint32_t bA_full = smem_u32(bars), bA_empty = bA_full + 8 * NA;
  const uint32_t bW_full = bA_empty + 8 * NA, bW_empty = bW_full + 8 * NW;
  const uint32_t bAcc_full = bW_empty + 8 * NW, bAcc_empty = bAcc_full + 16;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(bA_full + 8 * i, 1); mbar_init(bA_empty + 8 * i, 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(bW_full + 8 * i, 1); mbar_init(bW_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bAcc_full + 8 * i, 1); mbar_init(bAcc_empty + 8 * i, kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < C_OUT; i += kThreads) sBias[i] = a.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================= producer: bulk copies (activation K-blocks, weight stages) =================
    uint32_t a_it = 0, w_it = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
      const int b = tile / a.tiles_per_sample, t = tile - b * a.tiles_per_sample;
      const size_t row0 = (size_t)t * 128;  // padded row of the first halo row (l0 - 4 + 4)
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const int kc = Cfg::kb_size(kb) / 8;
        const uint32_t slot = a_it % NA, ph = (a_it / NA) & 1;
        mbar_wait(bA_empty + 8 * slot, ph ^ 1);
        if (elect_one()) mbar_expect_tx(bA_full + 8 * slot, (uint32_t)Cfg::PARTS * kc * kRows * 16);
        __syncwarp();
        if (lane < Cfg::PARTS * kc) {  // one bulk copy per lane: issued back to back by a single lane they cost ~65 cycles each
          const int c = FMT ? lane : (lane >> 1), part = FMT ? 0 : (lane & 1);
          const uint32_t dst = smem_u32(sA) + slot * kASlotBytes + (part ? kALoOff : 0) + c * kRows * 16;
          const size_t off = (((size_t)b * (C_IN / 8) + kb * 8 + c) * a.npad_in + row0) * 8;
          bulk_g2s(dst, (part ? a.in_lo : a.in_hi) + off, kRows * 16, bA_full + 8 * slot);
        }
        __syncwarp();
        ++a_it;
        if (RESIDENT && ti > 0) continue;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t ws = RESIDENT ? (uint32_t)(kb * 9 + tap) : w_it % NW, wph = (w_it / NW) & 1;
          if (!RESIDENT) mbar_wait(bW_empty + 8 * ws, wph ^ 1);
          if (elect_one()) {
            mbar_expect_tx(bW_full + 8 * ws, Cfg::stage_bytes(kb));
            bulk_g2s(smem_u32(sW) + (RESIDENT ? (uint32_t)Cfg::stage_offset(kb, tap) : ws * Cfg::STAGE_MAX),
                     a.w + Cfg::stage_offset(kb, tap), Cfg::stage_bytes(kb), bW_full + 8 * ws);
          }
          __syncwarp();
          ++w_it;
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues =================
    constexpr uint32_t idesc = FMT ? umma_idesc_f16(C_OUT) : umma_idesc_bf16(C_OUT), idesc_cat = umma_idesc_bf16(2 * C_OUT);
    constexpr uint32_t bLbo = Cfg::PARTS * C_OUT * 16;  // chunk image = [Bh rows][Bl rows] (FMT 0) or [B rows] (FMT 1)
    uint32_t a_it = 0, w_it = 0, acc_it = 0;
    int ti = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x, ++ti) {
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      mbar_wait(bAcc_empty + 8 * as, aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + as * ACC_STRIDE;
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        const int ksteps = Cfg::kb_size(kb) / 16;
        const uint32_t slot = a_it % NA;
        mbar_wait(bA_full + 8 * slot, (a_it / NA) & 1);
        const uint32_t aLo = umma_desc_lo(smem_u32(sA) + slot * kASlotBytes, kRows * 16);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t ws = RESIDENT ? (uint32_t)(kb * 9 + tap) : w_it % NW;
          if (!RESIDENT || ti == 0) mbar_wait(bW_full + 8 * ws, RESIDENT ? 0u : ((w_it / NW) & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint32_t bLo = umma_desc_lo(smem_u32(sW) + (RESIDENT ? (uint32_t)Cfg::stage_offset(kb, tap) : ws * Cfg::STAGE_MAX), bLbo);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < ksteps) {
                const uint32_t ao = aLo + ks * ((2 * kRows * 16) >> 4) + tap;  // + tap rows of 16 B
                const uint64_t ah = umma_desc64(ao), al = umma_desc64(ao + (kALoOff >> 4));
                const uint64_t bh = umma_desc64(bLo + ks * ((2 * bLbo) >> 4));
                const uint32_t accum = (kb | tap | ks) != 0 ? 1u : 0u;
                if (FMT) {
                  umma_bf16(d_tmem, ah, bh, idesc, accum);      // fp16 A * fp16 B, the only product
                } else if (CONCAT) {
                  umma_bf16(d_tmem, ah, bh, idesc_cat, accum);  // [Ah*Bh | Ah*Bl]
                  umma_bf16(d_tmem, al, bh, idesc, 1u);         // += Al*Bh into the first block
                } else {
                  const uint64_t bl = umma_desc64(bLo + ks * ((2 * bLbo) >> 4) + ((C_OUT * 16) >> 4));
                  umma_bf16(d_tmem, ah, bh, idesc, accum);
                  umma_bf16(d_tmem, al, bh, idesc, 1u);
                  umma_bf16(d_tmem, ah, bl, idesc, 1u);
                }
              }
            }
            if (!RESIDENT) umma_commit(bW_empty + 8 * ws);
            if (tap == 8) {
              umma_commit(bA_empty + 8 * slot);
              if (kb == NKB - 1) umma_commit(bAcc_full + 8 * as);
            }
          }
          __syncwarp();
          ++w_it;
        }
        ++a_it;
      }
      ++acc_it;
    }
  } else {
    // ================= epilogue warps (TMEM lane quarter = warp % 4) =================
    // The epilogue is instruction-latency bound per warp (one warp per scheduler, long dependent chains), so
    // two warps share each lane quarter: warp (q, h) handles the 32-column groups h, h+2, ...
    const int eset = (warp - 2) / kEpiWarps;  // which accumulator's tiles this warp drains (FMT 1: two sets)
    const int q = warp & 3, h = ((warp - 2) % kEpiWarps) >> 2;
    if (blockIdx.x == 0 && a.out_hi) {  // zero the pad rows of the output planes (they are the next layer's padding)
      const int et = (warp - 2) * 32 + lane;
      const int planes = a.nb * (C_OUT / 8);
      const int tail0 = a.n_out + 4, ntail = a.npad_out - tail0;
      const int per_plane = 4 + ntail;
      for (int i = et; i < planes * per_plane; i += 32 * kEpiWarps * epi_sets(FMT, C_OUT)) {
        const int p = i / per_plane, j = i - p * per_plane;
        const size_t r = (size_t)p * a.npad_out + (j < 4 ? j : tail0 + (j - 4));
        reinterpret_cast<uint4*>(a.out_hi)[r] = make_uint4(0, 0, 0, 0);
        if (a.out_lo) reinterpret_cast<uint4*>(a.out_lo)[r] = make_uint4(0, 0, 0, 0);
      }
    }
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int b = tile / a.tiles_per_sample, t = tile - b * a.tiles_per_sample;
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      if (epi_sets(FMT, C_OUT) == 2 && (int)as != eset) { ++acc_it; continue; }  // the other set's tile
      const int l = t * 128 + q * 32 + lane;  // position within the sample
      const bool valid = l < a.n;
      const size_t r_in = (size_t)l + 4;
      // residual(s) of a 32-channel group, fetched one group ahead (the first one before the accumulator
      // wait) so that their HBM/L2 latency overlaps the tile's MMAs instead of serialising the epilogue
      auto load_res = [&](int c0, float* dst) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dst[j] = 0.f;
        if (a.res_hi && valid) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const size_t off = (((size_t)b * (C_OUT / 8) + (c0 >> 3) + ch) * a.npad_in + r_in) * 8;
            if (FMT) {  // residuals share the input's format
              add_h8(dst + 8 * ch, a.res_hi + off);
              if (a.res2_hi) add_h8(dst + 8 * ch, a.res2_hi + off);
            } else {
              add_hilo8(dst + 8 * ch, a.res_hi + off, a.res_lo + off);
              if (a.res2_hi) add_hilo8(dst + 8 * ch, a.res2_hi + off, a.res2_lo + off);
            }
          }
        }
      };
      float rcur[32];
      if (32 * h < C_OUT) load_res(32 * h, rcur);
      mbar_wait(bAcc_full + 8 * as, aph);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < (C_OUT + 63) / 64; ++g) {
        const int c0 = 32 * h + 64 * g;
        if (c0 >= C_OUT) break;
        uint32_t raw[32];
        float v[32], rnext[32];
        if (c0 + 64 < C_OUT) load_res(c0 + 64, rnext);
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE + c0, raw);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (CONCAT) {
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + as * ACC_STRIDE + C_OUT + c0, raw);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(raw[j]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = v[j] + sBias[c0 + j];
          v[j] = a.relu ? fmaxf(x, 0.f) : x;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += rcur[j];
        if (c0 + 64 < C_OUT) {
#pragma unroll
          for (int j = 0; j < 32; ++j) rcur[j] = rnext[j];
        }
        if (FMT && a.out_hi && !a.out_lo) guard_h<32>(v, a.sat);
        if (FMT && a.pool > 1 && a.out_hi && !a.out_lo) {
          // Fused max-pool of the single-pass format through shared memory.  TMEM lane = position, so pooling over
          // 2 / 4 neighbouring positions with warp shuffles costs ~4 instructions per value and made the residual+pool
          // layers twice as slow as the plain ones (ncu: 567 vs 285 us).  Instead every lane rounds its 32 channels
          // to fp16 (max commutes with the monotonic rounding), parks them in a [chunk][position][16 B] buffer
          // (position index XOR-swizzled so that both the writes and the strided reads are bank-conflict free), and
          // each lane then reduces one (pooled row, 8-channel chunk) with packed fp16 max and stores 16 coalesced bytes.
          constexpr int CHP = Cfg::POOL_CHUNKS;  // chunks per pass (the buffer holds CHP x 32 positions x 16 B per warp)
          uint8_t* scr = sPool + (warp - 2) * (CHP * 512);
          const int fpos = (lane & ~7) | ((lane ^ (lane >> 3)) & 7);
          const int P = a.pool, rows = 32 / P;
#pragma unroll
          for (int pass = 0; pass < 4 / CHP; ++pass) {
#pragma unroll
            for (int ch = 0; ch < CHP; ++ch) store_h8(v + 8 * (pass * CHP + ch), scr + (ch * 32 + fpos) * 16);
            __syncwarp();
            for (int item = lane; item < rows * CHP; item += 32) {
              const int pr = item & (rows - 1), ch = item / rows;
              __half2 m[4];
#pragma unroll 1
              for (int k = 0; k < P; ++k) {
                const int pos = pr * P + k, sp = (pos & ~7) | ((pos ^ (pos >> 3)) & 7);
                const uint4 u = *reinterpret_cast<const uint4*>(scr + (ch * 32 + sp) * 16);
                const __half2* hv = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] = k == 0 ? hv[j] : __hmax2(m[j], hv[j]);
              }
              const size_t orow = (size_t)(t * 128 + q * 32) / P + pr;
              if (orow < (size_t)a.n_out) {
                const size_t off = (((size_t)b * (C_OUT / 8) + (c0 >> 3) + pass * CHP + ch) * a.npad_out + orow + 4) * 8;
                *reinterpret_cast<uint4*>(a.out_hi + off) = *reinterpret_cast<const uint4*>(m);
              }
            }
            __syncwarp();
          }
          continue;
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
          }
          if (a.out_hi) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const size_t off = (((size_t)b * (C_OUT / 8) + (c0 >> 3) + ch) * a.npad_out + lo_row + 4) * 8;
              if (a.out_lo) split_store8(v + 8 * ch, a.out_hi + off, a.out_lo + off);  // bf16 hi/lo planes
              else store_h8(v + 8 * ch, a.out_hi + off);                               // one fp16 plane
            }
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
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

template <int C_IN, int C_OUT, bool CONCAT, int FMT>
int launch_tc_impl(const TcKArgs& a, int sms, cudaStream_t s) {
  using Cfg = TcCfg<C_IN, C_OUT, FMT>;
  static bool configured_dev[32] = {};  // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 31];
  if (!configured) {
    ORCA_CUDA_OK(cudaFuncSetAttribute(conv1d_tc_kernel<C_IN, C_OUT, CONCAT, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    configured = true;
  }
  const int grid = a.total_tiles < sms ? a.total_tiles : sms;
  conv1d_tc_kernel<C_IN, C_OUT, CONCAT, FMT><<<grid, threads_of(FMT, C_OUT), Cfg::SMEM, s>>>(a);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

// bit i of the mask selects CONCAT for C_OUT = 64 (bit 0), 96 (bit 1), 128 (bit 2); env ORCA_B200_TC_CONCAT
int concat_mask() {
  static int mask = -1;
  if (mask < 0) {
    const char* e = getenv("ORCA_B200_TC_CONCAT");
    mask = e ? atoi(e) : 7;
  }
  return mask;
}

template <int C_IN, int C_OUT>
int launch_tc(const TcKArgs& a, int sms, cudaStream_t s, int fmt) {
  if (fmt) {
    if (C_IN == 96 && C_OUT == 96 && a.pool > 1 && a.out_hi && !a.out_lo) return launch_tc_impl<C_IN, C_OUT, false, (C_IN == 96 && C_OUT == 96) ? 2 : 1>(a, sms, s);
    return launch_tc_impl<C_IN, C_OUT, false, 1>(a, sms, s);
  }
  const int bit = C_OUT == 64 ? 1 : (C_OUT == 96 ? 2 : 4);
  if (concat_mask() & bit) return launch_tc_impl<C_IN, C_OUT, true, 0>(a, sms, s);
  return launch_tc_impl<C_IN, C_OUT, false, 0>(a, sms, s);
}

// ---- first layer (4 -> 64) writing chunk planes; pool-5 on planes ---------------------------------
__global__ void __launch_bounds__(256) conv_first_planes_kernel(const SeqIn in, long long Ltot, long long l_begin,
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
  for (int i = tid; i < 9 * 4 * 64; i += 256) Ws[i] = __ldg(w + i);
  for (int j = tid; j < NX; j += 256) {  // one position per thread: fp32 view or packed bases (seq_in.cuh)
    const long long l = l_begin + t0 - HALO + j;
    *reinterpret_cast<float4*>(&Xs[j][0]) = (l >= 0 && l < Ltot) ? seq_load(in, b, l) : make_float4(0.f, 0.f, 0.f, 0.f);
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
  // in_lo == NULL: the input is one fp16 plane; out_lo == NULL: write one fp16 plane (this kernel is also where the
  // single-pass encoder stages hand over to the bf16 hi/lo stages)
  const int per_plane = npad_out;  // also writes the zero pad rows
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)planes * per_plane;
       i += (long long)gridDim.x * blockDim.x) {
    const int pl = (int)(i / per_plane), r = (int)(i - (long long)pl * per_plane);
    const int lo_row = r - 4;
    float m[8];
    if (lo_row < 0 || lo_row >= n_out) {
      reinterpret_cast<uint4*>(out_hi)[i] = make_uint4(0, 0, 0, 0);
      if (out_lo) reinterpret_cast<uint4*>(out_lo)[i] = make_uint4(0, 0, 0, 0);
      continue;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int k = 0; k < p; ++k) {
      float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const size_t off = ((size_t)pl * npad_in + (size_t)lo_row * p + k + 4) * 8;
      if (in_lo) add_hilo8(v, in_hi + off, in_lo + off);
      else add_h8(v, in_hi + off);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
    if (out_lo) split_store8(m, out_hi + (size_t)i * 8, out_lo + (size_t)i * 8);
    else store_h8(m, out_hi + (size_t)i * 8);
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

// Stage images in consumption order: for K-block kb, for tap: [k-chunk][Bh rows | Bl rows][8] bf16, i.e. one
// K-major operand of 2*c_out rows whose first/second half are the hi/lo parts of the folded weights.
int tc_pack_layer(ConvLayer& L, const float* w /*[tap][c_in][c_out]*/, std::vector<void*>& allocs) {
  if (!tc_layer_eligible(L)) return ORCA_B200_OK;
  const int nkb = (L.c_in + 63) / 64;
  std::vector<uint16_t> img;
  img.reserve((size_t)9 * L.c_in * L.c_out * 2);
  for (int kb = 0; kb < nkb; ++kb) {
    const int ks = (kb == nkb - 1) ? L.c_in - 64 * (nkb - 1) : 64;
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < ks / 8; ++c)
        for (int part = 0; part < 2; ++part)
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
  // fp16 image for the single-pass format: same stage order, [k-chunk][c_out rows][8] fp16 per (K-block, tap)
  std::vector<uint16_t> img16;
  img16.reserve((size_t)9 * L.c_in * L.c_out);
  for (int kb = 0; kb < nkb; ++kb) {
    const int ks = (kb == nkb - 1) ? L.c_in - 64 * (nkb - 1) : 64;
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < ks / 8; ++c)
        for (int n = 0; n < L.c_out; ++n)
          for (int j = 0; j < 8; ++j) {
            const __half hv = __float2half_rn(w[((size_t)tap * L.c_in + kb * 64 + c * 8 + j) * L.c_out + n]);
            uint16_t bits;
            memcpy(&bits, &hv, 2);
            img16.push_back(bits);
          }
  }
  void* d16 = nullptr;
  ORCA_CUDA_OK(cudaMalloc(&d16, img16.size() * 2));
  allocs.push_back(d16);
  ORCA_CUDA_OK(cudaMemcpy(d16, img16.data(), img16.size() * 2, cudaMemcpyHostToDevice));
  L.tc_w16 = d16;
  return ORCA_B200_OK;
}

bool tc_supported(const ConvLayer&, const ConvCall&) { return false; }  // fp32 channel-last calls stay on SIMT
int conv_tc(const ConvLayer&, const ConvCall&, cudaStream_t) {
  set_error("conv_tc: the tcgen05 path takes chunk-plane activations (tc_conv1d)");
  return ORCA_B200_EUNSUPPORTED;
}

static int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int tc_conv1d(const ConvLayer& L, const TcAct& in, const TcAct* res, TcAct* out_planes, float* out_f32, int pool,
              int relu, cudaStream_t s, const TcAct* res2) {
  const int fmt = in.fmt;
  if (!(fmt ? L.tc_w16 : L.tc_w)) { set_error("tc_conv1d: layer %d->%d has no tensor-core weights", L.c_in, L.c_out); return ORCA_B200_EUNSUPPORTED; }
  if ((res && res->fmt != fmt) || (res2 && res2->fmt != fmt)) { set_error("tc_conv1d: residual format differs from the input format"); return ORCA_B200_EINVAL; }
  if (in.C != L.c_in || (pool != 1 && pool != 2 && pool != 4) || in.n % pool != 0 || (out_planes == nullptr && out_f32 == nullptr) || (res2 && !res)) {
    set_error("tc_conv1d: bad call (C=%d c_in=%d pool=%d n=%lld)", in.C, L.c_in, pool, (long long)in.n);
    return ORCA_B200_EINVAL;
  }
  TcKArgs a;
  a.in_hi = static_cast<const __nv_bfloat16*>(in.hi); a.in_lo = static_cast<const __nv_bfloat16*>(in.lo);
  a.w = static_cast<const uint8_t*>(fmt ? L.tc_w16 : L.tc_w); a.bias = L.b;
  a.res_hi = res ? static_cast<const __nv_bfloat16*>(res->hi) : nullptr;
  a.res_lo = res ? static_cast<const __nv_bfloat16*>(res->lo) : nullptr;
  a.res2_hi = res2 ? static_cast<const __nv_bfloat16*>(res2->hi) : nullptr;
  a.res2_lo = res2 ? static_cast<const __nv_bfloat16*>(res2->lo) : nullptr;
  a.out_hi = out_planes ? static_cast<__nv_bfloat16*>(out_planes->hi) : nullptr;
  a.out_lo = (out_planes && out_planes->fmt == 0) ? static_cast<__nv_bfloat16*>(out_planes->lo) : nullptr;  // NULL: one fp16 plane
  a.out_f32 = out_f32;
  a.sat = (out_planes && out_planes->fmt == 1) ? out_planes->sat : nullptr;
  a.nb = in.nb; a.n = (int)in.n; a.npad_in = (int)in.npad; a.n_out = (int)(in.n / pool);
  a.npad_out = out_planes ? (int)out_planes->npad : 0;
  a.pool = pool; a.relu = relu;
  a.tiles_per_sample = (int)((in.n + 127) / 128); a.total_tiles = a.tiles_per_sample * in.nb;
  if (res2 && (res2->C != L.c_out || res2->n != in.n || res2->npad != in.npad || res2->nb != in.nb)) { set_error("tc_conv1d: residual-2 geometry mismatch"); return ORCA_B200_EINVAL; }
  if (res && (res->C != L.c_out || res->n != in.n || res->npad != in.npad || res->nb != in.nb)) { set_error("tc_conv1d: residual geometry mismatch"); return ORCA_B200_EINVAL; }
  if (out_planes && (out_planes->C != L.c_out || out_planes->n != in.n / pool || out_planes->nb != in.nb)) { set_error("tc_conv1d: output geometry mismatch"); return ORCA_B200_EINVAL; }
  if (a.total_tiles <= 0) return ORCA_B200_OK;
  const int sms = sm_count();
  const int key = L.c_in * 1000 + L.c_out;
  switch (key) {
    case 64064: return launch_tc<64, 64>(a, sms, s, fmt);
    case 64096: return launch_tc<64, 96>(a, sms, s, fmt);
    case 96096: return launch_tc<96, 96>(a, sms, s, fmt);
    case 96128: return launch_tc<96, 128>(a, sms, s, fmt);
    case 128128: return launch_tc<128, 128>(a, sms, s, fmt);
    default: set_error("tc_conv1d: no kernel for %d->%d", L.c_in, L.c_out); return ORCA_B200_EUNSUPPORTED;
  }
}

int tc_conv_first(const ConvLayer& L, const SeqIn& in, int nb, int64_t Ltot, int64_t l_begin, int64_t n, TcAct* out,
                  cudaStream_t s) {
  if (L.c_in != 4 || L.c_out != 64 || out->C != 64 || out->n != n || out->nb != nb || out->fmt != 0) { set_error("tc_conv_first: bad geometry"); return ORCA_B200_EINVAL; }
  dim3 grid((unsigned)((n + 127) / 128), (unsigned)nb), block(256);
  conv_first_planes_kernel<<<grid, block, 0, s>>>(in, Ltot, l_begin, n, (int)out->npad, L.w, L.b,
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
  pool_planes_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(in.hi), in.fmt ? nullptr : static_cast<const __nv_bfloat16*>(in.lo),
                                          static_cast<__nv_bfloat16*>(out->hi), out->fmt ? nullptr : static_cast<__nv_bfloat16*>(out->lo), planes,
                                          (int)out->n, (int)in.npad, (int)out->npad, p);
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
