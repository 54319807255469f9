// Standalone micro-benchmark (one CTA): issue rate of back-to-back tcgen05.mma (M = 128, K = 16, bf16) as a function of
// the shared-memory operand layout (K-major SWIZZLE_NONE [k-chunk][row][16 B] vs K-major SWIZZLE_128B rows of 128 B), of
// a row shift of the A view (the conv taps) and of N.  Answers: are the N = 32 / 64 MMAs of the decoder / stage-1 encoder
// limited by the shared-memory operand port, and does a shifted (mis-aligned) A view cost extra?
// Results on B200 (round 1, one CTA, zero operands):
//  * tile pattern (36 MMAs per elect block, one commit per tile, accumulators alternating per tile):
//      N =  64: 53.6 cycles per MMA (1929 per tile; tensor floor 32)     N = 128: 68.2 (2457 per tile; floor 64)
//    which fits  T_mma = max(tensor floor, shared-memory operand bytes / 128 B per clock)  with operand bytes =
//    (128 + N) rows x 32 B per K = 16 step: N = 64 -> 48 cycles (operand bound, 67 % tensor utilisation at best),
//    N = 128 -> 64 cycles (balanced), plus ~150-200 cycles per tile of issue overhead.  36 distinct vs 4 repeated weight
//    images: no difference.
//  * flat loop (4 MMAs per elect block): ~105 / 113 / 130 / 171 cycles per K-step for N = 32 / 64 / 128 / 256, the same
//    for both layouts, for row shifts 0 / 1 / 8 and for 1 / 2 / 4 independent accumulators: an elect_one() + __syncwarp()
//    issue block costs ~360 cycles when it is not hidden behind MMA execution.  (Consequences for the decoder program
//    kernel, which issues one 24-MMA block per run: DESIGN.md section 9.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate_probe umma_rate_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc_none(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
// one elected lane of a converged warp: keeps the issue loop warp-uniform so that ptxas emits back-to-back UTCHMMA
// with uniform-register operands (a divergent `if (tid == 0)` wraps every MMA in an ELECT / R2UR / BRA.U.ANY loop that
// costs ~110 cycles per MMA and hides what this probe wants to measure)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

constexpr int ROWS_A = 160, K = 64;

// layout 0: SWIZZLE_NONE, 1: SWIZZLE_128B.  Per iteration: K/16 = 4 MMAs of width N (+ 4 of width N2 if N2 > 0, the "lo" product)
__global__ void __launch_bounds__(128) rate_kernel(int layout, int shift, int N, int N2, int iters, long long* out, int nacc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                    // 160 rows x 128 B = 20 KB either layout
  uint8_t* sB = smem + ROWS_A * 128;     // up to 256 rows x 128 B = 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 256 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = smem_u32(&bars[0]);
  for (int i = tid; i < (ROWS_A * 128 + 256 * 128) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((N2 > 0 ? N2 : 8) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint64_t ad[4], bd[4];
    for (int k = 0; k < 4; ++k) {
      if (layout == 0) {
        ad[k] = desc_none(smem_u32(sA) + 2 * k * ROWS_A * 16 + shift * 16, ROWS_A * 16);
        bd[k] = desc_none(smem_u32(sB) + 2 * k * 256 * 16, 256 * 16);
      } else {
        ad[k] = desc_sw128(smem_u32(sA) + shift * 128 + k * 32);
        bd[k] = desc_sw128(smem_u32(sB) + k * 32);
      }
    }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // nacc > 1: successive K-steps accumulate into different TMEM column blocks (independent chains)
          umma_bf16(tmem + (uint32_t)((k % nacc) * 64), ad[k], bd[k], idesc);
          if (N2 > 0) umma_bf16(tmem + 256 + (uint32_t)((k % nacc) * 64), ad[k], bd[k], idesc2);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    if (tid == 0) out[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// The conv1d 64->64 tile pattern: 36 MMAs per tile = 9 row-shifted views (taps) x 4 K-steps of A against 36 different
// 2 KB weight images, two accumulators used alternately, one commit per tile (nobody waits on it until the end).
__global__ void __launch_bounds__(128) tile_pattern_kernel(int N, int tiles, int distinct_b, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                         // [8 chunks][160 rows][16 B]
  uint8_t* sB = smem + 8 * ROWS_A * 16;       // 36 x [2 chunks][N rows][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 36 * 2 * 128 * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar = smem_u32(&bars[0]), bar_tile = smem_u32(&bars[1]);
  for (int i = tid; i < (8 * ROWS_A * 16 + 36 * 2 * 128 * 16) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_tile, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const long long t0 = clock64();
    for (int tile = 0; tile < tiles; ++tile) {
      if (elect_one()) {
        const uint32_t d = tmem + (uint32_t)((tile & 1) * 128);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = desc_none(a0 + 2 * k * ROWS_A * 16 + tap * 16, ROWS_A * 16);
            const uint64_t bd = desc_none(b0 + (distinct_b ? (tap * 4 + k) : k) * (2 * 128 * 16), 128 * 16);
            umma_bf16(d, ad, bd, idesc);
          }
        umma_commit(bar_tile);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    if (tid == 0) out[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main() {
  long long* d;
  CK(cudaMalloc(&d, 8));
  {
    const size_t smem2 = 8 * ROWS_A * 16 + 36 * 2 * 128 * 16 + 64 + 1024;
    CK(cudaFuncSetAttribute(tile_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    for (int N : {64, 128})
      for (int distinct_b : {1, 0}) {
        tile_pattern_kernel<<<1, 128, smem2>>>(N, 4, distinct_b, d);
        tile_pattern_kernel<<<1, 128, smem2>>>(N, 400, distinct_b, d);
        CK(cudaDeviceSynchronize());
        long long cyc;
        CK(cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost));
        printf("tile pattern N=%3d %s weight images : %6.1f cycles per MMA (%7.0f per 36-MMA tile; tensor floor %4.0f per MMA)\n", N,
               distinct_b ? "36 distinct" : "4 repeated ", (double)cyc / (400.0 * 36), (double)cyc / 400.0, 128.0 * N / 256.0);
      }
  }
  const size_t smem = ROWS_A * 128 + 256 * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 2000;
  const int cfgs[][2] = {{32, 0}, {64, 0}, {128, 0}, {256, 0}, {64, 32}, {128, 64}};
  for (int layout = 0; layout < 2; ++layout)
    for (auto& c : cfgs)
      for (int nacc : {1, 2, 4}) {
        if (nacc > 1 && c[0] > 64) continue;  // the independent chains use 64-column blocks
        const int shift = 1;
        rate_kernel<<<1, 128, smem>>>(layout, shift, c[0], c[1], 10, d, nacc);  // warm-up
        rate_kernel<<<1, 128, smem>>>(layout, shift, c[0], c[1], iters, d, nacc);
        CK(cudaDeviceSynchronize());
        long long cyc;
        CK(cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost));
        const double per = (double)cyc / (iters * 4.0);  // per K-step (one MMA, or one N + N2 pair)
        const double floor_cyc = 128.0 * c[0] / 256.0 + (c[1] ? 128.0 * c[1] / 256.0 : 0.0);
        printf("layout=%s N=%3d%s independent accumulators=%d : %6.1f cycles per K-step (tensor floor %5.1f) -> %4.0f%% of floor\n",
               layout ? "SW128" : "NONE ", c[0], c[1] ? (c[1] == 32 ? "+32" : "+64") : "   ", nacc, per, floor_cyc, 100.0 * floor_cyc / per);
      }
  return 0;
}
