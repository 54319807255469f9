// Standalone probe of the tcgen05 building blocks used by conv_tc.cu (run on a B200):
//   * K-major, no-swizzle ("interleaved") shared-memory operand layout [k-chunk(16 B)][row][16 B]
//     described by LBO = rows*16 B, SBO = 128 B, so that a ROW-SHIFTED window of A (the 9 conv taps)
//     is just a start-address offset of shift*16 B;
//   * operands written by cp.async.bulk (UBLKCP) with mbarrier complete_tx, or by threads
//     (generic proxy) + fence.proxy.async;
//   * tcgen05.mma kind::f16 (bf16 x bf16 -> f32 in TMEM), tcgen05.commit -> mbarrier,
//     tcgen05.ld 32x32b epilogue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// K-major, SWIZZLE_NONE descriptor: LBO = byte stride between the two 16-byte K chunks of one MMA,
// SBO = byte stride between 8-row groups.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;                // base_offset 0, layout_type 0 = SWIZZLE_NONE
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

constexpr int M = 128, ROWS_A = 136;

// mode 0: threads write smem; mode 1: cp.async.bulk
template <int N, int K>
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ Ag /*[K/8][ROWS_A][8]*/,
                                                    const __nv_bfloat16* __restrict__ Bg /*[K/8][N][8]*/,
                                                    float* __restrict__ D /*[M][N]*/, int shift, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int KC = K / 8;
  constexpr uint32_t TCOLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : N <= 256 ? 256 : 512;  // power of two >= 32
  constexpr uint32_t A_BYTES = KC * ROWS_A * 16, B_BYTES = KC * N * 16;
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_load = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

  if (tid == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (mode == 0) {
    for (uint32_t i = tid; i < A_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(Ag)[i];
    for (uint32_t i = tid; i < B_BYTES / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Bg)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  } else {
    if (tid == 0) {
      mbar_expect_tx(bar_load, A_BYTES + B_BYTES);
      for (int c = 0; c < KC; ++c) bulk_g2s(smem_u32(sA) + c * ROWS_A * 16, Ag + (size_t)c * ROWS_A * 8, ROWS_A * 16, bar_load);
      bulk_g2s(smem_u32(sB), Bg, B_BYTES, bar_load);
    }
  }
  if (tid == 0) {
    if (mode == 1) mbar_wait(bar_load, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t ad = make_desc(smem_u32(sA) + (2 * k) * ROWS_A * 16 + shift * 16, ROWS_A * 16, 128);
      const uint64_t bd = make_desc(smem_u32(sB) + (2 * k) * N * 16, N * 16, 128);
      umma_bf16(tmem, ad, bd, idesc, k > 0);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes [32w, 32w+32), 32 columns at a time
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS));
}

template <int N, int K>
int run(int shift, int mode) {
  constexpr int KC = K / 8;
  std::vector<float> A((size_t)ROWS_A * K), B((size_t)N * K);
  std::vector<__nv_bfloat16> Ap((size_t)KC * ROWS_A * 8), Bp((size_t)KC * N * 8);
  srand(1234 + shift * 7 + N);
  for (auto& v : A) v = (float)(rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (float)(rand() % 2001 - 1000) / 1000.f;
  for (int r = 0; r < ROWS_A; ++r)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16(A[(size_t)r * K + k]);
      A[(size_t)r * K + k] = __bfloat162float(h);
      Ap[((size_t)(k / 8) * ROWS_A + r) * 8 + k % 8] = h;
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      __nv_bfloat16 h = __float2bfloat16(B[(size_t)n * K + k]);
      B[(size_t)n * K + k] = __bfloat162float(h);
      Bp[((size_t)(k / 8) * N + n) * 8 + k % 8] = h;
    }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, Ap.size() * 2)); CK(cudaMalloc(&dB, Bp.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));
  const size_t smem = (size_t)KC * ROWS_A * 16 + (size_t)KC * N * 16 + 64;
  CK(cudaFuncSetAttribute(probe_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<N, K><<<1, 128, smem>>>(dA, dB, dD, shift, mode);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D((size_t)M * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)(m + shift) * K + k] * B[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(ref - D[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  const bool ok = maxerr <= 1e-4 * maxref + 1e-5;
  printf("probe N=%d K=%d shift=%d mode=%s : maxerr %.3e (max |ref| %.3f) %s\n", N, K, shift, mode ? "bulk" : "threads", maxerr, maxref,
         ok ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return ok ? 0 : 1;
}

int main() {
  int fails = 0;
  for (int mode = 0; mode < 2; ++mode)
    for (int shift : {0, 1, 3, 4, 8}) {
      fails += run<64, 64>(shift, mode);
      fails += run<128, 64>(shift, mode);
    }
  fails += run<96, 32>(5, 1);
  fails += run<192, 64>(2, 1);
  fails += run<256, 128>(7, 1);
  printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
  return fails ? 1 : 0;
}
