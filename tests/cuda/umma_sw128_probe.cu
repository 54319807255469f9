// Standalone probe for the next decoder layout: K-major SWIZZLE_128B shared-memory operands whose
// rows are 128 bytes (64 bf16 channels of one pixel), read through ROW-SHIFTED views (the conv taps) by advancing the
// descriptor start address by shift*128 B.  Question answered: does a shifted start need the descriptor's base_offset
// field ((addr >> 7) & 7), or is the XOR pattern taken from the absolute address bits?  Both variants are run.
// Operands are written by threads with the swizzle applied by hand (16-byte chunk index XOR (row & 7)), which is the
// pattern a SWIZZLE_128B TMA tensor copy produces for 128-byte rows in a 1024-byte aligned buffer.
// Result on B200 (round 1): every shift in {0,1,2,3,4,8,11,16} is exact with base_offset = 0; with base_offset =
// (addr >> 7) & 7 only the shifts that are multiples of 8 rows are.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_sw128_probe umma_sw128_probe.cu
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// K-major SWIZZLE_128B descriptor: SBO = 1024 B between 8-row groups, LBO unused, version 1, layout_type 2 (bits 61..63)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
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

constexpr int M = 128, ROWS_A = 144, K = 64, N = 64;

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ Ag /*[ROWS_A][K]*/,
                                                    const __nv_bfloat16* __restrict__ Bg /*[N][K]*/,
                                                    float* __restrict__ D /*[M][N]*/, int shift, int use_base_offset) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // ROWS_A x 128 B
  uint8_t* sB = smem + ROWS_A * 128;   // N x 128 B (ROWS_A * 128 = 18432 = 18 * 1024: still 1024-aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + N * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_mma = smem_u32(&bars[0]);
  if (tid == 0) {
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // operands: 16-byte chunk c of row r lands at r*128 + ((c ^ (r & 7)) * 16)
  for (int i = tid; i < ROWS_A * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(Ag + (size_t)r * K + c * 8);
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(Bg + (size_t)r * K + c * 8);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int k = 0; k < K / 16; ++k) {
      const uint32_t a_addr = smem_u32(sA) + shift * 128 + k * 32;
      const uint64_t ad = make_desc_sw128(a_addr, use_base_offset ? ((a_addr >> 7) & 7) : 0);
      const uint64_t bd = make_desc_sw128(smem_u32(sB) + k * 32, 0);
      umma_bf16(tmem, ad, bd, idesc, k > 0);
    }
    umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int run(int shift, int use_base_offset) {
  std::vector<float> A((size_t)ROWS_A * K), B((size_t)N * K);
  std::vector<__nv_bfloat16> Ap(A.size()), Bp(B.size());
  srand(99 + shift);
  for (size_t i = 0; i < A.size(); ++i) { Ap[i] = __float2bfloat16((float)(rand() % 2001 - 1000) / 1000.f); A[i] = __bfloat162float(Ap[i]); }
  for (size_t i = 0; i < B.size(); ++i) { Bp[i] = __float2bfloat16((float)(rand() % 2001 - 1000) / 1000.f); B[i] = __bfloat162float(Bp[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, Ap.size() * 2)); CK(cudaMalloc(&dB, Bp.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));
  const size_t smem = ROWS_A * 128 + N * 128 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, shift, use_base_offset);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe shift=%d base_offset=%d : CUDA error %s\n", shift, use_base_offset, cudaGetErrorString(e)); return 1; }
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
  printf("sw128 probe shift=%2d base_offset=%s : maxerr %.3e (max |ref| %.3f) %s\n", shift, use_base_offset ? "(addr>>7)&7" : "0", maxerr, maxref,
         ok ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return ok ? 0 : 1;
}

int main() {
  for (int bo = 0; bo < 2; ++bo)
    for (int shift : {0, 1, 2, 3, 4, 8, 11, 16}) run(shift, bo);
  return 0;
}
