// Device-side building blocks of the tcgen05 kernels: mbarrier, bulk copy (UBLKCP), UMMA descriptors,
// tcgen05.mma / commit / ld, and bf16 hi/lo split helpers.  (Inline PTX; no CUTLASS dependency.)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace orca {
namespace tcdev {

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
// SBO>>4 [32,46) = stride between 8-row groups (128 B: rows are 16 B apart), version=1 [46,48).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, M=128, N=n
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// same with A=B=fp16 (format code 0): the single-pass encoder stages (see conv_tc.cu, FMT = 1)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (x = hi + lo up to 2^-17 relative), one 16-byte store each.
// Packed conversions: cvt.rn.bf16x2.f32 handles two values per instruction.
__device__ __forceinline__ void split_pack2(float x0, float x1, uint32_t& h, uint32_t& l) {
  const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);  // .x = x0 (low half), .y = x1
  h = *reinterpret_cast<const uint32_t*>(&hh);
  const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xFFFF0000u);
  const __nv_bfloat162 ll = __floats2bfloat162_rn(r0, r1);
  l = *reinterpret_cast<const uint32_t*>(&ll);
}
__device__ __forceinline__ void split_store8(const float* v, __nv_bfloat16* hi_dst, __nv_bfloat16* lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_pack2(v[2 * j], v[2 * j + 1], h[j], l[j]);
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}
// fp16 single-plane variants: 8 fp32 -> 8 fp16 (round to nearest), one 16-byte store; v[0..8) += 8 fp16
__device__ __forceinline__ void store_h8(const float* v, void* dst) {
  uint32_t h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
}
// fp16 range guard of the single-pass stages: flag (never clamp) anything that would round to +-inf or sits within 10 %
// of the fp16 maximum (65504).  NaN cannot appear before an overflow has been flagged at its source layer.
template <int N>
__device__ __forceinline__ void guard_h(const float* v, unsigned int* sat) {
  if (!sat) return;
  float m = 0.f;
#pragma unroll
  for (int j = 0; j < N; ++j) m = fmaxf(m, fabsf(v[j]));
  if (!(m <= 60000.f)) atomicOr(sat, 1u);
}
__device__ __forceinline__ void add_h8(float* v, const void* src) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(src));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
    v[2 * j] += f.x;
    v[2 * j + 1] += f.y;
  }
}
// v[0..8) += hi + lo
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

}  // namespace tcdev
}  // namespace orca

namespace orca {
namespace tcdev {
// One elected lane of a fully-converged warp (the CUTLASS elect_one_sync idiom).  Role loops run
// warp-uniformly and only the issue instructions sit under this predicate, so ptxas keeps descriptors in
// uniform registers instead of wrapping every UTCHMMA / UBLKCP in a VOTEU/ELECT/BRA.U.ANY loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// descriptor halves: hi word is constant (SBO = 128 B, version 1); lo word = start>>4 | (LBO>>4)<<16,
// so advancing the start address by `bytes` is lo += bytes >> 4.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
constexpr uint32_t kUmmaDescHi = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint64_t umma_desc64(uint32_t lo) { return ((uint64_t)kUmmaDescHi << 32) | lo; }
}  // namespace tcdev
}  // namespace orca
