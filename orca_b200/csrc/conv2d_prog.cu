// Decoder "program" kernel: ALL dilated 3x3 convolutions of one Decoder / Decoder_1m call (118 / 76 layers)
// in ONE persistent tcgen05 kernel with a grid-wide barrier between layers.
//
// Why: a single 250x250 conv is only ~3.4 tiles per SM; launched one kernel per layer, ~10 us of every
// ~19 us launch is fixed latency (launch, TMEM allocation, barrier setup, first loads, drain) -- measured with
// tools/decoder_scan.py.  The layers of a decoder are strictly sequential (dilated taps read the whole previous
// map), so the latency chain is also what bounds multi-GPU scaling.  Here the CTAs stay resident (1 per SM,
// cooperative launch), keep their TMEM allocation and mbarriers, and only pay a software grid barrier
// (atomic counter in global memory) per layer.
//
// Measured in this round (tools/decoder_phases.py and an event trace of CTA 0): a layer is
// paced by the MMA issuer -- the 64->32 layers run their MMAs at ~140 cycles per (concat, lo) pair against a 48-cycle
// tensor floor because every tap re-reads the 128-row A operand from shared memory, mostly at row offsets that are
// not multiples of 8 rows (two 128-byte wavefronts per core matrix) -- not by the A-run loads: a 3-D TMA tensor copy
// per run (slower: 16-byte inner extent), copies issued by 16 lanes at once, and a chain-ordered tile walk with a
// sliding window of runs (one new run per tile instead of three) each left the layer time unchanged or worse.
//
// Per layer the kernel re-carves shared memory (A-run slots / weight stages sized for that layer's c_in, c_out and
// dilation, exactly like conv2d_tc_kernel) and re-initialises the A/W ring barriers; the two TMEM accumulator
// barriers run across layers.  The per-layer math (row-run A operands, [Bh;Bl] concat MMAs, epilogue) is the
// single-layer kernel's (conv2d_tc.cu).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"
#include "tc.h"
#include "tc_device.cuh"

namespace orca {
namespace {
using namespace tcdev;

constexpr int kPX = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreadsP = 64 + 32 * kEpiWarps;
constexpr int kMaxNA = 4, kMaxNW = 6;

struct Tc2dLayer {
  const __nv_bfloat16* in_hi; const __nv_bfloat16* in_lo;
  const uint8_t* w;
  const float* bias;
  const __nv_bfloat16* res_hi; const __nv_bfloat16* res_lo;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int c_in, c_out, d, relu;
  int NA, NW, resident, a_slot_bytes, w_stage_bytes, pad_;
};

struct Tc2dGeom {
  long long plane_rows;
  int nb, S, Wp, tiles_per_row, total_tiles, n_layers, oper_bytes;
  int experiment;  // diagnostic (env ORCA_B200_DEC_EXPERIMENT; results are WRONG when non-zero): bit 0 epilogue without
                   // global loads/stores, bit 1 producer copies 16 B per run chunk, bit 2 no MMAs, bit 3 no TMEM loads
};

struct Bars {
  uint32_t a_full, a_empty, w_full, w_empty, acc_full, acc_empty;
};

template <int C_OUT, int KSTEPS, bool FIRST>
__device__ __forceinline__ void issue_stage(uint32_t d_tmem, uint32_t aLo, uint32_t bLo, uint32_t aStep, uint32_t aLoStep,
                                            uint32_t tapStep, uint32_t d, bool first_kb) {
  constexpr uint32_t idesc = umma_idesc_bf16(C_OUT), idesc_cat = umma_idesc_bf16(2 * C_OUT);
  constexpr uint32_t bStep = (2u * 2 * C_OUT * 16) >> 4;
#pragma unroll
  for (int dxi = 0; dxi < 3; ++dxi) {
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const uint32_t ao = aLo + ks * aStep + dxi * d;
      const uint32_t bo = bLo + dxi * tapStep + ks * bStep;
      const uint32_t accum = (FIRST && dxi == 0 && ks == 0) ? (first_kb ? 0u : 1u) : 1u;
      umma_bf16(d_tmem, umma_desc64(ao), umma_desc64(bo), idesc_cat, accum);
      umma_bf16(d_tmem, umma_desc64(ao + aLoStep), umma_desc64(bo), idesc, 1u);
    }
  }
}

__device__ __forceinline__ void tile_coords(const Tc2dGeom& g, int tile, int& b, int& y, int& x0) {
  const int tiles_per_img = g.S * g.tiles_per_row;
  b = tile / tiles_per_img;
  const int rem = tile - b * tiles_per_img;
  y = rem / g.tiles_per_row;
  const int tx = rem - y * g.tiles_per_row;
  x0 = tx * 128;
  if (x0 + 128 > g.S) x0 = g.S > 128 ? g.S - 128 : 0;
}

// ---- producer: bulk copies of A runs and weight stages for one layer ----------------------------------
// `preloaded`: the layer's resident weight stages were already requested at the previous layer boundary
__device__ __forceinline__ void producer_layer(const Tc2dLayer& L, const Tc2dGeom& g, uint32_t sA, uint32_t sW, const Bars& B,
                                               bool preloaded) {
  const int nkb = (L.c_in + 63) / 64, R = 128 + 2 * L.d, kc = (L.c_in < 64 ? L.c_in : 64) / 8;
  const uint32_t aLoOff = (uint32_t)kc * R * 16, tapBytes = 2u * kc * L.c_out * 16;
  uint32_t a_it = 0, w_it = 0, loaded = (preloaded && L.resident) ? 0xFFFFFFFFu : 0u;
  const int lane = threadIdx.x & 31;
  for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
    int b, y, x0;
    tile_coords(g, tile, b, y, x0);
#pragma unroll 1
    for (int si = 0; si < 3; ++si) {
      const int dyi = si == 0 ? 1 : (si == 1 ? 0 : 2);  // centre row first (always inside the image)
      const int yy = y + (dyi - 1) * L.d;
      if (yy < 0 || yy >= g.S) continue;
      const long long row0 = (long long)yy * g.Wp + x0 + kPX - L.d;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t slot = a_it % L.NA, ph = (a_it / L.NA) & 1;
        mbar_wait(B.a_empty + 8 * slot, ph ^ 1);
        const int sid = dyi * nkb + kb;
        const bool need_w = L.resident ? !((loaded >> sid) & 1u) : true;
        const uint32_t ws = L.resident ? (uint32_t)sid : w_it % L.NW;
        if (!L.resident) mbar_wait(B.w_empty + 8 * ws, ((w_it / L.NW) & 1) ^ 1);
        // One bulk copy per (8-channel chunk, hi|lo) = up to 16 per run.  Issued by one lane they cost ~65 cycles
        // each back to back and the producer warp paced the whole layer (tools/decoder_phases.py: 1.65 us per tile
        // with every other role switched off); issued by 16 lanes of the warp at once they overlap.  A 3-D TMA tensor
        // copy (box {16 B, R, kc}) was measured SLOWER than either: its 16-byte inner extent starves the TMA engine.
        const uint32_t run_bytes = (g.experiment & 2) ? 16u : (uint32_t)R * 16;
        if (elect_one()) {
          mbar_expect_tx(B.a_full + 8 * slot, 2u * kc * run_bytes);
          if (need_w) {
            mbar_expect_tx(B.w_full + 8 * ws, 3 * tapBytes);
            bulk_g2s(sW + ws * L.w_stage_bytes, L.w + (size_t)sid * 3 * tapBytes, 3 * tapBytes, B.w_full + 8 * ws);
          }
        }
        __syncwarp();
        if (lane < 2 * kc) {
          const int c = lane >> 1, part = lane & 1;
          const uint32_t dst = sA + slot * L.a_slot_bytes + (part ? aLoOff : 0u) + c * R * 16;
          const long long plane = (long long)b * (L.c_in / 8) + kb * 8 + c;
          const long long off = (plane * g.plane_rows + row0) * 8;
          bulk_g2s(dst, (part ? L.in_lo : L.in_hi) + off, run_bytes, B.a_full + 8 * slot);
        }
        __syncwarp();
        ++a_it;
        if (L.resident) loaded |= 1u << sid; else ++w_it;
      }
    }
  }
}

// ---- MMA issuer for one layer ---------------------------------------------------------------------------
template <int C_OUT, int KSTEPS>
__device__ __forceinline__ void mma_layer(const Tc2dLayer& L, const Tc2dGeom& g, uint32_t sA, uint32_t sW, const Bars& B,
                                          uint32_t tmem, uint32_t& acc_it, bool preloaded) {
  const int nkb = (L.c_in + 63) / 64, R = 128 + 2 * L.d, kc = (L.c_in < 64 ? L.c_in : 64) / 8;
  const uint32_t aLoOff = (uint32_t)kc * R * 16, tapBytes = 2u * kc * C_OUT * 16;
  const uint32_t aStep = (uint32_t)(2 * R * 16) >> 4, aLoStep = aLoOff >> 4, tapStep = tapBytes >> 4;
  uint32_t a_it = 0, w_it = 0, waited = 0;
  if (preloaded && L.resident) {  // every preloaded stage must have landed before this layer's barriers are re-initialised
    for (int sid = 0; sid < 3 * nkb; ++sid) mbar_wait(B.w_full + 8 * sid, 0);
    waited = 0xFFFFFFFFu;
  }
  for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
    int b, y, x0;
    tile_coords(g, tile, b, y, x0);
    const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
    mbar_wait(B.acc_empty + 8 * as, aph ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = tmem + as * 128;
    const int last_dyi = (y + L.d < g.S) ? 2 : ((y - L.d >= 0) ? 0 : 1);
#pragma unroll 1
    for (int si = 0; si < 3; ++si) {
      const int dyi = si == 0 ? 1 : (si == 1 ? 0 : 2);
      const int yy = y + (dyi - 1) * L.d;
      if (yy < 0 || yy >= g.S) continue;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t slot = a_it % L.NA;
        mbar_wait(B.a_full + 8 * slot, (a_it / L.NA) & 1);
        const int sid = dyi * nkb + kb;
        uint32_t ws;
        if (L.resident) {
          ws = sid;
          if (!((waited >> sid) & 1u)) { waited |= 1u << sid; mbar_wait(B.w_full + 8 * ws, 0); }
        } else {
          ws = w_it % L.NW;
          mbar_wait(B.w_full + 8 * ws, (w_it / L.NW) & 1);
        }
        tc_fence_after();
        const uint32_t aLo = __shfl_sync(0xffffffffu, umma_desc_lo(sA + slot * L.a_slot_bytes, R * 16), 0);
        const uint32_t bLo = __shfl_sync(0xffffffffu, umma_desc_lo(sW + ws * L.w_stage_bytes, 2 * C_OUT * 16), 0);
        if (elect_one()) {
          if (g.experiment & 4) {}
          else if (si == 0) issue_stage<C_OUT, KSTEPS, true>(d_tmem, aLo, bLo, aStep, aLoStep, tapStep, (uint32_t)L.d, kb == 0);
          else issue_stage<C_OUT, KSTEPS, false>(d_tmem, aLo, bLo, aStep, aLoStep, tapStep, (uint32_t)L.d, false);
          if (!L.resident) umma_commit(B.w_empty + 8 * ws);
          umma_commit(B.a_empty + 8 * slot);
          if (dyi == last_dyi && kb == nkb - 1) umma_commit(B.acc_full + 8 * as);
        }
        __syncwarp();
        if (!L.resident) ++w_it;
        ++a_it;
      }
    }
    ++acc_it;
  }
}

// ---- epilogue for one layer -----------------------------------------------------------------------------
template <int C_OUT>
__device__ __forceinline__ void epilogue_layer(const Tc2dLayer& L, const Tc2dGeom& g, const Bars& B, uint32_t tmem,
                                               const float* sBias, uint32_t& acc_it, int warp, int lane) {
  const int q = warp & 3, h = (warp - 2) >> 2;
  constexpr int UNITS = C_OUT / 16, MYU = UNITS / 2;
  for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x) {
    int b, y, x0;
    tile_coords(g, tile, b, y, x0);
    const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
    const int x = x0 + q * 32 + lane;
    const bool valid = x < g.S && !(g.experiment & 1);
    const long long r = (long long)y * g.Wp + kPX + x;
    float res[MYU][16];
#pragma unroll
    for (int u = 0; u < MYU; ++u) {
#pragma unroll
      for (int j = 0; j < 16; ++j) res[u][j] = 0.f;
      if (L.res_hi && valid) {
        const int c0 = 16 * (h + 2 * u);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const long long off = (((long long)b * (C_OUT / 8) + (c0 >> 3) + ch) * g.plane_rows + r) * 8;
          // plain loads: the residual was written by OTHER CTAs earlier in this same kernel (no read-only path)
          const uint4 hh = *reinterpret_cast<const uint4*>(L.res_hi + off), ll = *reinterpret_cast<const uint4*>(L.res_lo + off);
          const uint32_t hw[4] = {hh.x, hh.y, hh.z, hh.w}, lw[4] = {ll.x, ll.y, ll.z, ll.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            res[u][8 * ch + 2 * j] = __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
            res[u][8 * ch + 2 * j + 1] = __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
          }
        }
      }
    }
    mbar_wait(B.acc_full + 8 * as, aph);
    tc_fence_after();
#pragma unroll
    for (int u = 0; u < MYU; ++u) {
      const int c0 = 16 * (h + 2 * u);
      uint32_t raw[16], raw2[16];
      if (!(g.experiment & 8)) {
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + as * 128 + c0, raw);
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + as * 128 + C_OUT + c0, raw2);
      }
      if (valid) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float t = __uint_as_float(raw[j]) + __uint_as_float(raw2[j]) + sBias[c0 + j];
          v[j] = (L.relu ? fmaxf(t, 0.f) : t) + res[u][j];
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const long long off = (((long long)b * (C_OUT / 8) + (c0 >> 3) + ch) * g.plane_rows + r) * 8;
          split_store8(v + 8 * ch, L.out_hi + off, L.out_lo + off);
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(B.acc_empty + 8 * as);  // one arrival per warp
    ++acc_it;
  }
}

__global__ void __launch_bounds__(kThreadsP, 1) conv2d_program_kernel(const Tc2dLayer* __restrict__ layers, const Tc2dGeom g,
                                                                      unsigned int* counter) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* sBias = reinterpret_cast<float*>(smem + g.oper_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 64);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxNA + 2 * kMaxNW + 4);
  Bars B;
  B.a_full = smem_u32(bars); B.a_empty = B.a_full + 8 * kMaxNA;
  B.w_full = B.a_empty + 8 * kMaxNA; B.w_empty = B.w_full + 8 * kMaxNW;
  B.acc_full = B.w_empty + 8 * kMaxNW; B.acc_empty = B.acc_full + 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  auto init_rings = [&]() {
    for (int i = 0; i < kMaxNA; ++i) { mbar_init(B.a_full + 8 * i, 1); mbar_init(B.a_empty + 8 * i, 1); }
    for (int i = 0; i < kMaxNW; ++i) { mbar_init(B.w_full + 8 * i, 1); mbar_init(B.w_empty + 8 * i, 1); }
  };
  if (tid == 0) {
    init_rings();
    for (int i = 0; i < 2; ++i) { mbar_init(B.acc_full + 8 * i, 1); mbar_init(B.acc_empty + 8 * i, kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid >= 64 && tid < 128) sBias[tid - 64] = (tid - 64 < layers[0].c_out) ? layers[0].bias[tid - 64] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  uint32_t acc_it = 0;  // runs across layers (same tile sequence in the MMA and epilogue roles)

#pragma unroll 1
  for (int l = 0; l < g.n_layers; ++l) {
    const Tc2dLayer L = layers[l];
    const uint32_t sA = smem_u32(smem), sW = sA + (uint32_t)L.NA * L.a_slot_bytes;
    const int variant = L.c_out == 32 ? 0 : (L.c_in == 32 ? 2 : 1);  // (32,4) (64,4) (64,2)
    if (warp == 0) {
      producer_layer(L, g, sA, sW, B, l > 0);
    } else if (warp == 1) {
      if (variant == 0) mma_layer<32, 4>(L, g, sA, sW, B, tmem, acc_it, l > 0);
      else if (variant == 1) mma_layer<64, 4>(L, g, sA, sW, B, tmem, acc_it, l > 0);
      else mma_layer<64, 2>(L, g, sA, sW, B, tmem, acc_it, l > 0);
    } else {
      if (variant == 0) epilogue_layer<32>(L, g, B, tmem, sBias, acc_it, warp, lane);
      else epilogue_layer<64>(L, g, B, tmem, sBias, acc_it, warp, lane);
    }
    if (l + 1 == g.n_layers) break;
    // ---- layer boundary: this CTA is done (all its MMAs completed: the epilogue saw every accumulator) ----
    __syncthreads();
    if (tid == 0) {
      // publish this CTA's output pixels: bar.sync ordered the other warps' stores before this point, and a release
      // at gpu scope is cumulative over them (cheaper than a full __threadfence + relaxed atomic)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      init_rings();                          // A/W rings restart from phase 0 for the next layer's carve
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      {
        // This CTA's shared memory is idle now (all its MMAs have completed), and weights do not depend on the
        // other CTAs: fetch the next layer's resident weight stages while waiting at the grid barrier.
        const Tc2dLayer& N = layers[l + 1];
        if (N.resident) {
          const int nkbN = (N.c_in + 63) / 64, kcN = (N.c_in < 64 ? N.c_in : 64) / 8;
          const uint32_t stageN = 3u * 2u * kcN * N.c_out * 16;
          const uint32_t sWN = smem_u32(smem) + (uint32_t)N.NA * N.a_slot_bytes;
          for (int sid = 0; sid < 3 * nkbN; ++sid) {
            mbar_expect_tx(B.w_full + 8 * sid, stageN);
            bulk_g2s(sWN + sid * N.w_stage_bytes, N.w + (size_t)sid * stageN, stageN, B.w_full + 8 * sid);
          }
        }
      }
      const unsigned int target = (unsigned int)(l + 1) * gridDim.x;
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      } while (seen < target);
      asm volatile("fence.proxy.async;" ::: "memory");  // later bulk copies (async proxy) read other CTAs' stores
    }
    if (tid >= 64 && tid < 128) {
      const Tc2dLayer& N = layers[l + 1];
      sBias[tid - 64] = (tid - 64 < N.c_out) ? N.bias[tid - 64] : 0.f;
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct Tc2dProgram::Impl {
  std::vector<Tc2dLayer> layers;
  Tc2dGeom g{};
  int oper_bytes = 0;
  double flop = 0.0;
};

Tc2dProgram::Tc2dProgram() : impl(new Impl) {}
Tc2dProgram::~Tc2dProgram() { delete impl; }
int Tc2dProgram::size() const { return (int)impl->layers.size(); }
double Tc2dProgram::flop() const { return impl->flop; }

int Tc2dProgram::add(const ConvLayer& L, const TcMap& in, const TcMap* res, TcMap* out, int relu) {
  if (!L.tc_w || !tc_layer2d_eligible(L)) { set_error("Tc2dProgram: layer %d->%d has no tensor-core weights", L.c_in, L.c_out); return ORCA_B200_EUNSUPPORTED; }
  if (in.C != L.c_in || out->C != L.c_out || out->S != in.S || out->nb != in.nb || (res && (res->C != L.c_out || res->S != in.S))) {
    set_error("Tc2dProgram: geometry mismatch");
    return ORCA_B200_EINVAL;
  }
  Tc2dGeom& g = impl->g;
  if (impl->layers.empty()) {
    g.plane_rows = in.plane_rows; g.nb = in.nb; g.S = in.S; g.Wp = in.Wp;
    g.tiles_per_row = (in.S + 127) / 128; g.total_tiles = in.nb * in.S * g.tiles_per_row;
  } else if (g.S != in.S || g.nb != in.nb) {
    set_error("Tc2dProgram: all layers must share the map geometry");
    return ORCA_B200_EINVAL;
  }
  Tc2dLayer t{};
  t.in_hi = static_cast<const __nv_bfloat16*>(in.hi); t.in_lo = static_cast<const __nv_bfloat16*>(in.lo);
  t.w = static_cast<const uint8_t*>(L.tc_w); t.bias = L.b;
  t.res_hi = res ? static_cast<const __nv_bfloat16*>(res->hi) : nullptr;
  t.res_lo = res ? static_cast<const __nv_bfloat16*>(res->lo) : nullptr;
  t.out_hi = static_cast<__nv_bfloat16*>(out->hi); t.out_lo = static_cast<__nv_bfloat16*>(out->lo);
  t.c_in = L.c_in; t.c_out = L.c_out; t.d = L.dil; t.relu = relu;
  const int nkb = (L.c_in + 63) / 64, kc = (L.c_in < 64 ? L.c_in : 64) / 8, R = 128 + 2 * L.dil;
  t.a_slot_bytes = 2 * kc * R * 16;
  t.w_stage_bytes = 3 * 2 * kc * L.c_out * 16;
  const int n_stages = 3 * nkb;
  const int limit = 227 * 1024 - 2048;
  if (n_stages * t.w_stage_bytes + 2 * t.a_slot_bytes <= limit && n_stages <= kMaxNW) {
    t.resident = 1; t.NW = n_stages;
  } else {
    t.resident = 0; t.NW = 2;
    if (2 * t.w_stage_bytes + t.a_slot_bytes > limit) t.NW = 1;
  }
  int na = (limit - t.NW * t.w_stage_bytes) / t.a_slot_bytes;
  if (na < 1) { set_error("Tc2dProgram: shared memory budget exceeded"); return ORCA_B200_EUNSUPPORTED; }
  t.NA = na > kMaxNA ? kMaxNA : na;
  const int oper = t.NA * t.a_slot_bytes + t.NW * t.w_stage_bytes;
  if (oper > impl->oper_bytes) impl->oper_bytes = oper;
  impl->flop += 2.0 * in.nb * in.S * in.S * (double)L.c_in * L.c_out * 9;
  impl->layers.push_back(t);
  return ORCA_B200_OK;
}

size_t Tc2dProgram::scratch_bytes(int max_layers) { return (size_t)max_layers * sizeof(Tc2dLayer) + 256; }

int Tc2dProgram::run(void* scratch, size_t scratch_bytes_, cudaStream_t s) {
  const int n = (int)impl->layers.size();
  if (n == 0) return ORCA_B200_OK;
  if (scratch_bytes_ < scratch_bytes(n)) { set_error("Tc2dProgram: scratch too small"); return ORCA_B200_EWORKSPACE; }
  Tc2dGeom g = impl->g;
  g.n_layers = n;
  g.oper_bytes = (impl->oper_bytes + 127) & ~127;
  {
    const char* e = getenv("ORCA_B200_DEC_EXPERIMENT");
    g.experiment = e ? atoi(e) : 0;
  }
  unsigned int* counter = static_cast<unsigned int*>(scratch);
  Tc2dLayer* d_layers = reinterpret_cast<Tc2dLayer*>(static_cast<char*>(scratch) + 256);
  ORCA_CUDA_OK(cudaMemsetAsync(counter, 0, 256, s));
  ORCA_CUDA_OK(cudaMemcpyAsync(d_layers, impl->layers.data(), (size_t)n * sizeof(Tc2dLayer), cudaMemcpyHostToDevice, s));
  const int smem = g.oper_bytes + 64 * 4 + (2 * kMaxNA + 2 * kMaxNW + 4) * 8 + 16 + 128;
  static bool configured_dev[32] = {};  // cudaFuncSetAttribute is per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 31];
  static int sms = 148;
  if (!configured) {
    ORCA_CUDA_OK(cudaFuncSetAttribute(conv2d_program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  const int grid = g.total_tiles < sms ? g.total_tiles : sms;
  const Tc2dLayer* lp = d_layers;
  void* args[] = {(void*)&lp, (void*)&g, (void*)&counter};
  // cooperative launch: the software grid barrier needs every CTA resident (1 per SM)
  ORCA_CUDA_OK(cudaLaunchCooperativeKernel((const void*)conv2d_program_kernel, dim3(grid), dim3(kThreadsP), args, (size_t)smem, s));
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
