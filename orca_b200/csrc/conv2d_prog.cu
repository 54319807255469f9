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
  int NA, NW, resident, a_slot_bytes, w_stage_bytes, share;  // share: sliding window of runs across chained tiles
};

struct Tc2dGeom {
  long long plane_rows;
  int nb, S, Wp, tiles_per_row, total_tiles, n_layers, oper_bytes;
  long long* trace;  // diagnostic (env ORCA_B200_DEC_TRACE): CTA 0 logs (event code, clock64) per role, 4 x 2048 entries
  int experiment;  // diagnostic (env ORCA_B200_DEC_EXPERIMENT; results are WRONG when non-zero): bit 0 epilogue without
                   // global loads/stores, bit 1 producer copies 16 B per run chunk, bit 2 no MMAs, bit 3 no TMEM loads
};

struct Bars {
  uint32_t a_full, a_empty, w_full, w_empty, acc_full, acc_empty;
};

// event trace of CTA 0 (one writer thread per role): entry = code << 48 | clock64
struct Trace {
  long long* p = nullptr;
  int n = 0;
  __device__ __forceinline__ void log(int code) {
    if (p && n < 2048) p[n++] = ((long long)code << 48) | (clock64() & 0xFFFFFFFFFFFFll);
  }
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

// ---- tile order and run sharing --------------------------------------------------------------------------
// A tile is 128 pixels of one image row y and reads the rows y-d, y, y+d ("runs").  Tiles are visited in CHAIN order:
// within one (image, column block) the rows are walked class by class of y mod d, so consecutive tiles are d rows
// apart and share two of their three runs.  Every CTA owns a contiguous piece of that sequence and keeps a sliding
// window of runs in shared memory: one new run per tile instead of three.  The layer is bound by the latency of the
// run-load -> MMA -> release loop times the runs per tile (tools/decoder_phases.py), so this is a 2-3x shorter
// critical path, and 2-3x less L2 traffic.
__device__ __forceinline__ void cta_range(const Tc2dGeom& g, int& t0, int& t1) {
  t0 = (int)((long long)g.total_tiles * blockIdx.x / gridDim.x);
  t1 = (int)((long long)g.total_tiles * (blockIdx.x + 1) / gridDim.x);
}

__device__ __forceinline__ void tile_coords(const Tc2dGeom& g, int d, int tile, int& b, int& tx, int& y, int& x0) {
  const int per_img = g.S * g.tiles_per_row;
  b = tile / per_img;
  const int rem = tile - b * per_img;
  tx = rem / g.S;
  const int p = rem - tx * g.S;
  const int q = g.S / d, m = g.S - q * d;  // m classes hold q+1 rows, the other d-m classes q rows
  int r, k;
  if (p < m * (q + 1)) {
    r = p / (q + 1);
    k = p - r * (q + 1);
  } else {
    const int pp = p - m * (q + 1);
    r = m + pp / q;
    k = pp - (pp / q) * q;
  }
  y = r + k * d;
  x0 = tx * 128;
  if (x0 + 128 > g.S) x0 = g.S > 128 ? g.S - 128 : 0;
}

// Walks a CTA's contiguous piece of the chain-ordered tile sequence without per-tile integer divisions (they sat on the
// MMA issuer's critical path: ~1700 cycles per tile for the two tile_coords calls of the first version).
struct TileCursor {
  int t, t_end, b, tx, r, y, x0;
  __device__ __forceinline__ void init(const Tc2dGeom& g, int d) {
    cta_range(g, t, t_end);
    if (t < t_end) {
      tile_coords(g, d, t, b, tx, y, x0);
      r = y % d;
    }
  }
  __device__ __forceinline__ bool done() const { return t >= t_end; }
  // the tile after this one in chain order (does not depend on the CTA range)
  __device__ __forceinline__ void peek(const Tc2dGeom& g, int d, int& nb, int& ntx, int& ny) const {
    nb = b; ntx = tx; ny = y + d;
    if (ny >= g.S) {
      int nr = r + 1;
      if (nr >= d || nr >= g.S) { nr = 0; ntx = tx + 1; if (ntx >= g.tiles_per_row) { ntx = 0; nb = b + 1; } }
      ny = nr;
    }
  }
  __device__ __forceinline__ void next(const Tc2dGeom& g, int d) {
    ++t;
    int nb, ntx, ny;
    peek(g, d, nb, ntx, ny);
    if (ny != y + d) r = ny;  // a new class starts at row r = its index
    b = nb; tx = ntx; y = ny;
    x0 = tx * 128;
    if (x0 + 128 > g.S) x0 = g.S > 128 ? g.S - 128 : 0;
  }
};

// Sliding window bookkeeping, evaluated identically by the producer and the MMA issuer.  Loads are numbered in issue
// order (`id`; id + kb for the K-blocks of one run) and live in ring slot id % NA; a tile that continues the chain of
// the previous one re-uses the loads of its rows y-d and y.
struct Window {
  uint32_t id_y = 0, id_yd = 0;
  int b = -1, tx = -1, y = 0;
  bool valid = false;  // the previous tile loaded its row y+d
};
__device__ __forceinline__ void window_step(Window& w, bool share, int nkb, int d, int S, int b, int tx, int y, uint32_t& a_it,
                                            uint32_t (&ids)[3], bool (&isnew)[3]) {
  const bool cont = share && w.valid && w.b == b && w.tx == tx && y == w.y + d;
#pragma unroll
  for (int i = 0; i < 3; ++i) { ids[i] = 0; isnew[i] = false; }
  if (y - d >= 0) {
    if (cont) ids[0] = w.id_y;
    else { ids[0] = a_it; a_it += nkb; isnew[0] = true; }
  }
  if (cont) ids[1] = w.id_yd;
  else { ids[1] = a_it; a_it += nkb; isnew[1] = true; }
  if (y + d < S) { ids[2] = a_it; a_it += nkb; isnew[2] = true; }
  w.valid = y + d < S; w.id_y = ids[1]; w.id_yd = ids[2]; w.b = b; w.tx = tx; w.y = y;
}

// ---- producer: bulk copies of A runs and weight stages for one layer ----------------------------------
// `preloaded`: the layer's resident weight stages were already requested at the previous layer boundary
__device__ __forceinline__ void producer_layer(const Tc2dLayer& L, const Tc2dGeom& g, uint32_t sA, uint32_t sW, const Bars& B,
                                               bool preloaded, Trace& tr) {
  const int nkb = (L.c_in + 63) / 64, R = 128 + 2 * L.d, kc = (L.c_in < 64 ? L.c_in : 64) / 8;
  const uint32_t aLoOff = (uint32_t)kc * R * 16, tapBytes = 2u * kc * L.c_out * 16;
  const int lane = threadIdx.x & 31;
  tr.log(1);
  const bool share = L.share != 0;
  if (L.resident && !preloaded && elect_one()) {  // first layer of the program: request every weight stage up front
    for (int sid = 0; sid < 3 * nkb; ++sid) {
      mbar_expect_tx(B.w_full + 8 * sid, 3 * tapBytes);
      bulk_g2s(sW + sid * L.w_stage_bytes, L.w + (size_t)sid * 3 * tapBytes, 3 * tapBytes, B.w_full + 8 * sid);
    }
  }
  __syncwarp();
  uint32_t a_it = 0, w_it = 0;
  Window win;
  TileCursor cur;
  for (cur.init(g, L.d); !cur.done(); cur.next(g, L.d)) {
    const int b = cur.b, tx = cur.tx, y = cur.y, x0 = cur.x0;
    uint32_t ids[3];
    bool isnew[3];
    window_step(win, share, nkb, L.d, g.S, b, tx, y, a_it, ids, isnew);
#pragma unroll 1
    for (int dyi = 0; dyi < 3; ++dyi) {
      const int yy = y + (dyi - 1) * L.d;
      if (yy < 0 || yy >= g.S) continue;
      const long long row0 = (long long)yy * g.Wp + x0 + kPX - L.d;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        if (!L.resident) {  // streamed weights: one stage per (run, K-block) use, whether or not the run itself is re-used
          const uint32_t ws = w_it % L.NW;
          mbar_wait(B.w_empty + 8 * ws, ((w_it / L.NW) & 1) ^ 1);
          if (elect_one()) {
            const int sid = dyi * nkb + kb;
            mbar_expect_tx(B.w_full + 8 * ws, 3 * tapBytes);
            bulk_g2s(sW + ws * L.w_stage_bytes, L.w + (size_t)sid * 3 * tapBytes, 3 * tapBytes, B.w_full + 8 * ws);
          }
          __syncwarp();
          ++w_it;
        }
        if (!isnew[dyi]) continue;  // still in the window from the previous tile
        const uint32_t id = ids[dyi] + kb, slot = id % L.NA;
        mbar_wait(B.a_empty + 8 * slot, ((id / L.NA) & 1) ^ 1);
        tr.log(2);
        // One bulk copy per (8-channel chunk, hi|lo), issued by 2*kc lanes at once.  (A 3-D TMA tensor copy with box
        // {16 B, R, kc} was measured SLOWER: its 16-byte inner extent starves the TMA engine.)
        const uint32_t run_bytes = (g.experiment & 2) ? 16u : (uint32_t)R * 16;
        if (elect_one()) mbar_expect_tx(B.a_full + 8 * slot, 2u * kc * run_bytes);
        __syncwarp();
        if (lane < 2 * kc) {
          const int c = lane >> 1, part = lane & 1;
          const uint32_t dst = sA + slot * L.a_slot_bytes + (part ? aLoOff : 0u) + c * R * 16;
          const long long plane = (long long)b * (L.c_in / 8) + kb * 8 + c;
          const long long off = (plane * g.plane_rows + row0) * 8;
          bulk_g2s(dst, (part ? L.in_lo : L.in_hi) + off, run_bytes, B.a_full + 8 * slot);
        }
        __syncwarp();
        tr.log(3);
      }
    }
  }
}

// ---- MMA issuer for one layer ---------------------------------------------------------------------------
template <int C_OUT, int KSTEPS>
__device__ __forceinline__ void mma_layer(const Tc2dLayer& L, const Tc2dGeom& g, uint32_t sA, uint32_t sW, const Bars& B,
                                          uint32_t tmem, uint32_t& acc_it, Trace& tr) {
  tr.log(10);
  const int nkb = (L.c_in + 63) / 64, R = 128 + 2 * L.d, kc = (L.c_in < 64 ? L.c_in : 64) / 8;
  const uint32_t aLoOff = (uint32_t)kc * R * 16, tapBytes = 2u * kc * C_OUT * 16;
  const uint32_t aStep = (uint32_t)(2 * R * 16) >> 4, aLoStep = aLoOff >> 4, tapStep = tapBytes >> 4;
  const bool share = L.share != 0;
  uint32_t a_it = 0, w_it = 0;
  if (L.resident)  // every stage was requested at the layer boundary (or by the producer for the first layer)
    for (int sid = 0; sid < 3 * nkb; ++sid) mbar_wait(B.w_full + 8 * sid, 0);
  Window win;
  TileCursor cur;
  for (cur.init(g, L.d); !cur.done(); cur.next(g, L.d)) {
    const int b = cur.b, tx = cur.tx, y = cur.y;
    uint32_t ids[3];
    bool isnew[3];
    window_step(win, share, nkb, L.d, g.S, b, tx, y, a_it, ids, isnew);
    // does the next tile of this CTA continue the chain (re-use rows y and y+d)?  In chain order that is exactly "y + d
    // is still inside the image", provided the tile belongs to this CTA.
    const bool cont_next = share && cur.t + 1 < cur.t_end && y + L.d < g.S;
    const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
    tr.log(11);
    mbar_wait(B.acc_empty + 8 * as, aph ^ 1);
    tc_fence_after();
    tr.log(12);
    const uint32_t d_tmem = tmem + as * 128;
    const int first_dyi = (y - L.d >= 0) ? 0 : 1, last_dyi = (y + L.d < g.S) ? 2 : 1;
#pragma unroll 1
    for (int dyi = 0; dyi < 3; ++dyi) {
      const int yy = y + (dyi - 1) * L.d;
      if (yy < 0 || yy >= g.S) continue;
      const bool release = dyi == 0 || !cont_next;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t id = ids[dyi] + kb, slot = id % L.NA;
        mbar_wait(B.a_full + 8 * slot, (id / L.NA) & 1);  // re-used runs: the phase completed earlier, returns at once
        tr.log(13);
        const int sid = dyi * nkb + kb;
        uint32_t ws = sid;
        if (!L.resident) {
          ws = w_it % L.NW;
          mbar_wait(B.w_full + 8 * ws, (w_it / L.NW) & 1);
        }
        tc_fence_after();
        const uint32_t aLo = __shfl_sync(0xffffffffu, umma_desc_lo(sA + slot * L.a_slot_bytes, R * 16), 0);
        const uint32_t bLo = __shfl_sync(0xffffffffu, umma_desc_lo(sW + ws * L.w_stage_bytes, 2 * C_OUT * 16), 0);
        if (elect_one()) {
          if (g.experiment & 4) {}
          else if (dyi == first_dyi) issue_stage<C_OUT, KSTEPS, true>(d_tmem, aLo, bLo, aStep, aLoStep, tapStep, (uint32_t)L.d, kb == 0);
          else issue_stage<C_OUT, KSTEPS, false>(d_tmem, aLo, bLo, aStep, aLoStep, tapStep, (uint32_t)L.d, false);
          if (!L.resident) umma_commit(B.w_empty + 8 * ws);
          if (release) umma_commit(B.a_empty + 8 * slot);
          if (dyi == last_dyi && kb == nkb - 1) umma_commit(B.acc_full + 8 * as);
        }
        __syncwarp();
        tr.log(14);
        if (!L.resident) ++w_it;
      }
    }
    ++acc_it;
  }
}

// ---- epilogue for one layer -----------------------------------------------------------------------------
template <int C_OUT>
__device__ __forceinline__ void epilogue_layer(const Tc2dLayer& L, const Tc2dGeom& g, const Bars& B, uint32_t tmem,
                                               const float* sBias, uint32_t& acc_it, int warp, int lane, Trace& tr) {
  tr.log(20);
  const int q = warp & 3, h = (warp - 2) >> 2;
  constexpr int UNITS = C_OUT / 16, MYU = UNITS / 2;
  TileCursor cur;
  for (cur.init(g, L.d); !cur.done(); cur.next(g, L.d)) {
    const int b = cur.b, y = cur.y, x0 = cur.x0;
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
    tr.log(21);
    mbar_wait(B.acc_full + 8 * as, aph);
    tc_fence_after();
    tr.log(22);
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
      if (lane == 0) mbar_arrive(B.acc_empty + 8 * as);  // one arrival per warp: 256 same-word atomics per tile serialise
    tr.log(23);
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
  Trace tr;
  if (g.trace && blockIdx.x == 0 && (tid == 0 || tid == 32 || tid == 64)) tr.p = g.trace + (tid >> 5) * 2048;
  Trace trb;  // layer boundary events (thread 0 doubles as the producer's logger: separate region)
  if (g.trace && blockIdx.x == 0 && tid == 0) trb.p = g.trace + 3 * 2048;

#pragma unroll 1
  for (int l = 0; l < g.n_layers; ++l) {
    const Tc2dLayer L = layers[l];
    const uint32_t sA = smem_u32(smem), sW = sA + (uint32_t)L.NA * L.a_slot_bytes;
    const int variant = L.c_out == 32 ? 0 : (L.c_in == 32 ? 2 : 1);  // (32,4) (64,4) (64,2)
    if (warp == 0) {
      producer_layer(L, g, sA, sW, B, l > 0, tr);
    } else if (warp == 1) {
      if (variant == 0) mma_layer<32, 4>(L, g, sA, sW, B, tmem, acc_it, tr);
      else if (variant == 1) mma_layer<64, 4>(L, g, sA, sW, B, tmem, acc_it, tr);
      else mma_layer<64, 2>(L, g, sA, sW, B, tmem, acc_it, tr);
    } else {
      if (variant == 0) epilogue_layer<32>(L, g, B, tmem, sBias, acc_it, warp, lane, tr);
      else epilogue_layer<64>(L, g, B, tmem, sBias, acc_it, warp, lane, tr);
    }
    if (l + 1 == g.n_layers) break;
    // ---- layer boundary: this CTA is done (all its MMAs completed: the epilogue saw every accumulator) ----
    __syncthreads();
    trb.log(30);
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
      trb.log(31);
      const unsigned int target = (unsigned int)(l + 1) * gridDim.x;
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      } while (seen < target);
      trb.log(32);
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
static long long* g_trace_buf = nullptr;
long long* tc2d_trace_buffer() { return g_trace_buf; }

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
  {
    const char* e = getenv("ORCA_B200_DEC_SHARE");  // 0 disables the sliding window (A/B comparison)
    t.share = (nkb == 1 && t.NA >= 3 && !(e && atoi(e) == 0)) ? 1 : 0;
  }
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
    g.trace = nullptr;
    static long long* trace_buf = nullptr;  // diagnostic only: one process-wide buffer, dumped by orca_b200_debug_trace
    if (getenv("ORCA_B200_DEC_TRACE")) {
      if (!trace_buf) ORCA_CUDA_OK(cudaMalloc(&trace_buf, 4 * 2048 * sizeof(long long)));
      ORCA_CUDA_OK(cudaMemsetAsync(trace_buf, 0, 4 * 2048 * sizeof(long long), s));
      g.trace = trace_buf;
      g_trace_buf = trace_buf;
    }
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
