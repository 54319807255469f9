// Decoder stream kernel: ALL dilated 3x3 convolutions of one Decoder / Decoder_1m call (orca_modules.py:461-488,
// :782-800; 116-120 / 77 layers) in ONE persistent tcgen05 kernel.  Layout and host API: dec_stream.h.
//
// What bounds a decoder on B200 is not tensor FLOP but (a) the shared-memory operand port -- an M = 128 MMA with
// N = 32 / 64 reads (128 + N) x 32 B per K = 16 step -- (b) L2 -> SM bandwidth when every 128-pixel tile re-reads
// three (128 + 2d)-pixel runs of the previous map, and (c) the fixed cost of 118 dependent layers (fill, drain, grid
// barrier).  The schedule below attacks (b) and (c):
//
//  * Input-stationary chains.  Output rows y, y + d, y + 2d, ... of one 128-pixel column ("chain") share their input
//    runs: the run of row rho is the +d tap of tile rho - d, the centre tap of tile rho and the -d tap of tile rho + d.
//    A CTA owns a contiguous piece of the chain-major tile order, loads every run ONCE (one TMA box {64 ch, 128 + 2d px},
//    SWIZZLE_128B; the dx taps are descriptor offsets) and issues its up to 3 x 3 x KS x 2 MMAs into up to three live
//    TMEM accumulators (4 x 128 columns).  A-run traffic drops ~3x and there is ONE issue block per run.
//  * No grid barrier.  A tile's completion is published per image row (flags[layer][image][row], release/acquire at
//    gpu scope); a run is loaded as soon as ITS row of the previous layer is complete, so CTAs drift across layer
//    boundaries instead of waiting for the slowest one, and the partial last wave of a layer overlaps the next layer.
//    Buffers are recycled two or more layers later; the store warp checks a per-layer completion counter before its
//    first store of a layer (write-after-read), which in steady state never waits.
//  * TMA both ways.  Epilogue warps read the residual tile from / write the output tile to a swizzled shared-memory
//    staging tile; a dedicated warp issues the tensor-map stores (clipped at the image edge by the hardware), waits for
//    them and publishes the row flag.  Maps carry no padding pixels (out-of-bounds zero fill), so the four hot buffers
//    of a batch-2 decoder (2 x 32 MB + 2 x 16 MB) fit the 126 MB L2.
//
// Roles per CTA (384 threads, 1 CTA / SM, cooperative launch so that every CTA is resident): warp 0 = TMA producer
// (flag polling, runs, weights, residual tiles), warp 1 = MMA issuer, warps 2-9 = epilogue, warps 10-11 = store + publish.
#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

#include "common.h"
#include "dec_stream.h"
#include "tc_device.cuh"

namespace orca {
namespace {
using namespace tcdev;

constexpr int kEpiWarps = 8;
constexpr int kStoreWarps = 2;
constexpr int kThreadsS = 32 * (2 + kEpiWarps + kStoreWarps);
constexpr int kMaxNA = 4;
constexpr int kAcc = 4;              // TMEM accumulator slots of 128 columns
constexpr int kSmemCarve = 230400;   // 225 KB of operand / staging space (1024-aligned base; barriers + bias above it)
constexpr int kSmemTotal = kSmemCarve + 1024 + 896;  // + 64 B of static shared memory (diagnostics) = 227 KB

struct SLayer {
  const CUtensorMap* tm_in;   // box {64 ch, R px, 1 row} on the input buffer
  const CUtensorMap* tm_res;  // box {64, 128, 1} on the residual buffer (nullptr: none)
  const CUtensorMap* tm_out;  // box {64, 128, 1} on the output buffer
  const uint8_t* w;
  const float* bias;          // nullptr: no bias (first K half of a split 128-channel conv)
  int c_in, c_out, d, relu;
  int in_c0, in_c1;           // inner (channel) coordinates of the hi and lo input boxes (c_in = 32: one box, in_c1 unused)
  int in_layer, res_layer, war_layer;
  int NA, NS, drain_before, rot, split;
  int offW, offA, offStg0, offStg1;
  int a_box_bytes, a_slot_bytes, w_bytes, R;
};

struct SGeom {
  unsigned int* flags;  // [n_layers][nb][S] completed tiles per image row
  unsigned int* done;   // [n_layers] CTAs that finished the layer
  const int2* parts;    // [grid][nb] (first tile, tile count) of every CTA in every image's chain-major tile order
  int nb, S, tpr, total_tiles, n_layers;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned int* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// Bounded waits: a protocol bug must surface as an error, never as a hung GPU.  Production: trap after ~2 s.  With the
// debug switch on (ds_debug_enable, tests/cuda/dec_stream_test.cu) the first wait that times out records
// {1, source line, block, warp, layer-ish tag} in g_ds_debug and every later wait returns at once, so the kernel ends
// and the host can read where it stalled.
__device__ unsigned int g_ds_debug[4 + 3 * 60];  // {abort flag, -, -, -, (line << 8 | layer, count, one block) x 60}
__device__ unsigned int g_ds_debug_on;
__shared__ int s_role_layer[16];                  // layer each warp is working on (diagnostics only)
constexpr long long kSpinLimit = 4000000000LL;
__device__ __noinline__ bool wait_timed_out(long long& t0, int line) {
  const long long now = clock64();
  if (t0 == 0) { t0 = now; return false; }
  if (!g_ds_debug_on) {
    if (now - t0 > kSpinLimit) __trap();
    return false;
  }
  const bool aborted = *(volatile unsigned int*)&g_ds_debug[0] != 0;
  const long long limit = kSpinLimit / 8;
  if (!aborted && now - t0 <= limit) return false;
  if (now - t0 > limit / 2) {  // this wait was stuck too: count it under (source line, layer)
    const unsigned int key = ((unsigned)line << 8) | (unsigned)(s_role_layer[threadIdx.x >> 5] & 255);
    for (int i = 0; i < 60; ++i) {
      const unsigned int old = atomicCAS(&g_ds_debug[4 + 3 * i], 0u, key);
      if (old == 0u || old == key) { atomicAdd(&g_ds_debug[5 + 3 * i], 1u); g_ds_debug[6 + 3 * i] = blockIdx.x; break; }
    }
  }
  atomicExch(&g_ds_debug[0], 1u);
  return true;
}
__device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity, int line) {
  uint32_t done;
  long long t0 = 0;
  uint32_t spins = 0;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if ((++spins & 255u) == 0 && wait_timed_out(t0, line)) return;
  }
}
__device__ __forceinline__ void spin_until_b(const unsigned int* p, unsigned int need, int line) {
  long long t0 = 0;
  uint32_t spins = 0;
  while (ld_acquire(p) < need) {
    __nanosleep(64);
    if ((++spins & 255u) == 0 && wait_timed_out(t0, line)) return;
  }
}
#define spin_until(p, need) spin_until_b(p, need, __LINE__)
__device__ __forceinline__ void smem_wait_ge_b(const volatile uint32_t* p, uint32_t need, int line) {
  long long t0 = 0;
  uint32_t spins = 0;
  while ((int32_t)(*p - need) < 0) {
    __nanosleep(32);
    if ((++spins & 255u) == 0 && wait_timed_out(t0, line)) return;
  }
}
#define smem_wait_ge(p, need) smem_wait_ge_b(p, need, __LINE__)
#define mbar_wait(bar, parity) mbar_wait_b(bar, parity, __LINE__)
// K-major SWIZZLE_128B descriptor (rows of 128 B, 8-row groups 1024 B apart; base_offset 0: the XOR pattern follows
// the absolute shared-memory address bits, so a start address advanced by whole rows is a row-shifted view)
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t desc_sw128(uint32_t lo) { return ((uint64_t)kDescHiSw128 << 32) | (lo & 0x3FFFu); }

// ---- the CTA's share of one layer, in chain-major tile order -------------------------------------------------
// Linear tile index = (image * tpr + tx) * S + p, p enumerating the rows of one 128-pixel column chain by chain:
// residue r = y mod d first, then k = y div d.  A segment = consecutive tiles of one chain, rows r + k*d, k0 <= k < k1.
struct Seg { int b, tx, r, k0, k1, len; };
struct Walk {
  int S, d, q, rem, tpr, cur, end, n_tiles;
  int cur2, end2;  // second piece of a rotated range (single image only)
  int part, nb, img_tiles;
  bool whole;  // one contiguous piece of the whole tile order (SLayer::split == 0)
  const int2* parts;
  // A CTA owns one contiguous piece of EVERY image (SGeom::parts) and walks them image by image.  With two or more
  // images the tiles of image b in layer l+1 depend on what this and the other CTAs finished half a layer ago (they have
  // all been working on the other images since), so the store -> flag -> load latency of the dependent chain and the
  // all-to-all dependency at a change of dilation are hidden behind useful work instead of stalling every layer.
  // With a single image the piece is walked from `rot` instead ([t0 + rot, t1) then [t0, t0 + rot), rot advancing by 2
  // per layer of equal dilation): the first tiles of a layer then depend on tiles the previous layer finished early.
  __device__ __forceinline__ void load_part(int rot) {
    const int2 pr = parts[part];
    const int t0 = part * img_tiles + pr.x, t1 = t0 + pr.y;
    const int sft = (nb == 1 && pr.y > 0) ? rot % pr.y : 0;
    cur = t0 + sft; end = t1;
    cur2 = t0; end2 = t0 + sft;
  }
  // split = 1: per-image pieces (used for the first layer of a new dilation, whose dependencies are all-to-all);
  // split = 0: ONE contiguous piece of the whole order, walked from `rot` (two halo runs per layer instead of 2 x nb)
  __device__ __forceinline__ void init(const SGeom& g, int d_, int rot, int split) {
    S = g.S; d = d_; tpr = g.tpr; nb = g.nb; img_tiles = g.tpr * g.S;
    q = S / d; rem = S - q * d;
    parts = g.parts + (size_t)blockIdx.x * g.nb;
    whole = nb > 1 && !split;
    rot_ = rot;
    if (whole) {
      const int t0 = (int)(((long long)blockIdx.x * g.total_tiles) / gridDim.x);
      const int t1 = (int)(((long long)(blockIdx.x + 1) * g.total_tiles) / gridDim.x);
      n_tiles = t1 - t0;
      const int sft = n_tiles > 0 ? rot % n_tiles : 0;
      cur = t0 + sft; end = t1;
      cur2 = t0; end2 = t0 + sft;
      part = nb;  // no further pieces
      return;
    }
    n_tiles = 0;
    for (int b = 0; b < nb; ++b) n_tiles += parts[b].y;
    part = 0;
    load_part(rot);
  }
  int rot_;
  __device__ __forceinline__ bool next(Seg& s) {
    while (cur >= end) {
      if (cur2 < end2) { cur = cur2; end = end2; cur2 = end2; break; }
      if (++part >= nb) return false;
      load_part(rot_);
    }
    const int img = cur / S, p = cur - img * S;
    s.b = img / tpr; s.tx = img - s.b * tpr;
    int k;
    if (p < rem * (q + 1)) { s.r = p / (q + 1); k = p - s.r * (q + 1); s.len = q + 1; }
    else { const int pp = p - rem * (q + 1); const int rr = pp / q; s.r = rem + rr; k = pp - rr * q; s.len = q; }
    const int left = end - cur, n = left < s.len - k ? left : s.len - k;
    s.k0 = k; s.k1 = k + n;
    cur += n;
    return true;
  }
};
// runs a segment loads: rows k0-1 .. k1 of the chain, clipped to the chain
__device__ __forceinline__ int run_first(const Seg& s) { return s.k0 > 0 ? s.k0 - 1 : 0; }
__device__ __forceinline__ int run_last(const Seg& s) { return s.k1 < s.len ? s.k1 : s.len - 1; }
// last run that contributes to tile kt
__device__ __forceinline__ int tile_last_run(const Seg& s, int kt) { return kt + 1 < s.len ? kt + 1 : kt; }

struct Bars {
  uint32_t a_full, a_empty, w_full, mma_done, acc_full, acc_empty, res_full, stg_free, stg_full;
  // Plain shared-memory counters for the two waits whose waiter (the producer) can run MANY phases ahead of the
  // signaller (the store warp), where an mbarrier parity wait would alias: stores completed per staging slot
  // (cnt[0], cnt[1]) and layers whose stores have all drained (cnt[2]).
  volatile uint32_t* cnt;
};

// staging slot bookkeeping, identical in every role: slot = tile parity when the layer has two slots
// (cnt = uses of the slot so far -> phase of its stg_free / stg_full barriers; rcnt = uses by layers WITH a residual ->
// phase of its res_full barrier, which only those layers complete)
struct Stg {
  uint32_t cnt[2], rcnt[2];
  __device__ __forceinline__ void take(int NS, uint32_t tile_it, bool has_res, uint32_t& s, uint32_t& use, uint32_t& ruse) {
    s = NS == 2 ? (tile_it & 1u) : 0u;
    use = cnt[s]++;
    ruse = rcnt[s];
    if (has_res) ++rcnt[s];
  }
  // a role that does not touch the staging tiles in this layer still counts them
  __device__ __forceinline__ void skip(int NS, uint32_t n_tiles) {
    if (NS == 2) { cnt[0] += (n_tiles + 1) >> 1; cnt[1] += n_tiles >> 1; }
    else cnt[0] += n_tiles;
  }
};

// Optional event trace of one CTA (compiled in with -DDS_TRACE, tests/cuda/dec_stream_test.cu "trace"): per role and
// per run / tile {id, t0, t1, t2} in clock64 ticks -- where the time of a layer goes, measured, not modelled.
#ifdef DS_TRACE
__device__ unsigned long long* g_ds_trace;  // [4 roles][kTraceEv][4]
__device__ int g_ds_trace_block = DS_TRACE;
constexpr int kTraceEv = 8192;
#define TR(role, idx, slot, val)                                                                     \
  do {                                                                                               \
    if (blockIdx.x == g_ds_trace_block && g_ds_trace && (idx) < kTraceEv)                                    \
      g_ds_trace[((size_t)(role) * kTraceEv + (idx)) * 4 + (slot)] = (unsigned long long)(val);      \
  } while (0)
#define TR_NOW() clock64()
#else
#define TR(role, idx, slot, val) do { } while (0)
#define TR_NOW() 0
#endif

// A-run slot ring.  The number of slots changes from layer to layer, so the phase of a slot's barriers is tracked by a
// per-slot use counter (identical in the producer and the MMA issuer), and every layer starts again at slot 0.
struct ARing {
  uint32_t cnt[kMaxNA];
  uint32_t pos, total;
  __device__ __forceinline__ void begin_layer() { pos = 0; }
  __device__ __forceinline__ void take(int NA, uint32_t& slot, uint32_t& use) {
    slot = pos;
    pos = pos + 1 == (uint32_t)NA ? 0u : pos + 1;
    use = cnt[slot]++;
    ++total;
  }
};

// ---- producer ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void producer_layer(const SLayer& L, int l, const SGeom& g, uint32_t smem0, const Bars& B, ARing& ring,
                                               Stg& stg, int lane) {
  ring.begin_layer();
  // the previous layer's MMAs have all completed (its weights and A slots are free) ...
  if (l > 0) mbar_wait(B.mma_done, (uint32_t)(l - 1) & 1u);
  // ... and, when this layer's carve overlaps the previous layer's staging tiles, so have its stores
  if (L.drain_before) smem_wait_ge(B.cnt + 2, (uint32_t)l);
  if (elect_one()) {
    mbar_expect_tx(B.w_full, (uint32_t)L.w_bytes);
    const uint32_t third = (uint32_t)L.w_bytes / 3;  // one bulk copy per dy row of taps
    for (int i = 0; i < 3; ++i) bulk_g2s(smem0 + L.offW + i * third, L.w + (size_t)i * third, third, B.w_full);
  }
  __syncwarp();
  const unsigned int* in_flags = L.in_layer >= 0 ? g.flags + (size_t)L.in_layer * g.nb * g.S : nullptr;
  const unsigned int* res_flags = (L.tm_res && L.res_layer >= 0) ? g.flags + (size_t)L.res_layer * g.nb * g.S : nullptr;
  const uint32_t slot_tx = (uint32_t)(L.c_in == 64 ? 2 : 1) * (uint32_t)L.R * 128u;
  const uint32_t stg_tx = (uint32_t)L.c_out * 4u * 128u;  // 128 pixels x (hi + lo) x c_out x 2 B
  Walk w;
  w.init(g, L.d, L.rot, L.split);
  if (!L.tm_res) stg.skip(L.NS, (uint32_t)w.n_tiles);
  Seg s;
  uint32_t tile_it = 0;
  while (w.next(s)) {
    const int k_first = run_first(s), k_last = run_last(s), n_runs = k_last - k_first + 1;
    const int x0 = s.tx * 128;
    const int row_base = s.b * g.S + s.r + k_first * L.d;  // row of run j = row_base + j * d
    // Row flags are polled for a WINDOW of up to 32 upcoming runs at once (lane t <-> run wb + t: one L2 round trip for the
    // whole window) and remembered as bit masks; a run whose bit is still clear when its turn comes is waited for, then the
    // window is polled again.  A poll per run in series (~1 us each) made the producer the slowest role of the CTA.
    uint32_t in_ok = 0, res_ok = 0;
    int wb = 0;
    auto poll_window = [&]() {
      const int j = wb + lane;
      bool oi = true, orr = true;
      if (j < n_runs) {
        if (in_flags) oi = ld_acquire(in_flags + row_base + j * L.d) >= (unsigned)g.tpr;
        if (res_flags) orr = ld_acquire(res_flags + row_base + j * L.d) >= (unsigned)g.tpr;
      }
      in_ok = __ballot_sync(0xffffffffu, oi);
      res_ok = __ballot_sync(0xffffffffu, orr);
      fence_async_all();  // the TMA loads that follow (async proxy) must observe what the acquires made visible
    };
    if (in_flags || res_flags) poll_window();
    else in_ok = res_ok = 0xffffffffu;
#pragma unroll 1
    for (int kk = k_first; kk <= k_last; ++kk) {
      const int row = s.b * g.S + s.r + kk * L.d;
      const int j = kk - k_first;
      // residual tile that goes with this run: tile kk-1 (its last run is this one), plus tile kk when the chain ends here
      const int rt0 = (kk - 1 >= s.k0 && kk - 1 < s.k1) ? kk - 1 : -1;
      const int rt1 = (kk == s.len - 1 && kk >= s.k0 && kk < s.k1) ? kk : -1;
      TR(0, ring.total, 0, ((unsigned long long)l << 32) | (unsigned)kk);
      TR(0, ring.total, 1, TR_NOW());
      if (in_flags || res_flags) {
        if (j - wb >= 32) { wb = j; poll_window(); }
        const bool need_in = in_flags && !((in_ok >> (j - wb)) & 1u);
        // (a residual row one run back can only precede the window right after the window moved: re-check it then)
        const bool need_r0 = res_flags && rt0 >= 0 && (j - 1 < wb || !((res_ok >> (j - 1 - wb)) & 1u));
        const bool need_r1 = res_flags && rt1 >= 0 && !((res_ok >> (j - wb)) & 1u);
        if (need_in || need_r0 || need_r1) {
          const unsigned int* f = nullptr;
          if (lane == 0 && need_in) f = in_flags + row;
          if (lane == 1 && need_r0) f = res_flags + row - L.d;
          if (lane == 2 && need_r1) f = res_flags + row;
          if (f) spin_until(f, (unsigned)g.tpr);
          __syncwarp();
          poll_window();  // refresh: later rows have usually completed meanwhile
        }
      }
      TR(0, ring.total, 2, TR_NOW());
      uint32_t slot, ause;
      ring.take(L.NA, slot, ause);
      mbar_wait(B.a_empty + 8 * slot, (ause & 1u) ^ 1u);
      TR(0, ring.total - 1, 3, TR_NOW());
      if (elect_one()) {
        mbar_expect_tx(B.a_full + 8 * slot, slot_tx);
        const uint32_t dst = smem0 + L.offA + slot * L.a_slot_bytes;
        tma_load_3d(dst, L.tm_in, L.in_c0, x0 - L.d, row, B.a_full + 8 * slot);
        if (L.c_in == 64) tma_load_3d(dst + L.a_box_bytes, L.tm_in, L.in_c1, x0 - L.d, row, B.a_full + 8 * slot);
      }
      __syncwarp();
      if (L.tm_res) {
#pragma unroll 1
        for (int e = 0; e < 2; ++e) {
          const int kt = e == 0 ? rt0 : rt1;
          if (kt < 0) continue;
          uint32_t ss, use, ruse;
          stg.take(L.NS, tile_it, true, ss, use, ruse);
          ++tile_it;
          smem_wait_ge(B.cnt + ss, use);  // every earlier store out of this staging tile has been read
          if (elect_one()) {
            mbar_expect_tx(B.res_full + 8 * ss, stg_tx);
            const uint32_t dst = smem0 + (ss ? L.offStg1 : L.offStg0);
            const int trow = s.b * g.S + s.r + kt * L.d;
            tma_load_3d(dst, L.tm_res, 0, x0, trow, B.res_full + 8 * ss);
            if (L.c_out == 64) tma_load_3d(dst + 16384, L.tm_res, 64, x0, trow, B.res_full + 8 * ss);
          }
          __syncwarp();
        }
      }
    }
  }
}

// ---- MMA issuer ---------------------------------------------------------------------------------------------
// all MMAs of one (run, dy): 3 dx taps x KS K steps x { A_hi x [Bh;Bl] (N = 2*C_OUT), A_lo x Bh (N = C_OUT) }
template <int C_IN, int C_OUT>
__device__ __forceinline__ void issue_dy(uint32_t d_tmem, uint32_t aHi, uint32_t aLo, uint32_t bTap, uint32_t dshift, bool fresh) {
  constexpr int KS = C_IN / 16;
  constexpr uint32_t idesc = umma_idesc_bf16(C_OUT), idesc_cat = umma_idesc_bf16(2 * C_OUT);
  constexpr uint32_t tapStep = (2u * (C_IN / 8) * C_OUT * 16) >> 4;
  constexpr uint32_t bStep = (2u * 2 * C_OUT * 16) >> 4;
#pragma unroll
  for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t ao = dx * dshift + ks * 2;  // dx tap = dx*d rows of 128 B; K step = 32 B inside the swizzled row
      const uint32_t bo = bTap + dx * tapStep + ks * bStep;
      umma_bf16(d_tmem, desc_sw128(aHi + ao), umma_desc64(bo), idesc_cat, (fresh && dx == 0 && ks == 0) ? 0u : 1u);
      umma_bf16(d_tmem, desc_sw128(aLo + ao), umma_desc64(bo), idesc, 1u);
    }
  }
}

template <int C_IN, int C_OUT>
__device__ __forceinline__ void mma_layer(const SLayer& L, int l, const SGeom& g, uint32_t smem0, const Bars& B, uint32_t tmem,
                                          ARing& ring, uint32_t& acc_it) {
  ring.begin_layer();
  constexpr uint32_t tapStep = (2u * (C_IN / 8) * C_OUT * 16) >> 4;
  mbar_wait(B.w_full, (uint32_t)l & 1u);
  tc_fence_after();
  const uint32_t bBase = umma_desc_lo(smem0 + L.offW, 2 * C_OUT * 16);
  const uint32_t dshift = ((uint32_t)L.d * 128u) >> 4;
  Walk w;
  w.init(g, L.d, L.rot, L.split);
  Seg s;
  while (w.next(s)) {
    const int k_first = run_first(s), k_last = run_last(s);
    const uint32_t acc_base = acc_it;  // accumulator sequence number of tile k0
#pragma unroll 1
    for (int kk = k_first; kk <= k_last; ++kk) {
      uint32_t slot, ause;
      ring.take(L.NA, slot, ause);
      TR(1, ring.total - 1, 0, ((unsigned long long)l << 32) | (unsigned)kk);
      TR(1, ring.total - 1, 1, TR_NOW());
      // tiles this run contributes to: kt = kk+1 (dy = -d, image row above the output), kk, kk-1
      bool val[3], fresh[3], last[3];
      uint32_t acc[3];
#pragma unroll
      for (int dyi = 0; dyi < 3; ++dyi) {
        const int kt = kk + 1 - dyi;
        val[dyi] = kt >= s.k0 && kt < s.k1;
        const int first_run = kt > 0 ? kt - 1 : 0;
        fresh[dyi] = val[dyi] && kk == first_run;
        last[dyi] = val[dyi] && kk == tile_last_run(s, kt);
        acc[dyi] = acc_base + (uint32_t)(kt - s.k0);
        if (fresh[dyi]) {  // first contribution: the accumulator slot must have been drained by the epilogue
          mbar_wait(B.acc_empty + 8 * (acc[dyi] % kAcc), ((acc[dyi] / kAcc) & 1u) ^ 1u);
        }
      }
      TR(1, ring.total - 1, 2, TR_NOW());
      mbar_wait(B.a_full + 8 * slot, ause & 1u);
      tc_fence_after();
      TR(1, ring.total - 1, 3, TR_NOW());
      const uint32_t aBase = smem0 + L.offA + slot * L.a_slot_bytes;
      const uint32_t aHi = __shfl_sync(0xffffffffu, aBase >> 4, 0);
      const uint32_t aLo = __shfl_sync(0xffffffffu, (aBase + (C_IN == 64 ? (uint32_t)L.a_box_bytes : 64u)) >> 4, 0);
      if (elect_one()) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const int dyi = 2 - o;  // oldest tile first: it completes with this run
          if (!val[dyi]) continue;
          const uint32_t d_tmem = tmem + (acc[dyi] % kAcc) * 128u;
          issue_dy<C_IN, C_OUT>(d_tmem, aHi, aLo, bBase + dyi * 3 * tapStep, dshift, fresh[dyi]);
          if (last[dyi]) umma_commit(B.acc_full + 8 * (acc[dyi] % kAcc));
        }
        umma_commit(B.a_empty + 8 * slot);
      }
      __syncwarp();
    }
    acc_it += (uint32_t)(s.k1 - s.k0);
  }
  if (elect_one()) umma_commit(B.mma_done);  // every MMA of this layer (hence every read of its weights / runs) has completed
  __syncwarp();
}

// ---- epilogue -------------------------------------------------------------------------------------------------
template <int C_OUT>
__device__ __forceinline__ void epilogue_layer(const SLayer& L, const SGeom& g, uint32_t smem0, uint8_t* smem_gen, const Bars& B,
                                               uint32_t tmem, const float* sBias, uint32_t& acc_it, Stg& stg, int warp, int lane) {
  const int q = warp & 3, h = (warp - 2) >> 2;
  constexpr int NCH = C_OUT / 2;           // channels per warp
  constexpr int NCHUNK = NCH / 8;          // 16-byte chunks per plane per thread
  const int row = q * 32 + lane;           // pixel within the tile = TMEM lane
  const uint32_t sw = (uint32_t)(row & 7);
  Walk w;
  w.init(g, L.d, L.rot, L.split);
  Seg s;
  uint32_t tile_it = 0;
  while (w.next(s)) {
#pragma unroll 1
    for (int kt = s.k0; kt < s.k1; ++kt) {
      const uint32_t as = acc_it % kAcc, aph = (acc_it / kAcc) & 1u;
      uint32_t ss, use, ruse;
      stg.take(L.NS, tile_it, L.tm_res != nullptr, ss, use, ruse);
      ++tile_it;
      if (warp == 2) { TR(2, acc_it, 0, ((unsigned long long)L.c_out << 32) | (unsigned)kt); TR(2, acc_it, 1, TR_NOW()); }
      mbar_wait(B.acc_full + 8 * as, aph);
      tc_fence_after();
      if (warp == 2) TR(2, acc_it, 2, TR_NOW());
      float v[NCH];
      {
        uint32_t raw[NCH], raw2[NCH];
        const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + as * 128u + (uint32_t)(h * NCH);
        if constexpr (NCH == 32) { tmem_ld32(t0, raw); tmem_ld32(t0 + C_OUT, raw2); }
        else { tmem_ld16(t0, raw); tmem_ld16(t0 + C_OUT, raw2); }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.acc_empty + 8 * as);  // accumulator drained into registers
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const float t = __uint_as_float(raw[j]) + __uint_as_float(raw2[j]) + sBias[h * NCH + j];
          v[j] = L.relu ? fmaxf(t, 0.f) : t;
        }
      }
      // staging tile: residual in (TMA-loaded), output out, in place.  Row = pixel, 128 B per box row, chunk c of a row
      // at ((c ^ (row & 7)) << 4).  C_OUT = 64: hi box then lo box (16 KB each); C_OUT = 32: one box [hi 32 | lo 32].
      uint8_t* stile = smem_gen + (ss ? L.offStg1 : L.offStg0);
      uint8_t* hi_row = stile + row * 128;
      uint8_t* lo_row = C_OUT == 64 ? stile + 16384 + row * 128 : hi_row;
      const uint32_t hi_c0 = (uint32_t)(h * NCHUNK), lo_c0 = (C_OUT == 64 ? 0u : 4u) + (uint32_t)(h * NCHUNK);
      if (L.tm_res) mbar_wait(B.res_full + 8 * ss, ruse & 1u);
      else mbar_wait(B.stg_free + 8 * ss, (use & 1u) ^ 1u);
      if (L.tm_res) {
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          const uint4 hh = *reinterpret_cast<const uint4*>(hi_row + (((hi_c0 + c) ^ sw) << 4));
          const uint4 ll = *reinterpret_cast<const uint4*>(lo_row + (((lo_c0 + c) ^ sw) << 4));
          const uint32_t hw[4] = {hh.x, hh.y, hh.z, hh.w}, lw[4] = {ll.x, ll.y, ll.z, ll.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[8 * c + 2 * j] += __uint_as_float(hw[j] << 16) + __uint_as_float(lw[j] << 16);
            v[8 * c + 2 * j + 1] += __uint_as_float(hw[j] & 0xFFFF0000u) + __uint_as_float(lw[j] & 0xFFFF0000u);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c)
        split_store8(v + 8 * c, reinterpret_cast<__nv_bfloat16*>(hi_row + (((hi_c0 + c) ^ sw) << 4)),
                     reinterpret_cast<__nv_bfloat16*>(lo_row + (((lo_c0 + c) ^ sw) << 4)));
      fence_async_smem();  // generic-proxy writes -> visible to the TMA store
      __syncwarp();
      if (warp == 2) TR(2, acc_it, 3, TR_NOW());
      if (lane == 0) mbar_arrive(B.stg_full + 8 * ss);
      ++acc_it;
    }
  }
}

// ---- store + publish ---------------------------------------------------------------------------------------------
// Two store warps: warp `sw` owns staging slot `sw` (layers with one staging tile: warp 0 alone), so the ~1.5 us a store
// takes to be read out of shared memory, performed and published overlaps with the other warp's tile instead of
// serialising (measured: one warp made 32->64 layers store-bound at ~2.3 us per tile).
__device__ __forceinline__ void store_layer(const SLayer& L, int l, const SGeom& g, uint32_t smem0, const Bars& B, Stg& stg, int lane,
                                            int sw) {
  if (L.war_layer >= 0 && lane == 0) {
    // write-after-read: every CTA must have finished the last layer that READ the buffer this layer overwrites
    spin_until(g.done + L.war_layer, gridDim.x);
  }
  __syncwarp();
  unsigned int* flags = g.flags + (size_t)l * g.nb * g.S;
  Walk w;
  w.init(g, L.d, L.rot, L.split);
  Seg s;
  uint32_t tile_it = 0;
  while (w.next(s)) {
#pragma unroll 1
    for (int kt = s.k0; kt < s.k1; ++kt) {
      uint32_t ss, use, ruse;
      stg.take(L.NS, tile_it, false, ss, use, ruse);
      ++tile_it;
      if ((int)ss != sw) continue;              // the other store warp's tile
      mbar_wait(B.stg_full + 8 * ss, use & 1u);
      if (lane == 0) {
        TR(3, stg.cnt[0] + stg.cnt[1] - 1, 0, ((unsigned long long)l << 32) | (unsigned)kt);
        TR(3, stg.cnt[0] + stg.cnt[1] - 1, 1, TR_NOW());
        const uint32_t src = smem0 + (ss ? L.offStg1 : L.offStg0);
        const int trow = s.b * g.S + s.r + kt * L.d, x0 = s.tx * 128;
        tma_store_3d(L.tm_out, 0, x0, trow, src);
        if (L.c_out == 64) tma_store_3d(L.tm_out, 64, x0, trow, src + 16384);
        bulk_commit();
        bulk_wait_read0();                      // staging tile read: it may be refilled
        TR(3, stg.cnt[0] + stg.cnt[1] - 1, 2, TR_NOW());
        B.cnt[ss] = use + 1;                    // (single writer) for the producer's residual loads
        mbar_arrive(B.stg_free + 8 * ss);       // for the epilogue of layers without a residual
        bulk_wait0();                           // writes performed
        fence_async_all();
        __threadfence();
        red_release(flags + trow);              // publish: one more tile of this image row is complete
        TR(3, stg.cnt[0] + stg.cnt[1] - 1, 3, TR_NOW());
      }
      __syncwarp();
    }
  }
  __syncwarp();
  asm volatile("bar.sync 2, %0;" ::"n"(32 * kStoreWarps) : "memory");  // both store warps have published their tiles of this layer
  if (sw == 0 && lane == 0) {
    B.cnt[2] = (uint32_t)l + 1;
    red_release(g.done + l);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kThreadsS, 1) conv2d_stream_kernel(const SLayer* __restrict__ layers, const SGeom g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sBias = reinterpret_cast<float*>(smem + kSmemCarve);       // [2][64]: double-buffered by layer parity
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemCarve + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  const uint32_t smem0 = smem_u32(smem);
  Bars B;
  B.a_full = smem_u32(bars); B.a_empty = B.a_full + 8 * kMaxNA;
  B.w_full = B.a_empty + 8 * kMaxNA; B.mma_done = B.w_full + 8;
  B.acc_full = B.mma_done + 8; B.acc_empty = B.acc_full + 8 * kAcc;
  B.res_full = B.acc_empty + 8 * kAcc; B.stg_free = B.res_full + 16; B.stg_full = B.stg_free + 16;
  B.cnt = reinterpret_cast<volatile uint32_t*>(bars + 36);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kMaxNA; ++i) { mbar_init(B.a_full + 8 * i, 1); mbar_init(B.a_empty + 8 * i, 1); }
    mbar_init(B.w_full, 1); mbar_init(B.mma_done, 1);
    B.cnt[0] = 0; B.cnt[1] = 0; B.cnt[2] = 0;
    for (int i = 0; i < kAcc; ++i) { mbar_init(B.acc_full + 8 * i, 1); mbar_init(B.acc_empty + 8 * i, kEpiWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(B.res_full + 8 * i, 1); mbar_init(B.stg_free + 8 * i, 1); mbar_init(B.stg_full + 8 * i, kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  uint32_t acc_it = 0;  // accumulator ring position runs across layers
  ARing ring;
  for (int i = 0; i < kMaxNA; ++i) ring.cnt[i] = 0;
  ring.pos = 0; ring.total = 0;
  Stg stg;
  stg.cnt[0] = stg.cnt[1] = stg.rcnt[0] = stg.rcnt[1] = 0;
  if (warp == 0) {
#pragma unroll 1
    for (int l = 0; l < g.n_layers; ++l) { s_role_layer[warp] = l; producer_layer(layers[l], l, g, smem0, B, ring, stg, lane); }
  } else if (warp == 1) {
#pragma unroll 1
    for (int l = 0; l < g.n_layers; ++l) {
      const SLayer& L = layers[l];
      s_role_layer[warp] = l;
      if (L.c_in == 64 && L.c_out == 32) mma_layer<64, 32>(L, l, g, smem0, B, tmem, ring, acc_it);
      else if (L.c_in == 32) mma_layer<32, 64>(L, l, g, smem0, B, tmem, ring, acc_it);
      else mma_layer<64, 64>(L, l, g, smem0, B, tmem, ring, acc_it);
    }
  } else if (warp < 2 + kEpiWarps) {
    const int et = tid - 64;  // 0..255
#pragma unroll 1
    for (int l = 0; l < g.n_layers; ++l) {
      const SLayer& L = layers[l];
      // bias of layer l in half (l & 1): the epilogue warps may be one layer apart (named barrier among them only)
      float* bl = sBias + (l & 1) * 64;
      s_role_layer[warp] = l;
      if (et < 64) bl[et] = (L.bias && et < L.c_out) ? L.bias[et] : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      if (L.c_out == 32) epilogue_layer<32>(L, g, smem0, smem, B, tmem, bl, acc_it, stg, warp, lane);
      else epilogue_layer<64>(L, g, smem0, smem, B, tmem, bl, acc_it, stg, warp, lane);
    }
  } else {
#pragma unroll 1
    for (int l = 0; l < g.n_layers; ++l) { s_role_layer[warp] = l; store_layer(layers[l], l, g, smem0, B, stg, lane, warp - (2 + kEpiWarps)); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- driver entry point for tensor maps (no link-time dependency on libcuda) --------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// box {64 channels, rows pixels, 1 image row} over a DMap [nb*S][S][2C] bf16, SWIZZLE_128B, zero fill out of bounds
int make_map(const DMap& m, int rows, CUtensorMap* out) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return ORCA_B200_ECUDA; }
  const cuuint64_t dims[3] = {(cuuint64_t)(2 * m.C), (cuuint64_t)m.S, (cuuint64_t)m.nb * m.S};
  const cuuint64_t strides[2] = {(cuuint64_t)m.C * 4, (cuuint64_t)m.C * 4 * m.S};
  const cuuint32_t box[3] = {64, (cuuint32_t)rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, m.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for a (%d, %d, %d, %d) map, box rows %d", (int)r, m.nb, m.C, m.S, m.S, rows);
    return ORCA_B200_ECUDA;
  }
  return ORCA_B200_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static inline uint16_t ds_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float ds_bf16_f(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int ds_debug_enable(int on) {
  const unsigned int v = on ? 1u : 0u;
  static const unsigned int zero[184] = {};
  ORCA_CUDA_OK(cudaMemcpyToSymbol(g_ds_debug_on, &v, sizeof v));
  ORCA_CUDA_OK(cudaMemcpyToSymbol(g_ds_debug, zero, sizeof zero));
  return ORCA_B200_OK;
}
int ds_debug_read(unsigned int out[184]) {
  ORCA_CUDA_OK(cudaMemcpyFromSymbol(out, g_ds_debug, 184 * sizeof(unsigned int)));
  return ORCA_B200_OK;
}

#ifdef DS_TRACE
int ds_trace_set(unsigned long long* dev_buf, int block) {
  ORCA_CUDA_OK(cudaMemcpyToSymbol(g_ds_trace, &dev_buf, sizeof dev_buf));
  if (block >= 0) ORCA_CUDA_OK(cudaMemcpyToSymbol(g_ds_trace_block, &block, sizeof block));
  return ORCA_B200_OK;
}
#endif

bool ds_layer_eligible(const ConvLayer& L) {
  return L.kh == 3 && L.kw == 3 && L.dil >= 1 && L.dil <= 64 &&
         ((L.c_in == 32 && L.c_out == 64) || ((L.c_in == 64 || L.c_in == 128) && (L.c_out == 32 || L.c_out == 64)));
}

int ds_pack_layer(ConvLayer& L, const float* w /*[tap][c_in][c_out]*/, std::vector<void*>& allocs) {
  if (!ds_layer_eligible(L)) return ORCA_B200_OK;
  const int halves = L.c_in == 128 ? 2 : 1, kc = (L.c_in / halves) / 8;
  std::vector<uint16_t> img;
  img.reserve((size_t)9 * L.c_in * L.c_out * 2);
  for (int hf = 0; hf < halves; ++hf)
    for (int tap = 0; tap < 9; ++tap)
      for (int c = 0; c < kc; ++c)
        for (int part = 0; part < 2; ++part)
          for (int n = 0; n < L.c_out; ++n)
            for (int j = 0; j < 8; ++j) {
              const int ci = hf * 64 + c * 8 + j;
              const float v = w[((size_t)tap * L.c_in + ci) * L.c_out + n];
              const uint16_t hb = ds_bf16_rn(v);
              img.push_back(part == 0 ? hb : ds_bf16_rn(v - ds_bf16_f(hb)));
            }
  void* d = nullptr;
  ORCA_CUDA_OK(cudaMalloc(&d, img.size() * 2));
  allocs.push_back(d);
  ORCA_CUDA_OK(cudaMemcpy(d, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  L.tc_w = d;
  L.tc_w_bytes = img.size() * 2;
  return ORCA_B200_OK;
}

struct DecStream::Impl {
  std::vector<SLayer> layers;
  std::vector<CUtensorMap> maps;
  std::map<std::tuple<const void*, int>, int> map_index;  // (buffer, box rows) -> index into maps
  // per layer: indices into maps (pointers are patched once the table's device address is known)
  std::vector<int> i_in, i_res, i_out;
  std::map<const void*, int> last_writer, last_reader;
  SGeom g{};
  double flop = 0.0;
  int stg0_prev = -1, stg1_prev = -1, stg_size_prev = 0, ns_prev = 0;

  int map_for(const DMap& m, int rows, int* idx) {
    const auto key = std::make_tuple((const void*)m.p, rows);
    auto it = map_index.find(key);
    if (it != map_index.end()) { *idx = it->second; return ORCA_B200_OK; }
    CUtensorMap tm;
    ORCA_TRY(make_map(m, rows, &tm));
    maps.push_back(tm);
    *idx = (int)maps.size() - 1;
    map_index[key] = *idx;
    return ORCA_B200_OK;
  }
};

DecStream::DecStream() : impl(new Impl) {}
DecStream::~DecStream() { delete impl; }
int DecStream::size() const { return (int)impl->layers.size(); }
double DecStream::flop() const { return impl->flop; }

static bool overlaps(int a0, int a1, int b0, int b1) { return a0 < b1 && b0 < a1; }

int DecStream::add(const ConvLayer& L, int k_half, int use_bias, const DMap& in, const DMap* res, DMap* out, int relu) {
  if (!L.tc_w || !ds_layer_eligible(L)) { set_error("DecStream: layer %d->%d has no stream weights", L.c_in, L.c_out); return ORCA_B200_EUNSUPPORTED; }
  const int c_in = L.c_in == 128 ? 64 : L.c_in;
  if ((L.c_in == 128) != (k_half >= 0) || in.C != L.c_in || out->C != L.c_out || out->S != in.S || out->nb != in.nb ||
      (res && (res->C != L.c_out || res->S != in.S || res->nb != in.nb)) || out->p == in.p || (res && res->p == out->p)) {
    set_error("DecStream: geometry mismatch (%d->%d, k_half %d)", L.c_in, L.c_out, k_half);
    return ORCA_B200_EINVAL;
  }
  SGeom& g = impl->g;
  if (impl->layers.empty()) {
    g.nb = in.nb; g.S = in.S; g.tpr = (in.S + 127) / 128; g.total_tiles = in.nb * g.tpr * in.S;
  } else if (g.S != in.S || g.nb != in.nb) {
    set_error("DecStream: all layers must share the map geometry");
    return ORCA_B200_EINVAL;
  }
  const int l = (int)impl->layers.size();
  SLayer t{};
  t.c_in = c_in; t.c_out = L.c_out; t.d = L.dil; t.relu = relu;
  t.R = 128 + 2 * L.dil;
  const int R8 = (t.R + 7) & ~7;
  t.a_box_bytes = R8 * 128;
  t.a_slot_bytes = (c_in == 64 ? 2 : 1) * t.a_box_bytes;
  t.w_bytes = 9 * (c_in / 8) * 2 * L.c_out * 16;
  t.w = static_cast<const uint8_t*>(L.tc_w) + (k_half > 0 ? (size_t)t.w_bytes : 0);
  t.bias = use_bias ? L.b : nullptr;
  if (in.C == 32) { t.in_c0 = 0; t.in_c1 = 0; }
  else { t.in_c0 = (k_half > 0 ? 64 : 0); t.in_c1 = in.C + t.in_c0; }
  // ---- shared-memory carve (bytes from the 1024-aligned base) ----
  const int stg_size = L.c_out * 4 * 128;  // 16 KB (c_out 32) / 32 KB (c_out 64)
  t.offW = 0;
  if (t.w_bytes <= 73728) {
    t.offStg0 = 73728; t.offStg1 = 73728 + 32768; t.NS = 2;
    t.offA = 73728 + 65536;
    int na = (kSmemCarve - t.offA) / t.a_slot_bytes;
    if (na < 2 && L.c_out == 32) {  // wide runs of the 64->32 layers: staging tiles are 16 KB, start the A slots lower
      t.offA = t.offStg1 + stg_size;
      na = (kSmemCarve - t.offA) / t.a_slot_bytes;
      if (na < 2) { t.NS = 1; t.offA = t.offStg0 + stg_size; na = (kSmemCarve - t.offA) / t.a_slot_bytes; }
    }
    t.NA = na > kMaxNA ? kMaxNA : na;
  } else {
    t.offStg0 = t.w_bytes; t.offStg1 = t.offStg0; t.NS = 1;
    t.offA = t.offStg0 + stg_size;
    const int na = (kSmemCarve - t.offA) / t.a_slot_bytes;
    t.NA = na > kMaxNA ? kMaxNA : na;
  }
  if (t.NA < 1) { set_error("DecStream: shared memory budget exceeded (%d->%d, d=%d)", L.c_in, L.c_out, L.dil); return ORCA_B200_EUNSUPPORTED; }
  // Must this layer wait for the previous layer's stores to drain before touching shared memory?  Yes when its
  // weights / runs overlap a staging tile of the previous carve, or when one of its staging tiles overlaps a
  // DIFFERENT slot of the previous carve (the same slot index is serialised by that slot's free / full barriers).
  t.drain_before = 0;
  if (l > 0) {
    const int w1 = t.offW + t.w_bytes, a1 = t.offA + t.NA * t.a_slot_bytes;
    const int pst[2] = {impl->stg0_prev, impl->stg1_prev}, nst[2] = {t.offStg0, t.offStg1};
    const int ps = impl->stg_size_prev;
    bool hit = false;
    for (int sp = 0; sp < impl->ns_prev; ++sp) {
      hit = hit || overlaps(t.offW, w1, pst[sp], pst[sp] + ps) || overlaps(t.offA, a1, pst[sp], pst[sp] + ps);
      for (int sn = 0; sn < t.NS; ++sn)
        if (sn != sp) hit = hit || overlaps(nst[sn], nst[sn] + stg_size, pst[sp], pst[sp] + ps);
    }
    t.drain_before = hit ? 1 : 0;
  }
  impl->stg0_prev = t.offStg0; impl->stg1_prev = t.offStg1; impl->stg_size_prev = stg_size; impl->ns_prev = t.NS;
  // rotation of the tile walk: +2 per consecutive layer with the same dilation (see Walk)
  t.rot = (l > 0 && impl->layers[l - 1].d == t.d) ? impl->layers[l - 1].rot + 2 : 0;
  t.split = (l == 0 || impl->layers[l - 1].d != t.d) ? 1 : 0;  // per-image pieces where the dependencies are all-to-all
  // ---- dependencies ----
  auto find = [](const std::map<const void*, int>& m, const void* p) { auto it = m.find(p); return it == m.end() ? -1 : it->second; };
  t.in_layer = find(impl->last_writer, in.p);
  t.res_layer = res ? find(impl->last_writer, res->p) : -1;
  t.war_layer = find(impl->last_reader, out->p);
  {
    const int ww = find(impl->last_writer, out->p);  // write-after-write: order behind the previous writer's readers at least
    if (ww > t.war_layer) t.war_layer = ww;
  }
  impl->last_reader[in.p] = l;
  if (res) impl->last_reader[res->p] = l;
  impl->last_writer[out->p] = l;
  int idx;
  ORCA_TRY(impl->map_for(in, t.R > 256 ? 256 : t.R, &idx));
  impl->i_in.push_back(idx);
  if (res) { ORCA_TRY(impl->map_for(*res, 128, &idx)); } else idx = -1;
  impl->i_res.push_back(idx);
  ORCA_TRY(impl->map_for(*out, 128, &idx));
  impl->i_out.push_back(idx);
  impl->flop += 2.0 * in.nb * in.S * in.S * (double)c_in * L.c_out * 9;
  impl->layers.push_back(t);
  return ORCA_B200_OK;
}

size_t DecStream::scratch_bytes(int max_layers, int nb, int S) {
  // [flags + done | tensor maps (3 per layer at most, 128 B each) | layer table | per-CTA tile ranges]
  const size_t counters = ((size_t)max_layers * nb * S + max_layers) * 4;
  return ((counters + 255) & ~size_t(255)) + (size_t)3 * max_layers * sizeof(CUtensorMap) + (size_t)max_layers * sizeof(SLayer) +
         (size_t)256 * nb * sizeof(int2) + 2048;
}

int DecStream::run(void* scratch, size_t scratch_bytes_, cudaStream_t s) {
  const int n = (int)impl->layers.size();
  if (n == 0) return ORCA_B200_OK;
  SGeom g = impl->g;
  g.n_layers = n;
  const size_t counters = (((size_t)n * g.nb * g.S + n) * 4 + 255) & ~size_t(255);
  const size_t maps_bytes = impl->maps.size() * sizeof(CUtensorMap);
  const size_t parts_off = counters + ((maps_bytes + 255) & ~size_t(255)) + (((size_t)n * sizeof(SLayer) + 255) & ~size_t(255));
  const size_t need = parts_off + (size_t)256 * g.nb * sizeof(int2) + 256;
  if (scratch_bytes_ < need || (reinterpret_cast<uintptr_t>(scratch) & 255)) { set_error("DecStream: scratch too small or misaligned"); return ORCA_B200_EWORKSPACE; }
  char* base = static_cast<char*>(scratch);
  g.flags = reinterpret_cast<unsigned int*>(base);
  g.done = g.flags + (size_t)n * g.nb * g.S;
  CUtensorMap* d_maps = reinterpret_cast<CUtensorMap*>(base + counters);
  SLayer* d_layers = reinterpret_cast<SLayer*>(base + counters + ((maps_bytes + 255) & ~size_t(255)));
  for (int l = 0; l < n; ++l) {
    SLayer& t = impl->layers[l];
    t.tm_in = d_maps + impl->i_in[l];
    t.tm_res = impl->i_res[l] >= 0 ? d_maps + impl->i_res[l] : nullptr;
    t.tm_out = d_maps + impl->i_out[l];
  }
  ORCA_CUDA_OK(cudaMemsetAsync(base, 0, counters, s));
  ORCA_CUDA_OK(cudaMemcpyAsync(d_maps, impl->maps.data(), maps_bytes, cudaMemcpyHostToDevice, s));
  ORCA_CUDA_OK(cudaMemcpyAsync(d_layers, impl->layers.data(), (size_t)n * sizeof(SLayer), cudaMemcpyHostToDevice, s));
  static bool configured_dev[64] = {};
  static int sms_dev[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!configured_dev[dev & 63]) {
    ORCA_CUDA_OK(cudaFuncSetAttribute(conv2d_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    cudaDeviceGetAttribute(&sms_dev[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    configured_dev[dev & 63] = true;
  }
  const int sms = sms_dev[dev & 63] > 0 ? sms_dev[dev & 63] : 148;
  int grid = g.total_tiles < sms ? g.total_tiles : sms;
  if (grid > 256) grid = 256;
  {
    // Every CTA gets a contiguous piece of every image.  Images 0..nb-2 are split evenly; the last image's pieces make
    // each CTA's TOTAL equal to the even split of all tiles (so the per-CTA totals differ by at most one tile).
    const int T_img = g.tpr * g.S;
    std::vector<int2> parts((size_t)grid * g.nb);
    bool ok = true;
    int last_start = 0;
    for (int c = 0; c < grid && ok; ++c) {
      int total = (int)(((long long)(c + 1) * g.total_tiles) / grid - ((long long)c * g.total_tiles) / grid);
      for (int b = 0; b < g.nb - 1; ++b) {
        const int a0 = (int)(((long long)c * T_img) / grid), a1 = (int)(((long long)(c + 1) * T_img) / grid);
        parts[(size_t)c * g.nb + b] = make_int2(a0, a1 - a0);
        total -= a1 - a0;
      }
      if (total < 0) { ok = false; break; }
      parts[(size_t)c * g.nb + g.nb - 1] = make_int2(last_start, total);
      last_start += total;
    }
    if (!ok || last_start != T_img) {  // tiny maps with many images: plain even split of every image
      for (int c = 0; c < grid; ++c)
        for (int b = 0; b < g.nb; ++b) {
          const int a0 = (int)(((long long)c * T_img) / grid), a1 = (int)(((long long)(c + 1) * T_img) / grid);
          parts[(size_t)c * g.nb + b] = make_int2(a0, a1 - a0);
        }
    }
    int2* d_parts = reinterpret_cast<int2*>(base + parts_off);
    ORCA_CUDA_OK(cudaMemcpyAsync(d_parts, parts.data(), parts.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
    g.parts = d_parts;
  }
  const SLayer* lp = d_layers;
  void* args[] = {(void*)&lp, (void*)&g};
  // cooperative launch: the row flags are spin-waited, so every CTA must be resident (1 per SM)
  ORCA_CUDA_OK(cudaLaunchCooperativeKernel((const void*)conv2d_stream_kernel, dim3(grid), dim3(kThreadsS), args, (size_t)kSmemTotal, s));
  ORCA_LAUNCH_OK();
  return ORCA_B200_OK;
}

}  // namespace orca
