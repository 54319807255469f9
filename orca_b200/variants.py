"""
Variant-level batching: the structural-variant entry points of the reference
(`orca_predict.process_del / process_dup / process_inv / process_ins / process_single_breakpoint`,
orca_predict.py:1172-3160) call `genomepredict` three or four times per variant -- ref.l, ref.r, alt(.l/.r) -- on 32 Mb
windows that share most of their base pairs (e.g. orca_predict.py:1673, :1728, :1794 for a deletion).  The reference encodes
every window from scratch (>85 % of a pass).  Here:

  * `EncoderBlockCache` keeps the 4 kb-resolution encoder output per aligned block of bins, keyed on a hash of the bases the
    block can see (the block plus the encoder's 112 kb halo on each side, and whether the window is clipped by a true
    sequence end).  A later window that contains the same bases at the same 4 kb phase re-uses the block; only the blocks
    around a breakpoint (and everything downstream of an indel whose length is not a multiple of 4 kb) are encoded again.
    Re-used and freshly encoded blocks are bit-identical to a monolithic encode: the chunked encoder equals the single pass
    bit for bit (tests/test_gpu_parity.py::test_encoder_layouts_and_chunks).
  * `predict_variant_windows` runs the decoder cascades of ALL windows and both strands as ONE batched chain (2 x n lanes per
    decoder call, each lane with its own zoom path), instead of 2 x n separate chains.

The hash is computed on the device from the packed bases (two independent 64-bit multiplicative hashes, 128 bits per key), so
a 32 Mb window costs one 32 MB upload and a few short kernels before any encoder work is decided.
"""
import numpy as np
import torch

from . import feeder, predict

HALO_BP = 112000 + 4000  # as orca_b200.parallel: the encoder's halo (+1 bin of slack for the k = 9 taps at the edge)


class EncoderBlockCache:
    """Per (encoder weights, strand) cache of encoder outputs in blocks of `block_bins` 4 kb bins."""

    def __init__(self, block_bins=250, max_blocks=4096, seed=0x0CA):
        self.block_bins = int(block_bins)
        self.max_blocks = int(max_blocks)
        self._store = {}      # key -> (block_bins, 128) float32 device tensor
        self._weights = {}    # (device, n) -> two int64 weight vectors for the hash
        self._seed = seed
        self.hits = 0
        self.misses = 0

    # -- hashing ----------------------------------------------------------------------------------------
    def _hash_weights(self, device, n):
        key = (str(device), n)
        if key not in self._weights:
            g = torch.Generator().manual_seed(self._seed + n)
            w = torch.randint(-(1 << 62), 1 << 62, (2, n), generator=g, dtype=torch.int64) | 1  # odd multipliers
            self._weights[key] = w.to(device)
        return self._weights[key]

    def _block_keys(self, codes_dev, L, weights_version, reverse):
        """One key per block of this window: (weights, strand, clip flags, length, 128-bit hash of the visible bases)."""
        bb = self.block_bins
        n_blocks = (L // 4000 + bb - 1) // bb
        span = bb * 4000 + 2 * HALO_BP
        w = self._hash_weights(codes_dev.device, span)
        c = codes_dev.to(torch.int64) + 1  # 1..5: a run of 'A' (code 0) must not hash like an empty window
        sums = []
        meta = []
        for b in range(n_blocks):
            lo, hi = max(b * bb * 4000 - HALO_BP, 0), min((b + 1) * bb * 4000 + HALO_BP, L)
            off = lo - (b * bb * 4000 - HALO_BP)  # where the clipped window starts inside the nominal span
            seg = c[lo:hi]
            sums.append((seg[None, :] * w[:, off:off + seg.numel()]).sum(1))  # int64 wrap-around arithmetic = mod 2^64
            meta.append((lo == 0, hi == L, hi - lo))
        h = torch.stack(sums).cpu().numpy()  # one small D2H for the whole window
        return [(weights_version, bool(reverse), m, int(h[i, 0]), int(h[i, 1])) for i, m in enumerate(meta)]

    # -- encode ------------------------------------------------------------------------------------------
    def encode(self, net0, codes_dev, reverse=False):
        """Encoder output (1, 128, L / 4000) of one strand of a packed sequence (uint8 (L,) device tensor, codes 0..4),
        re-using cached blocks.  `reverse`: the reverse-complement strand, read in place from the same bytes."""
        L = codes_dev.numel()
        if L % 4000:
            raise ValueError("sequence length must be a multiple of 4000")
        P, bb = L // 4000, self.block_bins
        dev = codes_dev.device
        net0.native_handle(dev)
        keys = self._block_keys(codes_dev, L, (id(net0), net0._handle_version), reverse)
        out = torch.empty((1, P, 128), dtype=torch.float32, device=dev)
        x = codes_dev[None]
        missing = []
        for b, k in enumerate(keys):
            b0, b1 = b * bb, min((b + 1) * bb, P)
            # the reverse strand's bins are the mirror image of the forward window's
            o0, o1 = (P - b1, P - b0) if reverse else (b0, b1)
            hit = self._store.get(k)
            if hit is not None:
                out[0, o0:o1] = hit[:o1 - o0]
                self.hits += 1
            else:
                missing.append((b, o0, o1))
                self.misses += 1
        # encode runs of consecutive missing blocks with one call each (one halo recompute per run, not per block)
        i = 0
        while i < len(missing):
            j = i
            while j + 1 < len(missing) and missing[j + 1][0] == missing[j][0] + 1:
                j += 1
            lo_bin = min(missing[i][1], missing[j][1])
            hi_bin = max(missing[i][2], missing[j][2])
            net0(x, bin_range=(lo_bin, hi_bin), out=out, reverse_complement=reverse, guard=False)
            i = j + 1
        for b, o0, o1 in missing:
            if len(self._store) >= self.max_blocks:
                self._store.pop(next(iter(self._store)))
            self._store[keys[b]] = out[0, o0:o1].clone()
        return out.transpose(1, 2)


def _as_codes(sequence):
    """One sequence as packed codes: (L,) uint8 numpy from text / bytes / codes / a (1, L, 4) or (L, 4) one-hot array."""
    if isinstance(sequence, np.ndarray) and sequence.dtype != np.uint8:
        arr = sequence[None] if sequence.ndim == 2 else sequence
        return feeder.from_onehot(arr)[0]
    c = feeder.codes(sequence) if isinstance(sequence, (str, bytes, bytearray, memoryview)) else np.asarray(sequence)
    c = c[0] if c.ndim == 2 else c
    if c.max(initial=0) > 4:
        c = feeder.codes(c.tobytes())
    return np.ascontiguousarray(c, dtype=np.uint8)


def predict_variant_windows(windows, model, cache=None):
    """genomepredict for several windows of ONE variant call with one model.

    windows: [(sequence, mchr, mpos, wpos), ...] with `sequence` a 32 Mb window as the reference passes it ((1, L, 4)
    one-hot), or packed bases.  Returns (outputs, cache): one dict per window in orca_predict.genomepredict's format
    (`predictions[0]` = the six strand-averaged maps).  Pass the returned cache to the next call of the same variant set
    (or of neighbouring variants) to keep re-using blocks."""
    device = predict._device_of(model)
    cache = cache if cache is not None else EncoderBlockCache()
    with torch.no_grad(), torch.cuda.device(device):
        for attempt in range(2):
            lanes, meta = [], []
            for sequence, mchr, mpos, wpos in windows:
                codes = torch.from_numpy(_as_codes(sequence)).to(device, non_blocking=True)
                enc = torch.cat([cache.encode(model.net0, codes, False), cache.encode(model.net0, codes, True)], 0)
                outs = model.net(enc)
                for i, rev in enumerate((False, True)):
                    lanes.append(({lvl: t[i:i + 1] for lvl, t in zip([1, 2, 4, 8, 16, 32], outs)}, rev, mpos, wpos))
                meta.append((mchr, mpos, wpos))
            preds, starts = predict.cascade_32mb_lanes(model, lanes, None, None)
            host = []
            for w in range(len(windows)):
                avg = predict._average_strands([p[2 * w:2 * w + 1] for p in preds], [p[2 * w + 1:2 * w + 2] for p in preds])
                host.append(torch.stack(avg).cpu().numpy())
            if not predict.check_fp16_guard([model]):
                break
            cache._store.clear()  # encoded at the wrong precision
    outputs = [predict._output_32mb([h], starts[2 * w], meta[w][0], meta[w][2], [model]) for w, h in enumerate(host)]
    return outputs, cache
