"""
Multi-GPU forward: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The Encoder is >85 % of a pass and shards along the sequence: rank r owns the
4 kb bins [r*P/R, (r+1)*P/R) and needs the input bp of that span plus the 112 kb halo the
reference's own block overlap uses (orca_modules.py:931-932) -- every rank slices its window
out of the host copy, so there is no input exchange at all.  The one real exchange step is an
all-gather of the (P/R, 128) fp32 encodings (0.5 MB per rank at 32 Mb, 4.1 MB at 256 Mb), after
which Encoder2/Encoder3 and the decoder cascades (<12 % of the FLOPs, global receptive field) run
un-sharded: the forward-strand cascade on rank 0, the reverse-complement cascade on rank 1.

For the reverse strand a rank encodes the MIRRORED bins [P-b1, P-b0): they read exactly the same
forward-strand window walked backwards, so one upload serves both strands.
"""
import threading

import numpy as np
import torch
import torch.distributed as dist

from . import predict

HALO_BP = 112000 + 4000  # the encoder's 112 kb halo (+1 bin of slack for the k=9 taps at the edge)


def shard_bins(P, rank, world):
    """Contiguous, near-equal bin ranges; equal when world divides P (8000 and 64000 / 1,2,4,8)."""
    base, rem = divmod(P, world)
    b0 = rank * base + min(rank, rem)
    return b0, b0 + base + (1 if rank < rem else 0)


def shard_window(L, b0, b1):
    """Forward-strand bp window a shard must hold to encode bins [b0, b1) (and their mirror)."""
    return max(b0 * 4000 - HALO_BP, 0), min(b1 * 4000 + HALO_BP, L)


class ShardedForward:
    """genomepredict-equivalent forward of one shell, sequence-sharded over `world` ranks: the 32 Mb shells
    (6 maps, 32 -> 1 Mb) and the 256 Mb shells (4 maps, 256 -> 32 Mb; call set_background first)."""

    def __init__(self, shell, L, rank=0, world=1, device=None):
        self.shell, self.L, self.rank, self.world = shell, L, rank, world
        self.device = device if device is not None else predict._device_of(shell)
        self.P = L // 4000
        if world > 1 and self.P % world != 0:
            raise ValueError("number of 4 kb bins (%d) must be divisible by the world size (%d)" % (self.P, world))
        self.b0, self.b1 = shard_bins(self.P, rank, world)
        self.s0, self.s1 = shard_window(L, self.b0, self.b1)
        self.window = None
        self.background, self.chrlen = None, None
        self._copy_stream = None
        self._uploader, self._staging, self._staging_ev = None, None, None
        self._ready = None
        self._cuts = [self.b0, self.b1]
        # upload pieces (fractions of this rank's bins) overlapped with the encoder: a small first piece (the encoder starts
        # after 1/32 of the upload), then growing ones; every extra cut costs one 2 x 112 kb halo recompute
        self.pieces = (0.0, 0.03125, 0.125, 0.5, 1.0)
        self.concurrent_strands = False  # measured: no gain (the big conv kernels fill the GPU) and 2x workspace
        # how a single GPU runs the two strand cascades: "batch" = the two strands as the two batch elements of ONE
        # chain (every decoder call is one persistent stream kernel at batch 2); "serial" = one strand after the other
        # at batch 1, which is what each cascade rank of a multi-GPU run does
        self.cascade_mode = "batch"
        self.h2d_bytes = 0
        # maps per level: 1, or num_2d for the multi-dataset shells of orca_leukemia.py
        denets = getattr(shell, "denets", None) or {}
        self.n_ch = int(getattr(next(iter(denets.values()), None), "num_2d", 1))
        self.d2h_bytes = 6 * self.n_ch * 250 * 250 * 4 if rank == 0 else 0

    def set_background(self, normmat, chrlen):
        """256 Mb shells: the caller's (8000, 8000) background matrix at 32 kb bins (as passed to
        genomepredict_256Mb) and the chromosome length; uploaded once per rank that runs a cascade."""
        self.background = predict.prepare_background(normmat, self.device) if self.rank <= 1 else None
        self.chrlen = chrlen

    def upload(self, seq_host):
        """seq_host: (1, L, 4) float32 CPU tensor, or packed bases as a (1, L) uint8 CPU tensor (orca_b200.feeder;
        16x fewer bytes over PCIe) -- pinned for full-speed copies; uploads this rank's window.

        On CUDA the window goes up in `self.pieces` pieces on a copy stream: the forward-strand encoder starts on
        the first range of bins as soon as piece 1 (those bins + halo) has landed, while the rest is in flight."""
        packed = seq_host.dtype == torch.uint8
        if packed and seq_host.dim() != 2:
            raise ValueError("packed sequence must be a (1, L) uint8 tensor")
        sl = seq_host[:, self.s0:self.s1]
        self.h2d_bytes = sl.numel() * sl.element_size()
        self._ready = None
        if self.device.type != "cuda":
            self.window = sl.to(self.device)
            return self.window
        n = sl.shape[1]
        # a small first piece (the encoder can start after ~1/8 of the upload), then pieces aligned with the
        # encoder's own chunk boundaries so that no extra halo is recomputed
        span = self.b1 - self.b0
        self._cuts = sorted({self.b0 + int(span * f) for f in self.pieces})  # bin boundaries
        ends = [min(max(c * 4000 + HALO_BP - self.s0, 0), n) for c in self._cuts[1:-1]] + [n]  # window rows per piece
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        if self.window is None or self.window.shape[1] != n or self.window.dtype != sl.dtype:
            self.window = torch.empty((1, n) if packed else (1, n, 4), dtype=sl.dtype, device=self.device)
        cs = self._copy_stream
        cs.wait_stream(main)  # earlier kernels may still be reading the previous contents
        self._join_uploader()
        if not sl.is_pinned() and sl.numel() * sl.element_size() >= (64 << 20):
            # PAGEABLE host memory (what orca_predict's callers hold): a helper thread copies it through two pinned
            # staging buffers piece by piece; the main thread goes on to enqueue the encoder, which starts on piece 1
            # while the helper is still staging the rest (torch copies release the GIL).
            events = [torch.cuda.Event() for _ in ends]
            recorded = [threading.Event() for _ in ends]
            self._uploader = threading.Thread(target=self._upload_pageable, args=(sl, ends, events, recorded), daemon=True)
            self._uploader.start()
            self._ready = list(zip(events, recorded))
            return self.window
        events, lo = [], 0
        with torch.cuda.stream(cs):
            for hi in ends:
                if hi > lo:
                    self.window[:, lo:hi].copy_(sl[:, lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                events.append((ev, None))
                lo = max(lo, hi)
        self._ready = events
        return self.window

    _STAGE_ROWS = 1 << 21  # rows (bp) per pinned staging buffer: 32 MB of fp32 one-hot

    def _upload_pageable(self, sl, ends, events, recorded):
        cs = self._copy_stream
        row_shape = tuple(sl.shape[2:])
        if self._staging is None or self._staging[0].dtype != sl.dtype or tuple(self._staging[0].shape[1:]) != row_shape:
            self._staging = [torch.empty((self._STAGE_ROWS,) + row_shape, dtype=sl.dtype).pin_memory() for _ in range(2)]
            self._staging_ev = [torch.cuda.Event() for _ in range(2)]
        k, lo = 0, 0
        try:
            with torch.cuda.device(self.device), torch.cuda.stream(cs):
                for i, hi in enumerate(ends):
                    while lo < hi:
                        n = min(self._STAGE_ROWS, hi - lo)
                        b = k & 1
                        if k >= 2:
                            self._staging_ev[b].synchronize()  # the DMA that last read this staging buffer is done
                        self._staging[b][:n].copy_(sl[0, lo:lo + n])
                        self.window[0, lo:lo + n].copy_(self._staging[b][:n], non_blocking=True)
                        self._staging_ev[b].record(cs)
                        lo += n
                        k += 1
                    events[i].record(cs)
                    recorded[i].set()
        finally:
            for r in recorded:
                r.set()

    def _join_uploader(self):
        if self._uploader is not None:
            self._uploader.join()
            self._uploader = None

    def _encode_local(self, reverse):
        """This rank's bins of one strand -> (1, P, 128) buffer (other bins undefined)."""
        enc = torch.empty((1, self.P, 128), dtype=torch.float32, device=self.device)
        bins = (self.P - self.b1, self.P - self.b0) if reverse else (self.b0, self.b1)
        kw = dict(out=enc, reverse_complement=reverse, window=(self.s0, self.L))
        if hasattr(self.shell.net0, "fp16_guard_fired"):
            kw["guard"] = False  # native Encoder: the fp16 range guard is checked once per pass, see fp16_guard()
        x = self.window if self.window.dtype == torch.uint8 else self.window.transpose(1, 2)
        ready, self._ready = (self._ready, None) if not reverse else (None, self._ready)
        def wait(item):
            ev, recorded = item
            if recorded is not None:
                recorded.wait()  # the helper thread has recorded the event (pageable upload)
            torch.cuda.current_stream(self.device).wait_event(ev)
        if ready is not None and len(ready) > 1:  # first use after a staged upload (forward strand)
            for i, item in enumerate(ready):
                wait(item)
                self.shell.net0(x, bin_range=(self._cuts[i], self._cuts[i + 1]), **kw)
        else:
            pending = ready or self._ready
            if pending:
                wait(pending[-1])
                self._ready = None
            self.shell.net0(x, bin_range=bins, **kw)
        return enc

    def fp16_guard(self):
        """Call after the maps of a forward() have been read back: True if this rank's encoder tripped its fp16 range
        guard (it then runs fp32-grade from now on and the forward must be repeated -- on EVERY rank, so callers
        all-reduce the flag when world > 1)."""
        return predict.check_fp16_guard([self.shell])

    def _gather(self, enc, reverse):
        P, n = self.P, self.b1 - self.b0
        if self.world > 1:
            bins = (P - self.b1, P - self.b0) if reverse else (self.b0, self.b1)
            mine = enc[0, bins[0]:bins[1]].contiguous()
            gathered = torch.empty((self.world * n, 128), dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(gathered, mine)  # concatenation along dim 0 (NCCL and gloo agree)
            gathered = gathered.view(self.world, n, 128)
            if reverse:  # rank r holds the mirrored range -> chunk order is reversed
                gathered = torch.flip(gathered, [0])
            enc = gathered.reshape(1, P, 128)
        return enc.transpose(1, 2)

    def _encode(self, reverse):
        return self._gather(self._encode_local(reverse), reverse)

    def forward(self, mpos, wpos):
        """Returns the strand-averaged maps on rank 0 (None elsewhere): (6, 250, 250) for the 32 Mb shells,
        (4, 250, 250) for the 256 Mb shells, (6, num_2d, 250, 250) for multi-dataset shells."""
        shell, world, rank = self.shell, self.world, self.rank
        with torch.no_grad():
            if self.concurrent_strands and self.device.type == "cuda":
                # the two strands' encoders are independent: two streams let one strand's small late-stage
                # kernels and launch gaps hide behind the other's big ones; the collectives stay on one stream
                loc_f, loc_r = predict.run_concurrent([lambda: self._encode_local(False), lambda: self._encode_local(True)],
                                                      self.device)
            else:
                loc_f, loc_r = self._encode_local(False), self._encode_local(True)
            enc_f, enc_r = self._gather(loc_f, False), self._gather(loc_r, True)
            rev_rank = 1 if world > 1 else 0
            has_1m = hasattr(shell, "denet_1_pt")
            # where the independent Decoder_1m term of each strand runs: beside the cascades on this GPU (extra
            # streams) when there are idle ranks none, on ranks 2 / 3 when the box has them
            x_rank = {False: 2, True: 3} if world >= 4 else {False: 0, True: rev_rank}
            preds, extras = {}, {}
            enc_of = {False: enc_f, True: enc_r}
            nets = {}

            is256 = str(getattr(shell, "kind", "")).endswith("256m")
            n_maps = 4 if is256 else 6

            def finest(rev):
                if rev not in nets:
                    if is256:  # net(net1(net0(x))[-1])  (orca_predict.py:675-683); only the pooling half of net1 is needed
                        e128 = shell.net1(enc_of[rev], coarsest_only=True)[-1]
                        nets[rev] = dict(zip([32, 64, 128, 256], shell.net(e128)))
                    else:
                        nets[rev] = dict(zip([1, 2, 4, 8, 16, 32], shell.net(enc_of[rev])))
                return nets[rev]

            def cascade(rev):
                if is256:
                    if self.background is None:
                        raise RuntimeError("256 Mb shells need set_background(normmat, chrlen) before forward()")
                    p = predict.cascade_256mb(shell, finest(rev), 1, self.background, self.chrlen, mpos, wpos, rev)[0]
                else:
                    p, _ = predict.cascade_32mb(shell, finest(rev), 1, mpos, wpos, rev, inline_1m=False)
                return torch.stack([t[0] for t in p], 0)  # (n_maps, C, 250, 250)

            if world == 1 and self.cascade_mode == "batch":
                # the U-nets of the two strands as one batch-2 call (their ~40 small launches per call are latency-bound)
                both_enc = torch.cat([enc_f, enc_r], 0)
                if is256:
                    outs = shell.net(shell.net1(both_enc, coarsest_only=True)[-1])
                    levels = [32, 64, 128, 256]
                else:
                    outs = shell.net(both_enc)
                    levels = [1, 2, 4, 8, 16, 32]
                for i, rev in enumerate((False, True)):
                    nets[rev] = {lvl: t[i:i + 1] for lvl, t in zip(levels, outs)}
                lanes = [(finest(False), False), (finest(True), True)]
                if is256:
                    if self.background is None:
                        raise RuntimeError("256 Mb shells need set_background(normmat, chrlen) before forward()")
                    p, _ = predict.cascade_256mb_lanes(shell, lanes, self.background, self.chrlen, mpos, wpos)
                else:
                    p, _ = predict.cascade_32mb_lanes(shell, lanes, mpos, wpos, inline_1m=True)
                both = torch.stack(p, 0)  # (n_maps, 2 strands, C, 250, 250)
                out = 0.5 * both[:, 0] + 0.5 * torch.flip(both[:, 1], [2, 3])
                return out[:, 0] if self.n_ch == 1 else out

            jobs = []
            for rev, owner in ((False, 0), (True, rev_rank)):
                if rank == owner:
                    finest(rev)  # Encoder2 once, on the main stream, before the concurrent jobs read it
                    jobs.append(("c", rev))
                if has_1m and rank == x_rank[rev]:
                    finest(rev)
                    jobs.append(("x", rev))
            # every decoder call is one persistent kernel that owns all SMs: this rank's chains run back to back
            outs = [cascade(r) if kind == "c" else predict.level1_extra(shell, finest(r), mpos, wpos, r)[0] for kind, r in jobs]
            for (kind, rev), o in zip(jobs, outs):
                (preds if kind == "c" else extras)[rev] = o
            if world > 1:  # collect on rank 0: the reverse cascade, and the Decoder_1m terms computed elsewhere
                # ONE batched group of point-to-point operations (ncclGroupStart/End underneath) instead of three
                # serialised send/recv pairs
                ops, landing = [], []

                def move(t_dict, rev, src, shape):
                    if src == 0:
                        return
                    if rank == src:
                        ops.append(dist.P2POp(dist.isend, t_dict[rev].contiguous(), 0))
                    elif rank == 0:
                        buf = torch.empty(shape, dtype=torch.float32, device=self.device)
                        ops.append(dist.P2POp(dist.irecv, buf, src))
                        landing.append((t_dict, rev, buf))
                move(preds, True, rev_rank, (n_maps, self.n_ch, 250, 250))
                if has_1m:
                    move(extras, False, x_rank[False], (self.n_ch, 250, 250))
                    move(extras, True, x_rank[True], (self.n_ch, 250, 250))
                if ops:
                    for req in dist.batch_isend_irecv(ops):
                        req.wait()
                for t_dict, rev, buf in landing:
                    t_dict[rev] = buf
            if rank != 0:
                return None
            if has_1m:
                for rev in (False, True):
                    preds[rev][5] += extras[rev]
            out = 0.5 * preds[False] + 0.5 * torch.flip(preds[True], [2, 3])
            return out[:, 0] if self.n_ch == 1 else out
