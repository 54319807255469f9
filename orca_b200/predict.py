"""
B200-side multiscale drivers: the same call contract and return dict as
`orca_predict.genomepredict` (orca_predict.py:231-540) and `genomepredict_256Mb` (:543-878),
restated for the native modules so that a pass costs ONE host->device upload:

  * the sequence is uploaded once; the reverse-complement strand is read in place (the RC of a
    (L,4) one-hot array is the same buffer walked backwards, orca_predict.py:324-329), instead
    of `sequence[:, ::-1, ::-1].copy()` + a second upload per model;
  * `log(normmats[level])` lives on the device per shell (the reference re-uploads it on every
    eval_step, :349-353); the 256 Mb background levels are block-averaged on the GPU
    (orca_b200_background_forward) instead of float64 numpy on the host (:724-737);
  * strand averaging `0.5*fwd + 0.5*rev[::-1, ::-1]` (:510-523) happens on the device and the
    maps come back in one device->host copy.

The unmodified reference drivers also work with these shells (`models=[shell]`), because the
modules keep the reference signatures; this file is the fast path, not a requirement.
`targets` / `annotation` (plot-only bookkeeping) are out of scope.
"""
import numpy as np
import torch

from . import _lib, feeder


def _device_of(model):
    for p in model.parameters():
        return p.device
    raise RuntimeError("model has no parameters")


def is_packed(sequence):
    """Packed-base input (orca_b200.feeder): text, bytes, or a uint8 (L,) / (B, L) array or tensor."""
    if isinstance(sequence, (str, bytes, bytearray, memoryview)):
        return True
    return getattr(sequence, "dtype", None) in (np.uint8, torch.uint8)


def _to_device_sequence(sequence, device):
    """One upload: (B, L, 4) float32 host array / tensor as the reference takes it (16 B/bp), or packed bases
    (str / bytes / uint8 (L,) or (B, L); codes 0..4 or ASCII, 1 B/bp) -> device tensor."""
    if is_packed(sequence):
        if not isinstance(sequence, torch.Tensor):
            sequence = torch.from_numpy(np.ascontiguousarray(feeder.as_bases(sequence)))
        t = sequence.contiguous()
        if t.dim() == 1:
            t = t[None]
        if t.dim() != 2:
            raise ValueError("packed sequence must be (L,) or (B, L), got %s" % (tuple(t.shape),))
        return t.to(device, non_blocking=True)
    if isinstance(sequence, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(sequence, dtype=np.float32))
    else:
        t = sequence.float().contiguous()
    if t.dim() != 3 or t.size(2) != 4:
        raise ValueError("sequence must be (B, L, 4), got %s" % (tuple(t.shape),))
    if t.is_cuda or t.is_pinned() or torch.device(device).type != "cuda" or t.numel() * 4 < (_STAGE_BYTES << 1):
        return t.to(device, non_blocking=True)
    return _upload_pageable(t, device)


_STAGE_BYTES = 32 << 20
_STAGING = {}


def _upload_pageable(t, device):
    """Pageable host tensor -> device through two pinned staging buffers: the (multi-threaded) host copy of chunk
    i+1 overlaps the DMA of chunk i, instead of cudaMemcpy's serial bounce of the whole 0.5-4 GB array (the
    reference hands genomepredict a pageable numpy array, orca_predict.py:334)."""
    device = torch.device(device)
    key = str(device)
    if key not in _STAGING:
        _STAGING[key] = ([torch.empty(_STAGE_BYTES // 4, dtype=torch.float32).pin_memory() for _ in range(2)],
                         [torch.cuda.Event() for _ in range(2)], torch.cuda.Stream(device=device))
    bufs, events, cs = _STAGING[key]
    out = torch.empty(t.shape, dtype=torch.float32, device=device)
    src, dst = t.view(-1), out.view(-1)
    n, step = src.numel(), _STAGE_BYTES // 4
    main = torch.cuda.current_stream(device)
    cs.wait_stream(main)
    for i, lo in enumerate(range(0, n, step)):
        hi = min(lo + step, n)
        b = i & 1
        if i >= 2:
            events[b].synchronize()  # the DMA that last read this staging buffer has finished
        bufs[b][:hi - lo].copy_(src[lo:hi])
        with torch.cuda.stream(cs):
            dst[lo:hi].copy_(bufs[b][:hi - lo], non_blocking=True)
            events[b].record(cs)
    main.wait_stream(cs)
    out.record_stream(main)
    return out


def _log_normmat(model, level, device):
    """log(normmats[level]) on the device (orca_predict.py:349-353 recomputes and re-uploads it on every eval_step).
    Cached per (level, device) and re-validated against the SOURCE array on every call -- same object and same
    checksum -- so that replacing or editing `model.normmats[level]` takes effect as it does in the reference."""
    cache = model.__dict__.setdefault("_distenc_cache", {})
    key = (level, str(device))
    src = model.normmats[level]
    stamp = float(np.sum(src))
    hit = cache.get(key)
    if hit is None or hit[0] is not src or hit[1] != stamp:
        nm = src[(None,) * (4 - src.ndim)]  # orca_predict.py:350
        t = torch.log(torch.as_tensor(nm, dtype=torch.float32).to(device))
        # the cached tensor is shared by cascades running on different streams (run_concurrent): make sure it is
        # complete before any other stream can pick it up (one-time cost per model and level)
        if torch.device(device).type == "cuda":
            torch.cuda.current_stream(device).synchronize()
        cache[key] = hit = (src, stamp, t)
    return hit[2]


_SIDE_STREAMS = {}


def run_concurrent(jobs, device):
    """Run independent GPU jobs (callables returning tensors) on separate CUDA streams and join them on the
    current stream.  The decoder cascades of different (model, strand) pairs are independent chains of ~20 us
    kernels; interleaving two chains hides each launch's prologue/tail latency behind the other's work."""
    if len(jobs) <= 1 or torch.device(device).type != "cuda":
        return [job() for job in jobs]
    main = torch.cuda.current_stream(device)
    pool = _SIDE_STREAMS.setdefault(str(device), [])
    while len(pool) < len(jobs):
        pool.append(torch.cuda.Stream(device=device))
    results = []
    for job, st in zip(jobs, pool):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            results.append(job())
    for st in pool[:len(jobs)]:
        main.wait_stream(st)

    def mark(obj):  # outputs were allocated on a side stream but live on with the main stream
        if isinstance(obj, torch.Tensor):
            obj.record_stream(main)
        elif isinstance(obj, (list, tuple)):
            for o in obj:
                mark(o)
        elif isinstance(obj, dict):
            for o in obj.values():
                mark(o)
    mark(results)
    return results


def encode_strand(model, seq_dev, reverse):
    """net0 on one strand of the uploaded tensor: float32 (B, L, 4) or packed uint8 (B, L)."""
    x = seq_dev if seq_dev.dtype == torch.uint8 else seq_dev.transpose(1, 2)
    if hasattr(model.net0, "fp16_guard_fired"):  # native Encoder: the range guard is checked once per pass (check_fp16_guard)
        return model.net0(x, reverse_complement=reverse, guard=False)
    return model.net0(x, reverse_complement=reverse)


def check_fp16_guard(models):
    """After a pass has been read back (the device is idle anyway): did any native encoder trip its fp16 range guard?
    Such encoders are switched to the fp32-grade format; returns True when the pass has to be repeated."""
    fired = False
    for m in models:
        net0 = getattr(m, "net0", None)
        if hasattr(net0, "fp16_guard_fired") and net0.fp16_guard_fired():
            net0._fall_back_to_fp32_grade()
            fired = True
    return fired


def cascade_starts_32mb(mpos, wpos, reverse):
    """Start bins of the six zoom levels: host integer arithmetic only (orca_predict.py:470-499)."""
    starts = [0]
    for j, level in enumerate([32, 16, 8, 4, 2, 1]):
        starts.append(starts[j] + _next_index_32mb(level, starts[j], mpos, wpos, reverse) * level)
    return starts[:-1]


def level1_extra(model, encodings, mpos, wpos, reverse):
    """Decoder_1m term of the 1 Mb level (orca_predict.py:362-366).  It depends only on the level-1 encoding and
    the (host-computed) start bin, so it can run beside the cascade -- on another stream or another rank."""
    s = int(cascade_starts_32mb(mpos, wpos, reverse)[5])
    return model.denet_1_pt.forward(encodings[1][:, :, s:s + 250])


def cascade_32mb(model, encodings, batch, mpos, wpos, reverse, inline_1m=True):
    """Decoder cascade 32 -> 1 Mb of one (model, strand) (orca_predict.py:348-500).  With inline_1m=False the
    caller adds level1_extra(...) to the last map itself."""
    device = encodings[1].device
    preds, starts = [], [0]
    start_index = 0
    for j, level in enumerate([32, 16, 8, 4, 2, 1]):
        distenc = _log_normmat(model, level, device).expand(batch, -1, -1, -1)
        s = int(starts[j] / level)
        xl = encodings[level][:, :, s:s + 250]
        coarse = None if j == 0 else preds[j - 1][:, :, start_index:start_index + 125, start_index:start_index + 125]
        pred = model.denets[level].forward(xl, distenc, coarse)
        if level == 1 and j > 0 and hasattr(model, "denet_1_pt") and inline_1m:
            pred = pred + model.denet_1_pt.forward(xl)
        start_index = _next_index_32mb(level, starts[j], mpos, wpos, reverse)
        starts.append(starts[j] + start_index * level)
        preds.append(pred)
    return preds, starts[:-1]


def _next_index_32mb(level, start, mpos, wpos, reverse):
    """Crop index of the next zoom level (orca_predict.py:470-497) -- the ONE statement of this formula."""
    if not reverse:
        v = np.floor(((mpos - level * 1000000 / 4) - (wpos - 16000000 + start * 4000)) / (4000 * level))
    else:
        v = np.ceil(((wpos + 16000000 - start * 4000) - (mpos + level * 1000000 / 4)) / (4000 * level))
    return int(np.clip(v, 0, 125))


def _next_index_256mb(level, start, chrlen, mpos, wpos, reverse):
    """Crop index of the next zoom level of the 256 Mb cascade (orca_predict.py:813-835): `chrlen` clipping, and the
    mirrored index 250 - (i + 125) on the reverse strand."""
    if not reverse:
        proposed = (mpos - level * 1000000 / 4) - (wpos - 128000000 + start * 4000 * 8)
    else:
        proposed = (mpos - level * 1000000 / 4) - (wpos + 128000000 - start * 4000 * 8 - level * 1000000)
    if chrlen is not None:
        bounds = [0 - (wpos - 128000000), chrlen - level * 1000000 / 2 - (wpos - 128000000)]
        proposed = np.clip(proposed, bounds[0], bounds[1]) if bounds[0] < bounds[1] else bounds[0]
    i = int(np.clip(np.floor(proposed / (4000 * level)), 0, 125))
    return 250 - (i + 125) if reverse else i


def cascade_32mb_lanes(model, lanes, mpos, wpos, inline_1m=True):
    """Several independent cascades of ONE model (e.g. its two strands) as one batched chain: every decoder call
    carries one batch element per lane, so the 118 dependent layers of a level are paid once instead of per strand
    (each lane keeps its own crop windows; the arithmetic per lane is exactly cascade_32mb's).

    lanes: [(encodings {level: (B, 128, P/level)}, reverse[, mpos, wpos]), ...] -- a lane may carry its own zoom target
    (orca_b200.variants batches the windows of one variant call); otherwise the call's mpos / wpos apply.  Returns
    (preds: list over levels of (n_lanes * B, C, 250, 250), lane-major; starts: per-lane start bins)."""
    n = len(lanes)
    target = [(ln[2], ln[3]) if len(ln) > 2 else (mpos, wpos) for ln in lanes]
    lanes = [(ln[0], ln[1]) for ln in lanes]
    device = lanes[0][0][1].device
    B = lanes[0][0][1].shape[0]
    starts, sidx, preds = [[0] for _ in lanes], [0] * n, []
    for j, level in enumerate([32, 16, 8, 4, 2, 1]):
        distenc = _log_normmat(model, level, device).expand(n * B, -1, -1, -1)
        xl = torch.cat([enc[level][:, :, int(st[j] / level):int(st[j] / level) + 250] for (enc, _), st in zip(lanes, starts)], 0)
        coarse = None if j == 0 else torch.cat([preds[j - 1][i * B:(i + 1) * B, :, si:si + 125, si:si + 125] for i, si in enumerate(sidx)], 0)
        pred = model.denets[level].forward(xl, distenc, coarse)
        if level == 1 and hasattr(model, "denet_1_pt") and inline_1m:
            pred = pred + model.denet_1_pt.forward(xl)
        for i, (_, rev) in enumerate(lanes):
            sidx[i] = _next_index_32mb(level, starts[i][j], target[i][0], target[i][1], rev)
            starts[i].append(starts[i][j] + sidx[i] * level)
        preds.append(pred)
    return preds, [st[:-1] for st in starts]


def _average_strands(fwd, rev):
    """0.5*fwd + 0.5*rev[::-1, ::-1] for batch element 0 (orca_predict.py:510-523)."""
    out = []
    for f, r in zip(fwd, rev):
        a = 0.5 * f[0] + 0.5 * torch.flip(r[0], [1, 2])
        out.append(a[0] if a.shape[0] == 1 else a)
    return out


def genomepredict(sequence, mchr, mpos=-1, wpos=-1, models=(), targets=None, annotation=None, use_cuda=True,
                  nan_thresh=1):
    """Drop-in for orca_predict.genomepredict with native shells (see module docstring)."""
    if targets is not None or annotation is not None:
        raise NotImplementedError("targets / annotation are plot-only inputs; use orca_predict.genomepredict for them")
    if not use_cuda:
        raise RuntimeError("orca_b200 has no CPU path")
    models = list(models)
    if not models or not all(isinstance(m, torch.nn.Module) for m in models):
        raise ValueError("models must be a non-empty list of shell modules")
    device = _device_of(models[0])
    with torch.cuda.device(device):
        return _genomepredict_on(device, sequence, mchr, mpos, wpos, models)


def _strand_lanes_32mb(model, seq_dev, mpos, wpos):
    """One model, both strands: two encoder passes over the one uploaded tensor, ONE U-net call and ONE decoder
    chain at batch 2B (the strands are the lanes).  Returns (per-level strand-averaged maps, forward-strand starts)."""
    B = seq_dev.shape[0]
    enc = torch.cat([encode_strand(model, seq_dev, False), encode_strand(model, seq_dev, True)], 0)
    outs = model.net(enc)
    lanes = [({lvl: t[i * B:(i + 1) * B] for lvl, t in zip([1, 2, 4, 8, 16, 32], outs)}, rev) for i, rev in enumerate((False, True))]
    preds, starts = cascade_32mb_lanes(model, lanes, mpos, wpos)
    return _average_strands([p[:B] for p in preds], [p[B:] for p in preds]), starts[0]


def _staged_runners(device, sequence, models):
    """Single 32 Mb-class sequence held on the HOST + native shells: one cached sharded runner (world size 1) per model, whose
    staged upload overlaps the PCIe transfer (and, for pageable arrays, the pinned staging) with the first encoder chunks.
    Returns None when the call does not qualify (device input, batch > 1, foreign modules ...)."""
    if torch.device(device).type != "cuda" or isinstance(sequence, (str, bytes, bytearray, memoryview)):
        return None
    t = torch.from_numpy(sequence) if isinstance(sequence, np.ndarray) else sequence
    if not isinstance(t, torch.Tensor) or t.is_cuda:
        return None
    if t.dtype == torch.uint8:
        t = t[None] if t.dim() == 1 else t
        if t.dim() != 2:
            return None
    elif t.dtype != torch.float32 or t.dim() != 3 or t.size(2) != 4 or not t.is_contiguous():
        return None
    L = t.shape[1]
    if t.shape[0] != 1 or L % 4000 or L // 4000 < 8000:
        return None
    if not all(hasattr(getattr(m, "net0", None), "fp16_guard_fired") for m in models):
        return None
    from . import parallel
    runners = []
    for m in models:
        cache = m.__dict__.setdefault("_staged_runner", {})
        key = (L, str(device))
        if key not in cache:
            cache[key] = parallel.ShardedForward(m, L, 0, 1, torch.device(device))
        runners.append(cache[key])
    return t, runners


def _genomepredict_on(device, sequence, mchr, mpos, wpos, models):
    """Body of genomepredict for shells living on `device` (the host logic is device-agnostic: tests/test_host.py
    drives it on the CPU with stand-in networks against fixtures from the unmodified reference driver)."""
    staged = _staged_runners(device, sequence, models)
    if staged is not None:
        t, runners = staged
        with torch.no_grad():
            for attempt in range(2):
                runners[0].upload(t)
                for r in runners[1:]:  # every model reads the one uploaded window
                    r.window, r._ready, r._cuts = runners[0].window, None, runners[0]._cuts
                outs = []
                for i, r in enumerate(runners):
                    if i > 0:
                        torch.cuda.current_stream(device).wait_stream(runners[0]._copy_stream)
                    outs.append(r.forward(mpos, wpos))
                host = [o.cpu().numpy() for o in outs]
                runners[0]._join_uploader()
                if not check_fp16_guard(models):
                    break
        starts0 = cascade_starts_32mb(mpos, wpos, False)
        return _output_32mb(host, starts0, mchr, wpos, models)
    with torch.no_grad():
        seq_dev = _to_device_sequence(sequence, device)
        for attempt in range(2):
            results = [_strand_lanes_32mb(model, seq_dev, mpos, wpos) for model in models]
            stacked = [torch.stack(avg) for avg, _ in results]
            starts0 = results[0][1]
            host = [s.cpu().numpy() for s in stacked]
            if not check_fp16_guard(models):  # else: the encoders now run fp32-grade; repeat the pass once
                break
    return _output_32mb(host, starts0, mchr, wpos, models)


def _output_32mb(host, starts0, mchr, wpos, models):
    """The dict orca_predict.genomepredict returns (orca_predict.py:500-540)."""
    output = {"predictions": [[h[j] for j in range(h.shape[0])] for h in host], "experiments": None}
    output["start_coords"] = [wpos - 16000000 + s * 4000 for s in starts0]
    output["end_coords"] = [int(output["start_coords"][ii] + 32000000 / 2 ** ii) for ii in range(6)]
    output["chr"] = mchr
    output["annos"] = None
    output["normmats"] = [[m.normmats[ii] for ii in [32, 16, 8, 4, 2, 1]] for m in models]
    return output


def assemble_background(regionlist, background_cis, background_trans, device, binsize=32000):
    """Background matrix of a multi-region input, built on the device: the normmat branch of
    orca_predict._retrieve_multi (orca_predict.py:936-965) without the host numpy assembly and the 512 MB upload.

    regionlist: [(chrom, start, end[, strand]), ...]; background_cis / background_trans: the shell's 32 kb curve
    (NaN-padded, orca_models.py:626-633) and trans constant.  Returns an (n, n) float64 device tensor that
    genomepredict_256Mb(normmats=[...]) accepts directly."""
    import ctypes
    device = torch.device(device)
    ids = {}
    regs = (_lib.Region * len(regionlist))()
    for r, region in zip(regs, regionlist):
        chrom, start, end = region[:3]
        strand = region[3] if len(region) > 3 else "+"
        r.chrom, r.reverse, r.start, r.end = ids.setdefault(chrom, len(ids)), 1 if strand == "-" else 0, int(start), int(end)
    lib = _lib.lib()
    n = int(lib.orca_b200_background_bins(regs, len(regs), int(binsize)))
    if n < 0:
        _lib.check(n)
    with torch.cuda.device(device):
        cis = torch.as_tensor(np.asarray(background_cis, dtype=np.float64)).to(device)
        out = torch.empty((n, n), dtype=torch.float64, device=device)
        ws = torch.empty(12 * n + 512, dtype=torch.uint8, device=device)
        _lib.check(lib.orca_b200_background_assemble(regs, len(regs), cis.data_ptr(), cis.numel(), float(background_trans),
                                                     int(binsize), out.data_ptr(), n, ws.data_ptr(), ws.numel(),
                                                     ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return out


def prepare_background(normmat, device):
    """Caller's (n, n) float64 background matrix (host array, or a device tensor from assemble_background) -> device,
    NaNs replaced by the minimum (orca_predict.py:668-671)."""
    if isinstance(normmat, torch.Tensor):
        t = normmat.to(device=device, dtype=torch.float64)
    else:
        t = torch.as_tensor(np.asarray(normmat, dtype=np.float64)).to(device)
    nan = torch.isnan(t)
    if bool(nan.any()):
        t = torch.where(nan, t[~nan].min(), t)
    return t


def background_level(normmat_dev, r0, f, flip=False, size=250, with_mean=False):
    """log(block-nanmean) of an (n, n) float64 device matrix -> (1, 1, size, size) float32 (flipped if `flip`);
    with_mean=True also returns the float64 block mean itself, (1, size, size), never flipped (orca_predict.py:724-737)."""
    out = torch.empty((1, 1, size, size), dtype=torch.float32, device=normmat_dev.device)
    mean = torch.empty((1, size, size), dtype=torch.float64, device=normmat_dev.device) if with_mean else None
    with torch.cuda.device(normmat_dev.device):
        _lib.check(_lib.lib().orca_b200_background_level(
            normmat_dev.data_ptr(), normmat_dev.shape[0], int(r0), int(f), size, 1 if flip else 0, out.data_ptr(),
            mean.data_ptr() if with_mean else None, torch.cuda.current_stream(normmat_dev.device).cuda_stream))
    return (out, mean) if with_mean else out


def cascade_256mb(model, encodings, batch, normmat_dev, chrlen, mpos, wpos, reverse):
    """Decoder cascade 256 -> 32 Mb of one (model, strand) (orca_predict.py:692-836).
    Returns (preds, starts, per-level float64 block-mean background matrices (1, 250, 250), device tensors)."""
    preds, starts, ns = [], [0], {}
    start_index = 0
    for j, level in enumerate([256, 128, 64, 32]):
        f = level // 8
        logbg, ns[level] = background_level(normmat_dev, starts[j], f, flip=reverse, with_mean=True)
        distenc = logbg.expand(batch, -1, -1, -1)
        s = int(starts[j] / f)
        xl = encodings[level][:, :, s:s + 250]
        coarse = None if j == 0 else preds[j - 1][:, :, start_index:start_index + 125, start_index:start_index + 125]
        pred = model.denets[level].forward(xl, distenc, coarse)
        start_index = _next_index_256mb(level, starts[j], chrlen, mpos, wpos, reverse)
        starts.append(starts[j] + start_index * level // 8)
        preds.append(pred)
    return preds, starts[:-1], ns


def cascade_256mb_lanes(model, lanes, normmat_dev, chrlen, mpos, wpos):
    """cascade_256mb for several lanes (strands) of one model as one batched chain; see cascade_32mb_lanes.
    Returns (preds: list over levels of (n_lanes, C, 250, 250), per-lane starts)."""
    n = len(lanes)
    starts, sidx, preds = [[0] for _ in lanes], [0] * n, []
    for j, level in enumerate([256, 128, 64, 32]):
        f = level // 8
        dist, xs = [], []
        for (enc, rev), st in zip(lanes, starts):
            dist.append(background_level(normmat_dev, st[j], f, flip=rev)[0])
            s = int(st[j] / f)
            xs.append(enc[level][:, :, s:s + 250])
        coarse = None if j == 0 else torch.stack([preds[j - 1][i, :, si:si + 125, si:si + 125] for i, si in enumerate(sidx)], 0)
        pred = model.denets[level].forward(torch.cat(xs, 0), torch.stack(dist, 0), coarse)
        for i, (_, rev) in enumerate(lanes):
            sidx[i] = _next_index_256mb(level, starts[i][j], chrlen, mpos, wpos, rev)
            starts[i].append(starts[i][j] + sidx[i] * level // 8)
        preds.append(pred)
    return preds, [st[:-1] for st in starts]


def genomepredict_256Mb(sequence, mchr, normmats, chrlen, mpos=-1, wpos=-1, models=(), targets=None, annotation=None,
                        padding_chr=None, use_cuda=True, nan_thresh=1):
    """Drop-in for orca_predict.genomepredict_256Mb with native shells."""
    if targets is not None or annotation is not None:
        raise NotImplementedError("targets / annotation are plot-only inputs; use orca_predict.genomepredict_256Mb")
    if not use_cuda:
        raise RuntimeError("orca_b200 has no CPU path")
    models = list(models)
    device = _device_of(models[0])
    with torch.cuda.device(device):
        return _genomepredict_256mb_on(device, sequence, mchr, normmats, chrlen, mpos, wpos, models, padding_chr)


def _genomepredict_256mb_on(device, sequence, mchr, normmats, chrlen, mpos, wpos, models, padding_chr=None):
    with torch.no_grad():
        seq_dev = _to_device_sequence(sequence, device)
        B = seq_dev.shape[0]
        nm_dev = [prepare_background(nm, device) for nm in normmats]
        for attempt in range(2):
            per_strand, allns, starts0 = [], [], None
            for reverse in (False, True):
                for ii, model in enumerate(models):
                    enc4k = encode_strand(model, seq_dev, reverse)
                    enc128k = model.net1(enc4k, coarsest_only=True)[-1]
                    encs = dict(zip([32, 64, 128, 256], model.net(enc128k)))
                    preds, starts, ns = cascade_256mb(model, encs, B, nm_dev[ii], chrlen, mpos, wpos, reverse)
                    per_strand.append(preds)
                    allns.append(ns)
                    if not reverse and starts0 is None:
                        starts0 = starts
            n = len(models)
            host = [torch.stack(_average_strands(per_strand[i], per_strand[i + n])).cpu().numpy() for i in range(n)]
            ns_host = [{lvl: t.cpu().numpy() for lvl, t in ns.items()} for ns in allns]
            if not check_fp16_guard(models):
                break
    output = {"predictions": [[h[j] for j in range(h.shape[0])] for h in host], "experiments": None}
    output["start_coords"] = [wpos - 128000000 + s * 32000 for s in starts0]
    output["end_coords"] = [np.fmin(int(output["start_coords"][ii] + 256000000 / 2 ** ii), chrlen) for ii in range(4)]
    output["annos"] = None
    output["chr"] = mchr
    output["padding_chr"] = padding_chr
    output["normmats"] = ns_host
    return output
