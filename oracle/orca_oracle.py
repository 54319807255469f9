"""
ORACLE -- test infrastructure only.  A CPU fp32 restatement of the Orca forward path in plain
``torch.nn.functional`` calls, operating directly on a reference-format ``state_dict``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s comparator legs -- cpu_baseline / ``--impl reference``
(this graph on the host CPU) and ``gpu_reference`` (the same stock-operator graph on CUDA tensors, i.e. the reference's own
eager PyTorch + cuDNN GPU path, timed beside ours) -- may import this module, and only as the checker / the baseline;
nothing under ``orca_b200/`` imports it and the product path never falls back to it.

Why a restatement: the reference (pure Python, /root/reference/orca_modules.py) is importable in
the build container but does not exist on the GPU box.  The arithmetic itself lives in a
third-party dependency, PyTorch (unpinned by the reference, README.md:41; here torch 2.11.0 CPU /
oneDNN), so each function below re-expresses one reference ``forward`` with the same torch
operators and cites the lines it follows.

Pinning: the reference ships no tests, golden vectors or fixtures, so the oracle
is pinned against outputs of the reference itself: ``oracle/make_golden.py`` imports the unmodified
reference classes from /root/reference in the build container, runs them on seeded inputs and
commits the outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this file against
those fixtures everywhere, and against the live reference classes whenever /root/reference exists.
"""
import numpy as np
import torch
import torch.nn.functional as F

BLOCKSIZE = 4000 * 200  # orca_modules.py:13


def _conv1(x, sd, p):
    return F.conv1d(x, sd[p + ".weight"], sd[p + ".bias"], padding=4)


def _conv2(x, sd, p, d=1, k=3):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=d if k == 3 else 0, dilation=d)


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _lin1d(x, sd, p, o):
    """Conv BN Conv BN with the first conv at Sequential index `o` (1 when a pool/upsample sits at 0)."""
    x = _bn(_conv1(x, sd, "%s.%d" % (p, o)), sd, "%s.%d" % (p, o + 1))
    return _bn(_conv1(x, sd, "%s.%d" % (p, o + 2)), sd, "%s.%d" % (p, o + 3))


def _relu1d(x, sd, p, last_bn=True):
    """Conv BN ReLU Conv [BN] ReLU."""
    x = F.relu(_bn(_conv1(x, sd, p + ".0"), sd, p + ".1"))
    x = _conv1(x, sd, p + ".3")
    if last_bn:
        x = _bn(x, sd, p + ".4")
    return F.relu(x)


# ------------------------------------------------------------------------------------------------
def encoder_run(sd, x, prefix=""):
    """Encoder.forward's inner `run` (orca_modules.py:935-950)."""
    pools = [None, 4, 4, 5, 5, 5, 2]  # :811-927
    cur = x
    out = None
    for k in range(1, 8):
        if pools[k - 1] is not None:
            cur = F.max_pool1d(cur, pools[k - 1], pools[k - 1])
        lout = _lin1d(cur, sd, "%slconv%d" % (prefix, k), 0 if k == 1 else 1)
        out = _relu1d(lout, sd, "%sconv%d" % (prefix, k))
        cur = out + lout
    return out  # out7 alone (:949-950)


def encoder_forward(sd, x, blocksize=BLOCKSIZE):
    """Encoder.forward (orca_modules.py:929-980): 800 kb blocks, 112 kb overlap, trim 28 bins."""
    binsize, x_padding = 4000, 112000
    segs = []
    starts = np.arange(0, x.size(2), blocksize)
    for start in starts:
        start = int(start)
        if start == starts[0]:
            segs.append(encoder_run(sd, x[:, :, start:start + blocksize + x_padding])[:, :, : blocksize // binsize])
        elif start == starts[-1]:
            segs.append(encoder_run(sd, x[:, :, start - x_padding:])[:, :, x_padding // binsize:])
        else:
            segs.append(encoder_run(sd, x[:, :, start - x_padding:start + blocksize + x_padding])
                        [:, :, x_padding // binsize:(blocksize + x_padding) // binsize])
    return torch.cat(segs, 2)


def encoder2_forward(sd, x, n=5, up=True):
    """Encoder2 (n=5), Encoder3 (n=3) and, with up=False, Encoder2b
    (orca_modules.py:1151-1169, :1388-1406, :1266-1276)."""
    out = x
    encodings = [out]
    for i in range(n):
        lout = _lin1d(F.max_pool1d(out, 2, 2), sd, "lblocks.%d" % i, 1)
        out = _relu1d(lout, sd, "blocks.%d" % i) + lout
        encodings.append(out)
    if not up:
        return encodings
    encodings2 = [out]
    for j, enc in enumerate(reversed(encodings[:-1])):
        lout = _lin1d(F.interpolate(out, scale_factor=2), sd, "downlblocks.%d" % j, 1)  # nn.Upsample, nearest
        out = _relu1d(lout, sd, "downblocks.%d" % j, last_bn=False) + lout  # no BN after 2nd conv (:1115-1120)
        out = enc + out
        encodings2.append(out)
    encodings2.reverse()
    return encodings2


def _lin2d(x, sd, p, o, d):
    x = _bn(_conv2(x, sd, "%s.%d" % (p, o), d), sd, "%s.%d" % (p, o + 1))
    return _bn(_conv2(x, sd, "%s.%d" % (p, o + 2), d), sd, "%s.%d" % (p, o + 3))


def _relu2d(x, sd, p, d):
    x = F.relu(_bn(_conv2(x, sd, p + ".0", d), sd, p + ".1"))
    return F.relu(_bn(_conv2(x, sd, p + ".3", d), sd, p + ".4"))


def _final(cur, sd, prefix=""):
    cur = F.relu(_bn(_conv2(cur, sd, prefix + "final.0", 1, k=1), sd, prefix + "final.1"))
    cur = _conv2(cur, sd, prefix + "final.3", 1, k=1)
    return 0.5 * cur + 0.5 * cur.transpose(2, 3)


DECODER_DILATIONS = [1, 2, 4, 8, 16, 32, 64] * 4
DECODER_1M_DILATIONS = [1, 2, 4, 8, 16, 32, 64] + [2, 4, 8, 16, 32, 64] * 2


def decoder_forward(sd, x, distenc, y=None, upsample_mode="nearest"):
    """Decoder.forward (orca_modules.py:461-488)."""
    mat = x[:, :, :, None] + x[:, :, None, :]
    mat = torch.cat([mat, distenc], dim=1)
    mat = _lin2d(mat, sd, "lcombinerD", 0, 1)
    mat = _relu2d(mat, sd, "combinerD", 1) + mat
    if y is not None:
        if upsample_mode == "bilinear":
            up = F.interpolate(y, scale_factor=(2, 2), mode="bilinear", align_corners=False)
        else:
            up = F.interpolate(y, scale_factor=(2, 2), mode="nearest")
        mat = torch.cat([mat, up], dim=1)
    cur = mat
    for i, d in enumerate(DECODER_DILATIONS):
        if i == 0:
            if y is not None:
                cur = _lin2d(cur, sd, "lcombiner", 1, 1)       # Dropout at index 0
                cur = _relu2d(cur, sd, "combiner", 1) + cur
            else:
                cur = _lin2d(cur, sd, "lconvtwos.0", 1, d)     # Dropout at index 0; no residual
                cur = _relu2d(cur, sd, "convtwos.0", d) + cur
        else:
            cur = _lin2d(cur, sd, "lconvtwos.%d" % i, 0, d) + cur
            cur = _relu2d(cur, sd, "convtwos.%d" % i, d) + cur
    return _final(cur, sd)


def decoder_1m_body(sd, mat, prefix=""):
    """Shared by Decoder_1m.forward (:782-800) and Net.forward's run1..run3 (:1866-1890)."""
    cur = mat
    for i, d in enumerate(DECODER_1M_DILATIONS):
        if i == 0:
            cur = _lin2d(cur, sd, prefix + "lconvtwos.0", 1, d)  # 128 -> 32 -> 64, shape changes: no residual
        else:
            cur = _lin2d(cur, sd, prefix + "lconvtwos.%d" % i, 0, d) + cur
        cur = _relu2d(cur, sd, prefix + "convtwos.%d" % i, d) + cur
    return _final(cur, sd, prefix)


def decoder_1m_forward(sd, x):
    return decoder_1m_body(sd, x[:, :, :, None] + x[:, :, None, :])


def net_forward(sd, x, num_1d=None):
    """Net.forward (orca_modules.py:1833-1900)."""
    out7 = encoder_run(sd, x)
    pred = decoder_1m_body(sd, out7[:, :, :, None] + out7[:, :, None, :])
    if num_1d:
        h = F.relu(_bn(F.conv1d(out7, sd["final_1d.0.weight"], sd["final_1d.0.bias"]), sd, "final_1d.1"))
        return pred, torch.sigmoid(F.conv1d(h, sd["final_1d.3.weight"], sd["final_1d.3.bias"]))
    return pred


def background_level(normmat, r0, f, size=250, flip=False):
    """orca_predict.py:724-737 (block nanmean, inner axis first) then :693-697 (float32 log), :703 (flip)."""
    blk = normmat[r0:r0 + size * f, r0:r0 + size * f]
    r = np.nanmean(np.nanmean(np.reshape(blk, (1, size, f, size, f)), axis=4), axis=2)
    d = torch.log(torch.FloatTensor(r[None, :, :]))
    return torch.flip(d, [2, 3]) if flip else d


def assemble_background(regionlist, background_cis, background_trans, binsize=32000):
    """Normmat branch of orca_predict._retrieve_multi (orca_predict.py:936-965): the background matrix of a
    multi-region input.  regionlist: [(chrom, start, end, strand), ...]."""
    rows = []
    for chrom, start, end, strand in regionlist:
        b = []
        for chrom2, start2, end2, strand2 in regionlist:
            if chrom2 != chrom:
                b.append(np.full((int((end - start) / binsize), int((end2 - start2) / binsize)), background_trans))
            else:
                acoor = np.linspace(start, end, int((end - start) / binsize) + 1)[:-1]
                bcoor = np.linspace(start2, end2, int((end2 - start2) / binsize) + 1)[:-1]
                blk = background_cis[(np.abs(acoor[:, None] - bcoor[None, :]) / binsize).astype(int)]
                if strand == "-":
                    blk = blk[::-1, :]
                if strand2 == "-":
                    blk = blk[:, ::-1]
                b.append(blk)
        rows.append(b)
    return np.vstack([np.hstack(l) for l in rows])


# ------------------------------------------------------------------------------------------------
# Shell-level restatement used to check whole genomepredict passes without the reference tree.
# ------------------------------------------------------------------------------------------------
def strip_prefix(sd, prefix="module."):
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in sd.items()}


def cascade_32mb(sds, seq_bl4, normmats, mpos, wpos, upsample_mode="bilinear", reverse=False):
    """One (model, strand) pass of genomepredict (orca_predict.py:324-500) on CPU.

    sds: dict with state_dicts 'net0', 'net', 'denet_1_pt' and 'd1'..'d32'.
    Returns (list of 6 prediction tensors (B,1,250,250), list of start bins)."""
    with torch.no_grad():
        seq = torch.from_numpy(np.ascontiguousarray(seq_bl4[:, ::-1, ::-1]) if reverse else seq_bl4)
        enc0 = encoder_forward(sds["net0"], seq.transpose(1, 2))
        encs = dict(zip([1, 2, 4, 8, 16, 32], encoder2_forward(sds["net"], enc0)))
        preds, starts = [], [0]
        start_index = 0
        for j, level in enumerate([32, 16, 8, 4, 2, 1]):
            distenc = torch.log(torch.FloatTensor(normmats[level][None, None])).expand(seq.shape[0], -1, -1, -1)
            s = int(starts[j] / level)
            xl = encs[level][:, :, s:s + 250]
            coarse = None if j == 0 else preds[j - 1][:, :, start_index:start_index + 125, start_index:start_index + 125]
            pred = decoder_forward(sds["d%d" % level], xl, distenc, coarse, upsample_mode)
            if level == 1:
                pred = pred + decoder_1m_forward(sds["denet_1_pt"], xl)
            if not reverse:  # :470-497
                start_index = int(np.clip(np.floor(((mpos - level * 1000000 / 4) - (wpos - 16000000 + starts[j] * 4000))
                                                   / (4000 * level)), 0, 125))
            else:
                start_index = int(np.clip(np.ceil(((wpos + 16000000 - starts[j] * 4000) - (mpos + level * 1000000 / 4))
                                                  / (4000 * level)), 0, 125))
            starts.append(starts[j] + start_index * level)
            preds.append(pred)
        return preds, starts[:-1]
