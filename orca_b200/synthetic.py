"""Seeded synthetic weights, inputs and background matrices.

The reference's trained weights and resources are not distributable with the repo
(orca_models.py:53-58 loads them from a Zenodo download), so benchmarks, smoke tests and
parity tests use random-init networks.  Everything here is generated with numpy's PCG64
(`default_rng`), which is stable across numpy/torch versions and machines, so the same
seed gives bit-identical tensors in this container and on the GPU box.
"""
import numpy as np
import torch


def fill_state_dict(shapes, seed, recipe="default"):
    """Return {key: tensor} for an ordered mapping key -> shape/dtype template.

    recipe: "default" (below); "wide_bn" = BatchNorm folded scale gamma/sqrt(var) log-uniform in [0.1, 10] per
    channel (trained checkpoints can carry such scales; they stress the fp16 range of the single-pass encoder
    stages); "heavy_tail" = conv weights from a Student-t (3 d.o.f.) instead of a uniform distribution.

    `shapes` is any ordered mapping whose values have `.shape` and `.dtype`
    (e.g. `module.state_dict()`).  Recipe: conv weight/bias
    U(-1/sqrt(fan_in), +1/sqrt(fan_in)) (PyTorch's default bound); BatchNorm gamma U(0.5,1.5),
    beta N(0,0.1), running_mean N(0,0.1), running_var U(0.5,1.5) -- randomised statistics so
    that BN-fold mistakes cannot hide behind an identity BatchNorm.
    """
    rng = np.random.default_rng(seed)
    out = {}
    keys = list(shapes.keys())
    # first pass: remember which prefixes are BatchNorms (they own a running_mean)
    bn_prefixes = {k[: -len(".running_mean")] for k in keys if k.endswith(".running_mean")}
    fan_in = {}
    for k in keys:
        t = shapes[k]
        shape = tuple(t.shape)
        prefix, _, leaf = k.rpartition(".")
        if leaf == "num_batches_tracked":
            out[k] = torch.zeros(shape, dtype=torch.int64)
            continue
        if prefix in bn_prefixes:
            if leaf == "weight":
                v = rng.uniform(0.5, 1.5, size=shape)
                if recipe == "wide_bn":
                    v = np.exp(rng.uniform(np.log(0.1), np.log(10.0), size=shape))  # x 1/sqrt(var ~ U(0.5,1.5)) below
            elif leaf == "bias":
                v = rng.normal(0.0, 0.1, size=shape)
            elif leaf == "running_mean":
                v = rng.normal(0.0, 0.1, size=shape)
            elif leaf == "running_var":
                v = rng.uniform(0.5, 1.5, size=shape)
            else:
                raise KeyError(k)
        else:
            if leaf == "weight":
                fan_in[prefix] = int(np.prod(shape[1:]))
                bound = 1.0 / np.sqrt(fan_in[prefix])
            elif leaf == "bias":
                bound = 1.0 / np.sqrt(fan_in[prefix])
            else:
                raise KeyError(k)
            v = rng.uniform(-bound, bound, size=shape)
            if recipe == "heavy_tail" and leaf == "weight":
                v = rng.standard_t(3, size=shape) * bound / np.sqrt(3.0)  # same variance as U(-bound, bound) x 3 d.o.f. tail
        out[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return out


def init_module(module, seed):
    """Load seeded synthetic weights into an nn.Module (reference class or orca_b200 mirror)."""
    sd = fill_state_dict(module.state_dict(), seed)
    module.load_state_dict(sd, strict=True)
    module.eval()
    return module


def random_sequence(batch, length, seed, n_fraction=0.0):
    """(B, L, 4) float32 one-hot in ACGT order; a fraction of positions may be 'N' (0.25 x 4),
    the encoding selene_utils2.MemmapGenome uses for unknown bases (selene_utils2.py:216-230)."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, 4, size=(batch, length))
    seq = np.zeros((batch, length, 4), dtype=np.float32)
    np.put_along_axis(seq, idx[..., None], 1.0, axis=2)
    if n_fraction > 0:
        mask = rng.random((batch, length)) < n_fraction
        seq[mask] = 0.25
    return seq


def hard_sequence(batch, length, seed, kind):
    """Inputs that stress the encoder's early stages: "alln" (every base unknown: 0.25 x 4, selene_utils2.py:216-230),
    "homopolymer" (poly-A with a poly-T block), "nruns" (random bases with 10 % of the positions inside long N runs),
    "random"."""
    rng = np.random.default_rng(seed)
    seq = random_sequence(batch, length, seed)
    if kind == "alln":
        seq[:] = 0.25
    elif kind == "homopolymer":
        seq[:] = 0.0
        seq[:, :, 0] = 1.0
        seq[:, length // 3: length // 2] = [0.0, 0.0, 0.0, 1.0]
    elif kind == "nruns":
        n_runs = max(1, length // 20000)
        for b in range(batch):
            for s0 in rng.integers(0, length - 2000, size=n_runs):
                seq[b, s0:s0 + 2000] = 0.25
    elif kind != "random":
        raise ValueError(kind)
    return seq


def random_codes(batch, length, seed, n_fraction=0.0):
    """The same sequence as random_sequence(batch, length, seed, n_fraction) as packed codes (B, L) uint8:
    0..3 = A, C, G, T, 4 = N (orca_b200.feeder); feeder.to_onehot(random_codes(...)) == random_sequence(...)."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=(batch, length)).astype(np.uint8)
    if n_fraction > 0:
        codes[rng.random((batch, length)) < n_fraction] = 4
    return codes


def expected_log(n=8000):
    """Synthetic log expected-contact curve e[d] = -0.8 ln(d + 1) - 3."""
    return -0.8 * np.log(np.arange(n, dtype=np.float64) + 1.0) - 3.0


def normmats_32mb(elog=None):
    """Per-level 250x250 background matrices exactly as orca_models.py:135-166 builds them."""
    elog = expected_log(8000) if elog is None else elog
    d = np.abs(np.arange(8000)[None, :] - np.arange(8000)[:, None])
    normmat = np.exp(elog[d])
    mats, epss = {}, {}
    for level in (1, 2, 4, 8, 16, 32):
        n = 250 * level
        r = np.reshape(normmat[:n, :n], (250, level, 250, level)).mean(axis=1).mean(axis=2)
        mats[level], epss[level] = r, np.min(r)
    return mats, epss


def normmat_256mb(chrlen_bins=8000, trans=np.exp(-9.0)):
    """8000x8000 caller-side background at 32 kb bins for genomepredict_256Mb: cis curve for
    the first `chrlen_bins` bins, `background_trans` elsewhere (mimics orca_predict.py:951-965)."""
    elog = -0.8 * np.log(np.arange(8000, dtype=np.float64) + 1.0) - 3.0
    d = np.abs(np.arange(8000)[None, :] - np.arange(8000)[:, None])
    nm = np.exp(elog[d])
    if chrlen_bins < 8000:
        nm[chrlen_bins:, :] = trans
        nm[:, chrlen_bins:] = trans
    return nm
