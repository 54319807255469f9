"""
Model shells: the attribute protocol `orca_predict.genomepredict*` drives
(`.net0 .net [.net1] .denets{level} [.denet_1_pt] .normmats .epss [.background_cis/.background_trans]`),
mirroring /root/reference/orca_models.py (H1esc :17-175, Hff :178-333, HCTnoc :335-446,
H1esc_1M :449-494, Hff_1M :497-542, H1esc_256M :545-649, Hff_256M :652-760) and the two multi-dataset
shells of /root/reference/orca_leukemia.py (OrcaLeukemiaA :1604-1733, OrcaLeukemiaB :1736-1876).

The reference shells load trained weights from `ORCA_PATH/models/*.statedict` and background
curves from `ORCA_PATH/resources/*.npy` (a Zenodo download, not in the repository).  Here:
  * `weights="synthetic"` (default when no `orca_path` is given) builds seeded random-init
    networks and the synthetic background of SURVEY.md 8d;
  * `orca_path=...` loads the reference files with the reference's own key filtering
    (orca_models.py:103-123: net0 / denet_1_pt are carved out of the stage-a file by key,
    with a "module." prefix).

`build_shell(classes, ...)` is class-agnostic: pass `orca_b200.modules` for the native
CUDA modules or the reference's `orca_modules` to get the same shell on stock torch.nn
(used by oracle/make_golden.py and the CPU baseline).
"""
import os

import numpy as np
import torch
from torch import nn

from . import synthetic

LEVELS_32M = (1, 2, 4, 8, 16, 32)
LEVELS_256M = (32, 64, 128, 256)

# seed offsets inside a shell, so every sub-network gets its own reproducible stream
_SEED_NET0, _SEED_NET, _SEED_NET1, _SEED_D1PT, _SEED_DEC = 0, 1, 2, 3, 10


class Shell(nn.Module):
    """Container with the reference attribute protocol; sub-networks are registered modules so
    `.cuda()`, `.eval()` and `.state_dict()` behave as for the reference shells."""

    def __init__(self, kind):
        super().__init__()
        self.kind = kind
        self.normmats = {}
        self.epss = {}

    @property
    def denets(self):
        return {int(k[len("denet_"):]): m for k, m in self._modules.items()
                if k.startswith("denet_") and k != "denet_1_pt"}

    def forward(self, x):  # H1esc_1M.forward / Hff_1M.forward  (orca_models.py:491-494)
        if self.kind not in ("h1esc_1m", "hff_1m"):
            raise RuntimeError("only the 1 Mb shells are callable; use orca_predict.genomepredict* for the others")
        out = self.net.forward(x)
        return out[0] if isinstance(out, tuple) else out


def _strip(sd, prefix="module."):
    """Drop every leading "module." (nn.DataParallel / AveragedModel wrappers): the reference's stage-a file carries
    the prefix twice (orca_models.py:111 indexes it with "module." + a DataParallel key), the others once."""
    out = {}
    for k, v in sd.items():
        while k.startswith(prefix):
            k = k[len(prefix):]
        out[k] = v
    return out


def _load_file(path):
    return torch.load(path, map_location=torch.device("cpu"))


def _init(module, seed, orca_path, fname, filter_keys=False):
    if orca_path is None:
        return synthetic.init_module(module, seed)
    sd = _strip(_load_file(os.path.join(orca_path, "models", fname)))
    if filter_keys:  # orca_models.py:111, :120-122
        sd = {k: sd[k] for k in module.state_dict()}
    module.load_state_dict(sd, strict=True)
    module.eval()
    return module


def build_shell(classes, kind="h1esc", seed=0, orca_path=None):
    """Build a shell of `kind` from network classes `classes` (a module/namespace exposing
    Encoder, Encoder2, Encoder2b, Encoder3, Decoder, Decoder_1m, Net)."""
    kind = kind.lower()
    base = seed * 1000
    sh = Shell(kind)
    cell = kind.split("_")[0]  # h1esc / hff / hctnoc
    if kind in ("h1esc", "hff", "hctnoc"):
        mode = "nearest" if kind == "hctnoc" else "bilinear"  # orca_models.py:45 vs :364
        sh.net0 = _init(classes.Encoder(), base + _SEED_NET0, orca_path, "orca_%s.net0.statedict" % cell, True)
        net_cls = classes.Encoder2b if kind == "hctnoc" else classes.Encoder2
        sh.net = _init(net_cls(), base + _SEED_NET, orca_path, "orca_%s.net.statedict" % cell)
        for i, level in enumerate(LEVELS_32M):
            setattr(sh, "denet_%d" % level,
                    _init(classes.Decoder(upsample_mode=mode), base + _SEED_DEC + i, orca_path,
                          "orca_%s.d%d.statedict" % (cell, level)))
        if kind != "hctnoc":
            sh.denet_1_pt = _init(classes.Decoder_1m(), base + _SEED_D1PT, orca_path,
                                  "orca_%s.net0.statedict" % cell, True)
        sh.normmats, sh.epss = _background_32mb(orca_path, cell)
    elif kind in ("h1esc_1m", "hff_1m"):
        num_1d = 32 if cell == "h1esc" else 22  # orca_models.py:468 (H1esc_1M) / :516 (Hff_1M)
        sh.net = _init(classes.Net(num_1d=num_1d), base + _SEED_NET0, orca_path, "orca_%s.net0.statedict" % cell, True)
        mats, epss = _background_32mb(orca_path, cell, res1000=True)
        sh.normmats, sh.epss = mats, epss
    elif kind in ("h1esc_256m", "hff_256m"):
        sh.net0 = _init(classes.Encoder(), base + _SEED_NET0, orca_path, "orca_%s.net0.statedict" % cell, True)
        sh.net1 = _init(classes.Encoder2(), base + _SEED_NET1, orca_path, "orca_%s.net.statedict" % cell)
        sh.net = _init(classes.Encoder3(), base + _SEED_NET, orca_path, "orca_%s_256m.net.statedict" % cell)
        for i, level in enumerate(LEVELS_256M):
            setattr(sh, "denet_%d" % level,
                    _init(classes.Decoder(upsample_mode="bilinear"), base + _SEED_DEC + i, orca_path,
                          "orca_%s_256m.d%d.statedict" % (cell, level)))
        sh.background_cis, sh.background_trans = _background_256mb(orca_path, cell)
    elif kind in ("leukemia_a", "leukemia_b"):
        # orca_leukemia.py:1604-1876 (OrcaLeukemiaA: 2 datasets, OrcaLeukemiaB: 6); `classes` must expose the
        # orca_leukemia signatures (orca_b200.leukemia, or the reference's orca_leukemia module)
        n_files = 2 if kind == "leukemia_a" else 6
        stem = "orca_leukemia" + kind[-1].upper()
        sh.net0 = _init(classes.Encoder(), base + _SEED_NET0, orca_path, stem + ".net0.statedict", True)
        sh.net = _init(classes.Encoder2(), base + _SEED_NET, orca_path, stem + ".net.statedict")
        for i, level in enumerate(LEVELS_32M):
            setattr(sh, "denet_%d" % level,
                    _init(classes.Decoder(n_files), base + _SEED_DEC + i, orca_path, "%s.d%d.statedict" % (stem, level)))
        sh.denet_1_pt = _init(classes.Decoder_1m(n_files), base + _SEED_D1PT, orca_path, stem + ".net0.statedict", True)
        sh.normmats, sh.epss = _background_leukemia(orca_path, kind, n_files)
    else:
        raise ValueError("unknown shell kind %r" % kind)
    sh.eval()
    return sh


# resource file stems (orca_models.py:136, :295, :408)
_RES = {"h1esc": "4DNFI9GMP2J8", "hff": "4DNFI643OYP9", "hctnoc": "4DNFILP99QJS.HCT_auxin6h"}


def _background_32mb(orca_path, cell, res1000=False):
    if res1000:  # orca_models.py:478-488: 1 kb curve, first 1000 bins, 4x4 block mean
        if orca_path is None:
            elog = synthetic.expected_log(1000)
        else:
            elog = np.load(os.path.join(orca_path, "resources", _RES[cell] + ".rebinned.mcool.expected.res1000.npy"))[:1000]
        d = np.abs(np.arange(1000)[None, :] - np.arange(1000)[:, None])
        r = np.reshape(np.exp(elog[d]), (250, 4, 250, 4)).mean(axis=1).mean(axis=2)
        return {1: r}, {1: np.min(r)}
    elog = None
    if orca_path is not None:
        elog = np.load(os.path.join(orca_path, "resources", _RES[cell] + ".rebinned.mcool.expected.res4000.npy"))
    return synthetic.normmats_32mb(elog)


_LEUKEMIA_RES = {  # orca_leukemia.py:1631-1632, :1763-1768
    "leukemia_a": ["GSE134761_TALL_all.hg38.no_filter.1000.mcool.expected.res4000.npy",
                   "THP1.hg38.no_filter.1000.mcool.expected.res4000.npy"],
    "leukemia_b": ["4DNFIXP4QG5B.mcool.rebinned.mcool.expected.res4000.npy",
                   "NALM6.hg38.no_filter.1000.mcool.expected.res4000.npy",
                   "GSE146901_T_ALL_NonETP.hg38.no_filter.1000.mcool.expected.res4000.npy",
                   "GSE146901_T_ALL_ETP.hg38.no_filter.1000.mcool.expected.res4000.npy",
                   "GSE63525_K562.hg38.no_filter.1000.mcool.expected.res4000.npy",
                   "GSE63525_KBM7.hg38.no_filter.1000.mcool.expected.res4000.npy"],
}


def _background_leukemia(orca_path, kind, n_files):
    """(n_files, 250, 250) block-mean backgrounds per level (orca_leukemia.py:1636-1643, :1706-1720)."""
    if orca_path is None:
        # synthetic: one power-law curve per dataset with slightly different exponents
        elogs = [-(0.8 + 0.05 * i) * np.log(np.arange(8000, dtype=np.float64) + 1.0) - 3.0 for i in range(n_files)]
    else:
        elogs = [np.load(os.path.join(orca_path, "resources", f))[:8000] for f in _LEUKEMIA_RES[kind]]
    per = [synthetic.normmats_32mb(e) for e in elogs]
    mats = {lvl: np.stack([m[lvl] for m, _ in per], axis=0) for lvl in LEVELS_32M}
    return mats, {lvl: np.min(mats[lvl]) for lvl in LEVELS_32M}


def _background_256mb(orca_path, cell):
    # orca_models.py:626-633: exp(mono curve) padded with 2000 NaNs; exp(trans scalar)
    if orca_path is None:
        cis = np.exp(-0.8 * np.log(np.arange(8000, dtype=np.float64) + 1.0) - 3.0)
        trans = np.exp(-9.0)
    else:
        cis = np.exp(np.load(os.path.join(orca_path, "resources", _RES[cell] + ".rebinned.mcool.expected.res32000.mono.npy")))
        trans = np.exp(np.load(os.path.join(orca_path, "resources", _RES[cell] + ".rebinned.mcool.expected.res32000.trans.npy")))
    return np.hstack([cis, np.repeat(np.nan, 2000)]), trans


def _native(kind, seed, orca_path, device):
    if kind.startswith("leukemia"):
        from . import leukemia as modules
    else:
        from . import modules
    sh = build_shell(modules, kind, seed, orca_path)
    return sh.to(device) if device is not None else sh


def H1esc(seed=0, orca_path=None, device="cuda"):
    return _native("h1esc", seed, orca_path, device)


def Hff(seed=1, orca_path=None, device="cuda"):
    return _native("hff", seed, orca_path, device)


def HCTnoc(seed=2, orca_path=None, device="cuda"):
    return _native("hctnoc", seed, orca_path, device)


def H1esc_1M(seed=0, orca_path=None, device="cuda"):
    return _native("h1esc_1m", seed, orca_path, device)


def Hff_1M(seed=1, orca_path=None, device="cuda"):
    return _native("hff_1m", seed, orca_path, device)


def H1esc_256M(seed=0, orca_path=None, device="cuda"):
    return _native("h1esc_256m", seed, orca_path, device)


def Hff_256M(seed=1, orca_path=None, device="cuda"):
    return _native("hff_256m", seed, orca_path, device)


def OrcaLeukemiaA(seed=4, orca_path=None, device="cuda"):
    return _native("leukemia_a", seed, orca_path, device)


def OrcaLeukemiaB(seed=5, orca_path=None, device="cuda"):
    return _native("leukemia_b", seed, orca_path, device)
