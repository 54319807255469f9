"""Pin the model shells (orca_b200.models) against the reference's OWN shell constructors (orca_models.py:17-760).

A fake ORCA_PATH tree (seeded synthetic `.statedict` files and background curves, named exactly as the reference
names them) is written to a temporary directory; the unmodified reference constructors and
`orca_b200.models.build_shell(..., orca_path=tree)` are then both pointed at it.  Compared: every weight tensor of
every sub-network, `normmats`, `epss` (row a10: orca_models.py:135-166, :297-324, :411-438, :478-488) and the 256 Mb
background curve (:626-633).  This also checks the resource / statedict file names of every shell kind
(the HCTnoc stem) because the reference constructors open the files by their own names.

Needs /root/reference (build container only); CPU only."""
import os
import sys
import types

import numpy as np
import pytest
import torch
from torch import nn

from conftest import REFERENCE
from orca_b200 import models, modules, synthetic

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present")


def _ref():
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import orca_models
    import orca_modules
    return orca_models, orca_modules


def _save(path, module, seed, prefix):
    sd = synthetic.fill_state_dict(module.state_dict(), seed)
    torch.save({prefix + k: v for k, v in sd.items()}, path)


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    """ORCA_PATH/models + ORCA_PATH/resources as the reference constructors expect them."""
    _, rm = _ref()
    root = tmp_path_factory.mktemp("orca_path")
    os.makedirs(root / "models")
    os.makedirs(root / "resources")
    rng = np.random.default_rng(5)
    seed = 100
    for cell in ("h1esc", "hff"):
        # stage-a file: AveragedModel(DataParallel(Net)) -> "module.module." (orca_models.py:111, :120-122, :474)
        _save(root / "models" / ("orca_%s.net0.statedict" % cell), rm.Net(num_1d=32 if cell == "h1esc" else 22), seed, "module.module."); seed += 1
        _save(root / "models" / ("orca_%s.net.statedict" % cell), rm.Encoder2(), seed, "module."); seed += 1
        for lvl in (1, 2, 4, 8, 16, 32):
            _save(root / "models" / ("orca_%s.d%d.statedict" % (cell, lvl)), rm.Decoder(upsample_mode="bilinear"), seed, "module."); seed += 1
        _save(root / "models" / ("orca_%s_256m.net.statedict" % cell), rm.Encoder3(), seed, "module."); seed += 1
        for lvl in (32, 64, 128, 256):
            _save(root / "models" / ("orca_%s_256m.d%d.statedict" % (cell, lvl)), rm.Decoder(upsample_mode="bilinear"), seed, "module."); seed += 1
    _save(root / "models" / "orca_hctnoc.net0.statedict", rm.Encoder(), seed, "module."); seed += 1
    _save(root / "models" / "orca_hctnoc.net.statedict", rm.Encoder2b(), seed, "module."); seed += 1
    for lvl in (1, 2, 4, 8, 16, 32):
        _save(root / "models" / ("orca_hctnoc.d%d.statedict" % lvl), rm.Decoder(), seed, "module."); seed += 1
    for stem in ("4DNFI9GMP2J8", "4DNFI643OYP9"):
        curve = lambda n: -(0.7 + 0.2 * rng.random()) * np.log(np.arange(n, dtype=np.float64) + 1.0) - 3.0 + 0.01 * rng.standard_normal(n)
        np.save(root / "resources" / (stem + ".rebinned.mcool.expected.res4000.npy"), curve(8000))
        np.save(root / "resources" / (stem + ".rebinned.mcool.expected.res1000.npy"), curve(1200))
        np.save(root / "resources" / (stem + ".rebinned.mcool.expected.res32000.mono.npy"), curve(8000))
        np.save(root / "resources" / (stem + ".rebinned.mcool.expected.res32000.trans.npy"), np.float64(-9.0 - rng.random()))
    np.save(root / "resources" / "4DNFILP99QJS.HCT_auxin6h.rebinned.mcool.expected.res4000.npy",
            -0.9 * np.log(np.arange(8000, dtype=np.float64) + 1.0) - 2.5)
    return str(root)


def _same_weights(ref_mod, ours):
    ref_sd = models._strip(ref_mod.state_dict())
    our_sd = ours.state_dict()
    assert list(ref_sd.keys()) == list(our_sd.keys())
    for k in ref_sd:
        assert torch.equal(ref_sd[k], our_sd[k]), k


@pytest.mark.parametrize("kind,ref_name", [("h1esc", "H1esc"), ("hff", "Hff"), ("hctnoc", "HCTnoc"),
                                           ("h1esc_1m", "H1esc_1M"), ("hff_1m", "Hff_1M"),
                                           ("h1esc_256m", "H1esc_256M"), ("hff_256m", "Hff_256M")])
def test_shell_matches_reference_constructor(tree, kind, ref_name, monkeypatch):
    om, _ = _ref()
    monkeypatch.setattr(om, "ORCA_PATH", tree)
    monkeypatch.setattr(nn.Module, "cuda", lambda self, device=None: self)  # HCTnoc.__init__ calls .cuda() (:396)
    ref = getattr(om, ref_name)()
    ours = models.build_shell(modules, kind, seed=0, orca_path=tree)
    for attr in ("net0", "net", "net1", "denet_1_pt"):
        assert hasattr(ref, attr) == hasattr(ours, attr), attr
        if hasattr(ref, attr):
            _same_weights(getattr(ref, attr), getattr(ours, attr))
    if hasattr(ref, "denets"):
        assert sorted(ref.denets) == sorted(ours.denets)
        for lvl in ref.denets:
            _same_weights(ref.denets[lvl], ours.denets[lvl])
            assert ref.denets[lvl].module.upsample.mode == ours.denets[lvl].upsample.mode
    if hasattr(ref, "normmats"):
        assert sorted(ref.normmats) == sorted(ours.normmats)
        for lvl in ref.normmats:
            assert np.array_equal(ref.normmats[lvl], ours.normmats[lvl]), lvl  # same numpy expressions: bit-exact
            assert ref.epss[lvl] == ours.epss[lvl]
    if hasattr(ref, "background_cis"):
        assert np.array_equal(ref.background_cis, ours.background_cis, equal_nan=True)
        assert float(ref.background_trans) == float(ours.background_trans)


def test_synthetic_normmats_follow_reference_expression():
    """The synthetic shells (no orca_path) use the same block-mean code path on the synthetic curve."""
    mats, epss = synthetic.normmats_32mb()
    elog = synthetic.expected_log(8000)
    normmat = np.exp(elog[np.abs(np.arange(8000)[None, :] - np.arange(8000)[:, None])])  # orca_models.py:139
    for lvl in (1, 4, 32):
        n = 250 * lvl
        want = np.reshape(normmat[:n, :n], (250, lvl, 250, lvl)).mean(axis=1).mean(axis=2)  # :141-150
        assert np.array_equal(mats[lvl], want)
        assert epss[lvl] == np.min(want)
