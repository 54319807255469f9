"""Pin the oracle (oracle/orca_oracle.py) against fixtures produced by the unmodified reference
(oracle/make_golden.py), and against the live reference classes when /root/reference exists."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, REFERENCE, relerr
import orca_oracle as oracle
from orca_b200 import modules, synthetic

TOL = 2e-6  # same torch operators on the same CPU build give 0; leave room for other oneDNN builds


def gold(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def randn(shape, seed, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))


def sd_for(module, seed):
    return synthetic.fill_state_dict(module.state_dict(), seed)


@pytest.mark.parametrize("name", ["encoder_24k", "encoder_1mb"])
def test_encoder(name):
    g = gold(name)
    sd = sd_for(modules.Encoder(), int(g["weight_seed"]))
    x = torch.from_numpy(synthetic.random_sequence(1, int(g["L"]), int(g["seq_seed"]), float(g["n_fraction"]))).transpose(1, 2)
    with torch.no_grad():
        y = oracle.encoder_forward(sd, x)
    assert y.shape == g["out"].shape
    assert relerr(y.numpy(), g["out"]) <= TOL


@pytest.mark.parametrize("name,cls,n,up", [("encoder2_p256", modules.Encoder2, 5, True),
                                           ("encoder3_p64", modules.Encoder3, 3, True),
                                           ("encoder2b_p64", modules.Encoder2b, 5, False)])
def test_unets(name, cls, n, up):
    g = gold(name)
    sd = sd_for(cls(), int(g["weight_seed"]))
    x = randn((2, 128, int(g["P"])), int(g["x_seed"]))
    with torch.no_grad():
        ys = oracle.encoder2_forward(sd, x, n=n, up=up)
    assert len(ys) == n + 1
    for i, y in enumerate(ys):
        assert relerr(y.numpy(), g["out%d" % i]) <= TOL


@pytest.mark.parametrize("name", ["decoder_nocoarse_250", "decoder_coarse_bilinear_250", "decoder_coarse_nearest_64",
                                  "decoder_nocoarse_nearest_30"])
def test_decoder(name):
    g = gold(name)
    mode, S, B = str(g["mode"]), int(g["S"]), int(g["B"])
    sd = sd_for(modules.Decoder(upsample_mode=mode), int(g["weight_seed"]))
    mats, _ = synthetic.normmats_32mb()
    x = randn((B, 128, S), int(g["x_seed"]), 0.5)
    distenc = torch.log(torch.FloatTensor(mats[int(g["level"])][:S, :S][None, None])).expand(B, -1, -1, -1)
    yc = randn((B, 1, S // 2, S // 2), int(g["y_seed"])) if bool(g["coarse"]) else None
    with torch.no_grad():
        y = oracle.decoder_forward(sd, x, distenc, yc, mode)
    assert relerr(y.numpy(), g["out"]) <= TOL
    # the decoder output is exactly symmetric (orca_modules.py:487-488)
    assert torch.equal(y, y.transpose(2, 3))


@pytest.mark.parametrize("name", ["decoder1m_250", "decoder1m_40"])
def test_decoder_1m(name):
    g = gold(name)
    sd = sd_for(modules.Decoder_1m(), int(g["weight_seed"]))
    x = randn((int(g["B"]), 128, int(g["S"])), int(g["x_seed"]), 0.5)
    with torch.no_grad():
        y = oracle.decoder_1m_forward(sd, x)
    assert relerr(y.numpy(), g["out"]) <= TOL


def test_net():
    g = gold("net_48k")
    sd = sd_for(modules.Net(num_1d=32), int(g["weight_seed"]))
    x = torch.from_numpy(synthetic.random_sequence(int(g["B"]), int(g["L"]), int(g["seq_seed"]), float(g["n_fraction"]))).transpose(1, 2)
    with torch.no_grad():
        pred, p1d = oracle.net_forward(sd, x, num_1d=32)
    assert relerr(pred.numpy(), g["out"]) <= TOL
    assert relerr(p1d.numpy(), g["out_1d"]) <= TOL


@pytest.mark.parametrize("name", ["leukemia_decoder_n2_64", "leukemia_decoder_n6_48", "leukemia_decoder_n6_30_nocoarse",
                                  "leukemia_decoder_n2_250"])
def test_leukemia_decoder(name):
    """Multi-map decoders of orca_leukemia.py (num_2d = 2 / 6): same oracle functions, wider tensors."""
    from orca_b200 import leukemia
    g = gold(name)
    n2d, S, B = int(g["num_2d"]), int(g["S"]), int(g["B"])
    sd = sd_for(leukemia.Decoder(n2d), int(g["weight_seed"]))
    x = randn((B, 128, S), int(g["x_seed"]), 0.5)
    distenc = randn((1, n2d, S, S), int(g["d_seed"])).expand(B, -1, -1, -1)
    yc = randn((B, n2d, S // 2, S // 2), int(g["y_seed"])) if bool(g["coarse"]) else None
    with torch.no_grad():
        y = oracle.decoder_forward(sd, x, distenc, yc, "nearest")
    assert tuple(y.shape) == (B, n2d, S, S)
    assert relerr(y.numpy(), g["out"]) <= TOL


def test_leukemia_decoder_1m_and_net():
    from orca_b200 import leukemia
    g = gold("leukemia_decoder1m_n2_40")
    sd = sd_for(leukemia.Decoder_1m(2), int(g["weight_seed"]))
    with torch.no_grad():
        y = oracle.decoder_1m_forward(sd, randn((2, 128, 40), int(g["x_seed"]), 0.5))
    assert relerr(y.numpy(), g["out"]) <= TOL
    g = gold("leukemia_net_n6_24k")
    sd = sd_for(leukemia.Net(6, 8), int(g["weight_seed"]))
    x = torch.from_numpy(synthetic.random_sequence(1, int(g["L"]), int(g["seq_seed"]), float(g["n_fraction"]))).transpose(1, 2)
    with torch.no_grad():
        pred, p1d = oracle.net_forward(sd, x, num_1d=8)
    assert relerr(pred.numpy(), g["out"]) <= TOL and relerr(p1d.numpy(), g["out_1d"]) <= TOL


def test_background_levels():
    g = gold("background")
    nm = synthetic.normmat_256mb(chrlen_bins=int(g["chrlen_bins"]))
    for tag, r0, level, flip in [("l256", 0, 256, False), ("l64_r", 1500, 64, True), ("l32", 4100, 32, False)]:
        d = oracle.background_level(nm, r0, level // 8, 250, flip)
        assert relerr(d.numpy(), g[tag]) <= 1e-7


def test_background_assemble():
    """Multi-region background matrix (orca_predict._retrieve_multi, :936-965) vs the fixture from the reference."""
    from orca_b200 import models
    g = gold("background_assemble")
    regions = [(str(c), int(a), int(b), str(s)) for c, a, b, s in zip(g["chroms"], g["starts"], g["ends"], g["strands"])]
    cis, trans = models._background_256mb(None, "h1esc")
    nm = oracle.assemble_background(regions, cis, trans)
    assert nm.shape == g["normmat"].shape and np.array_equal(nm, g["normmat"], equal_nan=True)


def test_blockwise_equals_monolithic():
    """The reference's 800 kb blocks with 112 kb overlap are exact w.r.t. one monolithic pass
    (receptive field 104,016 bp < 112,000 bp): property used by the chunked / sharded CUDA encoder."""
    sd = sd_for(modules.Encoder(), 3)
    x = torch.from_numpy(synthetic.random_sequence(1, 480000, 5)).transpose(1, 2)
    with torch.no_grad():
        a = oracle.encoder_forward(sd, x, blocksize=160000)  # 3 blocks
        b = oracle.encoder_run(sd, x)
    assert relerr(a.numpy(), b.numpy()) <= 1e-6


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present (GPU box)")
def test_against_live_reference():
    """Same weights, same input: oracle == unmodified reference classes; mirrors share the key set."""
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import orca_modules as om
    pairs = [(om.Encoder, modules.Encoder, {}), (om.Encoder2, modules.Encoder2, {}), (om.Encoder2b, modules.Encoder2b, {}),
             (om.Encoder3, modules.Encoder3, {}), (om.Decoder, modules.Decoder, {"upsample_mode": "bilinear"}),
             (om.Decoder_1m, modules.Decoder_1m, {}), (om.Net, modules.Net, {"num_1d": 32}), (om.Net, modules.Net, {})]
    for ref_cls, our_cls, kw in pairs:
        ref_sd, our_sd = ref_cls(**kw).state_dict(), our_cls(**kw).state_dict()
        assert list(ref_sd.keys()) == list(our_sd.keys()), ref_cls.__name__
        assert all(ref_sd[k].shape == our_sd[k].shape for k in ref_sd), ref_cls.__name__
    m = synthetic.init_module(om.Decoder(upsample_mode="bilinear"), 21)
    x, d, yc = randn((1, 128, 48), 1, 0.5), randn((1, 1, 48, 48), 2), randn((1, 1, 24, 24), 3)
    with torch.no_grad():
        assert relerr(oracle.decoder_forward(m.state_dict(), x, d, yc, "bilinear").numpy(), m(x, d, yc).numpy()) <= TOL
