"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the
nn.Module mirrors keep the reference state_dict layout, argument validation fails loudly, and the
synthetic generators are deterministic.  No compute entry point is called (no GPU here)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, relerr
from orca_b200 import _lib, models, modules, parallel, synthetic


def header_symbols():
    text = open(os.path.join(ROOT, "include", "orca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(orca_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "liborca_b200.so does not export %s" % s
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == syms
    assert b"sm_100a" in _lib.lib().orca_b200_version()


def test_state_dict_layout_matches_reference_fixture():
    """tests/golden/state_dict_keys.json was dumped from the reference classes (make_golden.py)."""
    want = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    ctors = {"Encoder": modules.Encoder, "Encoder2": modules.Encoder2, "Encoder2b": modules.Encoder2b,
             "Encoder3": modules.Encoder3, "Decoder": lambda: modules.Decoder(upsample_mode="bilinear"),
             "Decoder_1m": modules.Decoder_1m, "Net32": lambda: modules.Net(num_1d=32), "Net": modules.Net}
    for name, ctor in ctors.items():
        sd = ctor().state_dict()
        got = [[k, list(v.shape)] for k, v in sd.items()]
        assert got == want[name], name
    assert len(want["Encoder"]) == 196 and len(want["Decoder"]) == 849


def test_leukemia_state_dict_layout_matches_reference_fixture():
    """orca_leukemia.py classes (num_2d = 2 / 6): same trees, wider combiner / final convs."""
    from orca_b200 import leukemia
    want = json.load(open(os.path.join(GOLDEN, "state_dict_keys_leukemia.json")))
    ctors = {"Decoder2": lambda: leukemia.Decoder(2), "Decoder6": lambda: leukemia.Decoder(6),
             "Decoder_1m2": lambda: leukemia.Decoder_1m(2), "Net6_8": lambda: leukemia.Net(6, 8),
             "Encoder": leukemia.Encoder, "Encoder2": leukemia.Encoder2}
    for name, ctor in ctors.items():
        got = [[k, list(v.shape)] for k, v in ctor().state_dict().items()]
        assert got == want[name], name
    with pytest.raises(ValueError):
        leukemia.Decoder(9)
    sh = models.build_shell(leukemia, "leukemia_a", seed=2)
    assert sorted(sh.denets) == [1, 2, 4, 8, 16, 32] and sh.normmats[4].shape == (2, 250, 250)
    assert sh.denets[1].num_2d == 2 and sh.denet_1_pt.num_2d == 2 and isinstance(sh.net, modules.Encoder2b)


def test_packed_feeder_matches_reference_encoding():
    """orca_b200.feeder: packed codes / ASCII <-> the reference's one-hot rows (selene_utils2.py:216-230 semantics:
    ACGT one-hot, everything else 0.25), and the strand flip of orca_predict.py:324-329 on packed codes."""
    from orca_b200 import feeder
    seq = synthetic.random_sequence(2, 5000, 7, 0.03)
    codes = feeder.from_onehot(seq)
    assert codes.dtype == np.uint8 and codes.shape == (2, 5000) and set(np.unique(codes)) == {0, 1, 2, 3, 4}
    assert np.array_equal(feeder.to_onehot(codes), seq)
    text = "ACGTNacgtnRYKM-*"
    want = np.array([0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 4, 4, 4, 4, 4, 4], dtype=np.uint8)
    assert np.array_equal(feeder.codes(text), want)
    assert np.array_equal(feeder.codes(text.encode()), want)
    assert np.array_equal(feeder.as_bases(text), np.frombuffer(text.encode(), dtype=np.uint8))  # ASCII passes through
    assert np.array_equal(feeder.to_onehot(feeder.reverse_complement(codes)), seq[:, ::-1, ::-1])
    bad = seq.copy()
    bad[0, 3] = [0.5, 0.5, 0.0, 0.0]
    with pytest.raises(ValueError):
        feeder.from_onehot(bad)
    with pytest.raises(TypeError):
        feeder.as_bases(np.zeros(4, dtype=np.int64))


def test_cpu_tensors_are_rejected():
    enc = modules.Encoder()
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc(torch.zeros(1, 4, 4000))
    with pytest.raises(RuntimeError, match="no CPU path"):
        modules.Decoder()(torch.zeros(1, 128, 8), torch.zeros(1, 1, 8, 8))
    with pytest.raises(RuntimeError):
        enc.train()


def test_module_create_validates_architecture():
    lib = _lib.lib()
    handle = ctypes.c_void_p()
    arr = (_lib.ConvParams * 3)()
    st = lib.orca_b200_module_create(_lib.ENCODER, arr, 3, 0, 0, ctypes.byref(handle))
    assert st == -1 and b"expects 28" in lib.orca_b200_last_error()
    st = lib.orca_b200_module_create(99, arr, 3, 0, 0, ctypes.byref(handle))
    assert st == -1 and b"unknown module kind" in lib.orca_b200_last_error()
    arr = (_lib.ConvParams * 28)()
    for p in arr:
        p.c_in, p.c_out, p.kh, p.kw, p.dilation = 64, 64, 1, 9, 1
    st = lib.orca_b200_module_create(_lib.ENCODER, arr, 28, 0, 0, ctypes.byref(handle))
    assert st == -1 and b"conv 0" in lib.orca_b200_last_error()
    assert lib.orca_b200_module_set_option(None, _lib.OPT_IMPL, 1) == -1 and b"NULL module" in lib.orca_b200_last_error()


def test_abi_argument_checks_without_a_gpu():
    """Entry points that validate before touching the device: the multi-map module spec (num_2d read off the final
    conv), the region arithmetic of the background assembly, the packed-input stride check."""
    lib = _lib.lib()
    handle = ctypes.c_void_p()
    # a Decoder table whose final conv claims 9 maps: outside [1, 8]
    arr = (_lib.ConvParams * 122)()
    arr[113].c_out = 9
    assert lib.orca_b200_module_create(_lib.DECODER, arr, 122, 0, 0, ctypes.byref(handle)) == -1
    assert b"num_2d" in lib.orca_b200_last_error()
    # region arithmetic: int((end - start) / binsize) bins per region, as orca_predict.py:948-957 counts them
    regs = (_lib.Region * 3)()
    for r, (c, s0, e0, rv) in zip(regs, [(0, 0, 3_200_000, 0), (0, 8_000_000, 9_600_000, 1), (1, 40_000_000, 40_816_000, 1)]):
        r.chrom, r.start, r.end, r.reverse = c, s0, e0, rv
    assert lib.orca_b200_background_bins(regs, 3, 32000) == 100 + 50 + 25
    regs[1].end = regs[1].start
    assert lib.orca_b200_background_bins(regs, 3, 32000) == -1 and b"empty" in lib.orca_b200_last_error()
    # packed encoder input with a zero position stride is rejected before any pointer is looked at
    assert lib.orca_b200_encoder_forward_packed(None, None, 1, 4000, 4000, 0, 0, 0, 4000, None, 0, 1, 0, None, 0, None) == -1
    # kernel selection / precision defaults live on the Python side and reach the library per handle
    assert _lib.set_encoder_fp16_stages(0) == 4  # previous effective setting: library default, stages 1-4 single-pass
    assert _lib.set_encoder_fp16_stages(-1) == 0
    with pytest.raises(ValueError):
        _lib.set_impl("cpu")


def test_synthetic_is_deterministic():
    a = synthetic.fill_state_dict(modules.Encoder2b().state_dict(), 5)
    b = synthetic.fill_state_dict(modules.Encoder2b().state_dict(), 5)
    assert all(torch.equal(a[k], b[k]) for k in a)
    s = synthetic.random_sequence(2, 1000, 3, 0.05)
    assert s.shape == (2, 1000, 4) and np.allclose(s.sum(-1), 1.0)
    assert set(np.unique(s)) <= {0.0, 0.25, 1.0}
    mats, epss = synthetic.normmats_32mb()
    assert sorted(mats) == [1, 2, 4, 8, 16, 32] and mats[32].shape == (250, 250)
    assert np.allclose(mats[4], mats[4].T) and epss[1] == mats[1].min()


def test_shell_protocol():
    sh = models.build_shell(modules, "h1esc", seed=3)
    assert isinstance(sh, torch.nn.Module)
    assert sorted(sh.denets) == [1, 2, 4, 8, 16, 32]
    for attr in ("net0", "net", "denet_1_pt", "normmats", "epss"):
        assert hasattr(sh, attr)
    sh256 = models.build_shell(modules, "h1esc_256m", seed=3)
    assert sorted(sh256.denets) == [32, 64, 128, 256] and hasattr(sh256, "net1")
    assert sh256.background_cis.shape == (10000,) and np.isnan(sh256.background_cis[-1])
    hct = models.build_shell(modules, "hctnoc", seed=3)
    assert isinstance(hct.net, modules.Encoder2b) and not hasattr(hct, "denet_1_pt")
    assert hct.denets[1].upsample.mode == "nearest"


def test_shard_geometry():
    for P, world in [(8000, 1), (8000, 2), (8000, 8), (64000, 8), (250, 4)]:
        ranges = [parallel.shard_bins(P, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == P
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    s0, s1 = parallel.shard_window(32_000_000, 1000, 2000)
    assert s0 == 1000 * 4000 - 116000 and s1 == 2000 * 4000 + 116000
    assert parallel.shard_window(32_000_000, 0, 8000) == (0, 32_000_000)


class _FakeDecoder(torch.nn.Module):
    """Stand-in with Decoder.forward's signature: a cheap deterministic function of all three inputs."""

    def __init__(self, k):
        super().__init__()
        self.k = k

    def forward(self, x, distenc, y=None):
        m = x.mean(1)
        out = (m[:, :, None] + m[:, None, :])[:, None] * self.k + 0.1 * distenc
        if y is not None:
            out = out + torch.nn.functional.interpolate(y, scale_factor=(2, 2), mode="nearest")
        return out


def test_lane_batched_cascades_equal_separate_cascades(monkeypatch):
    """Host logic of predict.cascade_*_lanes (strands as batch elements of one chain): per-lane crop windows,
    coarse crops and start bins must equal running each strand's cascade on its own."""
    import orca_oracle as oracle
    from orca_b200 import predict
    rng = np.random.default_rng(0)
    shell = torch.nn.Module()
    mats, _ = synthetic.normmats_32mb()
    shell.normmats = mats
    dec = {lvl: _FakeDecoder(1.0 + 0.1 * i) for i, lvl in enumerate([1, 2, 4, 8, 16, 32])}
    shell.denets = dec
    shell.denet_1_pt = type("D1", (), {"forward": staticmethod(lambda x: (x.mean(1)[:, :, None] * x.mean(1)[:, None, :])[:, None])})()
    encs = [{lvl: torch.from_numpy(rng.standard_normal((1, 128, 8000 // lvl)).astype(np.float32)) for lvl in [1, 2, 4, 8, 16, 32]}
            for _ in range(2)]
    mpos, wpos = 17_300_000, 16_000_000
    lanes_p, lanes_s = predict.cascade_32mb_lanes(shell, [(encs[0], False), (encs[1], True)], mpos, wpos)
    for i, rev in enumerate((False, True)):
        p, st = predict.cascade_32mb(shell, encs[i], 1, mpos, wpos, rev)
        assert st == lanes_s[i]
        for a, b in zip(p, lanes_p):
            assert torch.equal(a[0], b[i])
    assert lanes_s[0] != lanes_s[1]  # the two strands really zoom into different windows
    assert lanes_s[0] == predict.cascade_starts_32mb(mpos, wpos, False)

    # 256 Mb analogue; the device background kernel is replaced by the oracle's numpy restatement
    nm = synthetic.normmat_256mb(chrlen_bins=7000)
    def fake_level(normmat, r0, f, flip=False, size=250, with_mean=False):
        log = oracle.background_level(normmat, r0, f, size, flip)
        return (log, torch.exp(oracle.background_level(normmat, r0, f, size, False)[0].double())) if with_mean else log
    monkeypatch.setattr(predict, "background_level", fake_level)
    shell.denets = {lvl: _FakeDecoder(1.0 + 0.1 * i) for i, lvl in enumerate([32, 64, 128, 256])}
    encs = [{lvl: torch.from_numpy(rng.standard_normal((1, 128, 8000 * 32 // lvl)).astype(np.float32)) for lvl in [32, 64, 128, 256]}
            for _ in range(2)]
    mpos, wpos, chrlen = 100_000_000, 128_000_000, 7000 * 32000
    lanes_p, lanes_s = predict.cascade_256mb_lanes(shell, [(encs[0], False), (encs[1], True)], nm, chrlen, mpos, wpos)
    for i, rev in enumerate((False, True)):
        p, st, _ = predict.cascade_256mb(shell, encs[i], 1, nm, chrlen, mpos, wpos, rev)
        assert st == lanes_s[i]
        for a, b in zip(p, lanes_p):
            assert torch.equal(a[0], b[i])


def test_cascade_index_math_matches_reference_drivers(monkeypatch):
    """Driver logic at the edges (crop index clipped to 0 and 125, mpos = wpos, shifted windows, chromosomes shorter
    than the window so that bounds[0] >= bounds[1], both strands): orca_b200.predict's host logic on stand-in networks
    (tests/fakes.py) against maps the UNMODIFIED orca_predict.genomepredict / genomepredict_256Mb produced with the
    same stand-ins (oracle/make_golden.py `cascade_index_cases`)."""
    import fakes
    import orca_oracle as oracle
    from orca_b200 import predict
    g = np.load(os.path.join(GOLDEN, "cascade_index_cases.npz"))
    cpu = torch.device("cpu")
    sh = fakes.FakeShell("h1esc")
    seq = fakes.stub_sequence(32_000_000, int(g["seq32_seed"]))
    for i, (mpos, wpos) in enumerate(fakes.CASES_32MB):
        out = predict._genomepredict_on(cpu, seq, "chrS", mpos, wpos, [sh])
        assert out["start_coords"] == [int(v) for v in g["s32_%d" % i]], (mpos, wpos)
        got = np.stack(out["predictions"][0])[:, ::5, ::5]  # the fixture keeps every 5th row / column
        assert relerr(got, g["p32_%d" % i]) <= 1e-6, (mpos, wpos)
    starts = [predict.cascade_starts_32mb(m, w, False) for m, w in fakes.CASES_32MB]
    assert any(s[-1] == s[-2] for s in starts) and any((s[1] - s[0]) == 125 * 32 for s in starts)  # both clip edges were hit

    def fake_level(normmat, r0, f, flip=False, size=250, with_mean=False):
        nm = normmat.numpy() if isinstance(normmat, torch.Tensor) else normmat
        log = oracle.background_level(nm, r0, f, size, flip)
        blk = nm[r0:r0 + size * f, r0:r0 + size * f]
        mean = np.nanmean(np.nanmean(np.reshape(blk, (1, size, f, size, f)), axis=4), axis=2)
        return (log, torch.from_numpy(mean)) if with_mean else log
    monkeypatch.setattr(predict, "background_level", fake_level)
    sh = fakes.FakeShell("h1esc_256m")
    seq = fakes.stub_sequence(64000, int(g["seq256_seed"]))
    for i, (mpos, wpos, chrlen) in enumerate(fakes.CASES_256MB):
        nm = synthetic.normmat_256mb(chrlen_bins=min(8000, chrlen // 32000))
        out = predict._genomepredict_256mb_on(cpu, seq, "chrS", [nm], chrlen, mpos, wpos, [sh])
        assert out["start_coords"] == [int(v) for v in g["s256_%d" % i]], (mpos, wpos, chrlen)
        assert [int(v) for v in out["end_coords"]] == [int(v) for v in g["e256_%d" % i]]
        assert relerr(np.stack(out["predictions"][0])[:, ::5, ::5], g["p256_%d" % i]) <= 1e-6, (mpos, wpos, chrlen)
        nms = np.stack([np.stack([ns[l][0] for l in (256, 128, 64, 32)]) for ns in out["normmats"]])[:, :, ::10, ::10]
        assert nms.dtype == np.float64 and np.array_equal(nms, g["n256_%d" % i])


def test_log_normmat_cache_follows_the_source_array():
    """the cached log(normmat) must track replacement and in-place edits of model.normmats[level]."""
    import fakes
    from orca_b200 import predict
    sh = fakes.FakeShell("h1esc")
    cpu = torch.device("cpu")
    a = predict._log_normmat(sh, 4, cpu)
    assert predict._log_normmat(sh, 4, cpu) is a
    sh.normmats[4] = sh.normmats[4] * 2.0
    b = predict._log_normmat(sh, 4, cpu)
    assert torch.allclose(b, a + np.log(2.0), atol=1e-6)
    sh.normmats[4][10, 10] *= 3.0
    c = predict._log_normmat(sh, 4, cpu)
    assert c is not b and abs(float(c[0, 0, 10, 10] - b[0, 0, 10, 10]) - np.log(3.0)) < 1e-5


def test_variant_window_inputs_normalise_to_packed_codes():
    """orca_b200.variants accepts a window as the reference passes it ((1, L, 4) one-hot), as text / bytes, or as packed
    codes; all forms must give the same packed bases (the cache keys are hashes of these)."""
    from orca_b200 import feeder, variants
    seq = synthetic.random_sequence(1, 8000, 3, 0.02)
    codes = feeder.from_onehot(seq)[0]
    text = "".join("ACGTN"[c] for c in codes)
    for form in (seq, seq[0], codes, codes[None], text, text.encode(), np.frombuffer(text.encode(), dtype=np.uint8)):
        got = variants._as_codes(form)
        assert got.dtype == np.uint8 and got.shape == (8000,) and np.array_equal(got, codes)
    assert variants.HALO_BP == parallel.HALO_BP
