"""GPU parity: the CUDA path (through the C ABI, via the nn.Module mirrors) against the golden
fixtures produced by the unmodified reference and against the oracle on the same seeded inputs.

Bar (BASELINE.json north star): max|ours - ref| / max|ref| <= 1e-3 per output, fp32.
The fp32 SIMT path is additionally held to 5e-5 (it differs from the reference only by fp32
summation order and the BatchNorm fold)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT, relerr
import orca_oracle as oracle
from orca_b200 import _lib, modules, synthetic

pytestmark = pytest.mark.gpu

TOL = 1e-3        # north-star tolerance
TOL_SIMT = 5e-5   # exact-fp32 path

IMPLS = ["simt", "auto"]


def gold(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def randn(shape, seed, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))


def native(module, seed):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return synthetic.init_module(module, seed).cuda()


def tol(impl):
    return TOL_SIMT if impl == "simt" else TOL


@pytest.fixture(params=IMPLS)
def impl(request):
    _lib.set_impl(request.param)
    yield request.param
    _lib.set_impl("auto")


@pytest.mark.parametrize("name", ["encoder_24k", "encoder_1mb"])
def test_encoder_golden(name, impl):
    g = gold(name)
    m = native(modules.Encoder(), int(g["weight_seed"]))
    seq = synthetic.random_sequence(1, int(g["L"]), int(g["seq_seed"]), float(g["n_fraction"]))
    x = torch.from_numpy(seq).transpose(1, 2).cuda()  # strided view, as genomepredict passes it (orca_predict.py:334)
    assert x.stride() == (4 * int(g["L"]), 1, 4)
    n0 = _lib.launch_count()
    y = m(x)
    assert _lib.launch_count() > n0
    assert tuple(y.shape) == g["out"].shape
    e = relerr(y.cpu().numpy(), g["out"])
    print(name, impl, "relerr %.2e" % e)
    assert e <= tol(impl)
    if impl == "auto":  # encoder precision settings: three-product bf16 everywhere (0), default (3), single-pass fp16 everywhere (7)
        errs = {}
        for stages in (0, 3, 4, 5, 7):
            _lib.set_encoder_fp16_stages(stages)
            try:
                errs[stages] = relerr(m(x).cpu().numpy(), g["out"])
            finally:
                _lib.set_encoder_fp16_stages(-1)
        print(name, "relerr by fp16 stages", {k: "%.2e" % v for k, v in errs.items()})
        assert errs[0] <= 5e-5 and errs[3] <= 1e-4 and errs[4] <= 1e-4 and errs[7] <= TOL


def test_encoder_layouts_and_chunks(impl):
    """Contiguous (B,4,L) input == channel-last view; chunked == single pass; batch handled."""
    m = native(modules.Encoder(), 5)
    L = 480000
    seq = synthetic.random_sequence(2, L, 9, 0.01)
    xt = torch.from_numpy(seq).cuda()
    a = m(xt.transpose(1, 2))
    b = m(xt.transpose(1, 2).contiguous())
    assert torch.equal(a, b)
    m.chunk_bp = 120000  # 4 chunks per sample, 112 kb halo each side
    c = m(xt.transpose(1, 2))
    m.chunk_bp = 0
    assert torch.equal(a, c), "chunked encoder must be bit-identical to the single pass"
    with torch.no_grad():
        ref = oracle.encoder_forward(synthetic.fill_state_dict(m.state_dict(), 5), torch.from_numpy(seq).transpose(1, 2))
    assert relerr(a.cpu().numpy(), ref.numpy()) <= tol(impl)
    # bin sub-range (what a sequence shard computes)
    part = m(xt.transpose(1, 2), bin_range=(30, 90))
    assert torch.equal(part[:, :, 30:90], a[:, :, 30:90])


def test_packed_bases_equal_onehot_input(impl):
    """the encoder fed with 1 B/bp packed bases (codes or raw ASCII, forward and
    reverse-complement strand, whole sequence or a shard window) is BIT-identical to the fp32 one-hot path."""
    from orca_b200 import feeder
    m = native(modules.Encoder(), 5)
    L = 480000
    seq = synthetic.random_sequence(2, L, 9, 0.02)
    codes = feeder.from_onehot(seq)
    ascii_ = np.frombuffer(b"ACGTN", dtype=np.uint8)[codes]
    ascii_[0, ::7] += 32  # some lower-case bases
    xf = torch.from_numpy(seq).cuda().transpose(1, 2)
    for rc in (False, True):
        want = m(xf, reverse_complement=rc)
        for packed in (codes, ascii_):
            got = m(torch.from_numpy(packed).cuda(), reverse_complement=rc)
            assert torch.equal(got, want), (impl, rc)
    # a shard: only a window of the forward strand is resident (orca_b200.parallel)
    pos0, pos1 = 80000, 400000
    part_f = m(xf[:, :, pos0:pos1], bin_range=(50, 70), window=(pos0, L))
    part_p = m(torch.from_numpy(codes[:, pos0:pos1]).cuda(), bin_range=(50, 70), window=(pos0, L))
    assert torch.equal(part_p[:, :, 50:70], part_f[:, :, 50:70])
    part_f = m(xf[:, :, pos0:pos1], bin_range=(50, 70), window=(pos0, L), reverse_complement=True)
    part_p = m(torch.from_numpy(codes[:, pos0:pos1]).cuda(), bin_range=(50, 70), window=(pos0, L), reverse_complement=True)
    assert torch.equal(part_p[:, :, 50:70], part_f[:, :, 50:70])
    # and against the oracle on the reference-format array the codes stand for
    with torch.no_grad():
        ref = oracle.encoder_forward(synthetic.fill_state_dict(m.state_dict(), 5), torch.from_numpy(feeder.to_onehot(codes)).transpose(1, 2))
    assert relerr(m(torch.from_numpy(codes).cuda()).cpu().numpy(), ref.numpy()) <= tol(impl)
    with pytest.raises(RuntimeError):
        m(torch.from_numpy(codes))  # CPU tensor
    # Net (Orca-1Mb) takes packed windows too
    net = native(modules.Net(), 17)
    s1 = synthetic.random_sequence(2, 48000, 104, 0.01)
    assert torch.equal(net(torch.from_numpy(feeder.from_onehot(s1)).cuda()), net(torch.from_numpy(s1).transpose(1, 2).cuda()))


@pytest.mark.parametrize("name,cls", [("encoder2_p256", modules.Encoder2), ("encoder3_p64", modules.Encoder3),
                                      ("encoder2b_p64", modules.Encoder2b)])
def test_unets_golden(name, cls, impl):
    g = gold(name)
    m = native(cls(), int(g["weight_seed"]))
    x = randn((2, 128, int(g["P"])), int(g["x_seed"])).cuda()
    ys = m(x)
    for i, y in enumerate(ys):
        assert tuple(y.shape) == g["out%d" % i].shape
        assert relerr(y.cpu().numpy(), g["out%d" % i]) <= tol(impl), (name, i)
    # strided (channel-last) input gives the same result
    ys2 = m(x.transpose(1, 2).contiguous().transpose(1, 2))
    assert all(torch.equal(a, b) for a, b in zip(ys, ys2))
    if cls is modules.Encoder2:  # what genomepredict_256Mb consumes: net1(...)[-1]
        last = m(x, coarsest_only=True)[-1]
        assert torch.equal(last, ys[-1])


@pytest.mark.parametrize("name", ["decoder_nocoarse_250", "decoder_coarse_bilinear_250", "decoder_coarse_nearest_64",
                                  "decoder_nocoarse_nearest_30"])
def test_decoder_golden(name, impl):
    g = gold(name)
    mode, S, B = str(g["mode"]), int(g["S"]), int(g["B"])
    m = native(modules.Decoder(upsample_mode=mode), int(g["weight_seed"]))
    mats, _ = synthetic.normmats_32mb()
    x = randn((B, 128, S), int(g["x_seed"]), 0.5).cuda()
    distenc = torch.log(torch.FloatTensor(mats[int(g["level"])][:S, :S][None, None]).cuda()).expand(B, -1, -1, -1)
    yc = randn((B, 1, S // 2, S // 2), int(g["y_seed"])).cuda() if bool(g["coarse"]) else None
    y = m(x, distenc, yc)
    e = relerr(y.cpu().numpy(), g["out"])
    print(name, impl, "relerr %.2e" % e)
    assert e <= tol(impl)
    assert torch.equal(y, y.transpose(2, 3)), "symmetrised output must be exactly symmetric"


def test_decoder_strided_inputs(impl):
    """x as a slice of a channel-last encoding, distenc flipped / expanded, y as a crop of a
    previous prediction -- the views genomepredict actually passes (orca_predict.py:356-401, :703)."""
    m = native(modules.Decoder(upsample_mode="bilinear"), 15)
    S = 48
    enc = randn((1, 128, 200), 1, 0.5).cuda().transpose(1, 2).contiguous().transpose(1, 2)
    x = enc[:, :, 100:100 + S]
    d = randn((1, 1, S, S), 2).cuda().expand(1, -1, -1, -1)
    prev = randn((1, 1, 2 * S, 2 * S), 3).cuda()
    yc = prev[:, :, 7:7 + S // 2, 7:7 + S // 2]
    y = m(x, torch.flip(d, [2, 3]), yc)
    sd = synthetic.fill_state_dict(m.state_dict(), 15)
    with torch.no_grad():
        ref = oracle.decoder_forward(sd, x.cpu(), torch.flip(d, [2, 3]).cpu(), yc.cpu(), "bilinear")
    assert relerr(y.cpu().numpy(), ref.numpy()) <= tol(impl)


@pytest.mark.parametrize("name", ["decoder1m_250", "decoder1m_40"])
def test_decoder_1m_golden(name, impl):
    g = gold(name)
    m = native(modules.Decoder_1m(), int(g["weight_seed"]))
    x = randn((int(g["B"]), 128, int(g["S"])), int(g["x_seed"]), 0.5).cuda()
    y = m(x)
    assert relerr(y.cpu().numpy(), g["out"]) <= tol(impl)


def test_net_golden(impl):
    g = gold("net_48k")
    m = native(modules.Net(num_1d=32), int(g["weight_seed"]))
    seq = synthetic.random_sequence(int(g["B"]), int(g["L"]), int(g["seq_seed"]), float(g["n_fraction"]))
    pred, p1d = m(torch.from_numpy(seq).transpose(1, 2).cuda())
    assert relerr(pred.cpu().numpy(), g["out"]) <= tol(impl)
    assert relerr(p1d.cpu().numpy(), g["out_1d"]) <= tol(impl)
    m2 = native(modules.Net(), 17)  # no final_1d head: returns the map only (orca_modules.py:1897-1900)
    assert isinstance(m2(torch.from_numpy(seq).transpose(1, 2).cuda()), torch.Tensor)


@pytest.mark.parametrize("name", ["leukemia_decoder_n2_64", "leukemia_decoder_n6_48", "leukemia_decoder_n6_30_nocoarse",
                                  "leukemia_decoder_n2_250"])
def test_leukemia_decoder_golden(name, impl):
    """orca_leukemia.Decoder(num_2d): num_2d distance / coarse channels in, num_2d maps out."""
    from orca_b200 import leukemia
    g = gold(name)
    n2d, S, B = int(g["num_2d"]), int(g["S"]), int(g["B"])
    m = native(leukemia.Decoder(n2d), int(g["weight_seed"]))
    x = randn((B, 128, S), int(g["x_seed"]), 0.5).cuda()
    distenc = randn((1, n2d, S, S), int(g["d_seed"])).cuda().expand(B, -1, -1, -1)
    yc = randn((B, n2d, S // 2, S // 2), int(g["y_seed"])).cuda() if bool(g["coarse"]) else None
    y = m(x, distenc, yc)
    assert tuple(y.shape) == (B, n2d, S, S)
    e = relerr(y.cpu().numpy(), g["out"])
    print(name, impl, "relerr %.2e" % e)
    assert e <= tol(impl)
    assert torch.equal(y, y.transpose(2, 3))
    if yc is not None:  # coarse map as a crop of a larger prediction (channel stride != (S/2)^2), as the cascade passes it
        big = torch.zeros((B, n2d, S, S), device="cuda")
        big[:, :, 3:3 + S // 2, 5:5 + S // 2] = yc
        assert torch.equal(m(x, distenc, big[:, :, 3:3 + S // 2, 5:5 + S // 2]), y)


def test_leukemia_decoder_1m_and_net_golden(impl):
    from orca_b200 import leukemia
    g = gold("leukemia_decoder1m_n2_40")
    m = native(leukemia.Decoder_1m(2), int(g["weight_seed"]))
    y = m(randn((2, 128, 40), int(g["x_seed"]), 0.5).cuda())
    assert relerr(y.cpu().numpy(), g["out"]) <= tol(impl)
    g = gold("leukemia_net_n6_24k")
    m = native(leukemia.Net(6, 8), int(g["weight_seed"]))
    seq = synthetic.random_sequence(1, int(g["L"]), int(g["seq_seed"]), float(g["n_fraction"]))
    pred, p1d = m(torch.from_numpy(seq).transpose(1, 2).cuda())
    assert tuple(pred.shape) == (1, 6, 6, 6)
    assert relerr(pred.cpu().numpy(), g["out"]) <= tol(impl)
    assert relerr(p1d.cpu().numpy(), g["out_1d"]) <= tol(impl)


def test_background_levels():
    import ctypes
    g = gold("background")
    nm = torch.from_numpy(synthetic.normmat_256mb(chrlen_bins=int(g["chrlen_bins"]))).cuda()
    for tag, r0, level, flip in [("l256", 0, 256, False), ("l64_r", 1500, 64, True), ("l32", 4100, 32, False)]:
        out = torch.empty((250, 250), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib().orca_b200_background_forward(nm.data_ptr(), 8000, r0, level // 8, 250, int(flip),
                                                           out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        assert relerr(out.cpu().numpy(), g[tag][0, 0]) <= 1e-6


def test_background_assemble_golden():
    """Device assembly of the multi-region background matrix: bit-exact against orca_predict._retrieve_multi
    (fixture), including reverse-strand regions, a second chromosome and a region that is not a multiple of 32 kb;
    and at the full 8000 x 8000 size against the oracle's numpy restatement through the level kernel."""
    from orca_b200 import models, predict
    g = gold("background_assemble")
    regions = [(str(c), int(a), int(b), str(s)) for c, a, b, s in zip(g["chroms"], g["starts"], g["ends"], g["strands"])]
    cis, trans = models._background_256mb(None, "h1esc")
    nm = predict.assemble_background(regions, cis, trans, "cuda")
    assert nm.dtype == torch.float64 and tuple(nm.shape) == g["normmat"].shape
    assert np.array_equal(nm.cpu().numpy(), g["normmat"], equal_nan=True)
    big = [("chr7", 0, 160_000_000, "+"), ("chr9", 1_000_000, 97_000_000, "-")]
    ref = oracle.assemble_background(big, cis, trans)
    got = predict.assemble_background(big, cis, trans, "cuda")
    assert np.array_equal(got.cpu().numpy(), ref, equal_nan=True)
    # feeds the level kernel like a host matrix does (NaN pads replaced by the minimum first, orca_predict.py:668-671)
    a = predict.background_level(predict.prepare_background(got, torch.device("cuda")), 1000, 16)
    b = predict.background_level(predict.prepare_background(ref, torch.device("cuda")), 1000, 16)
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):  # a lookup beyond the curve (numpy raises IndexError there)
        predict.assemble_background([("chr1", 0, 400_000_000, "+")], cis, trans, "cuda")


def test_errors_are_loud():
    m = native(modules.Encoder(), 1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 8000))  # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 4001, device="cuda"))
    d = native(modules.Decoder(), 2)
    with pytest.raises(RuntimeError):
        d(torch.zeros(1, 128, 10, device="cuda"), torch.zeros(1, 1, 11, 11, device="cuda"))


# ---------------------------------------------------------------------------------------------------
# whole passes: the B200 drivers (orca_b200.predict / orca_b200.parallel) against fixtures produced by
# the UNMODIFIED orca_predict.genomepredict / genomepredict_256Mb on reference-module shells
# ---------------------------------------------------------------------------------------------------
def test_genomepredict_32mb_golden():
    from orca_b200 import models, parallel, predict
    g = gold("genomepredict_32mb")
    shell = models.H1esc(seed=int(g["shell_seed"]))
    seq = synthetic.random_sequence(1, 32_000_000, int(g["seq_seed"]))
    mpos, wpos = int(g["mpos"]), int(g["wpos"])
    out = predict.genomepredict(seq, "chrS", mpos, wpos, models=[shell])
    assert out["start_coords"] == [int(v) for v in g["start_coords"]]
    errs = [relerr(p, r) for p, r in zip(out["predictions"][0], g["predictions"])]
    print("genomepredict 32 Mb relerr per level (32..1 Mb):", ["%.1e" % e for e in errs])
    assert max(errs) <= TOL
    # the bench's sharded runner is the same computation (world size 1)
    runner = parallel.ShardedForward(shell, 32_000_000, 0, 1, torch.device("cuda"))
    runner.upload(torch.from_numpy(seq))
    maps = runner.forward(mpos, wpos).cpu().numpy()
    assert max(relerr(maps[i], g["predictions"][i]) for i in range(6)) <= TOL
    # one strand cascade at a time (what a multi-GPU rank runs) computes bit-identical maps to the batched-strand chain:
    # the arithmetic of a tile does not depend on the batch or on the schedule
    runner.cascade_mode = "serial"
    maps_s = runner.forward(mpos, wpos).cpu().numpy()
    runner.cascade_mode = "batch"
    assert np.array_equal(maps_s, maps)
    # packed-base feeder (1 B/bp): identical maps from both drivers
    from orca_b200 import feeder
    codes = feeder.from_onehot(seq)
    out_p = predict.genomepredict(codes, "chrS", mpos, wpos, models=[shell])
    assert all(np.array_equal(a, b) for a, b in zip(out_p["predictions"][0], out["predictions"][0]))
    runner.upload(torch.from_numpy(codes))
    assert runner.h2d_bytes == 32_000_000
    assert np.array_equal(runner.forward(mpos, wpos).cpu().numpy(), maps)


def test_genomepredict_32mb_leukemia_golden():
    """OrcaLeukemiaA-like shell (2 datasets per map, pooling-only Encoder2, nearest upsample, (2,250,250) normmats)
    against the unmodified orca_predict.genomepredict on the reference's orca_leukemia classes."""
    from orca_b200 import models, predict
    g = gold("genomepredict_32mb_leukemia")
    shell = models.OrcaLeukemiaA(seed=int(g["shell_seed"]))
    seq = synthetic.random_sequence(1, 32_000_000, int(g["seq_seed"]))
    out = predict.genomepredict(seq, "chrS", int(g["mpos"]), int(g["wpos"]), models=[shell])
    assert out["start_coords"] == [int(v) for v in g["start_coords"]]
    assert out["predictions"][0][0].shape == (2, 250, 250)
    errs = [relerr(p, r) for p, r in zip(out["predictions"][0], g["predictions"])]
    print("genomepredict 32 Mb leukemia-A relerr per level (32..1 Mb):", ["%.1e" % e for e in errs])
    assert max(errs) <= TOL


def test_genomepredict_256mb_driver_golden():
    """Driver logic of the 256 Mb path (net1(...)[-1], Encoder3, background levels on the GPU, cascade index
    math with chrlen clipping, strand flip) with the fixture's stub 4 kb encoding in place of net0."""
    from orca_b200 import models, predict
    g = gold("genomepredict_256mb_stub")
    shell = models.H1esc_256M(seed=int(g["shell_seed"]))

    class StubNet0(torch.nn.Module):
        def forward(self, x, reverse_complement=False):
            e = np.random.default_rng(int(g["enc_seed"])).standard_normal((x.shape[0], 128, 64000)) * 0.5
            return torch.from_numpy(e.astype(np.float32)).cuda()
    shell.net0 = StubNet0()
    seq = synthetic.random_sequence(1, 4000, 107)
    nm = synthetic.normmat_256mb(chrlen_bins=int(g["chrlen_bins"]))
    out = predict.genomepredict_256Mb(seq, "chrS", [nm], int(g["chrlen"]), int(g["mpos"]), int(g["wpos"]), models=[shell])
    assert out["start_coords"] == [int(v) for v in g["start_coords"]]
    errs = [relerr(p, r) for p, r in zip(out["predictions"][0], g["predictions"])]
    print("genomepredict 256 Mb (stub encoder) relerr per level (256..32 Mb):", ["%.1e" % e for e in errs])
    assert max(errs) <= TOL


def test_encoder_locality_at_scale():
    """Size-independent property at a size the oracle cannot reach in full: bins of an 8 Mb encode equal the
    oracle's encode of a 1.2 Mb window around them (receptive field 104,016 bp < 112,000 bp halo), on both
    strands (the reverse strand is read in place with negative strides)."""
    m = native(modules.Encoder(), 6)
    L = 8_000_000
    seq = synthetic.random_sequence(1, L, 12)
    xd = torch.from_numpy(seq).cuda()
    sd = synthetic.fill_state_dict(m.state_dict(), 6)
    for rc in (False, True):
        full = m(xd.transpose(1, 2), reverse_complement=rc).cpu()
        s = np.ascontiguousarray(seq[:, ::-1, ::-1]) if rc else seq
        for b0 in (0, 977, 1990):  # first, interior (chunk boundary at bin 1000), last
            lo, hi = max(b0 * 4000 - 112000, 0), min((b0 + 10) * 4000 + 112000, L)
            with torch.no_grad():
                ref = oracle.encoder_run(sd, torch.from_numpy(s[:, lo:hi]).transpose(1, 2))
            ref = ref[:, :, b0 - lo // 4000: b0 - lo // 4000 + 10]
            assert relerr(full[:, :, b0:b0 + 10].numpy(), ref.numpy()) <= TOL, (rc, b0)


def test_two_models_concurrent_cascades_are_consistent():
    """Default genomepredict runs two models x two strands = four independent cascades, interleaved on four
    CUDA streams here; every model's maps must equal its single-model run exactly (no cross-stream races)."""
    from orca_b200 import models, predict
    a, b = models.H1esc(seed=7), models.Hff(seed=9)
    seq = synthetic.random_sequence(1, 32_000_000, 105)
    mpos, wpos = 16_500_000, 16_000_000
    both = predict.genomepredict(seq, "chrS", mpos, wpos, models=[a, b])
    only_a = predict.genomepredict(seq, "chrS", mpos, wpos, models=[a])
    only_b = predict.genomepredict(seq, "chrS", mpos, wpos, models=[b])
    for i in range(6):
        assert np.array_equal(both["predictions"][0][i], only_a["predictions"][0][i]), i
        assert np.array_equal(both["predictions"][1][i], only_b["predictions"][0][i]), i
    g = gold("genomepredict_32mb")
    assert max(relerr(p, r) for p, r in zip(both["predictions"][0], g["predictions"])) <= TOL


def test_sharded_runner_256mb_matches_driver():
    """The sharded runner on a 256 Mb shell (world size 1) is the same computation as
    predict.genomepredict_256Mb: a full-size 256 Mb pass (320 reference blocks' worth of sequence)."""
    from orca_b200 import models, parallel, predict
    shell = models.H1esc_256M(seed=3)
    L = 256_000_000
    seq = synthetic.random_sequence(1, L, 21)
    nm = synthetic.normmat_256mb(chrlen_bins=7500)
    mpos, wpos, chrlen = 100_000_000, 128_000_000, 7500 * 32000
    ref = predict.genomepredict_256Mb(seq, "chrS", [nm], chrlen, mpos, wpos, models=[shell])
    runner = parallel.ShardedForward(shell, L, 0, 1, torch.device("cuda"))
    runner.set_background(nm, chrlen)
    runner.upload(torch.from_numpy(seq))
    maps = runner.forward(mpos, wpos).cpu().numpy()
    assert maps.shape == (4, 250, 250) and np.isfinite(maps).all()
    for i in range(4):
        assert relerr(maps[i], ref["predictions"][0][i]) <= 1e-6, i
    runner.cascade_mode = "serial"
    maps_s = runner.forward(mpos, wpos).cpu().numpy()
    assert np.array_equal(maps_s, maps)


# ---------------------------------------------------------------------------------------------------
# round 2: batch 4, full-size Orca-1Mb, hard cases for the single-pass fp16 stages, HCTnoc shell, strand-dependent
# 256 Mb stub, and the UNMODIFIED reference drivers on native shells
# ---------------------------------------------------------------------------------------------------
def test_batch4_golden():
    g = gold("encoder_b4_24k")
    m = native(modules.Encoder(), int(g["weight_seed"]))
    seq = synthetic.random_sequence(4, int(g["L"]), int(g["seq_seed"]), float(g["n_fraction"]))
    y = m(torch.from_numpy(seq).transpose(1, 2).cuda())
    assert relerr(y.cpu().numpy(), g["out"]) <= TOL
    g = gold("decoder_b4_40")
    d = native(modules.Decoder(upsample_mode="bilinear"), int(g["weight_seed"]))
    x = randn((4, 128, 40), int(g["x_seed"]), 0.5).cuda()
    distenc, yc = randn((4, 1, 40, 40), int(g["d_seed"])).cuda(), randn((4, 1, 20, 20), int(g["y_seed"])).cuda()
    y = d(x, distenc, yc)
    assert relerr(y.cpu().numpy(), g["out"]) <= TOL
    for b in range(4):  # a batch element does not depend on its neighbours
        assert torch.equal(d(x[b:b + 1], distenc[b:b + 1], yc[b:b + 1]), y[b:b + 1])


def test_net_1mb_golden():
    """Orca-1Mb at full size (README.md:204-219 screen path): Net.forward on 1 Mb vs the reference class."""
    g = gold("net_1mb")
    m = native(modules.Net(num_1d=32), int(g["weight_seed"]))
    seq = synthetic.random_sequence(1, 1_000_000, int(g["seq_seed"]), float(g["n_fraction"]))
    pred, p1d = m(torch.from_numpy(seq).transpose(1, 2).cuda())
    e2, e1 = relerr(pred.cpu().numpy(), g["out"]), relerr(p1d.cpu().numpy(), g["out_1d"])
    print("net_1mb relerr 2d %.2e 1d %.2e" % (e2, e1))
    assert e2 <= TOL and e1 <= TOL


@pytest.mark.parametrize("name", ["encoder_hard_alln", "encoder_hard_homopolymer", "encoder_hard_nruns", "encoder_hard_widebn",
                                  "encoder_hard_heavytail"])
def test_fp16_stages_on_hard_inputs(name):
    """the single-pass fp16 stages on inputs / weights that stress them -- all-N, homopolymer, long N
    runs, BatchNorm scales in [0.1, 10], heavy-tailed (Student-t, 3 d.o.f.) conv weights -- against the reference.  At the
    DEFAULT setting every case must meet the 1e-3 bar.  Three protections decide the format by themselves: the library's
    static conditioning check at module creation, the module's one-time self-calibration on its first input (fast vs
    fp32-grade on a 96 kb window), and the runtime fp16 range guard."""
    g = gold(name)
    m = modules.Encoder()
    m.load_state_dict(synthetic.fill_state_dict(m.state_dict(), int(g["weight_seed"]), recipe=str(g["recipe"])))
    m = m.eval().cuda()
    seq = synthetic.hard_sequence(1, int(g["L"]), int(g["seq_seed"]), str(g["seqkind"]))
    x = torch.from_numpy(seq).transpose(1, 2).cuda()
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        y = m(x)
    fell_back = m.options.get("encoder_fp16_stages") == 0            # the one-time self-calibration or the range guard said no
    effective = _lib.lib().orca_b200_module_get_option(m.native_handle(x.device), _lib.OPT_ENCODER_FP16_STAGES)
    static_off = not fell_back and effective == 0                       # module_create judged the weights ill-conditioned
    e = relerr(y.cpu().numpy(), g["out"])
    print(name, "relerr %.2e" % e, "| self-calibration / range guard -> fp32-grade" if fell_back else
          ("| ill-conditioned weights -> fp32-grade by default" if static_off else "| single-pass fp16 stages kept"))
    assert np.isfinite(y.cpu().numpy()).all() and e <= TOL
    assert fell_back == any("fp32-grade format" in str(i.message) for i in w)
    if str(g["recipe"]) in ("wide_bn", "heavy_tail"):
        assert fell_back or static_off
        # forced back to single-pass fp16 without the guard, this case misses the bar or trips the guard: the protection matters
        m.options["encoder_fp16_stages"] = 4
        bad = m(x, guard=False).cpu().numpy()
        fired = m.fp16_guard_fired()
        print(name, "forced fp16 stages: relerr %.2e, guard fired: %s" % (relerr(bad, g["out"]), fired))
        assert fired or relerr(bad, g["out"]) > 0.5 * TOL
    elif not fell_back:  # the fp32-grade path agrees as well
        m.options["encoder_fp16_stages"] = 0
        assert relerr(m(x).cpu().numpy(), g["out"]) <= 5e-5


def test_genomepredict_32mb_hctnoc_golden():
    """HCTnoc-like shell end to end (orca_models.py:335-446: Encoder2b, nearest upsampling, no Decoder_1m), zooming into
    the left edge of the window (crop index clipped to 0).  Fixture: the unmodified driver with a zero Decoder_1m term
    attached (as written, orca_predict.py:362 raises AttributeError on the reference HCTnoc shell)."""
    from orca_b200 import models, predict
    g = gold("genomepredict_32mb_hctnoc")
    shell = models.HCTnoc(seed=int(g["shell_seed"]))
    assert not hasattr(shell, "denet_1_pt") and isinstance(shell.net, modules.Encoder2b)
    seq = synthetic.random_sequence(1, 32_000_000, int(g["seq_seed"]))
    out = predict.genomepredict(seq, "chrS", int(g["mpos"]), int(g["wpos"]), models=[shell])
    assert out["start_coords"] == [int(v) for v in g["start_coords"]]
    errs = [relerr(p, r) for p, r in zip(out["predictions"][0], g["predictions"])]
    print("genomepredict 32 Mb HCTnoc-like relerr per level (32..1 Mb):", ["%.1e" % e for e in errs])
    assert max(errs) <= TOL


def test_genomepredict_256mb_strand_dependent_stub():
    """genomepredict_256Mb with a stub net0 that depends on its input (the two strands differ), a chromosome shorter than
    the window, an off-centre zoom; also checks end_coords and the float64 block-mean backgrounds it returns."""
    from orca_b200 import models, predict
    g = gold("genomepredict_256mb_stub2")
    shell = models.H1esc_256M(seed=int(g["shell_seed"]))

    class StubNet0(torch.nn.Module):
        def forward(self, x, reverse_complement=False):
            if reverse_complement:
                x = x.flip(1).flip(2)
            w = torch.from_numpy(np.random.default_rng(int(g["w_seed"])).standard_normal((128, 4)).astype(np.float32)).to(x.device)
            e = torch.einsum("kc,bcl->bkl", w, x)
            return 0.5 * (e + 0.5 * torch.roll(e, 1, 2) + 0.25 * torch.roll(e, -3, 2))
    shell.net0 = StubNet0()
    seq = synthetic.random_sequence(1, 64000, int(g["seq_seed"]), 0.01)
    nm = synthetic.normmat_256mb(chrlen_bins=int(g["chrlen_bins"]))
    out = predict.genomepredict_256Mb(seq, "chrS", [nm], int(g["chrlen"]), int(g["mpos"]), int(g["wpos"]), models=[shell])
    assert out["start_coords"] == [int(v) for v in g["start_coords"]]
    assert [int(v) for v in out["end_coords"]] == [int(v) for v in g["end_coords"]]
    errs = [relerr(p, r) for p, r in zip(out["predictions"][0], g["predictions"])]
    print("genomepredict 256 Mb (strand-dependent stub) relerr per level:", ["%.1e" % e for e in errs])
    assert max(errs) <= TOL
    nms = np.stack([np.stack([ns[l][0] for l in (256, 128, 64, 32)]) for ns in out["normmats"]])
    assert nms.dtype == np.float64 and nms.shape == g["normmats"].shape
    assert np.allclose(nms, g["normmats"], rtol=1e-12, atol=0)


def _reference_tree():
    """The reference sources as TEST DATA: /root/reference in the build container, or the git-ignored payload copy that
    tools/ship_reference.sh places under tests/_reference_payload/ for a gpurun call (never committed)."""
    for p in (os.environ.get("ORCA_REFERENCE"), "/root/reference", os.path.join(ROOT, "tests", "_reference_payload")):
        if p and os.path.isfile(os.path.join(p, "orca_predict.py")):
            return p
    return None


@pytest.mark.skipif(_reference_tree() is None, reason="reference sources not available on this box")
def test_unmodified_reference_drivers_on_native_shells():
    """The drop-in promise of INTEGRATION.md section 1: the UNMODIFIED orca_predict.genomepredict /
    genomepredict_256Mb, imported from the reference tree, driven with native shells (models=[shell], use_cuda=True),
    against the fixtures the same drivers produced on the reference's own modules."""
    import types
    ref = _reference_tree()
    if ref not in sys.path:
        sys.path.insert(0, ref)
    for name, attrs in {"selene_utils2": ["MemmapGenome", "Genomic2DFeatures"], "selene_sdk": [], "selene_sdk.sequences": ["Genome"],
                        "orca_utils": ["genomeplot", "genomeplot_256Mb", "StructuralChange2", "process_anno", "coord_round", "coord_clip"]}.items():
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for a in attrs:
                setattr(mod, a, type(a, (), {}))
            sys.modules[name] = mod
    sys.modules["selene_sdk"].sequences = sys.modules["selene_sdk.sequences"]
    import orca_predict
    from orca_b200 import models
    g = gold("genomepredict_32mb")
    shell = models.H1esc(seed=int(g["shell_seed"]))
    seq = synthetic.random_sequence(1, 32_000_000, int(g["seq_seed"]))
    n0 = _lib.launch_count()
    out = orca_predict.genomepredict(seq, "chrS", int(g["mpos"]), int(g["wpos"]), models=[shell], use_cuda=True)
    assert _lib.launch_count() > n0
    assert out["start_coords"] == [int(v) for v in g["start_coords"]]
    errs = [relerr(p, r) for p, r in zip(out["predictions"][0], g["predictions"])]
    print("UNMODIFIED orca_predict.genomepredict on a native H1esc shell: relerr per level", ["%.1e" % e for e in errs])
    assert max(errs) <= TOL
    # 256 Mb driver with the stub 4 kb encoding of the fixture
    g = gold("genomepredict_256mb_stub")
    shell = models.H1esc_256M(seed=int(g["shell_seed"]))

    class StubNet0(torch.nn.Module):
        def forward(self, x):
            e = np.random.default_rng(int(g["enc_seed"])).standard_normal((x.shape[0], 128, 64000)) * 0.5
            return torch.from_numpy(e.astype(np.float32)).cuda()
    shell.net0 = StubNet0()
    nm = synthetic.normmat_256mb(chrlen_bins=int(g["chrlen_bins"]))
    out = orca_predict.genomepredict_256Mb(synthetic.random_sequence(1, 4000, 107), "chrS", [nm], int(g["chrlen"]), int(g["mpos"]),
                                           int(g["wpos"]), models=[shell], use_cuda=True)
    assert out["start_coords"] == [int(v) for v in g["start_coords"]]
    errs = [relerr(p, r) for p, r in zip(out["predictions"][0], g["predictions"])]
    print("UNMODIFIED orca_predict.genomepredict_256Mb on a native H1esc_256M shell: relerr per level", ["%.1e" % e for e in errs])
    assert max(errs) <= TOL


def test_variant_windows_reuse_encoder_blocks():
    """the windows of one structural-variant call (reference + a deletion allele, as
    orca_predict.process_del builds them, orca_predict.py:1673-1794) through orca_b200.variants: encoder blocks upstream
    of the breakpoint are re-used from the cache, all four strand cascades run as one batched chain, and every map is
    BIT-identical to a separate genomepredict call on that window."""
    from orca_b200 import feeder, models, predict, variants
    shell = models.H1esc(seed=7)
    L = 32_000_000
    ref = synthetic.random_codes(1, L, 105)[0]
    bp, dlen = 20_300_123, 37_000           # breakpoint and deletion length (not a multiple of the 4 kb bin)
    filler = synthetic.random_codes(1, dlen, 106)[0]
    alt = np.concatenate([ref[:bp], ref[bp + dlen:], filler])   # window start unchanged: downstream bases shift by dlen
    assert alt.shape == ref.shape
    wins = [(ref, "chrS", bp, 16_000_000), (alt, "chrS", bp, 16_000_000)]
    outs, cache = variants.predict_variant_windows(wins, shell)
    n_blocks = L // 4000 // cache.block_bins
    upstream = (bp - variants.HALO_BP) // (cache.block_bins * 4000)          # blocks that cannot see the breakpoint
    assert cache.hits >= 2 * (upstream - 1) and cache.misses <= 2 * (2 * n_blocks - (upstream - 1))
    print("encoder blocks: %d re-used, %d encoded (of %d)" % (cache.hits, cache.misses, 4 * n_blocks))
    for (seq, mchr, mpos, wpos), out in zip(wins, outs):
        single = predict.genomepredict(seq[None], mchr, mpos, wpos, models=[shell])
        assert out["start_coords"] == single["start_coords"]
        for a, b in zip(out["predictions"][0], single["predictions"][0]):
            assert np.array_equal(a, b)
    # a second call with the same cache encodes nothing new
    h0, m0 = cache.hits, cache.misses
    outs2, _ = variants.predict_variant_windows(wins, shell, cache)
    assert cache.misses == m0 and cache.hits == h0 + 4 * n_blocks
    assert all(np.array_equal(a, b) for o, o2 in zip(outs, outs2) for a, b in zip(o["predictions"][0], o2["predictions"][0]))


def _nccl_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from orca_b200 import models, parallel
    g = gold("genomepredict_32mb")
    shell = models.H1esc(seed=int(g["shell_seed"]), device=dev)
    seq = torch.from_numpy(synthetic.random_sequence(1, 32_000_000, int(g["seq_seed"])))
    runner = parallel.ShardedForward(shell, 32_000_000, rank, world, dev)
    runner.upload(seq)
    enc_f = runner._encode(False).contiguous()
    maps = runner.forward(int(g["mpos"]), int(g["wpos"]))
    if rank == 0:
        ret["enc"], ret["maps"] = enc_f.cpu(), maps.cpu()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_forward_matches_single_gpu():
    """World size 2 over NCCL on real GPUs: the all-gathered sharded encoding is BIT-equal to the single-GPU encoding
    (the chunked encoder is bit-identical to the single pass) and the maps meet the reference fixture."""
    import torch.multiprocessing as mp
    from orca_b200 import models, parallel
    g = gold("genomepredict_32mb")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, 29650 + os.getpid() % 300, ret), nprocs=2, join=True)
    dev = torch.device("cuda", 0)
    shell = models.H1esc(seed=int(g["shell_seed"]), device=dev)
    seq = torch.from_numpy(synthetic.random_sequence(1, 32_000_000, int(g["seq_seed"])))
    single = parallel.ShardedForward(shell, 32_000_000, 0, 1, dev)
    single.upload(seq)
    assert torch.equal(single._encode(False).cpu(), ret["enc"])
    maps1 = single.forward(int(g["mpos"]), int(g["wpos"])).cpu().numpy()
    maps2 = ret["maps"].numpy()
    assert max(relerr(maps2[i], maps1[i]) for i in range(6)) <= 1e-4
    assert max(relerr(maps2[i], g["predictions"][i]) for i in range(6)) <= TOL
