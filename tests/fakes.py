"""Cheap stand-in networks for driver-logic tests (test infrastructure).

The cascade index arithmetic of orca_predict.genomepredict / genomepredict_256Mb (crop windows, clip edges, chrlen
bounds, strand mirroring) depends only on host integers, never on tensor values.  These fakes make every output
pixel depend on WHERE each level was cropped, on the strand, on the background window and on the coarse crop, so a
full driver pass with them costs milliseconds and any index slip changes the maps.  oracle/make_golden.py runs the
UNMODIFIED reference drivers on a shell of these fakes (fixture `cascade_index_cases.npz`); tests/test_host.py runs
orca_b200.predict's host logic on the same shell and compares.  All float32 CPU, deterministic."""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F


class FakeNet0(nn.Module):
    """(B, 4, L) -> (B, 128, L / bin): base-composition track modulated by the position (strand-asymmetric)."""

    def __init__(self, bin_bp):
        super().__init__()
        self.bin_bp = bin_bp
        self.register_buffer("w", torch.tensor([0.1, 0.2, 0.3, 0.4]))
        self.register_buffer("ch", torch.linspace(0.5, 1.5, 128))

    def forward(self, x, reverse_complement=False, bin_range=None, out=None, window=None):
        """Reference call: forward(x).  The keyword extensions are those of orca_b200.modules.Encoder.forward (strand
        flip in place, a shard's bin range / window / output buffer) so that the sharded runner can drive this fake."""
        if window is not None:  # a shard holds positions [pos0, pos0 + n) of an L_total sequence
            pos0, L_total = window
            full = torch.zeros((x.shape[0], 4, L_total), dtype=x.dtype)
            full[:, :, pos0:pos0 + x.shape[2]] = x
            x = full
        if reverse_complement:  # orca_b200's native signature; the reference passes the flipped copy instead
            x = x.flip(1).flip(2)
        B, _, L = x.shape
        P = L // self.bin_bp
        if bin_range is not None:
            enc = self.forward(x).transpose(1, 2)  # (B, P, 128)
            b0, b1 = bin_range
            out[:, b0:b1] = enc[:, b0:b1]
            return out.transpose(1, 2)
        t = (x * self.w[None, :, None]).sum(1).reshape(B, P, self.bin_bp).mean(2)
        pos = torch.arange(P, dtype=torch.float32) / P
        e = t * (1.0 + pos)[None, :] + 0.05 * torch.sin(pos * 40.0)[None, :]
        return e[:, None, :] * self.ch[None, :, None]


class FakePools(nn.Module):
    """Encoder2 / Encoder3 stand-in: the input average-pooled by 1, 2, 4, ... (n_out tensors, finest first).
    first > 1 pre-pools (Encoder2 in the 256 Mb route: only [-1] is used, orca_predict.py:675-683)."""

    def __init__(self, n_out, first=1):
        super().__init__()
        self.n_out, self.first = n_out, first
        self.register_buffer("dummy", torch.zeros(1))

    def forward(self, x, coarsest_only=False):
        if self.first > 1:
            x = F.avg_pool1d(x, self.first)
        return [F.avg_pool1d(x, 1 << i) if i else x for i in range(self.n_out)]


class FakeDecoder(nn.Module):
    def __init__(self, k):
        super().__init__()
        self.k = float(k)
        self.register_buffer("dummy", torch.zeros(1))

    def forward(self, x, distenc, y=None):
        a = x[:, 0, :]
        m = a[:, :, None] + 0.5 * a[:, None, :] + 0.01 * self.k * distenc[:, 0]
        if y is not None:
            m = m + 0.3 * F.interpolate(y, scale_factor=2, mode="nearest")[:, 0]
        return m[:, None]


class FakeDecoder1m(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("dummy", torch.zeros(1))

    def forward(self, x):
        return (0.2 * x[:, 1, :, None] * x[:, 1, None, :])[:, None]


class FakeShell(nn.Module):
    def __init__(self, kind):
        super().__init__()
        from orca_b200 import synthetic
        self.kind = kind
        if kind == "h1esc":
            self.net0 = FakeNet0(4000)
            self.net = FakePools(6)
            levels = [1, 2, 4, 8, 16, 32]
            self.denet_1_pt = FakeDecoder1m()
            self.normmats, self.epss = synthetic.normmats_32mb()
        else:  # "h1esc_256m": 1 bp of the stub sequence = one 4 kb bin (64000 "bp" stand for 256 Mb)
            self.net0 = FakeNet0(1)
            self.net1 = FakePools(6, first=1)
            self.net = FakePools(4)
            levels = [32, 64, 128, 256]
        for i, lvl in enumerate(levels):
            setattr(self, "denet_%d" % lvl, FakeDecoder(1.0 + i))
        self._levels = levels

    @property
    def denets(self):
        return {lvl: getattr(self, "denet_%d" % lvl) for lvl in self._levels}


CASES_32MB = [  # (mpos, wpos): centre, default offset, both clip edges, shifted windows
    (16_000_000, 16_000_000), (16_500_000, 16_000_000), (300_000, 16_000_000), (31_900_000, 16_000_000),
    (5_000_000, 20_000_000), (27_123_456, 14_000_000), (15_999_999, 16_000_000),
]
# (mpos, wpos, chrlen): chrlen < 128 Mb makes bounds[0] >= bounds[1] at level 256 (orca_predict.py:818-825)
CASES_256MB = [
    (100_000_000, 128_000_000, 6000 * 32000), (128_000_000, 128_000_000, 8000 * 32000), (10_000_000, 128_000_000, 100_000_000),
    (240_000_000, 128_000_000, 250_000_000), (60_000_000, 128_000_000, 70_000_000),
]


def stub_sequence(L, seed):
    from orca_b200 import synthetic
    return synthetic.random_sequence(1, L, seed)
