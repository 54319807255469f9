"""world_size-2 CPU test (gloo) of the sequence-sharded forward's host logic: window slicing,
mirrored reverse-complement bins, all-gather placement.  The device encoder is replaced by the
oracle (CPU) behind the same net0(...) keyword interface; the result must equal the un-sharded
oracle encoding on both strands."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

L = 480000


class OracleNet0(torch.nn.Module):
    """CPU stand-in with Encoder.forward's extended signature (bin_range/out/reverse_complement/window)."""

    def __init__(self, sd):
        super().__init__()
        self.sd = sd
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x, bin_range=None, out=None, reverse_complement=False, window=None):
        import orca_oracle as oracle
        pos0, Ltot = window
        b0, b1 = bin_range
        full = torch.zeros(1, 4, Ltot)
        full[:, :, pos0:pos0 + x.shape[2]] = x          # positions outside the window stay 0 (never read)
        if reverse_complement:
            full = torch.flip(full, [1, 2])
        lo, hi = max(b0 * 4000 - 112000, 0), min(b1 * 4000 + 112000, Ltot)
        with torch.no_grad():
            y = oracle.encoder_run(self.sd, full[:, :, lo:hi])
        out[0, b0:b1] = y[0, :, b0 - lo // 4000: b1 - lo // 4000].T
        return out.transpose(1, 2)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from orca_b200 import modules, parallel, synthetic
    sd = synthetic.fill_state_dict(modules.Encoder().state_dict(), 4)
    seq = torch.from_numpy(synthetic.random_sequence(1, L, 6, 0.01))
    shell = torch.nn.Module()
    shell.net0 = OracleNet0(sd)
    fwd = parallel.ShardedForward(shell, L, rank, world, torch.device("cpu"))
    fwd.upload(seq)
    assert fwd.window.shape[1] < L  # each rank holds only its window
    ef, er = fwd._encode(False), fwd._encode(True)
    if rank == 0:
        ret["fwd"], ret["rev"] = ef.clone(), er.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_encode_matches_unsharded():
    import orca_oracle as oracle
    from orca_b200 import modules, synthetic
    world, port = 2, 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    sd = synthetic.fill_state_dict(modules.Encoder().state_dict(), 4)
    seq = synthetic.random_sequence(1, L, 6, 0.01)
    with torch.no_grad():
        ref_f = oracle.encoder_run(sd, torch.from_numpy(seq).transpose(1, 2))
        ref_r = oracle.encoder_run(sd, torch.from_numpy(np.ascontiguousarray(seq[:, ::-1, ::-1])).transpose(1, 2))
    assert torch.allclose(ret["fwd"], ref_f, atol=2e-6)
    assert torch.allclose(ret["rev"], ref_r, atol=2e-6)


def _cascade_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import fakes
    from orca_b200 import parallel
    shell = fakes.FakeShell("h1esc")
    seq = torch.from_numpy(fakes.stub_sequence(32_000_000, 111))
    fwd = parallel.ShardedForward(shell, 32_000_000, rank, world, torch.device("cpu"))
    fwd.upload(seq)
    maps = fwd.forward(16_500_000, 16_000_000)
    assert (maps is None) == (rank != 0)
    if rank == 0:
        ret["maps"] = maps.clone()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_forward_distribution_matches_single_process(world):
    """The WHOLE sharded forward on `world` ranks (gloo): sharded encode + all-gather, the strand cascades on ranks 0 / 1,
    the Decoder_1m terms on ranks 2 / 3 when they exist, the batched point-to-point collection on rank 0 and the strand
    average -- against the single-process driver on the same stand-in networks (tests/fakes.py), whose maps depend on
    every crop index of both strands."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fakes
    from orca_b200 import predict
    port = 31500 + os.getpid() % 2000 + world
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cascade_worker, args=(world, port, ret), nprocs=world, join=True)
    shell = fakes.FakeShell("h1esc")
    seq = fakes.stub_sequence(32_000_000, 111)
    ref = predict._genomepredict_on(torch.device("cpu"), seq, "chrS", 16_500_000, 16_000_000, [shell])
    want = np.stack(ref["predictions"][0])
    got = ret["maps"].numpy()
    assert got.shape == want.shape == (6, 250, 250)
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
