"""Single-GPU timing of how the two strand cascades are scheduled (diagnostic): four chains on four streams
(one launch per conv) vs the two strands as batch elements of one chain (one persistent program kernel per decoder)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orca_b200 import models, parallel, predict, synthetic
dev = torch.device("cuda:0")
which = sys.argv[1:] or ["32mb"]


def timed(fn, n=3):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


with torch.no_grad():
    if "32mb" in which:
        L = 32_000_000
        shell = models.H1esc(seed=0, device=dev)
        run = parallel.ShardedForward(shell, L, 0, 1, dev)
        run.upload(torch.from_numpy(synthetic.random_sequence(1, L, 0)).pin_memory())
        res = {}
        for mode in ("streams", "batch"):
            run.cascade_mode = mode
            t, r = timed(lambda: run.forward(L // 2 + 300_000, L // 2))
            res[mode] = (t, r)
        ef, er = run._encode(False), run._encode(True)
        fin = [dict(zip([1, 2, 4, 8, 16, 32], shell.net(e))) for e in (ef, er)]
        t_lanes, _ = timed(lambda: predict.cascade_32mb_lanes(shell, [(fin[0], False), (fin[1], True)], L // 2, L // 2))
        t_one, _ = timed(lambda: predict.cascade_32mb(shell, fin[0], 1, L // 2, L // 2, False))
        d = (res["streams"][1] - res["batch"][1]).abs().max().item()
        print("32mb: step streams %.2f ms | step batch %.2f ms | maxabs diff %.3g | batched 2-strand cascade alone %.2f ms | "
              "one strand cascade alone (program) %.2f ms" % (res["streams"][0], res["batch"][0], d, t_lanes, t_one))
    if "256mb" in which:
        L = 256_000_000
        shell = models.H1esc_256M(seed=0, device=dev)
        run = parallel.ShardedForward(shell, L, 0, 1, dev)
        run.set_background(synthetic.normmat_256mb(chrlen_bins=7500), 7500 * 32000)
        run.upload(torch.from_numpy(synthetic.random_sequence(1, L, 0)))
        res = {}
        for mode in ("streams", "batch"):
            run.cascade_mode = mode
            t, r = timed(lambda: run.forward(100_000_000, 128_000_000), 2)
            res[mode] = (t, r)
        d = (res["streams"][1] - res["batch"][1]).abs().max().item()
        print("256mb: step streams %.2f ms | step batch %.2f ms | maxabs diff %.3g" % (res["streams"][0], res["batch"][0], d))
