"""One-GPU validation + timing of the other BASELINE.json configs (not the bench workload):
  config 3: Hff-like 32 Mb model, batch 8 (modules called directly: Encoder -> Encoder2 -> 6-level cascade, B=8)
  config 4: H1esc_256M-like 256 Mb forward through orca_b200.predict.genomepredict_256Mb
  config 5: in-silico screen, 1 Mb windows through Encoder + level-1 Decoder (+ Decoder_1m), micro-batched
Prints one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from orca_b200 import models, predict, synthetic

dev = torch.device("cuda:0")
which = sys.argv[1:] or ["5", "3", "4"]


def timed(fn, n=2):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, r


with torch.no_grad():
    if "5" in which:
        shell = models.H1esc(seed=0, device=dev)
        W, MB = 64, 16  # 64 windows, 16 per micro-batch
        seq = torch.from_numpy(synthetic.random_sequence(MB, 1_000_000, 3)).to(dev)
        d1 = predict._log_normmat(shell, 1, dev)

        def screen():
            out = []
            for _ in range(W // MB):
                e = shell.net0(seq.transpose(1, 2))
                out.append(shell.denets[1](e, d1.expand(MB, -1, -1, -1)) + shell.denet_1_pt(e))
            return out
        t, r = timed(screen)
        print(json.dumps({"config": 5, "windows": W, "micro_batch": MB, "s_per_window": t / W, "windows_per_s": W / t,
                          "mbp_per_s": W / t, "finite": bool(torch.isfinite(r[0]).all())}))
    if "3" in which:
        shell = models.Hff(seed=1, device=dev)
        B, L = 8, 32_000_000
        seq = torch.from_numpy(synthetic.random_sequence(1, L, 4)).to(dev).expand(B, -1, -1)  # same sample x8 (4.1 GB if materialised)

        def batch8():
            e = shell.net0(seq.transpose(1, 2))
            encs = dict(zip([1, 2, 4, 8, 16, 32], shell.net(e)))
            return predict.cascade_32mb(shell, encs, B, L // 2, L // 2, False)[0]
        t, r = timed(batch8, 1)
        same = float((r[-1][0] - r[-1][7]).abs().max())
        print(json.dumps({"config": 3, "batch": B, "s_per_pass": t, "mbp_per_s": B * L / t / 1e6, "maps_per_s": 6 * B / t,
                          "batch_consistency_maxabs": same, "finite": bool(torch.isfinite(r[-1]).all())}))
    if "4" in which:
        shell = models.H1esc_256M(seed=0, device=dev)
        L = 256_000_000
        seq = synthetic.random_sequence(1, L, 5)
        nm = synthetic.normmat_256mb(chrlen_bins=7500)
        t, out = timed(lambda: predict.genomepredict_256Mb(seq, "chrS", [nm], 7500 * 32000, 100_000_000, 128_000_000, models=[shell]), 1)
        p = out["predictions"][0]
        print(json.dumps({"config": 4, "s_per_call_e2e": t, "mbp_per_s_e2e": 2 * L / t / 1e6, "start_coords": [int(s) for s in out["start_coords"]],
                          "finite": bool(all(np.isfinite(m).all() for m in p)), "absmax": float(max(np.abs(m).max() for m in p))}))
