"""Short single-GPU workload for ncu: one 4 Mb Encoder pass (all 7 stages, one chunk), Encoder2 and one
6-level decoder cascade of an H1esc-like shell.  Usage (under gpurun):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py
  ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -c 2 -o gpurun_out/prof python tools/ncu_target.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orca_b200 import models, predict, synthetic

L = int(os.environ.get("NCU_SEQ_LEN", 4_000_000))
dev = torch.device("cuda:0")
shell = models.H1esc(seed=0, device=dev)
seq = torch.from_numpy(synthetic.random_sequence(1, L, 0)).to(dev)
reps = int(os.environ.get("NCU_REPS", 1))
enc = torch.zeros(1, 128, 1)
with torch.no_grad():
    for _ in range(reps):
        enc = shell.net0(seq.transpose(1, 2))
        torch.cuda.synchronize()
    if L >= 32_000_000 or os.environ.get("NCU_CASCADE", "1") == "1":
        e = torch.from_numpy(synthetic.random_sequence(1, 32_000_000 // 4000 * 4, 1)).to(dev)  # dummy, unused
        enc8000 = torch.randn(1, 128, 8000, device=dev) * 0.5 if enc.shape[2] != 8000 else enc
        encs = dict(zip([1, 2, 4, 8, 16, 32], shell.net(enc8000)))
        preds, _ = predict.cascade_32mb(shell, encs, 1, 16_000_000, 16_000_000, False)
    torch.cuda.synchronize()
print("done", enc.shape)
