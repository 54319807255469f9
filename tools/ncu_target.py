"""Short single-GPU workload for ncu: one 4 Mb Encoder pass per strand (all 7 stages, one chunk, default precision:
stages 1-3 single-pass fp16), Encoder2 and the strand-batched 6-level decoder cascade (+ Decoder_1m) of an
H1esc-like shell -- the kernels of one bench step at a size ncu can replay.  Usage (under gpurun):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py
  ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -c 2 -o gpurun_out/prof python tools/ncu_target.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orca_b200 import feeder, models, predict, synthetic

L = int(os.environ.get("NCU_SEQ_LEN", 4_000_000))
dev = torch.device("cuda:0")
shell = models.H1esc(seed=0, device=dev)
host = synthetic.random_sequence(1, L, 0)
seq = torch.from_numpy(host).to(dev)
packed = torch.from_numpy(feeder.from_onehot(host)).to(dev)
with torch.no_grad():
    enc = shell.net0(seq.transpose(1, 2))
    enc_r = shell.net0(packed, reverse_complement=True)  # packed-base input, reverse strand read in place
    torch.cuda.synchronize()
    if os.environ.get("NCU_CASCADE", "1") == "1":
        fin = []
        for seed in (1, 2):
            g = torch.Generator(device=dev).manual_seed(seed)
            e8000 = torch.randn(1, 128, 8000, device=dev, generator=g) * 0.5
            fin.append(dict(zip([1, 2, 4, 8, 16, 32], shell.net(e8000))))
        preds, _ = predict.cascade_32mb_lanes(shell, [(fin[0], False), (fin[1], True)], 16_300_000, 16_000_000)
    torch.cuda.synchronize()
print("done", enc.shape, enc_r.shape)
