"""Single-GPU workload for ncu at the BENCH shapes: one 32 Mb Encoder pass per strand (one chunk, all 7 stages, default
precision: stages 1-4 single-pass fp16; fp32 one-hot forward strand, packed-base reverse strand read in place), Encoder2 on
both strands, the strand-batched 6-level decoder cascade (+ Decoder_1m) of an H1esc-like shell, and the 256 Mb background
kernels (assembly of an 8000 x 8000 matrix + one block-mean level).  Usage (under gpurun):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py
  ncu --set full --clock-control none --import-source on -k regex:conv1d_tc -c 2 -o gpurun_out/prof python tools/ncu_target.py
NCU_SEQ_LEN overrides the sequence length; NCU_CASCADE=0 / NCU_ENCODER=0 skip a part (faster targeted captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orca_b200 import feeder, models, predict, synthetic

L = int(os.environ.get("NCU_SEQ_LEN", 32_000_000))
dev = torch.device("cuda:0")
shell = models.H1esc(seed=0, device=dev)
shell.net0._calibrated_version = shell.net0._handle_version  # no calibration launches in the capture
with torch.no_grad():
    if os.environ.get("NCU_ENCODER", "1") == "1":
        host = synthetic.random_sequence(1, L, 0)
        seq = torch.from_numpy(host).to(dev)
        packed = torch.from_numpy(feeder.from_onehot(host)).to(dev)
        shell.net0.native_handle(dev)
        shell.net0._calibrated_version = shell.net0._handle_version
        enc = shell.net0(seq.transpose(1, 2), guard=False)
        enc_r = shell.net0(packed, reverse_complement=True, guard=False)  # packed-base input, reverse strand read in place
        torch.cuda.synchronize()
        print("encoder done", enc.shape, enc_r.shape)
    if os.environ.get("NCU_CASCADE", "1") == "1":
        fin = []
        for seed in (1, 2):
            g = torch.Generator(device=dev).manual_seed(seed)
            e8000 = torch.randn(1, 128, 8000, device=dev, generator=g) * 0.5
            fin.append(dict(zip([1, 2, 4, 8, 16, 32], shell.net(e8000))))
        preds, _ = predict.cascade_32mb_lanes(shell, [(fin[0], False), (fin[1], True)], 16_300_000, 16_000_000)
        torch.cuda.synchronize()
    if os.environ.get("NCU_BACKGROUND", "1") == "1":
        sh256 = models.build_shell(__import__("orca_b200.modules", fromlist=["x"]), "h1esc_256m", 0)
        regions = [("chr1", 0, 128_000_000, "+"), ("chr2", 0, 128_000_000, "-")]
        nm = predict.assemble_background(regions, sh256.background_cis, sh256.background_trans, dev)
        lvl = predict.background_level(predict.prepare_background(nm, dev), 0, 32, with_mean=True)
        torch.cuda.synchronize()
print("done")
