"""Decoder timing scan on one GPU: time per Decoder.forward for map sizes S and batch B (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orca_b200 import modules, synthetic
dev = torch.device("cuda:0")
dec = synthetic.init_module(modules.Decoder(upsample_mode="bilinear"), 2).to(dev)
for S, B in [(32, 1), (64, 1), (128, 1), (250, 1), (250, 2), (250, 4), (250, 8)]:
    x = torch.randn(B, 128, S, device=dev) * 0.5
    d = torch.randn(B, 1, S, S, device=dev)
    y = torch.randn(B, 1, S // 2, S // 2, device=dev)
    with torch.no_grad():
        for _ in range(3):
            dec(x, d, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            dec(x, d, y)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("S=%3d B=%d  %.3f ms/call  %.2f us per conv  %.3f ms per map" % (S, B, ms, ms * 1e3 / 118, ms / B))
