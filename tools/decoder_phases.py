"""Where a decoder layer's time goes (diagnostic; one GPU): Decoder.forward at S=250 timed with parts of the program
kernel switched off through ORCA_B200_DEC_EXPERIMENT (bit 0 epilogue without global loads/stores, bit 1 producer copies
16 B per run chunk, bit 2 no MMAs, bit 3 no TMEM loads).  The outputs of the switched-off runs are garbage."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from orca_b200 import modules, synthetic
    dev = torch.device("cuda:0")
    dec = synthetic.init_module(modules.Decoder(upsample_mode="bilinear"), 2).to(dev)
    out = []
    for B in (1, 2):
        x = torch.randn(B, 128, 250, device=dev) * 0.5
        d = torch.randn(B, 1, 250, 250, device=dev)
        y = torch.randn(B, 1, 125, 125, device=dev)
        with torch.no_grad():
            for _ in range(3):
                dec(x, d, y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                dec(x, d, y)
            e1.record()
            torch.cuda.synchronize()
        out.append("B=%d %.3f ms (%.2f us/conv)" % (B, e0.elapsed_time(e1) / 5, e0.elapsed_time(e1) / 5 * 1e3 / 118))
    print("experiment=%s: %s" % (os.environ.get("ORCA_B200_DEC_EXPERIMENT", "0"), " | ".join(out)))
else:
    exps = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 9, 3, 6, 7, 15]
    for exp in exps:
        env = dict(os.environ, ORCA_B200_DEC_EXPERIMENT=str(exp))
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, timeout=120)
