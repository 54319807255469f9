"""Event trace of one decoder-program launch (diagnostic): CTA 0 logs clock64 at the hand-over points of its producer,
MMA-issuer and epilogue roles; prints per-tile intervals for a few mid-program layers."""
import ctypes, os, sys
os.environ["ORCA_B200_DEC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from orca_b200 import _lib, modules, synthetic
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dec = synthetic.init_module(modules.Decoder(upsample_mode="bilinear"), 2).to(dev)
x = torch.randn(B, 128, 250, device=dev) * 0.5
d = torch.randn(B, 1, 250, 250, device=dev)
y = torch.randn(B, 1, 125, 125, device=dev)
with torch.no_grad():
    for _ in range(3):
        dec(x, d, y)
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * (4 * 2048))()
fn = ctypes.CDLL(_lib.LIB_PATH).orca_b200_debug_trace
fn.argtypes = [ctypes.c_void_p]
n = fn(buf)
a = np.frombuffer(buf, dtype=np.int64).reshape(4, 2048)
names = {1: "P.layer", 2: "P.slot_free", 3: "P.issued", 10: "M.layer", 11: "M.tile", 12: "M.acc_free", 13: "M.a_full", 14: "M.issued",
         20: "E.layer", 21: "E.tile", 22: "E.acc_full", 23: "E.arrived", 30: "B.sync", 31: "B.arrive", 32: "B.released"}
ev = []
for r in range(4):
    for v in a[r]:
        if v == 0:
            continue
        ev.append((int(v & 0xFFFFFFFFFFFF), int(v >> 48), r))
ev.sort()
t0 = ev[0][0]
# print the events between the 4th and 6th layer boundary
bounds = [t for t, c, r in ev if c == 32]
lo, hi = (bounds[3], bounds[5]) if len(bounds) > 5 else (ev[0][0], ev[-1][0])
prev = lo
for t, c, r in ev:
    if lo <= t <= hi:
        print("%8d (+%5d)  %s" % (t - lo, t - prev, names.get(c, c)))
        prev = t
