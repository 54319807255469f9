"""Phase timing of one 32 Mb step on one GPU (diagnostic): encoders, Encoder2, decoder cascades (2 streams)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from orca_b200 import models, parallel, predict, synthetic
dev = torch.device("cuda:0")
L = 32_000_000
shell = models.H1esc(seed=0, device=dev)
run = parallel.ShardedForward(shell, L, 0, 1, dev)
run.upload(torch.from_numpy(synthetic.random_sequence(1, L, 0)).pin_memory())
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r
with torch.no_grad():
    t_enc, ef = timed(lambda: run._encode(False))
    t_encr, er = timed(lambda: run._encode(True))
    t_net, encs = timed(lambda: dict(zip([1, 2, 4, 8, 16, 32], shell.net(ef))))
    t_c1, _ = timed(lambda: predict.cascade_32mb(shell, encs, 1, L // 2, L // 2, False))
    def both():
        return predict.run_concurrent([lambda: predict.cascade_32mb(shell, encs, 1, L // 2, L // 2, False),
                                       lambda: predict.cascade_32mb(shell, encs, 1, L // 2, L // 2, True)], dev)
    t_c2, _ = timed(both)
    t_all, _ = timed(lambda: run.forward(L // 2, L // 2))
print("encoder fwd %.2f ms, rev %.2f ms | Encoder2 %.2f ms | one cascade %.2f ms | two cascades concurrently %.2f ms | full step %.2f ms"
      % (t_enc, t_encr, t_net, t_c1, t_c2, t_all))
