"""Numerics simulation behind DESIGN.md section 3 (CPU only, minutes): what each reduced-pass tensor-core scheme does to
the Encoder output and to a Decoder map, with seeded synthetic weights, operands quantised the way the
kernels would (fp16 / bf16 hi+lo / fp8 e4m3 correction terms) and products accumulated in float64.

    python tools/sim_precision.py encoder [L=400000]     # fp32, bf16x3, fp16x2, fp16+fp8, fp16x1, mixN (fp16x1 in stages 1..N)
    python tools/sim_precision.py decoder [S=96]         # bf16x3, fp16x1, fp16x2w, fp16x2a, fp16+fp8

Uses the oracle (test infrastructure) as the fp32 reference; nothing in the product imports this file."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.nn.functional as F
import orca_oracle as oracle
from orca_b200 import modules, synthetic

torch.set_num_threads(os.cpu_count() or 1)
q16 = lambda x: x.half().float()
qb16 = lambda x: x.bfloat16().float()
q8 = lambda x: x.to(torch.float8_e4m3fn).float()
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())


def products(c, x, w, scheme):
    """c(a, b) = exact convolution; returns the scheme's sum of tensor-core products."""
    if scheme == "fp32":
        return c(x, w)
    if scheme == "bf16x3":
        xh, wh = qb16(x), qb16(w)
        xl, wl = qb16(x - xh), qb16(w - wh)
        return c(xh, wh) + c(xl, wh) + c(xh, wl)
    if scheme == "fp16x1":
        return c(q16(x), q16(w))
    if scheme == "fp16x2" or scheme == "fp16x2w":  # activations rounded, weights split
        xh, wh = q16(x), q16(w)
        return c(xh, wh) + c(xh, q16(w - wh))
    if scheme == "fp16x2a":  # activations split, weights rounded
        xh, wh = q16(x), q16(w)
        return c(xh, wh) + c(q16(x - xh), wh)
    if scheme == "fp16+fp8":  # fp16 main product + two fp8 (e4m3) correction products with power-of-two scaling
        xh, wh = q16(x), q16(w)
        xl, wl = x - xh, w - wh
        p = 12
        q = int(np.floor(np.log2(256.0 / float(w.abs().max()))))
        t = q + 12
        return c(xh, wh) + (c(q8(xl * 2.0 ** p), q8(wh * 2.0 ** q)) + c(q8(xh), q8(wl * 2.0 ** t))) * 2.0 ** -(p + q)
    raise ValueError(scheme)


def fold(sd, conv, bn, nd):
    w, b = sd[conv + ".weight"].double(), sd[conv + ".bias"].double()
    if bn is not None:
        sc = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + 1e-5)
        w = w * sc.reshape((-1,) + (1,) * nd)
        b = (b - sd[bn + ".running_mean"].double()) * sc + sd[bn + ".bias"].double()
    return w.float(), b.float()


# ---------------------------------------------------------------- encoder
def conv1(x, w, scheme):
    return products(lambda a, b: F.conv1d(a.double(), b.double(), padding=4).float(), x, w, scheme)


def encoder(sd, x, scheme_all):
    pools = [None, 4, 4, 5, 5, 5, 2]
    cur, out = x, None
    for k in range(1, 8):
        scheme = scheme_all
        if scheme_all.startswith("mix"):
            scheme = "fp16x1" if k <= int(scheme_all[3:]) else "bf16x3"
        if pools[k - 1]:
            cur = F.max_pool1d(cur, pools[k - 1], pools[k - 1])
        o = 0 if k == 1 else 1
        w, b = fold(sd, "lconv%d.%d" % (k, o), "lconv%d.%d" % (k, o + 1), 2); h = conv1(cur, w, scheme) + b[None, :, None]
        w, b = fold(sd, "lconv%d.%d" % (k, o + 2), "lconv%d.%d" % (k, o + 3), 2); lout = conv1(h, w, scheme) + b[None, :, None]
        w, b = fold(sd, "conv%d.0" % k, "conv%d.1" % k, 2); h = F.relu(conv1(lout, w, scheme) + b[None, :, None])
        w, b = fold(sd, "conv%d.3" % k, "conv%d.4" % k, 2); out = F.relu(conv1(h, w, scheme) + b[None, :, None])
        cur = out + lout
    return out


def run_encoder(L):
    sd = synthetic.fill_state_dict(modules.Encoder().state_dict(), 11)
    x = torch.from_numpy(synthetic.random_sequence(1, L, 102)).transpose(1, 2).contiguous()
    S = L // 4000
    mats, _ = synthetic.normmats_32mb()
    dist = torch.log(torch.FloatTensor(mats[1][:S, :S][None, None]))
    sdd = synthetic.fill_state_dict(modules.Decoder(upsample_mode="bilinear").state_dict(), 15)
    with torch.no_grad():
        ref = oracle.encoder_run(sd, x)
        p0 = oracle.decoder_forward(sdd, ref, dist, None, "bilinear")
        for s in ["fp32", "bf16x3", "fp16x2", "fp16+fp8", "fp16x1", "mix1", "mix2", "mix3", "mix4", "mix5"]:
            y = encoder(sd, x, s)
            p = oracle.decoder_forward(sdd, y, dist, None, "bilinear")  # propagate the encoder error through an exact decoder
            print("%-9s encoder relerr %.2e   -> decoder map relerr %.2e" % (s, rel(y, ref), rel(p, p0)), flush=True)


# ---------------------------------------------------------------- decoder
DIL = [1, 2, 4, 8, 16, 32, 64] * 4


def conv2(x, w, d, scheme):
    return products(lambda a, b: F.conv2d(a.double(), b.double(), padding=d, dilation=d).float(), x, w, scheme)


def lin(x, sd, p, o, d, scheme):
    w, b = fold(sd, "%s.%d" % (p, o), "%s.%d" % (p, o + 1), 3); x = conv2(x, w, d, scheme) + b[None, :, None, None]
    w, b = fold(sd, "%s.%d" % (p, o + 2), "%s.%d" % (p, o + 3), 3); return conv2(x, w, d, scheme) + b[None, :, None, None]


def relu2(x, sd, p, d, scheme):
    w, b = fold(sd, p + ".0", p + ".1", 3); x = F.relu(conv2(x, w, d, scheme) + b[None, :, None, None])
    w, b = fold(sd, p + ".3", p + ".4", 3); return F.relu(conv2(x, w, d, scheme) + b[None, :, None, None])


def decoder(sd, x, distenc, y, scheme, head="bf16x3"):
    mat = torch.cat([x[:, :, :, None] + x[:, :, None, :], distenc], 1)
    mat = lin(mat, sd, "lcombinerD", 0, 1, head)
    cur = relu2(mat, sd, "combinerD", 1, head) + mat
    for i, d in enumerate(DIL):
        if i == 0:
            if y is not None:
                cur = torch.cat([cur, F.interpolate(y, scale_factor=(2, 2), mode="bilinear", align_corners=False)], 1)
                cur = lin(cur, sd, "lcombiner", 1, 1, head)
                cur = relu2(cur, sd, "combiner", 1, head) + cur
            else:
                cur = lin(cur, sd, "lconvtwos.0", 1, d, scheme)
                cur = relu2(cur, sd, "convtwos.0", d, scheme) + cur
        else:
            cur = lin(cur, sd, "lconvtwos.%d" % i, 0, d, scheme) + cur
            cur = relu2(cur, sd, "convtwos.%d" % i, d, scheme) + cur
    return oracle._final(cur, sd)


def run_decoder(S):
    rng = np.random.default_rng(1)
    x = torch.from_numpy((rng.standard_normal((1, 128, S)) * 0.5).astype(np.float32))
    mats, _ = synthetic.normmats_32mb()
    dist = torch.log(torch.FloatTensor(mats[4][:S, :S][None, None]))
    yc = torch.from_numpy(rng.standard_normal((1, 1, S // 2, S // 2)).astype(np.float32))
    sd = synthetic.fill_state_dict(modules.Decoder(upsample_mode="bilinear").state_dict(), 15)
    with torch.no_grad():
        for y in (None, yc):
            ref = oracle.decoder_forward(sd, x, dist, y, "bilinear")
            for s in ["bf16x3", "fp16x1", "fp16x2w", "fp16x2a", "fp16+fp8"]:
                print("%-8s %-8s decoder relerr %.2e" % ("coarse" if y is not None else "nocoarse", s, rel(decoder(sd, x, dist, y, s), ref)), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "encoder"
    if what == "encoder":
        run_encoder(int(sys.argv[2]) if len(sys.argv) > 2 else 400000)
    else:
        run_decoder(int(sys.argv[2]) if len(sys.argv) > 2 else 96)
