#!/usr/bin/env python
"""
bench.py -- headline benchmark of the Orca forward hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W              # our CUDA path
    python bench.py --impl reference --steps K --warmup W      # the reference algorithm on the host CPU

Workload (BASELINE.json configs[1]): H1esc-like 32 Mb multiscale forward, batch 1, synthetic one-hot
sequence, random-init weights.  One step = one genomepredict pass of one model: both strands
(forward + reverse complement) x [Encoder (32 Mb -> 8000 bins) + Encoder2 + 6-level Decoder cascade
+ Decoder_1m] = 64 Mbp encoded and 12 contact maps decoded (6 after strand averaging).

  value   Mbp/s with the sequence already resident in HBM (device-timed, CUDA events)
  e2e     same metric through orca_b200.predict.genomepredict with the HOST fp32 (1, L, 4) array in
          pinned memory: the H2D upload and the D2H read of the 6 maps are inside the timed region
  N > 1   the 8000 4-kb bins are sharded over the ranks (112 kb halo recompute), the encodings are
          all-gathered over NCCL, strand cascades run on different ranks ("scaling": "strong")
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEQ_LEN = 32_000_000
METRIC = "Mbp/s encoded (32 Mb genomepredict forward, both strands; contact maps/s reported beside it)"
FLOP_PER_BP_ENCODER = 465_555.4       # SURVEY.md 8d (monolithic-convolution definition)
FLOP_PER_STRAND = 16.84e12            # Encoder 14.898 + Encoder2 0.0274 + decoders 1.911 TFLOP
MAPS_PER_STEP = 12                    # raw decoder maps per step (6 levels x 2 strands)


def load_peaks():
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        burst = float(p["bf16_tflops"])
        return {"hbm_gbs": float(p.get("hbm_gbs", fallback["hbm_gbs"])), "bf16_tflops": burst,
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", burst)), "source": "measured"}
    except (OSError, ValueError, KeyError, TypeError):
        return fallback


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ----------------------------------------------------------------------------------------------------
# reference algorithm on the host CPU (oracle port; test/baseline infrastructure)
# ----------------------------------------------------------------------------------------------------
CPU_BLOCKS = 20  # encoder blocks per CPU sample (of the 40 per strand at 32 Mb)
CPU_DECODERS = 6  # Decoder calls per CPU sample (all 6 of a strand)
CPU_SAMPLE = ("%d of 40 encoder blocks per strand (912 kb each incl. the 112 kb halo), Encoder2@8000, %d of 6 Decoder calls, "
              "Decoder_1m" % (CPU_BLOCKS, CPU_DECODERS))


def cpu_sample(threads):
    """Time a bounded sample (~10-20 s of CPU work on a 16-core host) of the workload with the oracle port on `threads` host threads
    and extrapolate to one full step.  Sample: CPU_BLOCKS 800 kb encoder blocks with their 112 kb halo
    (orca_modules.py:957-977), Encoder2 on 8000 bins, CPU_DECODERS Decoder calls and one Decoder_1m call; returns
    per-unit seconds (block, Encoder2, Decoder, Decoder_1m)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orca_oracle as oracle
    from orca_b200 import modules, synthetic
    torch.set_num_threads(threads)
    sd_e = synthetic.fill_state_dict(modules.Encoder().state_dict(), 0)
    sd_n = synthetic.fill_state_dict(modules.Encoder2().state_dict(), 1)
    sd_d = synthetic.fill_state_dict(modules.Decoder(upsample_mode="bilinear").state_dict(), 10)
    sd_m = synthetic.fill_state_dict(modules.Decoder_1m().state_dict(), 3)
    x = torch.from_numpy(synthetic.random_sequence(1, 912000, 0)).transpose(1, 2)
    rng = np.random.default_rng(0)
    e = torch.from_numpy(rng.standard_normal((1, 128, 8000)).astype(np.float32) * 0.5)
    d = torch.from_numpy(rng.standard_normal((1, 1, 250, 250)).astype(np.float32))
    y = torch.from_numpy(rng.standard_normal((1, 1, 125, 125)).astype(np.float32))

    def one():
        with torch.no_grad():
            t0 = time.perf_counter()
            for _ in range(CPU_BLOCKS):
                oracle.encoder_run(sd_e, x)
            t1 = time.perf_counter(); encs = oracle.encoder2_forward(sd_n, e)
            t2 = time.perf_counter()
            for _ in range(CPU_DECODERS):
                oracle.decoder_forward(sd_d, encs[-1], d, y, "bilinear")
            t3 = time.perf_counter(); oracle.decoder_1m_forward(sd_m, encs[-1])
            t4 = time.perf_counter()
        return (t1 - t0) / CPU_BLOCKS, t2 - t1, (t3 - t2) / CPU_DECODERS, t4 - t3
    return one


def cpu_extrapolate(t_block, t_enc2, t_dec, t_d1m):
    """Seconds for one full step (2 strands) from the sample timings: 40 blocks of 800 kb per strand."""
    per_strand = 40 * t_block + t_enc2 + 6 * t_dec + t_d1m
    return 2 * per_strand


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    one = cpu_sample(threads)
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    acc = np.zeros(4)
    for _ in range(args.steps):
        acc += np.array(one())
    wall = time.perf_counter() - t0
    tb, te, td, tm = acc / args.steps
    full = cpu_extrapolate(tb, te, td, tm)
    value = 2 * SEQ_LEN / full / 1e6
    sample = ("per step: " + CPU_SAMPLE + "; extrapolated linearly to 2 strands x (40 blocks + Encoder2 + 6 Decoder + Decoder_1m)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mbp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "H1esc-like 32 Mb multiscale forward, batch 1 (BASELINE configs[1])",
                       "seq_len": SEQ_LEN, "strands": 2, "models": 1},
            "contact_maps_per_s": MAPS_PER_STEP / full,
            "cpu_baseline": {"value": value, "unit": "Mbp/s", "cores": threads, "kind": "port", "sample": sample,
                             "sample_wall_s": wall / max(args.steps, 1)},
            "e2e": {"value": value, "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# our CUDA path
# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernels", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--seq-len", type=int, default=SEQ_LEN)
    ap.add_argument("--workload", default="32mb", choices=["32mb", "256mb"],
                    help="32mb = BASELINE configs[1] (default, what the driver runs); 256mb = configs[3], the "
                         "H1esc_256M-like genomepredict_256Mb forward (256 Mb, 4 levels), sequence-sharded the same way")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--enc-fp16-stages", type=int, default=-1,
                    help="leading encoder stages in single-pass fp16 (0 = three-product bf16 everywhere; default: library default, 3)")
    ap.add_argument("--cascade-mode", default=None, choices=["serial", "batch"])
    ap.add_argument("--chunk-bp", type=int, default=0, help="encoder chunk length in bp (0 = library default)")
    ap.add_argument("--concurrent-strands", action="store_true", help="encode the two strands on two CUDA streams")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    import torch
    import torch.distributed as dist
    from orca_b200 import _lib, models, parallel, predict, synthetic

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a CUDA device; there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.set_impl(args.kernels)
    enc16 = _lib.set_encoder_fp16_stages(args.enc_fp16_stages)  # returns the previous (= default) setting, 3
    if args.enc_fp16_stages >= 0:
        enc16 = min(args.enc_fp16_stages, 7)
    peaks = load_peaks()
    big = args.workload == "256mb"
    L = 256_000_000 if big else args.seq_len
    maps_per_step = 8 if big else MAPS_PER_STEP
    flop_per_strand = 120.6e12 if big else FLOP_PER_STRAND * (L / SEQ_LEN)
    workload = ("H1esc_256M-like 256 Mb multiscale forward (genomepredict_256Mb), batch 1 (BASELINE configs[3])" if big
                else "H1esc-like 32 Mb multiscale forward, batch 1 (BASELINE configs[1])")

    shell = (models.H1esc_256M if big else models.H1esc)(seed=0, device=dev)
    shell.net0.chunk_bp = args.chunk_bp
    seq_host = torch.from_numpy(synthetic.random_sequence(1, L, 0)).pin_memory()
    runner = parallel.ShardedForward(shell, L, rank, world, dev)
    runner.concurrent_strands = args.concurrent_strands
    if args.cascade_mode:
        runner.cascade_mode = args.cascade_mode
    runner.upload(seq_host)  # device-resident input for the `value` leg
    if big:
        runner.set_background(synthetic.normmat_256mb(chrlen_bins=7500), 7500 * 32000)
        runner.d2h_bytes = 4 * 250 * 250 * 4 if rank == 0 else 0
        mpos, wpos = 100_000_000, 128_000_000
    else:
        mpos = wpos = L // 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def step_device():
        return runner.forward(mpos, wpos)

    def step_e2e():
        runner.upload(seq_host)
        maps = runner.forward(mpos, wpos)
        return maps.cpu() if maps is not None else None

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    n0 = _lib.launch_count()
    ms = timed(step_device, args.steps)
    launches = _lib.launch_count() - n0
    ms_step = ms / args.steps
    value = 2 * L / (ms_step * 1e-3) / 1e6

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    e2e_value = 2 * L / (ms_e2e * 1e-3) / 1e6
    h2d_fp32 = int(runner.h2d_bytes)

    # the same end-to-end step fed with packed bases (1 B/bp, orca_b200.feeder) instead of the reference's fp32 one-hot
    codes_host = torch.from_numpy(synthetic.random_codes(1, L, 0)).pin_memory()

    def step_e2e_packed():
        runner.upload(codes_host)
        maps = runner.forward(mpos, wpos)
        return maps.cpu() if maps is not None else None

    step_e2e_packed()
    ms_e2e_packed = timed(step_e2e_packed, args.steps) / args.steps
    h2d_packed = int(runner.h2d_bytes)
    runner.upload(seq_host)  # back to the fp32 window for the roofline leg

    # roofline leg: same steps with per-launch CUDA events around every conv kernel
    _lib.profile_enable(True)
    ms_prof = timed(step_device, args.steps) / args.steps
    prof = _lib.profile_summary()
    _lib.profile_enable(False)
    clocks = sampler.summary()  # sampled across the three timed legs (value, e2e, per-kernel events)
    if rank == 0 and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        with open(os.path.join(ROOT, "gpurun_out", "conv_profile_n%d.json" % world), "w") as f:
            json.dump({"steps": args.steps, "ms_per_step_profiled": ms_prof, "kernels": prof}, f, indent=1)
    # group the per-shape records by kernel: (c_in, c_out, Conv1d|Conv2d, tcgen05|simt); Conv1d records carry dil = 0
    groups = {}
    for r in prof:
        key = (r["c_in"], r["c_out"], "conv1d_k9" if r["dil"] == 0 else ("decoder_program" if r["c_in"] < 0 else "conv2d_3x3"),
               {0: "simt fp32", 1: "tcgen05 bf16x3", 2: "tcgen05 fp16x1"}[r["tc"]])
        grp = groups.setdefault(key, {"ms": 0.0, "launches": 0, "flop": 0.0})
        grp["ms"] += r["ms"]; grp["launches"] += r["launches"]; grp["flop"] += r["flop"]
    roofline = None
    if groups:
        key, dom = max(groups.items(), key=lambda kv: kv[1]["ms"])
        dur_ms = dom["ms"] / dom["launches"]
        achieved = dom["flop"] / dom["launches"] / (dur_ms * 1e-3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        split = 3 if key[3] == "tcgen05 bf16x3" else 1  # bf16 hi/lo split: 3 tensor-core products per algorithmic product
        # DRAM bytes per launch from the committed ncu --set full captures (profiles/r01_prof_*.md, dram__bytes_read.sum +
        # dram__bytes_write.sum), scaled to this run's work per launch:
        #   decoder_program (118-conv Decoder at batch 2, S = 250): 0.69 GB read + 2.29 GB written per launch -- every layer's
        #     output map is written back (the rotating activation buffers of two images exceed what L2 keeps dirty)
        #   conv1d 64->64 fp16x1: 0.512 GB read + 0.471 GB written per 4.224 M positions (= algorithmic 2 x 64 ch x 2 B)
        #   conv1d 64->64 bf16x3: 1.024 GB + 0.978 GB per 4.224 M positions (r01 capture of the three-product kernel)
        traffic = None
        per_launch_flop = dom["flop"] / dom["launches"]
        if key[2] == "decoder_program":
            traffic = 2.98e9 * per_launch_flop / (2 * 290.5e9)   # ncu launch: 2 maps x 290.5 GFLOP
        elif key[:3] == (64, 64, "conv1d_k9") and key[3] == "tcgen05 fp16x1":
            traffic = 232.8 * per_launch_flop / (2 * 9 * 64 * 64)
        elif key[:3] == (64, 64, "conv1d_k9") and key[3] == "tcgen05 bf16x3":
            traffic = 474.0 * per_launch_flop / (2 * 9 * 64 * 64)
        label = lambda k: ("%s, %d convs (%s)" % (k[2], k[1], k[3])) if k[2] == "decoder_program" else "%s %d->%d (%s)" % (k[2], k[0], k[1], k[3])
        breakdown = []
        for k, gk in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
            ach = gk["flop"] / (gk["ms"] * 1e-3) / 1e12
            sp = 3 if k[3] == "tcgen05 bf16x3" else 1
            breakdown.append({"kernel": label(k), "ms_per_step": gk["ms"] / args.steps, "launches_per_step": gk["launches"] // args.steps,
                              "achieved_tflops": ach, "frac": ach / peak, "issued_mma_frac": ach * sp / peak})
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": "%s bf16 sustained (kernel timed inside a long step)" % peaks["source"],
                    "kernel": label(key),
                    "issued_mma_tflops": achieved * split, "issued_mma_frac": achieved * split / peak,
                    "note": "achieved = ALGORITHMIC conv FLOP (2*positions*c_in*c_out*9) per launch / launch time of the kernel with "
                            "the largest share of the step; it issues %dx that many tensor-core FLOP (%s). `kernels` lists every "
                            "conv kernel family the same way." % (split, "fp32-parity bf16 hi/lo split" if split == 3 else "single fp16 product"),
                    "avg_launch_ms": dur_ms, "launches_per_step": dom["launches"] // args.steps,
                    "share_of_step": dom["ms"] / args.steps / ms_prof, "ms_per_step_profiled": ms_prof,
                    "conv_ms_per_step": sum(r["ms"] for r in prof) / args.steps, "kernels": breakdown}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Mbp/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": ("f32" if args.kernels == "simt" else
                          "bf16x3 (fp32 operands split hi+lo, 3 tcgen05 products, fp32 accumulate)" if enc16 == 0 else
                          "fp16 (1 tcgen05 product, fp32 accumulate) in encoder stages 1-%d; bf16x3 (operands split hi+lo, 3 products) "
                          "in the other stages, the U-nets and the decoders" % enc16),
                "data": "synthetic",
                "config": {"workload": workload,
                           "seq_len": L, "strands": 2, "models": 1, "kernels": args.kernels, "encoder_fp16_stages": enc16,
                           "l2": "inputs (512 MB) and stage activations (>1 GB per chunk) exceed the 126 MB L2",
                           "parallelism": "sequence-sharded encoder x%d + all-gather, strand-parallel cascades" % world
                           if world > 1 else "single GPU"},
                "contact_maps_per_s": maps_per_step / (ms_step * 1e-3),
                "algorithmic_tflops": 2 * flop_per_strand / (ms_step * 1e-3) / 1e12,
                "e2e": {"value": e2e_value, "unit": "Mbp/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": h2d_fp32, "d2h_bytes_per_step": int(runner.d2h_bytes),
                        "packed_bases": {"value": 2 * L / (ms_e2e_packed * 1e-3) / 1e6, "ms_per_step": ms_e2e_packed,
                                         "h2d_bytes_per_step": h2d_packed,
                                         "note": "same step with the sequence uploaded as 1 B/bp packed bases (orca_b200.feeder)"}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            one = cpu_sample(threads)
            one()
            t0 = time.perf_counter()
            tb, te, td, tm = one()
            # 256 Mb: 320 blocks, pooling half of Encoder2 at 64000 bins (8 x 1/3 of the timed 8000-bin U-net), 4 decoders
            full = 2 * (320 * tb + 8 * te / 3 + 4 * td) if big else cpu_extrapolate(tb, te, td, tm)
            line["cpu_baseline"] = {
                "value": 2 * L / full / 1e6, "unit": "Mbp/s", "cores": threads, "kind": "port",
                "sample": CPU_SAMPLE + "; extrapolated linearly to the full 2-strand step",
                "sample_wall_s": time.perf_counter() - t0,
                "maps_per_s": maps_per_step / full}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
