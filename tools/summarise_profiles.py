"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep} into the tracked summaries under profiles/ (run in the build
container: `python tools/summarise_profiles.py r01`)."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

HBM_PEAK_GBS = 6448.1  # MEASURED_PEAKS.json (driver-measured STREAM-style copy on this pool's B200s)
KEY = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
       "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum"]


def launches():
    path = os.path.join(OUT, "launches.csv")
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", tag + "_launches.md"), "w") as f:
        f.write("# %s launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n" % tag)
        f.write("Workload: `tools/ncu_target.py` at the BENCH shapes = one 32 Mb Encoder pass per strand (single chunk, all 7 stages; fp32\n"
                "one-hot forward strand, packed-base reverse strand) + Encoder2 + the strand-batched 6-level decoder cascade (7 decoder\n"
                "calls at batch 2) of an H1esc-like shell + the 256 Mb background kernels. Cold-cache, serialised launches: compare SHARES.\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (n[:110], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
        f.write("\ntotal %.1f us over %d launches\n" % (tot, sum(a[0] for a in agg.values())))


def ncu_raw(name):
    path = os.path.join(OUT, name + ".csv")
    rows = [r for r in csv.reader(open(path)) if r]
    hi = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    hdr, units = rows[hi], rows[hi + 1]
    with open(os.path.join(ROOT, "profiles", "%s_%s.md" % (tag, name)), "w") as f:
        f.write("# %s %s (ncu --set full --clock-control none), one row per captured launch\n\n" % (tag, name))
        cols = [k for k in KEY if k in hdr]
        f.write("| kernel | grid | " + " | ".join(c.split(".")[0].replace("__", " ") for c in cols) + " | DRAM GB/s (frac of measured %.0f) |\n" % HBM_PEAK_GBS)
        f.write("|---|---|" + "---:|" * (len(cols) + 1) + "\n")

        def val(r, c):
            v, u = float(r[hdr.index(c)].replace(",", "")), units[hdr.index(c)]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
            return v * scale
        out = []
        for r in rows[hi + 2:]:
            kn = r[hdr.index("Kernel Name")]
            dram = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
            dur = val(r, "gpu__time_duration.sum")
            gbs = dram / dur / 1e9
            f.write("| `%s` | %s | " % (kn[-80:], r[hdr.index("Grid Size")]) + " | ".join("%s %s" % (r[hdr.index(c)], units[hdr.index(c)]) for c in cols)
                    + " | %.0f (%.2f) |\n" % (gbs, gbs / HBM_PEAK_GBS))
            out.append((kn, dram, dur))
        return out


launches()
captured = {}
for fn in sorted(os.listdir(OUT)):
    if fn.startswith("prof_") and fn.endswith(".csv"):
        captured[fn[len("prof_"):-len(".csv")]] = ncu_raw(fn[:-len(".csv")])

# DRAM traffic table for bench.py's roofline.traffic (bytes per captured launch + the algorithmic conv FLOP of that launch)
import json
traffic = {}
src = "profiles/%s_prof_%%s.md (ncu --set full, tools/ncu_target.py at the bench shapes)" % tag
if captured.get("decoder_stream"):
    kn, dram, dur = captured["decoder_stream"][0]  # first decoder call of the cascade: 117 convs, batch 2, no coarse input
    px = 2 * 250 * 250
    flop = px * (5 * 2 * 9 * 64 * 64 + 112 * 2 * 9 * 64 * 32)
    traffic["decoder_stream"] = {"dram_bytes": dram, "flop": flop, "ms_under_ncu": dur * 1e3, "source": src % "decoder_stream"}
for kn, dram, dur in captured.get("conv1d", []):
    import re
    m = re.search(r"conv1d_tc_kernel<(\d+), (\d+), (?:\(bool\))?(\d+|true|false), (\d+)>", kn)
    if not m:
        continue
    ci, co, fmt = int(m.group(1)), int(m.group(2)), int(m.group(4))
    key = "conv1d_k9_%d_%d_%s" % (ci, co, "fp16" if fmt else "bf16x3")
    if key not in traffic:  # first launch of the family: the full-resolution layer of its stage
        traffic[key] = {"dram_bytes": dram, "flop": None, "ms_under_ncu": dur * 1e3, "source": src % "conv1d", "kernel": kn[-60:]}
# positions of the first launch of each family at L = 32 Mb (+ the 8 pad rows): stage 1 = L, stage 2 = L/4, stage 3 = L/16
pos = {"conv1d_k9_64_64_fp16": 32_000_000, "conv1d_k9_64_96_fp16": 8_000_000, "conv1d_k9_96_96_fp16": 8_000_000,
       "conv1d_k9_96_128_fp16": 2_000_000, "conv1d_k9_128_128_fp16": 2_000_000}
for k, n in pos.items():
    if k in traffic:
        ci, co = int(k.split("_")[2]), int(k.split("_")[3])
        traffic[k]["flop"] = 2.0 * 9 * ci * co * n
with open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
if os.path.exists(os.path.join(OUT, "sass_opcodes.txt")):
    import shutil
    shutil.copy(os.path.join(OUT, "sass_opcodes.txt"), os.path.join(ROOT, "profiles", tag + "_sass_opcodes.txt"))
print(os.listdir(os.path.join(ROOT, "profiles")))
