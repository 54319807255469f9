"""Turn gpurun_out/{launches.csv, prof_*.ncu-rep} into the tracked summaries under profiles/ (run in the build
container: `python tools/summarise_profiles.py r01`)."""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

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
        f.write("Workload: `tools/ncu_target.py` = one 4 Mb Encoder pass per strand (single chunk, all 7 stages; fp32 one-hot forward\n"
                "strand, packed-base reverse strand) + Encoder2 + the strand-batched 6-level decoder cascade (7 decoder calls at\n"
                "batch 2) of an H1esc-like shell. Cold-cache, serialised launches: compare SHARES.\n\n")
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
        f.write("| kernel | " + " | ".join(c.split(".")[0].replace("__", " ") for c in cols) + " |\n")
        f.write("|---|" + "---:|" * len(cols) + "\n")
        for r in rows[hi + 2:]:
            kn = r[hdr.index("Kernel Name")]
            f.write("| `%s` | " % kn[-70:] + " | ".join("%s %s" % (r[hdr.index(c)], units[hdr.index(c)]) for c in cols) + " |\n")


launches()
for fn in sorted(os.listdir(OUT)):
    if fn.startswith("prof_") and fn.endswith(".csv"):
        ncu_raw(fn[:-len(".csv")])
print(os.listdir(os.path.join(ROOT, "profiles")))
