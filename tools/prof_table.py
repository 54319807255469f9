"""Print gpurun_out/conv_profile_n1.json (written by bench.py) as a table; 1D convs have dil=0."""
import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_profile_n1.json"))
print("ms/step (profiled) %.2f" % d["ms_per_step_profiled"])
tot = 0
rows = {}
for k in d["kernels"]:
    key = (k["c_in"], k["c_out"], "1d" if k["dil"] == 0 else "2d", k["tc"])
    r = rows.setdefault(key, {"ms": 0, "n": 0, "flop": 0})
    r["ms"] += k["ms"] / d["steps"]; r["n"] += k["launches"] / d["steps"]; r["flop"] += k["flop"] / d["steps"]
for key, r in sorted(rows.items(), key=lambda kv: -kv[1]["ms"]):
    tot += r["ms"]
    print("%3d->%3d %s tc=%d launches/step=%5d ms/step=%8.3f us/launch=%8.1f TFLOP/s=%7.1f" % (key + (r["n"], r["ms"], r["ms"] / r["n"] * 1e3, r["flop"] / r["ms"] / 1e9)))
print("conv total ms/step %.2f" % tot)
