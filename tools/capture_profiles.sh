#!/bin/bash
# Run under gpurun (one GPU): launch list of tools/ncu_target.py (bench shapes) and ncu --set full captures of every kernel
# family.  The .ncu-rep files are converted to raw-page CSV on the box and deleted (gpurun_out/ is capped at 64 MiB), except
# the decoder stream kernel's report (source page: per-line stall reasons);
# post-process here with `python tools/summarise_profiles.py r02`.
set -u
mkdir -p gpurun_out
T=tools/ncu_target.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python $T > gpurun_out/ncu_launches.log 2>&1
cap() {  # name, kernel regex, launch count, env
  env $4 ncu --set full --clock-control none -k "regex:$2" -c $3 -f -o /tmp/prof_$1 python $T > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.csv 2>/dev/null
  rm -f /tmp/prof_$1.ncu-rep
}
cap conv1d conv1d_tc_kernel 14 "NCU_CASCADE=0 NCU_BACKGROUND=0"
cap first "lconv1_tc|lconv1_edge|pool_planes" 5 "NCU_CASCADE=0 NCU_BACKGROUND=0"
cap decoder_stream conv2d_stream_kernel 2 "NCU_ENCODER=0 NCU_BACKGROUND=0"
cap glue "ds_outer_sum|ds_extra_conv|ds_final_head|symmetrise|upsample2_planes|from_channel_last|to_channel_last" 12 "NCU_ENCODER=0 NCU_BACKGROUND=0"
cap background "background_" 2 "NCU_ENCODER=0 NCU_CASCADE=0"
for f in gpurun_out/ncu_*.log; do tail -n 1 $f; done
cuobjdump -sass orca_b200/liborca_b200.so > /tmp/sass.txt 2>/dev/null
python - <<'PY'
import collections, re
cur, hist = None, collections.OrderedDict()
for line in open('/tmp/sass.txt'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        hist[cur][m.group(1).split('.')[0]] += 1
with open('gpurun_out/sass_opcodes.txt', 'w') as f:
    for k, c in hist.items():
        keys = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'HMMA', 'LDGSTS', 'LDG', 'STG', 'LDS', 'STS', 'RED', 'ATOM']
        f.write(k[:120] + '\n    ' + ' '.join('%s=%d' % (o, c[o]) for o in keys if c[o]) + '  total=%d\n' % sum(c.values()))
PY
ls -la gpurun_out | tail -20
