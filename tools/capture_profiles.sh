#!/bin/bash
# Run under gpurun (one GPU): launch list of tools/ncu_target.py and ncu --set full captures of every kernel family.
# The .ncu-rep files are converted to raw-page CSV on the box and deleted (gpurun_out/ is capped at 64 MiB);
# post-process here with `python tools/summarise_profiles.py r01`.
set -u
mkdir -p gpurun_out
T=tools/ncu_target.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python $T > gpurun_out/ncu_launches.log 2>&1
cap() {  # name, kernel regex, launch count
  ncu --set full --clock-control none -k "regex:$2" -c $3 -f -o /tmp/prof_$1 python $T > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/prof_$1.csv 2>/dev/null
  rm -f /tmp/prof_$1.ncu-rep
}
cap conv1d conv1d_tc_kernel 11
cap decoder_program conv2d_program_kernel 1
cap first lconv1_tc 2
cap glue "extra_conv_planes|outer_sum_planes|final_head_planes|pool_planes" 6
for f in gpurun_out/ncu_*.log; do tail -n 1 $f; done
ls -la gpurun_out
