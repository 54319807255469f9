#!/bin/bash
# Run under gpurun (one GPU): launch list of one 4 Mb encoder pass + one decoder cascade, and ncu --set full
# captures of the three tcgen05 kernels.  Outputs land in gpurun_out/ (post-process with tools/summarise_profiles.py).
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_kernel -c 8 -f -o gpurun_out/prof_conv1d \
    python tools/ncu_target.py > gpurun_out/ncu_conv1d.log 2>&1
NCU_REPS=0 ncu --set full --clock-control none --import-source on -k regex:conv2d_tc_kernel -s 30 -c 6 -f \
    -o gpurun_out/prof_conv2d python tools/ncu_target.py > gpurun_out/ncu_conv2d.log 2>&1
ncu --set full --clock-control none -k regex:lconv1_tc -c 1 -f -o gpurun_out/prof_first \
    python tools/ncu_target.py > gpurun_out/ncu_first.log 2>&1
for f in gpurun_out/ncu_*.log; do tail -n 1 $f; done
