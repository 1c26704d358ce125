#!/bin/bash
# ncu evidence for bench.py's kernels (run on the GPU box under gpurun; ncu only after the plain run exited 0).
# Leaves in gpurun_out/: launches.csv (per-launch durations of the same command), prof_myers.ncu-rep and
# prof_classify.ncu-rep (--set full, one launch each).  tools/ncu_summary.py turns them into profiles/*.csv.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ab --no-file-e2e"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hx_myers_pair -s 2 -c 1 -o gpurun_out/prof_myers -f $CMD > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^hx_classify_kernel$" -s 2 -c 1 -o gpurun_out/prof_classify -f $CMD > gpurun_out/ncu_classify.log 2>&1
ls -la gpurun_out | tail -8
