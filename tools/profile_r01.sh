#!/bin/bash
# ncu evidence for bench.py's dominant kernel (run on the GPU box under gpurun, after a plain run exited 0)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ab"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hx_myers_pair -s 1 -c 1 -o gpurun_out/prof_myers -f $CMD > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:hx_classify -s 1 -c 1 -o gpurun_out/prof_classify -f $CMD > gpurun_out/ncu_classify.log 2>&1
ls -la gpurun_out
