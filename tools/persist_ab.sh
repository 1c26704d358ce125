#!/bin/bash
# A/B of the scan grid: persistent CTAs fetching batches from a device counter (default), persistent with static rounds
# (HYPEREX_SCAN_GRID=static), one CTA per batch (HYPEREX_SCAN_GRID=full).  Step time from bench.py, DRAM traffic and the
# stall picture of the graded kernel from ncu.  Run on the GPU box under gpurun; leaves gpurun_out/persist_*.
set -u
mkdir -p gpurun_out
MODES=${1:-"dynamic static full"}
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-file-e2e"
PCMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ab --no-file-e2e"
for mode in $MODES; do
  if [ $mode = dynamic ]; then unset HYPEREX_SCAN_GRID; else export HYPEREX_SCAN_GRID=$mode; fi
  $CMD > gpurun_out/persist_bench_$mode.json 2> gpurun_out/persist_bench_$mode.err || { echo "$mode run failed"; tail -5 gpurun_out/persist_bench_$mode.err; exit 1; }
  [ $mode = static ] && continue
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:hx_myers_pair_kernel<unsigned int, .int.15, .bool.0, .int.-1, .bool.0, .bool.0, .bool.0, .int.2>" -s 2 -c 1 \
      -o gpurun_out/persist_$mode -f $PCMD > gpurun_out/persist_ncu_$mode.log 2>&1
  ncu -i gpurun_out/persist_$mode.ncu-rep --page raw --csv > gpurun_out/persist_$mode.raw.csv 2>/dev/null
  rm -f gpurun_out/persist_$mode.ncu-rep
  python tools/ncu_brief.py gpurun_out/persist_$mode.raw.csv
done
