#!/bin/bash
# Round-2 ncu evidence (run on the GPU box under gpurun; every ncu pass only after the same command exited 0 without ncu).
# Leaves in gpurun_out/: r02_launches.csv (per-launch durations of `bench.py --steps 2 --warmup 1`, the kernel's SHARE of
# the step must agree with bench.py's live figure) and prof_<tag>.{raw,src}.csv for the graded Myers kernel and the
# library-default kernels.  tools/ncu_summary.py r02 turns them into profiles/r02_*.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ab --no-file-e2e"
$CMD > gpurun_out/r02_prof_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_prof_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
# full captures at the bench size (1 M reads): the graded kernel (traffic figure) ...
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:hx_myers_pair_kernel<unsigned int, .int.15, .bool.0, .int.-1, .bool.0, .bool.0, .bool.0, .int.2>" -s 2 -c 1 -o gpurun_out/prof_myers_1M -f $CMD > gpurun_out/r02_ncu_myers.log 2>&1
ncu -i gpurun_out/prof_myers_1M.ncu-rep --page raw --csv > gpurun_out/prof_myers_1M.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_myers_1M.ncu-rep --page source --csv > gpurun_out/prof_myers_1M.src.csv 2>/dev/null
rm -f gpurun_out/prof_myers_1M.ncu-rep
# ... and at 200 k reads: the library-default kernels of -m 0 / 1 / 2 (packed automata), the packed-text variant, classify
bash tools/profile_fast.sh 0 16 1 1 0 shiftor_k0_packed
bash tools/profile_fast.sh 1 16 1 1 0 wm_k1_packed
bash tools/profile_fast.sh 2 16 1 1 0 wm_k2_packed
bash tools/profile_fast.sh 0 16 1 1 1 shiftor_k0_packed_nibble
ls -la gpurun_out | grep -E "prof_|r02_" | tail -20
