#!/bin/bash
# ncu capture of the small-k fast-path kernel: tools/profile_fast.sh <k> [np]   (np: automata per thread of the group to
# capture; the built-in regions are one group of 15 at -m 0..1 and groups of 8 + 7 at -m 2)
K=$1
NP=${2:-15}
CMD="python bench.py --k $K --steps 1 --warmup 1 --records 200000 --no-cpu-baseline --no-file-e2e"
$CMD > gpurun_out/prof_fast_plain_$K.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:hx_myers_pair_kernel<unsigned int, .int.$NP, .bool.0, .int.$K, .bool.0>" -s 1 -c 1 -o gpurun_out/prof_fast_k$K -f $CMD > gpurun_out/ncu_fast_$K.log 2>&1
ncu -i gpurun_out/prof_fast_k$K.ncu-rep --page raw --csv > gpurun_out/prof_fast_k$K.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_fast_k$K.ncu-rep --page source --csv > gpurun_out/prof_fast_k$K.src.csv 2>/dev/null
rm -f gpurun_out/prof_fast_k$K.ncu-rep
grep -c . gpurun_out/prof_fast_k$K.raw.csv
