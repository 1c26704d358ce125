#!/bin/bash
# ncu capture of one instantiation of the scan kernel from a short bench.py run:
#   tools/profile_fast.sh <k> <np> [lf pk nib] [tag] [ppw]
# k: -1 Myers, 0..2 Shift-Or / Wu-Manber; np: automata (pattern halves when pk = 1) per thread; lf / pk / nib: the suffix
# filter, pattern packing and packed-text template switches (0 / 1); ppw: automata per word of the packed kernels (2, or 4 at -m 0).  The built-in regions are one group of 15 automata
# (16 halves when packed).  Leaves gpurun_out/prof_<tag>.{raw,src}.csv.
K=$1
NP=${2:-15}
LF=${3:-0}
PK=${4:-0}
NIB=${5:-0}
TAG=${6:-k${K}_np${NP}_lf${LF}_pk${PK}_nib${NIB}}
PPW=${7:-2}
BK=$K; [ "$K" -lt 0 ] && BK=0
CMD="python bench.py --k $BK --steps 1 --warmup 1 --records 200000 --no-cpu-baseline --no-file-e2e"
$CMD > gpurun_out/prof_plain_$TAG.log 2>&1 || { echo plain failed; tail -3 gpurun_out/prof_plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:hx_myers_pair_kernel<unsigned int, .int.$NP, .bool.0, .int.$K, .bool.$LF, .bool.$PK, .bool.$NIB, .int.$PPW>" -s 1 -c 1 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_$TAG.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_$TAG.src.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
grep -c . gpurun_out/prof_$TAG.raw.csv
