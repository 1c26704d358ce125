#!/bin/bash
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ab"
$CMD > gpurun_out/prof_cls_plain.log 2>&1 || { echo plain failed; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:^hx_classify_kernel$" -s 1 -c 1 -o gpurun_out/prof_cls -f $CMD > gpurun_out/ncu_cls.log 2>&1
ncu -i gpurun_out/prof_cls.ncu-rep --page raw --csv > gpurun_out/prof_cls.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_cls.ncu-rep --page source --csv > gpurun_out/prof_cls.src.csv 2>/dev/null
rm -f gpurun_out/prof_cls.ncu-rep
