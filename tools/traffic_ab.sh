#!/bin/bash
# DRAM traffic + duration of the graded scan kernel at the bench size for the library named by $1 (tag $2)
#   bash tools/traffic_ab.sh hyperex_b200/_lib/libhyperex_b200_expldpol3.so ldpol3
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ab --no-file-e2e"
export HYPEREX_B200_LIB=$1
$CMD > gpurun_out/traffic_plain_$2.json 2> gpurun_out/traffic_plain_$2.err || { echo plain run failed; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct \
    --clock-control none --kernel-name-base demangled -k "regex:hx_myers_pair_kernel<unsigned int, .int.15, .bool.0, .int.-1, .bool.0, .bool.0, .bool.0, .int.2>" -s 2 -c 1 --csv \
    --log-file gpurun_out/traffic_$2.csv $CMD > gpurun_out/traffic_ncu_$2.log 2>&1
grep -v "^==" gpurun_out/traffic_$2.csv | cut -d, -f5,13- | tail -7
python -c "
import json; d=json.load(open('gpurun_out/traffic_plain_$2.json')); print('$2', 'kernel ms', d['roofline']['kernel_ms_per_launch'], 'step ms', d['ms_per_step'])"
