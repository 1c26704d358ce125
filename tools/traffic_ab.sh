#!/bin/bash
# DRAM traffic of the scan kernel at the bench size for the library named by $1 (tag $2)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ab"
export HYPEREX_B200_LIB=$1
$CMD > gpurun_out/traffic_plain_$2.log 2>&1 || { echo plain run failed; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:hx_myers_pair -s 1 -c 1 --csv --log-file gpurun_out/traffic_$2.csv $CMD > gpurun_out/traffic_ncu_$2.log 2>&1
grep -v "^==" gpurun_out/traffic_$2.csv | tail -4
