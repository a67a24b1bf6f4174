#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
./tools/bin/microbench B > gpurun_out/micro_B.txt 2>&1; cat gpurun_out/micro_B.txt
for g in 0 32 64; do
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --csv --log-file gpurun_out/micro_A_$g.csv ./tools/bin/microbench A $g > gpurun_out/micro_A_$g.txt 2>&1
cat gpurun_out/micro_A_$g.txt
done
