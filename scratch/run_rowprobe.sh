#!/bin/bash
# first GPU call of round 2: what bounds the random row access?  (results -> gpurun_out/rowprobe_*.log)
set -x
B=scratch/bin/rowprobe
for g in 0 32 64 128; do $B $g 10 > gpurun_out/rowprobe_g$g.log 2>&1; done
$B 0 2.5 > gpurun_out/rowprobe_g0_n2p5.log 2>&1
$B 0 40 > gpurun_out/rowprobe_g0_n40.log 2>&1
$B 0 10 quick > gpurun_out/rowprobe_quick.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --csv --log-file gpurun_out/rowprobe_ncu_g0.csv $B 0 10 quick > /dev/null 2>&1
$B 32 10 quick > gpurun_out/rowprobe_quick32.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --csv --log-file gpurun_out/rowprobe_ncu_g32.csv $B 32 10 quick > /dev/null 2>&1
tail -n +1 gpurun_out/rowprobe_g0.log
