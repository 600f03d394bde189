#!/bin/bash
# ncu --set full of the row kernels of one bench step (lookup / streamed forward / update), after a plain run of the same command
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --profile-steps 0"
$CMD > gpurun_out/ncu_rows_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"lookup_rows_kernel|rows_staged_kernel|rows_short_kernel" -s 6 -c 4 -o gpurun_out/prof_r2_rows2 -f $CMD > gpurun_out/ncu_rows.log 2>&1
tail -3 gpurun_out/ncu_rows.log
