#!/bin/bash
# final state of round 2: launch list + ncu --set full of the co-located update (two lanes per record), catch-up and gather
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --profile-steps 0 --no-configs --no-eager-gpu --windows 0"
$CMD > gpurun_out/r2g_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2g.csv $CMD > /dev/null 2>&1
$CMD > gpurun_out/r2g_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"group_fwd_kernel|group_rows2_kernel|replay_rows_kernel" -s 2 -c 3 -o gpurun_out/prof_r2g -f $CMD > gpurun_out/ncu_r2g.log 2>&1
ls -la gpurun_out/prof_r2g.ncu-rep gpurun_out/launches_r2g.csv
