#!/bin/bash
# round 2 evidence: launch list of one bench step + ncu --set full of the top kernels (after a plain run of the same command)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --profile-steps 0 --no-configs --no-eager-gpu --windows 0"
$CMD > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv $CMD > /dev/null 2>&1
$CMD > gpurun_out/r2_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"rows_short_kernel|embed_fwd_kernel|replay_rows_kernel|gemm3x_tma|gemv_bwd|rows_scalar|embed_fwd_scalar|rows_catchup_scalar" -s 28 -c 20 -o gpurun_out/prof_r2_final -f $CMD > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log; ls -la gpurun_out/prof_r2_final.ncu-rep gpurun_out/launches_r2.csv
