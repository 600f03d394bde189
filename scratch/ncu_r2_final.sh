#!/bin/bash
# round 2 (final state) evidence: launch list of the C2 bench step + ncu --set full of its kernels + the FFM kernels
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --profile-steps 0 --no-configs --no-eager-gpu --windows 0"
$CMD > gpurun_out/r2f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2f.csv $CMD > /dev/null 2>&1
$CMD > gpurun_out/r2f_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"group_fwd_kernel|rows_short_kernel|replay_rows_kernel|gemm3x_tma|gemv_bwd" -s 12 -c 11 -o gpurun_out/prof_r2f -f $CMD > gpurun_out/ncu_r2f.log 2>&1
python scratch/ffm_probe.py 8 > gpurun_out/r2f_ffm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"ffm_fwd_kernel|rows_wide_kernel|replay_kernel" -s 12 -c 3 -o gpurun_out/prof_r2f_ffm -f python scratch/ffm_probe.py 8 > gpurun_out/ncu_r2f_ffm.log 2>&1
tail -2 gpurun_out/ncu_r2f.log gpurun_out/ncu_r2f_ffm.log; ls -la gpurun_out/prof_r2f*.ncu-rep gpurun_out/launches_r2f.csv
