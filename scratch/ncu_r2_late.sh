#!/bin/bash
# late round 2: launch list of the final build + ncu --set full of the tower GEMMs (incl. the masked-dgrad instantiation)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --profile-steps 0 --no-configs --no-eager-gpu --windows 0"
$CMD > gpurun_out/r2r_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2r.csv $CMD > /dev/null 2>&1
$CMD > gpurun_out/r2r_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"gemm3x_tma_kernel" -s 12 -c 6 -o gpurun_out/prof_r2r -f $CMD > gpurun_out/ncu_r2r.log 2>&1
ls -la gpurun_out/prof_r2r.ncu-rep gpurun_out/launches_r2r.csv
