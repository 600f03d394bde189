#!/bin/bash
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --profile-steps 0"
for mode in 0 1; do
RLCTR_LOOKUP=$mode $CMD > gpurun_out/ab3_plain$mode.log 2>&1 && \
RLCTR_LOOKUP=$mode ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_lookup$mode.csv $CMD > /dev/null 2>&1
done
ls -la gpurun_out/launches_r2_lookup*.csv
