#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_colocated.py -x -q -m gpu 2>&1 | tail -5
for rep in 1 2; do for val in 0 1; do
  RLCTR_GROUP_UNSORTED_CATCHUP=$val python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-eager-gpu --no-cpu-baseline > gpurun_out/ab_$val.log 2>gpurun_out/ab_$val.err || tail -5 gpurun_out/ab_$val.err
  python - <<EOF2
import json
d=None
for l in open("gpurun_out/ab_$val.log"):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l)
if d: print("UNSORTED=$val", "value %.2f M  ms %.4f  steady %.4f  e2e %.2f M" % (d["value"]/1e6, d["ms_per_step"], d["steady_state"]["ms_per_step"], d["e2e"]["value"]/1e6))
EOF2
done; done
