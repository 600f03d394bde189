#!/bin/bash
# A/B of one environment switch on the same box: scratch/ab_bench.sh VAR A B
V=$1; A=$2; B=$3
for rep in 1 2; do
for val in $A $B; do
  env $V=$val python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/ab_$val.log 2>/dev/null
  python - <<EOF2
import json
for l in open("gpurun_out/ab_$val.log"):
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l)
g=d["gemm"]
print("$V=$val", "value %.2f M  ms %.4f  steady %.4f  e2e %.2f M" % (d["value"]/1e6, d["ms_per_step"], d["steady_state"]["ms_per_step"], d["e2e"]["value"]/1e6),
      "bwd300x200 %.4f fwd150x300 %.4f" % (g["rlctr_linear_bwd[300x200]"]["mean_ms"], g["rlctr_linear_fwd[150x300]"]["mean_ms"]))
EOF2
done
done
