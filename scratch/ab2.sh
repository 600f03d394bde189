#!/bin/bash
for mode in 0 1; do for fork in "" "--no-fork"; do
RLCTR_LOOKUP=$mode python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --profile-steps 0 $fork > gpurun_out/ab2.log 2>gpurun_out/ab2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ab2.log").read().strip().splitlines()[-1])
print("lookup=$mode fork='$fork'", round(d["value"]/1e6,2), "M samples/s", round(d["ms_per_step"],4), "ms/step", d["gpu_launches"])
PY
done; done
