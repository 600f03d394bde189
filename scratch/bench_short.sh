#!/bin/bash
# short bench + per-kernel table:  bench_short.sh <tag> [extra bench args]
tag=$1; shift
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err || tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
print("$tag", round(d["value"]/1e6,2), "M samples/s", round(d["ms_per_step"],4), "ms/step")
for k,v in d["roofline"]["all_kernels"].items(): print("  ", k, round(v["mean_ms"]*1e3,1), "us", round(v["frac_of_hbm_peak"] or 0,3))
print("  ", d["roofline"]["share_of_step_by_entry_point"])
PY
