#!/bin/bash
RLCTR_LOOKUP=0 bash scratch/bench_short.sh r2e_old
bash scratch/bench_short.sh r2e_new
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --profile-steps 0"
$CMD > gpurun_out/ncu_rows_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:"lookup_rows_kernel|rows_staged|rows_short_kernel|embed_fwd_kernel|replay_rows" -s 6 -c 12 --csv --log-file gpurun_out/ncu_nocache_new.csv $CMD > /dev/null 2>&1
RLCTR_LOOKUP=0 $CMD > gpurun_out/ncu_rows_plain2.log 2>&1 && \
RLCTR_LOOKUP=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:"lookup_rows_kernel|rows_staged|rows_short_kernel|embed_fwd_kernel|replay_rows" -s 6 -c 12 --csv --log-file gpurun_out/ncu_nocache_old.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
for f in ("gpurun_out/ncu_nocache_new.csv","gpurun_out/ncu_nocache_old.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); ii=hdr.index("ID")
    d={}
    for r in rows[1:]: d.setdefault((int(r[ii]),r[ik][:40]),{})[r[im]]=float(r[iv].replace(",",""))
    print(f)
    for k,v in sorted(d.items()): print("  ",k[1].ljust(40), round(v["gpu__time_duration.sum"]/1e3,1),"us rd",round(v["dram__bytes_read.sum"]/1e6,1),"wr",round(v["dram__bytes_write.sum"]/1e6,1))
PY
