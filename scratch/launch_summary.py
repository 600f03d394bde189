"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: the last STEP of the bench (between the
last two rlctr sort launches... approximated as the last `n` launches) by kernel.  usage: launch_summary.py csv [tail_launches]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iK, iV, iID = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
body = [r for r in rows if r is not hdr and r[iID].isdigit()]
names = [r[iK] for r in body]
# one step = from a sort_prep_kernel launch to the next one
starts = [i for i, n in enumerate(names) if "sort_prep_kernel" in n]
lo, hi = (starts[-1], len(body)) if len(sys.argv) < 3 else (starts[-2], starts[-1])
agg = collections.OrderedDict()
for r in body[lo:hi]:
    n = re.sub(r"\(.*", "", r[iK])[:110]
    t = float(r[iV].replace(",", "")) / 1e3
    a = agg.setdefault(n, [0.0, 0])
    a[0] += t; a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"launches {hi - lo}, total {tot:.1f} us")
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {t:8.1f} | {c:3d} | {100 * t / tot:5.1f}% | `{n}` |")
