"""Kernel timeline of one CUDA-graph replay of the co-located C2 step (torch.profiler / CUPTI): busy time, gaps, per-kernel list."""
import sys, json, torch
sys.path.insert(0, ".")
from rl_ctr_prediction_b200 import optim, p_model, colocated, graphs
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
N, B, F, D = 10_000_000, 65536, 15, 10
gen = torch.Generator(device=dev).manual_seed(1)
def batch():
    per = N // F
    x = torch.randint(0, per, (B, F), generator=gen, device=dev, dtype=torch.int64) + torch.arange(F, device=dev) * per
    y = (torch.rand(B, generator=gen, device=dev) < 0.05).long()
    return x, y
ms = [p_model.LR(N, device=dev), p_model.FM(N, D, device=dev), p_model.DeepFM(N, F, D, device=dev)]
for m in ms:
    with torch.no_grad():
        m.table.mul_(0.1)
    m.train()
g = colocated.colocate(ms)
opt = optim.Adam(g.parameters(), lr=1e-3, weight_decay=1e-5)
step = graphs.GraphedTrainStep([(g, opt)])
bs = [batch() for _ in range(12)]
for x, y in bs[:8]:
    step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(12):
        step(*bs[i % 12])
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
sp = [e.time_range.start for e in ev if "sort_prep" in e.name]
print("replay periods (us):", [round(b - a, 1) for a, b in zip(sp, sp[1:])])
# what sits between the end of one replay and the start of the next
ends = [i for i, e in enumerate(ev) if "steps_advance" in e.name]
for i in ends[3:5]:
    for e in ev[i:i + 6]:
        print(f"   between: {e.time_range.start - ev[i].time_range.start:8.1f} +{e.time_range.end - e.time_range.start:6.1f} {e.name[:60]}")
# last replay: from the last sort_prep kernel on
starts = [i for i, e in enumerate(ev) if "sort_prep" in e.name]
seg = ev[starts[-1]:]
t0, t1 = seg[0].time_range.start, max(e.time_range.end for e in seg)
busy = sum(e.time_range.end - e.time_range.start for e in seg)
print(f"replay span {t1 - t0:.1f} us, kernels {len(seg)}, busy {busy:.1f} us, gaps {t1 - t0 - busy:.1f} us")
prev_end = t0
for e in seg:
    if e.time_range.end - e.time_range.start > 30 or "splitk" in e.name or "bce" in e.name:
        print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:7.1f}  gap {e.time_range.start - prev_end:6.1f}  {e.name[:70]}")
    prev_end = max(prev_end, e.time_range.end)
