"""Per-call times of the FFM training step (eager, CUDA events around every library call).  usage: ffm_probe.py [steps]"""
import sys, torch
sys.path.insert(0, ".")
from rl_ctr_prediction_b200 import _lib, optim, p_model, graphs
dev = torch.device("cuda", 0)
N, B, F, D = 10_000_000, 65536, 15, 10
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 14
gen = torch.Generator(device=dev).manual_seed(1)
def batch():
    per = N // F
    x = torch.randint(0, per, (B, F), generator=gen, device=dev, dtype=torch.int64) + torch.arange(F, device=dev) * per
    y = (torch.rand(B, generator=gen, device=dev) < 0.05).long()
    return x, y
m = p_model.FFM(N, F, D, device=dev)
with torch.no_grad():
    m.table.mul_(0.1)
m.train()
opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
lossf = torch.nn.BCELoss()
bs = [batch() for _ in range(steps)]
for x, y in bs[:6]:
    graphs.eager_step(m, opt, lossf, x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for x, y in bs[6:]:
    graphs.eager_step(m, opt, lossf, x, y)
e1.record(); torch.cuda.synchronize()
print(f"eager step {e0.elapsed_time(e1) / (steps - 6) * 1e3:.1f} us")
prof = _lib.KernelTimer()
_lib.set_timer(prof)
for x, y in bs[6:]:
    graphs.eager_step(m, opt, lossf, x, y)
_lib.set_timer(None)
for k, (n, ms_, _) in prof.summary().items():
    print(f"{k:32s} {n:3d} {ms_ * 1e3:8.1f} us")
