"""Time the co-located step's kernels on their own (eager, CUDA events): group fwd / catch-up / update.  usage: group_probe.py [steps]"""
import sys, torch
sys.path.insert(0, ".")
from rl_ctr_prediction_b200 import _lib, optim, p_model, colocated, graphs
dev = torch.device("cuda", 0)
N, B, F, D = 10_000_000, 65536, 15, 10
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
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
bs = [batch() for _ in range(steps)]
for x, y in bs[:4]:
    g.train_step(x, y, opt)
prof = _lib.KernelTimer()
_lib.set_timer(prof)
for x, y in bs[4:]:
    g.train_step(x, y, opt)
_lib.set_timer(None)
for k, (n, ms_, _) in prof.summary().items():
    print(f"{k:32s} {n:3d} {ms_ * 1e3:8.1f} us")
