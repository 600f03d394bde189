"""Train-step throughput of every p_model class on the C2 shape (B = 65536, F = 15, D = 10, N = 10M uniform ids) plus the per-call
time of the tail kernels (KernelTimer).  usage: tails_bench.py [steps]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import _lib, graphs, optim, pretrain_main as PM

dev = torch.device("cuda:0")
B, F, D, N = 65536, 15, 10, 10_000_000
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
gen = torch.Generator(device=dev).manual_seed(1)
per = N // F


def batch():
    x = torch.randint(0, per, (B, F), generator=gen, device=dev, dtype=torch.int64) + torch.arange(F, device=dev) * per
    return x, (torch.rand(B, generator=gen, device=dev) < 0.05).to(torch.int64)


lossf = torch.nn.BCELoss()
for name in ("LR", "FM", "FFM", "DeepFM", "W&D", "FNN", "IPNN", "OPNN", "DCN", "AFM"):
    torch.manual_seed(1)
    m = PM.get_model(name, N, F, D)
    with torch.no_grad():
        m.table.mul_(0.1)
    m = m.to(dev).train()
    opt = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    gs = graphs.GraphedTrainStep([(m, opt)], lossf)
    data = [batch() for _ in range(steps + 4)]
    for x, y in data[:4]:
        gs(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for x, y in data[4:]:
        gs(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # eager pass with per-call timing for the tail kernels
    t = _lib.KernelTimer()
    _lib.set_timer(t)
    for x, y in data[:3]:
        graphs.eager_step(m, opt, lossf, x, y)
    _lib.set_timer(None)
    calls = {k: round(v[1] * 1e3, 1) for k, v in t.summary().items()
             if any(s in k for s in ("afm", "cross", "fieldsq", "pairdots", "ffm"))}
    print(json.dumps({"model": name, "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3), "graph": gs.graph is not None,
                      "tail_kernels_us": calls}), flush=True)
    del m, opt, gs, data
    torch.cuda.empty_cache()
