import ctypes as C, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import _lib
from rl_ctr_prediction_b200.tables import Geometry, table_struct
lib = _lib.load()
dev = "cuda:0"
N, n = 10_000_000, 983040
ids = torch.randint(0, N, (n,), device=dev)
def run(rs, reps=20):
    tab = torch.randn(N, rs, device=dev)
    out = torch.empty(n, rs, device=dev)
    g = Geometry(N, rs, 0, 1, 10)
    t = table_struct(tab, g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        lib.rlctr_gather_rows(_lib.ptr(ids), n, C.byref(t), _lib.ptr(out), _lib.stream())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0.record()
        lib.rlctr_gather_rows(_lib.ptr(ids), n, C.byref(t), _lib.ptr(out), _lib.stream())
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts)//2] * 1e3
for rs in (12, 16, 32, 48, 64, 128):
    us = run(rs)
    print(f"row {rs*4:4d} B: {us:7.1f} us  {n/us/1e3:6.2f} G rows/s  read {n*rs*4/us/1e3:7.1f} GB/s (+ same written)")
