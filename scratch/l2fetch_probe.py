"""Does cudaLimitMaxL2FetchGranularity change the cost of the random 64 B row gather / the 192 B record RMW?"""
import ctypes, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import _lib
import ctypes as C
lib = _lib.load()
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name); break
    except OSError:
        pass
dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
LIMIT = 0x05   # cudaLimitMaxL2FetchGranularity
def get():
    v = ctypes.c_size_t(0); rc = rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT); return rc, v.value
N, rs = 10_000_000, 48
tab = torch.randn(N, rs, device=dev)
ids = torch.randint(0, N, (983040,), device=dev, dtype=torch.int64)
out = torch.empty(ids.numel(), 16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run():
    t = _lib.Table(tab.data_ptr(), N, 16, 0, 1, 10, rs)
    ts = []
    for _ in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.rlctr_gather_rows(ids.data_ptr(), ids.numel(), C.byref(t), out.data_ptr(), _lib.stream()); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
print("default", get(), "gather us", round(run(), 1))
for g in (32, 64, 128):
    rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(g))
    print("set", g, "rc", rc, "now", get(), "gather us", round(run(), 1))
