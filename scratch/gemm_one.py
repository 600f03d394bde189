"""One shape, one form, a few launches: the ncu target for the TMA GEMM.  usage: gemm_one.py K N ld form(fwd|dgrad|wgrad)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
K, N, ld, form = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
B = int(os.environ.get("GB", 65536))
x = torch.randn(B, ld, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
y = torch.empty(B, N, device=dev); gy = torch.randn(B, N, device=dev); dx = torch.empty(B, K, device=dev); dw = torch.empty(N, K, device=dev)
wsb = lib.rlctr_mlp_ws_bytes(B, K, N); ws = torch.empty(wsb, dtype=torch.uint8, device=dev); st = _lib.stream()
for _ in range(4):
    if form == "fwd":
        rc = lib.rlctr_linear_fwd(x.data_ptr(), ld, w.data_ptr(), b.data_ptr(), y.data_ptr(), B, K, N, 1, 0.0, None, ws.data_ptr(), wsb, st)
    elif form == "dgrad":
        rc = lib.rlctr_linear_bwd(x.data_ptr(), ld, w.data_ptr(), None, gy.data_ptr(), dx.data_ptr(), None, None, B, K, N, 0, 1.0, 1.0, ws.data_ptr(), wsb, st)
    else:
        rc = lib.rlctr_linear_bwd(x.data_ptr(), ld, w.data_ptr(), None, gy.data_ptr(), None, dw.data_ptr(), None, B, K, N, 0, 1.0, 1.0, ws.data_ptr(), wsb, st)
    assert rc == 0, rc
torch.cuda.synchronize()
print("ok")
