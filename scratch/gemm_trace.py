"""Per-stage clock stamps of block 0 of gemm3x_tma_kernel (RLCTR_GEMM_DBG): where a pipeline stage's round trip goes.
usage: gemm_trace.py K N ld form"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
K, N, ld, form = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
B = 65536
x = torch.randn(B, ld, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
y = torch.empty(B, N, device=dev); gy = torch.randn(B, N, device=dev); dx = torch.empty(B, K, device=dev); dw = torch.empty(N, K, device=dev)
wsb = lib.rlctr_mlp_ws_bytes(B, K, N); ws = torch.empty(wsb, dtype=torch.uint8, device=dev); st = _lib.stream()
dbg = torch.zeros(256 * 16 + 4 * 160, dtype=torch.int64, device=dev)
def run():
    if form == "fwd":
        rc = lib.rlctr_linear_fwd(x.data_ptr(), ld, w.data_ptr(), b.data_ptr(), y.data_ptr(), B, K, N, 1, 0.0, None, ws.data_ptr(), wsb, st)
    elif form == "dgrad":
        rc = lib.rlctr_linear_bwd(x.data_ptr(), ld, w.data_ptr(), None, gy.data_ptr(), dx.data_ptr(), None, None, B, K, N, 0, 1.0, 1.0, ws.data_ptr(), wsb, st)
    else:
        rc = lib.rlctr_linear_bwd(x.data_ptr(), ld, w.data_ptr(), None, gy.data_ptr(), None, dw.data_ptr(), None, B, K, N, 0, 1.0, 1.0, ws.data_ptr(), wsb, st)
    assert rc == 0, rc
for _ in range(3):
    run()
torch.cuda.synchronize()
os.environ["RLCTR_GEMM_DBG"] = hex(dbg.data_ptr())
run()
torch.cuda.synchronize()
os.environ["RLCTR_GEMM_DBG"] = ""
blk = dbg[4096:].view(160, 4).cpu().numpy()
blk = blk[blk[:, 0] > 0]
if len(blk) and os.environ.get("TRACE_BLOCKS", "1") != "0":
    import numpy as np
    t_start = blk[:, 0].min()
    dur = (blk[:, 1] - blk[:, 0]) / 1e3
    print("blocks", len(blk), "kernel span us", (blk[:, 1].max() - t_start) / 1e3, "start skew us", (blk[:, 0].max() - t_start) / 1e3)
    print("per-CTA duration us: min %.1f p25 %.1f median %.1f p75 %.1f max %.1f" % (dur.min(), np.percentile(dur, 25), np.median(dur), np.percentile(dur, 75), dur.max()))
    order = np.argsort(dur)
    print("fastest (block, sm, us):", [(int(i), int(blk[i, 2]), round(float(dur[i]), 1)) for i in order[:6]])
    print("slowest (block, sm, us):", [(int(i), int(blk[i, 2]), round(float(dur[i]), 1)) for i in order[-6:]])
    print("block 0 us", float(dur[0]))
d = dbg[:4096].view(256, 16).cpu().numpy()
t0 = d[0, 0]
print("it  tma_issue raw_ready conv_done mma_saw_full mma_issued | tile: wait_acc got_acc | epi(start/end at it=tile*kb, +1)")
for i in range(int(os.environ.get('TRACE_ROWS', '40'))):
    r = [int(v - t0) if v else -1 for v in d[i]]
    print(i, "tma", r[0], "raw", r[1], "w3raw", r[12], "conv", r[2], "w3conv", r[11], "| mma: raw", r[8], "full", r[9], "fence", r[3], "mmas", r[10], "commit", r[4], "| tile", r[5], r[6], "| epi", r[7])
