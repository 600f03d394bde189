import ctypes as C, sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
torch.set_printoptions(linewidth=200, precision=0, sci_mode=False)
def probe(B, K, N):
    # dX[B,N] = dY[B,K] @ W[K,N];  W[k,n] = 100*k + n ; dY one-hot rows: row b selects k = b % K
    W = (torch.arange(K).view(K, 1) * 100 + torch.arange(N).view(1, N)).float().to(dev).contiguous()
    dy = torch.zeros(B, K, device=dev)
    dy[torch.arange(B), torch.arange(B) % K] = 1
    x = torch.zeros(B, N, device=dev)
    dx = torch.full((B, N), -1.0, device=dev)
    ws = torch.empty(lib.rlctr_mlp_ws_bytes(B, N, K), dtype=torch.uint8, device=dev)
    rc = lib.rlctr_linear_bwd(_lib.ptr(x), _lib.ptr(W), None, _lib.ptr(dy), _lib.ptr(dx), None, None, B, N, K, 0, _lib.ptr(ws), ws.numel(), _lib.stream())
    torch.cuda.synchronize()
    ref = dy @ W
    print("rc", rc, "B,K,N", B, K, N, "max err", (dx - ref).abs().max().item())
    print("got  row0..3:", dx[:4, :min(N, 12)].cpu())
    print("want row0..3:", ref[:4, :min(N, 12)].cpu())
    print("got  row 9,17:", dx[[9, 17], :min(N, 12)].cpu())
probe(128, 32, 16)
probe(128, 32, 32)
probe(128, 8, 64)

def probe_w(B, K, N):
    # dW[K_out, N_in] = dY^T X ; dY[b,k] one-hot k = b % K ... use small ints so sums are exact
    x = (torch.arange(B).view(B, 1) % 7 + torch.arange(N).view(1, N)).float().to(dev).contiguous()
    dy = torch.zeros(B, K, device=dev); dy[torch.arange(B), torch.arange(B) % K] = 1
    W = torch.zeros(K, N, device=dev)
    dw = torch.full((K, N), -1.0, device=dev)
    ws = torch.empty(lib.rlctr_mlp_ws_bytes(B, N, K), dtype=torch.uint8, device=dev)
    rc = lib.rlctr_linear_bwd(_lib.ptr(x), _lib.ptr(W), None, _lib.ptr(dy), None, _lib.ptr(dw), None, B, N, K, 0, _lib.ptr(ws), ws.numel(), _lib.stream())
    torch.cuda.synchronize()
    ref = dy.t() @ x
    print("WGRAD rc", rc, "B,K,N", B, K, N, "max err", (dw - ref).abs().max().item())
    print("got :", dw[:3, :12].cpu()); print("want:", ref[:3, :12].cpu())
probe_w(64, 32, 16)
