import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import mlp
torch.manual_seed(0)
for B in (64, 256, 257, 512):
    for K, pitch in ((255, 256), (256, 256), (259, 259), (259, 260), (300, 300)):
        for N in (300, 3, 2):
            lin = mlp.Linear(K, N, device="cuda")
            buf = torch.randn(B, pitch, device="cuda")
            x = buf[:, :K]
            y = lin(x, relu=False)
            ref = (x.double() @ lin.weight.double().t() + lin.bias.double())
            err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
            flag = "" if err < 1e-5 else "   <<<<<< BAD"
            print(f"B={B} K={K} ld={pitch} N={N} relerr={err:.2e}{flag}")
