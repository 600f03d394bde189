import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import p_model
B, F, D = 65536, 15, 10
rows = (torch.randn(B, F * D, device="cuda") * 0.5).requires_grad_(True)
packed = (torch.randn(D * D + 3 * D + 2, device="cuda") * 0.3).requires_grad_(True)
rng = torch.tensor([1, 0], dtype=torch.int64, device="cuda")
g = torch.randn(B, 1, device="cuda")
for _ in range(3):
    y = p_model._AFMAttention.apply(rows, packed, F, D, 0.2, rng, None)
    y.backward(g)
torch.cuda.synchronize()
print("ok")
