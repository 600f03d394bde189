import sys, copy, torch
sys.path.insert(0, ".")
from rl_ctr_prediction_b200 import mlp, DDQN_model
DEV = "cuda:0"
for B, N in [(256, 300), (16, 5), (1000, 33), (4096, 128)]:
    torch.manual_seed(1)
    net = DDQN_model.bn_mlp(N, 3, hidden=(64, 32), device=DEV).train()
    twin = copy.deepcopy(net)
    inp = torch.randn(B, N, device=DEV)
    out = net(inp); out.sum().backward()
    saved, mlp.BN_FUSED_MAX_BATCH = mlp.BN_FUSED_MAX_BATCH, 0
    out_ref = twin(inp); out_ref.sum().backward()
    mlp.BN_FUSED_MAX_BATCH = saved
    print(B, N, "out", float((out - out_ref).abs().max()), float(out_ref.abs().max()))
    for (k, p), (_, q) in zip(net.named_parameters(), twin.named_parameters()):
        print("   ", k, float((p.grad - q.grad).abs().max()), float(q.grad.abs().max()))
    # layer by layer forward
    h1 = net[0](inp)
    a = mlp.bn_relu(copy.deepcopy(net[1]), h1, True)
    b = torch.relu(copy.deepcopy(twin[1])(h1))
    print("    bn1 fwd diff", float((a - b).abs().max()), "flips", int(((a > 0) != (b > 0)).sum()))
