import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import state_from_golden
from rl_ctr_prediction_b200 import DDPG_for_PG_model
G = np.load("tests/golden/ref_golden_grads.npz")
DEV = "cuda:0"
b = G["s0"].shape[0]
dp = DDPG_for_PG_model.DDPG(1000, 15, 10, action_nums=3, memory_size=512, batch_size=b, device=DEV)
for nm in ("Actor", "Critic", "Actor_", "Critic_"):
    getattr(dp, nm).load_state_dict({k: torch.as_tensor(v) for k, v in state_from_golden(G, f"ddpg/{nm}_init").items()})
t = lambda k: torch.as_tensor(G[k]).to(DEV)
da = t("a0").float()
def rel(a, k):
    b_ = G[k]; a = a.detach().cpu().numpy()
    return float(np.abs(a - b_).max() / np.abs(b_).max())
print("actor_target", rel(dp.Actor_.forward(t("s1"), da), "ddpg/probe_actor_target"))
print("q_target", rel(t("r0") + dp.gamma * dp.Critic_.forward(t("s1"), dp.Actor_.forward(t("s1"), da), da), "ddpg/probe_q_target"))
print("q", rel(dp.Critic.forward(t("s0"), t("w0"), da), "ddpg/probe_q"))
print("modes", dp.Actor_.training, dp.Critic_.training, dp.Critic.training)
q_target = t("r0") + dp.gamma * dp.Critic_.forward(t("s1"), dp.Actor_.forward(t("s1"), da), da).detach()
q = dp.Critic.forward(t("s0"), t("w0"), da)
print("td manual", dp.loss_func(q, q_target).item(), "golden", float(G["ddpg/td_error"]))
print("td from golden probes", float(((G["ddpg/probe_q"] - G["ddpg/probe_q_target"]) ** 2).mean()))
# ---- stock torch replica of the Critic from the golden init (CPU)
import torch.nn as nn
sd = state_from_golden(G, "ddpg/Critic_init")
layers, d = [], 259
for w in (300, 300, 300):
    layers += [nn.Linear(d, w), nn.BatchNorm1d(w), nn.ReLU()]
    d = w
layers.append(nn.Linear(d, 3))
rep = nn.Module(); rep.mlp = nn.Sequential(*layers); rep.bn_input = nn.BatchNorm1d(1)
rep.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
rep.train()
s0c, w0c, dac = torch.as_tensor(G["s0"]), torch.as_tensor(G["w0"]), torch.as_tensor(G["a0"]).float()
x = torch.cat([torch.cat([s0c, rep.bn_input(dac)], 1), w0c], 1)
print("replica q vs golden", float((rep.mlp(x) - torch.as_tensor(G["ddpg/probe_q"])).abs().max()))
# ours, layer by layer
C_ = dp.Critic
C_.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
C_.train()
xo = torch.cat([torch.cat([t("s0"), C_.bn_input(da)], 1), t("w0")], 1)
print("input diff", float((xo.cpu() - x).abs().max()))
h_ref, h = x, xo
mods_r, mods_o = list(rep.mlp), list(C_.mlp)
for i, (mr, mo) in enumerate(zip(mods_r, mods_o)):
    h_ref = mr(h_ref)
    h = mo(h)
    print(i, type(mo).__name__, "maxdiff", float((h.detach().cpu() - h_ref.detach()).abs().max()), "scale", float(h_ref.abs().max()))
print("whole tower", float((C_.mlp(xo).detach().cpu() - rep.mlp(x).detach()).abs().max()))
