#!/usr/bin/env python
"""Golden vectors for one ``learn`` call (k_epochs = 3 update steps) of the reference's hybrid PPO agent.

    python tests/golden/make_golden_ppo.py        # build container only: needs /root/reference (read-only)

The REAL ``Hybrid_PPO_Model`` of ``src/models/Hybrid_PPO_model.py`` on CPU (its hard-coded ``.cuda()`` patched to the identity).
``learn`` is deterministic given its inputs (the rollout is an argument); the discrete head's bias is raised by 6 so that the raw
outputs the reference feeds to ``Categorical`` as probabilities (:91) are positive.  Stores the initial network, the rollout, the
GAE advantages of the reference's Python loop (recomputed here from the same deltas), the returned loss and the final network
(large tensors as every 17th element + their sum).  Writes ``tests/golden/ref_golden_ppo.npz``.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RLCTR_REF_PATH", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(HERE, "ref_golden_ppo.npz")
G = {}


def put(key, val):
    if isinstance(val, torch.Tensor):
        val = val.detach().cpu().numpy()
    G[key] = np.array(val, copy=True)


def state(mod, prefix, compact=False):
    for k, v in mod.state_dict().items():
        v = v.detach().cpu().numpy()
        if compact and v.size > 5000:
            put(f"{prefix}/{k}/sub", v.reshape(-1)[::17])
            put(f"{prefix}/{k}/sum", np.float64(v.astype(np.float64).sum()))
        else:
            put(f"{prefix}/{k}", v)


def main():
    torch.set_num_threads(1)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        P = importlib.import_module("src.models.Hybrid_PPO_model")
        F_, D, A, n = 15, 10, 3, 96
        in_dims = F_ * (F_ - 1) // 2 + F_ * D
        torch.manual_seed(8)
        agent = P.Hybrid_PPO_Model(500, F_, D, A, memory_size=128, batch_size=32, init_lr=1e-3, device="cpu")
        with torch.no_grad():
            agent.hybrid_actor_critic.Discrete_Actor.bias.add_(6.0)
        state(agent.hybrid_actor_critic, "init")
        s = torch.randn(n, in_dims) * 0.5
        old_c_a = torch.softmax(torch.randn(n, A), dim=-1) + 0.3 * torch.randn(n, A)
        old_c_lp = -0.5 * torch.randn(n, A) ** 2 - 0.92
        old_d_a = torch.randint(0, A - 1, (n, 1))
        old_d_lp = torch.log(torch.full((n, 1), 1.0 / (A - 1))) + 0.05 * torch.randn(n, 1)
        rewards = (torch.rand(n, 1) < 0.5).float() * 2 - 1
        for k, v in (("states", s), ("old_c_a", old_c_a), ("old_c_lp", old_c_lp), ("old_d_a", old_d_a), ("old_d_lp", old_d_lp),
                     ("rewards", rewards)):
            put(f"in/{k}", v)
        # the advantages of the reference's loop, from the same deltas (train-mode evaluate, as learn() does)
        agent.hybrid_actor_critic.train()
        with torch.no_grad():
            probe = P.Hybrid_Actor_Critic(in_dims, A)
            probe.load_state_dict(agent.hybrid_actor_critic.state_dict())
            probe.train()
            v_ = probe.evaluate(s, old_c_a, old_d_a)[0]
            v = probe.evaluate(s, old_c_a, old_d_a)[0]
            deltas = rewards + agent.gamma * v_ - v
        adv = torch.zeros(n, 1)
        a = 0.0
        for i, d in enumerate(reversed(deltas)):
            a = agent.gamma * agent.lamda * a + d.item()
            adv[i, :] = a
        put("gae/deltas", deltas)
        put("gae/advantages", adv)
        loss = agent.learn(s, s, old_c_a, old_c_lp, old_d_a, old_d_lp, rewards)
        put("loss", np.float64(loss))
        state(agent.hybrid_actor_critic, "final", compact=True)
    finally:
        torch.Tensor.cuda = real_cuda
    put("meta/torch_version", np.array(torch.__version__))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays", os.path.getsize(OUT), "bytes; loss", loss)


if __name__ == "__main__":
    main()
