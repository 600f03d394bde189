#!/usr/bin/env python
"""Golden vectors for two ``learn`` steps of the reference's latest hybrid TD3 agent.

    python tests/golden/make_golden_td3.py        # build container only: needs /root/reference (read-only)

The REAL ``Hybrid_TD3_Model`` of ``src/models/v10_Hybrid_TD3_model_PER.py`` on CPU (its hard-coded ``.cuda()`` patched to the identity,
nothing in /root/reference edited), ``policy_freq = 1`` so that both steps run the delayed actor update.  Random draws are pinned
without touching the reference code: ``np.random.choice`` (replay indices) returns fixed indices; ``torch.normal`` is wrapped -- the
first call of a ``learn`` (the perturbation of the next-state discrete logits) returns ``mean + std * eps`` with a recorded eps, the
calls inside ``to_next_state_c_actions`` return ``mean`` (zero noise); the two Gumbel uniforms are the first two draws of the global
generator after the seed and are re-drawn and stored.  Large tensors of the final states are stored as every 17th element plus their
sum (``.../sub``, ``.../sum``) to keep the file small.  Writes ``tests/golden/ref_golden_td3.npz``.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RLCTR_REF_PATH", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(HERE, "ref_golden_td3.npz")
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
    real_cuda, real_normal, real_choice = torch.Tensor.cuda, torch.normal, np.random.choice
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        T = importlib.import_module("src.models.v10_Hybrid_TD3_model_PER")
        FE = importlib.import_module("src.models.Feature_embedding")
        F_, D, A, N, B = 15, 10, 3, 500, 32
        agent = T.Hybrid_TD3_Model(N, F_, D, A, memory_size=64, batch_size=B, device="cpu")
        agent.policy_freq = 1
        torch.manual_seed(5)
        fe = FE.Feature_Embedding(N, F_, D)
        put("fe/feature_embedding.weight", fe.feature_embedding.weight)
        state(agent.Hybrid_Actor, "init/actor")
        state(agent.Hybrid_Critic, "init/critic")
        rs = np.random.default_rng(0)
        tr = np.concatenate([rs.integers(0, N, (48, F_)), np.tanh(rs.standard_normal((48, A))), rs.random((48, A)),
                             rs.integers(1, A + 1, (48, 1)), rs.integers(0, 2, (48, 1))], axis=1).astype(np.float32)
        put("transitions", tr)
        agent.store_transition(torch.as_tensor(tr[:30]))
        agent.store_transition(torch.as_tensor(tr[30:]))
        put("memory/prioritys_after_store", agent.memory.prioritys_)
        gen = torch.Generator().manual_seed(99)
        for step in range(2):
            idx = np.sort(rs.permutation(48)[:B])
            eps_d = torch.randn(B, A, generator=gen)
            calls = [0]

            def normal(mean, std, *a, **k):
                calls[0] += 1
                return mean + std * eps_d if calls[0] == 1 else mean + std * 0.0
            torch.normal = normal
            np.random.choice = lambda *a_, **k_: idx
            torch.manual_seed(1000 + step)
            U_next, U_now = torch.FloatTensor(B, A).uniform_(), torch.FloatTensor(B, A).uniform_()
            torch.manual_seed(1000 + step)
            loss = agent.learn(fe)
            torch.normal, np.random.choice = real_normal, real_choice
            put(f"step{step}/idx", idx); put(f"step{step}/eps_d", eps_d)
            put(f"step{step}/U_next", U_next); put(f"step{step}/U_now", U_now)
            put(f"step{step}/critic_loss", np.float64(loss))
            put(f"step{step}/prioritys_", agent.memory.prioritys_)
        for name, net in (("actor", agent.Hybrid_Actor), ("critic", agent.Hybrid_Critic), ("actor_target", agent.Hybrid_Actor_),
                          ("critic_target", agent.Hybrid_Critic_)):
            state(net, f"final/{name}", compact=True)
    finally:
        torch.Tensor.cuda, torch.normal, np.random.choice = real_cuda, real_normal, real_choice
    put("meta/torch_version", np.array(torch.__version__))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
