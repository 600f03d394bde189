#!/usr/bin/env python
"""Golden vectors for the hybrid SAC agent's networks and prioritized memory.

    python tests/golden/make_golden_sac.py        # build container only: needs /root/reference (read-only)

The REAL reference classes of ``src/models/Hybrid_SAC_model.py`` on CPU, fp32, seeded: C_Actor / D_Actor / Hybrid_Q_network
forward passes (train-mode BatchNorm: batch statistics), ``C_Actor.sample`` with the Gaussian draw of ``Normal.rsample`` recorded
(``torch.distributions.Normal.rsample`` is wrapped so that the eps it draws is stored beside the outputs), ``evaluate``, the
parameter gradients of a critic loss, and the ``Memory`` semantics (``add`` wrap-around with priority max(old, 1),
``batch_update``, IS weights and beta schedule of ``stochastic_sample`` with ``np.random.choice`` replaced by fixed indices).
``Hybrid_RL_Model.learn`` itself cannot be pinned: it raises ``RuntimeError: one of the variables needed for gradient computation
has been modified by an inplace operation`` under the installed torch (its entropy-tuning losses back-propagate through the actor
graphs after the actor optimizers have stepped, :425-447); ``meta/learn_error`` records the message.
Writes ``tests/golden/ref_golden_sac.npz``.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RLCTR_REF_PATH", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(HERE, "ref_golden_sac.npz")
G = {}


def put(key, val):
    if isinstance(val, torch.Tensor):
        val = val.detach().cpu().numpy()
    G[key] = np.array(val, copy=True)


def state(mod, prefix):
    for k, v in mod.state_dict().items():
        put(f"{prefix}/{k}", v)


def main():
    torch.set_num_threads(1)
    S = importlib.import_module("src.models.Hybrid_SAC_model")
    F_, D, A, B = 15, 10, 3, 64
    in_dims = F_ * (F_ - 1) // 2 + F_ * D
    torch.manual_seed(3)
    s = torch.randn(B, in_dims) * 0.5
    a = torch.tanh(torch.randn(B, A))
    put("in/state", s)
    put("in/action", a)
    # ---- continuous actor
    torch.manual_seed(11)
    ca = S.C_Actor(in_dims, A)
    state(ca, "c_actor/init")
    ca.train()
    mean, log_std = ca.forward(s)
    put("c_actor/mean", mean)
    put("c_actor/log_std", log_std)
    real_rsample = S.Normal.rsample
    drawn = []

    def rsample(self, sample_shape=torch.Size()):
        eps = torch.randn(self.loc.shape)
        drawn.append(eps)
        return self.loc + eps * self.scale
    S.Normal.rsample = rsample
    torch.manual_seed(5)
    act, logp = ca.sample(s)
    S.Normal.rsample = real_rsample
    put("c_actor/eps", drawn[0])
    put("c_actor/sample_actions", act)
    put("c_actor/sample_log_prob", logp)
    loss = (logp * 0.3 - act.sum(-1, keepdim=True)).mean()
    ca.zero_grad()
    loss.backward()
    for k, p in ca.named_parameters():
        put(f"c_actor/grad/{k}", p.grad)
    state(ca, "c_actor/after_train_fwd")             # BatchNorm running statistics after the two train-mode forwards
    ca.eval()
    put("c_actor/evaluate", ca.evaluate(s))
    # ---- discrete actor
    torch.manual_seed(12)
    da = S.D_Actor(in_dims, A)
    state(da, "d_actor/init")
    put("d_actor/probs", da.forward(s))
    put("d_actor/evaluate", da.evaluate(s))
    # ---- critics
    torch.manual_seed(13)
    q = S.Hybrid_Q_network(in_dims, A)
    state(q, "critic/init")
    c1, d1, c2, d2 = q.forward(s, a)
    for k, v in (("c_q1", c1), ("d_q1", d1), ("c_q2", c2), ("d_q2", d2)):
        put(f"critic/{k}", v)
    tgt = torch.randn(B, 1)
    put("critic/target", tgt)
    da_idx = torch.randint(0, A, (B, 1))
    put("critic/disc", da_idx)
    closs = ((c1 - tgt).pow(2) + (c2 - tgt).pow(2) + (d1.gather(1, da_idx) - tgt).pow(2) + (d2.gather(1, da_idx) - tgt).pow(2)).mean()
    q.zero_grad()
    closs.backward()
    put("critic/loss", closs)
    for k, p in q.named_parameters():
        put(f"critic/grad/{k}", p.grad)
    # ---- memory
    mem = S.Memory(10, 4, "cpu")
    rs = np.random.default_rng(0)
    adds = []
    for n in (4, 4, 5, 7):
        tr = torch.as_tensor(rs.standard_normal((n, 4)).astype(np.float32))
        adds.append(tr.numpy())
        mem.add(tr)
    for i, t_ in enumerate(adds):
        put(f"memory/add{i}", t_)
    put("memory/after_add/memory", mem.memory)
    put("memory/after_add/priorities", mem.priorities_)
    idx = torch.tensor([7, 2, 9, 0])
    td = torch.tensor([[0.5], [-2.0], [0.01], [1.5]])
    mem.batch_update(idx, td)
    put("memory/update_idx", idx)
    put("memory/update_td", td)
    put("memory/after_update/priorities", mem.priorities_)
    real_choice = np.random.choice
    fixed = np.array([3, 7, 0, 2])
    np.random.choice = lambda *a_, **k_: fixed
    si, batch, isw = mem.stochastic_sample(4)
    np.random.choice = real_choice
    put("memory/sample_idx", si)
    put("memory/sample_batch", batch)
    put("memory/sample_isw", isw)
    put("memory/beta_after", np.float64(mem.beta))
    # ---- learn(): record that the reference cannot run it
    FE = importlib.import_module("src.models.Feature_embedding")
    agent = S.Hybrid_RL_Model(500, F_, D, A, memory_size=64, batch_size=16, device="cpu")
    fe = FE.Feature_Embedding(500, F_, D)
    tr = torch.cat([torch.randint(0, 500, (40, F_)).float(), torch.randn(40, A), torch.randint(1, A + 1, (40, 1)).float(),
                    torch.randint(0, 2, (40, 1)).float() * 2 - 1], dim=1)
    agent.store_transition(tr)
    try:
        agent.learn(fe)
        msg = "ran"
    except RuntimeError as e:
        msg = str(e).split("\n")[0][:200]
    put("meta/learn_error", np.array(msg))
    put("meta/torch_version", np.array(torch.__version__))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays", os.path.getsize(OUT), "bytes; learn():", msg)


if __name__ == "__main__":
    main()
