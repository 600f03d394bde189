#!/usr/bin/env python
"""Golden vectors for the network heads of the hybrid TD3 (v10) and PPO agents.

    python tests/golden/make_golden_heads.py      # build container only: needs /root/reference (read-only)

The REAL reference classes on CPU, fp32, seeded.  Both modules hard-code ``.cuda()`` (``v10_Hybrid_TD3_model_PER.py:262``,
``Hybrid_PPO_model.py:49``): ``torch.Tensor.cuda`` is patched to the identity while they run (nothing in /root/reference is
edited).  Random draws are made reproducible without touching the reference code: ``torch.normal`` is wrapped so that the
standard-normal eps behind every ``torch.normal(mean, std)`` call is recorded (it returns ``mean + std * eps``), and the uniform of
``gumbel_softmax_sample`` (``torch.FloatTensor(*shape).uniform_()``, the first generator draw after the seed) is re-drawn with the
same seed and stored.  Writes ``tests/golden/ref_golden_heads.npz``.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RLCTR_REF_PATH", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(HERE, "ref_golden_heads.npz")
G = {}


def put(key, val):
    if isinstance(val, torch.Tensor):
        val = val.detach().cpu().numpy()
    G[key] = np.array(val, copy=True)


def state(mod, prefix, only_buffers=False):
    for k, v in mod.state_dict().items():
        if only_buffers and not ("running_" in k or "num_batches" in k):
            continue                                  # weights are unchanged by a forward: keep the file small
        put(f"{prefix}/{k}", v)


def grads(mod, prefix):
    for k, p in mod.named_parameters():
        put(f"{prefix}/{k}", p.grad if p.grad is not None else torch.zeros_like(p))


def main():
    torch.set_num_threads(1)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    real_normal = torch.normal
    try:
        T = importlib.import_module("src.models.v10_Hybrid_TD3_model_PER")
        P = importlib.import_module("src.models.Hybrid_PPO_model")
        F_, D, A, B = 15, 10, 3, 48
        in_dims = F_ * (F_ - 1) // 2 + F_ * D
        torch.manual_seed(21)
        s = torch.randn(B, in_dims) * 0.5
        ca = torch.tanh(torch.randn(B, A))
        dact = torch.softmax(torch.randn(B, A), dim=-1)
        put("in/state", s); put("in/c_actions", ca); put("in/d_actions", dact)
        # ---- TD3 critic
        torch.manual_seed(31)
        cr = T.Hybrid_Critic(in_dims, A)
        state(cr, "td3_critic/init")
        cr.train()
        q1, q2 = cr.evaluate(s, ca, dact)
        put("td3_critic/q1", q1); put("td3_critic/q2", q2)
        tgt = torch.randn(B, 1)
        put("td3_critic/target", tgt)
        loss = (torch.nn.functional.mse_loss(q1, tgt, reduction="none") + torch.nn.functional.mse_loss(q2, tgt, reduction="none")).mean()
        cr.zero_grad(); loss.backward()
        put("td3_critic/loss", loss); grads(cr, "td3_critic/grad")
        state(cr, "td3_critic/after", only_buffers=True)
        cr.eval()
        put("td3_critic/q1_eval", cr.evaluate_q_1(s, ca, dact))
        # ---- TD3 actor
        torch.manual_seed(32)
        ac = T.Hybrid_Actor(in_dims, A)
        state(ac, "td3_actor/init")
        ac.train()
        drawn = []

        def normal(mean, std, *a, **k):
            eps = torch.randn(mean.shape)
            drawn.append(eps)
            return mean + std * eps
        torch.normal = normal
        torch.manual_seed(77)
        e1, e2 = torch.randn(B, A), torch.randn(B, A)         # what the two wrapped calls will draw ...
        U = torch.FloatTensor(B, A).uniform_()                # ... and then the Gumbel uniform
        torch.manual_seed(77)
        c_means, ens_c, d_action, ens_d = ac.act(s, 0.7)
        torch.normal = real_normal
        assert torch.equal(drawn[0], e1) and torch.equal(drawn[1], e2)
        put("td3_actor/eps_c", e1); put("td3_actor/eps_d", e2); put("td3_actor/U", U)
        put("td3_actor/act/c_means", c_means); put("td3_actor/act/ens_c", ens_c)
        put("td3_actor/act/d_action", d_action); put("td3_actor/act/ens_d", ens_d)
        loss = (ens_c * ca).sum(-1).mean() + (d_action * dact).sum(-1).mean() + (c_means ** 2).mean()
        ac.zero_grad(); loss.backward()
        grads(ac, "td3_actor/grad")
        state(ac, "td3_actor/after", only_buffers=True)
        ac.eval()
        c_e, d_e = ac.evaluate(s)
        put("td3_actor/eval/c", c_e); put("td3_actor/eval/d", d_e)
        put("td3/boltzmann", T.boltzmann_softmax(c_e, 0.5))
        # ---- PPO actor-critic head
        torch.manual_seed(41)
        pp = P.Hybrid_Actor_Critic(in_dims, A)
        with torch.no_grad():
            pp.Discrete_Actor.bias.add_(6.0)                  # Categorical(probs=raw outputs) (:91) needs non-negative outputs
        state(pp, "ppo/init")
        pp.train()
        d_a = torch.randint(0, A - 1, (B, 1))
        put("ppo/d_a", d_a)
        sv, clp, cent, dlp, dent = pp.evaluate(s, ca, d_a)
        for k, v in (("state_value", sv), ("c_logprob", clp), ("c_entropy", cent), ("d_logprob", dlp), ("d_entropy", dent)):
            put(f"ppo/evaluate/{k}", v)
        loss = (sv ** 2).mean() - clp.sum(-1).mean() * 0.1 - dlp.mean() - 0.01 * dent.mean()
        pp.zero_grad(); loss.backward()
        grads(pp, "ppo/grad")
        state(pp, "ppo/after", only_buffers=True)
        pp.eval()
        bc, bd = pp.best_a(s)
        put("ppo/best/c", bc); put("ppo/best/d", bd)
    finally:
        torch.Tensor.cuda = real_cuda
        torch.normal = real_normal
    put("meta/torch_version", np.array(torch.__version__))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
