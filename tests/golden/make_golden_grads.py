#!/usr/bin/env python
"""First-step parameter GRADIENTS of the src/all_main policy nets from the REAL reference classes (build container only).

    python tests/golden/make_golden_grads.py      # needs /root/reference (read-only); writes ref_golden_grads.npz

After several Adam steps two correct fp32 implementations disagree wherever an element's gradient is rounding-level noise
(Adam normalises every step to ~lr), which is why the multi-step DDQN / DDPG states are held to a statistical bar.  The
gradients themselves are well conditioned: this file pins them -- DoubleDQN.learn (DDQN_model.py:198-224), DDPG.learn_c /
learn_a (DDPG_for_PG_model.py:227-251) on one batch from a known initial state -- so that the CUDA path can be held to 1e-5.
"""
import importlib
import os
import sys

import numpy as np
import torch

REF = os.environ.get("RLCTR_REF_PATH", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden_grads.npz")
G = {}


def put(key, val):
    # .copy(): a CPU tensor's .numpy() ALIASES its storage, and the optimizer steps below update the parameters in place
    G[key] = val.detach().cpu().numpy().copy() if isinstance(val, torch.Tensor) else np.array(val)


def state(model, prefix):
    for k, v in model.state_dict().items():
        put(f"{prefix}/{k}", v)


def grads(model, prefix):
    for k, p in model.named_parameters():
        put(f"{prefix}/{k}", p.grad)


def main():
    DQ = importlib.import_module("src.models.DDQN_model")
    DP = importlib.import_module("src.models.DDPG_for_PG_model")
    Fn, Dn, M, b = 15, 10, 3, 256
    g = torch.Generator().manual_seed(31)
    s0 = torch.randn(b, 255, generator=g) * 0.3
    s1 = torch.randn(b, 255, generator=g) * 0.3
    a0 = torch.randint(2, M + 1, (b, 1), generator=g)
    r0 = (torch.rand(b, 1, generator=g) < 0.5).float() * 2 - 1
    w0 = torch.softmax(torch.randn(b, M, generator=g), dim=1)
    for k, v in (("s0", s0), ("s1", s1), ("a0", a0), ("r0", r0), ("w0", w0)):
        put(k, v)
    torch.manual_seed(15)
    dq = DQ.DoubleDQN(1000, Fn, Dn, action_nums=M, memory_size=512, batch_size=b, device="cpu")
    state(dq.eval_net, "ddqn/eval_init")
    dq.learn(s0, a0, r0, s1)                          # returns nothing in the reference
    grads(dq.eval_net, "ddqn/grad")                   # Adam.step does not clear .grad: these are the first step's gradients
    torch.manual_seed(16)
    dp = DP.DDPG(1000, Fn, Dn, action_nums=M, memory_size=512, batch_size=b, device="cpu")
    for nm in ("Actor", "Critic", "Actor_", "Critic_"):
        state(getattr(dp, nm), f"ddpg/{nm}_init")
    da = a0.float()
    import copy
    probe = copy.deepcopy(dp)                         # forward intermediates of learn_c, on a copy (train-mode BN updates buffers)
    put("ddpg/probe_actor_target", probe.Actor_.forward(s1, da))
    put("ddpg/probe_q_target", r0 + probe.gamma * probe.Critic_.forward(s1, probe.Actor_.forward(s1, da), da))
    put("ddpg/probe_q", probe.Critic.forward(s0, w0, da))
    put("ddpg/td_error", dp.learn_c(s0, w0, r0, s1, da))
    grads(dp.Critic, "ddpg/critic_grad")
    state(dp.Critic, "ddpg/Critic_after_c")           # the critic the actor step differentiates through
    # The actor's gradient is ill-conditioned by construction: d a_loss / d Q is the SAME constant for every sample, and
    # every BatchNorm on the way back removes the batch-constant part of the gradient, so what reaches the actor is a small
    # remainder of cancelling terms.  Two fp32 evaluations differ there by ~1e-4 relative; the arbiter is the same
    # computation in float64 (the reference's own modules, cast with .double()).
    d64 = copy.deepcopy(dp)
    d64.Actor.double(); d64.Critic.double()
    a64 = -d64.Critic.forward(s0.double(), d64.Actor.forward(s0.double(), da.double()), da.double()).mean()
    d64.Actor.zero_grad()
    a64.backward()
    grads(d64.Actor, "ddpg/actor_grad64")
    put("ddpg/a_loss", dp.learn_a(s0, da))
    grads(dp.Actor, "ddpg/actor_grad")
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
