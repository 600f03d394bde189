#!/usr/bin/env python
"""Generate golden vectors from the REAL reference modules (build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

Imports ``src.models.p_model`` / ``Feature_embedding`` / ``PG_model`` and
``src.all_main.main`` / ``hybrid_td3_main_per`` from the reference tree, runs them on
small seeded inputs on CPU (torch fp32) and writes ``tests/golden/ref_golden.npz``.
The reference has no tests or fixtures of its own, so these vectors are what pins
``oracle/`` (and, through it, the CUDA path).  The reference tree does not exist on
the GPU box; only the committed ``.npz`` travels.
"""
import importlib
import os
import sys

import numpy as np
import torch

REF = os.environ.get("RLCTR_REF_PATH", "/root/reference")
sys.path.insert(0, REF)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
G = {}


def put(key, val):
    if isinstance(val, torch.Tensor):
        val = val.detach().cpu().numpy()
    G[key] = np.array(val, copy=True)      # parameters are updated in place later: snapshot


def zipf_ids(rng, B, F, N, a=1.3):
    """ids with heavy collisions, disjoint-ish per-field ranges, some shared across fields."""
    per = N // F
    out = np.zeros((B, F), dtype=np.int64)
    for f in range(F):
        r = np.minimum(rng.zipf(a, size=B) - 1, per - 1)
        out[:, f] = f * per + r
    # a few cross-field repeats inside one sample (autograd accumulates both)
    out[0, 1] = out[0, 0]
    out[1, 5] = out[1, 9]
    return out


def scale_params(model, s):
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "embedding" in k or k == "linear.weight":
                p.mul_(s)


def state(model, prefix):
    for k, v in model.state_dict().items():
        put(f"{prefix}/{k}", v)


def build(P, name, N, F, D):
    torch.manual_seed(1)
    return {"LR": lambda: P.LR(N), "FM": lambda: P.FM(N, D), "FFM": lambda: P.FFM(N, F, D),
            "DeepFM": lambda: P.DeepFM(N, F, D)}[name]()


def main():
    torch.set_num_threads(1)
    P = importlib.import_module("src.models.p_model")
    FE = importlib.import_module("src.models.Feature_embedding")
    PM = importlib.import_module("src.main.pretrain_main")
    put("meta/torch_version", np.array(torch.__version__))

    # ---- 1. known-answer case of SURVEY section 4 ------------------------------------
    N, F, D, B = 64, 15, 10, 4
    x = torch.from_numpy(((np.arange(B * F).reshape(B, F) * 7 + 3) % N).astype(np.int64))
    y = torch.tensor([1, 0, 0, 1]).view(-1, 1)
    put("kat/x", x)
    put("kat/y", y)
    for name in ("LR", "FM", "FFM", "DeepFM"):
        m = build(P, name, N, F, D)
        scale_params(m, 0.1)
        m.eval()
        state(m, f"kat/{name}/init")
        p = m(x)
        loss = torch.nn.BCELoss()(p, y.float())
        m.zero_grad()
        loss.backward()
        put(f"kat/{name}/pctr", p)
        put(f"kat/{name}/loss", loss)
        put(f"kat/{name}/abs_dlinear", m.linear.weight.grad.abs().sum())
    torch.manual_seed(1)
    fe = FE.Feature_Embedding(N, F, D)
    scale_params(fe, 0.1)
    put("kat/FE/weight", fe.feature_embedding.weight)
    put("kat/FE/out", fe(x))

    # ---- 2. multi-step training trajectories (dense Adam + L2) ---------------
    rng = np.random.default_rng(1234)
    N, F, D, B, STEPS = 255, 15, 10, 48, 3
    xs = [torch.from_numpy(zipf_ids(rng, B, F, N)) for _ in range(STEPS)]
    ys = [torch.from_numpy((rng.random(B) < 0.3).astype(np.int64)) for _ in range(STEPS)]
    put("train/x", torch.stack(xs))
    put("train/y", torch.stack(ys))
    for name in ("LR", "FM", "FFM", "DeepFM"):
        m = build(P, name, N, F, D)
        scale_params(m, 0.1)
        m.eval()                                   # dropout off: deterministic tower
        state(m, f"train/{name}/init")
        opt = torch.optim.Adam(params=m.parameters(), lr=1e-3, weight_decay=1e-5)
        lossf = torch.nn.BCELoss()
        for s in range(STEPS):
            feats, labels = xs[s].long(), torch.unsqueeze(ys[s], 1)
            p = m(feats)                           # body of pretrain_main.train :96-102
            tl = lossf(p, labels.float())
            m.zero_grad()
            tl.backward()
            if s == 0:
                for k, prm in m.named_parameters():
                    put(f"train/{name}/grad0/{k}", prm.grad)
            opt.step()
            put(f"train/{name}/pctr{s}", p)
            put(f"train/{name}/loss{s}", tl)
        state(m, f"train/{name}/final")

    # the loop API itself: reference train()/test() drive a reference FM for one epoch
    m = build(P, "FM", N, F, D)
    scale_params(m, 0.1)
    opt = torch.optim.Adam(params=m.parameters(), lr=1e-3, weight_decay=1e-5)
    loader = [(xs[s], ys[s]) for s in range(STEPS)]
    avg = PM.train(m, opt, loader, torch.nn.BCELoss(), torch.device("cpu"))
    auc, tloss = PM.test(m, loader, torch.nn.BCELoss(), torch.device("cpu"))
    put("loop/FM/train_avg_loss", avg)
    put("loop/FM/test_auc", auc)
    put("loop/FM/test_loss", tloss)
    state(m, "loop/FM/final")

    # ---- 3. saturated regime: default N(0,1) init ---------------------------
    m = build(P, "FM", N, F, D)
    state(m, "sat/FM/init")
    opt = torch.optim.Adam(params=m.parameters(), lr=1e-3, weight_decay=1e-5)
    for s in range(2):
        p = m(xs[s])
        tl = torch.nn.BCELoss()(p, ys[s].view(-1, 1).float())
        m.zero_grad()
        tl.backward()
        if s == 0:
            put("sat/FM/grad0/feature_embedding.weight", m.feature_embedding.weight.grad)
            put("sat/FM/grad0/linear.weight", m.linear.weight.grad)
        opt.step()
        put(f"sat/FM/pctr{s}", p)
        put(f"sat/FM/loss{s}", tl)
    state(m, "sat/FM/final")

    # ---- 4. other latent dims (src/main default D=8; D=16) ------------------------------
    for D2 in (8, 16):
        m = build(P, "FM", N, F, D2)
        scale_params(m, 0.1)
        state(m, f"dims/FM{D2}/init")
        opt = torch.optim.Adam(params=m.parameters(), lr=1e-3, weight_decay=1e-5)
        p = m(xs[0])
        tl = torch.nn.BCELoss()(p, ys[0].view(-1, 1).float())
        m.zero_grad()
        tl.backward()
        opt.step()
        put(f"dims/FM{D2}/pctr0", p)
        put(f"dims/FM{D2}/loss0", tl)
        state(m, f"dims/FM{D2}/final")

    # ---- 5. Feature_Embedding on a bigger case -------------------------------------------
    torch.manual_seed(3)
    fe = FE.Feature_Embedding(N, F, D)
    put("fe/weight", fe.feature_embedding.weight)
    put("fe/x", xs[1])
    put("fe/out", fe(xs[1]))

    # ---- 6. generate_preds: canonical (all_main/main.py) and TD3-PER variant -------------
    AM = importlib.import_module("src.all_main.main")
    TD = importlib.import_module("src.all_main.hybrid_td3_main_per")

    class Frozen(torch.nn.Module):
        def __init__(self, col):
            super().__init__()
            self.col = col

        def forward(self, feats):
            return self.col

    for M in (3, 5, 6):
        Bp = 240
        g = torch.Generator().manual_seed(100 + M)
        pctr = torch.rand(Bp, M, generator=g)
        w = torch.softmax(torch.randn(Bp, M, generator=g) * 1.5, dim=1)
        lab = (torch.rand(Bp, 1, generator=g) < 0.4).long()
        feats = torch.zeros(Bp, F, dtype=torch.long)
        md = {i: Frozen(pctr[:, i:i + 1]) for i in range(M)}
        act = torch.randint(2, M + 1, (Bp, 1), generator=g)
        yy, ww, rr = AM.generate_preds(md, feats, act, w, lab, torch.device("cpu"), mode="train")
        for k, v in (("pctr", pctr), ("w", w), ("label", lab), ("action", act), ("y", yy), ("w_out", ww), ("reward", rr)):
            put(f"gp/v0_M{M}/{k}", v)
        act1 = torch.randint(1, M + 1, (Bp, 1), generator=g)
        yy, rr = TD.generate_preds(md, feats, act1, w, lab, torch.device("cpu"), mode="train")
        for k, v in (("action", act1), ("y", yy), ("reward", rr)):
            put(f"gp/v1_M{M}/{k}", v)

    # ---- 7. REINFORCE pieces (PG_model.py:104-107,139-154) -------------------------------
    PG = importlib.import_module("src.models.PG_model")
    pg = PG.PolicyGradient.__new__(PG.PolicyGradient)     # formulas only; Net is dimensionally broken (N9)
    pg.gamma = 1
    g = torch.Generator().manual_seed(7)
    A, Bq = 4, 64
    logits = torch.randn(Bq, A, generator=g, requires_grad=True)
    probs = torch.softmax(logits, dim=1)
    acts = torch.randint(1, A + 1, (Bq, 1), generator=g)
    rs = (torch.rand(Bq, generator=g) < 0.5).float() * 2 - 1
    pg.ep_rs = rs
    vt_norm = pg.discount_and_norm_rewards()
    vt = torch.FloatTensor(vt_norm)
    loss = pg.loss_func(probs, acts, vt)
    loss.backward()
    put("pg/logits", logits)
    put("pg/acts", acts)
    put("pg/rs", rs)
    put("pg/vt_norm", vt_norm)
    put("pg/loss_literal", loss)
    put("pg/dlogits_literal", logits.grad)
    put("pg/logp", torch.log(probs.gather(1, acts - 1)).view(-1))
    # non-degenerate second case: un-normalised returns
    vt2 = torch.randn(Bq, generator=g)
    logits2 = logits.detach().clone().requires_grad_(True)
    loss2 = pg.loss_func(torch.softmax(logits2, dim=1), acts, vt2)
    loss2.backward()
    put("pg/vt_raw", vt2)
    put("pg/loss_literal_raw", loss2)
    put("pg/dlogits_literal_raw", logits2.grad)

    # ---- 8. policy networks of src/all_main/main.py: DDQN (DDQN_model.py) and DDPG (DDPG_for_PG_model.py) ----
    DQ = importlib.import_module("src.models.DDQN_model")
    DP = importlib.import_module("src.models.DDPG_for_PG_model")
    Fn, Dn, M, b = 15, 10, 3, 64
    g = torch.Generator().manual_seed(21)
    torch.manual_seed(5)
    dq = DQ.DoubleDQN(1000, Fn, Dn, action_nums=M, memory_size=256, batch_size=b, device="cpu")
    state(dq.eval_net, "ddqn/eval_init")
    s0 = torch.randn(b, 255, generator=g) * 0.3
    s1 = torch.randn(b, 255, generator=g) * 0.3
    a0 = torch.randint(2, M + 1, (b, 1), generator=g)
    r0 = (torch.rand(b, 1, generator=g) < 0.5).float() * 2 - 1
    put("ddqn/s0", s0); put("ddqn/s1", s1); put("ddqn/a0", a0); put("ddqn/r0", r0)
    dq.eval_net.eval()
    put("ddqn/q_eval_mode", dq.eval_net(s0))
    put("ddqn/best_action", dq.choose_best_action(s0))
    dq.eval_net.train()
    for it in range(2):
        dq.learn(s0, a0, r0, s1)
    state(dq.eval_net, "ddqn/eval_final")
    state(dq.target_net, "ddqn/target_final")

    torch.manual_seed(6)
    dp = DP.DDPG(1000, Fn, Dn, action_nums=M, memory_size=256, batch_size=b, device="cpu")
    for nm in ("Actor", "Critic", "Actor_", "Critic_"):
        state(getattr(dp, nm), f"ddpg/{nm}_init")
    w0 = torch.softmax(torch.randn(b, M, generator=g), dim=1)
    da = a0.float()
    put("ddpg/w0", w0)
    dp.Actor.eval()
    put("ddpg/actor_eval", dp.Actor(s0, da))
    dp.Actor.train()
    tds, als = [], []
    for it in range(2):
        tds.append(dp.learn_c(s0, w0, r0, s1, da))
        als.append(dp.learn_a(s0, da))
        dp.soft_update(dp.Actor, dp.Actor_)
        dp.soft_update(dp.Critic, dp.Critic_)
    put("ddpg/td_errors", np.array(tds)); put("ddpg/a_losses", np.array(als))
    for nm in ("Actor", "Critic", "Actor_", "Critic_"):
        state(getattr(dp, nm), f"ddpg/{nm}_final")

    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
