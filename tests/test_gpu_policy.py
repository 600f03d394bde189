"""Policy networks of the src/all_main/main.py pipeline on the B200 path (tcgen05 GEMMs + fused dense Adam)
against golden vectors recorded from the REAL reference classes (DDQN_model.DoubleDQN,
DDPG_for_PG_model.DDPG; tests/golden/make_golden.py section 8), plus the whole RL+CTR step."""
import numpy as np
import pytest
import torch

from conftest import state_from_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(a, b, rtol=1e-5, atol=None):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    a, b = a.astype(np.float64), b.astype(np.float64)
    if atol is None:
        atol = rtol * (np.abs(b).max() if b.size else 0.0)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


# Linear biases directly followed by BatchNorm1d (DDQN_model.py:35-37, DDPG_for_PG_model.py:30-32,58-60) and the
# BatchNorm1d(1) shift that the next BatchNorm cancels
BN_SHADOWED = {"mlp.0.bias", "mlp.3.bias", "mlp.6.bias", "bn_input.bias"}


def load(net, golden, prefix):
    sd = {k: torch.as_tensor(v) for k, v in state_from_golden(golden, prefix).items()}
    net.load_state_dict(sd)
    return net


def check_state(net, golden, prefix, rtol=2e-5, travel=2e-3):
    ref = state_from_golden(golden, prefix)
    sd = net.state_dict()
    assert set(sd.keys()) == set(ref.keys())
    for k, v in ref.items():
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
            continue
        # Adam divides each element's step by its own gradient history, so wherever the gradient is
        # rounding-level noise the step direction itself is noise, in the reference as well, and two correct
        # fp32 implementations disagree there by a fraction of lr per step.  This is the case for a handful of
        # weights, and for many 1-D parameters: a bias feeding a BatchNorm has an EXACTLY zero true gradient, and
        # BatchNorm shifts sum cancelling terms over the batch.  Bar: weight matrices >= 99.8 % of the elements within
        # 2e-5 of the scale, and EVERY element of every parameter within the total Adam travel (lr * steps) band.  The
        # functional check (network outputs of the two final states agree) is in the tests below.
        a = sd[k].detach().cpu().numpy().astype(np.float64)
        b = v.astype(np.float64)
        tight = rtol * max(float(np.abs(b).max()), 1e-30) + rtol * np.abs(b)
        bad = np.abs(a - b) > tight
        if b.ndim == 2:
            assert bad.mean() <= 2e-3, (k, bad.mean())
        assert np.abs(a - b).max() <= 2 * travel, (k, np.abs(a - b).max())


def torch_replica(sd, in_dims, out_dims, with_bn_input=False):
    """stock-torch CPU copy of the reference architecture (Linear, BatchNorm1d, ReLU x3, Linear), eval mode."""
    import torch.nn as nn
    layers, d = [], in_dims
    for w in (300, 300, 300):
        layers += [nn.Linear(d, w), nn.BatchNorm1d(w), nn.ReLU()]
        d = w
    layers.append(nn.Linear(d, out_dims))
    net = nn.Module()
    net.mlp = nn.Sequential(*layers)
    if with_bn_input:
        net.bn_input = nn.BatchNorm1d(1)
    net.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    return net.eval()


def test_ddqn_matches_reference(golden):
    from rl_ctr_prediction_b200 import DDQN_model
    torch.manual_seed(0)
    dq = DDQN_model.DoubleDQN(1000, 15, 10, action_nums=3, memory_size=256, batch_size=64, device=DEV)
    load(dq.eval_net, golden, "ddqn/eval_init")
    s0, s1 = torch.as_tensor(golden["ddqn/s0"]).to(DEV), torch.as_tensor(golden["ddqn/s1"]).to(DEV)
    a0, r0 = torch.as_tensor(golden["ddqn/a0"]).to(DEV), torch.as_tensor(golden["ddqn/r0"]).to(DEV)
    dq.eval_net.eval()
    with torch.no_grad():
        close(dq.eval_net(s0), golden["ddqn/q_eval_mode"])
    assert np.array_equal(dq.choose_best_action(s0).cpu().numpy(), golden["ddqn/best_action"])
    dq.eval_net.train()
    for _ in range(2):
        dq.learn(s0, a0, r0, s1)
    check_state(dq.eval_net, golden, "ddqn/eval_final")
    check_state(dq.target_net, golden, "ddqn/target_final")
    # functional parity of the learned Q function: this run's final net vs the reference's final net
    ref = torch_replica(state_from_golden(golden, "ddqn/eval_final"), 255, 2)
    dq.eval_net.eval()
    with torch.no_grad():
        close(dq.eval_net(s1), ref.mlp(s1.cpu()), rtol=2e-3)      # noise-driven +-lr bias steps move Q by O(lr)


def test_ddpg_matches_reference(golden):
    from rl_ctr_prediction_b200 import DDPG_for_PG_model
    dp = DDPG_for_PG_model.DDPG(1000, 15, 10, action_nums=3, memory_size=256, batch_size=64, device=DEV)
    for nm in ("Actor", "Critic", "Actor_", "Critic_"):
        load(getattr(dp, nm), golden, f"ddpg/{nm}_init")
    s0, s1 = torch.as_tensor(golden["ddqn/s0"]).to(DEV), torch.as_tensor(golden["ddqn/s1"]).to(DEV)
    r0 = torch.as_tensor(golden["ddqn/r0"]).to(DEV)
    da = torch.as_tensor(golden["ddqn/a0"]).float().to(DEV)
    w0 = torch.as_tensor(golden["ddpg/w0"]).to(DEV)
    dp.Actor.eval()
    with torch.no_grad():
        close(dp.Actor(s0, da), golden["ddpg/actor_eval"])
    dp.Actor.train()
    tds, als = [], []
    for _ in range(2):
        tds.append(dp.learn_c(s0, w0, r0, s1, da))
        als.append(dp.learn_a(s0, da))
        dp.soft_update(dp.Actor, dp.Actor_)
        dp.soft_update(dp.Critic, dp.Critic_)
    close(np.array(tds), golden["ddpg/td_errors"], rtol=2e-5)
    close(np.array(als), golden["ddpg/a_losses"], rtol=2e-5)
    for nm in ("Actor", "Critic", "Actor_", "Critic_"):
        check_state(getattr(dp, nm), golden, f"ddpg/{nm}_final")
    ref = torch_replica(state_from_golden(golden, "ddpg/Actor_final"), 256, 3, with_bn_input=True)
    dp.Actor.eval()
    with torch.no_grad():
        mine = dp.Actor(s1, da)
        theirs = torch.softmax(ref.mlp(torch.cat([s1.cpu(), ref.bn_input(da.cpu())], dim=1)), dim=1)
    close(mine, theirs, rtol=2e-3)


def test_rl_ctr_step_runs_and_is_consistent():
    """The full src/all_main/main.py step: encoder -> DDQN/DDPG act -> generate_preds over {LR, FM, FFM} -> store ->
    learn.  Checks shapes, value ranges, the replay contents and that generate_preds inside the step equals the
    oracle on the same inputs."""
    from oracle import np_oracle as O
    from rl_ctr_prediction_b200 import all_main, ensemble, p_model
    from rl_ctr_prediction_b200.Feature_embedding import Feature_Embedding
    torch.manual_seed(3)
    N, F, D, B, M = 5000, 15, 10, 512, 3
    models = {0: p_model.LR(N), 1: p_model.FM(N, D), 2: p_model.FFM(N, F, D)}
    for m in models.values():
        with torch.no_grad():
            m.table.mul_(0.1)
        m.to(DEV).eval()
    fe = Feature_Embedding(N, F, D).to(DEV)
    fe.load_embedding(models[1].state_dict())
    ddqn, ddpg = all_main.get_model(M, N, F, D, 128, 4096, DEV, "1458")
    rng = np.random.default_rng(0)
    for it in range(3):
        x = torch.as_tensor(rng.integers(0, N, size=(B, F))).to(DEV)
        y = torch.as_tensor((rng.random((B, 1)) < 0.3).astype(np.int64)).to(DEV)
        y_preds, rewards, td, al = all_main.train_step(ddqn, ddpg, models, x, y, fe, 0.5, DEV)
        assert y_preds.shape == (B, 1) and rewards.shape == (B, 1)
        assert set(np.unique(rewards.cpu().numpy())) <= {-1.0, 1.0}
        assert np.isfinite(td) and np.isfinite(al)
    assert ddqn.memory_counter == 3 * B and ddpg.memory_counter == 3 * B
    stored = ddqn.memory[:3 * B]
    assert torch.equal(stored[2 * B:3 * B, :F].long(), x)                  # ids survive the float32 ring buffer (N < 2^24)
    assert set(np.unique(stored[:, F].cpu().numpy())) <= {2.0, 3.0}
    # generate_preds on the step's own inputs == oracle
    pctr = ensemble.score_models(models, x)
    w = torch.softmax(torch.randn(B, M, device=DEV), dim=1)
    act = torch.randint(2, M + 1, (B, 1), device=DEV)
    yv, wv, rv = ensemble.generate_preds(models, x, act, w, y, DEV, "train", pctr=pctr)
    yo, wo, ro, margin = O.generate_preds(pctr.cpu().numpy(), w.cpu().numpy(), act.cpu().numpy(), y.cpu().numpy(), return_margin=True)
    close(yv, yo)
    close(wv, wo)
    bad = (rv.cpu().numpy() != ro).reshape(-1)
    assert np.all(margin[bad] <= 2e-7), margin[bad]        # a reward may differ only where y == base to rounding (a tie)


@pytest.mark.parametrize("variant", ["literal", "per_sample"])
def test_reinforce_policy_gradient_learn(variant):
    """PG_model.PolicyGradient.learn (state encoder kernel -> tcgen05 MLP -> REINFORCE head kernel -> fused Adam) vs
    the reference's formulas evaluated with stock torch on the CPU (PG_model.py:53-58,104-107,156-179 with the N9
    input-dims fix): per-sample policy log-probs within 1e-5, loss, and the updated network."""
    import torch.nn as nn
    from oracle import torch_port as TP
    from rl_ctr_prediction_b200 import PG_model
    torch.manual_seed(11)
    N, F, D, A, B = 3000, 15, 10, 4, 512
    pg = PG_model.PolicyGradient(N, F, D, action_nums=A, device=DEV, loss_variant=variant)
    for m in pg.policy_net.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0                                      # deterministic comparison (torch's Philox masks differ across devices)
    with torch.no_grad():
        pg.policy_net.embedding_layer.table.mul_(0.2)
    # CPU replica of the same network
    fe = TP.PortFeatureEmbedding(N, F, D)
    fe.load_state_dict(pg.policy_net.embedding_layer.state_dict())
    dims = [255, 1024, 512, 256, 128]
    layers = []
    for i in range(4):
        layers += [nn.Linear(dims[i], dims[i + 1]), nn.ReLU(), nn.Dropout(p=0.0)]
    layers.append(nn.Linear(128, A))
    ref = nn.Sequential(*layers)
    ref.load_state_dict({k: v.cpu() for k, v in pg.policy_net.mlp.state_dict().items()})
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-4, weight_decay=1e-5)
    rng = np.random.default_rng(4)
    x = torch.as_tensor(rng.integers(0, N, size=(B, F)))
    a = torch.as_tensor(rng.integers(1, A + 1, size=(B, 1)))
    r = torch.as_tensor((rng.random(B) < 0.5).astype(np.float32) * 2 - 1)
    vt = torch.as_tensor(rng.standard_normal(B).astype(np.float32))            # raw returns: non-degenerate gradient (N9)
    # act: the acting path (softmax policy) agrees
    with torch.no_grad():
        close(pg.policy_net.forward(x.to(DEV)), torch.softmax(ref(fe(x)), dim=1))
    pg.store_transition(x.to(DEV), a.to(DEV), r.to(DEV))
    assert pg.ep_states.shape == (B, F)
    g = pg.discount_and_norm_rewards()
    gref = np.cumsum(r.numpy()[::-1].astype(np.float64))[::-1]
    close(g, (gref - gref.mean()) / gref.std(), rtol=1e-12)
    loss = pg.learn(vt=vt.to(DEV))
    probs = torch.softmax(ref(fe(x)), dim=1)
    logp_ref = torch.log(probs.gather(1, a - 1)).view(-1)
    if variant == "literal":
        loss_ref = torch.mean(torch.mul(torch.sum(-logp_ref), vt))             # PG_model.py:105-106
    else:
        loss_ref = torch.mean(-logp_ref * vt)
    ropt.zero_grad()
    loss_ref.backward()
    ropt.step()
    close(pg.last_logp, logp_ref.detach(), rtol=1e-5)                          # the north-star quantity
    close(loss, loss_ref.detach(), rtol=1e-4, atol=1e-4 * float(np.abs(logp_ref.detach().numpy()).sum()) * 1e-2)
    for (k, p), (_, q) in zip(pg.policy_net.mlp.state_dict().items(), ref.state_dict().items()):
        # one Adam step of size lr: every element moved by ~1e-4 in the same direction
        assert (p.cpu() - q).abs().max().item() <= 2.2e-4, k
        agree = ((p.cpu() - q).abs() <= 2e-6).float().mean().item()
        assert agree >= 0.98, (k, agree)
    assert pg.ep_states.numel() == 0                                            # episode cleared (:176-179)


# ------------------------------------------------------------------------------------------------
# replay memories sampled on the device
# ------------------------------------------------------------------------------------------------
def _successive_inclusion(w, k):
    """Exact inclusion probabilities of weighted sampling WITHOUT replacement (draw one by one, renormalise): what
    np.random.choice(n, k, p=w/sum(w), replace=False) samples.  Exponential in k: tiny cases only."""
    import itertools
    n = len(w)
    inc = np.zeros(n)
    for perm in itertools.permutations(range(n), k):
        p, rest = 1.0, w.sum()
        for i in perm:
            p *= w[i] / rest
            rest -= w[i]
        inc[list(perm)] += p
    return inc


def test_replay_uniform_sampling_is_a_permutation_prefix():
    from rl_ctr_prediction_b200 import replay
    rng = replay._rng(torch.device(DEV), seed=77)
    for n, k in ((1, 1), (5, 5), (1000, 1000), (100000, 512), (1 << 20, 4096)):
        idx = replay.sample_uniform(n, k, rng).cpu().numpy()
        assert idx.min() >= 0 and idx.max() < n
        assert len(np.unique(idx)) == k                                # distinct, like random.sample
    # same (seed, counter) -> same draw; the counter moves with every call
    a = replay.sample_uniform(5000, 64, replay._rng(torch.device(DEV), seed=5))
    b = replay.sample_uniform(5000, 64, replay._rng(torch.device(DEV), seed=5))
    assert torch.equal(a, b)
    r = replay._rng(torch.device(DEV), seed=5)
    c, d = replay.sample_uniform(5000, 64, r), replay.sample_uniform(5000, 64, r)
    assert not torch.equal(c, d)
    # uniform inclusion: every index of a pool of 50 is drawn ~ draws * k / n times
    r = replay._rng(torch.device(DEV), seed=9)
    n, k, draws = 50, 10, 4000
    cnt = np.zeros(n)
    for _ in range(draws):
        cnt[replay.sample_uniform(n, k, r).cpu().numpy()] += 1
    exp = draws * k / n
    assert np.abs(cnt - exp).max() < 5 * np.sqrt(exp * (1 - k / n))
    with pytest.raises(ValueError):
        replay.sample_uniform(10, 11, r)


def test_replay_memory_matches_reference_semantics():
    """Memory.add wrap-around, gather, IS weights and greedy top-k against a numpy restatement of the reference's Memory
    (v10_Hybrid_TD3_model_PER.py:19-110); stochastic_sample's inclusion frequencies against the exact successive-sampling law."""
    from rl_ctr_prediction_b200 import replay
    rs = np.random.default_rng(0)
    size, T = 10, 3
    mem = replay.Memory(size, T, DEV, seed=3)
    ref_mem, ref_pr, counter = np.zeros((size, T), np.float32), np.zeros((size, 2), np.float32), 0
    for n in (4, 4, 5, 7):                                            # the third and fourth writes wrap
        tr = rs.standard_normal((n, T)).astype(np.float32)
        td = np.abs(rs.standard_normal((n, 1))).astype(np.float32)
        mem.add(torch.as_tensor(td), torch.as_tensor(tr))
        for i in range(n):
            ref_mem[(counter + i) % size] = tr[i]
            ref_pr[(counter + i) % size] = td[i, 0]
        counter += n
    assert mem.memory_counter == counter
    assert np.array_equal(mem.memory.cpu().numpy(), ref_mem) and np.array_equal(mem.prioritys_.cpu().numpy(), ref_pr)
    # stochastic sample: distinct indices, rows = memory[idx], IS weights = (p / min p)^-beta with p = (|td| + eps)^alpha
    mem.beta = 0.7
    idx, batch, isw = mem.stochastic_sample(4)
    i = idx.cpu().numpy()
    assert len(np.unique(i)) == 4
    assert np.array_equal(batch.cpu().numpy(), ref_mem[i])
    pri = (np.abs(ref_pr[:, 0].astype(np.float64)) + 1e-3) ** 0.6
    close(isw[:, 0], (pri[i] / pri.min()) ** -0.7, rtol=1e-5)
    # inclusion frequencies == successive weighted sampling without replacement (np.random.choice(p=P, replace=False))
    small = replay.Memory(6, 1, DEV, seed=11)
    td6 = np.array([[0.05], [0.2], [0.4], [0.9], [1.5], [3.0]], np.float32)
    small.add(torch.as_tensor(td6), torch.zeros(6, 1))
    w = (np.abs(td6[:, 0].astype(np.float64)) + 1e-3) ** 0.6
    inc = _successive_inclusion(w, 3)
    draws, cnt = 6000, np.zeros(6)
    for _ in range(draws):
        cnt[small.stochastic_sample(3)[0].cpu().numpy()] += 1
    freq = cnt / draws
    assert np.abs(freq - inc).max() < 5 * np.sqrt(0.25 / draws), (freq, inc)
    # greedy: the largest raw priorities, IS weights on the raw priorities
    gi, gb, gw = mem.greedy_sample(3)
    order = np.argsort(-ref_pr[:, 0], kind="stable")[:3]
    assert np.array_equal(np.sort(gi.cpu().numpy()), np.sort(order))
    close(gw[:, 0], (ref_pr[gi.cpu().numpy(), 0].astype(np.float64) / ref_pr[:, 0].min()) ** -mem.beta, rtol=1e-5)
    # batch_update writes column 0 only
    mem.batch_update(gi, torch.full((3, 1), 9.0))
    ref_pr[gi.cpu().numpy(), 0] = 9.0
    assert np.array_equal(mem.prioritys_.cpu().numpy(), ref_pr)


def test_per_sampling_frequencies_at_a_million_slots():
    """Weighted sampling without replacement at the size the reference's agents run (memory_size ~ millions): the share of draws
    per priority class and the absence of a slot-index bias inside a class (the exponential clocks use all 32 hash bits:
    with a 24-bit uniform the winning keys were quantised to ~4 slots per level and ties went to the lower slot index)."""
    from rl_ctr_prediction_b200 import replay
    n, k, draws = 1 << 20, 256, 160
    mem = replay.Memory(n, 1, DEV, seed=5)
    td = np.full((n, 1), 0.1, np.float32)
    td[1::2] = 1.6                                                     # odd slots: the heavy class
    mem.add(torch.as_tensor(td), torch.zeros(n, 1))
    w = (np.abs(np.array([0.1, 1.6], np.float64)) + 1e-3) ** 0.6
    share = w[1] / w.sum()                                             # k << n: draws are ~independent, P(heavy) = w1 / (w0 + w1)
    heavy, pos = 0, []
    for _ in range(draws):
        i = mem.stochastic_sample(k)[0].cpu().numpy()
        assert len(np.unique(i)) == k
        heavy += int((i % 2 == 1).sum())
        pos.append(i[i % 2 == 1])
    tot = draws * k
    assert abs(heavy / tot - share) < 5 * np.sqrt(share * (1 - share) / tot), (heavy / tot, share)
    pos = np.concatenate(pos).astype(np.float64)
    # slot indices of the heavy draws are uniform over [0, n): mean n/2 within 5 sigma, no pile-up at low indices
    assert abs(pos.mean() - n / 2) < 5 * n / np.sqrt(12 * len(pos)), pos.mean()
    assert abs((pos < n / 16).mean() - 1 / 16) < 5 * np.sqrt((1 / 16) * (15 / 16) / len(pos))


def test_ddqn_device_sampling():
    """DoubleDQN with device-side replay sampling (RingMemory.device_sampling): distinct in-range indices, stored rows back."""
    from rl_ctr_prediction_b200 import DDQN_model
    torch.manual_seed(3)
    F_, D_ = 15, 10
    agent = DDQN_model.DoubleDQN(300, F_, D_, action_nums=3, memory_size=256, batch_size=32, device=DEV)
    rs = np.random.default_rng(1)
    tr = torch.as_tensor(np.concatenate([rs.integers(0, 300, (200, F_)), rs.integers(2, 4, (200, 1)), rs.integers(0, 2, (200, 1))],
                                        axis=1)).float().to(DEV)
    agent.store_transition(tr)
    agent._mem.device_sampling = True
    idx = agent._mem.sample_index(32, DEV).cpu().numpy()
    assert len(np.unique(idx)) == 32 and idx.min() >= 0 and idx.max() < 200
    b_s, b_a, b_r, b_s_ = agent.sample_batch()                     # the learn step consumes these (all_main/main.py:300-306)
    assert tuple(b_s.shape) == (32, F_) and tuple(b_a.shape) == (32, 1) and tuple(b_r.shape) == (32, 1)
    rows = {tuple(r) for r in tr[:, :F_].long().cpu().numpy().tolist()}
    assert all(tuple(r) in rows for r in b_s.cpu().numpy().tolist())


# ------------------------------------------------------------------------------------------------
# hybrid SAC agent (src/models/Hybrid_SAC_model.py) -- networks and memory against the real reference
# ------------------------------------------------------------------------------------------------
def _sac_load(mod, golden_sac, prefix):
    from conftest import state_from_golden
    sd = state_from_golden(golden_sac, prefix)
    assert set(mod.state_dict().keys()) == set(sd.keys())
    mod.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    return mod.to(DEV)


def test_sac_networks_match_reference(golden_sac):
    from rl_ctr_prediction_b200 import Hybrid_SAC_model as S
    g = golden_sac
    s = torch.as_tensor(g["in/state"]).to(DEV)
    a = torch.as_tensor(g["in/action"]).to(DEV)
    in_dims, A = s.shape[1], a.shape[1]
    # continuous actor: train-mode BatchNorm (batch statistics), sample with the recorded Gaussian draw, gradients
    ca = _sac_load(S.C_Actor(in_dims, A), g, "c_actor/init").train()
    mean, log_std = ca(s)
    close(mean, g["c_actor/mean"])
    close(log_std, g["c_actor/log_std"])
    act, logp = ca.sample(s, torch.as_tensor(g["c_actor/eps"]).to(DEV))
    close(act, g["c_actor/sample_actions"])
    # log(1 - tanh(x)^2 + 1e-6) cancels catastrophically where |tanh| -> 1: one ulp of tanh (CPU libm vs CUDA) moves it by ~1e-4
    close(logp, g["c_actor/sample_log_prob"], rtol=3e-4)
    loss = (logp * 0.3 - act.sum(-1, keepdim=True)).mean()
    ca.zero_grad()
    loss.backward()
    # (a Linear bias in front of a BatchNorm has an exactly-zero gradient: ~1e-9 of rounding noise, compared on the network's scale)
    gscale = max(float(np.abs(g[f"c_actor/grad/{k}"]).max()) for k, _ in ca.named_parameters())
    for k, p in ca.named_parameters():
        close(p.grad, g[f"c_actor/grad/{k}"], rtol=3e-4, atol=3e-4 * gscale)                # same cancellation, through the gradient
    ca2 = _sac_load(S.C_Actor(in_dims, A), g, "c_actor/after_train_fwd").eval()
    close(ca2.evaluate(s), g["c_actor/evaluate"])
    # discrete actor
    da = _sac_load(S.D_Actor(in_dims, A), g, "d_actor/init")
    close(da(s), g["d_actor/probs"])
    assert np.array_equal(da.evaluate(s).cpu().numpy(), g["d_actor/evaluate"])
    # twin hybrid critics: outputs, loss, every parameter gradient
    q = _sac_load(S.Hybrid_Q_network(in_dims, A), g, "critic/init")
    c1, d1, c2, d2 = q(s, a)
    for k, v in (("c_q1", c1), ("d_q1", d1), ("c_q2", c2), ("d_q2", d2)):
        close(v, g[f"critic/{k}"])
    tgt = torch.as_tensor(g["critic/target"]).to(DEV)
    di = torch.as_tensor(g["critic/disc"]).to(DEV)
    closs = ((c1 - tgt).pow(2) + (c2 - tgt).pow(2) + (d1.gather(1, di) - tgt).pow(2) + (d2.gather(1, di) - tgt).pow(2)).mean()
    close(closs, g["critic/loss"])
    q.zero_grad()
    closs.backward()
    gscale = max(float(np.abs(g[f"critic/grad/{k}"]).max()) for k, _ in q.named_parameters())
    for k, p in q.named_parameters():
        close(p.grad, g[f"critic/grad/{k}"], atol=1e-5 * gscale)


def test_sac_memory_matches_reference(golden_sac):
    from rl_ctr_prediction_b200 import Hybrid_SAC_model as S
    g = golden_sac
    mem = S.Memory(10, 4, DEV, seed=1)
    for i in range(4):
        mem.add(torch.as_tensor(g[f"memory/add{i}"]))
    assert np.array_equal(mem.memory.cpu().numpy(), g["memory/after_add/memory"])
    assert np.array_equal(mem.priorities_.cpu().numpy(), g["memory/after_add/priorities"])
    mem.batch_update(torch.as_tensor(g["memory/update_idx"]), torch.as_tensor(g["memory/update_td"]))
    close(mem.priorities_, g["memory/after_update/priorities"], rtol=1e-6)
    idx, batch, isw = mem.stochastic_sample(4, sample=g["memory/sample_idx"])
    assert np.array_equal(idx.cpu().numpy(), g["memory/sample_idx"])
    assert np.array_equal(batch.cpu().numpy(), g["memory/sample_batch"])
    close(isw, g["memory/sample_isw"], rtol=1e-5)
    assert abs(mem.beta - float(g["memory/beta_after"])) < 1e-9
    # device draw: distinct indices, probability proportional to the stored priority (inclusion frequencies, successive law)
    draws, cnt = 4000, np.zeros(10)
    for _ in range(draws):
        i = mem.stochastic_sample(3)[0].cpu().numpy()
        assert len(np.unique(i)) == 3
        cnt[i] += 1
    w = mem.priorities_[:, 0].double().cpu().numpy()
    inc = _successive_inclusion(w, 3)
    assert np.abs(cnt / draws - inc).max() < 5 * np.sqrt(0.25 / draws)


def test_sac_learn_steps_run(golden_sac):
    """learn(): the reference's own learn() raises under the installed torch (recorded in the golden file), so the step is checked
    for what can be checked: it runs, losses are finite, every network and both temperatures move, the sampled priorities are
    rewritten, a fixed (noise, sample) pair gives a reproducible step."""
    from rl_ctr_prediction_b200 import Hybrid_SAC_model as S
    from rl_ctr_prediction_b200.Feature_embedding import Feature_Embedding
    assert "inplace" in str(golden_sac["meta/learn_error"])
    F_, D_, A, N = 15, 10, 3, 500

    def run(seed):
        torch.manual_seed(seed)
        agent = S.Hybrid_RL_Model(N, F_, D_, A, memory_size=64, batch_size=16, device=DEV)
        agent.memory = S.Memory(64, F_ + A + 2, DEV, seed=2)
        fe = Feature_Embedding(N, F_, D_).to(DEV)
        rs = np.random.default_rng(0)
        tr = np.concatenate([rs.integers(0, N, (40, F_)), rs.standard_normal((40, A)), rs.integers(1, A + 1, (40, 1)),
                             rs.integers(0, 2, (40, 1)) * 2 - 1], axis=1).astype(np.float32)
        agent.store_transition(torch.as_tensor(tr))
        before = {k: v.clone() for k, v in agent.Critic.state_dict().items()}
        gen = torch.Generator(device="cpu").manual_seed(7)
        losses = []
        for it in range(3):
            noise = [torch.randn(16, A, generator=gen).to(DEV), torch.randn(16, A, generator=gen).to(DEV)]
            sample = np.arange(16) + it
            losses.append(agent.learn(fe, noise=noise, sample=sample))
        return agent, before, losses

    a1, before, l1 = run(3)
    a2, _, l2 = run(3)
    assert all(np.isfinite(l1)) and l1 == l2                              # reproducible with injected draws
    for k, v in a1.Critic.state_dict().items():
        assert torch.equal(v, a2.Critic.state_dict()[k])
    assert any(not torch.equal(v, before[k]) for k, v in a1.Critic.state_dict().items())
    assert float(a1.c_log_alpha.detach()) != 0.0 and float(a1.d_log_alpha.detach()) != 0.0
    pr = a1.memory.priorities_[:, 0].cpu().numpy()
    assert (pr[:18] != 1.0).all() and (pr[18:40] == 1.0).all()            # rows 0..17 were sampled and re-prioritised


# ------------------------------------------------------------------------------------------------
# TD3 (v10) and PPO network heads against the real reference
# ------------------------------------------------------------------------------------------------
def _load_init(mod, g, prefix):
    sd = state_from_golden(g, prefix)
    assert set(mod.state_dict().keys()) == set(sd.keys()), (sorted(mod.state_dict().keys()), sorted(sd.keys()))
    mod.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    return mod.to(DEV)


def _check_grads(mod, g, prefix, rtol=1e-5):
    gscale = max(float(np.abs(g[f"{prefix}/{k}"]).max()) for k, _ in mod.named_parameters())
    for k, p in mod.named_parameters():
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        close(got, g[f"{prefix}/{k}"], rtol=rtol, atol=rtol * gscale)


def test_td3_heads_match_reference(golden_heads):
    from rl_ctr_prediction_b200 import v10_Hybrid_TD3_model_PER as T
    g = golden_heads
    s = torch.as_tensor(g["in/state"]).to(DEV)
    ca = torch.as_tensor(g["in/c_actions"]).to(DEV)
    da = torch.as_tensor(g["in/d_actions"]).to(DEV)
    in_dims, A = s.shape[1], ca.shape[1]
    cr = _load_init(T.Hybrid_Critic(in_dims, A), g, "td3_critic/init").train()
    q1, q2 = cr.evaluate(s, ca, da)
    close(q1, g["td3_critic/q1"])
    close(q2, g["td3_critic/q2"])
    tgt = torch.as_tensor(g["td3_critic/target"]).to(DEV)
    loss = (torch.nn.functional.mse_loss(q1, tgt, reduction="none") + torch.nn.functional.mse_loss(q2, tgt, reduction="none")).mean()
    close(loss, g["td3_critic/loss"])
    cr.zero_grad()
    loss.backward()
    _check_grads(cr, g, "td3_critic/grad")
    for k, v in state_from_golden(g, "td3_critic/after").items():       # BatchNorm running statistics after the train-mode forward
        close(cr.state_dict()[k].float(), v, rtol=1e-5)
    cr.eval()
    close(cr.evaluate_q_1(s, ca, da), g["td3_critic/q1_eval"])
    # actor: act() with the recorded draws, gradients through the softmax / Gumbel heads, evaluate in eval mode
    ac = _load_init(T.Hybrid_Actor(in_dims, A), g, "td3_actor/init").train()
    noise = tuple(torch.as_tensor(g[f"td3_actor/{k}"]).to(DEV) for k in ("eps_c", "eps_d", "U"))
    c_means, ens_c, d_action, ens_d = ac.act(s, 0.7, noise=noise)
    close(c_means, g["td3_actor/act/c_means"])
    close(ens_c, g["td3_actor/act/ens_c"])
    close(d_action, g["td3_actor/act/d_action"], rtol=2e-5)
    assert np.array_equal(ens_d.cpu().numpy(), g["td3_actor/act/ens_d"])
    loss = (ens_c * ca).sum(-1).mean() + (d_action * da).sum(-1).mean() + (c_means ** 2).mean()
    ac.zero_grad()
    loss.backward()
    _check_grads(ac, g, "td3_actor/grad", rtol=2e-5)
    ac.eval()
    c_e, d_e = ac.evaluate(s)
    close(c_e, g["td3_actor/eval/c"])
    close(d_e, g["td3_actor/eval/d"])
    close(T.boltzmann_softmax(c_e, 0.5), g["td3/boltzmann"])
    hard = T.gumbel_softmax_sample(d_e, temprature=0.1, hard=True)
    assert torch.equal(hard.sum(-1), torch.ones(len(s), device=DEV)) and ((hard == 0) | (hard == 1)).all()
    # the reference module exports its Memory class too
    assert T.Memory is __import__("rl_ctr_prediction_b200.replay", fromlist=["Memory"]).Memory


def test_ppo_head_matches_reference(golden_heads):
    from rl_ctr_prediction_b200 import Hybrid_PPO_model as P
    g = golden_heads
    s = torch.as_tensor(g["in/state"]).to(DEV)
    ca = torch.as_tensor(g["in/c_actions"]).to(DEV)
    d_a = torch.as_tensor(g["ppo/d_a"]).to(DEV)
    pp = _load_init(P.Hybrid_Actor_Critic(s.shape[1], ca.shape[1]), g, "ppo/init").train()
    sv, clp, cent, dlp, dent = pp.evaluate(s, ca, d_a)
    for k, v in (("state_value", sv), ("c_logprob", clp), ("c_entropy", cent), ("d_logprob", dlp), ("d_entropy", dent)):
        close(v, g[f"ppo/evaluate/{k}"])
    loss = (sv ** 2).mean() - clp.sum(-1).mean() * 0.1 - dlp.mean() - 0.01 * dent.mean()
    pp.zero_grad()
    loss.backward()
    _check_grads(pp, g, "ppo/grad", rtol=2e-5)
    pp.eval()
    pp.load_state_dict({**pp.state_dict(), **{k: torch.as_tensor(v) for k, v in state_from_golden(g, "ppo/after").items()}})
    bc, bd = pp.best_a(s)
    close(bc, g["ppo/best/c"])
    close(bd, g["ppo/best/d"])
    (c_act, c_lp, ens_c), (d_draw, d_lp, ens_d) = pp.act(s, c_noise=torch.zeros_like(ca), d_draw=d_a.view(-1))
    close(c_act, bc)                                                    # zero noise: the sample is the mean
    assert torch.equal(ens_d, d_a + 2)


def test_td3_learn_steps_match_reference(golden_td3):
    """Two full ``Hybrid_TD3_Model.learn`` steps (critic update with gradient clipping, priority update, delayed actor update,
    Polyak targets) from the reference's initial networks, replay contents and recorded random draws: critic losses, priorities
    and the final actor / critic / target networks."""
    from rl_ctr_prediction_b200 import v10_Hybrid_TD3_model_PER as T
    from rl_ctr_prediction_b200.Feature_embedding import Feature_Embedding
    g = golden_td3
    F_, D_, A, N, B = 15, 10, 3, 500, 32
    agent = T.Hybrid_TD3_Model(N, F_, D_, A, memory_size=64, batch_size=B, device=DEV)
    agent.policy_freq = 1
    for net, tgt, name in ((agent.Hybrid_Actor, agent.Hybrid_Actor_, "actor"), (agent.Hybrid_Critic, agent.Hybrid_Critic_, "critic")):
        sd = {k: torch.as_tensor(v) for k, v in state_from_golden(g, f"init/{name}").items()}
        assert set(sd.keys()) == set(net.state_dict().keys())
        net.load_state_dict(sd)
        tgt.load_state_dict(sd)
    fe = Feature_Embedding(N, F_, D_).to(DEV)
    fe.load_embedding({"feature_embedding.weight": torch.as_tensor(g["fe/feature_embedding.weight"])})
    tr = torch.as_tensor(g["transitions"])
    agent.store_transition(tr[:30])
    agent.store_transition(tr[30:])
    assert np.array_equal(agent.memory.prioritys_.cpu().numpy(), g["memory/prioritys_after_store"])
    for step in range(2):
        noise = {"eps_d": torch.as_tensor(g[f"step{step}/eps_d"]).to(DEV), "U_next": torch.as_tensor(g[f"step{step}/U_next"]).to(DEV),
                 "U_now": torch.as_tensor(g[f"step{step}/U_now"]).to(DEV), "eps_c": torch.zeros(B, A, device=DEV)}
        loss = agent.learn(fe, noise=noise, sample=g[f"step{step}/idx"])
        close(loss, g[f"step{step}/critic_loss"], rtol=2e-5)
        close(agent.memory.prioritys_, g[f"step{step}/prioritys_"], rtol=1e-4, atol=2e-5)
    for name, net in (("actor", agent.Hybrid_Actor), ("critic", agent.Hybrid_Critic), ("actor_target", agent.Hybrid_Actor_),
                      ("critic_target", agent.Hybrid_Critic_)):
        for k, v in net.state_dict().items():
            v = v.detach().float().cpu().numpy()
            full = f"final/{name}/{k}" in g.files
            ref = g[f"final/{name}/{k}"] if full else g[f"final/{name}/{k}/sub"]
            got = v if full else v.reshape(-1)[::17]
            # Adam moves an element by ~lr per step whatever its gradient's size (lr_C = 1e-2, two steps): an element whose gradient
            # is rounding noise lands anywhere within a fraction of that step.  So: (almost) every element tight, none far.
            err = np.abs(got.astype(np.float64) - ref.astype(np.float64)).reshape(-1)
            tight = 1e-4 * np.abs(ref).reshape(-1) + max(1e-4 * float(np.abs(ref).max()), 2e-5)
            assert (err > tight).mean() <= 2e-3, (name, k, float((err > tight).mean()))
            assert err.max() <= 2e-3, (name, k, float(err.max()))
            if not full:
                assert abs(v.astype(np.float64).sum() - float(g[f"final/{name}/{k}/sum"])) <= 1e-3 * max(1.0, np.abs(v).sum() ** 0.5)


def test_ppo_learn_matches_reference(golden_ppo):
    """The GAE scan against the reference's reversed Python loop, then one full ``Hybrid_PPO_Model.learn`` (3 clipped-surrogate
    epochs): returned loss and final network."""
    from rl_ctr_prediction_b200 import Hybrid_PPO_model as P
    g = golden_ppo
    adv = P.gae_advantages(torch.as_tensor(g["gae/deltas"]).to(DEV), 1 * 0.95)
    close(adv, g["gae/advantages"], rtol=1e-6)
    # a long rollout: the scan equals the sequential fp64 recurrence (c = 1 too: no decay to hide behind)
    rs = np.random.default_rng(1)
    d = rs.standard_normal(200_003).astype(np.float32)
    for c in (0.95, 1.0):
        ref, a = np.empty_like(d, dtype=np.float64), 0.0
        for i, x in enumerate(d[::-1].astype(np.float64)):
            a = c * a + x
            ref[i] = a
        got = P.gae_advantages(torch.as_tensor(d).to(DEV), c)
        close(got[:, 0], ref.astype(np.float32), rtol=1e-6, atol=1e-6 * float(np.abs(ref).max()))
    F_, D_, A = 15, 10, 3
    agent = P.Hybrid_PPO_Model(500, F_, D_, A, memory_size=128, batch_size=32, init_lr=1e-3, device=DEV)
    sd = {k: torch.as_tensor(v) for k, v in state_from_golden(g, "init").items()}
    assert set(sd.keys()) == set(agent.hybrid_actor_critic.state_dict().keys())
    agent.hybrid_actor_critic.load_state_dict(sd)
    t = lambda k: torch.as_tensor(g[f"in/{k}"]).to(DEV)
    loss = agent.learn(t("states"), t("states"), t("old_c_a"), t("old_c_lp"), t("old_d_a"), t("old_d_lp"), t("rewards"))
    close(loss, g["loss"], rtol=5e-5)
    for k, v in agent.hybrid_actor_critic.state_dict().items():
        v = v.detach().float().cpu().numpy()
        full = f"final/{k}" in g.files
        ref = g[f"final/{k}"] if full else g[f"final/{k}/sub"]
        got = v if full else v.reshape(-1)[::17]
        err = np.abs(got.astype(np.float64) - ref.astype(np.float64)).reshape(-1)
        # 3 Adam steps at lr = 1e-3 (step scale 3e-3) through a clipped / min surrogate: an element whose gradient is rounding noise
        # moves by a fraction of a step either way, so: the bulk tight, the mean error two orders below the step, nothing far
        tight = 1e-4 * np.abs(ref).reshape(-1) + max(1e-4 * float(np.abs(ref).max()), 2e-5)
        assert (err > tight).mean() <= 0.05, (k, float((err > tight).mean()))
        assert err.mean() <= 2e-5, (k, float(err.mean()))
        assert err.max() <= 1.5e-3, (k, float(err.max()))
    # the rollout memory keeps the reference's (quirky) write semantics
    agent.store_memory(torch.ones(5, F_), torch.ones(5, A), torch.ones(5, A), torch.ones(5, 1), torch.ones(5, 1), torch.ones(5, 1))
    assert float(agent.memory_state[:5].sum()) == 5 * F_ and agent.memory_counter == 0


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: REINFORCE fused into the step; acting path of the BN policy nets
# ------------------------------------------------------------------------------------------------
def _frozen_models(N, F, D):
    from rl_ctr_prediction_b200 import p_model
    torch.manual_seed(5)
    ms = {0: p_model.LR(N, device=DEV), 1: p_model.FM(N, D, device=DEV), 2: p_model.FFM(N, F, D, device=DEV)}
    for m in ms.values():
        with torch.no_grad():
            m.table.mul_(0.1)
        m.eval()
    return ms


def test_reinforce_fused_step_equals_unfused():
    """PolicyGradient.fused_step == choose actions, generate_preds, store_transition, learn() for the same actions:
    same rewards (bit-exact), same normalised returns (device fp64 scan vs the host loop), same loss and update."""
    import torch.nn as nn
    from rl_ctr_prediction_b200 import PG_model, ensemble
    N, F, D, B = 2000, 15, 10, 777
    md = _frozen_models(N, F, D)
    M = len(md)
    rng = np.random.default_rng(2)
    x = torch.as_tensor(rng.integers(0, N, size=(B, F))).to(DEV)
    y = torch.as_tensor((rng.random((B, 1)) < 0.3).astype(np.int64)).to(DEV)
    acts = torch.as_tensor(rng.integers(1, M, size=(B, 1))).to(DEV)

    def make():
        torch.manual_seed(21)
        pg = PG_model.PolicyGradient(N, F, D, action_nums=M - 1, device=DEV)
        for m in pg.policy_net.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
        return pg

    a, b = make(), make()
    loss_a, r_a, act_a = a.fused_step(x, y, md, actions=acts)
    w = torch.full((B, M), 1.0 / M, device=DEV)
    _, _, r_b = ensemble.generate_preds(md, x, acts + 1, w, y, DEV, "train")
    assert torch.equal(r_a, r_b) and torch.equal(act_a, acts)
    b.store_transition(x, acts, r_b)
    close(a.returns_on_device(r_b), b.discount_and_norm_rewards(), rtol=1e-6)
    loss_b = b.learn()
    close(loss_a, loss_b, rtol=1e-5)
    for (k, p), (_, q) in zip(a.policy_net.mlp.state_dict().items(), b.policy_net.mlp.state_dict().items()):
        assert torch.equal(p, q), k                      # same kernels, same inputs: bit-identical update
    # own action draw: in range, and the epsilon-style rule of :110-121 picks argmax where rand >= max pi
    _, r2, act2 = a.fused_step(x, y, md)
    assert int(act2.min()) >= 1 and int(act2.max()) <= M - 1 and set(r2.unique().tolist()) <= {-1.0, 1.0}


def test_reinforce_fused_step_graph_replay_equals_eager():
    """graphs.GraphedCallable over fused_step: the replayed step (given actions) leaves the same network as the eager one."""
    import torch.nn as nn
    from rl_ctr_prediction_b200 import PG_model, graphs
    N, F, D, B = 2000, 15, 10, 512
    md = _frozen_models(N, F, D)
    rng = np.random.default_rng(3)
    batches = [(torch.as_tensor(rng.integers(0, N, size=(B, F))).to(DEV),
                torch.as_tensor((rng.random((B, 1)) < 0.3).astype(np.int64)).to(DEV),
                torch.as_tensor(rng.integers(1, 3, size=(B, 1))).to(DEV)) for _ in range(5)]

    def make():
        torch.manual_seed(22)
        pg = PG_model.PolicyGradient(N, F, D, action_nums=2, device=DEV)
        for m in pg.policy_net.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
        return pg

    a, b = make(), make()
    step = graphs.GraphedCallable(lambda x, y, act: a.fused_step(x, y, md, actions=act)[0], [a.optimizer])
    for x, y, act in batches:
        la = step(x, y, act).clone()
        lb = b.fused_step(x, y, md, actions=act)[0]
        close(la, lb, rtol=1e-6)
    assert step.graph is not None
    torch.cuda.synchronize()
    for (k, p), (_, q) in zip(a.policy_net.mlp.state_dict().items(), b.policy_net.mlp.state_dict().items()):
        assert torch.equal(p, q), k
    assert a.optimizer._dense_count == b.optimizer._dense_count


def test_eval_mode_batchnorm_folded_into_gemm_and_padded_state():
    """Acting path: Linear -> BatchNorm1d(eval) -> ReLU as one GEMM (mlp.Tower folds the running statistics into the layer)
    over the 256-float-pitch state view == stock torch on the CPU."""
    import torch.nn as nn
    from rl_ctr_prediction_b200 import DDQN_model
    from rl_ctr_prediction_b200.Feature_embedding import Feature_Embedding
    from oracle import np_oracle as O
    torch.manual_seed(9)
    N, F, D, B = 3000, 15, 10, 1000
    fe = Feature_Embedding(N, F, D, device=DEV)
    x = torch.randint(0, N, (B, F), device=DEV)
    s = fe(x)
    assert s.shape == (B, 255) and s.stride(0) == 256 and s.data_ptr() % 16 == 0          # TMA-addressable rows
    ref_state = O.feature_embedding(x.cpu().numpy(), fe.state_dict()["feature_embedding.weight"].cpu().numpy())
    close(s, ref_state)
    net = DDQN_model.Net(F, N, D, 2, device=DEV)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.3)
                m.running_var.uniform_(0.5, 2.0)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
    net.eval()
    ref = torch_replica({k: v.cpu().numpy() for k, v in net.state_dict().items()}, 255, 2)
    with torch.no_grad():
        q = net(s)                                                                          # folded path
        close(q, ref.mlp(torch.as_tensor(ref_state)), rtol=2e-5)
    q_grad = net(s)                                                                         # autograd on: module-by-module path
    close(q_grad, q, rtol=2e-5)


# the main of the v10 TD3 agent scores the ensemble with its own generate_preds
@pytest.mark.parametrize("M", [3, 6])
def test_generate_preds_v10_host_api(golden_gp10, M):
    """ensemble.generate_preds_v10(model_dict, features, actions, prob_weights, c_actions, labels, device, mode): the
    reference's argument order and return values (hybrid_td3_main_per_v10.py:54-55,164) on the reference's own outputs."""
    from rl_ctr_prediction_b200 import ensemble
    g = lambda k: torch.as_tensor(golden_gp10[f"gp10/M{M}/{k}"]).to(DEV)
    pctr = g("pctr")

    class Frozen(torch.nn.Module):
        def __init__(self, col):
            super().__init__()
            self.col = col

        def forward(self, feats):
            return self.col

    md = {i: Frozen(pctr[:, i:i + 1]) for i in range(M)}
    feats = torch.zeros(pctr.shape[0], 15, dtype=torch.long, device=DEV)
    y, r, c_out = ensemble.generate_preds_v10(md, feats, g("action"), g("w"), g("c"), g("label"), DEV, mode="train")
    assert y.shape == (pctr.shape[0], 1) and r.shape == y.shape and c_out.shape == pctr.shape
    np.testing.assert_allclose(y.cpu().numpy(), golden_gp10[f"gp10/M{M}/y"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(r.cpu().numpy(), golden_gp10[f"gp10/M{M}/reward"])
    assert np.array_equal(c_out.cpu().numpy(), golden_gp10[f"gp10/M{M}/c_out"])


def test_td3_v10_main_step_is_consistent():
    """hybrid_td3_main_per_v10.py:354-410 for a few batches: random acting while batch_index < 1000 (no learning), then acting by
    the actor and one learn step.  The stored transition is [features | return_c_actions | d_q_values | d_action | reward]
    (:366-367) with return_c_actions / rewards equal to the CPU restatement on the step's own draws."""
    from oracle import np_oracle as O
    from rl_ctr_prediction_b200 import ensemble, p_model, hybrid_td3_main_per_v10 as V10, v10_Hybrid_TD3_model_PER as T
    from rl_ctr_prediction_b200.Feature_embedding import Feature_Embedding
    torch.manual_seed(5)
    N, F, D, B, M = 5000, 15, 10, 384, 3
    models = {0: p_model.LR(N), 1: p_model.FM(N, D), 2: p_model.FFM(N, F, D)}
    for m in models.values():
        with torch.no_grad():
            m.table.mul_(0.1)
        m.to(DEV).eval()
    fe = Feature_Embedding(N, F, D).to(DEV)
    fe.load_embedding(models[1].state_dict())
    agent = T.Hybrid_TD3_Model(N, F, D, M, memory_size=4 * B, batch_size=64, device=DEV)
    drawn = []
    act0 = agent.choose_action

    def recording(state, random):
        out = act0(state, random)
        drawn.append(out)
        return out

    agent.choose_action = recording
    rng = np.random.default_rng(1)
    for it, (index, learn) in enumerate(((0, None), (999, None), (1000, None))):
        x = torch.as_tensor(rng.integers(0, N, size=(B, F))).to(DEV)
        y = torch.as_tensor((rng.random((B, 1)) < 0.3).astype(np.int64)).to(DEV)
        y_preds, rewards, critic_loss = V10.train_step(agent, models, x, y, fe, index, DEV)
        assert (critic_loss is None) == (index < V10.RANDOM_STEPS)
        if critic_loss is not None:
            assert np.isfinite(float(critic_loss))
        c_a, e_c_a, d_q, d_a = drawn[-1]
        assert d_a.shape == (B, 1) and int(d_a.min()) >= 1 and int(d_a.max()) <= M
        pctr = ensemble.score_models(models, x)
        yo, ro, co, margin = O.generate_preds_v10(pctr.cpu().numpy(), e_c_a.cpu().numpy(), c_a.cpu().numpy(), d_a.cpu().numpy(),
                                                  y.cpu().numpy(), return_margin=True)
        close(y_preds, yo)
        bad = (rewards.cpu().numpy() != ro).reshape(-1)
        assert np.all(margin[bad] <= 2e-7)
        stored = agent.memory.memory[it * B:(it + 1) * B]
        assert torch.equal(stored[:, :F].long(), x)
        assert np.array_equal(stored[:, F:F + M].cpu().numpy(), co)
        assert torch.equal(stored[:, F + M:F + 2 * M], d_q.float())
        assert torch.equal(stored[:, F + 2 * M].long(), d_a.reshape(-1))
        assert torch.equal(stored[:, -1], rewards.reshape(-1))
    assert agent.memory.memory_counter == 3 * B
    yt, rt, at, wt = V10.test_batch(agent, models, x, y, fe, DEV)
    assert yt.shape == (B, 1) and set(np.unique(rt.cpu().numpy())) <= {0.0, 1.0} and wt.shape == (B, M)
