"""Pin the CPU oracle (numpy restatement + CPU-torch port) against vectors produced by
the real reference modules (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import state_from_golden
from oracle import np_oracle as O
from oracle import torch_port as TP

RTOL = 1e-5          # BASELINE.json north_star: logits/loss within 1e-5 relative in fp32
F, D = 15, 10


def close(a, b, rtol=RTOL, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def ffm_tables(sd, F):
    return np.stack([sd[f"field_feature_embeddings.{t}.weight"] for t in range(F)])


def tower(sd):
    return [(sd[f"mlp.{i}.weight"], sd[f"mlp.{i}.bias"]) for i in (0, 3, 6)]


def oracle_logit(name, sd, x, dtype=np.float32):
    if name == "LR":
        return O.lr_logit(x, sd["linear.weight"], sd["bias"], dtype)
    if name == "FM":
        return O.fm_logit(x, sd["feature_embedding.weight"], sd["linear.weight"], sd["bias"], dtype)
    if name == "FFM":
        return O.ffm_logit(x, ffm_tables(sd, x.shape[1]), sd["linear.weight"], sd["bias"], dtype)
    if name == "DeepFM":
        return O.deepfm_logit(x, sd["feature_embedding.weight"], sd["linear.weight"], sd["bias"], tower(sd), None, dtype)[0]
    raise KeyError(name)


# values printed in SURVEY.md section 4 (independent transcription of the same recipe)
SURVEY_KAT = {
    "LR": ([0.39326188, 0.44343323, 0.52433133, 0.38520321], 0.80406642, 8.20987225),
    "FM": ([0.43657333, 0.50530714, 0.52672166, 0.33997881], 0.83989018, 8.45803738),
    "FFM": ([0.52958852, 0.45676178, 0.50888431, 0.36476862], 0.74135745, 7.76733351),
    "DeepFM": ([0.43912792, 0.51165366, 0.52986866, 0.34588599], 0.83902109, 8.46190643),
}


@pytest.mark.parametrize("name", ["LR", "FM", "FFM", "DeepFM"])
def test_kat_matches_survey_and_oracle(golden, name):
    x, y = golden["kat/x"], golden["kat/y"]
    pctr, loss, absd = SURVEY_KAT[name]
    close(golden[f"kat/{name}/pctr"].reshape(-1), pctr, atol=2e-8, rtol=2e-7)
    close(golden[f"kat/{name}/loss"], loss, rtol=2e-7)
    close(golden[f"kat/{name}/abs_dlinear"], absd, rtol=2e-7)
    sd = state_from_golden(golden, f"kat/{name}/init")
    z = oracle_logit(name, sd, x)
    p, l, dz = O.loss_head(z, y)
    close(p, golden[f"kat/{name}/pctr"])
    close(l, golden[f"kat/{name}/loss"])
    # |dL/dlinear| summed over rows: each of the F gathered rows of sample b receives dz_b
    dense = O.scatter_dense(x, np.broadcast_to(dz, x.shape).reshape(-1, 1), sd["linear.weight"].shape[0])
    close(np.abs(dense).sum(), golden[f"kat/{name}/abs_dlinear"])


def test_hand_kat_fm_identity():
    """F=2,D=2, v1=(1,2), v2=(3,4): <v1,v2> = 11 = 0.5*sum_d[(sum v)^2 - sum v^2]."""
    emb = np.array([[1, 2], [3, 4]], dtype=np.float32)
    ids = np.array([[0, 1]])
    assert O.fm_second_order(O.gather_rows(emb, ids))[0, 0] == 11.0
    assert O.feature_embedding(ids, emb)[0, 0] == 11.0


def test_feature_embedding(golden):
    close(O.feature_embedding(golden["kat/x"], golden["kat/FE/weight"]), golden["kat/FE/out"], atol=1e-7)
    close(O.feature_embedding(golden["fe/x"], golden["fe/weight"]), golden["fe/out"], atol=1e-6)
    fe = TP.PortFeatureEmbedding(255, F, D)
    fe.load_state_dict({"feature_embedding.weight": torch.from_numpy(golden["fe/weight"])})
    assert np.array_equal(fe(torch.from_numpy(golden["fe/x"])).numpy(), golden["fe/out"])


def oracle_train(name, golden, case, steps, lr=1e-3, wd=1e-5):
    """Re-run the golden training trajectory with the numpy oracle (dense Adam + L2)."""
    sd = {k: v.copy() for k, v in state_from_golden(golden, f"{case}/{name}/init").items()}
    xs, ys = golden["train/x"], golden["train/y"]
    N = sd["linear.weight"].shape[0]
    mom = {k: (np.zeros_like(v), np.zeros_like(v)) for k, v in sd.items()}
    out = []
    for s in range(steps):
        x, y = xs[s], ys[s]
        cache = None
        if name == "DeepFM":
            z, cache = O.deepfm_logit(x, sd["feature_embedding.weight"], sd["linear.weight"], sd["bias"], tower(sd))
        else:
            z = oracle_logit(name, sd, x)
        p, loss, dz = O.loss_head(z, y)
        grads = {"bias": dz.sum(dtype=np.float32).reshape(1),
                 "linear.weight": O.scatter_dense(x, np.broadcast_to(dz, x.shape).reshape(-1, 1), N)}
        if name in ("FM", "DeepFM"):
            demb, _ = O.fm_row_grads(dz, x, sd["feature_embedding.weight"])
            if name == "DeepFM":
                dx, mg = O.mlp_backward(dz, cache)
                demb = demb + dx.reshape(demb.shape)
                for i, (dW, db) in zip((0, 3, 6), mg):
                    grads[f"mlp.{i}.weight"], grads[f"mlp.{i}.bias"] = dW, db
            grads["feature_embedding.weight"] = O.scatter_dense(x, demb, N)
        if name == "FFM":
            G = O.ffm_row_grads(dz, x, ffm_tables(sd, x.shape[1]))
            for t in range(x.shape[1]):
                grads[f"field_feature_embeddings.{t}.weight"] = O.scatter_dense(x, G[t], N)
        out.append((p, loss, {k: g.copy() for k, g in grads.items()}))
        for k in sd:
            m, v = mom[k]
            sd[k], m, v = O.adam_step(sd[k], grads[k].reshape(sd[k].shape), m, v, s + 1, lr, wd)
            mom[k] = (m, v)
    return sd, out


@pytest.mark.parametrize("name", ["LR", "FM", "FFM", "DeepFM"])
def test_training_trajectory_numpy(golden, name):
    sd, out = oracle_train(name, golden, "train", 3)
    for s, (p, loss, grads) in enumerate(out):
        close(p, golden[f"train/{name}/pctr{s}"])
        close(loss, golden[f"train/{name}/loss{s}"])
        if s == 0:
            for k, g in grads.items():
                ref = golden[f"train/{name}/grad0/{k}"]
                close(g.reshape(ref.shape), ref, atol=1e-6 * max(1e-3, float(np.abs(ref).max())))
    final = state_from_golden(golden, f"train/{name}/final")
    for k, v in final.items():
        close(sd[k].reshape(v.shape), v, atol=2e-6)


def test_saturated_regime_numpy(golden):
    """default N(0,1) init: fp32 sigmoid saturates, BCE log clamp and zero grads."""
    p0 = golden["sat/FM/pctr0"]
    assert (p0 == 0).any() or (p0 == 1).any()
    sd, out = oracle_train("FM", golden, "sat", 2)
    for s, (p, loss, grads) in enumerate(out):
        close(p, golden[f"sat/FM/pctr{s}"], rtol=2e-5, atol=1e-30)
        close(loss, golden[f"sat/FM/loss{s}"], rtol=2e-5)
    g = out[0][2]["feature_embedding.weight"]
    ref = golden["sat/FM/grad0/feature_embedding.weight"]
    assert np.array_equal(g == 0, ref == 0)           # exactly-zero rows agree
    close(g, ref, rtol=1e-4, atol=1e-6 * float(np.abs(ref).max()))
    final = state_from_golden(golden, "sat/FM/final")
    for k, v in final.items():
        close(sd[k].reshape(v.shape), v, atol=1e-5)


@pytest.mark.parametrize("name", ["LR", "FM", "FFM", "DeepFM"])
def test_training_trajectory_torch_port(golden, name):
    """The CPU-torch port that bench.py times must reproduce the reference bit-for-bit
    (same ATen ops in the same order) from the same initial state."""
    N = golden[f"train/{name}/init/linear.weight"].shape[0]
    m = TP.PortCTR(name, N, F, D)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state_from_golden(golden, f"train/{name}/init").items()})
    m.eval()
    opt = TP.make_adam(m)
    xs, ys = torch.from_numpy(golden["train/x"]), torch.from_numpy(golden["train/y"])
    for s in range(3):
        loss = TP.ctr_train_step(m, opt, torch.nn.BCELoss(), xs[s], ys[s].view(-1, 1))
        close(loss, golden[f"train/{name}/loss{s}"], rtol=1e-6)
    for k, v in state_from_golden(golden, f"train/{name}/final").items():
        close(m.state_dict()[k].numpy(), v, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["WideAndDeep", "FNN", "InnerPNN", "OuterPNN", "DCN", "AFM"])
def test_training_trajectory_torch_port_tails(golden, golden_tails, name):
    """The restated W&D / FNN / IPNN / OPNN / DCN / AFM reproduce the real reference modules' trajectories
    (AFM with its always-on dropout switched off on both sides, see tests/golden/make_golden_tails.py)."""
    init = state_from_golden(golden_tails, f"train/{name}/init")
    N = [v for k, v in init.items() if k.endswith("embedding.weight")][0].shape[0]
    m = TP.make_port(name, N, F, D)
    if name == "AFM":
        m.dropout_p = 0.0
    m.load_state_dict({k: torch.from_numpy(v) for k, v in init.items()})
    m.eval()
    opt = TP.make_adam(m)
    xs, ys = torch.from_numpy(golden["train/x"]), torch.from_numpy(golden["train/y"])
    for s in range(3):
        loss = TP.ctr_train_step(m, opt, torch.nn.BCELoss(), xs[s], ys[s].view(-1, 1))
        close(loss, golden_tails[f"train/{name}/loss{s}"], rtol=1e-6)
    for k, v in state_from_golden(golden_tails, f"train/{name}/final").items():
        close(m.state_dict()[k].numpy(), v, rtol=1e-6, atol=1e-7)
    # same seed -> same initial parameters as the reference constructors (creation order)
    torch.manual_seed(1)
    fresh = TP.make_port(name, N, F, D)
    for k, v in fresh.state_dict().items():
        scale = 0.1 if ("embedding" in k or k == "linear.weight") else 1.0
        close(v.numpy() * scale, init[k], rtol=1e-6, atol=1e-8)


def test_afm_port_with_explicit_dropout_masks(golden, golden_tails):
    """AFM's dropout arithmetic (p_model.py:477,479) with the masks as an input: forward, loss and every gradient."""
    init = state_from_golden(golden_tails, "afm_mask/init")
    m = TP.make_port("AFM", init["linear.weight"].shape[0], F, D)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in init.items()})
    x, y = torch.from_numpy(golden["train/x"][0]), torch.from_numpy(golden["train/y"][0]).view(-1, 1)
    p = m(x, masks=torch.from_numpy(golden_tails["afm_mask/masks"]))
    loss = torch.nn.BCELoss()(p, y.float())
    loss.backward()
    close(p.detach().numpy(), golden_tails["afm_mask/pctr"], rtol=1e-6)
    close(loss.item(), golden_tails["afm_mask/loss"], rtol=1e-6)
    for k, prm in m.named_parameters():
        close(prm.grad.numpy(), golden_tails[f"afm_mask/grad/{k}"], rtol=1e-6, atol=1e-9)


def test_port_init_matches_reference_seed(golden):
    """Same torch.manual_seed -> same initial parameters as the reference constructors."""
    for name in ("LR", "FM", "FFM", "DeepFM"):
        torch.manual_seed(1)
        m = TP.PortCTR(name, 64, F, D)
        ref = state_from_golden(golden, f"kat/{name}/init")
        for k, v in m.state_dict().items():
            scale = 0.1 if ("embedding" in k or k == "linear.weight") else 1.0
            close(v.numpy() * scale, ref[k], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("M", [3, 5, 6])
def test_generate_preds(golden, M):
    g = lambda k: golden[f"gp/v0_M{M}/{k}"]
    y, w_out, r = O.generate_preds(g("pctr"), g("w"), g("action"), g("label"), O.GP_DDQN_DDPG)
    close(y, g("y"), atol=1e-7)
    close(w_out, g("w_out"), atol=1e-7)
    assert np.array_equal(r, g("reward"))
    y1, _, r1 = O.generate_preds(g("pctr"), g("w"), golden[f"gp/v1_M{M}/action"], g("label"), O.GP_TD3_PER)
    close(y1, golden[f"gp/v1_M{M}/y"], atol=1e-7)
    assert np.array_equal(r1, golden[f"gp/v1_M{M}/reward"])


@pytest.mark.parametrize("M", [3, 5, 6, 4])
def test_generate_preds_v10(golden_gp10, M):
    """hybrid_td3_main_per_v10.py:54-164, incl. the rank-indexed return_c_actions (:117); M = 4 is the single-action batch."""
    g = lambda k: golden_gp10[f"gp10/M{M}/{k}"]
    y, r, c_out = O.generate_preds_v10(g("pctr"), g("w"), g("c"), g("action"), g("label"))
    close(y, g("y"), atol=1.5e-7)
    assert np.array_equal(r, g("reward"))
    assert np.array_equal(c_out, g("c_out"))
    if M != 4:                                   # the rank quirk is visible: some returned value is not from the row's own c_actions
        own = np.sort(g("c"), axis=1)
        got = np.sort(np.where(c_out != 0, c_out, np.inf), axis=1)
        partial = g("action").reshape(-1) < M
        foreign = [(~np.isin(got[i][np.isfinite(got[i])], own[i])).any() for i in np.nonzero(partial)[0]]
        assert any(foreign)


def test_reinforce(golden):
    vt = O.discount_and_norm_rewards(golden["pg/rs"], 1.0)
    close(vt, golden["pg/vt_norm"], rtol=1e-12, atol=1e-12)
    logp, loss, dl = O.reinforce_loss(golden["pg/logits"], golden["pg/acts"], vt.astype(np.float32), O.RF_LITERAL)
    close(logp, golden["pg/logp"])
    # literal loss is (sum -logp) * mean(vt) with mean(vt) ~ 0: compare on the scale of sum|logp|
    scale = float(np.abs(golden["pg/logp"]).sum())
    close(loss, golden["pg/loss_literal"], atol=1e-5 * scale * 1e-6 + 1e-6)
    logp, loss, dl = O.reinforce_loss(golden["pg/logits"], golden["pg/acts"], golden["pg/vt_raw"], O.RF_LITERAL)
    close(loss, golden["pg/loss_literal_raw"], rtol=1e-5)
    close(dl, golden["pg/dlogits_literal_raw"], rtol=1e-5, atol=1e-7)


def test_gae_oracle_matches_reference_loop():
    """oracle gae_advantages == the advantages the reference's Python loop produced (tests/golden/make_golden_ppo.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden_ppo.npz"), allow_pickle=False)
    close(O.gae_advantages(g["gae/deltas"], 1 * 0.95), g["gae/advantages"], rtol=1e-6)


def test_per_weights_oracle_matches_reference_memory():
    """oracle IS weights == the reference Memory.stochastic_sample output for the recorded indices (SAC flavour: the stored
    priority is the weight; tests/golden/make_golden_sac.py)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden_sac.npz"), allow_pickle=False)
    pr = g["memory/after_update/priorities"][:, 0]
    close(O.per_is_weights(pr, g["memory/sample_idx"], 0.4 + 1e-5), g["memory/sample_isw"][:, 0], rtol=1e-5)
    # and the priorities batch_update wrote are (|td| + eps)^alpha
    close(pr[g["memory/update_idx"]], O.per_weights(g["memory/update_td"][:, 0]), rtol=1e-6)
