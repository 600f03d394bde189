"""numpy restatement of the RL_CTR_Prediction hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference ``file:line`` (relative to the reference
root) whose arithmetic it restates.  ``dtype=np.float32`` mirrors the
reference's fp32 evaluation op by op; ``dtype=np.float64`` re-evaluates the same
formula as a higher-precision arbiter for tolerance disputes.

Parity status: pinned against the reference's own modules by
``tests/golden/make_golden.py`` (the reference has no tests or golden vectors of
its own -- SURVEY.md section 4).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _as(x, dtype):
    return np.asarray(x, dtype=dtype)


# --------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------
def sigmoid(z, dtype=F32):
    """``torch.sigmoid`` as used at p_model.py:24,55,98,324: 1/(1+exp(-z)) in `dtype`.

    exp overflow to +inf is the intended fp32 behaviour (p becomes exactly 0)."""
    z = _as(z, dtype)
    with np.errstate(over="ignore"):
        return (dtype(1) / (dtype(1) + np.exp(-z))).astype(dtype)


def softmax(x, dtype=F32):
    """``torch.softmax(x, dim=1)`` (PG_model.py:56, all_main/main.py:233)."""
    x = _as(x, dtype)
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(dtype)


def gather_rows(table, ids):
    """``nn.Embedding.forward`` (p_model.py:23,47,87,303,320): pure row copy."""
    return np.asarray(table)[np.asarray(ids)]


def pair_index(field_nums):
    """Row/col pair lists, order (0,1),(0,2)..(F-2,F-1) (Feature_embedding.py:40-43)."""
    row, col = [], []
    for i in range(field_nums - 1):
        for j in range(i + 1, field_nums):
            row.append(i)
            col.append(j)
    return np.asarray(row), np.asarray(col)


# --------------------------------------------------------------------------
# CTR model logits (pre-sigmoid)
# --------------------------------------------------------------------------
def lr_logit(ids, lin, bias, dtype=F32):
    """LR.forward p_model.py:18-26: bias + sum_f lin[x_f].  lin: [N] or [N,1]."""
    w = _as(lin, dtype).reshape(-1)[np.asarray(ids)]           # [B,F]
    return (_as(bias, dtype).reshape(1, 1) + w.sum(axis=1, keepdims=True, dtype=dtype)).astype(dtype)


def fm_second_order(rows, dtype=F32):
    """0.5 * sum_d[(sum_f v)^2 - sum_f v^2]  (p_model.py:49-54, :305-311).  rows [B,F,D]."""
    rows = _as(rows, dtype)
    square_of_sum = rows.sum(axis=1, dtype=dtype) ** 2
    sum_of_square = (rows ** 2).sum(axis=1, dtype=dtype)
    ix = (square_of_sum - sum_of_square).sum(axis=1, keepdims=True, dtype=dtype)
    return (ix * dtype(0.5)).astype(dtype)


def fm_logit(ids, emb, lin, bias, dtype=F32):
    """FM.forward p_model.py:40-57 (logit only; sigmoid applied by the caller)."""
    rows = gather_rows(_as(emb, dtype), ids)
    return (lr_logit(ids, lin, bias, dtype) + fm_second_order(rows, dtype)).astype(dtype)


def ffm_logit(ids, tables, lin, bias, dtype=F32):
    """FFM.forward p_model.py:82-100.  tables: [F, N, D] (table t = field_feature_embeddings[t]).

    pair (i<j) contributes <tables[j][x_i], tables[i][x_j]>; no 0.5 factor (:97)."""
    ids = np.asarray(ids)
    tables = _as(tables, dtype)
    F = ids.shape[1]
    acc = np.zeros((ids.shape[0],), dtype=dtype)
    pairs = []
    for i in range(F - 1):
        for j in range(i + 1, F):
            pairs.append(tables[j][ids[:, i]] * tables[i][ids[:, j]])   # [B,D]
    second = np.stack(pairs, axis=1)                                    # [B,P,D]
    acc = second.sum(axis=1, dtype=dtype).sum(axis=1, keepdims=True, dtype=dtype)
    return (lr_logit(ids, lin, bias, dtype) + acc).astype(dtype)


def mlp_forward(x, layers, masks=None, keep=0.8, dtype=F32):
    """The [in->300->200->1] tower (p_model.py:276-293): Linear, ReLU, Dropout(.2) per hidden
    layer, then Linear.  `layers` = [(W[out,in], b[out]), ...]; `masks` = per-hidden-layer
    0/1 keep masks (train mode) or None (eval mode).  Returns (out, cache)."""
    h = _as(x, dtype)
    cache = []
    n = len(layers)
    for li, (W, b) in enumerate(layers):
        W = _as(W, dtype)
        b = _as(b, dtype)
        pre = (h @ W.T + b).astype(dtype)
        if li == n - 1:
            cache.append((h, W, None, None))
            h = pre
            break
        act = np.maximum(pre, dtype(0))
        mask = None
        if masks is not None:
            mask = _as(masks[li], dtype)
            out = (act * mask / dtype(keep)).astype(dtype)
        else:
            out = act
        cache.append((h, W, pre, mask))
        h = out
    return h, (cache, keep)


def mlp_backward(dout, cache_keep, dtype=F32):
    """Backward of :func:`mlp_forward`; returns (dx, [(dW, db), ...])."""
    cache, keep = cache_keep
    g = _as(dout, dtype)
    grads = [None] * len(cache)
    for li in range(len(cache) - 1, -1, -1):
        h_in, W, pre, mask = cache[li]
        if pre is not None:
            if mask is not None:
                g = (g * mask / dtype(keep)).astype(dtype)
            g = (g * (pre > 0)).astype(dtype)
        grads[li] = ((g.T @ h_in).astype(dtype), g.sum(axis=0, dtype=dtype))
        g = (g @ W).astype(dtype)
    return g, grads


def deepfm_logit(ids, emb, lin, bias, layers, masks=None, dtype=F32):
    """DeepFM.forward p_model.py:315-324: FM logit + MLP(concat_f v_f)."""
    rows = gather_rows(_as(emb, dtype), ids)
    B = rows.shape[0]
    mlp_out, cache = mlp_forward(rows.reshape(B, -1), layers, masks, dtype=dtype)
    z = (lr_logit(ids, lin, bias, dtype) + fm_second_order(rows, dtype) + mlp_out).astype(dtype)
    return z, cache


# --------------------------------------------------------------------------
# loss head: sigmoid + nn.BCELoss(mean)  (main/pretrain_main.py:167,98; p_model.py:55)
# --------------------------------------------------------------------------
def bce_loss(p, y, dtype=F32):
    """nn.BCELoss(mean) with torch's log >= -100 clamp."""
    p = _as(p, dtype).reshape(-1)
    y = _as(y, dtype).reshape(-1)
    with np.errstate(divide="ignore"):
        lp = np.maximum(np.log(p), dtype(-100))
        l1p = np.maximum(np.log1p(-p), dtype(-100))
    per = (y - dtype(1)) * l1p - y * lp
    return per.mean(dtype=dtype).astype(dtype)


def loss_head(z, y, dtype=F32):
    """p = sigmoid(z); L = BCE(p, y); dL/dz exactly as torch autograd evaluates it:
    (p - y) / max((1-p)*p, 1e-12) / B  then  * (1-p) * p   (SURVEY section 3.7 'Loss head')."""
    z = _as(z, dtype).reshape(-1)
    y = _as(y, dtype).reshape(-1)
    B = z.shape[0]
    p = sigmoid(z, dtype)
    loss = bce_loss(p, y, dtype)
    one = dtype(1)
    dp = (p - y) / np.maximum((one - p) * p, dtype(1e-12)) * (one / dtype(B))
    dz = (dp * (one - p) * p).astype(dtype)
    return p.reshape(-1, 1), loss, dz.reshape(-1, 1)


# --------------------------------------------------------------------------
# row gradients + dense scatter (aten::embedding_dense_backward)
# --------------------------------------------------------------------------
def fm_row_grads(dz, ids, emb, dtype=F32):
    """d z / d v_f = S - v_f ; d z / d w_f = 1   (SURVEY section 3.7 FM).  Returns
    (demb_rows [B,F,D], dlin_rows [B,F])."""
    rows = gather_rows(_as(emb, dtype), ids)
    S = rows.sum(axis=1, keepdims=True, dtype=dtype)
    dz = _as(dz, dtype).reshape(-1, 1, 1)
    demb = (dz * (S - rows)).astype(dtype)
    dlin = np.broadcast_to(dz.reshape(-1, 1), np.asarray(ids).shape).astype(dtype)
    return demb, dlin


def ffm_row_grads(dz, ids, tables, dtype=F32):
    """FFM: d z / d T_j[x_i] = T_i[x_j] for every i != j (p_model.py:87-97).
    Returns G [F_table, B, F_field, D]: gradient of the row of table t gathered at
    field f of sample b (zero on the diagonal t == f)."""
    ids = np.asarray(ids)
    tables = _as(tables, dtype)
    F = ids.shape[1]
    B = ids.shape[0]
    D = tables.shape[2]
    G = np.zeros((F, B, F, D), dtype=dtype)
    dzc = _as(dz, dtype).reshape(-1, 1)
    for t in range(F):
        for f in range(F):
            if t == f:
                continue
            # row tables[t][x_f] pairs with tables[f][x_t]
            G[t, :, f, :] = dzc * tables[f][ids[:, t]]
    return G


def scatter_dense(ids, rows, n_rows, dtype=F32):
    """embedding_dense_backward: dense[id] += row, sequential in slot order (b major, f minor)."""
    ids = np.asarray(ids).reshape(-1)
    rows = _as(rows, dtype).reshape(ids.shape[0], -1)
    dense = np.zeros((n_rows, rows.shape[1]), dtype=dtype)
    np.add.at(dense, ids, rows)
    return dense


# --------------------------------------------------------------------------
# torch.optim.Adam, _single_tensor_adam non-capturable branch (torch/optim/adam.py)
# used at main/pretrain_main.py:181,102  (dense, L2 folded into the gradient)
# --------------------------------------------------------------------------
def adam_schedule(step, lr, b1=0.9, b2=0.999):
    """Python-double scalars torch derives per step: (step_size, bias_correction2_sqrt)."""
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    return lr / bc1, bc2 ** 0.5


def adam_step(p, g, m, v, step, lr, wd=0.0, b1=0.9, b2=0.999, eps=1e-8, dtype=F32):
    """One dense Adam step on arrays (returns new p, m, v).  `g` may be None (== 0)."""
    p = _as(p, dtype)
    m = _as(m, dtype)
    v = _as(v, dtype)
    g = np.zeros_like(p) if g is None else _as(g, dtype)
    if wd != 0:
        g = (g + dtype(wd) * p).astype(dtype)
    m = (m + dtype(1 - b1) * (g - m)).astype(dtype)
    v = (v * dtype(b2) + dtype(1 - b2) * g * g).astype(dtype)
    step_size, bc2_sqrt = adam_schedule(step, lr, b1, b2)
    denom = (np.sqrt(v) / dtype(bc2_sqrt) + dtype(eps)).astype(dtype)
    p = (p + dtype(-step_size) * m / denom).astype(dtype)
    return p, m, v


# --------------------------------------------------------------------------
# RL state encoder (Feature_embedding.py:51-59)
# --------------------------------------------------------------------------
def feature_embedding(ids, emb, dtype=F32):
    """[B,F] ids -> [B, F(F-1)/2 + F*D]: pairwise inner products then flattened rows."""
    rows = gather_rows(_as(emb, dtype), ids)
    B, F, D = rows.shape
    r, c = pair_index(F)
    ip = (rows[:, r] * rows[:, c]).sum(axis=2, dtype=dtype)
    return np.concatenate([ip, rows.reshape(B, F * D)], axis=1).astype(dtype)


# --------------------------------------------------------------------------
# ensemble scoring + reward  (generate_preds)
# --------------------------------------------------------------------------
GP_DDQN_DDPG = 0     # src/all_main/main.py:183-271  (k in 2..M, rewards +1/-1, per-branch baseline)
GP_TD3_PER = 1       # src/all_main/hybrid_td3_main_per.py:56-133 (k in 1..M, rewards 1/0, all-model mean)


def generate_preds(pctr, w, action, label, variant=GP_DDQN_DDPG, dtype=F32, return_margin=False):
    """Per-sample ensemble prediction, returned weights and reward.

    pctr [B,M] frozen-model predictions, w [B,M] continuous action, action [B] int in
    {2..M} (variant 0) / {1..M} (variant 1), label [B] in {0,1}.
    Returns y [B,1], w_out [B,M], reward [B,1].  Ties in `w` are resolved by lower model
    index first (the reference's torch.sort is unstable -- SURVEY N12; avoid ties in tests).
    return_margin: also |y - base| [B], the distance of the reward's comparison from a tie (two correct
    fp32 evaluations may disagree on a reward only where this is at rounding level).
    """
    pctr = _as(pctr, dtype)
    w = _as(w, dtype)
    action = np.asarray(action).reshape(-1)
    label = np.asarray(label).reshape(-1)
    B, M = pctr.shape
    order = np.argsort(-w, axis=1, kind="stable")
    y = np.ones((B,), dtype=dtype)                 # all_main/main.py:185
    r = np.ones((B,), dtype=dtype)                 # :186
    w_out = np.zeros((B, M), dtype=dtype)          # :192
    margin = np.full((B,), np.inf, dtype=np.float64)
    mean_all = pctr.mean(axis=1, dtype=dtype)
    k_lo = 2 if variant == GP_DDQN_DDPG else 1
    pos, neg = (dtype(1), dtype(-1)) if variant == GP_DDQN_DDPG else (dtype(1), dtype(0))
    for k in range(k_lo, M + 1):
        sel = np.nonzero(action == k)[0]
        if sel.size == 0:
            continue
        if k == M:
            yk = (w[sel] * pctr[sel]).sum(axis=1, dtype=dtype)         # :213
            w_out[sel] = w[sel]                                       # :217
            base = mean_all[sel]                                      # :220-231
        elif k == 1:
            # hybrid_td3_main_per.py:74-85: the single best-weighted model, weight 1
            top = order[sel, 0]
            yk = pctr[sel, top]
            w_out[sel, top] = dtype(1)
            base = mean_all[sel]
        else:
            top = order[sel, :k]                                      # :201
            tw = np.take_along_axis(w[sel], top, axis=1)
            sw = softmax(tw, dtype)                                   # :233-235
            tp = np.take_along_axis(pctr[sel], top, axis=1)           # :243-249
            yk = (sw * tp).sum(axis=1, dtype=dtype)                   # :251
            wk = np.zeros((sel.size, M), dtype=dtype)
            np.put_along_axis(wk, top, sw, axis=1)                    # :238-239
            w_out[sel] = wk
            base = tp.mean(axis=1, dtype=dtype) if variant == GP_DDQN_DDPG else mean_all[sel]
        y[sel] = yk
        clk = label[sel] == 1
        good = np.where(clk, yk >= base, yk <= base)                   # :219-231,254-267
        r[sel] = np.where(good, pos, neg)
        margin[sel] = np.abs(yk.astype(np.float64) - base.astype(np.float64))
    if return_margin:
        return y.reshape(-1, 1), w_out, r.reshape(-1, 1), margin
    return y.reshape(-1, 1), w_out, r.reshape(-1, 1)


def generate_preds_v10(pctr, w, c_actions, action, label, dtype=F32, return_margin=False):
    """src/all_main/hybrid_td3_main_per_v10.py:54-164: y [B,1], reward [B,1], return_c_actions [B,M].

    Models are chosen by descending ``w`` (prob_weights, :62), the softmax runs over the k LARGEST ``c_actions`` in their own
    descending order (:63,:103-105) -- the two sorts are independent; k == M scores sum(w * pctr) and returns the row's c_actions
    (:88-96).  Rewards are 1 / 0 on strict comparisons with the mean over all M models (:127-146); actions outside 1..M keep
    the defaults y = 1, reward = 1, return_c_actions = 0 (:56-57,:73).
    As written (:117) the returned c_action of slot m of a partial ensemble is read from ``sort_c_actions`` at the row's RANK
    within its action group (a subset-relative index applied to the whole-batch tensor), not at the row itself: reproduced.
    """
    pctr = _as(pctr, dtype)
    w = _as(w, dtype)
    c = _as(c_actions, dtype)
    action = np.asarray(action).reshape(-1)
    label = np.asarray(label).reshape(-1)
    B, M = pctr.shape
    order = np.argsort(-w, axis=1, kind="stable")                   # :62
    c_desc = -np.sort(-c, axis=1, kind="stable")                    # :63 (values only)
    y = np.ones((B,), dtype=dtype)
    r = np.ones((B,), dtype=dtype)
    c_out = np.zeros((B, M), dtype=dtype)
    margin = np.full((B,), np.inf, dtype=np.float64)
    mean_all = pctr.mean(axis=1, dtype=dtype)
    for k in range(1, M + 1):
        sel = np.nonzero(action == k)[0]
        if sel.size == 0:
            continue
        if k == M:
            yk = (w[sel] * pctr[sel]).sum(axis=1, dtype=dtype)       # :89
            c_out[sel] = c[sel]                                     # :96
        else:
            sw = softmax(c_desc[sel, :k], dtype)                    # :103-105
            top = order[sel, :k]
            tp = np.take_along_axis(pctr[sel], top, axis=1)         # :110-115
            yk = (sw * tp).sum(axis=1, dtype=dtype)                 # :119
            rank = np.arange(sel.size)                              # :117: rows 0..len(sel)-1 of the WHOLE batch
            ck = np.zeros((sel.size, M), dtype=dtype)
            np.put_along_axis(ck, top, c_desc[rank, :k], axis=1)
            c_out[sel] = ck
        y[sel] = yk
        clk = label[sel] == 1
        good = np.where(clk, yk > mean_all[sel], yk < mean_all[sel])  # :127-146 (strict)
        r[sel] = np.where(good, dtype(1), dtype(0))
        margin[sel] = np.abs(yk.astype(np.float64) - mean_all[sel].astype(np.float64))
    if return_margin:
        return y.reshape(-1, 1), r.reshape(-1, 1), c_out, margin
    return y.reshape(-1, 1), r.reshape(-1, 1), c_out


# --------------------------------------------------------------------------
# REINFORCE (PG_model.py, with the N9 input-dim fix applied by the caller)
# --------------------------------------------------------------------------
RF_LITERAL = 0       # PG_model.py:104-107 as written: (sum_b -log pi_b) * mean_b(vt_b)
RF_PER_SAMPLE = 1    # textbook: mean_b(-log pi_b * vt_b)


def discount_and_norm_rewards(rs, gamma=1.0):
    """PG_model.py:139-154: float64 suffix returns over the stored episode, then (G-mean)/std."""
    rs = np.asarray(rs, dtype=np.float64).reshape(-1)
    out = np.zeros_like(rs)
    run = 0.0
    for i in range(rs.shape[0] - 1, -1, -1):
        run = run * gamma + rs[i]
        out[i] = run
    out -= out.mean()
    out /= out.std()
    return out


def reinforce_loss(logits, acts, vt, variant=RF_LITERAL, dtype=F32):
    """pi = softmax(logits); logp_b = log pi_b[a_b - 1] (PG_model.py:56,105).

    Returns (logp [B], loss scalar, dlogits [B,A])."""
    logits = _as(logits, dtype)
    acts = np.asarray(acts).reshape(-1)
    vt = _as(vt, dtype).reshape(-1)
    B, A = logits.shape
    pi = softmax(logits, dtype)
    idx = acts - 1
    pa = pi[np.arange(B), idx]
    logp = np.log(pa).astype(dtype)
    onehot = np.zeros((B, A), dtype=dtype)
    onehot[np.arange(B), idx] = 1
    if variant == RF_LITERAL:
        coef = np.full((B,), vt.mean(dtype=dtype), dtype=dtype)      # d loss / d(-logp_b)
        loss = (-logp).sum(dtype=dtype) * vt.mean(dtype=dtype)
    else:
        coef = (vt / dtype(B)).astype(dtype)
        loss = (-logp * vt).mean(dtype=dtype)
    # d(-log pi_a)/dlogits = pi - onehot
    dlogits = (coef.reshape(-1, 1) * (pi - onehot)).astype(dtype)
    return logp, dtype(loss), dlogits


# ---------------------------------------------------------------------------------------------
# SURVEY 8f.4 pieces of the RL agents that became device kernels
# ---------------------------------------------------------------------------------------------
def gae_advantages(deltas, gamma_lambda):
    """Hybrid_PPO_model.py:206-212 as written: ``adv = 0; for i, d in enumerate(reversed(deltas)): adv = c*adv + d;
    advantages[i] = adv`` with a Python-float (fp64) accumulator, stored as fp32 -- note advantages[i] belongs to sample n-1-i."""
    d = np.asarray(deltas, dtype=np.float64).reshape(-1)
    out = np.empty(d.size, dtype=np.float32)
    a = 0.0
    for i, x in enumerate(d[::-1]):
        a = gamma_lambda * a + float(np.float32(x))
        out[i] = a
    return out.reshape(-1, 1)


def per_weights(priorities, eps=1e-3, alpha=0.6):
    """Sampling weights of the prioritized memory (v10_Hybrid_TD3_model_PER.py:38-39,64-68): (|td| + eps)^alpha."""
    return (np.abs(np.asarray(priorities, dtype=np.float64)) + eps) ** alpha


def per_is_weights(p_all, idx, beta):
    """Importance-sampling weights (:80-82): (p_i / min_j p_j)^(-beta)."""
    p_all = np.asarray(p_all, dtype=np.float64)
    return (p_all[idx] / p_all.min()) ** (-beta)


def keep_top_d_actions(d_actions, c_actions, eps=None):
    """v10_Hybrid_TD3_model_PER.py:427-472 (the O(A^2) nonzero loops): sample b keeps the d_b largest continuous actions,
    d_b = argmax(d_actions[b]) + 1; kept entries become c (+ N(c, 0.2) = 2c + 0.2 eps when eps is given), the rest 0; clamp to
    [-1, 1]."""
    c = np.asarray(c_actions, dtype=np.float32)
    d = np.argmax(np.asarray(d_actions), axis=-1) + 1
    order = np.argsort(-c, axis=-1, kind="stable")
    out = np.zeros_like(c)
    for b in range(c.shape[0]):
        for m in order[b, :d[b]]:
            out[b, m] = c[b, m] if eps is None else c[b, m] + (c[b, m] + 0.2 * eps[b, m])
    return np.clip(out, -1, 1)
