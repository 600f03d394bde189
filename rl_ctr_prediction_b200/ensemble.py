"""Ensemble scoring + reward: ``generate_preds`` of the reference's RL mains
(``src/all_main/main.py:183-271``; variant 1 = ``src/all_main/hybrid_td3_main_per.py:56-133``;
``generate_preds_v10`` = ``src/all_main/hybrid_td3_main_per_v10.py:54-164``, the main of the v10 TD3 agent).

Same signature and return values.  The M frozen CTR models write their pCTRs straight into one
``[B, M]`` buffer, and one kernel (rlctr_generate_preds, a thread per sample with an in-register
sorting network) replaces the reference's O(M^2) masked ``nonzero`` / ``index_put`` passes and
their host synchronisations.
"""
from __future__ import annotations

import torch

from . import _lib

GP_DDQN_DDPG, GP_TD3_PER = 0, 1


def score_models(model_dict, features):
    """pctr[B, M]: column m = model_dict[m](features).detach() (main.py:194-196)."""
    M = len(model_dict)
    B = len(features)
    out = torch.empty(B, M, dtype=torch.float32, device=features.device)
    with torch.no_grad():
        for m in range(M):
            out[:, m:m + 1] = model_dict[m](features)
    return out


def generate_preds(model_dict, features, actions, prob_weights, labels, device=None, mode="train",
                   variant=GP_DDQN_DDPG, pctr=None):
    lib = _lib.load()
    if pctr is None:
        pctr = score_models(model_dict, features)
    B, M = pctr.shape
    dev = pctr.device
    w = prob_weights.detach().float().contiguous()
    act = actions.reshape(-1).long().contiguous()
    lab = labels.reshape(-1).long().contiguous()
    y = torch.empty(B, 1, dtype=torch.float32, device=dev)
    w_out = torch.empty(B, M, dtype=torch.float32, device=dev)
    r = torch.empty(B, 1, dtype=torch.float32, device=dev)
    _lib.check(lib.rlctr_generate_preds(_lib.ptr(pctr.contiguous()), _lib.ptr(w), _lib.ptr(act), _lib.ptr(lab),
                                        _lib.ptr(y), _lib.ptr(w_out), _lib.ptr(r), B, M, variant, _lib.stream()),
               "rlctr_generate_preds")
    if variant == GP_TD3_PER:
        return y, r                        # hybrid_td3_main_per.py:133 returns (y_preds, rewards)
    return y, w_out, r


def generate_preds_v10(model_dict, features, actions, prob_weights, c_actions, labels, device=None, mode="train", pctr=None):
    """``hybrid_td3_main_per_v10.py:54-164``: same arguments, returns ``(y_preds, rewards, return_c_actions)``.

    Models are chosen by descending ``prob_weights``; the softmax runs over the k largest ``c_actions`` in their own order;
    rewards are 1 / 0.  ``return_c_actions`` of a partial ensemble reads the sorted ``c_actions`` of the batch row given by the
    sample's rank inside its action group -- what the reference's ``sort_c_actions[choose_model_indexs, m]`` (:117) does.
    """
    lib = _lib.load()
    if pctr is None:
        pctr = score_models(model_dict, features)
    B, M = pctr.shape
    dev = pctr.device
    w = prob_weights.detach().float().contiguous()
    c = c_actions.detach().float().contiguous()
    act = actions.reshape(-1).long().contiguous()
    lab = labels.reshape(-1).long().contiguous()
    y = torch.empty(B, 1, dtype=torch.float32, device=dev)
    c_out = torch.empty(B, M, dtype=torch.float32, device=dev)
    r = torch.empty(B, 1, dtype=torch.float32, device=dev)
    ws_bytes = lib.rlctr_generate_preds_v10_ws_bytes(B)
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.rlctr_generate_preds_v10(_lib.ptr(pctr.contiguous()), _lib.ptr(w), _lib.ptr(c), _lib.ptr(act), _lib.ptr(lab),
                                            _lib.ptr(y), _lib.ptr(c_out), _lib.ptr(r), B, M, _lib.ptr(ws), ws_bytes,
                                            _lib.stream()), "rlctr_generate_preds_v10")
    return y, r, c_out
