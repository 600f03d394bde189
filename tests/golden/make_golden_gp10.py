#!/usr/bin/env python
"""Golden vectors of the v10 ensemble scoring (``src/all_main/hybrid_td3_main_per_v10.py:54-164``) from the REAL reference.

    python tests/golden/make_golden_gp10.py       # needs /root/reference (read-only); writes ref_golden_gp10.npz

The v10 ``generate_preds`` differs from ``hybrid_td3_main_per.py``: the softmax weights come from the independently sorted
``c_actions``, the chosen models from the sorted ``prob_weights``; it returns ``return_c_actions``; rewards are 1 / 0 on STRICT
comparisons with the all-model mean.  Its ``return_c_actions`` of a partial ensemble reads ``sort_c_actions`` at the row's RANK
inside its action group (a subset-relative index applied to the whole-batch tensor, :117) -- kept as it is: these vectors pin it.
"""
import importlib
import os
import sys

import numpy as np
import torch

REF = os.environ.get("RLCTR_REF_PATH", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden_gp10.npz")


class Frozen(torch.nn.Module):
    def __init__(self, col):
        super().__init__()
        self.col = col

    def forward(self, feats):
        return self.col


def main():
    V10 = importlib.import_module("src.all_main.hybrid_td3_main_per_v10")
    G = {}
    for M, Bp in ((3, 240), (5, 300), (6, 257), (4, 64)):
        g = torch.Generator().manual_seed(700 + M)
        pctr = torch.rand(Bp, M, generator=g)
        w = torch.softmax(torch.randn(Bp, M, generator=g) * 1.5, dim=1)
        c = torch.tanh(torch.randn(Bp, M, generator=g))            # the actor's continuous head (tanh range)
        lab = (torch.rand(Bp, 1, generator=g) < 0.4).long()
        act = torch.randint(1, M + 1, (Bp, 1), generator=g)
        if M == 4:
            act[:] = 2                                             # one action for the whole batch: rank == row
        feats = torch.zeros(Bp, 15, dtype=torch.long)
        md = {i: Frozen(pctr[:, i:i + 1]) for i in range(M)}
        y, r, rc = V10.generate_preds(md, feats, act, w, c, lab, torch.device("cpu"), mode="train")
        for k, v in (("pctr", pctr), ("w", w), ("c", c), ("label", lab), ("action", act), ("y", y), ("reward", r), ("c_out", rc)):
            G[f"gp10/M{M}/{k}"] = v.detach().cpu().numpy().copy()
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays")


if __name__ == "__main__":
    main()
