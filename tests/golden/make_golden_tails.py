#!/usr/bin/env python
"""Golden vectors for the remaining p_model tails on the same gather: WideAndDeep, FNN, InnerPNN.

    python tests/golden/make_golden_tails.py      # build container only: needs /root/reference (read-only)

Same recipe as make_golden.py (which it reuses): the REAL reference modules on small seeded inputs, CPU fp32, dropout off
(``eval()``), 3 steps of the loop body of src/main/pretrain_main.py:96-102 with dense ``torch.optim.Adam(lr=1e-3,
weight_decay=1e-5)``.  Writes ``tests/golden/ref_golden_tails.npz``; the ids / labels are those of ``ref_golden.npz``
(``train/x``, ``train/y``) so the two files describe the same batches.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RLCTR_REF_PATH", "/root/reference")
sys.path.insert(0, REF)
OUT = os.path.join(HERE, "ref_golden_tails.npz")
G = {}


def put(key, val):
    if isinstance(val, torch.Tensor):
        val = val.detach().cpu().numpy()
    G[key] = np.array(val, copy=True)


def scale_params(model, s):
    with torch.no_grad():
        for k, p in model.named_parameters():
            if "embedding" in k or k == "linear.weight":
                p.mul_(s)


def state(model, prefix):
    for k, v in model.state_dict().items():
        put(f"{prefix}/{k}", v)


def main():
    torch.set_num_threads(1)
    P = importlib.import_module("src.models.p_model")
    base = np.load(os.path.join(HERE, "ref_golden.npz"), allow_pickle=False)
    N, F, D = 255, 15, 10
    xs = [torch.from_numpy(a) for a in base["train/x"]]
    ys = [torch.from_numpy(a) for a in base["train/y"]]
    ctor = {"WideAndDeep": lambda: P.WideAndDeep(N, F, D), "FNN": lambda: P.FNN(N, F, D), "InnerPNN": lambda: P.InnerPNN(N, F, D)}
    for name, make in ctor.items():
        torch.manual_seed(1)
        m = make()
        scale_params(m, 0.1)
        m.eval()
        state(m, f"train/{name}/init")
        opt = torch.optim.Adam(params=m.parameters(), lr=1e-3, weight_decay=1e-5)
        lossf = torch.nn.BCELoss()
        for s in range(len(xs)):
            feats, labels = xs[s].long(), torch.unsqueeze(ys[s], 1)
            p = m(feats)
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
    put("meta/torch_version", np.array(torch.__version__))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
