#!/usr/bin/env python
"""Golden vectors for the remaining p_model tails on the same gather: WideAndDeep, FNN, InnerPNN,
OuterPNN, DCN, AFM.

    python tests/golden/make_golden_tails.py      # build container only: needs /root/reference (read-only)

Same recipe as make_golden.py (which it reuses): the REAL reference modules on small seeded inputs, CPU fp32, dropout off
(``eval()``), 3 steps of the loop body of src/main/pretrain_main.py:96-102 with dense ``torch.optim.Adam(lr=1e-3,
weight_decay=1e-5)``.  Writes ``tests/golden/ref_golden_tails.npz``; the ids / labels are those of ``ref_golden.npz``
(``train/x``, ``train/y``) so the two files describe the same batches.

Two reference quirks need a shim AROUND the unmodified reference code (nothing in /root/reference is edited):
* ``OuterPNN.__init__`` calls ``torch.ones(...).cuda()`` (p_model.py:236): ``torch.Tensor.cuda`` is patched to the identity
  while the module is constructed, so the constant kernel stays on the CPU;
* ``AFM.forward`` calls ``F.dropout(..., p=0.2)`` with the default ``training=True`` (p_model.py:477,479), stochastic even in
  ``eval()``: for the ``train/AFM/*`` trajectory ``p_model.F.dropout`` is patched to the identity ("dropout off", as ``eval()``
  gives for the other models); for ``afm_mask/*`` it is patched to multiply by explicit 0 / 1.25 masks that are stored in
  the file, so the dropout arithmetic itself is pinned too.
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
    def make_opnn():
        real = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            return P.OuterPNN(N, F, D)
        finally:
            torch.Tensor.cuda = real

    ctor = {"WideAndDeep": lambda: P.WideAndDeep(N, F, D), "FNN": lambda: P.FNN(N, F, D), "InnerPNN": lambda: P.InnerPNN(N, F, D),
            "OuterPNN": make_opnn, "DCN": lambda: P.DCN(N, F, D), "AFM": lambda: P.AFM(N, F, D)}
    real_dropout = P.F.dropout
    for name, make in ctor.items():
        torch.manual_seed(1)
        m = make()
        P.F.dropout = (lambda t, p=0.5, training=True, inplace=False: t) if name == "AFM" else real_dropout
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
    # AFM with explicit dropout masks (0 or 1/(1-p)): forward + first-step gradients
    torch.manual_seed(1)
    m = P.AFM(N, F, D)
    scale_params(m, 0.1)
    state(m, "afm_mask/init")
    npair = F * (F - 1) // 2
    gen = torch.Generator().manual_seed(11)
    B = xs[0].shape[0]
    masks = (torch.rand(B, npair + D, generator=gen) >= 0.2).float() / 0.8
    queue = [masks[:, :npair].unsqueeze(2), masks[:, npair:]]
    P.F.dropout = lambda t, p=0.5, training=True, inplace=False: t * queue.pop(0)
    p = m(xs[0].long())
    tl = torch.nn.BCELoss()(p, torch.unsqueeze(ys[0], 1).float())
    m.zero_grad()
    tl.backward()
    P.F.dropout = real_dropout
    put("afm_mask/masks", masks)
    put("afm_mask/pctr", p)
    put("afm_mask/loss", tl)
    for k, prm in m.named_parameters():
        put(f"afm_mask/grad/{k}", prm.grad)
    put("meta/torch_version", np.array(torch.__version__))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, len(G), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
