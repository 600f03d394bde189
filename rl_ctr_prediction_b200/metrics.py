"""AUC / log-loss on the device (rlctr_auc_logloss).

The reference's ``test()`` / ``submission()`` move every batch's predictions to the host (``.tolist()``) and call
``sklearn.metrics.roc_auc_score`` (src/main/pretrain_main.py:110-139); here the predictions stay on the GPU and one
sort + scan + bisection pass gives the same number (ties take average ranks like sklearn's trapezoidal area)."""
from __future__ import annotations

import torch

from . import _lib


def auc_logloss(pred: torch.Tensor, labels: torch.Tensor):
    """(auc, logloss) as a device tensor of two floats; ``pred`` fp32 probabilities, ``labels`` int64 or fp32, same numel."""
    lib = _lib.load()
    p = pred.reshape(-1).contiguous().float()
    y = labels.reshape(-1).contiguous()
    if y.numel() != p.numel():
        raise ValueError("pred and labels differ in size")
    n = p.numel()
    yi = y if y.dtype == torch.int64 else None
    yf = None if yi is not None else y.float()
    out = torch.empty(2, dtype=torch.float32, device=p.device)
    ws_bytes = lib.rlctr_auc_ws_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=p.device)
    _lib.call("rlctr_auc_logloss", lib.rlctr_auc_logloss, _lib.ptr(p), _lib.ptr(yi), _lib.ptr(yf), n, _lib.ptr(out), _lib.ptr(ws),
              ws_bytes, _lib.stream(), meta={"n": n})
    return out


def roc_auc_score(labels: torch.Tensor, pred: torch.Tensor) -> float:
    """sklearn's argument order and a Python float, for drop-in use in the loop functions."""
    return float(auc_logloss(pred, labels)[0].item())
