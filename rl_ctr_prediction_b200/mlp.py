"""Dense layers of the DeepFM tower and the policy networks (p_model.py:276-293; PG_model.py:42-51;
DDQN_model.py:20-52; DDPG_for_PG_model.py:20-81) on the tcgen05 tensor cores.

``Linear`` keeps ``nn.Linear``'s parameters, init and state_dict keys (``weight [out,in]``,
``bias [out]``), so reference checkpoints load unchanged and the same ``torch.manual_seed`` gives the
reference's initial weights.  Forward and backward are the 3xTF32 GEMM kernels of ``csrc/mlp.cu``
(rlctr_linear_fwd / rlctr_linear_bwd): fp32-grade accuracy (the reference runs fp32 SGEMM,
``allow_tf32=False``), no cuBLAS on the path.  ``Tower`` is an ``nn.Sequential`` that fuses each
``Linear -> ReLU`` pair into the GEMM epilogue; ``nn.Dropout`` stays a torch op between kernels so
train-mode masks are torch's Philox stream.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


def _rows_view(x):
    """(x as [rows, K] fp32 CUDA tensor, floats between rows).  A row-padded view (the tower input DeepFM gathers into a
    [B, round4(F*D)] buffer so that TMA can address it) is passed through with its pitch; anything else is made dense."""
    if not x.is_cuda:
        raise _lib.RlctrError("rl_ctr_prediction_b200.mlp.Linear runs on a CUDA (sm_100a) device only")
    K = x.shape[-1]
    if x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1 and x.stride(0) >= K and x.shape[0] > 1:
        return x, x.stride(0)
    x2 = x.reshape(-1, K).contiguous().float()
    return x2, K


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        lib = _lib.load()
        x2, ldx = _rows_view(x)
        B, K = x2.shape
        N = weight.shape[0]
        y = torch.empty(B, N, dtype=torch.float32, device=x.device)
        flags = _lib.RLCTR_MLP_RELU if relu else 0
        w = weight.detach().contiguous()
        ws_bytes = lib.rlctr_mlp_ws_bytes(B, K, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        _lib.call("rlctr_linear_fwd", lib.rlctr_linear_fwd, x2.data_ptr(), ldx, _lib.ptr(w),
                  _lib.ptr(bias.detach() if bias is not None else None), _lib.ptr(y), B, K, N, flags, _lib.ptr(ws), ws_bytes,
                  _lib.stream(), key=f"rlctr_linear_fwd[{K}x{N}]", meta={"B": B, "K": K, "N": N})
        ctx.relu = relu
        ctx.ldx = ldx
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x2, w, y if relu else None)
        ctx.in_shape = x.shape
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x2, w, y = ctx.saved_tensors
        B, K = x2.shape
        N = w.shape[0]
        gy2 = gy.reshape(B, N).contiguous().float()
        if ctx.relu:
            gy2 = gy2.clone() if gy2.data_ptr() == gy.data_ptr() else gy2    # masked in place by the kernel
        need_dx, need_dw, need_db = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dev = x2.device
        dx = torch.empty(B, K, dtype=torch.float32, device=dev) if need_dx else None
        dw = torch.empty(N, K, dtype=torch.float32, device=dev) if need_dw else None
        db = torch.empty(N, dtype=torch.float32, device=dev) if need_db else None
        ws_bytes = lib.rlctr_mlp_ws_bytes(B, K, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        flags = _lib.RLCTR_MLP_RELU if ctx.relu else 0
        _lib.call("rlctr_linear_bwd", lib.rlctr_linear_bwd, x2.data_ptr(), ctx.ldx, _lib.ptr(w), _lib.ptr(y), _lib.ptr(gy2), _lib.ptr(dx),
                  _lib.ptr(dw), _lib.ptr(db), B, K, N, flags, _lib.ptr(ws), ws_bytes, _lib.stream(),
                  key=f"rlctr_linear_bwd[{K}x{N}]", meta={"B": B, "K": K, "N": N, "dx": need_dx})
        if dx is not None:
            dx = dx.reshape(ctx.in_shape)
        return dx, dw, db, None


class Linear(nn.Linear):
    """y = x W^T + b in fp32-grade 3xTF32 on tcgen05 (rlctr_linear_fwd/bwd)."""

    def forward(self, x, relu: bool = False):
        return _LinearFn.apply(x, self.weight, self.bias, relu)


class Tower(nn.Sequential):
    """``nn.Sequential`` whose ``Linear -> nn.ReLU`` pairs run as one GEMM with a fused ReLU epilogue."""

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, Linear) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                x = m(x, relu=True)
                i += 2
            else:
                x = m(x)
                i += 1
        return x
