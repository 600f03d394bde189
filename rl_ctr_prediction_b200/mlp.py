"""Dense layers of the DeepFM tower and the policy networks (p_model.py:276-293; PG_model.py:42-51;
DDQN_model.py:20-52; DDPG_for_PG_model.py:20-81) on the tcgen05 tensor cores.

``Linear`` keeps ``nn.Linear``'s parameters, init and state_dict keys (``weight [out,in]``,
``bias [out]``), so reference checkpoints load unchanged and the same ``torch.manual_seed`` gives the
reference's initial weights.  Forward and backward are the 3xTF32 GEMM kernels of ``csrc/mlp_tma.cu`` /
``csrc/mlp.cu`` (rlctr_linear_fwd / rlctr_linear_bwd): fp32-grade accuracy (the reference runs fp32 SGEMM,
``allow_tf32=False``), no cuBLAS on the path.

``Tower`` is an ``nn.Sequential`` of ``Linear [-> ReLU] [-> Dropout]`` groups that runs as ONE autograd node:
every group is a single GEMM whose epilogue applies bias, ReLU and (in train mode) the dropout mask, and in the
backward the ReLU/dropout mask of layer i is applied inside the dgrad epilogue of layer i+1 -- no elementwise kernel
touches an activation.  The mask is a counter-based hash of (seed, element index) kept in device memory
(include/rlctr.h, rlctr_rng_advance): Bernoulli(1-p) scaled by 1/(1-p) exactly like ``nn.Dropout``, drawn from its
own stream (the reference's CPU Philox stream cannot be reproduced on a GPU either -- SURVEY N6); ``eval()`` is the
identity as in the reference.  A Sequential with any other module inside (BatchNorm in the DDQN/DDPG nets) runs
module by module with ``Linear -> ReLU`` pairs fused.
"""
from __future__ import annotations

import torch
import torch.nn as nn

import os

from . import _lib

# Batches of at most this many rows take the exact-fp32 CUDA-core GEMM (RLCTR_MLP_FP32) instead of 3xTF32 on the tensor cores:
# the learn steps of the policy nets run on replay batches of 32-256 rows, where tensor cores buy nothing and the BatchNorm
# backward needs the reference's fp32 SGEMM accuracy (csrc/mlp.cu `simt`).  0 disables it.
FP32_MAX_BATCH = int(os.environ.get("RLCTR_FP32_MAX_BATCH", "1024"))


# The dgrad of a tower layer reuses the (W_hi, W_lo) images its forward left in the workspace (RLCTR_MLP_W_PRESPLIT): one launch
# less per layer and step.  0 = split again (A/B switch; tests compare the two bit for bit).
REUSE_SPLIT = os.environ.get("RLCTR_MLP_REUSE_SPLIT", "1") != "0"


def _fp32_flag(B):
    return _lib.RLCTR_MLP_FP32 if B <= FP32_MAX_BATCH else 0


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


def _fwd(lib, x2, ldx, w, bias, relu, drop_p=0.0, rng=None, keep_ws=False):
    """keep_ws: also return the call's workspace -- it holds the (W_hi, W_lo) images the layer's dgrad can reuse (RLCTR_MLP_W_PRESPLIT)."""
    B, K = x2.shape
    N = w.shape[0]
    y = torch.empty(B, N, dtype=torch.float32, device=x2.device)
    flags = (_lib.RLCTR_MLP_RELU if relu else 0) | (_lib.RLCTR_MLP_DROPOUT if drop_p > 0.0 else 0) | _fp32_flag(B)
    ws_bytes = lib.rlctr_mlp_ws_bytes(B, K, N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x2.device)
    _lib.call("rlctr_linear_fwd", lib.rlctr_linear_fwd, x2.data_ptr(), ldx, _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), B, K, N,
              flags, float(drop_p), _lib.ptr(rng) if drop_p > 0.0 else None, _lib.ptr(ws), ws_bytes, _lib.stream(),
              key=f"rlctr_linear_fwd[{K}x{N}]", meta={"B": B, "K": K, "N": N})
    if drop_p > 0.0:
        _lib.check(lib.rlctr_rng_advance(_lib.ptr(rng), B * N, _lib.stream()), "rlctr_rng_advance")
    if keep_ws:
        return y, (ws if (N > 1 and not (flags & _lib.RLCTR_MLP_FP32)) else None)
    return y


def _bwd(lib, x2, ldx, w, y, gy2, need_dx, need_dw, need_db, relu, gy_scale=1.0, dx_mask=False, dx_scale=1.0, fwd_ws=None):
    """fwd_ws: the workspace of this layer's forward call (weights unchanged since): its weight split is reused."""
    B, K = x2.shape
    N = w.shape[0]
    dev = x2.device
    dx = torch.empty(B, K, dtype=torch.float32, device=dev) if need_dx else None
    dw = torch.empty(N, K, dtype=torch.float32, device=dev) if need_dw else None
    db = torch.empty(N, dtype=torch.float32, device=dev) if need_db else None
    ws_bytes = lib.rlctr_mlp_ws_bytes(B, K, N)
    reuse = REUSE_SPLIT and fwd_ws is not None and fwd_ws.numel() >= ws_bytes and not _fp32_flag(B)
    ws = fwd_ws if reuse else torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    flags = ((_lib.RLCTR_MLP_RELU if relu else 0) | (_lib.RLCTR_MLP_DX_MASK if dx_mask else 0) | _fp32_flag(B) |
             (_lib.RLCTR_MLP_W_PRESPLIT if reuse else 0))
    _lib.call("rlctr_linear_bwd", lib.rlctr_linear_bwd, x2.data_ptr(), ldx, _lib.ptr(w), _lib.ptr(y) if relu else None,
              _lib.ptr(gy2), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), B, K, N, flags, float(gy_scale), float(dx_scale),
              _lib.ptr(ws), ws_bytes, _lib.stream(), key=f"rlctr_linear_bwd[{K}x{N}]",
              meta={"B": B, "K": K, "N": N, "dx": need_dx})
    return dx, dw, db


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        lib = _lib.load()
        x2, ldx = _rows_view(x)
        N = weight.shape[0]
        w = weight.detach().contiguous()
        y = _fwd(lib, x2, ldx, w, bias.detach() if bias is not None else None, relu)
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
        dx, dw, db = _bwd(lib, x2, ctx.ldx, w, y, gy2, need_dx, need_dw, need_db, ctx.relu)
        if dx is not None:
            dx = dx.reshape(ctx.in_shape)
        return dx, dw, db, None


class Linear(nn.Linear):
    """y = x W^T + b in fp32-grade 3xTF32 on tcgen05 (rlctr_linear_fwd/bwd)."""

    def forward(self, x, relu: bool = False):
        return _LinearFn.apply(x, self.weight, self.bias, relu)


BN_FUSED_MAX_BATCH = 1 << 16


class _BnReluFn(torch.autograd.Function):
    """nn.BatchNorm1d (training mode) [+ nn.ReLU] as ONE library call each way: rlctr_bn_relu_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu):
        lib = _lib.load()
        x2, ldx = _rows_view(x)
        B, N = x2.shape
        dev = x2.device
        y = torch.empty(B, N, dtype=torch.float32, device=dev)
        mean = torch.empty(N, dtype=torch.float32, device=dev)
        invstd = torch.empty(N, dtype=torch.float32, device=dev)
        _lib.call("rlctr_bn_relu_fwd", lib.rlctr_bn_relu_fwd, x2.data_ptr(), ldx, _lib.ptr(weight.detach() if weight is not None else None),
                  _lib.ptr(bias.detach() if bias is not None else None), _lib.ptr(running_mean), _lib.ptr(running_var),
                  float(momentum), float(eps), _lib.ptr(y), N, _lib.ptr(mean), _lib.ptr(invstd), B, N, 1 if relu else 0,
                  _lib.stream(), meta={"B": B, "N": N})
        ctx.save_for_backward(x2, y, weight, mean, invstd)
        ctx.ldx, ctx.relu, ctx.in_shape, ctx.has_bias = ldx, relu, x.shape, bias is not None
        return y.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        x2, y, weight, mean, invstd = ctx.saved_tensors
        B, N = y.shape
        dev = y.device
        g2 = gy.reshape(B, N).contiguous().float()
        dx = torch.empty(B, N, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
        dgamma = torch.empty(N, dtype=torch.float32, device=dev) if (weight is not None and ctx.needs_input_grad[1]) else None
        dbeta = torch.empty(N, dtype=torch.float32, device=dev) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        _lib.call("rlctr_bn_relu_bwd", lib.rlctr_bn_relu_bwd, x2.data_ptr(), ctx.ldx, _lib.ptr(y), N, _lib.ptr(g2), N,
                  _lib.ptr(weight.detach() if weight is not None else None), _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(dx), N,
                  _lib.ptr(dgamma), _lib.ptr(dbeta), B, N, 1 if ctx.relu else 0, _lib.stream(), meta={"B": B, "N": N})
        if dx is not None:
            dx = dx.reshape(ctx.in_shape)
        return dx, dgamma, dbeta, None, None, None, None, None


def bn_relu(bn: nn.BatchNorm1d, x, relu: bool):
    """``relu(bn(x))`` for a training-mode ``nn.BatchNorm1d`` on the library kernels (same running-statistics bookkeeping as
    ``nn.BatchNorm1d.forward``: momentum None = cumulative average, num_batches_tracked)."""
    momentum = 0.0 if bn.momentum is None else bn.momentum
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if bn.momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    rm, rv = (bn.running_mean, bn.running_var) if bn.track_running_stats else (None, None)
    return _BnReluFn.apply(x, bn.weight if bn.affine else None, bn.bias if bn.affine else None, rm, rv, momentum, bn.eps, relu)


def apply_bn(bn, x):
    """``bn(x)``: on the library kernel when ``bn`` is a plain training-mode ``nn.BatchNorm1d`` on a CUDA batch, else the module
    itself (eval mode, cross-rank statistics, batches beyond the kernel's range)."""
    if (type(bn) is nn.BatchNorm1d and bn.training and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32
            and 1 < x.shape[0] <= BN_FUSED_MAX_BATCH):
        return bn_relu(bn, x, False)
    return bn(x)


class _TowerFn(torch.autograd.Function):
    """All layers of a Linear[-ReLU][-Dropout] stack as one autograd node (see the module docstring)."""

    @staticmethod
    def forward(ctx, x, spec, rng, *params):
        # spec: tuple of (relu, drop_p, has_bias) per layer; params: weight_0, bias_0, weight_1, bias_1, ... (bias may be None)
        lib = _lib.load()
        h, ld = _rows_view(x)
        acts, lds, ws_, fwd_ws = [], [], [], []
        for i, (relu, drop_p, has_bias) in enumerate(spec):
            w = params[2 * i].detach().contiguous()
            b = params[2 * i + 1].detach() if has_bias else None
            acts.append(h)
            lds.append(ld)
            ws_.append(w)
            h, fws = _fwd(lib, h, ld, w, b, relu, drop_p, rng, keep_ws=True)
            fwd_ws.append(fws)
            ld = h.shape[1]
        ctx.spec, ctx.lds = spec, lds
        ctx.fwd_ws = fwd_ws                                   # (W_hi, W_lo) of every layer: the dgrads reuse them
        ctx.in_shape = x.shape
        ctx.n_layers = len(spec)
        last_relu = spec[-1][0]
        ctx.save_for_backward(*acts, *ws_, *( [h] if last_relu else [] ))
        return h.reshape(*x.shape[:-1], h.shape[1])

    @staticmethod
    def backward(ctx, gy):
        lib = _lib.load()
        L = ctx.n_layers
        saved = ctx.saved_tensors
        acts, ws_ = saved[:L], saved[L:2 * L]
        spec = ctx.spec
        B = acts[0].shape[0]
        g = gy.reshape(B, -1).contiguous().float()
        grads = [None] * (2 * L)
        for i in range(L - 1, -1, -1):
            relu, drop_p, has_bias = spec[i]
            own_mask = relu and i == L - 1          # only the top layer masks its own incoming gradient
            if own_mask and g.data_ptr() == gy.data_ptr():
                g = g.clone()
            below = spec[i - 1] if i > 0 else None
            dx_mask = below is not None and below[0]                     # the layer below ended in ReLU (+ dropout)
            dx_scale = 1.0 / (1.0 - below[1]) if (below is not None and below[1] > 0.0) else 1.0
            if below is not None and not below[0] and below[1] > 0.0:
                raise _lib.RlctrError("Dropout without a preceding ReLU is not fused; use separate modules")
            need_dx = i > 0 or ctx.needs_input_grad[0]
            need_dw = ctx.needs_input_grad[3 + 2 * i]
            need_db = has_bias and ctx.needs_input_grad[3 + 2 * i + 1]
            gy_scale = 1.0 / (1.0 - drop_p) if (own_mask and drop_p > 0.0) else 1.0
            dx, dw, db = _bwd(lib, acts[i], ctx.lds[i], ws_[i], saved[2 * L] if own_mask else None, g, need_dx, need_dw,
                              need_db, own_mask, gy_scale, dx_mask, dx_scale, fwd_ws=ctx.fwd_ws[i])
            ctx.fwd_ws[i] = None
            grads[2 * i], grads[2 * i + 1] = dw, db
            g = dx
        dx0 = g.reshape(ctx.in_shape) if (g is not None and ctx.needs_input_grad[0]) else None
        return (dx0, None, None, *grads)


class Tower(nn.Sequential):
    """``nn.Sequential`` of Linear / ReLU / Dropout (reference state_dict indices kept: Linear layers at 0, 3, 6 ...)."""

    def _groups(self):
        """[(Linear, relu, dropout module or None)] if the Sequential is made only of Linear[-ReLU][-Dropout] groups."""
        mods, i, out = list(self), 0, []
        while i < len(mods):
            if not isinstance(mods[i], Linear):
                return None
            lin, relu, drop = mods[i], False, None
            i += 1
            if i < len(mods) and isinstance(mods[i], nn.ReLU):
                relu = True
                i += 1
            if i < len(mods) and isinstance(mods[i], nn.Dropout):
                if not relu:
                    return None
                drop = mods[i]
                i += 1
            out.append((lin, relu, drop))
        return out

    def _rng_state(self, device):
        st = getattr(self, "_rlctr_rng", None)
        if st is None or st.device != device:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())          # CPU default generator: follows torch.manual_seed
            st = torch.tensor([seed, 0], dtype=torch.int64, device=device)
            self._rlctr_rng = st
        return st

    def forward(self, x):
        groups = self._groups()
        if groups is None or not x.is_cuda:
            return self._forward_modules(x)
        spec, params, any_drop = [], [], False
        for lin, relu, drop in groups:
            p = float(drop.p) if (drop is not None and drop.training and drop.p > 0.0) else 0.0
            any_drop = any_drop or p > 0.0
            spec.append((relu, p, lin.bias is not None))
            params += [lin.weight, lin.bias]
        rng = self._rng_state(x.device) if any_drop else None
        return _TowerFn.apply(x, tuple(spec), rng, *params)

    def _forward_modules(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if (isinstance(m, Linear) and isinstance(nxt, nn.BatchNorm1d) and not nxt.training and nxt.track_running_stats
                    and not torch.is_grad_enabled() and x.is_cuda):
                # acting path (eval mode, no autograd; DDQN_model.py:148-151, DDPG_for_PG_model.py:178-181): BatchNorm with
                # running statistics is a per-column affine map, folded into the layer -- W' = W * s, b' = (b - mean) * s + beta,
                # s = gamma / sqrt(var + eps) -- so Linear -> BN -> ReLU is ONE GEMM with a fused epilogue instead of a GEMM and
                # two elementwise passes over [B, 300] (at B = 1M: 2.4 GB of traffic per layer)
                scale = nxt.weight / torch.sqrt(nxt.running_var + nxt.eps) if nxt.affine else torch.rsqrt(nxt.running_var + nxt.eps)
                bias0 = m.bias if m.bias is not None else torch.zeros_like(nxt.running_mean)
                shift = (bias0 - nxt.running_mean) * scale + (nxt.bias if nxt.affine else 0.0)
                relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                x2, ldx = _rows_view(x)
                y = _fwd(_lib.load(), x2, ldx, (m.weight * scale[:, None]).contiguous(), shift.contiguous(), relu)
                x = y.reshape(*x.shape[:-1], y.shape[1])
                i += 3 if relu else 2
                continue
            if (type(m) is nn.BatchNorm1d and m.training and x.is_cuda and x.dim() == 2 and x.dtype == torch.float32
                    and 1 < x.shape[0] <= BN_FUSED_MAX_BATCH):
                # learn steps (training mode): batch statistics, normalisation, affine map and the ReLU behind it in one kernel
                # (and one for their backward) instead of native_batch_norm + relu (+ three backward kernels)
                relu = isinstance(nxt, nn.ReLU)
                x = bn_relu(m, x, relu)
                i += 2 if relu else 1
                continue
            if isinstance(m, Linear) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                x = m(x, relu=True)
                i += 2
            else:
                x = m(x)
                i += 1
        return x
