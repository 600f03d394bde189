"""RL state encoder -- drop-in for the reference's ``src/models/Feature_embedding.py:31-59``.

``Feature_Embedding(feature_numbers, field_nums, latent_dims).forward(x)`` returns the detached
``[B, F(F-1)/2 + F*D]`` state: the pairwise inner products of the gathered rows in the order
(0,1),(0,2)..(F-2,F-1) (:40-43) followed by the flattened rows (:56-57).  One kernel
(rlctr_featemb_fwd) gathers, forms the dots from shared memory and writes the state row once,
instead of the reference's two fancy-index copies of ``[B, P, D]``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .tables import Geometry, table_struct


class Feature_Embedding(nn.Module):
    def __init__(self, feature_numbers, field_nums, latent_dims, device=None):
        super().__init__()
        self.field_nums, self.latent_dims = int(field_nums), int(latent_dims)
        self._geom = Geometry.fm(feature_numbers, self.latent_dims, with_linear=False)
        g = self._geom
        dev = torch.device(device) if device is not None else None
        data = torch.zeros(g.n_rows, g.row_stride, dtype=torch.float32, device=dev)
        if dev is not None and dev.type == "cuda":
            data[:, :g.dim].normal_()
        else:
            data[:, :g.dim].copy_(torch.empty(g.n_rows, g.dim).normal_())      # nn.Embedding init (:37)
        self.table = nn.Parameter(data, requires_grad=False)                   # frozen (:59 .detach())
        self.row, self.col = [], []
        for i in range(self.field_nums - 1):
            for j in range(i + 1, self.field_nums):
                self.row.append(i), self.col.append(j)

    @property
    def output_dims(self):
        return len(self.row) + self.field_nums * self.latent_dims

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        destination[prefix + "feature_embedding.weight"] = self.table.detach()[:, :self._geom.dim].clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        k = prefix + "feature_embedding.weight"
        if k not in state_dict:
            missing_keys.append(k)
            return
        with torch.no_grad():
            self.table.data[:, :self._geom.dim].copy_(state_dict[k])

    def load_embedding(self, pretrain_params):
        """Feature_embedding.py:45-49: copy the pre-trained FM's ``feature_embedding.weight``."""
        src = pretrain_params["feature_embedding.weight"]
        with torch.no_grad():
            self.table.data[:, :self._geom.dim].copy_(torch.as_tensor(src).detach())

    @torch.no_grad()
    def forward(self, x, out=None):
        lib = _lib.load()
        if not x.is_cuda:
            raise _lib.RlctrError("Feature_Embedding runs on a CUDA (sm_100a) device only; no CPU fallback")
        x = x.long().contiguous()
        B, F = x.shape
        width = self.output_dims
        if out is None:
            # rows at a pitch of round4(width) floats (256 for the 255-wide state): the [B, width] view handed out is
            # 16-byte aligned row by row, so the first Linear of a policy net fetches it by TMA (mlp._rows_view passes
            # the pitch on) instead of falling to the software-staged GEMM; the pad column is never read
            pitch = (width + 3) // 4 * 4
            out = torch.empty(B, pitch, dtype=torch.float32, device=x.device)[:, :width]
        t = table_struct(self.table.data, self._geom)
        if not out.is_cuda or out.stride(1) != 1:
            raise _lib.RlctrError("Feature_Embedding output must be a CUDA tensor with unit column stride")
        _lib.call("rlctr_featemb_fwd", lib.rlctr_featemb_fwd, _lib.ptr(x), C.byref(t), out.data_ptr(),
                  out.stride(0) if B > 1 else width, B, F, _lib.stream(),
                  meta={"B": B, "F": F, "dim": self._geom.dim, "rs": self._geom.row_stride, "n_rows": self._geom.n_rows, "lin": False})
        return out
