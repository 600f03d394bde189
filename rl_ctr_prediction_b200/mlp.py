"""Dense layers of the DeepFM tower and the policy networks (p_model.py:276-293;
DDQN_model.py:20-52; DDPG_for_PG_model.py:20-81; PG_model.py:24-58).

``Linear`` keeps ``nn.Linear``'s parameters, init and state_dict keys (``weight [out,in]``,
``bias [out]``) so reference checkpoints load unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class Linear(nn.Linear):
    """y = x W^T + b in fp32 (torch default allow_tf32=False semantics: the reference runs fp32 SGEMM)."""

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("rl_ctr_prediction_b200.mlp.Linear runs on a CUDA (sm_100a) device only")
        return F.linear(x, self.weight, self.bias)
