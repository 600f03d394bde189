"""Network heads and replay memory of the reference's latest hybrid TD3 agent (``src/models/v10_Hybrid_TD3_model_PER.py``;
SURVEY 8f.4) on the B200 path: ``Memory`` (:19-110, = :class:`rl_ctr_prediction_b200.replay.Memory`: device-side prioritized
sampling), ``Hybrid_Critic`` (:119-189), ``Hybrid_Actor`` (:191-256), ``gumbel_softmax_sample`` / ``boltzmann_softmax`` /
``onehot_from_logits`` (:258-286).  Same class names, constructor arguments, state_dict keys and method signatures; every
``nn.Linear`` is :class:`rl_ctr_prediction_b200.mlp.Linear` (tcgen05 3xTF32 GEMMs; ``Linear -> ReLU`` pairs fused by
``mlp.Tower``).  The reference hard-codes ``.cuda()`` in ``gumbel_softmax_sample`` (:262); here the noise lives on the input's
device.  Random draws (the two Gaussian perturbations of ``act`` and the uniform of the Gumbel sample) can be injected for
reproducible parity runs.  ``Hybrid_TD3_Model`` (the agent loop, :288-560) is not built in this round.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import mlp as _mlp
from .replay import Memory  # noqa: F401  (the reference module exports its Memory class too)


def hidden_init(layer):                                                          # :112-117
    fan_in = layer.weight.data.size()[0]
    lim = 1. / np.sqrt(fan_in)
    return (0, lim)


def _mlp3(in_dims, widths, out_dims, device):
    """Linear-ReLU-Linear-ReLU[-Linear] at Sequential indices 0, 2, 4 as in the reference."""
    mods = [_mlp.Linear(in_dims, widths[0], device=device), nn.ReLU(), _mlp.Linear(widths[0], widths[1], device=device), nn.ReLU()]
    if out_dims is not None:
        mods.append(_mlp.Linear(widths[1], out_dims, device=device))
    return _mlp.Tower(*mods)


class Hybrid_Critic(nn.Module):
    """:119-189  twin Q networks over [BatchNorm(state) | discrete action (soft one-hot) | continuous action]."""

    def __init__(self, input_dims, action_nums, device=None):
        super().__init__()
        self.input_dims, self.action_nums = input_dims, action_nums
        deep_input_dims = self.input_dims + self.action_nums * 2
        self.bn_input = nn.BatchNorm1d(self.input_dims, device=device)
        self.mlp_1 = _mlp3(deep_input_dims, [400, 300], 1, device)
        self.mlp_2 = _mlp3(deep_input_dims, [400, 300], 1, device)

    def reset_parameters(self):                                                  # :155-159
        for i in range(3):
            if i % 2 == 0:
                self.mlp_1[i].weight.data.uniform_(*hidden_init(self.mlp_1[i]))
                self.mlp_2[i].weight.data.uniform_(*hidden_init(self.mlp_2[i]))

    def evaluate(self, input, c_actions, d_actions):
        obs = self.bn_input(input)
        x = torch.cat([obs, d_actions, c_actions], dim=-1)
        return self.mlp_1(x), self.mlp_2(x)

    def evaluate_q_1(self, input, c_actions, d_actions):
        obs = self.bn_input(input)
        return self.mlp_1(torch.cat([obs, d_actions, c_actions], dim=-1))


def boltzmann_softmax(actions, temprature):                                     # :258-259
    return (actions / temprature).exp() / torch.sum((actions / temprature).exp(), dim=-1).view(-1, 1)


def gumbel_softmax_sample(logits, temprature=1.0, hard=False, eps=1e-20, uniform_seed=1.0, U=None):
    """:261-270; ``U`` = the uniform draw (made on the logits' device when not given)."""
    if U is None:
        U = torch.rand_like(logits)
    y = logits + -torch.log(-torch.log(U + eps) + eps)
    y = F.softmax(y / temprature, dim=-1)
    if hard:
        y_hard = onehot_from_logits(y)
        y = (y_hard - y).detach() + y
    return y


def onehot_from_logits(logits, eps=0.0):                                        # :272-286 (greedy branch; eps > 0 as the reference)
    argmax_acs = (logits == logits.max(1, keepdim=True)[0]).float()
    if eps == 0.0:
        return argmax_acs
    rand_acs = torch.eye(logits.shape[1], device=logits.device)[
        torch.as_tensor(np.random.choice(range(logits.shape[1]), size=logits.shape[0]), device=logits.device)]
    return torch.stack([argmax_acs[i] if r > eps else rand_acs[i] for i, r in enumerate(torch.rand(logits.shape[0]))])


class Hybrid_Actor(nn.Module):
    """:191-256  shared trunk -> tanh continuous head + discrete-logit head; ``act`` perturbs both and Gumbel-samples the discrete one."""

    def __init__(self, input_dims, action_nums, device=None):
        super().__init__()
        self.input_dims, self.action_dims = input_dims, action_nums
        self.bn_input = nn.BatchNorm1d(self.input_dims, device=device)
        self.mlp = _mlp3(self.input_dims, [400, 300], None, device)
        self.c_action_layer = nn.Sequential(_mlp.Linear(300, self.action_dims, device=device), nn.Tanh())
        self.d_action_layer = nn.Sequential(_mlp.Linear(300, self.action_dims, device=device))

    def reset_parameters(self):                                                  # :221-224
        for i in range(3):
            if i % 2 == 0:
                self.mlp[i].weight.data.uniform_(*hidden_init(self.mlp[i]))

    def act(self, input, temprature, noise=None):
        """:229-246; ``noise`` = (eps_c, eps_d, U): the standard-normal draws behind the two ``torch.normal(x, 0.2)`` calls
        (``torch.normal(x, 0.2) = x + 0.2 * eps``) and the uniform of the Gumbel sample."""
        feature_exact = self.mlp(self.bn_input(input))
        c_action_means = self.c_action_layer(feature_exact)
        eps_c, eps_d, U = noise if noise is not None else (torch.randn_like(c_action_means), None, None)
        ensemble_c_actions = torch.softmax(c_action_means + (c_action_means + 0.2 * eps_c).detach(), dim=-1)
        d_action_q_values = self.d_action_layer(feature_exact)
        if eps_d is None:
            eps_d = torch.randn_like(d_action_q_values)
        d_action = gumbel_softmax_sample(logits=d_action_q_values + (d_action_q_values + 0.2 * eps_d).detach(),
                                         temprature=temprature, hard=False, U=U)
        ensemble_d_actions = torch.argmax(d_action, dim=-1) + 1
        return c_action_means, ensemble_c_actions, d_action, ensemble_d_actions.view(-1, 1)

    def evaluate(self, input):
        feature_exact = self.mlp(self.bn_input(input))
        return self.c_action_layer(feature_exact), self.d_action_layer(feature_exact)
