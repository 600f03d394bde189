"""Actor-critic head of the reference's hybrid PPO agent (``src/models/Hybrid_PPO_model.py:20-96``) on the B200
path: the shared trunk ``(Linear, BatchNorm1d, ReLU) x 3`` with the value / continuous / discrete heads, ``best_a`` and
``evaluate`` (log-probabilities and entropies of the Normal / Categorical policies).  Same class name, constructor arguments and
state_dict keys; every ``nn.Linear`` is :class:`rl_ctr_prediction_b200.mlp.Linear`.  The reference builds ``action_std`` with
``.cuda()`` (:49); here it follows the module's device.  ``Hybrid_PPO_Model`` (the agent loop, :98-259) is not built in this round.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import mlp as _mlp


class Hybrid_Actor_Critic(nn.Module):
    def __init__(self, input_dims, action_nums, device=None):
        super().__init__()
        self.input_dims = input_dims
        neuron_nums = [300, 300, 300]
        mods, d = [], self.input_dims
        for w in neuron_nums:                                                    # :27-38
            mods += [_mlp.Linear(d, w, device=device), nn.BatchNorm1d(w, device=device), nn.ReLU()]
            d = w
        self.mlp = nn.Sequential(*mods)
        self.Critic = _mlp.Linear(neuron_nums[2], 1, device=device)
        self.Continuous_Actor = _mlp.Linear(neuron_nums[2], action_nums, device=device)
        self.Discrete_Actor = _mlp.Linear(neuron_nums[2], action_nums - 1, device=device)
        self.register_buffer("action_std", torch.ones(1, action_nums, device=device), persistent=False)   # :49 (not in the state_dict)
        self.action_nums = action_nums

    @staticmethod
    def _normal_log_prob(value, loc, scale):                                     # torch.distributions.Normal.log_prob
        return -((value - loc) ** 2) / (2 * scale ** 2) - scale.log() - math.log(math.sqrt(2 * math.pi))

    def act(self, input, c_noise=None, d_draw=None):
        """:52-70; ``c_noise`` = the standard-normal draw of ``Normal.sample``, ``d_draw`` = the categorical sample (0-based)."""
        mlp_out = self.mlp(input)
        c_action_means = torch.softmax(self.Continuous_Actor(mlp_out), dim=-1)
        if c_noise is None:
            c_noise = torch.randn_like(c_action_means)
        with torch.no_grad():
            c_actions = c_action_means + self.action_std * c_noise               # Normal.sample() carries no gradient
        c_action_logprobs = self._normal_log_prob(c_actions, c_action_means, self.action_std.expand_as(c_action_means))
        ensemble_c_actions = torch.softmax(c_actions, dim=-1)
        d_probs = torch.softmax(self.Discrete_Actor(mlp_out), dim=-1)
        if d_draw is None:
            d_draw = torch.multinomial(d_probs.detach(), 1).view(-1)
        d_action_logprobs = torch.log(d_probs / d_probs.sum(-1, keepdim=True)).gather(1, d_draw.view(-1, 1))
        return (c_actions, c_action_logprobs, ensemble_c_actions), (d_draw.view(-1, 1), d_action_logprobs, (d_draw + 2).view(-1, 1))

    def best_a(self, input):                                                     # :72-78
        mlp_out = self.mlp(input)
        return torch.softmax(self.Continuous_Actor(mlp_out), dim=-1), torch.softmax(self.Discrete_Actor(mlp_out), dim=-1)

    def evaluate(self, input, c_a, d_a):
        """:80-96.  Note the reference passes the RAW discrete-head outputs to ``Categorical`` as probabilities (:91), which
        normalises them by their sum (and rejects negative values); reproduced literally."""
        mlp_out = self.mlp(input)
        state_value = self.Critic(mlp_out)
        c_action_means = torch.softmax(self.Continuous_Actor(mlp_out), dim=-1)
        std = self.action_std.expand_as(c_action_means)
        c_action_logprobs = self._normal_log_prob(c_a, c_action_means, std)
        c_action_entropy = 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)   # Normal.entropy
        d_action_values = self.Discrete_Actor(mlp_out)
        dist = torch.distributions.Categorical(d_action_values)
        d_actions_logprobs = dist.log_prob(d_a.squeeze(1)).view(-1, 1)
        d_action_entropy = dist.entropy().view(-1, 1)
        return state_value, c_action_logprobs, c_action_entropy, d_actions_logprobs, d_action_entropy
