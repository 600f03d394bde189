"""Actor-critic head of the reference's hybrid PPO agent (``src/models/Hybrid_PPO_model.py:20-96``) on the B200
path: the shared trunk ``(Linear, BatchNorm1d, ReLU) x 3`` with the value / continuous / discrete heads, ``best_a`` and
``evaluate`` (log-probabilities and entropies of the Normal / Categorical policies).  Same class name, constructor arguments and
state_dict keys; every ``nn.Linear`` is :class:`rl_ctr_prediction_b200.mlp.Linear`.  The reference builds ``action_std`` with
``.cuda()`` (:49); here it follows the module's device.  ``Hybrid_PPO_Model`` (:98-259) is the agent: rollout memory, clipped
surrogate update for ``k_epochs``; its generalised-advantage loop (:206-212: a reversed Python loop with one ``.item()`` host
synchronisation per sample) is one fp64 scan on the device (``rlctr_gae_scan``).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib
from . import mlp as _mlp
from . import optim as _optim


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


def gae_advantages(deltas, gamma_lambda):
    """:206-212 on the device: ``adv = 0; for i, d in enumerate(reversed(deltas)): adv = c * adv + d; advantages[i] = adv``."""
    lib = _lib.load()
    d = deltas.detach().reshape(-1).float().contiguous()
    n = d.numel()
    out = torch.empty(n, 1, dtype=torch.float32, device=d.device)
    wsb = lib.rlctr_gae_ws_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=d.device)
    _lib.call("rlctr_gae_scan", lib.rlctr_gae_scan, _lib.ptr(d), n, float(gamma_lambda), _lib.ptr(out), _lib.ptr(ws), wsb, _lib.stream(),
              meta={"n": n})
    return out


class Hybrid_PPO_Model():
    """:98-259."""

    def __init__(self, feature_nums, field_nums=15, latent_dims=5, action_nums=2, campaign_id='1458', init_lr=1e-2, train_epochs=500,
                 reward_decay=1, lr_lamda=0.01, memory_size=4096000, batch_size=256, tau=0.005, k_epochs=3, eps_clip=0.2,
                 device='cuda:0'):
        self.feature_nums, self.field_nums, self.action_nums, self.campaign_id = feature_nums, field_nums, action_nums, campaign_id
        self.init_lr, self.train_epochs, self.gamma, self.latent_dims = init_lr, train_epochs, reward_decay, latent_dims
        self.lr_lamda, self.memory_size, self.batch_size, self.tau, self.device = lr_lamda, memory_size, batch_size, tau, device
        self.action_std = 0.5
        self.k_epochs, self.eps_clip = k_epochs, eps_clip
        self.lamda = 0.95
        self.memory_counter = 0
        self.input_dims = self.field_nums * (self.field_nums - 1) // 2 + self.field_nums * self.latent_dims
        z = lambda w: torch.zeros(size=[self.memory_size, w], device=self.device)
        self.memory_state, self.memory_c_a, self.memory_c_logprobs = z(self.field_nums), z(self.action_nums), z(self.action_nums)
        self.memory_d_a, self.memory_d_logprobs, self.memory_reward = z(1), z(1), z(1)
        self.hybrid_actor_critic = Hybrid_Actor_Critic(self.input_dims, self.action_nums).to(self.device)
        self.hybrid_actor_critic_old = Hybrid_Actor_Critic(self.input_dims, self.action_nums).to(self.device)
        self.optimizer = _optim.Adam(self.hybrid_actor_critic.parameters(), lr=self.init_lr, weight_decay=1e-5)   # :150
        self.loss_func = nn.MSELoss()

    def store_memory(self, states, c_a, c_logprobs, d_a, d_logprobs, rewards):
        """:154-170 (as written: ``index_end = counter + len`` and the counter is never advanced, so a rollout overwrites the
        head of the memory)."""
        n = len(states)
        lo = self.memory_counter % self.memory_size
        hi = self.memory_counter + n
        self.memory_state[lo:hi, :] = states
        self.memory_c_a[lo:hi, :] = c_a
        self.memory_c_logprobs[lo:hi, :] = c_logprobs
        self.memory_d_a[lo:hi, :] = d_a
        self.memory_d_logprobs[lo:hi, :] = d_logprobs
        self.memory_reward[lo:hi, :] = rewards

    def choose_a(self, state):
        self.hybrid_actor_critic.eval()
        with torch.no_grad():
            return self.hybrid_actor_critic.act(state)

    def choose_best_a(self, state):
        self.hybrid_actor_critic.eval()
        with torch.no_grad():
            ensemble_c_actions, d_actions = self.hybrid_actor_critic.best_a(state)
            ensemble_d_actions = torch.argsort(-d_actions)[:, 0] + 2
        return ensemble_c_actions, ensemble_d_actions.view(-1, 1)

    def memory(self):
        s = self.memory_state.long()
        return s, s, self.memory_c_a, self.memory_c_logprobs, self.memory_d_a, self.memory_d_logprobs, self.memory_reward

    def learn(self, states, states_, old_c_a, old_c_a_logprobs, old_d_a, old_d_a_logprobs, rewards):
        """:194-258."""
        ac = self.hybrid_actor_critic
        return_loss = 0
        old_d_a = old_d_a.long() if old_d_a.dtype != torch.int64 else old_d_a
        value_of_states_ = ac.evaluate(states_, old_c_a, old_d_a)
        value_of_states = ac.evaluate(states, old_c_a, old_d_a)
        td_target = rewards + self.gamma * value_of_states_[0]
        deltas = td_target - value_of_states[0]
        advantages = gae_advantages(deltas, self.gamma * self.lamda)              # :206-209, no per-sample host sync
        advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-5)
        for _ in range(self.k_epochs):
            state_values, c_a_logprobs, c_a_entropy, d_a_logprobs, d_a_entropy = ac.evaluate(states, old_c_a, old_d_a)
            ratios = torch.exp(c_a_logprobs - old_c_a_logprobs)
            c_a_loss = -torch.min(ratios * advantages, torch.clamp(ratios, 1 - self.eps_clip, 1 + self.eps_clip) * advantages).mean()
            c_a_entropy_loss = 0.01 * c_a_entropy.mean()
            ratios = torch.exp(d_a_logprobs - old_d_a_logprobs)
            d_a_loss = -torch.min(ratios * advantages, torch.clamp(ratios, 1 - self.eps_clip, 1 + self.eps_clip) * advantages).mean()
            d_a_entropy_loss = 0.01 * d_a_entropy.mean()
            critic_loss = self.loss_func(state_values, td_target.detach())
            loss = c_a_loss - c_a_entropy_loss + d_a_loss - d_a_entropy_loss + 0.5 * critic_loss
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
            return_loss = loss.mean().item()
        self.hybrid_actor_critic_old.load_state_dict(self.hybrid_actor_critic.state_dict())
        return return_loss
