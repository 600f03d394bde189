"""DDPG agent that emits the ensemble weights -- drop-in for the reference's
``src/models/DDPG_for_PG_model.py`` (``Actor`` :18-46, ``Critic`` :49-79, ``DDPG`` :81-251) as used by
``src/all_main/main.py``.

Actor: ``softmax(MLP(cat[state, BN_1(a_ddqn)]))`` over the M CTR models; Critic: ``MLP(cat[state, BN_1(a_ddqn),
weights])``; both ``-> 300 -> 300 -> 300 ->`` with BatchNorm1d + ReLU.  Exploration: per sample, with
probability epsilon the action is ``softmax(Normal(mu, epsilon))`` (:175-192).  ``learn_c`` / ``learn_a`` :227-251,
Polyak ``tau = 0.005`` over ``parameters()`` only -- BatchNorm running statistics are buffers and are NOT
soft-updated (:203-205), exactly like the reference.

Linear layers are :class:`.mlp.Linear` (tcgen05 3xTF32 GEMMs), optimizers :class:`.optim.Adam`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import mlp as _mlp
from . import optim as _optim
from .DDQN_model import RingMemory, bn_mlp, state_dims


class Actor(nn.Module):
    def __init__(self, input_dims, action_nums, feature_nums=None, field_nums=None, latent_dims=None, device=None):
        super().__init__()
        self.input_dims = input_dims
        self.bn_input = nn.BatchNorm1d(1, device=device)
        self.mlp = bn_mlp(input_dims + 1, action_nums, device=device)

    def forward(self, input, ddqn_a):
        obs = torch.cat([input, _mlp.apply_bn(self.bn_input, ddqn_a)], dim=1)
        return torch.softmax(self.mlp(obs), dim=1)


class Critic(nn.Module):
    def __init__(self, input_dims, action_nums, feature_nums=None, field_nums=None, latent_dims=None, device=None):
        super().__init__()
        self.bn_input = nn.BatchNorm1d(1, device=device)
        self.mlp = bn_mlp(input_dims + action_nums + 1, action_nums, device=device)

    def forward(self, input, action, ddqn_a):
        obs = torch.cat([input, _mlp.apply_bn(self.bn_input, ddqn_a)], dim=1)
        return self.mlp(torch.cat([obs, action], dim=1))


class DDPG:
    device_rng = False      # see DDQN_model.DoubleDQN
    grad_sync = None

    def __init__(self, feature_nums, field_nums=15, latent_dims=5, action_nums=2, campaign_id="1458", lr_A=1e-4, lr_C=1e-3,
                 reward_decay=1, memory_size=4096000, batch_size=256, tau=0.005, device="cuda:0"):
        self.feature_nums, self.field_nums, self.action_nums, self.campaign_id = feature_nums, field_nums, action_nums, campaign_id
        self.lr_A, self.lr_C, self.gamma, self.latent_dims = lr_A, lr_C, reward_decay, latent_dims
        self.memory_size, self.batch_size, self.tau, self.device = memory_size, batch_size, tau, device
        self.input_dims = state_dims(field_nums, latent_dims)
        self._mem = RingMemory(memory_size, [field_nums, action_nums + 1, 1], device)      # :116-118
        mk = lambda cls: cls(self.input_dims, action_nums, feature_nums, field_nums, latent_dims).to(device)
        self.Actor, self.Critic, self.Actor_, self.Critic_ = mk(Actor), mk(Critic), mk(Actor), mk(Critic)
        self.optimizer_a = _optim.Adam(self.Actor.parameters(), lr=self.lr_A, weight_decay=1e-5)
        self.optimizer_c = _optim.Adam(self.Critic.parameters(), lr=self.lr_C, weight_decay=1e-5)
        self.loss_func = nn.MSELoss(reduction="mean")

    @property
    def memory_state(self):
        return self._mem.bufs[0]

    @property
    def memory_action_reward(self):
        return self._mem.bufs[1]

    @property
    def memory_ddqn_action(self):
        return self._mem.bufs[2]

    @property
    def memory_counter(self):
        return self._mem.counter

    def store_transition(self, features, action_rewards, ddqn_actions):
        self._mem.store(features, action_rewards, ddqn_actions)

    def _actor_eval(self, state, ddqn_a):
        self.Actor.eval()
        with torch.no_grad():
            return self.Actor.forward(state, ddqn_a)

    def choose_action(self, state, ddqn_a, exploration_rate):
        """:175-192."""
        action = self._actor_eval(state, ddqn_a)
        self.Actor.train()
        random_seeds = torch.rand(len(state), 1, device=self.device if self.device_rng else None).to(self.device)
        random_action = torch.softmax(torch.normal(action, exploration_rate), dim=1)
        return torch.where(random_seeds >= exploration_rate, action, random_action)

    def choose_best_action(self, state, ddqn_a):
        action = self._actor_eval(state, ddqn_a)
        return action, torch.softmax(action, dim=1)

    def soft_update(self, net, net_target):
        with torch.no_grad():
            for pt, p in zip(net_target.parameters(), net.parameters()):
                pt.mul_(1.0 - self.tau).add_(p, alpha=self.tau)

    def sample_batch(self):
        idx = self._mem.sample_index(self.batch_size, self.device)
        b_s = self.memory_state[idx, :].long()
        ar = self.memory_action_reward[idx, :]
        return b_s, ar[:, :self.action_nums], torch.unsqueeze(ar[:, self.action_nums], 1), b_s, self.memory_ddqn_action[idx, :]

    def learn_c(self, b_s, b_a, b_r, b_s_, b_ddqn_a):
        q_target = b_r + self.gamma * self.Critic_.forward(b_s_, self.Actor_.forward(b_s_, b_ddqn_a), b_ddqn_a).detach()
        q = self.Critic.forward(b_s, b_a, b_ddqn_a)
        td_error = self.loss_func(q, q_target)
        self.optimizer_c.zero_grad()
        td_error.backward()
        if self.grad_sync is not None:
            self.grad_sync(self.Critic.parameters())
        self.optimizer_c.step()
        return td_error.item()

    def learn_a(self, b_s, b_ddqn_a):
        a_loss = -self.Critic.forward(b_s, self.Actor.forward(b_s, b_ddqn_a), b_ddqn_a).mean()
        self.optimizer_a.zero_grad()
        a_loss.backward()
        if self.grad_sync is not None:
            self.grad_sync(self.Actor.parameters())
        self.optimizer_a.step()
        return a_loss.item()
