"""Network heads and replay memory of the reference's latest hybrid TD3 agent (``src/models/v10_Hybrid_TD3_model_PER.py``;
SURVEY 8f.4) on the B200 path: ``Memory`` (:19-110, = :class:`rl_ctr_prediction_b200.replay.Memory`: device-side prioritized
sampling), ``Hybrid_Critic`` (:119-189), ``Hybrid_Actor`` (:191-256), ``gumbel_softmax_sample`` / ``boltzmann_softmax`` /
``onehot_from_logits`` (:258-286).  Same class names, constructor arguments, state_dict keys and method signatures; every
``nn.Linear`` is :class:`rl_ctr_prediction_b200.mlp.Linear` (tcgen05 3xTF32 GEMMs; ``Linear -> ReLU`` pairs fused by
``mlp.Tower``).  The reference hard-codes ``.cuda()`` in ``gumbel_softmax_sample`` (:262); here the noise lives on the input's
device.  Random draws (the two Gaussian perturbations of ``act`` and the uniform of the Gumbel sample) can be injected for
reproducible parity runs.  ``Hybrid_TD3_Model`` (:288-560) is the agent: prioritized replay sampled on the device, twin-critic update
with gradient clipping, delayed actor update, Polyak targets.  Its two per-sample action-masking helpers (:427-472: "keep the d
largest continuous actions of a sample, d = its discrete action") are O(A^2) ``nonzero`` / ``index_put`` loops with a host
synchronisation each in the reference; here they are one rank comparison (no loop, no sync, differentiable).
"""
from __future__ import annotations

import copy
import random

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import mlp as _mlp
from . import optim as _optim
from .replay import Memory  # noqa: F401  (the reference module exports its Memory class too)


def setup_seed(seed):                                                            # :12-17
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.deterministic = True


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


def _keep_top_d(d_actions, c_actions):
    """keep[b, m] = column m is among the d_b largest continuous actions of sample b, d_b = argmax(d_actions[b]) + 1
    (:428-443 / :451-466: argsort(-c), then per (d, m) nonzero loops)."""
    choose_d = torch.argmax(d_actions, dim=-1) + 1
    rank = torch.argsort(torch.argsort(-c_actions, dim=-1), dim=-1)              # position of column m in the descending order
    return rank < choose_d.unsqueeze(1)


class Hybrid_TD3_Model():
    """:288-560."""

    def __init__(self, feature_nums, field_nums=15, latent_dims=5, action_nums=2, campaign_id='1458', lr_C_A=1e-3, lr_D_A=1e-3,
                 lr_C=1e-2, data_len=10, train_batch_size=10, reward_decay=1.0, memory_size=4096000, batch_size=256, tau=0.01,
                 device='cuda:0'):
        self.feature_nums, self.field_nums, self.action_nums, self.campaign_id = feature_nums, field_nums, action_nums, campaign_id
        self.lr_C_A, self.lr_D_A, self.lr_C = lr_C_A, lr_D_A, lr_C
        self.data_len, self.train_batch_size = data_len, train_batch_size
        self.gamma, self.latent_dims = reward_decay, latent_dims
        self.memory_size, self.batch_size, self.tau, self.device = memory_size, batch_size, tau, device
        setup_seed(1)                                                            # :323
        self.memory_counter = 0
        self.input_dims = self.field_nums * (self.field_nums - 1) // 2 + self.field_nums * self.latent_dims
        self.memory = Memory(self.memory_size, self.field_nums + self.action_nums * 2 + 2, self.device)
        self.Hybrid_Actor = Hybrid_Actor(self.input_dims, self.action_nums).to(self.device)
        self.Hybrid_Critic = Hybrid_Critic(self.input_dims, self.action_nums).to(self.device)
        self.Hybrid_Actor_ = copy.deepcopy(self.Hybrid_Actor)
        self.Hybrid_Critic_ = copy.deepcopy(self.Hybrid_Critic)
        self.optimizer_a = _optim.Adam(self.Hybrid_Actor.parameters(), lr=self.lr_C_A)          # :338-339 (no weight decay)
        self.optimizer_c = _optim.Adam(self.Hybrid_Critic.parameters(), lr=self.lr_C)
        self.loss_func = nn.MSELoss(reduction='mean')
        self.learn_iter = 0
        self.policy_freq = 10
        self.temprature = 1.0
        self.temprature_min = 0.1
        self.anneal_rate = 1e-6

    def store_transition(self, transitions):                                     # :350-356
        transitions = transitions.to(self.device, torch.float32)
        n = len(transitions)
        top = torch.max(self.memory.prioritys_)
        first = torch.where(top == 0., torch.ones_like(top), top).expand(n, 1)   # 1 while the memory holds no priority yet
        td_errors = torch.cat([first, transitions[:, -1].view(-1, 1)], dim=-1)
        self.memory.add(td_errors, transitions)

    def choose_action(self, state, random):                                      # :389-408
        self.Hybrid_Actor.eval()
        with torch.no_grad():
            c_actions, ensemble_c_actions, d_q_values, ensemble_d_actions = self.Hybrid_Actor.act(state, self.temprature)
            if random:
                c_actions = torch.clamp(torch.randn_like(c_actions), -1, 1)
                ensemble_c_actions = torch.softmax(c_actions, dim=-1)
                d_q_values = torch.softmax(torch.randn_like(d_q_values), dim=-1)
                ensemble_d_actions = torch.argmax(d_q_values, dim=-1) + 1
                return c_actions, ensemble_c_actions, d_q_values, ensemble_d_actions.view(-1, 1)
        self.Hybrid_Actor.train()
        return c_actions, ensemble_c_actions, d_q_values, ensemble_d_actions

    def choose_best_action(self, state):                                         # :410-421
        self.Hybrid_Actor.eval()
        with torch.no_grad():
            c_action_means, d_q_values = self.Hybrid_Actor.evaluate(state)
        ensemble_c_actions = torch.softmax(c_action_means, dim=-1)
        ensemble_d_actions = gumbel_softmax_sample(d_q_values, temprature=self.temprature_min, hard=True)
        ensemble_d_actions = torch.argmax(ensemble_d_actions, dim=-1) + 1
        return ensemble_d_actions.view(-1, 1), c_action_means, ensemble_c_actions

    def soft_update(self, net, net_target):                                      # :423-425
        with torch.no_grad():
            for pt, p in zip(net_target.parameters(), net.parameters()):
                pt.copy_(pt * (1.0 - self.tau) + p * self.tau)

    def to_next_state_c_actions(self, next_d_actions, next_c_actions, eps=None):
        """:427-448: kept entries become c + N(c, 0.2) = 2c + 0.2 eps, the rest 0, clamped to [-1, 1]."""
        keep = _keep_top_d(next_d_actions, next_c_actions)
        if eps is None:
            eps = torch.randn_like(next_c_actions)
        noisy = next_c_actions + (next_c_actions + 0.2 * eps)
        return torch.clamp(torch.where(keep, noisy, torch.zeros_like(noisy)), -1, 1)

    def to_current_state_c_actions(self, next_d_actions, next_c_actions):        # :450-472
        keep = _keep_top_d(next_d_actions, next_c_actions)
        return torch.clamp(torch.where(keep, next_c_actions, torch.zeros_like(next_c_actions)), -1, 1)

    def learn(self, embedding_layer, noise=None, sample=None):
        """:474-560.  ``noise`` = dict of injected draws (``eps_d``: the N(0,1) behind ``torch.normal(d_next, 0.2)``; ``U_next`` /
        ``U_now``: the Gumbel uniforms; ``eps_c``: the N(0,1) of the next-state action noise), ``sample`` = replay indices."""
        noise = noise or {}
        self.learn_iter += 1
        if (self.learn_iter + 1) % 1000 == 0:                                    # :477-479
            self.temprature = max(self.temprature_min, self.temprature - (self.temprature - self.temprature_min) * self.learn_iter
                                  / (self.data_len // self.train_batch_size))
        F_, A = self.field_nums, self.action_nums
        choose_idx, batch_memory, ISweights = self.memory.stochastic_sample(self.batch_size, sample=sample)
        b_s = embedding_layer.forward(batch_memory[:, :F_].long())
        b_c_a = batch_memory[:, F_: F_ + A].contiguous()
        b_d_a = batch_memory[:, F_ + A: F_ + A * 2].contiguous()
        b_r = torch.unsqueeze(batch_memory[:, -1], 1)
        b_s_ = b_s
        with torch.no_grad():
            c_next, d_next = self.Hybrid_Actor_.evaluate(b_s_)
            eps_d = noise.get("eps_d")
            if eps_d is None:
                eps_d = torch.randn_like(d_next)
            next_d_actions = gumbel_softmax_sample(logits=d_next + (d_next + 0.2 * eps_d), temprature=self.temprature, hard=False,
                                                   U=noise.get("U_next"))
            next_c_actions = self.to_next_state_c_actions(next_d_actions, c_next, eps=noise.get("eps_c"))
            q1_target, q2_target = self.Hybrid_Critic_.evaluate(b_s_, next_c_actions, next_d_actions)
            q_target = b_r + self.gamma * torch.min(q1_target, q2_target)
        q1, q2 = self.Hybrid_Critic.evaluate(b_s, b_c_a, b_d_a)
        critic_td_error = (q_target * 2 - q1 - q2).detach() / 2
        critic_loss = (ISweights * (F.mse_loss(q1, q_target, reduction='none') + F.mse_loss(q2, q_target, reduction='none'))).mean()
        self.optimizer_c.zero_grad()
        critic_loss.backward()
        nn.utils.clip_grad_norm_(self.Hybrid_Critic.parameters(), 0.5)
        self.optimizer_c.step()
        critic_loss_r = critic_loss.item()
        self.memory.batch_update(choose_idx, critic_td_error)
        if self.learn_iter % self.policy_freq == 0:
            c_means, d_q_values = self.Hybrid_Actor.evaluate(b_s)
            d_q_values_ = gumbel_softmax_sample(logits=d_q_values, temprature=self.temprature_min, hard=False, U=noise.get("U_now"))
            c_means_ = self.to_current_state_c_actions(d_q_values_, c_means)
            c_reg = (c_means ** 2).mean()
            d_reg = (d_q_values ** 2).mean()
            a_critic_value = self.Hybrid_Critic.evaluate_q_1(b_s, c_means_, d_q_values_)
            c_a_loss = (ISweights * (- a_critic_value)).mean() + (c_reg + d_reg) * 1e-2
            self.optimizer_a.zero_grad()
            c_a_loss.backward()
            nn.utils.clip_grad_norm_(self.Hybrid_Actor.parameters(), 0.5)
            self.optimizer_a.step()
            self.soft_update(self.Hybrid_Critic, self.Hybrid_Critic_)
            self.soft_update(self.Hybrid_Actor, self.Hybrid_Actor_)
        return critic_loss_r
