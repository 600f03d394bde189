"""Hybrid soft actor-critic agent of the reference (``src/models/Hybrid_SAC_model.py``) on the B200 path.

Same classes, constructor arguments, state_dict keys and method names as the reference (``Memory``, ``C_Actor``, ``D_Actor``,
``Hybrid_Q_network``, ``Hybrid_RL_Model`` with ``store_transition / choose_action / choose_best_action / soft_update /
hard_update / learn``), so ``src/all_main/hybrid_sac_main.py`` switches over by changing its import.  What runs underneath:

* every ``nn.Linear`` is :class:`rl_ctr_prediction_b200.mlp.Linear` (tcgen05 3xTF32 GEMMs, ReLU fused into the epilogue);
* every ``torch.optim.Adam`` is :class:`rl_ctr_prediction_b200.optim.Adam` (one multi-tensor launch per optimizer);
* the prioritized memory samples on the device (``rlctr_replay_sample_per``: weighted sampling without replacement), where the
  reference copies every priority to the host for ``np.random.choice`` (:64-83).

Networks and memory are pinned by golden vectors from the real reference (``tests/golden/make_golden_sac.py``).  ``learn`` is the
reference's arithmetic line by line (cited), including its quirk that the actor losses use ``Critic(b_s, b_c_a)`` with the
STORED continuous action (:418), with ONE necessary change: the reference's ``learn`` does not run under the installed torch -- its
entropy-tuning losses (:434,444) back-propagate through the actor graphs AFTER ``optimizer_c_a.step()`` / ``optimizer_d_a.step()``
have modified the actors in place, which autograd rejects (``RuntimeError: ... modified by an inplace operation``; recorded in the
golden file).  Here the entropies are detached in those two losses (the standard SAC temperature update, and the only reading
under which the code can run); every other line is as written.  Stochastic draws (the Gaussian noise of ``rsample``, the replay
indices) can be injected (``noise=``, ``sample=``) for reproducible steps.
"""
from __future__ import annotations

import copy

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import mlp as _mlp
from . import optim as _optim
from . import replay as _replay


class Memory(object):
    """Hybrid_SAC_model.py:21-108: single-column priorities; ``add`` gives new transitions priority max(old, 1); sampling
    probability proportional to the STORED priority (``batch_update`` applies (|td| + eps)^alpha when it stores)."""

    def __init__(self, memory_size, transition_lens, device, seed=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RlctrError("Hybrid_SAC_model.Memory lives on a CUDA (sm_100a) device; there is no CPU fallback")
        self.transition_lens = transition_lens
        self.epsilon = 1e-3
        self.alpha = 0.6
        self.beta = 0.4
        self.beta_increment_per_sampling = 1e-5
        self.abs_err_upper = 1
        self.memory_size = memory_size
        self.memory_counter = 0
        self.priorities_ = torch.zeros(size=[memory_size, 1], device=self.device)
        self.memory = torch.zeros(size=[memory_size, transition_lens], device=self.device)
        self._rng = _replay._rng(self.device, seed)
        self._ws = None

    def get_priority(self, td_error):                                            # :40-41
        return torch.pow(torch.abs(td_error) + self.epsilon, self.alpha)

    def add(self, transitions):                                                  # :43-62
        lib = _lib.load()
        n = len(transitions)
        tr = transitions.to(self.device, torch.float32).contiguous()
        _lib.check(lib.rlctr_replay_store(_lib.ptr(self.memory), self.memory_size, self.transition_lens, self.memory_counter,
                                          _lib.ptr(tr), n, self.transition_lens, _lib.stream()), "rlctr_replay_store")
        idx = (torch.arange(n, device=self.device) + self.memory_counter) % self.memory_size
        self.priorities_[idx] = torch.clamp(self.priorities_[idx], min=1.0)     # max(old, 1) :51,55,59
        self.memory_counter += n

    def _valid(self):
        return self.memory_size if self.memory_counter >= self.memory_size else self.memory_counter

    def _is_weights(self, idx):
        n = self._valid()
        min_prob = torch.min(self.priorities_[:n])                               # :68,72,88-91
        return torch.pow(torch.div(self.priorities_[idx], min_prob), -self.beta)

    def stochastic_sample(self, batch_size, sample=None):                        # :64-83
        lib = _lib.load()
        n = self._valid()
        if sample is not None:                                                   # injected indices (tests)
            idx = torch.as_tensor(sample, device=self.device).long().reshape(-1)
        else:
            if batch_size > n:
                raise ValueError("Cannot take a larger sample than population when 'replace=False'")
            wsb = lib.rlctr_replay_per_ws_bytes(n)
            if self._ws is None or self._ws.numel() < wsb:
                self._ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
            idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
            # P proportional to the stored priority: weight (|p| + 0)^1
            _lib.call("rlctr_replay_sample_per", lib.rlctr_replay_sample_per, _lib.ptr(self.priorities_), 1, n, 0.0, 1.0, 0.0, 0,
                      int(batch_size), _lib.ptr(self._rng), _lib.ptr(idx), None, _lib.ptr(self._ws), self._ws.numel(),
                      _lib.stream(), meta={"n": n, "batch": batch_size})
            _lib.check(lib.rlctr_rng_advance(_lib.ptr(self._rng), n, _lib.stream()), "rlctr_rng_advance")
        self.beta = torch.min(torch.FloatTensor([1., self.beta + self.beta_increment_per_sampling])).item()   # fp32, as the reference      # :77
        batch = torch.empty(idx.numel(), self.transition_lens, dtype=torch.float32, device=self.device)
        _lib.check(lib.rlctr_replay_gather(_lib.ptr(self.memory), self.transition_lens, _lib.ptr(idx), idx.numel(), _lib.ptr(batch),
                                           _lib.stream()), "rlctr_replay_gather")
        return idx, batch, self._is_weights(idx)

    def greedy_sample(self, batch_size):                                         # :85-104
        self.beta = torch.min(torch.FloatTensor([1., self.beta + self.beta_increment_per_sampling])).item()   # fp32, as the reference
        idx = torch.sort(-self.priorities_, dim=0)[1][:batch_size, :].squeeze(1)
        return idx, self.memory[idx], self._is_weights(idx).detach()

    def batch_update(self, choose_idx, td_errors):                               # :106-108
        lib = _lib.load()
        p = self.get_priority(td_errors.to(self.device, torch.float32)).reshape(-1).contiguous()
        idx = choose_idx.to(self.device, torch.int64).reshape(-1).contiguous()
        _lib.check(lib.rlctr_replay_update(_lib.ptr(self.priorities_), 1, _lib.ptr(idx), _lib.ptr(p), idx.numel(), _lib.stream()),
                   "rlctr_replay_update")


class C_Actor(nn.Module):
    """:118-173  Gaussian policy: BatchNorm1d -> (Linear, BatchNorm1d, ReLU) x 2 -> mean / clamped log-std heads."""

    def __init__(self, input_dims, action_nums, device=None):
        super().__init__()
        self.input_dims, self.action_nums = input_dims, action_nums
        self.bn_input = nn.BatchNorm1d(self.input_dims, device=device)
        hidden_dims = [256, 256]
        self.mlp = nn.Sequential(
            _mlp.Linear(self.input_dims, hidden_dims[0], device=device), nn.BatchNorm1d(hidden_dims[0], device=device), nn.ReLU(),
            _mlp.Linear(hidden_dims[0], hidden_dims[1], device=device), nn.BatchNorm1d(hidden_dims[1], device=device), nn.ReLU())
        self.mean_linear = _mlp.Linear(hidden_dims[1], self.action_nums, device=device)
        self.log_std_linear = _mlp.Linear(hidden_dims[1], self.action_nums, device=device)

    def forward(self, state):
        x = self.mlp(self.bn_input(state))
        action_mean = self.mean_linear(x)
        log_std = torch.clamp(self.log_std_linear(x), min=-20, max=2)           # :149
        return action_mean, log_std

    def sample(self, state, eps=None):
        """:155-167; ``eps`` = the standard-normal draw of ``Normal.rsample`` (drawn here when not given)."""
        mean, log_std = self.forward(state)
        std = log_std.exp()
        if eps is None:
            eps = torch.randn_like(mean)
        x_t = mean + std * eps                                                   # rsample
        y_t = torch.tanh(x_t)
        normal_log_prob = -((x_t - mean) ** 2) / (2 * std ** 2) - log_std - float(np.log(np.sqrt(2 * np.pi)))   # Normal.log_prob
        log_prob = (normal_log_prob - torch.log(1 - y_t.pow(2) + 1e-6)).sum(-1, keepdim=True)                   # :165
        return y_t, log_prob

    def evaluate(self, state):
        mean, _ = self.forward(state)
        return torch.tanh(mean)


class D_Actor(nn.Module):
    """:176-226  categorical policy: Linear -> ReLU -> Linear -> ReLU -> Linear -> softmax (``bn_input`` exists but is unused)."""

    def __init__(self, input_dims, action_dims, device=None):
        super().__init__()
        self.input_dims, self.action_dims = input_dims, action_dims
        self.bn_input = nn.BatchNorm1d(self.input_dims, device=device)
        hidden_dims = [256, 256]
        self.mlp_l1 = _mlp.Linear(self.input_dims, hidden_dims[0], device=device)
        self.mlp_l2 = _mlp.Linear(hidden_dims[0], hidden_dims[1], device=device)
        self.policy_layer = _mlp.Linear(hidden_dims[1], self.action_dims, device=device)

    def forward(self, state):
        x = self.mlp_l1(state, relu=True)                                        # Linear + ReLU in one GEMM epilogue
        x = self.mlp_l2(x, relu=True)
        return F.softmax(self.policy_layer(x), dim=-1)

    def sample(self, state, draw=None):
        """:208-219; ``draw`` = injected categorical samples (0-based), else ``Categorical(probs).sample()``."""
        action_probs = self.forward(state)
        if draw is None:
            draw = torch.multinomial(action_probs.detach(), 1).view(-1)
        actions = draw.view(-1, 1) + 1
        mirror = (action_probs == 0.0).float() * 1e-6
        return actions, action_probs, torch.log(action_probs + mirror)

    def evaluate(self, state):
        return torch.argmax(self.forward(state), dim=-1, keepdim=True) + 1


class Hybrid_Q_network(nn.Module):
    """:229-282  twin hybrid critics: each trunk feeds a continuous-action Q (trunk | action -> 1) and per-discrete-action Qs."""

    def __init__(self, input_dims, action_dims, device=None):
        super().__init__()
        self.input_dims, self.action_dims = input_dims, action_dims
        hidden_dims = [256, 256]
        self.mlp_q1_l1 = _mlp.Linear(self.input_dims, hidden_dims[0], device=device)
        self.mlp_q1_l2 = _mlp.Linear(hidden_dims[0], hidden_dims[1], device=device)
        self.c_q1 = _mlp.Linear(hidden_dims[1] + self.action_dims, 1, device=device)
        self.d_q1 = _mlp.Linear(hidden_dims[1], self.action_dims, device=device)
        self.mlp_q2_l1 = _mlp.Linear(self.input_dims, hidden_dims[0], device=device)
        self.mlp_q2_l2 = _mlp.Linear(hidden_dims[0], hidden_dims[1], device=device)
        self.c_q2 = _mlp.Linear(hidden_dims[1] + self.action_dims, 1, device=device)
        self.d_q2 = _mlp.Linear(hidden_dims[1], self.action_dims, device=device)

    def forward(self, state, action):
        x1 = self.mlp_q1_l2(self.mlp_q1_l1(state, relu=True), relu=True)
        c_q1 = self.c_q1(torch.cat([x1, action], dim=-1))
        d_q1 = self.d_q1(x1)
        x2 = self.mlp_q2_l2(self.mlp_q2_l1(state, relu=True), relu=True)
        c_q2 = self.c_q2(torch.cat([x2, action], dim=-1))
        d_q2 = self.d_q2(x2)
        return c_q1, d_q1, c_q2, d_q2


class Hybrid_RL_Model():
    """:285-461."""

    def __init__(self, feature_nums, field_nums=15, latent_dims=5, action_nums=2, campaign_id='1458', lr_C_A=3e-4, lr_D_A=3e-4,
                 lr_C=3e-4, reward_decay=1, memory_size=4096000, batch_size=256, tau=0.005, device='cuda:0'):
        self.feature_nums, self.field_nums, self.action_nums, self.campaign_id = feature_nums, field_nums, action_nums, campaign_id
        self.lr_C_A, self.lr_D_A, self.lr_C = lr_C_A, lr_D_A, lr_C
        self.gamma, self.latent_dims = reward_decay, latent_dims
        self.memory_size, self.batch_size, self.tau, self.device = memory_size, batch_size, tau, device
        self.memory_counter = 0
        self.input_dims = self.field_nums * (self.field_nums - 1) // 2 + self.field_nums * self.latent_dims      # :318
        self.memory = Memory(self.memory_size, self.field_nums + self.action_nums + 2, self.device)
        self.Critic = Hybrid_Q_network(self.input_dims, self.action_nums).to(self.device)
        self.Critic_ = copy.deepcopy(self.Critic)
        self.D_Actor = D_Actor(self.input_dims, self.action_nums).to(self.device)
        self.C_Actor = C_Actor(self.input_dims, self.action_nums).to(self.device)
        adam = lambda params, lr: _optim.Adam(params, lr=lr, eps=1e-8, weight_decay=1e-2)                       # :330-332
        self.optimizer_c_a = adam(self.C_Actor.parameters(), self.lr_C_A)
        self.optimizer_d_a = adam(self.D_Actor.parameters(), self.lr_D_A)
        self.optimizer_c = adam(self.Critic.parameters(), self.lr_C)
        self.c_target_entropy = -float(self.action_nums)                         # -prod([action_nums, 1]) :335
        self.c_log_alpha = torch.zeros(1, requires_grad=True, device=self.device)
        self.c_alpha = self.c_log_alpha.exp()
        self.optimizer_c_alpha = adam([self.c_log_alpha], lr_C)
        self.d_target_entropy = -np.log(1.0 / self.action_nums) * 0.98          # :341-342
        self.d_log_alpha = torch.zeros(1, requires_grad=True, device=self.device)
        self.d_alpha = self.d_log_alpha.exp()
        self.optimizer_d_alpha = adam([self.d_log_alpha], lr_C)
        self.learn_iter = 0

    def store_transition(self, transitions):
        self.memory.add(transitions)

    def choose_action(self, state):
        with torch.no_grad():
            c_actions, _ = self.C_Actor.sample(state)
            d_actions, _, _ = self.D_Actor.sample(state)
        return c_actions, torch.softmax(c_actions, dim=-1), d_actions

    def choose_best_action(self, state):
        with torch.no_grad():
            c_actions = self.C_Actor.evaluate(state)
            d_actions = self.D_Actor.evaluate(state)
        return torch.softmax(c_actions, dim=-1), d_actions

    def soft_update(self, net, net_target):
        with torch.no_grad():
            for pt, p in zip(net_target.parameters(), net.parameters()):
                pt.copy_(pt * (1.0 - self.tau) + p * self.tau)

    def hard_update(self, net, net_target):
        net_target.load_state_dict(net.state_dict())

    def learn(self, embedding_layer, noise=None, sample=None):
        """:377-461.  ``noise`` = [eps_next, eps_now] (the two Gaussian draws of C_Actor.sample, in call order) and ``sample`` =
        the replay indices, for reproducible runs; both drawn on the device when not given."""
        self.learn_iter += 1
        F_, A = self.field_nums, self.action_nums
        choose_idx, batch_memory, ISweights = self.memory.stochastic_sample(self.batch_size, sample=sample)
        b_s = embedding_layer.forward(batch_memory[:, :F_].long())
        b_c_a = batch_memory[:, F_: F_ + A].contiguous()
        b_discrete_a = torch.unsqueeze(batch_memory[:, F_ + A] - 1, 1).long()
        b_r = torch.unsqueeze(batch_memory[:, -1], 1)
        b_s_ = b_s
        eps_next, eps_now = (noise[0], noise[1]) if noise is not None else (None, None)
        with torch.no_grad():
            c_action_next, c_log_probs_next = self.C_Actor.sample(b_s_, eps_next)
            _, d_action_probs_next, d_log_probs_next = self.D_Actor.sample(b_s_, draw=torch.zeros(len(b_s_), dtype=torch.long,
                                                                                                   device=b_s_.device))
            c_q1_next, d_q1_next, c_q2_next, d_q2_next = self.Critic_.forward(b_s_, c_action_next)
            q_c_next_target = b_r + self.gamma * (torch.min(c_q1_next, c_q2_next) - self.c_alpha * c_log_probs_next)
            q_d_next_target = (b_r + self.gamma * (torch.min(d_q1_next, d_q2_next) - self.d_alpha * d_log_probs_next)
                               * d_action_probs_next).mean(dim=-1).unsqueeze(-1)                                # :396
        c_q1, d_q1, c_q2, d_q2 = self.Critic.forward(b_s, b_c_a)
        c_critic_loss = ISweights * ((c_q1 - q_c_next_target).pow(2) + (c_q2 - q_c_next_target).pow(2))
        d_critic_loss = ISweights * ((d_q1.gather(1, b_discrete_a) - q_d_next_target).pow(2)
                                     + (d_q2.gather(1, b_discrete_a) - q_d_next_target).pow(2))
        critic_loss = (c_critic_loss + d_critic_loss).mean()
        self.optimizer_c.zero_grad()
        critic_loss.backward()
        self.optimizer_c.step()
        td_errors = ((2 * q_c_next_target - c_q1 - c_q2) / 2 + 1e-6) + \
                    ((2 * q_d_next_target - d_q1.gather(1, b_discrete_a) - d_q2.gather(1, b_discrete_a)) / 2 + 1e-6)
        self.memory.batch_update(choose_idx, td_errors.detach())
        c_action, c_log_probs = self.C_Actor.sample(b_s, eps_now)
        _, d_action_probs, d_log_probs = self.D_Actor.sample(b_s, draw=torch.zeros(len(b_s), dtype=torch.long, device=b_s.device))
        c_q1, d_q1, c_q2, d_q2 = self.Critic.forward(b_s, b_c_a)                 # :418 (the stored action, as the reference)
        c_actor_loss = (ISweights * (self.c_alpha * c_log_probs - torch.min(c_q1, c_q2))).mean()
        self.optimizer_c_a.zero_grad()
        c_actor_loss.backward(retain_graph=True)
        self.optimizer_c_a.step()
        c_entropies = c_log_probs
        d_actor_loss = (ISweights * (d_action_probs * (self.d_alpha * d_log_probs - torch.min(d_q1, d_q2)))).mean()
        self.optimizer_d_a.zero_grad()
        d_actor_loss.backward(retain_graph=True)
        self.optimizer_d_a.step()
        d_entropies = torch.sum(d_action_probs * d_log_probs, dim=-1)
        c_alpha_loss = -(self.c_log_alpha * (c_entropies.detach() + self.c_target_entropy)).mean()       # detached: see module doc
        self.optimizer_c_alpha.zero_grad()
        c_alpha_loss.backward()
        self.optimizer_c_alpha.step()
        self.c_alpha = self.c_log_alpha.exp()
        d_alpha_loss = -(self.d_log_alpha * (d_entropies.detach() + self.d_target_entropy)).mean()
        self.optimizer_d_alpha.zero_grad()
        d_alpha_loss.backward()
        self.optimizer_d_alpha.step()
        self.d_alpha = self.d_log_alpha.exp()
        if self.learn_iter % 100 == 0:
            self.hard_update(self.Critic, self.Critic_)
        return critic_loss.item()
