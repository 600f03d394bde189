"""Hybrid soft actor-critic agent on the B200 path (counterpart of the reference's ``src/models/Hybrid_SAC_model.py``).

Drop-in surface: the class names (``Memory``, ``C_Actor``, ``D_Actor``, ``Hybrid_Q_network``, ``Hybrid_RL_Model``), their
constructor arguments, ``state_dict`` keys and public methods are the reference's, so ``src/all_main/hybrid_sac_main.py``
switches over by changing one import.  Underneath:

* dense layers are :class:`rl_ctr_prediction_b200.mlp.Linear` (tcgen05 3xTF32 GEMMs; ``Linear -> ReLU`` in one epilogue);
* optimizers are :class:`rl_ctr_prediction_b200.optim.Adam` (one multi-tensor launch per optimizer step);
* the prioritized memory is :class:`rl_ctr_prediction_b200.replay.PrioritizedBuffer`: weighted sampling without replacement on
  the device instead of a D2H copy of all priorities and ``np.random.choice`` (reference ``:64-83``).

Parity: networks, their gradients and the memory are pinned by golden vectors from the real reference
(``tests/golden/make_golden_sac.py``).  ``learn`` follows the reference's update equations (``:377-461``, cited per block below)
including its quirk that the actor losses are taken against ``Critic(state, STORED action)`` (``:418``), but it cannot be pinned
and differs in ONE place: the reference's ``learn`` raises under the installed torch, because its temperature losses
(``:434,444``) back-propagate through the actor graphs after ``optimizer_c_a.step()`` / ``optimizer_d_a.step()`` have changed the
actors in place (the error message is stored in the golden file).  Here the entropies enter the temperature losses detached --
the standard SAC temperature update and the only reading under which those lines can run.  The stochastic draws (Gaussian
noise of ``rsample``, replay indices) can be injected for reproducible steps.
"""
from __future__ import annotations

import copy
import math

import torch
import torch.nn as nn

from . import mlp as _mlp
from . import optim as _optim
from .replay import PrioritizedBuffer

_HIDDEN = (256, 256)
_LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


class Memory(PrioritizedBuffer):
    """The SAC flavour of the prioritized buffer (see the table in :class:`replay.PrioritizedBuffer`): one priority column that
    already holds (|td| + eps)^alpha, new transitions enter with priority max(old, 1), beta 0.4 -> 1 in steps of 1e-5."""

    _COLS, _RAW, _PRIO_NAME, _BETA0, _BETA_INC = 1, False, "priorities_", 0.4, 1e-5

    def add(self, transitions):
        n = len(transitions)
        self._ring_write(self.memory, self.transition_lens, transitions)
        slots = (torch.arange(n, device=self.device) + self.memory_counter) % self.memory_size
        self.priorities_[slots] = self.priorities_[slots].clamp_min(1.0)
        self.memory_counter += n


def _trunk(in_dims, device, batch_norm):
    """(Linear[, BatchNorm1d], ReLU) x 2 with the reference's Sequential indices."""
    mods, d = [], in_dims
    for width in _HIDDEN:
        mods.append(_mlp.Linear(d, width, device=device))
        if batch_norm:
            mods.append(nn.BatchNorm1d(width, device=device))
        mods.append(nn.ReLU())
        d = width
    return nn.Sequential(*mods)


class C_Actor(nn.Module):
    """Gaussian policy (reference ``:118-173``): BatchNorm on the state, a batch-normalised trunk, a mean head and a log-std head
    clamped to [-20, 2]; actions are tanh-squashed."""

    def __init__(self, input_dims, action_nums, device=None):
        super().__init__()
        self.input_dims, self.action_nums = input_dims, action_nums
        self.bn_input = nn.BatchNorm1d(input_dims, device=device)
        self.mlp = _trunk(input_dims, device, batch_norm=True)
        self.mean_linear = _mlp.Linear(_HIDDEN[-1], action_nums, device=device)
        self.log_std_linear = _mlp.Linear(_HIDDEN[-1], action_nums, device=device)

    def forward(self, state):
        h = self.mlp(self.bn_input(state))
        return self.mean_linear(h), self.log_std_linear(h).clamp(-20, 2)

    def sample(self, state, eps=None):
        """Reparameterised sample and its log-density under the squashed Gaussian (``:155-167``).  ``eps``: the N(0, 1) draw."""
        mu, log_sigma = self.forward(state)
        sigma = log_sigma.exp()
        u = mu + sigma * (torch.randn_like(mu) if eps is None else eps)
        a = torch.tanh(u)
        gauss = -(u - mu).pow(2) / (2 * sigma.pow(2)) - log_sigma - _LOG_SQRT_2PI
        return a, (gauss - torch.log(1 - a.pow(2) + 1e-6)).sum(-1, keepdim=True)

    def evaluate(self, state):
        return torch.tanh(self.forward(state)[0])


class D_Actor(nn.Module):
    """Categorical policy (``:176-226``): two ReLU layers and a softmax head (the reference creates ``bn_input`` and never uses it;
    it is kept for the state_dict)."""

    def __init__(self, input_dims, action_dims, device=None):
        super().__init__()
        self.input_dims, self.action_dims = input_dims, action_dims
        self.bn_input = nn.BatchNorm1d(input_dims, device=device)
        self.mlp_l1 = _mlp.Linear(input_dims, _HIDDEN[0], device=device)
        self.mlp_l2 = _mlp.Linear(_HIDDEN[0], _HIDDEN[1], device=device)
        self.policy_layer = _mlp.Linear(_HIDDEN[1], action_dims, device=device)

    def forward(self, state):
        return torch.softmax(self.policy_layer(self.mlp_l2(self.mlp_l1(state, relu=True), relu=True)), dim=-1)

    def sample(self, state, draw=None):
        """(actions in 1..A, probabilities, log-probabilities with exact zeros nudged by 1e-6) -- ``:208-219``."""
        probs = self.forward(state)
        if draw is None:
            draw = torch.multinomial(probs.detach(), 1).view(-1)
        return draw.view(-1, 1) + 1, probs, torch.log(probs + (probs == 0.0).float() * 1e-6)

    def evaluate(self, state):
        return self.forward(state).argmax(dim=-1, keepdim=True) + 1


class Hybrid_Q_network(nn.Module):
    """Twin critics (``:229-282``); critic i has a trunk ``mlp_qi_l1/l2``, a continuous head ``c_qi`` on [trunk | action] and a
    discrete head ``d_qi`` with one value per discrete action."""

    def __init__(self, input_dims, action_dims, device=None):
        super().__init__()
        self.input_dims, self.action_dims = input_dims, action_dims
        for i in (1, 2):                     # creation order q1 trunk, q1 heads, q2 trunk, q2 heads: same seed -> same init
            setattr(self, f"mlp_q{i}_l1", _mlp.Linear(input_dims, _HIDDEN[0], device=device))
            setattr(self, f"mlp_q{i}_l2", _mlp.Linear(_HIDDEN[0], _HIDDEN[1], device=device))
            setattr(self, f"c_q{i}", _mlp.Linear(_HIDDEN[1] + action_dims, 1, device=device))
            setattr(self, f"d_q{i}", _mlp.Linear(_HIDDEN[1], action_dims, device=device))

    def _one(self, i, state, action):
        h = getattr(self, f"mlp_q{i}_l2")(getattr(self, f"mlp_q{i}_l1")(state, relu=True), relu=True)
        return getattr(self, f"c_q{i}")(torch.cat([h, action], dim=-1)), getattr(self, f"d_q{i}")(h)

    def forward(self, state, action):
        c1, d1 = self._one(1, state, action)
        c2, d2 = self._one(2, state, action)
        return c1, d1, c2, d2


class Hybrid_RL_Model():
    """The agent (``:285-461``)."""

    def __init__(self, feature_nums, field_nums=15, latent_dims=5, action_nums=2, campaign_id='1458', lr_C_A=3e-4, lr_D_A=3e-4,
                 lr_C=3e-4, reward_decay=1, memory_size=4096000, batch_size=256, tau=0.005, device='cuda:0'):
        self.feature_nums, self.field_nums, self.latent_dims, self.action_nums = feature_nums, field_nums, latent_dims, action_nums
        self.campaign_id, self.device = campaign_id, device
        self.lr_C_A, self.lr_D_A, self.lr_C = lr_C_A, lr_D_A, lr_C
        self.gamma, self.tau, self.memory_size, self.batch_size = reward_decay, tau, memory_size, batch_size
        self.memory_counter, self.learn_iter = 0, 0
        F_, A = field_nums, action_nums
        self.input_dims = F_ * (F_ - 1) // 2 + F_ * latent_dims
        self.memory = Memory(memory_size, F_ + A + 2, device)             # [features | continuous action | discrete action | reward]
        self.Critic = Hybrid_Q_network(self.input_dims, A).to(device)
        self.Critic_ = copy.deepcopy(self.Critic)
        self.D_Actor = D_Actor(self.input_dims, A).to(device)
        self.C_Actor = C_Actor(self.input_dims, A).to(device)
        make_opt = lambda params, lr: _optim.Adam(params, lr=lr, eps=1e-8, weight_decay=1e-2)
        self.optimizer_c_a = make_opt(self.C_Actor.parameters(), lr_C_A)
        self.optimizer_d_a = make_opt(self.D_Actor.parameters(), lr_D_A)
        self.optimizer_c = make_opt(self.Critic.parameters(), lr_C)
        # temperatures, tuned towards -|A| (continuous) and 0.98 * log|A| (discrete)
        self.c_target_entropy = -float(A)
        self.d_target_entropy = 0.98 * math.log(A)
        self.c_log_alpha = torch.zeros(1, requires_grad=True, device=device)
        self.d_log_alpha = torch.zeros(1, requires_grad=True, device=device)
        self.c_alpha, self.d_alpha = self.c_log_alpha.exp(), self.d_log_alpha.exp()
        self.optimizer_c_alpha = make_opt([self.c_log_alpha], lr_C)
        self.optimizer_d_alpha = make_opt([self.d_log_alpha], lr_C)

    # ---- acting ----------------------------------------------------------------------------------------------------
    def store_transition(self, transitions):
        self.memory.add(transitions)

    @torch.no_grad()
    def choose_action(self, state):
        c = self.C_Actor.sample(state)[0]
        return c, torch.softmax(c, dim=-1), self.D_Actor.sample(state)[0]

    @torch.no_grad()
    def choose_best_action(self, state):
        return torch.softmax(self.C_Actor.evaluate(state), dim=-1), self.D_Actor.evaluate(state)

    @torch.no_grad()
    def soft_update(self, net, net_target):
        for tgt, src in zip(net_target.parameters(), net.parameters()):
            tgt.copy_(tgt * (1.0 - self.tau) + src * self.tau)

    def hard_update(self, net, net_target):
        net_target.load_state_dict(net.state_dict())

    # ---- learning --------------------------------------------------------------------------------------------------
    @staticmethod
    def _step(optimizer, loss, retain=False):
        optimizer.zero_grad()
        loss.backward(retain_graph=retain)
        optimizer.step()

    @torch.no_grad()
    def _targets(self, s_next, reward, eps):
        """Soft Bellman targets of both heads (``:388-396``)."""
        a_next, logp_next = self.C_Actor.sample(s_next, eps)
        _, pi_next, log_pi_next = self.D_Actor.sample(s_next, draw=torch.zeros(len(s_next), dtype=torch.long, device=s_next.device))
        c1, d1, c2, d2 = self.Critic_(s_next, a_next)
        y_c = reward + self.gamma * (torch.min(c1, c2) - self.c_alpha * logp_next)
        y_d = (reward + self.gamma * (torch.min(d1, d2) - self.d_alpha * log_pi_next) * pi_next).mean(dim=-1, keepdim=True)
        return y_c, y_d

    def learn(self, embedding_layer, noise=None, sample=None):
        """One update of critics, both actors and both temperatures.  ``noise`` = (eps for the target sample, eps for the actor
        sample), ``sample`` = replay indices; drawn on the device when omitted.  Returns the critic loss."""
        self.learn_iter += 1
        F_, A = self.field_nums, self.action_nums
        picked, rows, w = self.memory.stochastic_sample(self.batch_size, sample=sample)
        s = embedding_layer.forward(rows[:, :F_].long())
        a_c = rows[:, F_:F_ + A].contiguous()
        a_d = (rows[:, F_ + A] - 1).long().unsqueeze(1)
        r = rows[:, -1:].contiguous()
        eps_t, eps_a = noise if noise is not None else (None, None)

        # critics (:398-412): IS-weighted squared errors of the four heads; priorities <- td errors
        y_c, y_d = self._targets(s, r, eps_t)                              # next state = state (:386)
        c1, d1, c2, d2 = self.Critic(s, a_c)
        q_d1, q_d2 = d1.gather(1, a_d), d2.gather(1, a_d)
        critic_loss = (w * ((c1 - y_c).pow(2) + (c2 - y_c).pow(2) + (q_d1 - y_d).pow(2) + (q_d2 - y_d).pow(2))).mean()
        self._step(self.optimizer_c, critic_loss)
        td = ((2 * y_c - c1 - c2) / 2 + 1e-6) + ((2 * y_d - q_d1 - q_d2) / 2 + 1e-6)
        self.memory.batch_update(picked, td.detach())

        # actors (:414-431): against the critics evaluated on the STORED continuous action, as the reference does
        _, logp = self.C_Actor.sample(s, eps_a)
        _, pi, log_pi = self.D_Actor.sample(s, draw=torch.zeros(len(s), dtype=torch.long, device=s.device))
        c1, d1, c2, d2 = self.Critic(s, a_c)
        self._step(self.optimizer_c_a, (w * (self.c_alpha * logp - torch.min(c1, c2))).mean(), retain=True)
        self._step(self.optimizer_d_a, (w * (pi * (self.d_alpha * log_pi - torch.min(d1, d2)))).mean(), retain=True)

        # temperatures (:433-447), entropies detached (see the module docstring)
        h_c, h_d = logp.detach(), (pi * log_pi).sum(dim=-1).detach()
        self._step(self.optimizer_c_alpha, -(self.c_log_alpha * (h_c + self.c_target_entropy)).mean())
        self._step(self.optimizer_d_alpha, -(self.d_log_alpha * (h_d + self.d_target_entropy)).mean())
        self.c_alpha, self.d_alpha = self.c_log_alpha.exp(), self.d_log_alpha.exp()

        if self.learn_iter % 100 == 0:
            self.hard_update(self.Critic, self.Critic_)
        return critic_loss.item()
