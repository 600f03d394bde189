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
    """The agent (reference ``:98-259``): a rollout memory of six column blocks, acting through the current network, and ``learn``:
    TD residuals -> generalised advantages (device scan) -> ``k_epochs`` clipped-surrogate steps on one Adam optimizer."""

    _COLUMNS = (("memory_state", "field_nums"), ("memory_c_a", "action_nums"), ("memory_c_logprobs", "action_nums"),
                ("memory_d_a", 1), ("memory_d_logprobs", 1), ("memory_reward", 1))

    def __init__(self, feature_nums, field_nums=15, latent_dims=5, action_nums=2, campaign_id='1458', init_lr=1e-2, train_epochs=500,
                 reward_decay=1, lr_lamda=0.01, memory_size=4096000, batch_size=256, tau=0.005, k_epochs=3, eps_clip=0.2,
                 device='cuda:0'):
        self.feature_nums, self.field_nums, self.latent_dims, self.action_nums = feature_nums, field_nums, latent_dims, action_nums
        self.campaign_id, self.device = campaign_id, device
        self.init_lr, self.train_epochs, self.lr_lamda = init_lr, train_epochs, lr_lamda
        self.gamma, self.lamda, self.tau = reward_decay, 0.95, tau
        self.memory_size, self.batch_size, self.memory_counter = memory_size, batch_size, 0
        self.k_epochs, self.eps_clip, self.action_std = k_epochs, eps_clip, 0.5
        self.input_dims = field_nums * (field_nums - 1) // 2 + field_nums * latent_dims
        for name, width in self._COLUMNS:
            w = getattr(self, width) if isinstance(width, str) else width
            setattr(self, name, torch.zeros(memory_size, w, device=device))
        self.hybrid_actor_critic = Hybrid_Actor_Critic(self.input_dims, action_nums).to(device)
        self.hybrid_actor_critic_old = Hybrid_Actor_Critic(self.input_dims, action_nums).to(device)
        self.optimizer = _optim.Adam(self.hybrid_actor_critic.parameters(), lr=init_lr, weight_decay=1e-5)      # :150
        self.loss_func = nn.MSELoss()

    def store_memory(self, states, c_a, c_logprobs, d_a, d_logprobs, rewards):
        """Writes the rollout at ``[counter % size, counter + len)``; the reference never advances the counter (:154-170), so
        every rollout overwrites the head of the memory -- kept."""
        lo, hi = self.memory_counter % self.memory_size, self.memory_counter + len(states)
        for (name, _), block in zip(self._COLUMNS, (states, c_a, c_logprobs, d_a, d_logprobs, rewards)):
            getattr(self, name)[lo:hi, :] = block

    def memory(self):
        ids = self.memory_state.long()
        return ids, ids, self.memory_c_a, self.memory_c_logprobs, self.memory_d_a, self.memory_d_logprobs, self.memory_reward

    @torch.no_grad()
    def choose_a(self, state):
        return self.hybrid_actor_critic.eval().act(state)

    @torch.no_grad()
    def choose_best_a(self, state):
        weights, d_probs = self.hybrid_actor_critic.eval().best_a(state)
        return weights, (torch.argsort(-d_probs)[:, 0] + 2).view(-1, 1)

    def _clipped(self, logp, old_logp, adv):
        """PPO's pessimistic surrogate -E[min(rho A, clip(rho, 1-e, 1+e) A)], rho = exp(logp - old_logp)."""
        rho = torch.exp(logp - old_logp)
        return -torch.min(rho * adv, rho.clamp(1 - self.eps_clip, 1 + self.eps_clip) * adv).mean()

    def learn(self, states, states_, old_c_a, old_c_a_logprobs, old_d_a, old_d_a_logprobs, rewards):
        """:194-258.  Returns the loss of the last epoch."""
        net = self.hybrid_actor_critic
        d_taken = old_d_a.long()
        v_next = net.evaluate(states_, old_c_a, d_taken)[0]
        v_now = net.evaluate(states, old_c_a, d_taken)[0]
        target = rewards + self.gamma * v_next
        adv = gae_advantages(target - v_now, self.gamma * self.lamda)             # :206-209 without the per-sample host sync
        adv = (adv - adv.mean()) / (adv.std() + 1e-5)
        target = target.detach()
        last = 0
        for _ in range(self.k_epochs):
            value, c_logp, c_ent, d_logp, d_ent = net.evaluate(states, old_c_a, d_taken)
            loss = (self._clipped(c_logp, old_c_a_logprobs, adv) - 0.01 * c_ent.mean()
                    + self._clipped(d_logp, old_d_a_logprobs, adv) - 0.01 * d_ent.mean()
                    + 0.5 * self.loss_func(value, target))
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
            last = loss.mean().item()
        self.hybrid_actor_critic_old.load_state_dict(net.state_dict())
        return last
