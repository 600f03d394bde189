"""Double DQN agent -- drop-in for the reference's ``src/models/DDQN_model.py`` (``Net`` :20-52,
``DoubleDQN`` :55-224) as used by ``src/all_main/main.py``.

Same constructor arguments, methods and numerics: the Q network is ``[255 -> 300 -> 300 -> 300 -> M-1]``
with BatchNorm1d + ReLU after each hidden Linear; the discrete action is the NUMBER of CTR models to
ensemble, ``argsort(-Q)[:, 0] + 2`` in ``{2..M}`` (:154-155); epsilon-greedy draws come from torch's CPU
generator exactly as in the reference (:153,155); replay is a float32 ring buffer on the device (:92) sampled with
Python's ``random.sample`` (:183-185); ``learn`` is double-Q with MSE, hard target copy every
``replace_target_iter`` calls (:198-224).

What runs on the B200 path: every Linear is :class:`.mlp.Linear` (tcgen05 3xTF32 GEMMs), the update is
:class:`.optim.Adam` (fused dense-Adam kernel).  BatchNorm1d + ReLU between the GEMMs: in training mode (the learn steps) one
library kernel each way (rlctr_bn_relu_fwd / _bwd: batch statistics, normalisation, affine map, ReLU, running statistics); in
eval mode (the acting path, :148-151) folded into the GEMM (mlp.Tower).
"""
from __future__ import annotations

import random

import torch
import torch.nn as nn

from . import mlp as _mlp
from . import optim as _optim


def state_dims(field_nums, latent_dims):
    """Width of the Feature_Embedding state: F(F-1)/2 pair dots + F*D flattened rows (DDQN_model.py:27)."""
    return field_nums * (field_nums - 1) // 2 + field_nums * latent_dims


def bn_mlp(in_dims, out_dims, hidden=(300, 300, 300), device=None):
    """Linear -> BatchNorm1d -> ReLU per hidden layer, then Linear (DDQN_model.py:32-46; DDPG_for_PG_model.py:27-40)."""
    layers, d = [], in_dims
    for width in hidden:
        layers += [_mlp.Linear(d, width, device=device), nn.BatchNorm1d(width, device=device), nn.ReLU()]
        d = width
    layers.append(_mlp.Linear(d, out_dims, device=device))
    return _mlp.Tower(*layers)


class Net(nn.Module):
    def __init__(self, field_nums, feature_nums, latent_dims, action_nums, device=None):
        super().__init__()
        self.field_nums, self.feature_nums, self.latent_dims = field_nums, feature_nums, latent_dims
        self.input_dims = state_dims(field_nums, latent_dims)
        self.mlp = bn_mlp(self.input_dims, action_nums, device=device)

    def forward(self, input):
        return self.mlp(input)


class RingMemory:
    """float32 ring buffer on the device with the reference's wrap-around write (DDQN_model.py:105-120)."""

    def __init__(self, size, widths, device):
        self.size, self.counter = size, 0
        self.bufs = [torch.zeros(size, w, device=device) for w in widths]

    def store(self, *cols):
        n = len(cols[0])
        start, end = self.counter % self.size, (self.counter + n) % self.size
        for buf, c in zip(self.bufs, cols):
            if end > start:
                buf[start:end] = c
            else:
                first = self.size - start
                buf[start:] = c[:first]
                buf[:n - first] = c[first:n]
        self.counter += n

    device_sampling = False        # True: draw the indices on the device (replay.sample_uniform) instead of with Python's RNG

    def sample_index(self, batch_size, device):
        pool = self.size if self.counter > self.size else self.counter
        if self.device_sampling:
            from . import replay
            if getattr(self, "_rng", None) is None:
                self._rng = replay._rng(self.bufs[0].device)
            return replay.sample_uniform(pool, batch_size, self._rng)                 # no host round trip
        return torch.LongTensor(random.sample(range(pool), batch_size)).to(device)     # host RNG, as the reference


class DoubleDQN:
    device_rng = False      # True: epsilon-greedy draws on the device (no CPU generator + H2D copy per batch); False: as the reference
    grad_sync = None        # data-parallel training: callable(parameters) that all-reduces the gradients (all_main.make_data_parallel)

    def __init__(self, feature_nums, field_nums, latent_dims, campaign_id="1458", action_nums=3, learning_rate=1e-3,
                 reward_decay=1, replace_target_iter=30, memory_size=300, batch_size=32, device="cuda:0"):
        self.action_nums = action_nums - 1                       # :70  Q-values for actions 2..M
        self.feature_nums, self.field_nums, self.latent_dims = feature_nums, field_nums, latent_dims
        self.lr, self.gamma, self.replace_target_iter = learning_rate, reward_decay, replace_target_iter
        self.memory_size, self.batch_size, self.device, self.campaign_id = memory_size, batch_size, device, campaign_id
        self.learn_step_counter = 0
        self._mem = RingMemory(memory_size, [field_nums + 2], device)          # [features | action | reward] :92
        self.eval_net = Net(field_nums, feature_nums, latent_dims, self.action_nums).to(device)
        self.target_net = Net(field_nums, feature_nums, latent_dims, self.action_nums).to(device)
        self.optimizer = _optim.Adam(self.eval_net.parameters(), lr=self.lr, weight_decay=1e-5)     # :99
        self.loss_func = nn.MSELoss()

    @property
    def memory(self):
        return self._mem.bufs[0]

    @property
    def memory_counter(self):
        return self._mem.counter

    def store_transition(self, transitions):
        self._mem.store(transitions)

    def _q_eval_mode(self, states):
        self.eval_net.eval()
        with torch.no_grad():
            q = self.eval_net.forward(states)
        return q

    def choose_action(self, states, exploration_rate):
        """:144-161."""
        action_values = self._q_eval_mode(states)
        self.eval_net.train()
        rdev = self.device if self.device_rng else None
        random_seeds = torch.rand(len(states), 1, device=rdev).to(self.device)
        # argsort(-Q)[:, 0] (:154) is the arg max of the row; a full sort of every row (0.7 ms at a 1M batch) is not needed for it
        # (rows with exactly tied Q-values have no defined winner in the reference either: its sort is unstable)
        max_action = torch.argmax(action_values, dim=1) + 2
        random_action = torch.randint(low=2, high=self.action_nums + 2, size=[len(states), 1], device=rdev).to(self.device)
        return torch.where(random_seeds >= exploration_rate, max_action.view(-1, 1), random_action)

    def choose_best_action(self, states):
        """:164-172 (leaves the net in eval mode, like the reference)."""
        action_values = self._q_eval_mode(states)
        return (torch.argmax(action_values, dim=1) + 2).view(-1, 1)

    def soft_update(self, net, net_target):
        with torch.no_grad():
            for pt, p in zip(net_target.parameters(), net.parameters()):
                pt.mul_(1.0 - 0.001).add_(p, alpha=0.001)

    def sample_batch(self):
        """:179-196 (ids come back from the float32 buffer with .long(): exact below 2^24)."""
        idx = self._mem.sample_index(self.batch_size, self.device)
        batch = self.memory[idx, :].long()
        F = self.field_nums
        b_s = batch[:, :F]
        return b_s, batch[:, F:F + 1], batch[:, F + 1].view(-1, 1).float(), b_s

    def learn(self, b_s, b_a, b_r, b_s_):
        """:198-224: double-Q target, MSE, Adam."""
        if self.learn_step_counter % self.replace_target_iter == 0:
            self.target_net.load_state_dict(self.eval_net.state_dict())
        self.learn_step_counter += 1
        q_eval = self.eval_net.forward(b_s).gather(1, b_a - 2)
        q_next = self.target_net.forward(b_s_).detach()
        q_eval_next = self.eval_net.forward(b_s_)
        max_b_a_next = torch.unsqueeze(torch.max(q_eval_next, 1)[1], 1)
        q_target = b_r + self.gamma * q_next.gather(1, max_b_a_next)
        loss = self.loss_func(q_eval, q_target)
        self.optimizer.zero_grad()
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync(self.eval_net.parameters())
        self.optimizer.step()
        return loss
