"""Replay memories sampled on the device.

``Memory`` mirrors the prioritized memory of the reference's TD3 / hybrid agents
(``src/models/v10_Hybrid_TD3_model_PER.py:19-110``: same constructor, attributes ``memory``, ``prioritys_``,
``memory_counter``, ``epsilon / alpha / beta``, methods ``get_priority / add / stochastic_sample / greedy_sample /
batch_update``).  What differs is where the sampling runs: the reference copies every priority to the host and calls
``np.random.choice(n, batch, p=P, replace=False)`` (:69-77); here ``rlctr_replay_sample_per`` draws the same distribution
(weighted sampling without replacement) on the device, stream-ordered, without a host round trip.  ``sample_uniform`` is the
device counterpart of ``random.sample(range(n), batch)`` (``DDQN_model.py:183-185``, ``DDPG_for_PG_model.py``).

The indices are random, so they cannot equal the host RNG's draw; parity is distributional (tests: inclusion frequencies against
the exact successive-sampling probabilities) and exact for everything downstream of the indices (gathered rows, IS weights).
"""
from __future__ import annotations

import torch

from . import _lib


def _rng(device, seed=None):
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())          # CPU default generator: follows torch.manual_seed
    return torch.tensor([int(seed), 0], dtype=torch.int64, device=device)


def sample_uniform(n_valid: int, batch_size: int, rng: torch.Tensor) -> torch.Tensor:
    """``batch_size`` distinct indices of ``range(n_valid)`` (random.sample semantics) as an int64 device tensor."""
    lib = _lib.load()
    if batch_size > n_valid:
        raise ValueError("Sample larger than population or is negative")      # random.sample's message
    out = torch.empty(batch_size, dtype=torch.int64, device=rng.device)
    _lib.call("rlctr_replay_sample_uniform", lib.rlctr_replay_sample_uniform, int(n_valid), int(batch_size), _lib.ptr(rng),
              _lib.ptr(out), _lib.stream(), meta={"n": n_valid, "batch": batch_size})
    _lib.check(lib.rlctr_rng_advance(_lib.ptr(rng), 1, _lib.stream()), "rlctr_rng_advance")
    return out


class Memory(object):
    """v10_Hybrid_TD3_model_PER.py:19-110 on the device."""

    def __init__(self, memory_size, transition_lens, device, seed=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RlctrError("rl_ctr_prediction_b200.replay.Memory lives on a CUDA (sm_100a) device; there is no CPU fallback")
        self.transition_lens = transition_lens
        self.epsilon = 1e-3
        self.alpha = 0.6
        self.beta = 1.0
        self.beta_increment_per_sampling = 1e-4
        self.abs_err_upper = 1
        self.memory_size = memory_size
        self.memory_counter = 0
        self.prioritys_ = torch.zeros(size=[memory_size, 2], device=self.device)
        self.memory = torch.zeros(size=[memory_size, transition_lens], device=self.device)
        self._rng = _rng(self.device, seed)
        self._ws = None

    def get_priority(self, td_error):                                            # :40-41
        return torch.pow(torch.abs(td_error) + self.epsilon, self.alpha)

    def add(self, td_error, transitions):                                        # :43-60
        lib = _lib.load()
        n = len(transitions)
        tr = transitions.to(self.device, torch.float32).contiguous()
        p = td_error.to(self.device, torch.float32).expand(n, 2).contiguous() if td_error.shape[-1] != 2 else \
            td_error.to(self.device, torch.float32).contiguous()
        st = _lib.stream()
        _lib.check(lib.rlctr_replay_store(_lib.ptr(self.memory), self.memory_size, self.transition_lens, self.memory_counter,
                                          _lib.ptr(tr), n, self.transition_lens, st), "rlctr_replay_store")
        _lib.check(lib.rlctr_replay_store(_lib.ptr(self.prioritys_), self.memory_size, 2, self.memory_counter, _lib.ptr(p), n, 2, st),
                   "rlctr_replay_store")
        self.memory_counter += n

    def _valid(self):
        return self.memory_size if self.memory_counter >= self.memory_size else self.memory_counter

    def _sample(self, batch_size, greedy, sample=None):
        lib = _lib.load()
        n = self._valid()
        if sample is not None:                          # injected indices (reproducible parity runs): everything downstream as usual
            idx = torch.as_tensor(sample, device=self.device).long().reshape(-1).contiguous()
            pri = self.get_priority(self.prioritys_[:n, 0:1])
            isw = torch.pow(torch.div(pri[idx], torch.min(pri)), -self.beta)                         # :80-82
            batch = torch.empty(idx.numel(), self.transition_lens, dtype=torch.float32, device=self.device)
            _lib.check(lib.rlctr_replay_gather(_lib.ptr(self.memory), self.transition_lens, _lib.ptr(idx), idx.numel(),
                                               _lib.ptr(batch), _lib.stream()), "rlctr_replay_gather")
            return idx, batch, isw
        if batch_size > n:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")     # numpy's message
        wsb = lib.rlctr_replay_per_ws_bytes(n)
        if self._ws is None or self._ws.numel() < wsb:
            self._ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
        idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
        isw = torch.empty(batch_size, 1, dtype=torch.float32, device=self.device)
        _lib.call("rlctr_replay_sample_per", lib.rlctr_replay_sample_per, _lib.ptr(self.prioritys_), 2, n, float(self.epsilon),
                  float(self.alpha), float(self.beta), 1 if greedy else 0, int(batch_size), _lib.ptr(self._rng), _lib.ptr(idx),
                  _lib.ptr(isw), _lib.ptr(self._ws), self._ws.numel(), _lib.stream(), meta={"n": n, "batch": batch_size})
        if not greedy:
            _lib.check(lib.rlctr_rng_advance(_lib.ptr(self._rng), n, _lib.stream()), "rlctr_rng_advance")
        batch = torch.empty(batch_size, self.transition_lens, dtype=torch.float32, device=self.device)
        _lib.check(lib.rlctr_replay_gather(_lib.ptr(self.memory), self.transition_lens, _lib.ptr(idx), batch_size, _lib.ptr(batch),
                                           _lib.stream()), "rlctr_replay_gather")
        return idx, batch, isw

    def stochastic_sample(self, batch_size, sample=None):                        # :62-85
        return self._sample(batch_size, False, sample)

    def greedy_sample(self, batch_size):                                         # :87-105 (top-batch of column 0)
        self.beta = torch.min(torch.FloatTensor([1., self.beta + self.beta_increment_per_sampling])).item()   # fp32, as the reference      # :94
        return self._sample(batch_size, True)

    def batch_update(self, choose_idx, td_errors):                               # :107-108
        lib = _lib.load()
        td = td_errors.to(self.device, torch.float32).reshape(-1).contiguous()
        idx = choose_idx.to(self.device, torch.int64).reshape(-1).contiguous()
        _lib.check(lib.rlctr_replay_update(_lib.ptr(self.prioritys_), 2, _lib.ptr(idx), _lib.ptr(td), idx.numel(), _lib.stream()),
                   "rlctr_replay_update")
