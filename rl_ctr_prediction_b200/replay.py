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


class PrioritizedBuffer(object):
    """Ring buffer of transitions with one priority per slot, sampled on the device.  The reference has two flavours of the
    same class, which differ in WHEN (|td| + eps)^alpha is applied and in the priority a new transition gets:

    ===========================  =============================================  ============================================
    flavour                      ``v10_Hybrid_TD3_model_PER.Memory`` (:19-110)   ``Hybrid_SAC_model.Memory`` (:21-108)
    ===========================  =============================================  ============================================
    priority array               ``prioritys_`` [size, 2] (raw td | label)       ``priorities_`` [size, 1] (already ^alpha)
    ``add``                      caller passes the [n, 2] rows                   new slots get max(old, 1)
    sampling weight              (|raw| + eps)^alpha                             the stored value
    ``batch_update``             stores the raw td error                         stores (|td| + eps)^alpha
    beta                         1.0, grows in ``greedy_sample`` only            0.4, grows in both samplers
    ===========================  =============================================  ============================================

    Subclasses set ``_COLS`` / ``_RAW`` and the attribute name the reference uses for the priority array."""

    _COLS, _RAW, _PRIO_NAME, _BETA0, _BETA_INC = 2, True, "prioritys_", 1.0, 1e-4

    def __init__(self, memory_size, transition_lens, device, seed=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.RlctrError(f"{type(self).__module__}.{type(self).__name__} lives on a CUDA (sm_100a) device; "
                                  "there is no CPU fallback")
        self.memory_size, self.transition_lens, self.memory_counter = memory_size, transition_lens, 0
        self.epsilon, self.alpha, self.abs_err_upper = 1e-3, 0.6, 1
        self.beta, self.beta_increment_per_sampling = self._BETA0, self._BETA_INC
        setattr(self, self._PRIO_NAME, torch.zeros(memory_size, self._COLS, device=self.device))
        self.memory = torch.zeros(memory_size, transition_lens, device=self.device)
        self._rng = _rng(self.device, seed)
        self._ws = None

    # ---- pieces ----------------------------------------------------------------------------------------------------
    @property
    def _prio(self):
        return getattr(self, self._PRIO_NAME)

    def get_priority(self, td_error):
        return torch.pow(torch.abs(td_error) + self.epsilon, self.alpha)

    def _valid(self):
        return min(self.memory_counter, self.memory_size)

    def _bump_beta(self):             # the reference rounds through a FloatTensor
        self.beta = torch.min(torch.FloatTensor([1., self.beta + self.beta_increment_per_sampling])).item()

    def _ring_write(self, dst, width, rows):
        lib = _lib.load()
        rows = rows.to(self.device, torch.float32).contiguous()
        _lib.check(lib.rlctr_replay_store(_lib.ptr(dst), self.memory_size, width, self.memory_counter, _lib.ptr(rows), len(rows),
                                          width, _lib.stream()), "rlctr_replay_store")

    def _rows_of(self, idx):
        lib = _lib.load()
        out = torch.empty(idx.numel(), self.transition_lens, dtype=torch.float32, device=self.device)
        _lib.check(lib.rlctr_replay_gather(_lib.ptr(self.memory), self.transition_lens, _lib.ptr(idx), idx.numel(), _lib.ptr(out),
                                           _lib.stream()), "rlctr_replay_gather")
        return out

    def _weights_of(self, idx, greedy=False):
        """(p_i / min_j p_j)^(-beta) over the valid range; p = the sampling weight (raw priority for the greedy sampler)."""
        col = self._prio[:self._valid(), 0:1]
        p = self.get_priority(col) if (self._RAW and not greedy) else col
        return torch.pow(torch.div(p[idx], torch.min(p)), -self.beta)

    def _draw(self, batch_size, greedy):
        lib = _lib.load()
        n = self._valid()
        if batch_size > n:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")     # numpy's message
        need = lib.rlctr_replay_per_ws_bytes(n)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        idx = torch.empty(batch_size, dtype=torch.int64, device=self.device)
        isw = torch.empty(batch_size, 1, dtype=torch.float32, device=self.device)
        eps, alpha = (float(self.epsilon), float(self.alpha)) if self._RAW else (0.0, 1.0)    # stored value IS the weight
        _lib.call("rlctr_replay_sample_per", lib.rlctr_replay_sample_per, _lib.ptr(self._prio), self._COLS, n, eps, alpha,
                  float(self.beta), 1 if greedy else 0, int(batch_size), _lib.ptr(self._rng), _lib.ptr(idx), _lib.ptr(isw),
                  _lib.ptr(self._ws), self._ws.numel(), _lib.stream(), meta={"n": n, "batch": batch_size})
        if not greedy:
            _lib.check(lib.rlctr_rng_advance(_lib.ptr(self._rng), n, _lib.stream()), "rlctr_rng_advance")
        return idx, isw                                  # IS weights (p_i / min p)^(-beta) from the same kernel

    # ---- the reference's surface -----------------------------------------------------------------------------------
    def stochastic_sample(self, batch_size, sample=None):
        """Weighted sampling without replacement (the reference: a D2H copy of every priority + ``np.random.choice``).
        ``sample`` = injected indices for reproducible parity runs."""
        if not self._RAW:
            self._bump_beta()                            # the SAC flavour raises beta before it weighs the sample (:77-81)
        if sample is not None:
            idx = torch.as_tensor(sample, device=self.device).long().reshape(-1).contiguous()
            return idx, self._rows_of(idx), self._weights_of(idx)
        idx, isw = self._draw(batch_size, False)
        return idx, self._rows_of(idx), isw

    def greedy_sample(self, batch_size):
        self._bump_beta()
        idx, isw = self._draw(batch_size, True)
        return idx, self._rows_of(idx), isw

    def batch_update(self, choose_idx, td_errors):
        lib = _lib.load()
        td = td_errors.to(self.device, torch.float32)
        val = (td if self._RAW else self.get_priority(td)).reshape(-1).contiguous()
        idx = choose_idx.to(self.device, torch.int64).reshape(-1).contiguous()
        _lib.check(lib.rlctr_replay_update(_lib.ptr(self._prio), self._COLS, _lib.ptr(idx), _lib.ptr(val), idx.numel(),
                                           _lib.stream()), "rlctr_replay_update")


class Memory(PrioritizedBuffer):
    """``v10_Hybrid_TD3_model_PER.Memory``: raw td errors stored beside the click label; ``add(td_error [n, 2], transitions)``."""

    def add(self, td_error, transitions):
        n = len(transitions)
        rows = td_error.to(self.device, torch.float32)
        if rows.shape[-1] != self._COLS:
            rows = rows.expand(n, self._COLS)
        self._ring_write(self.memory, self.transition_lens, transitions)
        self._ring_write(self._prio, self._COLS, rows)
        self.memory_counter += n
