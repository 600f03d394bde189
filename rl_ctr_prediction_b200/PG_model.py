"""REINFORCE policy -- drop-in for the reference's ``src/models/PG_model.py`` (``Net`` :24-58,
``PolicyGradient`` :60-179), with the input-dimension fix the reference needs to run at all.

Divergence from the reference, stated once: ``PG_model.Net`` sizes its first Linear for
``F(F-1)/2*D + F*D`` inputs (:34-39), the layout of an older Feature_Embedding, while the live
``Feature_Embedding`` emits ``F(F-1)/2 + F*D`` (Feature_embedding.py:54-57) -- the reference raises a
shape error on the first forward.  Here ``input_dims`` is what the encoder really emits (255 for
F=15, D=10).  Everything else is literal, including the loss ``(sum_b -log pi_b) * mean_b(vt_b)``
(:104-107; ``loss_variant='literal'``); ``loss_variant='per_sample'`` is the textbook
``mean_b(-log pi_b * vt_b)``.

The state encoder is rlctr_featemb_fwd, the softmax / log-prob / loss head and its gradient are
rlctr_reinforce_loss_bwd (one kernel pair instead of softmax, gather, log, sum, mul, mean and
their autograd), the Linear layers are :class:`.mlp.Linear` and the update is :class:`.optim.Adam`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import mlp as _mlp
from . import optim as _optim
from .Feature_embedding import Feature_Embedding

RF_LITERAL, RF_PER_SAMPLE = 0, 1


class _ReinforceHead(torch.autograd.Function):
    """logits[B,A], acts[B] in 1..A, vt[B] -> (loss scalar, logp[B]); grad flows to logits only."""

    @staticmethod
    def forward(ctx, logits, acts, vt, variant, ws):
        lib = _lib.load()
        logits = logits.contiguous()
        B, A = logits.shape
        dev = logits.device
        acts = acts.reshape(-1).long().contiguous()
        vt = vt.reshape(-1).float().contiguous()
        logp = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dlogits = torch.empty(B, A, dtype=torch.float32, device=dev)
        _lib.check(lib.rlctr_reinforce_loss_bwd(_lib.ptr(logits), _lib.ptr(acts), _lib.ptr(vt), _lib.ptr(logp),
                                                _lib.ptr(loss), _lib.ptr(dlogits), _lib.ptr(ws), B, A, variant,
                                                _lib.stream()), "rlctr_reinforce_loss_bwd")
        ctx.save_for_backward(dlogits)
        ctx.mark_non_differentiable(logp)
        return loss.reshape(()), logp

    @staticmethod
    def backward(ctx, gloss, _glogp):
        (dlogits,) = ctx.saved_tensors
        return dlogits * gloss, None, None, None, None


class Net(nn.Module):
    """PG_model.py:24-58: Feature_Embedding -> [in,1024,512,256,128,A] MLP (ReLU, Dropout .2) -> softmax."""

    def __init__(self, field_nums, feature_nums, latent_dims, action_numbers, campaign_id=None, device=None):
        super().__init__()
        self.field_nums, self.feature_nums, self.latent_dims = field_nums, feature_nums, latent_dims
        self.campaign_id = campaign_id
        self.embedding_layer = Feature_Embedding(feature_nums, field_nums, latent_dims, device=device)
        input_dims = self.embedding_layer.output_dims                  # N9 fix: what the encoder emits
        self.input_dims = input_dims
        layers, width = [], 1024
        for _ in range(4):
            layers += [_mlp.Linear(input_dims, width, device=device), nn.ReLU(), nn.Dropout(p=0.2)]
            input_dims, width = width, width // 2
        layers.append(_mlp.Linear(input_dims, action_numbers, device=device))
        self.mlp = _mlp.Tower(*layers)

    def logits(self, x):
        return self.mlp(self.embedding_layer.forward(x))

    def forward(self, x):
        return torch.softmax(self.logits(x), dim=1)                    # :56


class PolicyGradient:
    def __init__(self, feature_nums, field_nums, latent_dims, campaign_id=None, action_nums=2,
                 learning_rate=1e-4, reward_decay=1, device="cuda:0", loss_variant="literal"):
        self.action_nums, self.feature_nums, self.field_nums = action_nums, feature_nums, field_nums
        self.latent_dims, self.lr, self.gamma, self.device = latent_dims, learning_rate, reward_decay, device
        self.campaign_id = campaign_id
        self.variant = RF_LITERAL if loss_variant == "literal" else RF_PER_SAMPLE
        self._clear()
        self.policy_net = Net(field_nums, feature_nums, latent_dims, action_nums, campaign_id).to(device)
        self.optimizer = _optim.Adam(self.policy_net.parameters(), lr=self.lr, weight_decay=1e-5)   # :87
        self._ws = torch.zeros(_lib.RLCTR_REDUCE_WS_BYTES, dtype=torch.uint8, device=device)
        self.last_logp = None

    def _clear(self):
        self.ep_states = torch.empty(0, dtype=torch.long, device=self.device)
        self.ep_as = torch.empty(0, dtype=torch.long, device=self.device)
        self.ep_rs = torch.empty(0, dtype=torch.float32, device=self.device)

    def load_embedding(self, pretrain_params):
        self.policy_net.embedding_layer.load_embedding(pretrain_params)

    def loss_func(self, logits, acts, vt):
        """:104-107 on logits (softmax folded into the head kernel).  Returns the loss; the per-sample
        policy log-probs are left in ``self.last_logp``."""
        loss, self.last_logp = _ReinforceHead.apply(logits, acts, vt, self.variant, self._ws)
        return loss

    def choose_action(self, states):
        """:110-121 (host RNG exactly as the reference: CPU rand / randint)."""
        with torch.no_grad():
            prob_weights = self.policy_net.forward(states).cpu()
        random_seeds = torch.rand(len(states), 1)
        max_action = torch.argsort(-prob_weights)[:, 0] + 1
        random_action = torch.randint(low=1, high=self.action_nums + 1, size=[len(states), 1])
        actions = torch.where(random_seeds >= torch.max(prob_weights, 1)[0].view(-1, 1), max_action.view(-1, 1),
                              random_action)
        return actions.to(self.device)

    def choose_best_action(self, state):
        with torch.no_grad():
            prob_weights = self.policy_net.forward(state)
        return torch.max(prob_weights, 1)[1].view(-1, 1) + 1

    def store_transition(self, s, a, r):
        self.ep_states = torch.cat([self.ep_states, s], dim=0)
        self.ep_as = torch.cat([self.ep_as, a], dim=0)
        self.ep_rs = torch.cat([self.ep_rs, r.reshape(-1).float()], dim=0)

    def discount_and_norm_rewards(self):
        """:139-154: float64 suffix returns over the stored episode, then (G - mean) / std.  The
        reference walks the episode in a Python loop; here it is a reversed inclusive scan (numpy
        float64 on the host copy, as the reference also computes it on the host)."""
        rs = self.ep_rs.detach().cpu().numpy().astype(np.float64)
        if self.gamma == 1:
            g = np.cumsum(rs[::-1])[::-1].copy()
        else:
            g = np.zeros_like(rs)
            run = 0.0
            for i in range(len(rs) - 1, -1, -1):
                run = run * self.gamma + rs[i]
                g[i] = run
        g -= np.mean(g)
        g /= np.std(g)
        return g

    def learn(self, vt=None):
        """:156-179.  ``vt`` overrides the normalised returns (tests use raw returns)."""
        if vt is None:
            vt = torch.as_tensor(self.discount_and_norm_rewards(), dtype=torch.float32).to(self.device)
        self.policy_net.train()
        logits = self.policy_net.logits(self.ep_states)
        loss = self.loss_func(logits, self.ep_as, vt)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        self._clear()
        return loss

    # ------------------------------------------------------------------------------------------
    def returns_on_device(self, rewards):
        """``discount_and_norm_rewards`` (:139-154) without leaving the device: the float64 suffix returns
        ``G_i = r_i + gamma * G_{i+1}`` are one fp64 scan (rlctr_gae_scan, the kernel behind the PPO agent's advantage
        recurrence), then ``(G - mean) / std`` (population std, as ``np.std``) and the cast to fp32."""
        import ctypes as C
        lib = _lib.load()
        r = rewards.reshape(-1).float().contiguous()
        n = r.numel()
        out = torch.empty(n, dtype=torch.float32, device=r.device)
        ws_bytes = lib.rlctr_gae_ws_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=r.device)
        _lib.check(lib.rlctr_gae_scan(_lib.ptr(r), n, float(self.gamma), _lib.ptr(out), _lib.ptr(ws), ws_bytes, _lib.stream()),
                   "rlctr_gae_scan")
        g = out.flip(0).double()                       # the scan writes G of sample n-1-i at i (Hybrid_PPO_model.py:206-212)
        g = (g - g.mean()) / g.std(unbiased=False)
        return g.float()

    def fused_step(self, features, labels, model_dict, prob_weights=None, actions=None):
        """BASELINE.json configs[2]: the REINFORCE update fused into ONE training step over a batch, nothing leaving the
        device.  The policy's logits serve both the action draw (``choose_action`` :110-121, device RNG) and the loss;
        the action a in 1..A is the number of CTR models to ensemble minus one (k = a + 1 in 2..M, the action space of
        ``src/all_main/main.py:285``); ``generate_preds`` (:183-271) scores the M frozen models and returns the +-1
        reward; the batch is the episode: returns :139-154, loss :104-107, Adam :87.  Returns (loss, rewards, actions).

        The unfused sequence -- ``choose_action``, ``generate_preds``, ``store_transition``, ``learn`` -- gives the same
        update for the same actions (tests/test_gpu_policy.py::test_reinforce_fused_step_equals_unfused)."""
        from .ensemble import generate_preds
        x = features.long()
        self.policy_net.train()
        logits = self.policy_net.logits(x)
        B, M = x.shape[0], len(model_dict)
        if self.action_nums != M - 1:
            raise ValueError(f"fused_step: the policy picks k in 2..M, so action_nums must be M - 1 = {M - 1}")
        with torch.no_grad():
            if actions is None:
                pw = torch.softmax(logits.detach(), dim=1)
                seeds = torch.rand(B, 1, device=x.device)
                rand_a = torch.randint(1, self.action_nums + 1, (B, 1), device=x.device)
                mx, arg = torch.max(pw, 1)
                actions = torch.where(seeds >= mx.view(-1, 1), arg.view(-1, 1) + 1, rand_a)
            if prob_weights is None:                   # no weight-emitting actor in this configuration: equal weights
                prob_weights = torch.full((B, M), 1.0 / M, device=x.device)
            _, _, rewards = generate_preds(model_dict, x, actions + 1, prob_weights, labels, x.device, "train")
            vt = self.returns_on_device(rewards)
        loss = self.loss_func(logits, actions, vt)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss, rewards, actions
