"""CPU-PyTorch port of the reference's CTR modules and loop bodies.  TEST INFRASTRUCTURE ONLY.

This is the ``cpu_baseline`` (``kind: "port"``) that ``bench.py`` times beside the
B200 path: it runs the same stock ATen CPU kernels the reference runs
(``index_select`` gathers, ``embedding_dense_backward``, dense ``torch.optim.Adam``
with L2 over every row -- SURVEY N3) on all host threads.  The reference itself
is a pure-Python tree that cannot travel to the GPU box, so its modules are
restated here; ``tests/test_oracle_golden.py`` pins this port (and the numpy
oracle) against vectors produced by the real reference modules.

state_dict keys and shapes are the reference's so golden
parameters load directly.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def _tower(in_dims: int) -> nn.Sequential:
    """[in -> 300 -> 200 -> 1] with ReLU + Dropout(0.2) (p_model.py:276-293); Sequential
    indices 0,3,6 hold the Linear layers, as in the reference's state_dict."""
    mods = []
    d = in_dims
    for width in (300, 200):
        mods += [nn.Linear(d, width), nn.ReLU(), nn.Dropout(p=0.2)]
        d = width
    mods.append(nn.Linear(d, 1))
    return nn.Sequential(*mods)


class PortCTR(nn.Module):
    """LR / FM / FFM / DeepFM of src/models/p_model.py:9-100,256-324 in one class.

    Parameter creation order follows the reference constructors so that the same
    ``torch.manual_seed`` yields the same initial values."""

    def __init__(self, kind: str, feature_nums: int, field_nums: int = 15, latent_dims: int = 10):
        super().__init__()
        assert kind in ("LR", "FM", "FFM", "DeepFM")
        self.kind, self.F, self.D = kind, field_nums, latent_dims
        self.linear = nn.Embedding(feature_nums, 1)
        self.bias = nn.Parameter(torch.zeros(1))
        if kind in ("FM", "DeepFM"):
            self.feature_embedding = nn.Embedding(feature_nums, latent_dims)
        if kind == "FFM":
            self.field_feature_embeddings = nn.ModuleList(
                nn.Embedding(feature_nums, latent_dims) for _ in range(field_nums))
        if kind == "DeepFM":
            self.mlp = _tower(field_nums * latent_dims)

    def _first_order(self, x):
        return self.bias + self.linear(x).sum(dim=1)                       # p_model.py:23

    def _fm_term(self, x):
        v = self.feature_embedding(x)                                      # :47
        sq_of_sum = v.sum(dim=1) ** 2                                      # :49
        sum_of_sq = (v ** 2).sum(dim=1)                                    # :50
        return 0.5 * (sq_of_sum - sum_of_sq).sum(dim=1, keepdim=True)      # :52,54

    def logit(self, x):
        z = self._first_order(x)
        if self.kind == "FM":
            z = z + self._fm_term(x)
        elif self.kind == "DeepFM":
            e = self.feature_embedding(x)                                  # :320 (second gather)
            z = z + self._fm_term(x) + self.mlp(e.view(-1, self.F * self.D))
        elif self.kind == "FFM":
            g = [t(x) for t in self.field_feature_embeddings]              # :87
            terms = []
            for i in range(self.F - 1):
                for j in range(i + 1, self.F):
                    terms.append(g[j][:, i] * g[i][:, j])                  # :91
            z = z + torch.stack(terms, dim=1).sum(dim=1).sum(dim=1, keepdim=True)   # :95,97
        return z

    def forward(self, x):
        return torch.sigmoid(self.logit(x))


class PortTail(nn.Module):
    """WideAndDeep / FNN / InnerPNN of src/models/p_model.py:103-200,326-373: the same gather with another dense
    tail.  Parameter creation order and state_dict keys are the reference's."""

    def __init__(self, kind: str, feature_nums: int, field_nums: int = 15, latent_dims: int = 10):
        super().__init__()
        assert kind in ("WideAndDeep", "FNN", "InnerPNN")
        self.kind, self.F, self.D = kind, field_nums, latent_dims
        if kind == "WideAndDeep":
            self.linear = nn.Embedding(feature_nums, 1)                    # p_model.py:115
            self.bias = nn.Parameter(torch.zeros(1))                       # :116
            self.embedding = nn.Embedding(feature_nums, latent_dims)       # :118
            self.mlp = _tower(field_nums * latent_dims)
        elif kind == "FNN":
            self.feature_embedding = nn.Embedding(feature_nums, latent_dims)   # :336
            self.mlp = _tower(field_nums * latent_dims)
        else:
            self.feature_embedding = nn.Embedding(feature_nums, latent_dims)   # :158
            self.mlp = _tower(field_nums * latent_dims + field_nums * (field_nums - 1) // 2)
            idx = torch.triu_indices(field_nums, field_nums, offset=1)         # :178-181 (i < j, row-major)
            self.row, self.col = idx[0].tolist(), idx[1].tolist()

    def forward(self, x):
        FD = self.F * self.D
        if self.kind == "WideAndDeep":                                     # :135-144
            e = self.embedding(x)
            return torch.sigmoid(self.bias + torch.sum(self.linear(x), dim=1) + self.mlp(e.view(-1, FD)))
        e = self.feature_embedding(x)
        if self.kind == "FNN":                                             # :365-373
            return torch.sigmoid(self.mlp(e.view(-1, FD)))
        ip = torch.sum(torch.mul(e[:, self.row], e[:, self.col]), dim=2)   # :189
        return torch.sigmoid(self.mlp(torch.cat([e.view(-1, FD), ip], dim=1)))   # :194-198


def make_port(kind: str, feature_nums: int, field_nums: int = 15, latent_dims: int = 10) -> nn.Module:
    if kind in ("WideAndDeep", "FNN", "InnerPNN"):
        return PortTail(kind, feature_nums, field_nums, latent_dims)
    return PortCTR(kind, feature_nums, field_nums, latent_dims)


class PortFeatureEmbedding(nn.Module):
    """Feature_embedding.py:31-59."""

    def __init__(self, feature_numbers: int, field_nums: int, latent_dims: int):
        super().__init__()
        self.F, self.D = field_nums, latent_dims
        self.feature_embedding = nn.Embedding(feature_numbers, latent_dims)
        idx = torch.triu_indices(field_nums, field_nums, offset=1)
        self.row, self.col = idx[0].tolist(), idx[1].tolist()

    def forward(self, x):
        e = self.feature_embedding(x)
        ip = (e[:, self.row] * e[:, self.col]).sum(dim=2)
        return torch.cat([ip, e.view(-1, self.F * self.D)], dim=1).detach()


def ctr_train_step(model, optimizer, loss_fn, features, labels):
    """Body of train() at src/main/pretrain_main.py:96-103 for one pre-built batch
    (no DataLoader, as src/all_main/pretrain_main_2.py:71-72 feeds it)."""
    y = model(features)
    loss = loss_fn(y, labels.float())
    model.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.item()


def make_adam(model, lr=1e-3, weight_decay=1e-5):
    """torch.optim.Adam exactly as built at src/main/pretrain_main.py:181."""
    return torch.optim.Adam(params=model.parameters(), lr=lr, weight_decay=weight_decay)
