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
    """WideAndDeep / FNN / InnerPNN / OuterPNN / DCN / AFM of src/models/p_model.py:103-254,326-485: the same gather with
    another dense tail.  Parameter creation order and state_dict keys are the reference's.

    Two departures, both needed to run at all on a CPU and both stated in the tests: OuterPNN's constant ``kernel`` is
    created on the module's device instead of ``.cuda()`` (p_model.py:236), and AFM's two always-on
    ``F.dropout`` calls (:477,479) take ``dropout_p`` (0.2 like the reference; 0 for deterministic parity
    runs) and optional explicit masks."""

    def __init__(self, kind: str, feature_nums: int, field_nums: int = 15, latent_dims: int = 10):
        super().__init__()
        assert kind in TAIL_KINDS
        self.kind, self.F, self.D = kind, field_nums, latent_dims
        if kind == "WideAndDeep":
            self.linear = nn.Embedding(feature_nums, 1)                    # p_model.py:115
            self.bias = nn.Parameter(torch.zeros(1))                       # :116
            self.embedding = nn.Embedding(feature_nums, latent_dims)       # :118
            self.mlp = _tower(field_nums * latent_dims)
        elif kind == "FNN":
            self.feature_embedding = nn.Embedding(feature_nums, latent_dims)   # :336
            self.mlp = _tower(field_nums * latent_dims)
        elif kind == "InnerPNN":
            self.feature_embedding = nn.Embedding(feature_nums, latent_dims)   # :158
            self.mlp = _tower(field_nums * latent_dims + field_nums * (field_nums - 1) // 2)
            idx = torch.triu_indices(field_nums, field_nums, offset=1)         # :178-181 (i < j, row-major)
            self.row, self.col = idx[0].tolist(), idx[1].tolist()
        elif kind == "OuterPNN":
            self.feature_embedding = nn.Embedding(feature_nums, latent_dims)   # :214
            self.mlp = _tower(latent_dims + field_nums * latent_dims)          # :217-232
            self.kernel = torch.ones((latent_dims, latent_dims))               # :236 (not a parameter / buffer)
        elif kind == "DCN":
            fd = field_nums * latent_dims
            self.feature_embedding = nn.Embedding(feature_nums, latent_dims)   # :387
            mods, d = [], fd
            for width in (300, 200):                                           # :396-400: ends in ReLU + Dropout
                mods += [nn.Linear(d, width), nn.ReLU(), nn.Dropout(p=0.2)]
                d = width
            self.DN = nn.Sequential(*mods)
            self.num_neural_layers = 5                                         # :394
            self.cross_net_w = nn.ModuleList([nn.Linear(fd, 1, bias=False) for _ in range(5)])          # :409-411
            self.cross_net_b = nn.ParameterList([nn.Parameter(torch.zeros((fd,))) for _ in range(5)])   # :415-417
            self.linear = nn.Linear(200 + fd, 1)                               # :419
        else:                                                                  # AFM :438-470
            self.feature_embedding = nn.Embedding(feature_nums, latent_dims)
            idx = torch.triu_indices(field_nums, field_nums, offset=1)
            self.row, self.col = idx[0].tolist(), idx[1].tolist()
            self.attention_net = nn.Linear(latent_dims, latent_dims)
            self.attention_softmax = nn.Linear(latent_dims, 1)
            self.fc = nn.Linear(latent_dims, 1)
            self.linear = nn.Embedding(feature_nums, 1)
            self.bias = nn.Parameter(torch.zeros((1,)))
            self.dropout_p = 0.2

    def forward(self, x, masks=None):
        FD = self.F * self.D
        if self.kind == "AFM":                                             # :472-485
            e = self.feature_embedding(x)
            ip = torch.mul(e[:, self.row], e[:, self.col])
            sc = torch.relu(self.attention_net(ip))
            sc = torch.softmax(self.attention_softmax(sc), dim=1)
            P = len(self.row)
            if masks is not None:                                          # explicit masks: [B, P + D] multipliers
                sc = sc * masks[:, :P].unsqueeze(2)
            else:
                sc = torch.nn.functional.dropout(sc, p=self.dropout_p)
            out = torch.sum(torch.mul(sc, ip), dim=1)
            out = out * masks[:, P:] if masks is not None else torch.nn.functional.dropout(out, p=self.dropout_p)
            return torch.sigmoid(self.bias + torch.sum(self.linear(x), dim=1) + self.fc(out))
        if self.kind == "DCN":                                             # :421-435
            e = self.feature_embedding(x).view(-1, FD)
            x0, xl = e, e
            for i in range(self.num_neural_layers):
                xl = x0 * self.cross_net_w[i](xl) + self.cross_net_b[i] + xl
            return torch.sigmoid(self.linear(torch.cat([xl, self.DN(e)], dim=1)))
        if self.kind == "OuterPNN":                                        # :238-254
            e = self.feature_embedding(x)
            se = torch.sum(e, dim=1).unsqueeze(1)
            cross = torch.sum(torch.mul(torch.mul(se, self.kernel), se), dim=1)
            return torch.sigmoid(self.mlp(torch.cat([e.view(-1, FD), cross], dim=1)))
        if self.kind == "WideAndDeep":                                     # :135-144
            e = self.embedding(x)
            return torch.sigmoid(self.bias + torch.sum(self.linear(x), dim=1) + self.mlp(e.view(-1, FD)))
        e = self.feature_embedding(x)
        if self.kind == "FNN":                                             # :365-373
            return torch.sigmoid(self.mlp(e.view(-1, FD)))
        ip = torch.sum(torch.mul(e[:, self.row], e[:, self.col]), dim=2)   # :189
        return torch.sigmoid(self.mlp(torch.cat([e.view(-1, FD), ip], dim=1)))   # :194-198


TAIL_KINDS = ("WideAndDeep", "FNN", "InnerPNN", "OuterPNN", "DCN", "AFM")


def make_port(kind: str, feature_nums: int, field_nums: int = 15, latent_dims: int = 10) -> nn.Module:
    if kind in TAIL_KINDS:
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
