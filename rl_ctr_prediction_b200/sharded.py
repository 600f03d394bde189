"""Row-sharded embedding tables over the GPUs of one box (BASELINE.json configs[3]).

The reference is single-device; this is how its hot path scales.  One process per GPU
(``torch.distributed``, NCCL over NVLink/NVSwitch).  Samples are data-parallel; every table is
row-sharded: ``owner(id) = id mod G``, ``local_row(id) = id div G`` (modulo spreads the Zipf heads of a
dictionary-ranked vocabulary).  Per step and per model there are three all-to-alls:

    ids  -> owners        (8 B per gathered row; ``rlctr_bucket_by_owner`` lays out the send buffer)
    rows -> requesters    (row_stride * 4 B per gathered row; owners run ``rlctr_gather_rows``)
    row gradients -> owners, where ``rlctr_sort_ids`` + ``rlctr_rows_adam`` apply them to the shard

The interaction kernels run unchanged on the requester: the received rows ``[n, row_stride]`` are
addressed as a small table by the inverse permutation ``pos_of_slot``.  Optimizer state is sharded with
the rows; lazy-exact Adam works per shard (the owner catches its rows up before serving them).  Dense
parameters (bias, tower) are replicated and their gradients all-reduced.  Gradients are scaled by
1/G so that G ranks with local batch B/G take exactly the step one rank takes with batch B.

The bucketing is stable and the owner-side reduction walks a fixed order, so a sharded step is
bit-identical from run to run.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from . import p_model as Model
from .tables import Geometry, table_struct


# ---------------------------------------------------------------------------------------------------
# device operations used by the exchange (the product backend is the C ABI; tests inject a host one
# to exercise the routing logic over gloo without a GPU)
# ---------------------------------------------------------------------------------------------------
class CudaBackend:
    def bucket(self, ids_flat, world, n_rows):
        lib = _lib.load()
        n = ids_flat.numel()
        dev = ids_flat.device
        send_local = torch.empty(n, dtype=torch.int64, device=dev)
        pos_of_slot = torch.empty(n, dtype=torch.int64, device=dev)
        send_slots = torch.empty(n, dtype=torch.int32, device=dev)
        ends = torch.empty(world, dtype=torch.int64, device=dev)
        ws_bytes = lib.rlctr_bucket_ws_bytes(n, world)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.call("rlctr_bucket_by_owner", lib.rlctr_bucket_by_owner, _lib.ptr(ids_flat), n, world, n_rows,
                  _lib.ptr(send_local), _lib.ptr(pos_of_slot), _lib.ptr(send_slots), _lib.ptr(ends), _lib.ptr(ws), ws_bytes,
                  _lib.stream(), meta={"n": n})
        return send_local, pos_of_slot, send_slots, ends

    def gather(self, local_ids, table, geom):
        lib = _lib.load()
        n = local_ids.numel()
        out = torch.empty(n, geom.row_stride, dtype=torch.float32, device=table.device)
        t = table_struct(table, geom)
        _lib.call("rlctr_gather_rows", lib.rlctr_gather_rows, _lib.ptr(local_ids), n, C.byref(t), _lib.ptr(out), _lib.stream(),
                  meta={"n": n, "rs": geom.row_stride})
        return out


class ExchangePlan:
    __slots__ = ("n", "send_local", "pos_of_slot", "send_slots", "send_counts", "recv_counts", "recv_local", "n_recv",
                 "sorted_recv")


_PLAN_CACHE = {"key": None, "x": None, "plan": None}


def cached_exchange_plan(x, n_rows, group, backend) -> ExchangePlan:
    """Models that consume the same batch and shard the same vocabulary the same way (LR, FM, DeepFM...
    over one feature space) share the id exchange: the bucketing, the count exchange with its host
    synchronisation and the id all-to-all run once per batch, not once per model.  Keyed on the identity
    of the (unmodified) features tensor, which the cache keeps alive."""
    key = (id(x), x._version, x.data_ptr(), tuple(x.shape), int(n_rows), id(group))
    if _PLAN_CACHE["key"] == key and _PLAN_CACHE["x"] is x:
        return _PLAN_CACHE["plan"]
    plan = exchange_plan(x, n_rows, group, backend)
    _PLAN_CACHE.update(key=key, x=x, plan=plan)
    return plan


def _counts_from_ends(ends_host, n):
    """bucket_ends (-1 for empty buckets) -> per-owner counts."""
    counts, prev = [], 0
    for e in ends_host:
        e = prev if e < 0 else int(e)
        counts.append(e - prev)
        prev = e
    assert prev == n or n == 0, (prev, n)
    return counts


def exchange_plan(ids, n_rows, group, backend) -> ExchangePlan:
    """Route the ids of the local batch to their owners (all-to-all #1)."""
    world = dist.get_world_size(group)
    flat = ids.reshape(-1).contiguous()
    p = ExchangePlan()
    p.n = flat.numel()
    p.send_local, p.pos_of_slot, p.send_slots, ends = backend.bucket(flat, world, n_rows)
    p.send_counts = _counts_from_ends(ends.tolist(), p.n)            # one D2H sync: variable-size all-to-all
    sc = torch.tensor(p.send_counts, dtype=torch.int64, device=flat.device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc, group=group)
    p.recv_counts = rc.tolist()
    p.n_recv = int(sum(p.recv_counts))
    p.recv_local = torch.empty(p.n_recv, dtype=torch.int64, device=flat.device)
    dist.all_to_all_single(p.recv_local, p.send_local, p.recv_counts, p.send_counts, group=group)
    p.sorted_recv = None
    return p


def fetch_rows(plan, table, geom, group, backend):
    """Owners gather the requested rows and send them back (all-to-all #2): [n, row_stride] in send order."""
    served = backend.gather(plan.recv_local, table, geom)
    rows = torch.empty(plan.n, geom.row_stride, dtype=torch.float32, device=table.device)
    dist.all_to_all_single(rows, served, plan.send_counts, plan.recv_counts, group=group)
    return rows


def push_grads(plan, grad_rows, group):
    """Row gradients (send order) travel to the owners (all-to-all #3): [n_recv, row_stride]."""
    out = torch.empty(plan.n_recv, grad_rows.shape[1], dtype=torch.float32, device=grad_rows.device)
    dist.all_to_all_single(out, grad_rows.contiguous(), plan.recv_counts, plan.send_counts, group=group)
    return out


def shard_rows(n_rows, world, rank):
    return (n_rows - rank + world - 1) // world if n_rows > rank else 0


# ---------------------------------------------------------------------------------------------------
class ShardedCTR(nn.Module):
    """LR / FM / FFM / DeepFM with the table row-sharded over the process group.

    ``train_step(features, labels, optimizer)`` is the reference loop body (src/main/pretrain_main.py:96-102) for
    the local batch; ``forward(features)`` is the frozen scoring pass.  ``from_model`` shards an existing
    single-GPU model (for the 1-vs-G equivalence check); ``gather_state_dict`` rebuilds the reference-keyed
    state_dict on every rank."""

    def __init__(self, kind, feature_nums, field_nums, latent_dims, group=None, device=None):
        super().__init__()
        assert kind in ("LR", "FM", "FFM", "DeepFM")
        self.kind, self.feature_nums, self.field_nums, self.latent_dims = kind, int(feature_nums), int(field_nums), int(latent_dims)
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.backend = CudaBackend()
        n_local = max(shard_rows(self.feature_nums, self.world, self.rank), 1)
        if kind == "LR":
            g = Geometry.lr(n_local)
        elif kind == "FFM":
            g = Geometry.ffm(n_local, self.field_nums, self.latent_dims)
        else:
            g = Geometry.fm(n_local, self.latent_dims)
        g = g.with_state()
        self._geom = g
        self._kind = {"LR": "lr", "FM": "fm", "DeepFM": "fm", "FFM": "ffm"}[kind]
        dev = torch.device(device) if device is not None else None
        data = torch.zeros(g.n_rows, g.row_pitch, dtype=torch.float32, device=dev)
        used = [g.lin_col] if g.lin_col >= 0 else []
        used += list(range(g.emb_col, g.emb_col + g.dim))
        if used:
            data[:, used] = torch.randn(g.n_rows, len(used), device=dev)
        self.table = nn.Parameter(data)
        self.table._rlctr_owner = self
        self.bias = nn.Parameter(torch.zeros(1, device=dev))
        self.mlp = Model._tower(self.field_nums * self.latent_dims, device) if kind == "DeepFM" else None
        self._opt = None
        self._stash = None
        self._ws = {}

    # ---- protocol shared with optim.Adam --------------------------------------------------------
    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self.table._rlctr_owner = self
        self._ws = {}
        return out

    def _meta(self, B, F):
        g = self._geom
        return {"model": "Sharded" + self.kind, "B": B, "F": max(F, 1), "rs": g.row_stride, "dim": g.dim, "n_rows": g.n_rows,
                "lin": g.lin_col >= 0}

    def flush(self):
        if self._opt is not None:
            self._opt.flush(self.table.data)

    def _reduce_ws(self, dev):
        ws = self._ws.get("reduce")
        if ws is None:
            ws = torch.zeros(_lib.RLCTR_REDUCE_WS_BYTES, dtype=torch.uint8, device=dev)
            self._ws["reduce"] = ws
        return ws

    def zero_grad(self, set_to_none=True):
        self._stash = None
        return super().zero_grad(set_to_none)

    @classmethod
    def from_model(cls, model, group=None):
        """Shard a single-GPU p_model instance (same parameters on every rank) over the group."""
        kind = type(model).__name__
        self = cls(kind, model.feature_nums, getattr(model, "field_nums", 15), getattr(model, "latent_dims", 1),
                   group=group, device=model.table.device)
        with torch.no_grad():
            shard = model.table.data[self.rank::self.world]
            self.table.data[:shard.shape[0]].copy_(shard)
            self.bias.data.copy_(model.bias.data)
            if self.mlp is not None:
                self.mlp.load_state_dict(model.mlp.state_dict())
        return self

    def gather_table(self):
        """The full fused table [N, row_stride], rebuilt on every rank (tests / checkpointing)."""
        self.flush()
        n_max = shard_rows(self.feature_nums, self.world, 0)
        mine = torch.zeros(n_max, self._geom.row_pitch, dtype=torch.float32, device=self.table.device)
        mine[:self._geom.n_rows] = self.table.data
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        full = torch.empty(self.feature_nums, self._geom.row_pitch, dtype=torch.float32, device=mine.device)
        for r in range(self.world):
            cnt = shard_rows(self.feature_nums, self.world, r)
            full[r::self.world] = parts[r][:cnt]
        return full

    # ---- the step ---------------------------------------------------------------------------------
    def _lookup(self, x, train):
        lib = _lib.load()
        plan = cached_exchange_plan(x, self.feature_nums, self.group, self.backend)
        sorted_pair = None
        if train:
            if plan.n_recv and (plan.sorted_recv is None or plan.sorted_recv[0] != self._geom.n_rows):
                plan.sorted_recv = (self._geom.n_rows, Model._sort_ids(plan.recv_local, self._geom.n_rows))
            sorted_pair = plan.sorted_recv[1] if plan.n_recv else None
            opt = self._opt
            if sorted_pair is not None and opt is not None and opt.lazy and opt.dirty:
                t, a = table_struct(self.table.data, self._geom), opt.struct()
                _lib.call("rlctr_rows_catchup", lib.rlctr_rows_catchup, _lib.ptr(sorted_pair[0]), plan.n_recv, C.byref(t),
                          C.byref(a), _lib.stream(), key=f"rlctr_rows_catchup[Sharded{self.kind}]",
                          meta=self._meta(plan.n_recv, 1))
        else:
            self.flush()
        rows = fetch_rows(plan, self.table.data, self._geom, self.group, self.backend)
        return plan, sorted_pair, rows

    def _interact(self, x, plan, rows, train):
        """Run the single-GPU interaction kernels over the received rows (a table of n rows addressed by
        pos_of_slot).  Returns (logit[B], sums, partners, tower_rows)."""
        lib = _lib.load()
        B, F = x.shape
        dev = x.device
        g = self._geom
        gl = g.rows_only(max(plan.n, 1))
        t = table_struct(rows, gl)
        pos = plan.pos_of_slot.view(B, F)
        logit = torch.empty(B, dtype=torch.float32, device=dev)
        sums = partners = trows = None
        if self._kind == "ffm":
            if train:
                partners = torch.empty(B * F, g.row_stride, dtype=torch.float32, device=dev)
            _lib.call("rlctr_ffm_fwd", lib.rlctr_ffm_fwd, _lib.ptr(pos), C.byref(t), _lib.ptr(self.bias.data), _lib.ptr(logit),
                      None, 1, _lib.ptr(partners), B, F, self.latent_dims, _lib.stream(), key="rlctr_ffm_fwd[sharded]",
                      meta=self._meta(B, F))
        else:
            if train and self._kind == "fm":
                sums = torch.empty(B, g.row_stride, dtype=torch.float32, device=dev)
            if self.kind == "DeepFM":
                trows = torch.empty(B, F * g.dim, dtype=torch.float32, device=dev)
            flags = _lib.RLCTR_FM_TERM if self.kind in ("FM", "DeepFM") else 0
            _lib.call("rlctr_embed_fwd", lib.rlctr_embed_fwd, _lib.ptr(pos), C.byref(t), _lib.ptr(self.bias.data), _lib.ptr(logit),
                      None, 1, _lib.ptr(sums), _lib.ptr(trows), 0, B, F, flags, _lib.stream(),
                      key=f"rlctr_embed_fwd[Sharded{self.kind}]", meta=dict(self._meta(B, F), sums=sums is not None,
                                                                            rows=trows is not None))
        return logit, sums, partners, trows, gl

    @torch.no_grad()
    def forward(self, x):
        x = Model._check_ids(x)
        plan, _, rows = self._lookup(x, train=False)
        logit, _, _, trows, _ = self._interact(x, plan, rows, train=False)
        if self.mlp is not None:
            logit = logit + self.mlp(trows).reshape(-1)
        return torch.sigmoid(logit).reshape(-1, 1)

    def train_step(self, features, labels, optimizer):
        """One step on the local batch; the update equals the single-GPU step on the concatenated global
        batch.  Returns the local mean BCE loss (device scalar)."""
        lib = _lib.load()
        x = Model._check_ids(features)
        B, F = x.shape
        dev = x.device
        st = _lib.stream()
        g = self._geom
        y = labels.reshape(-1).contiguous()
        plan, sorted_pair, rows = self._lookup(x, train=True)
        logit, sums, partners, trows, gl = self._interact(x, plan, rows, train=True)
        tower_out = None
        if self.mlp is not None:
            trows.requires_grad_(True)
            with torch.enable_grad():
                tower_out = self.mlp(trows).reshape(-1)
            logit = logit + tower_out.detach()
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dlogit = torch.empty(B, dtype=torch.float32, device=dev)
        dbias = torch.empty(1, dtype=torch.float32, device=dev)
        yi = y if y.dtype == torch.int64 else None
        yf = None if yi is not None else y.float()
        _lib.check(lib.rlctr_bce_fwd_bwd(_lib.ptr(logit), _lib.ptr(yi), _lib.ptr(yf), None, _lib.ptr(loss), _lib.ptr(dlogit),
                                         _lib.ptr(dbias), _lib.ptr(self._reduce_ws(dev)), B, st), "rlctr_bce_fwd_bwd")
        if self.world > 1:                                    # gradient of the GLOBAL mean loss
            dlogit.mul_(1.0 / self.world)
            dbias.mul_(1.0 / self.world)
        extra = None
        for p in self.parameters():
            p.grad = None
        if tower_out is not None:
            tower_out.backward(dlogit)
            extra = trows.grad.contiguous()
        # row gradients in send-buffer order: position k of the send buffer <- slot send_slots[k]
        n = plan.n
        gbuf = torch.empty(n, g.row_stride, dtype=torch.float32, device=dev)
        ident = self._ws.get("ident")
        if ident is None or ident.numel() < n:
            ident = torch.arange(n, dtype=torch.int32, device=dev)
            self._ws["ident"] = ident
        grad = _lib.RowGrad(_lib.ptr(partners), _lib.ptr(dlogit), _lib.ptr(sums), _lib.ptr(extra), F,
                            _lib.RLCTR_STAGED_PARTNER if partners is not None else 0)
        t = table_struct(rows, gl)
        ws_bytes = lib.rlctr_rows_ws_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.call("rlctr_rows_grad_dense", lib.rlctr_rows_grad_dense, _lib.ptr(ident), _lib.ptr(plan.send_slots), n,
                  C.byref(grad), C.byref(t), _lib.ptr(gbuf), _lib.ptr(ws), ws_bytes, st,
                  key=f"rlctr_rows_grad_dense[Sharded{self.kind}]", meta=self._meta(B, F))
        recv_grads = push_grads(plan, gbuf, self.group)
        if sorted_pair is not None:
            self._stash = Model.RowsStash(sorted_ids=sorted_pair[0], sorted_slots=sorted_pair[1], n=plan.n_recv, dlogit=None,
                                          sums=None, extra=None, staged=recv_grads, fields=1, flags=0)
        self.bias.grad = dbias
        if self.world > 1:
            dense = [p for p in self.parameters() if p is not self.table and p.grad is not None]
            flat = torch.cat([p.grad.reshape(-1) for p in dense])
            dist.all_reduce(flat, group=self.group)
            o = 0
            for p in dense:
                p.grad = flat[o:o + p.numel()].view_as(p).clone()
                o += p.numel()
        optimizer.step()
        return loss.reshape(())
