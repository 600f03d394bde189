"""Row-sharded embedding tables over the GPUs of one box (BASELINE.json configs[3]).

The reference is single-device; this is how its hot path scales.  One process per GPU (``torch.distributed``, NCCL
over NVLink/NVSwitch).  Samples are data-parallel; every table is row-sharded: ``owner(id) = id mod G``,
``local_row(id) = id div G`` (modulo spreads the Zipf heads of a dictionary-ranked vocabulary); optimizer state
is sharded with the rows.  Dense parameters (bias, tower) are replicated and their gradients all-reduced (NCCL).

The lookup and the gradient exchange are NOT collectives: every shard (and every rank's small gradient-side
buffers) lives in symmetric memory -- peer-mapped over NVLink -- and the kernels address it directly:

    forward   rlctr_embed_fwd / rlctr_ffm_fwd read row ``id`` at ``peers[id % G] + (id / G) * pitch``: the gather
              IS the exchange (64 B NVLink reads; measured 529 GB/s for random rows, profiles/r1_symm_probe.md);
    backward  every rank leaves dlogit / column sums / tower-input gradients in its symmetric buffers; the OWNER of a
              row pulls them (``rlctr_rowgrad.peer_*``) inside the sort / segment-reduce / fused-Adam kernel and
              recomputes the row gradient exactly as the single-GPU kernel does.

What does travel as a collective is the id list: one fixed-size ``all_gather`` of the batch's ids per step (4 B per
gathered row, shared by every model fed with that batch).  Each rank keeps the ids it owns (sentinel keys for the
rest), sorts them once (``rlctr_sort_ids_sharded``) and uses that view for the lazy-Adam catch-up and for the
update.  Two device-side barriers per model step order the phases (catch-up | forward reads ... backward | owner
update); nothing synchronises with the host and no count ever leaves the device, so the sharded step is as
graph-capturable as the single-GPU one.  The owner walks occurrences in (source rank, slot) order: a sharded step
is bit-identical from run to run and equal to the single-GPU step on the concatenated batch.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from . import p_model as Model
from .tables import Geometry, table_struct


def shard_rows(n_rows, world, rank):
    return (n_rows - rank + world - 1) // world if n_rows > rank else 0


def owned_sorted_view_host(ids_all: torch.Tensor, world: int, rank: int, n_rows: int):
    """Host restatement of rlctr_sort_ids_sharded (tests; CPU tensors): (sorted local rows, sorted global slots) of the
    ids this rank owns, in stable (row, position) order -- the owned prefix only, without the sentinel tail."""
    flat = ids_all.reshape(-1).to(torch.int64)
    pos = torch.arange(flat.numel(), dtype=torch.int64)
    mine = (flat >= 0) & (flat < n_rows) & (flat % world == rank)
    rows, pos = flat[mine] // world, pos[mine]
    order = torch.sort(rows, stable=True).indices
    return rows[order], pos[order]


class PeerMemory:
    """One symmetric allocation: the same number of elements on every rank, peer-mapped into every process.
    ``local`` is this rank's tensor, ``ptrs[r]`` the address of rank r's copy in THIS process.  world == 1: a plain
    tensor."""

    def __init__(self, numel, dtype, device, group):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            self.local = symm.empty(int(numel), dtype=dtype, device=device)
            self.handle = symm.rendezvous(self.local, group if group is not None else dist.group.WORLD)
            self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        else:
            self.local = torch.empty(int(numel), dtype=dtype, device=device)
            self.handle = None
            self.ptrs = [self.local.data_ptr()]

    def barrier(self):
        """All ranks' work enqueued before this point is complete before any rank's work enqueued after it starts
        (device side, stream ordered: symmetric-memory signal pads, no host involvement)."""
        if self.handle is not None:
            self.handle.barrier(channel=0)


_VIEW_CACHE = {"key": None, "x": None, "view": None}
_ROUTE = {}          # (n, world, device) -> receive buffers of the owner routing


def route_capacity(n, world):
    """Slots each source rank gets in every owner's receive buffer: RLCTR_ROUTE_CAP x the balanced share n / G (default 1.5;
    ids spread by ``id mod G``, so a bucket is Binomial(n, 1/G): 1.5x is > 100 sigma for the uniform benchmark ids, and a
    heavy-hitter id that alone exceeds it trips the overflow flag instead of corrupting the step silently)."""
    f = float(os.environ.get("RLCTR_ROUTE_CAP", "1.5"))
    return ((int(f * n / world) + 1024) + 255) // 256 * 256


def _route_buffers(n, world, dev, group):
    key = (n, world, str(dev))
    rb = _ROUTE.get(key)
    if rb is None:
        cap = route_capacity(n, world)
        keys = PeerMemory(world * cap, torch.int32, dev, group)
        vals = PeerMemory(world * cap, torch.int32, dev, group)
        keys.local.fill_(-1)
        vals.local.zero_()
        rb = {"cap": cap, "keys": keys, "vals": vals, "overflow": torch.zeros(1, dtype=torch.int32, device=dev),
              "kp": (C.c_void_p * 8)(*[int(p) for p in keys.ptrs]), "vp": (C.c_void_p * 8)(*[int(p) for p in vals.ptrs])}
        keys.barrier()
        _ROUTE[key] = rb
    return rb


def check_route_overflow():
    """Host-side check of the routing overflow flags (one device -> host read each; called from ShardedCTR.flush)."""
    for (n, world, _), rb in _ROUTE.items():
        if int(rb["overflow"].item()) != 0:
            raise _lib.RlctrError(f"owner routing overflow: some rank sent more than {rb['cap']} of its {n} ids to one owner "
                                  f"(skewed ids); the steps since the last check are incomplete.  Rerun with a larger "
                                  f"RLCTR_ROUTE_CAP (now {os.environ.get('RLCTR_ROUTE_CAP', '1.5')}) or RLCTR_SHARD_ROUTE=0")


def shared_sorted_view(x, n_rows, group):
    """(sorted local rows u32[n_all], sorted global slots u32[n_all], n_all) of the GLOBAL batch as seen by this owner.
    Models that consume the same batch and shard the same vocabulary share it (one exchange + one sort per batch).

    Default (RLCTR_SHARD_ROUTE=1): owner routing -- every rank writes the (local row, global slot) pairs of the ids a peer
    owns straight into that peer's receive buffer (rlctr_route_ids: posted NVLink writes, fixed capacity per source), one
    device barrier, and the owner sorts its G * cap ~ 1.5 n received pairs (rlctr_sort_routed): per-rank work independent
    of G.  RLCTR_SHARD_ROUTE=0: all_gather of the ids + rlctr_sort_ids_sharded over all G * n of them (the first design)."""
    # (the group is not part of the key: models with process groups of their own -- same ranks -- share the view)
    key = (id(x), x._version, x.data_ptr(), tuple(x.shape), int(n_rows))
    if _VIEW_CACHE["key"] == key and _VIEW_CACHE["x"] is x:
        return _VIEW_CACHE["view"]
    lib = _lib.load()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = x.device
    n = x.numel()
    st = _lib.stream()
    if world > 1 and os.environ.get("RLCTR_SHARD_ROUTE", "1") == "1":
        rb = _route_buffers(n, world, dev, group)
        cap = rb["cap"]
        ws_bytes = lib.rlctr_route_ws_bytes(n, world)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.call("rlctr_route_ids", lib.rlctr_route_ids, _lib.ptr(x), n, world, rank, n_rows, cap, rb["kp"], rb["vp"],
                  _lib.ptr(rb["overflow"]), _lib.ptr(ws), ws_bytes, st, meta={"n": n, "world": world, "cap": cap})
        rb["keys"].barrier()                                     # every source's segment of my receive buffer is complete
        n_all = world * cap
        srows = torch.empty(n_all, dtype=torch.int32, device=dev)
        sslots = torch.empty(n_all, dtype=torch.int32, device=dev)
        n_local = shard_rows(n_rows, world, rank)
        ws2_bytes = lib.rlctr_sort_ws_bytes(n_all, max(n_local, 1))
        ws2 = torch.empty(ws2_bytes, dtype=torch.uint8, device=dev)
        _lib.call("rlctr_sort_routed", lib.rlctr_sort_routed, _lib.ptr(rb["keys"].local), _lib.ptr(rb["vals"].local), n_all,
                  max(n_local, 1), _lib.ptr(srows), _lib.ptr(sslots), _lib.ptr(ws2), ws2_bytes, st, key="rlctr_sort_ids",
                  meta={"n": n_all})
        view = (srows, sslots, n_all)
        _VIEW_CACHE.update(key=key, x=x, view=view)
        return view
    ids32 = x.reshape(-1).clamp(-1, n_rows).to(torch.int32)          # out-of-range ids stay out of range in 32 bits
    if world > 1:
        all32 = torch.empty(world * n, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all32, ids32, group=group)
    else:
        all32 = ids32
    n_all = world * n
    srows = torch.empty(n_all, dtype=torch.int32, device=dev)
    sslots = torch.empty(n_all, dtype=torch.int32, device=dev)
    ws_bytes = lib.rlctr_sort_ws_bytes(n_all, n_rows)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("rlctr_sort_ids_sharded", lib.rlctr_sort_ids_sharded, _lib.ptr(all32), n_all, world, rank, n_rows, _lib.ptr(srows),
              _lib.ptr(sslots), _lib.ptr(ws), ws_bytes, st, key="rlctr_sort_ids", meta={"n": n_all})
    view = (srows, sslots, n_all)
    _VIEW_CACHE.update(key=key, x=x, view=view)
    return view


# ---------------------------------------------------------------------------------------------------
class ShardedCTR(nn.Module):
    """LR / FM / FFM / DeepFM with the table row-sharded over the process group.

    ``train_step(features, labels, optimizer)`` is the reference loop body (src/main/pretrain_main.py:96-102) for
    the local batch; ``forward(features)`` is the frozen scoring pass.  ``from_model`` shards an existing
    single-GPU model (for the 1-vs-G equivalence check); ``gather_table`` rebuilds the full fused table on every
    rank.  Every rank must call the same methods in the same order (they contain device barriers)."""

    def __init__(self, kind, feature_nums, field_nums, latent_dims, group=None, device=None):
        super().__init__()
        assert kind in ("LR", "FM", "FFM", "DeepFM")
        self.kind, self.feature_nums, self.field_nums, self.latent_dims = kind, int(feature_nums), int(field_nums), int(latent_dims)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world not in (1, 2, 4, 8):
            raise _lib.RlctrError("row sharding supports 1, 2, 4 or 8 GPUs (owner = id & (G-1))")
        n_local = max(shard_rows(self.feature_nums, self.world, self.rank), 1)
        n_alloc = max(shard_rows(self.feature_nums, self.world, 0), 1)           # same allocation size on every rank
        if kind == "LR":
            g = Geometry.lr(n_local)
        elif kind == "FFM":
            g = Geometry.ffm(n_local, self.field_nums, self.latent_dims)
        else:
            g = Geometry.fm(n_local, self.latent_dims)
        g = g.with_state()
        self._geom = g
        self._kind = {"LR": "lr", "FM": "fm", "DeepFM": "fm", "FFM": "ffm"}[kind]
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._pm = PeerMemory(n_alloc * g.row_pitch, torch.float32, dev, group)
        data = self._pm.local.view(n_alloc, g.row_pitch)
        data.zero_()
        used = [g.lin_col] if g.lin_col >= 0 else []
        used += list(range(g.emb_col, g.emb_col + g.dim))
        if used:
            data[:g.n_rows, used] = torch.randn(g.n_rows, len(used), device=dev)
        self.table = nn.Parameter(data)
        self.table._rlctr_owner = self
        self.bias = nn.Parameter(torch.zeros(1, device=dev))
        self.mlp = Model._tower(self.field_nums * self.latent_dims, dev) if kind == "DeepFM" else None
        self._opt = None
        self._stash = None
        self._ws = {}
        self._gbuf = {}
        self._xbuf = {}
        # how the owners get the gradient side of remote samples: "push" = replicate the per-sample rows (dlogit / column sums)
        # with one all_gather and let the source ranks WRITE the per-occurrence rows (DeepFM's tower-input gradients) into the
        # owners' memory, so that the update kernel reads local memory only; "pull" = the owners read everything through the peer
        # mapping inside the update kernel (2-3 us NVLink round trips on its dependent path; kept for comparison and for FFM's
        # 600-byte partner rows)
        self.exchange = os.environ.get("RLCTR_SHARD_EXCHANGE", "push")
        # True when `group` is used by this model alone: its step may then run as a parallel branch of a captured graph
        # (graphs.GraphedTrainStep), overlapping its NVLink-bound gathers with another model's GEMMs
        self.fork_ok = False
        self._after_gather = None

    # ---- protocol shared with optim.Adam --------------------------------------------------------
    def _apply(self, fn, recurse=True):
        # the shard lives in symmetric memory: .to()/.cuda() must not move it
        saved = self._parameters.pop("table")
        out = super()._apply(fn, recurse)
        self._parameters["table"] = saved
        self.table._rlctr_owner = self
        return out

    def _meta(self, B, F):
        g = self._geom
        return {"model": "Sharded" + self.kind, "B": B, "F": max(F, 1), "rs": g.row_stride, "dim": g.dim, "n_rows": g.n_rows,
                "lin": g.lin_col >= 0}

    def flush(self):
        if self._opt is not None:
            self._opt.flush(self.table.data)
        check_route_overflow()

    def _reduce_ws(self, dev):
        ws = self._ws.get("reduce")
        if ws is None:
            ws = torch.zeros(_lib.RLCTR_REDUCE_WS_BYTES, dtype=torch.uint8, device=dev)
            self._ws["reduce"] = ws
        return ws

    def zero_grad(self, set_to_none=True):
        self._stash = None
        return super().zero_grad(set_to_none)

    _rows_ws = Model._TableModel._rows_ws

    def barrier(self):
        self._pm.barrier()

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
        self.barrier()
        return self

    def gather_table(self):
        """The full fused table [N, row_pitch], rebuilt on every rank (tests / checkpointing)."""
        self.flush()
        mine = self.table.data.contiguous()
        if self.world == 1:
            return mine[:self.feature_nums].clone()
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        full = torch.empty(self.feature_nums, self._geom.row_pitch, dtype=torch.float32, device=mine.device)
        for r in range(self.world):
            cnt = shard_rows(self.feature_nums, self.world, r)
            full[r::self.world] = parts[r][:cnt]
        return full

    # ---- table structs ----------------------------------------------------------------------------
    def _local_struct(self):
        """The shard as the optimizer kernels see it: n_local rows, local row ids."""
        return table_struct(self.table.data, self._geom)

    def _global_struct(self):
        """The whole table as the forward gathers see it: global ids, peers[] = every rank's shard."""
        g = self._geom
        t = _lib.Table(self.table.data.data_ptr(), self.feature_nums, g.row_stride, g.lin_col, g.emb_col, g.dim, g.row_pitch)
        if self.world > 1:
            t.world = self.world
            for r, p in enumerate(self._pm.ptrs):
                t.peers[r] = p
        return t

    def _grad_buffers(self, B, F):
        """dlogit [B] | sums [B, rs] | extra [B, F*D] | staged [B*F, rs] in ONE symmetric allocation per batch size:
        what the owners of this rank's rows pull after the backward."""
        buf = self._gbuf.get(B)
        if buf is not None:
            return buf
        g = self._geom
        sizes = {"dlogit": (B + 3) // 4 * 4,
                 "sums": B * g.row_stride if self._kind == "fm" else 0,
                 "extra": B * F * g.dim if self.kind == "DeepFM" else 0,
                 "staged": B * F * g.row_stride if self._kind == "ffm" else 0}
        total = sum(sizes.values())
        pm = PeerMemory(total, torch.float32, self.table.device, self.group)
        buf, off = {"pm": pm}, 0
        for k, sz in sizes.items():
            if sz:
                buf[k] = pm.local[off:off + sz]
                buf[k + "_ptrs"] = [p + 4 * off for p in pm.ptrs]
            else:
                buf[k], buf[k + "_ptrs"] = None, None
            off += sz
        self._gbuf[B] = buf
        return buf

    def _exchange_buffers(self, B, F):
        """Receive side of the "push" exchange for batch size B: the all-gathered per-sample rows of every rank (plain local
        tensors) and the symmetric buffer the sources write DeepFM's per-occurrence rows into."""
        xb = self._xbuf.get(B)
        if xb is not None:
            return xb
        g, dev, G = self._geom, self.table.device, self.world
        bp = (B + 3) // 4 * 4
        xb = {"dlogit_all": torch.empty(G * bp, dtype=torch.float32, device=dev), "bp": bp, "sums_all": None, "recv": None}
        if self._kind == "fm":
            xb["sums_all"] = torch.empty(G * B * g.row_stride, dtype=torch.float32, device=dev)
        if self.kind == "DeepFM":
            pm = PeerMemory(G * B * F * g.dim, torch.float32, dev, self.group)
            pm.local.zero_()
            xb["recv"] = pm
            xb["recv_ptrs"] = (C.c_void_p * 8)(*[int(p) for p in pm.ptrs])
        self._xbuf[B] = xb
        return xb

    # ---- the step ---------------------------------------------------------------------------------
    def _interact(self, x, buf, train):
        """The single-GPU interaction kernels reading rows through the peer table.  Returns (logit[B], tower rows)."""
        lib = _lib.load()
        B, F = x.shape
        dev = x.device
        g = self._geom
        t = self._global_struct()
        logit = torch.empty(B, dtype=torch.float32, device=dev)
        trows = None
        if self._kind == "ffm":
            partners = buf["staged"] if train else None
            _lib.call("rlctr_ffm_fwd", lib.rlctr_ffm_fwd, _lib.ptr(x), C.byref(t), _lib.ptr(self.bias.data), _lib.ptr(logit),
                      None, 1, _lib.ptr(partners), B, F, self.latent_dims, _lib.stream(), key="rlctr_ffm_fwd[sharded]",
                      meta=self._meta(B, F))
        else:
            sums = buf["sums"] if (train and self._kind == "fm") else None
            pitch = 0
            if self.kind == "DeepFM":
                pitch = (F * g.dim + 3) // 4 * 4
                trows = torch.empty(B, pitch, dtype=torch.float32, device=dev)
            flags = _lib.RLCTR_FM_TERM if self.kind in ("FM", "DeepFM") else 0
            _lib.call("rlctr_embed_fwd", lib.rlctr_embed_fwd, _lib.ptr(x), C.byref(t), _lib.ptr(self.bias.data), _lib.ptr(logit),
                      None, 1, _lib.ptr(sums), _lib.ptr(trows), pitch, B, F, flags, _lib.stream(),
                      key=f"rlctr_embed_fwd[Sharded{self.kind}]",
                      meta=dict(self._meta(B, F), sums=sums is not None, rows=trows is not None, n_rows=self.feature_nums))
            if trows is not None and pitch != F * g.dim:
                trows = trows[:, :F * g.dim]
        return logit, trows

    @torch.no_grad()
    def forward(self, x):
        x = Model._check_ids(x)
        self.flush()
        self.barrier()                                        # every shard is settled before anyone reads it
        logit, trows = self._interact(x, None, train=False)
        if self.mlp is not None:
            logit = logit + self.mlp(trows).reshape(-1)
        self.barrier()                                        # nobody starts writing rows while a peer still reads
        return torch.sigmoid(logit).reshape(-1, 1)

    def train_step(self, features, labels, optimizer):
        """One step on the local batch; the update equals the single-GPU step on the concatenated global
        batch.  Returns the local mean BCE loss (device scalar)."""
        lib = _lib.load()
        x = Model._check_ids(features)
        B, F = x.shape
        dev = x.device
        st = _lib.stream()
        y = labels.reshape(-1).contiguous()
        buf = self._grad_buffers(B, F)
        srows, sslots, n_all = shared_sorted_view(x, self.feature_nums, self.group)
        opt = self._opt
        if opt is not None and opt.lazy and opt.dirty:
            t, a = self._local_struct(), opt.struct()
            _lib.call("rlctr_rows_catchup", lib.rlctr_rows_catchup, _lib.ptr(srows), n_all, C.byref(t), C.byref(a), st,
                      key=f"rlctr_rows_catchup[Sharded{self.kind}]", meta=self._meta(n_all // F, F))
        self.barrier()                                        # B1: all owners caught their rows up | forward reads
        logit, trows = self._interact(x, buf, train=True)
        hook, self._after_gather = self._after_gather, None     # graphs.GraphedTrainStep: fork point of the captured step
        if hook is not None:
            hook()
        tower_out = None
        if self.mlp is not None:
            trows.requires_grad_(True)
            with torch.enable_grad():
                tower_out = self.mlp(trows).reshape(-1)
            logit = logit + tower_out.detach()
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dlogit = buf["dlogit"][:B]
        dbias = torch.empty(1, dtype=torch.float32, device=dev)
        yi = y if y.dtype == torch.int64 else None
        yf = None if yi is not None else y.float()
        _lib.check(lib.rlctr_bce_fwd_bwd(_lib.ptr(logit), _lib.ptr(yi), _lib.ptr(yf), None, _lib.ptr(loss), _lib.ptr(dlogit),
                                         _lib.ptr(dbias), _lib.ptr(self._reduce_ws(dev)), B, st), "rlctr_bce_fwd_bwd")
        if self.world > 1:                                    # gradient of the GLOBAL mean loss
            dlogit.mul_(1.0 / self.world)
            dbias.mul_(1.0 / self.world)
        dz_in_sums = self.world > 1 and self._kind == "fm"
        if dz_in_sums:                                        # one aligned 64 B line per occurrence for the owners to pull
            buf["sums"].view(B, -1)[:, -1] = dlogit
        for p in self.parameters():
            p.grad = None
        if tower_out is not None:
            tower_out.backward(dlogit)
            buf["extra"].view(B, -1).copy_(trows.grad)
        self.bias.grad = dbias
        if self.world > 1:
            dense = [p for p in self.parameters() if p is not self.table and p.grad is not None]
            flat = torch.cat([p.grad.reshape(-1) for p in dense])
            dist.all_reduce(flat, group=self.group)
            o = 0
            for p in dense:
                p.grad = flat[o:o + p.numel()].view_as(p).clone()
                o += p.numel()
        peer = None
        if self.world > 1:
            peer = {"world": self.world, "n_per_rank": B * F, "staged": buf["staged_ptrs"], "dlogit": buf["dlogit_ptrs"],
                    "sums": buf["sums_ptrs"], "extra": buf["extra_ptrs"]}
            if self.exchange == "push":
                xb = self._exchange_buffers(B, F)
                g = self._geom
                if dz_in_sums:                                # dlogit rides in the sums rows: one all_gather serves both
                    dist.all_gather_into_tensor(xb["sums_all"], buf["sums"], group=self.group)
                    per = 4 * B * g.row_stride
                    peer["sums"] = [xb["sums_all"].data_ptr() + r * per for r in range(self.world)]
                    peer["dlogit"] = [xb["dlogit_all"].data_ptr() + r * 4 * xb["bp"] for r in range(self.world)]   # not read
                else:
                    dist.all_gather_into_tensor(xb["dlogit_all"], buf["dlogit"], group=self.group)
                    peer["dlogit"] = [xb["dlogit_all"].data_ptr() + r * 4 * xb["bp"] for r in range(self.world)]
                if xb["recv"] is not None:                    # per-occurrence rows: posted writes into the owners' memory
                    n = B * F
                    _lib.call("rlctr_push_rows", lib.rlctr_push_rows, _lib.ptr(x), n, self.world, self.rank, self.feature_nums,
                              _lib.ptr(buf["extra"]), g.dim, xb["recv_ptrs"], st, meta={"n": n, "width": g.dim})
                    mine = xb["recv"].local.data_ptr()
                    peer["extra"] = [mine + r * 4 * n * g.dim for r in range(self.world)]
        self.barrier()                                        # B2: every rank's gradient-side buffers are complete
        self._stash = Model.RowsStash(sorted_ids=srows, sorted_slots=sslots, n=n_all, dlogit=dlogit, sums=buf["sums"],
                                      extra=buf["extra"], staged=buf["staged"], fields=F,
                                      flags=(_lib.RLCTR_STAGED_PARTNER if buf["staged"] is not None else 0) |
                                            (_lib.RLCTR_DZ_IN_SUMS if dz_in_sums else 0), peer=peer)
        optimizer.step()
        return loss.reshape(())


def routed_view_pos(x, n_rows, group):
    """The routed sorted view with RECEIVE POSITIONS as values (rlctr_sort_routed_pos), plus what the routed gradient push needs:
    {"rows", "pos", "n_all", "cap", "route_ws" (bucket order of this rank's ids), "slot_of" (global slot at every receive position)}."""
    lib = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev, n, st = x.device, x.numel(), _lib.stream()
    rb = _route_buffers(n, world, dev, group)
    cap = rb["cap"]
    ws_bytes = lib.rlctr_route_ws_bytes(n, world)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.call("rlctr_route_ids", lib.rlctr_route_ids, _lib.ptr(x), n, world, rank, n_rows, cap, rb["kp"], rb["vp"],
              _lib.ptr(rb["overflow"]), _lib.ptr(ws), ws_bytes, st, meta={"n": n, "world": world, "cap": cap})
    rb["keys"].barrier()                                         # every source's segment of my receive buffer is complete
    n_all = world * cap
    srows = torch.empty(n_all, dtype=torch.int32, device=dev)
    spos = torch.empty(n_all, dtype=torch.int32, device=dev)
    n_local = max(shard_rows(n_rows, world, rank), 1)
    ws2_bytes = lib.rlctr_sort_ws_bytes(n_all, n_local)
    ws2 = torch.empty(ws2_bytes, dtype=torch.uint8, device=dev)
    _lib.call("rlctr_sort_routed_pos", lib.rlctr_sort_routed_pos, _lib.ptr(rb["keys"].local), n_all, n_local, _lib.ptr(srows),
              _lib.ptr(spos), _lib.ptr(ws2), ws2_bytes, st, key="rlctr_sort_ids", meta={"n": n_all})
    return {"rows": srows, "pos": spos, "n_all": n_all, "cap": cap, "route_ws": ws, "slot_of": rb["vals"].local}


# ---------------------------------------------------------------------------------------------------
class _MemberGeom:
    """Stand-in carrying the stand-alone geometry of a member kind for colocated._layout."""

    def __init__(self, kind, n, d):
        self._geom = Geometry.lr(n) if kind == "LR" else Geometry.fm(n, d)


class ShardedGroup(nn.Module):
    """Co-located records (colocated.py) over a row-sharded joint table: LR / FM / DeepFM trained on the same id stream share one
    384-byte record per id, ``owner(id) = id mod G``.  Per step and rank: ONE routed sorted view (shared_sorted_view), one
    catch-up of the owned records, one gather through the peer mapping (rlctr_group_fwd: one NVLink read per (sample, field)
    for all members instead of one per member), one all_gather of the per-sample rows [B, 32] = column sums + every member's
    dL/dlogit, one push of the tower-input gradients, one all_reduce of all dense gradients, two device barriers, one update
    (rlctr_group_rows_adam).  The update equals the single-GPU group step on the concatenated global batch."""

    SUMS_PITCH = 32                    # one 128-byte line per sample: S (<= 28 floats) | dL/dlogit of members 0..3

    def __init__(self, kinds, feature_nums, field_nums, latent_dims, group=None, device=None):
        super().__init__()
        from . import colocated as _co
        kinds = tuple(kinds)
        if not all(k in ("LR", "FM", "DeepFM") for k in kinds) or not 1 <= len(kinds) <= _lib.RLCTR_GROUP_MAX:
            raise _lib.RlctrError("ShardedGroup members: LR, FM, DeepFM (up to four)")
        self.kinds, self.feature_nums, self.field_nums, self.latent_dims = kinds, int(feature_nums), int(field_nums), int(latent_dims)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world not in (1, 2, 4, 8):
            raise _lib.RlctrError("row sharding supports 1, 2, 4 or 8 GPUs (owner = id & (G-1))")
        n_local = max(shard_rows(self.feature_nums, self.world, self.rank), 1)
        n_alloc = max(shard_rows(self.feature_nums, self.world, 0), 1)
        self._cols, used = _co._layout([_MemberGeom(k, n_local, self.latent_dims) for k in kinds])
        rs = (used + 1 + 3) // 4 * 4
        if rs > self.SUMS_PITCH - _lib.RLCTR_GROUP_MAX:
            raise _lib.RlctrError("joint row too wide for the per-sample exchange row (28 floats)")
        self._geom = Geometry(n_local, rs, -1, 0, used, pitch=96, block=32, stamp_at=used)
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._pm = PeerMemory(n_alloc * 96, torch.float32, dev, group)
        data = self._pm.local.view(n_alloc, 96)
        data.zero_()
        for lin, emb, dim in self._cols:
            cols = ([lin] if lin >= 0 else []) + list(range(emb, emb + dim))
            data[:n_local, cols] = torch.randn(n_local, len(cols), device=dev)
        self.table = nn.Parameter(data)
        self.table._rlctr_owner = self
        self.biases = nn.ParameterList([nn.Parameter(torch.zeros(1, device=dev)) for _ in kinds])
        self.mlps = nn.ModuleList([Model._tower(self.field_nums * self.latent_dims, dev) if k == "DeepFM" else nn.Identity()
                                   for k in kinds])
        self._opt, self._stash, self._ws, self._buf = None, None, {}, {}
        self.fork_ok = False

    # ---- protocol shared with optim.Adam / graphs ----------------------------------------------------------------------
    def _apply(self, fn, recurse=True):
        saved = self._parameters.pop("table")            # the shard lives in symmetric memory: .to() must not move it
        out = super()._apply(fn, recurse)
        self._parameters["table"] = saved
        self.table._rlctr_owner = self
        return out

    def _meta(self, B, F):
        g = self._geom
        return {"model": "Sharded" + "+".join(self.kinds), "B": B, "F": max(F, 1), "rs": g.row_stride, "dim": g.dim,
                "n_rows": g.n_rows, "lin": False, "members": [(k, c[2]) for k, c in zip(self.kinds, self._cols)]}

    def flush(self):
        if self._opt is not None:
            self._opt.flush(self.table.data)
        check_route_overflow()

    _reduce_ws = ShardedCTR._reduce_ws
    _rows_ws = Model._TableModel._rows_ws

    def zero_grad(self, set_to_none=True):
        self._stash = None
        return super().zero_grad(set_to_none)

    def barrier(self):
        self._pm.barrier()

    @classmethod
    def from_group(cls, cg, group=None):
        """Shard a single-GPU colocated.ColocatedCTR (same parameters on every rank) over the process group."""
        kinds = [type(m).__name__ for m in cg.members]
        m0 = cg.members[-1]
        self = cls(kinds, cg.feature_nums, getattr(m0, "field_nums", 15), max(getattr(m, "latent_dims", 1) for m in cg.members),
                   group=group, device=cg.table.device)
        assert self._cols == cg._cols
        cg.flush()
        with torch.no_grad():
            shard = cg.table.data[self.rank::self.world]
            self.table.data[:shard.shape[0]].copy_(shard)
            for i, m in enumerate(cg.members):
                self.biases[i].data.copy_(m.bias.data)
                if getattr(m, "mlp", None) is not None:
                    self.mlps[i].load_state_dict(m.mlp.state_dict())
        self.barrier()
        return self

    def gather_table(self):
        """The full joint table [N, 96], rebuilt on every rank (tests / checkpointing)."""
        self.flush()
        mine = self.table.data.contiguous()
        if self.world == 1:
            return mine[:self.feature_nums].clone()
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        full = torch.empty(self.feature_nums, 96, dtype=torch.float32, device=mine.device)
        for r in range(self.world):
            full[r::self.world] = parts[r][:shard_rows(self.feature_nums, self.world, r)]
        return full

    def _global_struct(self):
        g = self._geom
        t = _lib.Table(self.table.data.data_ptr(), self.feature_nums, g.row_stride, g.lin_col, g.emb_col, g.dim, g.row_pitch)
        if self.world > 1:
            t.world = self.world
            for r, p in enumerate(self._pm.ptrs):
                t.peers[r] = p
        return t

    def _exchange(self, B, F):
        """Per batch size: this rank's per-sample exchange rows [B, 32], their all-gathered copy [G*B, 32], and per tower member
        the symmetric receive buffer [G*n, D] the source ranks push their tower-input gradients into."""
        b = self._buf.get(B)
        if b is None:
            dev, G, D = self.table.device, self.world, self.latent_dims
            b = {"sums": torch.zeros(B, self.SUMS_PITCH, dtype=torch.float32, device=dev),
                 "sums_all": torch.zeros(G * B, self.SUMS_PITCH, dtype=torch.float32, device=dev), "recv": {}}
            for i, k in enumerate(self.kinds):
                if k == "DeepFM" and G > 1:                   # dense per-source segments: [G * cap, D] (routed push)
                    pm = PeerMemory(G * route_capacity(B * F, G) * D, torch.float32, dev, self.group)
                    pm.local.zero_()
                    b["recv"][i] = (pm, (C.c_void_p * 8)(*[int(p) for p in pm.ptrs]))
            self._buf[B] = b
        return b

    def _members(self, B, F, dev):
        arr = (_lib.Member * len(self.kinds))()
        logits, rows = [], []
        for i, (k, (lin, emb, dim)) in enumerate(zip(self.kinds, self._cols)):
            s = arr[i]
            s.lin_col, s.emb_col, s.dim = lin, emb, dim
            s.flags = _lib.RLCTR_FM_TERM if k in ("FM", "DeepFM") else 0
            s.bias = _lib.ptr(self.biases[i].data)
            z = torch.empty(B, dtype=torch.float32, device=dev)
            s.logit = _lib.ptr(z)
            logits.append(z)
            r = None
            if k == "DeepFM":
                pitch = (F * dim + 3) // 4 * 4
                r = torch.empty(B, pitch, dtype=torch.float32, device=dev)
                s.rows_out, s.rows_pitch = r.data_ptr(), pitch
            rows.append(r)
        return arr, logits, rows

    @torch.no_grad()
    def forward(self, x):
        """pCTR of every member, float32 [B, M]."""
        lib = _lib.load()
        x = Model._check_ids(x)
        B, F = x.shape
        self.flush()
        self.barrier()
        arr, logits, rows = self._members(B, F, x.device)
        t = self._global_struct()
        _lib.call("rlctr_group_fwd", lib.rlctr_group_fwd, _lib.ptr(x), C.byref(t), arr, len(self.kinds), None, 0, B, F, _lib.stream(),
                  key="rlctr_group_fwd[sharded infer]", meta=self._meta(B, F))
        out = []
        for i, k in enumerate(self.kinds):
            z = logits[i]
            if rows[i] is not None:
                z = z + self.mlps[i](rows[i][:, :F * self._cols[i][2]]).reshape(-1)
            out.append(torch.sigmoid(z))
        self.barrier()
        return torch.stack(out, dim=1)

    def train_step(self, features, labels, optimizer):
        """One step of every member on the local batch; returns the local mean BCE losses, float32 [M]."""
        lib = _lib.load()
        opt = self._opt
        if opt is None:
            raise _lib.RlctrError("build rl_ctr_prediction_b200.optim.Adam(group.parameters(), ...) before training the group")
        x = Model._check_ids(features)
        B, F = x.shape
        dev, st, G, M = x.device, _lib.stream(), self.world, len(self.kinds)
        y = labels.reshape(-1).contiguous()
        yi = y if y.dtype == torch.int64 else None
        yf = None if yi is not None else y.float()
        buf = self._exchange(B, F)
        view = None
        if G > 1:                                             # routed exchange: positions in the owner's receive buffers
            view = routed_view_pos(x, self.feature_nums, self.group)
            srows, sslots, n_all = view["rows"], view["pos"], view["n_all"]
        else:
            srows, sslots, n_all = shared_sorted_view(x, self.feature_nums, self.group)
        if opt.lazy and opt.dirty:
            t, a = table_struct(self.table.data, self._geom), opt.struct()
            _lib.call("rlctr_rows_catchup", lib.rlctr_rows_catchup, _lib.ptr(srows), n_all, C.byref(t), C.byref(a), st,
                      key="rlctr_rows_catchup[ShardedGroup]", meta=self._meta(n_all // F, F))
        self.barrier()                                        # B1: all owners caught their rows up | forward reads
        arr, logits, rows = self._members(B, F, dev)
        t = self._global_struct()
        sums = buf["sums"]
        _lib.call("rlctr_group_fwd", lib.rlctr_group_fwd, _lib.ptr(x), C.byref(t), arr, M, _lib.ptr(sums), self.SUMS_PITCH, B, F, st,
                  key="rlctr_group_fwd[ShardedGroup]", meta=dict(self._meta(B, F), n_rows=self.feature_nums))
        losses = torch.empty(M, dtype=torch.float32, device=dev)
        ws = self._reduce_ws(dev)
        for p in self.parameters():
            p.grad = None
        extras = [None] * M
        for i, k in enumerate(self.kinds):
            dl = torch.empty(B, dtype=torch.float32, device=dev)
            dbias = torch.empty(1, dtype=torch.float32, device=dev)
            z, tower_out, leaf = logits[i], None, None
            if rows[i] is not None:
                leaf = rows[i][:, :F * self._cols[i][2]].detach().requires_grad_(True)
                with torch.enable_grad():
                    tower_out = self.mlps[i](leaf).reshape(-1)
                z = z + tower_out.detach()
            _lib.check(lib.rlctr_bce_fwd_bwd(_lib.ptr(z), _lib.ptr(yi), _lib.ptr(yf), None, losses.data_ptr() + 4 * i, _lib.ptr(dl),
                                             _lib.ptr(dbias), _lib.ptr(ws), B, st), "rlctr_bce_fwd_bwd")
            if G > 1:                                         # gradient of the GLOBAL mean loss
                dl.mul_(1.0 / G)
                dbias.mul_(1.0 / G)
            sums[:, self.SUMS_PITCH - _lib.RLCTR_GROUP_MAX + i] = dl    # dL/dlogit rides in the sample's exchange row
            if tower_out is not None:
                tower_out.backward(dl)
                extras[i] = leaf.grad.contiguous()
            self.biases[i].grad = dbias
        if G > 1:
            dense = [p for p in self.parameters() if p is not self.table and p.grad is not None]
            flat = torch.cat([p.grad.reshape(-1) for p in dense])
            dist.all_reduce(flat, group=self.group)
            o = 0
            for p in dense:
                p.grad = flat[o:o + p.numel()].view_as(p).clone()
                o += p.numel()
            dist.all_gather_into_tensor(buf["sums_all"], sums, group=self.group)
            sums_all = buf["sums_all"]
            for i, (pm, ptrs) in buf["recv"].items():          # per-occurrence rows, in the routed bucket order: coalesced posted
                n = B * F                                      # writes into dense per-source segments of the owners' memory
                _lib.call("rlctr_push_rows_routed", lib.rlctr_push_rows_routed, _lib.ptr(view["route_ws"]), n, G, self.rank,
                          view["cap"], _lib.ptr(extras[i]), self.latent_dims, ptrs, st, key="rlctr_push_rows",
                          meta={"n": n, "width": self.latent_dims})
                extras[i] = pm.local
        else:
            sums_all = sums
        self.barrier()                                        # B2: every rank's gradient-side buffers are complete
        self._stash = {"sorted_ids": srows, "sorted_slots": sslots, "n": n_all, "fields": F, "sums": sums_all, "extra": extras,
                       "slot_of": view["slot_of"] if view is not None else None}
        optimizer.step()
        return losses

    def _group_update(self, stash, opt_state, st):
        lib = _lib.load()
        arr = (_lib.Member * len(self.kinds))()
        for i, (k, (lin, emb, dim)) in enumerate(zip(self.kinds, self._cols)):
            s = arr[i]
            s.lin_col, s.emb_col, s.dim = lin, emb, dim
            s.flags = _lib.RLCTR_FM_TERM if k in ("FM", "DeepFM") else 0
            s.dlogit = None                                   # in the exchange rows
            s.extra = _lib.ptr(stash["extra"][i])
        t, a = table_struct(self.table.data, self._geom), opt_state.struct()
        n, F = stash["n"], stash["fields"]
        ws_bytes = lib.rlctr_rows_ws_bytes(n)
        ws = self._rows_ws(ws_bytes)
        _lib.call("rlctr_group_rows_adam", lib.rlctr_group_rows_adam, _lib.ptr(stash["sorted_ids"]), _lib.ptr(stash["sorted_slots"]),
                  n, C.byref(t), C.byref(a), arr, len(self.kinds), _lib.ptr(stash["sums"]), self.SUMS_PITCH, F, self.world,
                  _lib.ptr(stash["slot_of"]), _lib.ptr(ws), ws_bytes, st, key="rlctr_group_rows_adam[ShardedGroup]", meta=self._meta(n // F, F))
