"""Co-located records: several CTR models trained on the same id stream share ONE row record per id.

The reference trains its models one after the other on the same encoded batches (``get_model`` names in
src/main/pretrain_main.py:25-45, loop :96-103) and the RL ensemble scores M of them on every sample
(src/all_main/main.py:183-271); each keeps its own ``nn.Embedding`` tables, so a batch costs one random gather and one
random optimizer update PER MODEL per (sample, field).  On the B200 a random HBM access costs the same whether it returns
4 or 128 bytes (profiles/r2_rowprobe.md: ~30-36 G accesses/s for 16 .. 128-byte rows), so the models' parameters of one id
are laid side by side in one record

    [ FM: w v0..v9 | LR: w ][ DeepFM: w v0..v9 | stamp ] (pad)     p block       (one 128-byte line)
    [ exp_avg of the same columns                       ] (pad)     exp_avg block
    [ exp_avg_sq                                        ] (pad)     exp_avg_sq block

and one gather (rlctr_group_fwd), one catch-up (rlctr_rows_catchup) and one update (rlctr_group_rows_adam) serve all of
them: 13 random line operations per (sample, field) and step instead of 23 for LR + FM + DeepFM.  Each member keeps the
chunk alignment of its stand-alone row, so its logits and its updated parameters are bit-identical to the stand-alone
model's (tests/test_gpu_colocated.py).

    lr, fm, dfm = (get_model(name, ...).to(dev) for name in ("LR", "FM", "DeepFM"))
    group = colocate([lr, fm, dfm])                       # re-homes the three tables into one; the members stay usable
    opt = optim.Adam(group.parameters(), lr=1e-3, weight_decay=1e-5)
    losses = group.train_step(x, y, opt)                  # float32 [3]: one BCE loss per member
    fm(x); fm.state_dict()                                # inference and reference-keyed checkpoints work per member
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib
from . import p_model as Model
from .tables import Geometry, round4, table_struct


# 1: the group's catch-up works from the ids in batch order and the sort of the batch runs on a side stream (see train_step);
# 0 (default): sort first, catch-up over the sorted run heads.  The two leave the same bits in the table.  Measured A/B on one
# B200 (C2 step): 1.347 ms sorted, 1.422 ms with the sort off the critical path -- the batch-order catch-up with its claim
# atomics and the concurrent sort cost more than the 66 us of sorting they hide, so it stays an option.
UNSORTED_CATCHUP = os.environ.get("RLCTR_GROUP_UNSORTED_CATCHUP", "0") != "0"


class GroupStash:
    """What a group step leaves for the optimizer: the sorted view, the joint column sums and each member's gradient side."""
    __slots__ = ("sorted_ids", "sorted_slots", "n", "fields", "sums", "dlogit", "extra")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


def _layout(members):
    """Columns of the joint row.  Vector members (FM-type rows: [w, v_0 .. v_{D-1}]) take consecutive 16-byte aligned blocks
    with their stand-alone layout; scalar members (LR) fill the padding columns those blocks leave, else open a new chunk.
    Returns ([(lin_col, emb_col, dim)] per member, used columns)."""
    cols, taken, nxt = [None] * len(members), set(), 0
    for i, m in enumerate(members):
        g = m._geom
        if g.row_stride == 1:
            continue
        if g.row_stride > 16:
            raise _lib.RlctrError(f"{type(m).__name__}: rows wider than 16 floats (FFM) cannot be co-located")
        lin = nxt + g.lin_col if g.lin_col >= 0 else -1
        cols[i] = (lin, nxt + g.emb_col, g.dim)
        for c in ([lin] if lin >= 0 else []) + list(range(nxt + g.emb_col, nxt + g.emb_col + g.dim)):
            taken.add(c)
        nxt += round4(g.used)
    for i, m in enumerate(members):
        if m._geom.row_stride != 1:
            continue
        free = [c for c in range(nxt) if c not in taken]
        c = free[0] if free else nxt
        if not free:
            nxt += 4
        taken.add(c)
        cols[i] = (c, 0, 0)
    used = max(taken) + 1
    if used % 4 == 0:
        used += 1                       # the stamp needs a padding column in the last active chunk: open one
    if round4(used + 1) > 32:
        raise _lib.RlctrError("co-located record wider than 32 floats (one 128-byte line): too many members")
    return cols, used


class ColocatedCTR(Model._TableModel):
    """``colocate(models)``: LR / FM-type drop-in models over one joint table.  The group owns the table, its lazy-exact
    Adam state and the training step; the members keep ``forward`` (inference), ``state_dict`` / ``load_state_dict``
    (reference keys) and their dense parameters."""

    _kind = "group"

    def __init__(self, models):
        super().__init__()
        models = list(models)
        if not 1 <= len(models) <= _lib.RLCTR_GROUP_MAX:
            raise ValueError(f"a co-located group holds 1..{_lib.RLCTR_GROUP_MAX} models")
        for m in models:
            # LR / FM: table only; DeepFM / W&D: logit = table part + tower(gathered rows) -- the forms the group step computes
            if type(m) not in (Model.LR, Model.FM, Model.DeepFM, Model.WideAndDeep) or m.__dict__.get("_group") is not None:
                raise _lib.RlctrError(f"{type(m).__name__} cannot join a co-located group (LR, FM, DeepFM, WideAndDeep; once)")
        n_rows = {m._geom.n_rows for m in models}
        devs = {m.table.device for m in models}
        if len(n_rows) != 1 or len(devs) != 1:
            raise ValueError("members of a group share the vocabulary (feature_nums) and the device")
        dev = devs.pop()
        if dev.type != "cuda":
            raise _lib.RlctrError("move the models to the CUDA device before co-locating them")
        self.feature_nums = n_rows.pop()
        cols, used = _layout(models)
        rs = round4(used + 1)
        self._geom = Geometry(self.feature_nums, rs, -1, 0, used, pitch=96, block=32, stamp_at=used)
        joint = torch.zeros(self.feature_nums, 96, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for m, (lin, emb, dim) in zip(models, cols):
                m.flush()                                    # settle what a previous optimizer still owes the stand-alone table
                g, src = m._geom, m.table.data
                if lin >= 0:
                    joint[:, lin].copy_(src[:, g.lin_col])
                if dim:
                    joint[:, emb:emb + dim].copy_(src[:, g.emb_col:g.emb_col + dim])
        self.table = nn.Parameter(joint)
        self.table._rlctr_owner = self
        self._opt, self._stash, self._ws = None, None, {}
        self._sort_stream, self._claim = None, None
        self._cols = cols
        for m, (lin, emb, dim) in zip(models, cols):
            m.table = self.table                             # shared Parameter: the member's view of the joint record
            m._geom = Geometry(self.feature_nums, rs, lin, emb, dim, pitch=96, block=32)
            m._opt, m._stash = None, None
            m.__dict__["_group"] = self                      # not a sub-module of the member: no cycle in parameters()
        self.members = nn.ModuleList(models)

    # ---- nn.Module plumbing ---------------------------------------------------------------------------------------
    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self.table._rlctr_owner = self
        return out

    def _ref_items(self):
        return []

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        self.flush()                                         # the members write the reference keys from the joint table

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        self.flush()

    def _meta(self, B, F):
        g = self._geom
        return {"model": "+".join(type(m).__name__ for m in self.members), "B": B, "F": F, "rs": g.row_stride, "dim": g.dim,
                "n_rows": g.n_rows, "lin": False, "members": [(type(m).__name__, c[2]) for m, c in zip(self.members, self._cols)]}

    # ---- forward of every member from one gather ---------------------------------------------------------------------
    def _member_structs(self, B, F, dev, train, pctr=None):
        arr = (_lib.Member * len(self.members))()
        logits, rows = [], []
        for i, (m, (lin, emb, dim)) in enumerate(zip(self.members, self._cols)):
            s = arr[i]
            s.lin_col, s.emb_col, s.dim = lin, emb, dim
            s.flags = _lib.RLCTR_FM_TERM if m._fm_term else 0
            bias = getattr(m, "bias", None)
            s.bias = _lib.ptr(bias.data) if bias is not None else None
            tower = getattr(m, "mlp", None) is not None
            z = torch.empty(B, dtype=torch.float32, device=dev) if (train or tower) else None
            logits.append(z)
            s.logit = _lib.ptr(z)
            if pctr is not None and not tower:
                s.pctr, s.pctr_stride = pctr.data_ptr() + 4 * i, pctr.shape[1]
            r = None
            if tower:
                pitch = round4(F * dim)                      # 16-byte aligned pitch: the tower's first GEMM fetches it by TMA
                r = torch.empty(B, pitch, dtype=torch.float32, device=dev)
                s.rows_out, s.rows_pitch = r.data_ptr(), pitch
            rows.append(r)
        return arr, logits, rows

    @torch.no_grad()
    def forward(self, x):
        """pCTR of every member on one gather: float32 [B, M] (column m = ``members[m](x)``)."""
        lib = _lib.load()
        x = Model._check_ids(x)
        B, F = x.shape
        self.flush()
        out = torch.empty(B, len(self.members), dtype=torch.float32, device=x.device)
        arr, logits, rows = self._member_structs(B, F, x.device, False, pctr=out)
        t = table_struct(self.table.data, self._geom)
        _lib.call("rlctr_group_fwd", lib.rlctr_group_fwd, _lib.ptr(x), C.byref(t), arr, len(self.members), None, 0, B, F,
                  _lib.stream(), key="rlctr_group_fwd[infer]", meta=self._meta(B, F))
        for i, m in enumerate(self.members):
            if rows[i] is not None:
                fd = F * self._cols[i][2]
                out[:, i:i + 1] = torch.sigmoid(logits[i].view(B, 1) + m.mlp(rows[i][:, :fd]))
        return out

    # ---- the training step ---------------------------------------------------------------------------------------
    def train_step(self, x, y, optimizer):
        """One step of src/main/pretrain_main.py:96-102 for EVERY member on the batch (x, y): sort -> catch-up -> one gather ->
        per member: (tower,) sigmoid + BCE and their gradient -> one segment-reduce + Adam over the joint record -> dense Adam.
        Returns the members' losses, float32 [M] on the device."""
        lib = _lib.load()
        opt = self._opt
        if opt is None:
            raise _lib.RlctrError("build rl_ctr_prediction_b200.optim.Adam(group.parameters(), ...) before training the group")
        x = Model._check_ids(x)
        B, F = x.shape
        dev, st, g = x.device, _lib.stream(), self._geom
        yy = y.reshape(-1).contiguous()
        yi = yy if yy.dtype == torch.int64 else None
        yf = None if yi is not None else yy.float()
        t = table_struct(self.table.data, g)
        # The sorted view is needed by the scatter at the END of the step only: with the catch-up working from the ids in batch
        # order (rlctr_rows_catchup_ids: a claim bit per row tells the occurrences of an id apart) the sort runs on its own
        # stream -- a parallel branch of a captured step -- beside catch-up, gather and tower instead of in front of them.
        side = None
        if UNSORTED_CATCHUP and opt.lazy and opt.stamp_col >= 0 and (g.used + 3) // 4 >= 5:    # joint records of 5..8 chunks
            cur = torch.cuda.current_stream(dev)
            side = self._sort_stream
            if side is None or side.device != dev:
                side = self._sort_stream = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                sid, sslot = Model.sort_ids(x, g.n_rows)
            if opt.dirty:
                a = opt.struct()
                cb = lib.rlctr_rows_claim_bytes(g.n_rows)
                if self._claim is None or self._claim.numel() < cb or self._claim.device != dev:
                    self._claim = torch.empty(cb, dtype=torch.uint8, device=dev)
                _lib.call("rlctr_rows_catchup_ids", lib.rlctr_rows_catchup_ids, _lib.ptr(x), x.numel(), C.byref(t), C.byref(a),
                          _lib.ptr(self._claim), cb, st, key="rlctr_rows_catchup[group]", meta=self._meta(B, F))
        else:
            sid, sslot = Model.sort_ids(x, g.n_rows)
            if opt.lazy and opt.dirty:
                a = opt.struct()
                _lib.call("rlctr_rows_catchup", lib.rlctr_rows_catchup, _lib.ptr(sid), x.numel(), C.byref(t), C.byref(a), st,
                          key="rlctr_rows_catchup[group]", meta=self._meta(B, F))
        M = len(self.members)
        arr, logits, rows = self._member_structs(B, F, dev, True)
        sums = torch.empty(B, g.row_stride, dtype=torch.float32, device=dev)
        _lib.call("rlctr_group_fwd", lib.rlctr_group_fwd, _lib.ptr(x), C.byref(t), arr, M, _lib.ptr(sums), 0, B, F, st,
                  key="rlctr_group_fwd[train]", meta=self._meta(B, F))
        losses = torch.empty(M, dtype=torch.float32, device=dev)
        dlogits, extras = [], []
        ws = self._reduce_ws(dev)
        for i, m in enumerate(self.members):
            dl = torch.empty(B, dtype=torch.float32, device=dev)
            dbias = torch.empty(1, dtype=torch.float32, device=dev)
            bias = getattr(m, "bias", None)
            if rows[i] is None:                              # LR, FM: the loss head and its gradient in one call
                _lib.check(lib.rlctr_bce_fwd_bwd(_lib.ptr(logits[i]), _lib.ptr(yi), _lib.ptr(yf), None,
                                                 losses.data_ptr() + 4 * i, _lib.ptr(dl), _lib.ptr(dbias), _lib.ptr(ws), B, st),
                           "rlctr_bce_fwd_bwd")
                extras.append(None)
            else:                                            # dense tail: autograd over the tower only
                fd = F * self._cols[i][2]
                leaf = rows[i][:, :fd].detach().requires_grad_(True)
                z = logits[i].view(B, 1) + m.mlp(leaf)
                zz = z.detach().reshape(-1).contiguous()
                # (the head's own sum of dlogit is the bias gradient: the same block partition and tree as the stand-alone
                # backward's rlctr_sigmoid_bwd sum -- the same bits, one launch less)
                _lib.check(lib.rlctr_bce_fwd_bwd(_lib.ptr(zz), _lib.ptr(yi), _lib.ptr(yf), None, losses.data_ptr() + 4 * i,
                                                 _lib.ptr(dl), _lib.ptr(dbias) if bias is not None else None, _lib.ptr(ws), B, st),
                           "rlctr_bce_fwd_bwd")
                m.zero_grad()
                z.backward(dl.view_as(z))
                extras.append(leaf.grad.contiguous())
            if bias is not None:
                bias.grad = dbias
            dlogits.append(dl)
        self._stash = GroupStash(sorted_ids=sid, sorted_slots=sslot, n=B * F, fields=F, sums=sums, dlogit=dlogits, extra=extras)
        if side is not None:
            torch.cuda.current_stream(dev).wait_stream(side)     # join: the scatter reads the sorted view
        optimizer.step()
        return losses

    def _group_update(self, stash, opt_state, st):
        """optim.Adam.step() for the joint table: rlctr_group_rows_adam."""
        lib = _lib.load()
        arr = (_lib.Member * len(self.members))()
        for i, (m, (lin, emb, dim)) in enumerate(zip(self.members, self._cols)):
            s = arr[i]
            s.lin_col, s.emb_col, s.dim = lin, emb, dim
            s.flags = _lib.RLCTR_FM_TERM if m._fm_term else 0
            s.dlogit = _lib.ptr(stash.dlogit[i])
            s.extra = _lib.ptr(stash.extra[i])
        t, a = table_struct(self.table.data, self._geom), opt_state.struct()
        ws_bytes = lib.rlctr_rows_ws_bytes(stash.n)
        ws = self._rows_ws(ws_bytes)
        _lib.call("rlctr_group_rows_adam", lib.rlctr_group_rows_adam, _lib.ptr(stash.sorted_ids), _lib.ptr(stash.sorted_slots),
                  stash.n, C.byref(t), C.byref(a), arr, len(self.members), _lib.ptr(stash.sums), 0, stash.fields, 1, None,
                  _lib.ptr(ws), ws_bytes, st, key="rlctr_group_rows_adam", meta=self._meta(stash.n // stash.fields, stash.fields))


def colocate(models) -> ColocatedCTR:
    return ColocatedCTR(models)
