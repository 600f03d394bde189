"""Co-located records (rl_ctr_prediction_b200/colocated.py): LR + FM + DeepFM trained on one joint table must give the numbers
of the three stand-alone models -- FM and DeepFM bit for bit (same reduction trees, same arithmetic), LR within rounding of
its 15-term logit sum (the stand-alone scalar kernel sums in another order) -- and therefore the reference's
(src/main/pretrain_main.py:96-102 with dense torch.optim.Adam): the golden-trajectory test at the end pins that directly.
"""
import copy

import numpy as np
import pytest
import torch

from conftest import state_from_golden

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
F, D = 15, 10


def _models(N, seed=3, names=("LR", "FM", "DeepFM"), eval_tower=True):
    from rl_ctr_prediction_b200 import pretrain_main as PM
    torch.manual_seed(seed)
    ms = [PM.get_model(n, N, F, D).to(DEV) for n in names]
    for m in ms:
        m.train()
        if eval_tower and getattr(m, "mlp", None) is not None:
            m.mlp.eval()                      # no dropout: the towers of the two arms then see the same arithmetic
    return ms


def _batches(N, B, steps, seed=11, zipf=False):
    g = torch.Generator().manual_seed(seed)
    per = N // F
    out = []
    for _ in range(steps):
        if zipf:
            x = (torch.rand(B, F, generator=g) ** 6 * per).long().clamp_(0, per - 1)      # heavy hitters: long runs of one id
        else:
            x = torch.randint(0, per, (B, F), generator=g)
        x = x + torch.arange(F) * per
        y = (torch.rand(B, generator=g) < 0.3).long()
        out.append((x.to(DEV), y.to(DEV)))
    return out


def _train_separate(models, batches, mode="lazy"):
    from rl_ctr_prediction_b200 import graphs, optim
    opts = [optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5, mode=mode) for m in models]
    loss_fn = torch.nn.BCELoss()
    losses = []
    for x, y in batches:
        losses.append([float(graphs.eager_step(m, o, loss_fn, x, y)) for m, o in zip(models, opts)])
    return np.array(losses)


def _train_group(models, batches, mode="lazy"):
    from rl_ctr_prediction_b200 import colocated, optim
    group = colocated.colocate(models)
    opt = optim.Adam(group.parameters(), lr=1e-3, weight_decay=1e-5, mode=mode)
    losses = [group.train_step(x, y, opt).cpu().numpy().copy() for x, y in batches]
    return group, np.array(losses)


@pytest.mark.parametrize("catchup", ["ids", "sorted"])
@pytest.mark.parametrize("rows2", ["1", "0"])
@pytest.mark.parametrize("zipf", [False, True])
def test_group_equals_standalone_models(zipf, rows2, catchup, monkeypatch):
    from rl_ctr_prediction_b200 import colocated
    monkeypatch.setenv("RLCTR_GROUP_ROWS2", rows2)       # two lanes per record (default) / eight lanes per record
    # catch-up from the ids in batch order with claim bits, the sort on a side stream (default) / sort first, run heads
    monkeypatch.setattr(colocated, "UNSORTED_CATCHUP", catchup == "ids")
    N, B, steps = 3000, 512, 6
    sep = _models(N)
    grp_members = [copy.deepcopy(m) for m in sep]
    batches = _batches(N, B, steps, zipf=zipf)
    l_sep = _train_separate(sep, batches)
    group, l_grp = _train_group(grp_members, batches)
    assert group._geom.row_stride == 24 and group._geom.stamp_col == 23
    # FM and DeepFM: same bits, losses and every parameter of every row
    assert np.array_equal(l_sep[:, 1:], l_grp[:, 1:])
    np.testing.assert_allclose(l_grp[:, 0], l_sep[:, 0], rtol=2e-6)
    for i, name in enumerate(("LR", "FM", "DeepFM")):
        a, b = sep[i].state_dict(), grp_members[i].state_dict()
        assert set(a) == set(b)
        for k in a:
            if i == 0:
                torch.testing.assert_close(b[k], a[k], rtol=0, atol=2e-6), (name, k)
            else:
                assert torch.equal(a[k], b[k]), (name, k)


def test_group_inference_and_member_views():
    N, B = 2000, 300
    sep = _models(N)
    members = [copy.deepcopy(m) for m in sep]
    batches = _batches(N, B, 3)
    _train_separate(sep, batches)
    group, _ = _train_group(members, batches)
    x = batches[0][0]
    with torch.no_grad():
        for m in sep + members:
            m.eval()
        want = torch.cat([m(x) for m in sep], dim=1)
        got = group(x)                                   # one gather, [B, 3]
        per_member = torch.cat([m(x) for m in members], dim=1)
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-7)
    # a member called on its own runs the stand-alone gather kernel over the wider joint row (8 lanes per row: another
    # summation order for the 15-term sums -- rounding-level differences in logits of size ~10)
    torch.testing.assert_close(per_member, want, rtol=1e-4, atol=1e-5)
    assert torch.equal(got[:, 1], want[:, 1])            # FM: the group kernel reproduces the stand-alone tree
    # a member refuses to train on its own; checkpoints round-trip through the reference keys
    members[1].train()
    with pytest.raises(Exception):
        members[1](x)
    sd = {k: v.clone() for k, v in members[2].state_dict().items()}
    assert set(sd) == set(sep[2].state_dict())
    members[2].load_state_dict(sd)
    for k, v in members[2].state_dict().items():
        assert torch.equal(v, sd[k])


def test_group_lazy_equals_dense_and_long_staleness():
    """lazy + flush == dense Adam on the joint record (two lanes per row in the replay kernel), with rows that stay untouched
    for many steps."""
    N, B, steps = 6000, 64, 40
    a = _models(N, seed=5)
    b = [copy.deepcopy(m) for m in a]
    batches = _batches(N, B, steps, seed=2)
    ga, la = _train_group(a, batches, mode="lazy")
    gb, lb = _train_group(b, batches, mode="dense")
    assert np.array_equal(la, lb)
    ga.flush()
    gb.flush()
    assert torch.equal(ga.table.data[:, :24], gb.table.data[:, :24])
    assert torch.equal(ga.table.data[:, 32:56], gb.table.data[:, 32:56])
    assert torch.equal(ga.table.data[:, 64:88], gb.table.data[:, 64:88])


def test_group_graphed_step_equals_eager():
    from rl_ctr_prediction_b200 import colocated, graphs, optim
    N, B, steps = 4000, 256, 7
    a = _models(N, seed=9, eval_tower=False)             # train-mode dropout: the device RNG state is replayed too
    b = [copy.deepcopy(m) for m in a]
    batches = _batches(N, B, steps, seed=4)
    ga, gb = colocated.colocate(a), colocated.colocate(b)
    oa = optim.Adam(ga.parameters(), lr=1e-3, weight_decay=1e-5)
    ob = optim.Adam(gb.parameters(), lr=1e-3, weight_decay=1e-5)
    step = graphs.GraphedTrainStep([(ga, oa)])
    torch.manual_seed(77)                                # the towers draw their dropout seed at their first training forward
    la = [step(x, y)[0].clone() for x, y in batches]
    torch.manual_seed(77)
    lb = [gb.train_step(x, y, ob).clone() for x, y in batches]
    assert step.graph is not None
    for u, v in zip(la, lb):
        assert torch.equal(u, v)
    ga.flush()
    gb.flush()
    assert torch.equal(ga.table.data, gb.table.data)
    for pa, pb in zip(ga.parameters(), gb.parameters()):
        assert torch.equal(pa, pb)


@pytest.mark.parametrize("names", [("LR", "FM", "DeepFM"), ("FM", "LR"), ("DeepFM", "FM", "LR", "LR")])
def test_group_matches_reference_trajectory(golden, names):
    """The reference's own three training steps (tests/golden/make_golden.py: real p_model classes, dense torch.optim.Adam)
    for every member, trained as ONE group."""
    from test_gpu_models import assert_state, build, close, load
    from rl_ctr_prediction_b200 import colocated, optim
    xs, ys = golden["train/x"], golden["train/y"]
    members = []
    for n in names:
        m = load(build(n, 255), state_from_golden(golden, f"train/{n}/init")).to(DEV).train()
        if getattr(m, "mlp", None) is not None:
            m.mlp.eval()                                 # the golden trajectories were recorded with dropout off
        members.append(m)
    group = colocated.colocate(members)
    opt = optim.Adam(group.parameters(), lr=1e-3, weight_decay=1e-5)
    for s in range(3):
        losses = group.train_step(torch.as_tensor(xs[s]).to(DEV), torch.as_tensor(ys[s]).to(DEV), opt)
        for i, n in enumerate(names):
            close(losses[i], golden[f"train/{n}/loss{s}"])
    for m, n in zip(members, names):
        assert_state(m, state_from_golden(golden, f"train/{n}/final"))


def test_sharded_group_world1_equals_colocated():
    """sharded.ShardedGroup on one rank (the exchange-row layout: sums at a 32-float pitch with every member's dL/dlogit riding in
    it, the capped grid, the peer-table forward) == colocated.ColocatedCTR, bit for bit."""
    from rl_ctr_prediction_b200 import colocated, optim, sharded
    N, B, steps = 3001, 384, 5
    a = _models(N, seed=21)
    cg = colocated.colocate(a)
    sg = sharded.ShardedGroup.from_group(cg)
    for mlp in sg.mlps:
        mlp.eval()
    oa = optim.Adam(cg.parameters(), lr=1e-3, weight_decay=1e-5)
    ob = optim.Adam(sg.parameters(), lr=1e-3, weight_decay=1e-5)
    for x, y in _batches(N, B, steps, seed=6, zipf=True):
        la = cg.train_step(x, y, oa)
        lb = sg.train_step(x, y, ob)
        assert torch.equal(la, lb)
    cg.flush()
    assert torch.equal(sg.gather_table(), cg.table.data)
    for i, m in enumerate(cg.members):
        assert torch.equal(sg.biases[i].data, m.bias.data)
        if getattr(m, "mlp", None) is not None:
            for (k, u), (_, v) in zip(sg.mlps[i].state_dict().items(), m.mlp.state_dict().items()):
                assert torch.equal(u, v), k
    x = _batches(N, B, 1, seed=8)[0][0]
    with torch.no_grad():
        for m in cg.members:
            m.eval()
        assert torch.equal(sg(x), cg(x))


def test_group_update_with_receive_positions():
    """The routed exchange of the row-sharded group (rlctr_group_rows_adam(slot_of=...)): the sorted view carries RECEIVE
    positions r, the global slot is slot_of[r] and the dense-tail rows are stored by r.  Emulated on one GPU with a random
    permutation as the receive order: same bits as the plain update."""
    import ctypes as C
    from rl_ctr_prediction_b200 import _lib, colocated, optim, tables
    lib = _lib.load()
    N, B, steps = 3000, 256, 3
    a = _models(N, seed=31)
    b = [copy.deepcopy(m) for m in a]
    ga, gb = colocated.colocate(a), colocated.colocate(b)
    oa = optim.Adam(ga.parameters(), lr=1e-3, weight_decay=1e-5)
    ob = optim.Adam(gb.parameters(), lr=1e-3, weight_decay=1e-5)
    gen = torch.Generator(device=DEV).manual_seed(5)

    def permuted_update(stash, opt_state, st):
        n = stash.n
        slot_of = torch.randperm(n, generator=gen, device=DEV).to(torch.int32)          # receive position -> slot
        inv = torch.empty(n, dtype=torch.int64, device=DEV)
        inv[slot_of.long()] = torch.arange(n, device=DEV)
        spos = inv[stash.sorted_slots.long()].to(torch.int32)                             # the sorted view, as positions
        arr = (_lib.Member * len(gb.members))()
        keep = []
        for i, (m, (lin, emb, dim)) in enumerate(zip(gb.members, gb._cols)):
            s = arr[i]
            s.lin_col, s.emb_col, s.dim = lin, emb, dim
            s.flags = _lib.RLCTR_FM_TERM if m._fm_term else 0
            s.dlogit = _lib.ptr(stash.dlogit[i])
            if stash.extra[i] is not None:
                ex = stash.extra[i].reshape(n, dim)[slot_of.long()].contiguous()          # dense-tail rows in receive order
                keep.append(ex)
                s.extra = _lib.ptr(ex)
        t, ad = tables.table_struct(gb.table.data, gb._geom), opt_state.struct()
        wsb = lib.rlctr_rows_ws_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
        _lib.check(lib.rlctr_group_rows_adam(_lib.ptr(stash.sorted_ids), _lib.ptr(spos), n, C.byref(t), C.byref(ad), arr,
                                             len(gb.members), _lib.ptr(stash.sums), 0, stash.fields, 1, _lib.ptr(slot_of), _lib.ptr(ws),
                                             wsb, st), "rlctr_group_rows_adam")
        torch.cuda.synchronize()

    gb._group_update = permuted_update
    for x, y in _batches(N, B, steps, seed=12, zipf=True):
        la = ga.train_step(x, y, oa)
        lb = gb.train_step(x, y, ob)
        assert torch.equal(la, lb)
    ga.flush()
    gb.flush()
    assert torch.equal(ga.table.data, gb.table.data)


def test_group_with_more_than_sixteen_fields():
    """20 fields: the gather walks two 16-field blocks per sample (the non-pipelined path of group_fwd_kernel) and must still
    reproduce the stand-alone kernels' summation order: FM / DeepFM bit for bit."""
    from rl_ctr_prediction_b200 import colocated, graphs, optim, pretrain_main as PM
    F20, N, B, steps = 20, 4000, 300, 4
    torch.manual_seed(17)
    sep = [PM.get_model(n, N, F20, D).to(DEV).train() for n in ("LR", "FM", "DeepFM")]
    sep[2].mlp.eval()
    grp = [copy.deepcopy(m) for m in sep]
    g = torch.Generator().manual_seed(23)
    batches = [(torch.randint(0, N, (B, F20), generator=g).to(DEV), (torch.rand(B, generator=g) < 0.3).long().to(DEV))
               for _ in range(steps)]
    opts = [optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5) for m in sep]
    lossf = torch.nn.BCELoss()
    l_sep = np.array([[float(graphs.eager_step(m, o, lossf, x, y)) for m, o in zip(sep, opts)] for x, y in batches])
    group = colocated.colocate(grp)
    gopt = optim.Adam(group.parameters(), lr=1e-3, weight_decay=1e-5)
    l_grp = np.array([group.train_step(x, y, gopt).cpu().numpy() for x, y in batches])
    assert np.array_equal(l_sep[:, 1:], l_grp[:, 1:])
    np.testing.assert_allclose(l_grp[:, 0], l_sep[:, 0], rtol=2e-6)
    for i in (1, 2):
        a, b = sep[i].state_dict(), grp[i].state_dict()
        for k in a:
            assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("B", [1, 33])
def test_group_ragged_batches_and_out_of_range_ids(B):
    """Tiny / odd batch sizes and ids outside the vocabulary (zero row, no update -- include/rlctr.h conventions): the group
    follows the stand-alone models there too."""
    N, steps = 500, 3
    sep = _models(N, seed=41)
    grp = [copy.deepcopy(m) for m in sep]
    g = torch.Generator().manual_seed(3)
    batches = []
    for _ in range(steps):
        x = torch.randint(0, N, (B, F), generator=g)
        x[0, 0] = N + 7                                   # out of range
        x[-1, 3] = -1
        batches.append((x.to(DEV), (torch.rand(B, generator=g) < 0.5).long().to(DEV)))
    l_sep = _train_separate(sep, batches)
    group, l_grp = _train_group(grp, batches)
    assert np.array_equal(l_sep[:, 1:], l_grp[:, 1:])
    np.testing.assert_allclose(l_grp[:, 0], l_sep[:, 0], rtol=2e-6)
    for i in (1, 2):
        a, b = sep[i].state_dict(), grp[i].state_dict()
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_group_at_benchmark_batch_against_cpu_oracle():
    """The benchmark's own path and batch (B = 65536; N = 1e6): two steps of LR + FM + DeepFM as ONE co-located group against the
    CPU oracle port (three stock-ATen models, dense torch Adam over every row): every loss, every touched row, 10^4 untouched
    rows of every table -- and, at this size, FM / DeepFM still bit-identical to the stand-alone CUDA models."""
    from oracle import torch_port as TP
    from rl_ctr_prediction_b200 import colocated, graphs, optim, pretrain_main as PM
    N, B = 1_000_000, 65536
    names = ("LR", "FM", "DeepFM")
    ports, members, alone = [], [], []
    for n in names:
        torch.manual_seed(1)
        port = TP.PortCTR(n, N, F, D).eval()                 # eval(): DeepFM's dropout off
        with torch.no_grad():
            for k, p in port.named_parameters():
                if "embedding" in k or k == "linear.weight":
                    p.mul_(0.1)
        ports.append((port, TP.make_adam(port)))
        for bag in (members, alone):
            m = PM.get_model(n, N, F, D)
            m.load_state_dict(port.state_dict())
            bag.append(m.to(DEV).eval())
    group = colocated.colocate(members)
    gopt = optim.Adam(group.parameters(), lr=1e-3, weight_decay=1e-5)
    aopts = [optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5) for m in alone]
    lossf = torch.nn.BCELoss()
    rng = np.random.default_rng(0)
    per = N // F
    touched = []
    for s in range(2):
        x = torch.as_tensor(rng.integers(0, per, size=(B, F)) + np.arange(F) * per)
        y = torch.as_tensor((rng.random(B) < 0.05).astype(np.int64))
        touched.append(x.reshape(-1))
        ref = [TP.ctr_train_step(port, popt, lossf, x, y.unsqueeze(1)) for port, popt in ports]
        got = group.train_step(x.to(DEV), y.to(DEV), gopt).cpu().numpy()
        sep = [float(graphs.eager_step(m, o, lossf, x.to(DEV), y.to(DEV))) for m, o in zip(alone, aopts)]
        for i in range(3):
            assert abs(got[i] - ref[i]) <= 1e-5 * abs(ref[i]), (s, names[i], got[i], ref[i])
        assert got[1] == np.float32(sep[1]) and got[2] == np.float32(sep[2])
    rows = torch.unique(torch.cat(touched))
    extra = torch.as_tensor(rng.choice(N, 10_000, replace=False))
    for i, (m, (port, _)) in enumerate(zip(members, ports)):
        ref_sd, sd, sd_alone = port.state_dict(), m.state_dict(), alone[i].state_dict()
        for k, v in ref_sd.items():
            if i > 0:
                assert torch.equal(sd[k], sd_alone[k]), (names[i], k)          # == the stand-alone CUDA model, bit for bit
            if not (v.dim() == 2 and v.shape[0] == N):
                continue                                                          # dense parameters: covered by the equality above
            got_t = sd[k].cpu()
            for sel in (rows, extra):
                a, b = got_t[sel].double().numpy(), v[sel].double().numpy()
                bad = np.abs(a - b) > 2e-5 * np.abs(b) + 2e-5 * float(np.abs(b).max())
                if names[i] == "DeepFM":      # tower-borne gradients at rounding level: see test_gpu_parity_scale.py
                    assert bad.mean() <= 1e-4 and np.abs(a - b).max() <= 2.2 * 1e-3 * 2
                else:
                    assert not bad.any(), (names[i], k, int(bad.sum()), float(np.abs(a - b).max()))
