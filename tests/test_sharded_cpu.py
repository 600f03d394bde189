"""world_size-2 gloo tests (CPU) of the row-sharding protocol in rl_ctr_prediction_b200/sharded.py:
all_gather of the ids -> every owner keeps what it owns, sorted by local row in (source rank, slot) order ->
the owner pulls each occurrence's gradient from the source rank's buffer (here: an all_gather standing in for the
peer-mapped reads) and reduces per row.  The device kernels (rlctr_sort_ids_sharded, the peer reads of
rlctr_embed_fwd / rlctr_rows_adam) are covered by the -m gpu tests, which emulate G ranks on one GPU; what is
checked here is the routing: ownership, local rows, global slots and their decomposition, across real ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, B, F, rs):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rl_ctr_prediction_b200 import sharded
        torch.manual_seed(0)                                   # same full table on every rank
        full = torch.randn(N, rs)
        n_local = sharded.shard_rows(N, world, rank)
        shard = full[rank::world].contiguous()
        assert shard.shape[0] == n_local
        g = torch.Generator().manual_seed(100 + rank)          # different batch per rank
        ids = torch.randint(0, N, (B, F), generator=g)
        ids[0, 0] = N + 5                                      # out-of-range id: zero row, no update
        ids[1, :] = ids[2, :]                                  # duplicates
        n = B * F
        # ---- the one collective of the lookup: fixed-size all_gather of the ids
        all_ids = [torch.empty_like(ids) for _ in range(world)]
        dist.all_gather(all_ids, ids)
        ids_all = torch.cat([a.reshape(-1) for a in all_ids])
        rows, gslots = sharded.owned_sorted_view_host(ids_all, world, rank, N)
        # every kept id is owned by this rank, addressed by its local row, in stable (row, rank, slot) order
        assert bool(((ids_all[gslots] % world) == rank).all()) and bool((ids_all[gslots] // world == rows).all())
        assert int(rows.max()) < n_local
        key = rows * (world * n) + gslots
        assert bool((key[1:] > key[:-1]).all())
        # the union over the owners covers every in-range occurrence exactly once
        cnt = torch.tensor([gslots.numel()])
        dist.all_reduce(cnt)
        assert int(cnt) == int(((ids_all >= 0) & (ids_all < N)).sum())
        # ---- forward: row id is read at peers[id % G] + (id // G): emulate the peer tables with an all_gather
        n_max = sharded.shard_rows(N, world, 0)
        padded = torch.zeros(n_max, rs)
        padded[:n_local] = shard
        shards = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(shards, padded)
        flat = ids.reshape(-1)
        ok = flat < N
        got = torch.zeros(n, rs)
        for r in range(world):
            sel = ok & (flat % world == r)
            got[sel] = shards[r][flat[sel] // world]
        want = torch.zeros(n, rs)
        want[ok] = full[flat[ok]]
        assert torch.equal(got, want)                          # bit-exact gather through the sharded addressing
        # ---- backward: the owner pulls grad[src rank][slot] for global slot = src * n + slot
        grad_slot = torch.randn(n, rs, generator=g)
        all_g = [torch.empty_like(grad_slot) for _ in range(world)]
        dist.all_gather(all_g, grad_slot)                      # stands in for the peer-mapped gradient buffers
        src, slot = gslots // n, gslots % n
        pulled = torch.stack([all_g[int(s)][int(k)] for s, k in zip(src, slot)]) if gslots.numel() else torch.zeros(0, rs)
        dense = torch.zeros(n_local, rs, dtype=torch.float64)
        dense.index_add_(0, rows, pulled.double())
        ref = torch.zeros(n_local, rs, dtype=torch.float64)
        for i_, g_ in zip(all_ids, all_g):
            f = i_.reshape(-1)
            mine = (f < N) & (f % world == rank)
            ref.index_add_(0, f[mine] // world, g_[mine].double())
        assert torch.allclose(dense, ref, rtol=0, atol=1e-12)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,B,F", [(101, 64, 15), (7, 16, 3)])
def test_sharded_protocol_world2(N, B, F):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), N, B, F, 12), nprocs=world, join=True)


def test_shard_rows_and_owned_view():
    from rl_ctr_prediction_b200 import sharded
    assert [sharded.shard_rows(10, 4, r) for r in range(4)] == [3, 3, 2, 2]
    assert sum(sharded.shard_rows(10_000_000, 8, r) for r in range(8)) == 10_000_000
    ids = torch.tensor([5, 2, 9, 2, 100, 6, 1])
    rows, pos = sharded.owned_sorted_view_host(ids, 2, 0, 10)          # even ids < 10: 2, 2, 6 -> rows 1, 1, 3
    assert rows.tolist() == [1, 1, 3] and pos.tolist() == [1, 3, 5]
    rows, pos = sharded.owned_sorted_view_host(ids, 2, 1, 10)          # odd ids: 5, 9, 1 -> rows 2, 4, 0
    assert rows.tolist() == [0, 2, 4] and pos.tolist() == [6, 0, 2]


# ---- data-parallel policy nets (all_main.make_data_parallel): cross-rank BatchNorm + averaged gradients == one process ------
def _dp_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch.nn as nn
        from rl_ctr_prediction_b200 import all_main

        def net():
            torch.manual_seed(3)
            return nn.Sequential(nn.Linear(6, 5), nn.BatchNorm1d(5), nn.ReLU(), nn.Linear(5, 2))

        torch.manual_seed(10)
        X, Y = torch.randn(2 * 8, 6), torch.randn(2 * 8, 2)                  # the global batch, same on both ranks
        ref = net()
        ((ref(X) - Y) ** 2).mean().backward()                                # one process, whole batch
        dp = net()
        for m in dp.modules():
            if type(m) is nn.BatchNorm1d:
                m.__class__ = all_main.CrossRankBatchNorm1d
        xs, ys = X[rank * 8:(rank + 1) * 8], Y[rank * 8:(rank + 1) * 8]
        ((dp(xs) - ys) ** 2).mean().backward()                               # each rank: its slice, its local mean loss

        class Agent:                                                         # the hook make_data_parallel installs
            pass
        a, b = Agent(), Agent()
        a.eval_net = a.target_net = b.Actor = b.Critic = b.Actor_ = b.Critic_ = dp
        sync = all_main.make_data_parallel(a, b)
        sync(dp.parameters())
        for (k, p), (_, q) in zip(dp.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7), (k, (p.grad - q.grad).abs().max())
        assert torch.allclose(dp[1].running_mean, ref[1].running_mean, atol=1e-6)
        assert torch.allclose(dp[1].running_var, ref[1].running_var, atol=1e-6)
        assert int(dp[1].num_batches_tracked) == 1
    finally:
        dist.destroy_process_group()


def test_data_parallel_policy_nets_world2():
    mp.spawn(_dp_worker, args=(2, _free_port()), nprocs=2, join=True)


def _gather_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch.nn as nn
        from rl_ctr_prediction_b200 import all_main

        class Agent:
            batch_size = 8
        a, b = Agent(), Agent()
        torch.manual_seed(rank)                                              # replicas start different: rank 0's are broadcast
        net = nn.Linear(3, 2)
        a.eval_net = a.target_net = b.Actor = b.Critic = b.Actor_ = b.Critic_ = net
        gather = all_main.make_gathered_replay(a, b)
        assert a.batch_size == 4 and b.batch_size == 4
        ref = nn.Linear(3, 2)
        torch.manual_seed(0)
        ref = nn.Linear(3, 2)
        assert torch.equal(net.weight, ref.weight)
        ids = torch.arange(4 * 5).reshape(4, 5) + 1000 * rank + (1 << 40)    # ids beyond 2^24 survive the packing exactly
        r = torch.full((4, 1), 0.5 + rank)
        gi, gr = gather(ids, r)
        assert gi.dtype == torch.int64 and gi.shape == (8, 5) and gr.shape == (8, 1)
        want = torch.cat([torch.arange(20).reshape(4, 5) + 1000 * k + (1 << 40) for k in range(world)])
        assert torch.equal(gi, want) and torch.equal(gr, torch.tensor([[0.5]] * 4 + [[1.5]] * 4))
    finally:
        dist.destroy_process_group()


def test_gathered_replay_world2():
    mp.spawn(_gather_worker, args=(2, _free_port()), nprocs=2, join=True)
