"""world_size-2 gloo tests (CPU) of the row-sharding host logic in rl_ctr_prediction_b200/sharded.py:
bucketing -> all-to-all #1 (ids) -> owner gather -> all-to-all #2 (rows) -> all-to-all #3 (row
gradients).  The device operations (bucket, gather) are injected as a host backend built from torch
ops -- the product backend is the CUDA C ABI and is covered by the -m gpu tests; what is checked here is
the routing: counts, split sizes, permutations and their inverses across ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class HostBackend:
    """Same contract as sharded.CudaBackend, on CPU tensors (TEST ONLY)."""

    def bucket(self, ids_flat, world, n_rows):
        n = ids_flat.numel()
        ok = (ids_flat >= 0) & (ids_flat < n_rows)
        owner = torch.where(ok, ids_flat % world, torch.zeros_like(ids_flat))
        order = torch.sort(owner, stable=True).indices
        send_local = torch.where(ok, ids_flat // world, torch.full_like(ids_flat, -1))[order]
        pos_of_slot = torch.empty(n, dtype=torch.int64)
        pos_of_slot[order] = torch.arange(n)
        counts = torch.bincount(owner, minlength=world)
        ends = torch.cumsum(counts, 0)
        ends = torch.where(counts > 0, ends, torch.full_like(ends, -1))
        return send_local, pos_of_slot, order.to(torch.int32), ends

    def gather(self, local_ids, table, geom):
        out = torch.zeros(local_ids.numel(), geom.row_stride)
        ok = (local_ids >= 0) & (local_ids < geom.n_rows)
        out[ok] = table[local_ids[ok]]
        return out


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
        from rl_ctr_prediction_b200.tables import Geometry
        torch.manual_seed(0)                                   # same full table on every rank
        full = torch.randn(N, rs)
        n_local = sharded.shard_rows(N, world, rank)
        shard = full[rank::world].contiguous()
        assert shard.shape[0] == n_local
        geom = Geometry(n_local, rs, 0, 1, rs - 1)
        g = torch.Generator().manual_seed(100 + rank)          # different batch per rank
        ids = torch.randint(0, N, (B, F), generator=g)
        ids[0, 0] = N + 5                                      # out-of-range id: zero row, no update
        ids[1, :] = ids[2, :]                                  # duplicates
        be = HostBackend()
        plan = sharded.exchange_plan(ids, N, None, be)
        assert sum(plan.send_counts) == B * F and plan.n_recv == sum(plan.recv_counts)
        # every received local row belongs to this shard
        ok = plan.recv_local >= 0
        assert int(plan.recv_local[ok].max()) < n_local
        rows = sharded.fetch_rows(plan, shard, geom, None, be)
        got = rows[plan.pos_of_slot].view(B, F, rs)            # back in slot order
        want = torch.zeros(B, F, rs)
        inr = ids < N
        want[inr] = full[ids[inr]]
        assert torch.equal(got, want)                          # bit-exact routed gather
        # gradients: send slot-indexed rows, the owner must see each against the right local row
        grad_slot = torch.randn(B * F, rs, generator=g)
        gbuf = grad_slot[plan.send_slots.long()]               # send-buffer order
        recv = sharded.push_grads(plan, gbuf, None)
        dense = torch.zeros(n_local, rs, dtype=torch.float64)
        dense.index_add_(0, plan.recv_local[ok], recv[ok].double())
        # reference: gather every rank's (ids, grads), keep the ones this rank owns
        all_ids = [torch.empty_like(ids) for _ in range(world)]
        all_g = [torch.empty_like(grad_slot) for _ in range(world)]
        dist.all_gather(all_ids, ids)
        dist.all_gather(all_g, grad_slot)
        ref = torch.zeros(n_local, rs, dtype=torch.float64)
        for i_, g_ in zip(all_ids, all_g):
            f = i_.reshape(-1)
            mine = (f < N) & (f % world == rank)
            ref.index_add_(0, f[mine] // world, g_[mine].double())
        assert torch.allclose(dense, ref, rtol=0, atol=1e-12)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,B,F", [(101, 64, 15), (7, 16, 3)])
def test_exchange_round_trip_world2(N, B, F):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), N, B, F, 12), nprocs=world, join=True)


def test_counts_from_ends_and_shard_rows():
    from rl_ctr_prediction_b200 import sharded
    assert sharded._counts_from_ends([3, -1, 10, -1], 10) == [3, 0, 7, 0]
    assert sharded._counts_from_ends([-1, -1], 0) == [0, 0]
    assert [sharded.shard_rows(10, 4, r) for r in range(4)] == [3, 3, 2, 2]
    assert sum(sharded.shard_rows(10_000_000, 8, r) for r in range(8)) == 10_000_000
