#!/usr/bin/env python
"""1-vs-G equivalence of the sharded path on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/dist_check.py

Every rank builds the same single-GPU model, shards it, trains G ranks x local batch B/G for a few steps,
and compares the gathered tables / dense parameters with the single-GPU model trained on the concatenated
global batch.  Prints one JSON line per model; exits non-zero on mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_ctr_prediction_b200 import optim, pretrain_main as PM, sharded  # noqa: E402


def within(errs):
    """Tables, biases: 2e-5 of the scale.  Tower weights: Adam normalises every element's update by its own gradient history,
    so a weight whose gradient is at rounding level moves by an ill-conditioned fraction of lr per step, and the G partial
    gradients are all-reduced in another order than the single-GPU sum: after 4 steps at lr = 1e-3 such elements differ by a
    few 1e-5 ABSOLUTE (weights ~0.08) while everything well-conditioned agrees to 1e-6 -- bound: 3e-4 of the scale."""
    return all(v <= (3e-4 if k.startswith("mlp.") and k.endswith("weight") else 2e-5) for k, v in errs.items())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N, F, D, B, STEPS = 5003, 15, 10, 256 * world, 4
    ok = True
    for name in ("LR", "FM", "FFM", "DeepFM"):
        torch.manual_seed(7)
        single = PM.get_model(name, N, F, D)
        with torch.no_grad():
            single.table.mul_(0.1)
        single.to(dev).eval()
        m = sharded.ShardedCTR.from_model(single)
        m.eval()
        opt_s = optim.Adam(single.parameters(), lr=1e-3, weight_decay=1e-5)
        opt_m = optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
        rng = np.random.default_rng(11)
        lossf = torch.nn.BCELoss()
        for s in range(STEPS):
            x = torch.as_tensor(rng.integers(0, N, size=(B, F))).to(dev)
            y = torch.as_tensor((rng.random(B) < 0.3).astype(np.int64)).to(dev)
            p = single(x)
            tl = lossf(p, y.unsqueeze(1).float())
            single.zero_grad()
            tl.backward()
            opt_s.step()
            lo, hi = rank * (B // world), (rank + 1) * (B // world)
            m.train_step(x[lo:hi].contiguous(), y[lo:hi].contiguous(), opt_m)
        single.flush()
        full = m.gather_table()
        scale = single.table.data.abs().max().item()
        err = (full - single.table.data).abs().max().item() / scale
        errs = {"table": err, "bias": (m.bias.data - single.bias.data).abs().max().item()}
        if m.mlp is not None:
            for (k, a), (_, b) in zip(m.mlp.state_dict().items(), single.mlp.state_dict().items()):
                errs["mlp." + k] = ((a - b).abs().max() / b.abs().max()).item()
        good = within(errs)
        ok = ok and good
        if rank == 0:
            print(json.dumps({"model": name, "world": world, "ok": good, "max_rel_err_vs_single_gpu": errs}))
    # ---- co-located group: ShardedGroup on G ranks x B/G == colocated.ColocatedCTR on the global batch
    from rl_ctr_prediction_b200 import colocated
    torch.manual_seed(9)
    members = [PM.get_model(n, N, F, D) for n in ("LR", "FM", "DeepFM")]
    for mm in members:
        with torch.no_grad():
            mm.table.mul_(0.1)
        mm.to(dev).train()
        if getattr(mm, "mlp", None) is not None:
            mm.mlp.eval()
    cg = colocated.colocate(members)
    sg = sharded.ShardedGroup.from_group(cg)
    for mlp in sg.mlps:
        mlp.eval()
    opt_s = optim.Adam(cg.parameters(), lr=1e-3, weight_decay=1e-5)
    opt_m = optim.Adam(sg.parameters(), lr=1e-3, weight_decay=1e-5)
    rng = np.random.default_rng(13)
    for s in range(STEPS):
        x = torch.as_tensor(rng.integers(0, N, size=(B, F))).to(dev)
        y = torch.as_tensor((rng.random(B) < 0.3).astype(np.int64)).to(dev)
        cg.train_step(x, y, opt_s)
        lo, hi = rank * (B // world), (rank + 1) * (B // world)
        sg.train_step(x[lo:hi].contiguous(), y[lo:hi].contiguous(), opt_m)
    cg.flush()
    full = sg.gather_table()
    scale = cg.table.data[:, :24].abs().max().item()
    errs = {"table": (full - cg.table.data)[:, :24].abs().max().item() / scale,
            "exp_avg_sq": ((full - cg.table.data)[:, 64:88].abs().max() / cg.table.data[:, 64:88].abs().max()).item()}
    for i, mm in enumerate(cg.members):
        errs[f"bias{i}"] = (sg.biases[i].data - mm.bias.data).abs().max().item()
        if getattr(mm, "mlp", None) is not None:
            for (k, a), (_, b) in zip(sg.mlps[i].state_dict().items(), mm.mlp.state_dict().items()):
                errs["mlp." + k] = ((a - b).abs().max() / b.abs().max()).item()
    good = within(errs)
    ok = ok and good
    if rank == 0:
        print(json.dumps({"model": "group(LR+FM+DeepFM)", "world": world, "ok": good, "max_rel_err_vs_single_gpu": errs}))
    # every rank has printed / compared; leave without tearing the symmetric-memory mappings down collectively
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
