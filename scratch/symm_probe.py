"""2-GPU probe: can torch symmetric memory give peer pointers + a device barrier here?  And how fast is a random 64 B
row gather over NVLink with the library's own gather kernel?"""
import os, sys, time, ctypes as C
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from rl_ctr_prediction_b200 import _lib
lib = _lib.load()
ok = {}
try:
    import torch.distributed._symmetric_memory as symm
    n_rows, rs = 2_000_000, 16
    t = symm.empty(n_rows * rs, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    t.copy_(torch.arange(n_rows * rs, device=dev, dtype=torch.float32) + rank * 1e9 % 7)
    t.fill_(float(rank + 1))
    hdl.barrier(channel=0)
    peer = (rank + 1) % world
    pb = hdl.get_buffer(peer, (n_rows, rs), torch.float32)
    ok["symm_peer_value"] = float(pb[5, 3].item())
    ok["buffer_ptrs"] = [hex(p) for p in hdl.buffer_ptrs]
    # random gather over NVLink with rlctr_gather_rows
    ids = torch.randint(0, n_rows, (983040,), device=dev, dtype=torch.int64)
    out = torch.empty(ids.numel(), rs, device=dev)
    for name, base in (("local", t), ("peer", pb)):
        tab = _lib.Table(base.data_ptr(), n_rows, rs, 0, 1, 10, 0)
        for _ in range(3):
            lib.rlctr_gather_rows(ids.data_ptr(), ids.numel(), C.byref(tab), out.data_ptr(), _lib.stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.rlctr_gather_rows(ids.data_ptr(), ids.numel(), C.byref(tab), out.data_ptr(), _lib.stream())
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        ok[f"gather_{name}_us"] = round(us, 1)
        ok[f"gather_{name}_GBps"] = round(ids.numel() * 64 / us / 1e3, 1)
        ok[f"gather_{name}_val"] = float(out[0, 0].item())
    # barrier cost
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(100):
        hdl.barrier(channel=0)
    e1.record(); torch.cuda.synchronize()
    ok["symm_barrier_us"] = round(e0.elapsed_time(e1) / 100 * 1e3, 2)
    one = torch.ones(1, device=dev)
    e0.record()
    for i in range(100):
        dist.all_reduce(one)
    e1.record(); torch.cuda.synchronize()
    ok["nccl_allreduce1_us"] = round(e0.elapsed_time(e1) / 100 * 1e3, 2)
    big = torch.empty(983040, dtype=torch.int32, device=dev)
    allb = torch.empty(world * 983040, dtype=torch.int32, device=dev)
    e0.record()
    for i in range(20):
        dist.all_gather_into_tensor(allb, big)
    e1.record(); torch.cuda.synchronize()
    ok["allgather_ids_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    # graph capture of barrier + all_gather
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=s):
        dist.all_gather_into_tensor(allb, big)
        hdl.barrier(channel=0)
    g.replay(); g.replay(); torch.cuda.synchronize()
    ok["graph_capture"] = True
except Exception as e:
    import traceback
    ok["symm_error"] = repr(e)[:400]
    traceback.print_exc()
print(rank, ok, flush=True)
dist.barrier()
dist.destroy_process_group()
