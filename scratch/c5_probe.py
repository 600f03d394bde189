"""Where the C5 step (src/all_main step at batch 1M) spends its time: library calls (CUDA events) against the whole step."""
import sys, torch
sys.path.insert(0, ".")
import bench
from rl_ctr_prediction_b200 import _lib, all_main
from rl_ctr_prediction_b200.DDQN_model import RingMemory
from rl_ctr_prediction_b200.Feature_embedding import Feature_Embedding
dev = torch.device("cuda", 0)
N, D, F = 10_000_000, 10, 15
B = 1 << 20
md = bench._frozen_ctr_models(N, D, dev)
M = len(md)
torch.manual_seed(5)
fe = Feature_Embedding(N, F, D, device=dev)
ddqn, ddpg = all_main.get_model(M, N, F, D, 256, 1 << 21, dev, "1458")
RingMemory.device_sampling = True
ddqn.device_rng = ddpg.device_rng = True
gen = torch.Generator(device=dev).manual_seed(19)
batches = [bench.make_batch(gen, B, N, dev) for _ in range(3)]
def run(i):
    x, y = batches[i % 3]
    all_main.train_step(ddqn, ddpg, md, x, y.reshape(-1, 1), fe, 0.9, dev)
for i in range(2):
    run(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(4):
    run(i)
e1.record(); torch.cuda.synchronize()
print(f"step {e0.elapsed_time(e1) / 4:.2f} ms")
prof = _lib.KernelTimer(); _lib.set_timer(prof)
for i in range(2):
    run(i)
_lib.set_timer(None)
tot = 0.0
for k, (n, ms_, _) in sorted(prof.summary().items(), key=lambda kv: -kv[1][0] * kv[1][1]):
    print(f"{k:40s} {n:4d} x {ms_ * 1e3:9.1f} us = {n * ms_ / 2:8.3f} ms/step")
    tot += n * ms_ / 2
print(f"library calls {tot:.2f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    run(0); torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
