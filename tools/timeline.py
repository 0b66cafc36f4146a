"""GPU timeline of ClipLoss fwd+bwd steps through the drop-in module (torch.profiler / CUPTI): every kernel and memcpy of
the last profiled step with its start offset and duration, the gaps between them and the host time of the step.

    python tools/timeline.py [B] [D]                                   (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/timeline.py [B] [D]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import flyp_b200
import _inputs as torch_port

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
b = B // world
I, T = torch_port.synthetic_pairs(B, D, seed=0, dtype=torch.bfloat16)
Id = I[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
Td = T[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
theta = torch.tensor(2.6592600369327783, device=dev, requires_grad=True)
fn = flyp_b200.ClipLoss(cache_labels=True, rank=rank, world_size=world)


def step():
    Id.grad = Td.grad = theta.grad = None
    loss = fn(Id, Td, theta.exp())
    loss.float().mean().backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
# host time per step when the GPU queue is never waited on
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
if rank == 0:
    print(f"B={B} D={D} world={world}: host enqueue {(t1 - t0) / 20 * 1e3:.3f} ms/step, wall {(t2 - t0) / 20 * 1e3:.3f} ms/step")

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        step()
        torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # last step = events after the last large gap
    starts = [e.time_range.start for e in evs]
    cut = 0
    for i in range(1, len(evs)):
        if starts[i] - evs[i - 1].time_range.end > 1500:      # us; the synchronize + barrier between steps
            cut = i
    last = evs[cut:]
    # drop the barrier's own kernels at the head (nccl all-reduce of the barrier) if any: keep everything, label it
    t_first = last[0].time_range.start
    prev_end = t_first
    print(f"{'start_us':>9} {'dur_us':>9} {'gap_us':>8}  name")
    busy = 0.0
    for e in last:
        s, en = e.time_range.start, e.time_range.end
        print(f"{s - t_first:9.1f} {en - s:9.1f} {s - prev_end:8.1f}  {e.name[:110]}")
        busy += en - s
        prev_end = max(prev_end, en)
    print(f"span {prev_end - t_first:.1f} us, sum of durations {busy:.1f} us, {len(last)} device activities")
if world > 1:
    dist.destroy_process_group()
