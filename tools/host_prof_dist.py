"""Host-side cost of the multi-rank ClipLoss step (cProfile on rank 0; run under torch.distributed.run)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist
import flyp_b200
import _inputs as torch_port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
b = B // world
I, T = torch_port.synthetic_pairs(B, D, dtype=torch.bfloat16)
Id = I[rank * b:(rank + 1) * b].to(dev).requires_grad_(True); Td = T[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
theta = torch.tensor(2.6592600369327783, device=dev, requires_grad=True)
fn = flyp_b200.ClipLoss(cache_labels=True, rank=rank, world_size=world)
def step():
    Id.grad = Td.grad = theta.grad = None
    loss = fn(Id, Td, theta.exp())
    loss.mean().backward()
for _ in range(20): step()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(200): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
if rank == 0:
    print(f"B={B} D={D} world={world}: host enqueue {(t1 - t0) / 200 * 1e6:.1f} us/step, wall {(time.perf_counter() - t0) / 200 * 1e6:.1f} us/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
if rank == 0:
    pstats.Stats(pr).sort_stats("tottime").print_stats(35)
dist.barrier(); dist.destroy_process_group()
