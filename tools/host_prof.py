"""Host-side cost of one ClipLoss fwd+bwd call (cProfile over many small steps; the GPU is never the bottleneck here)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import flyp_b200
import _inputs as torch_port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
I, T = torch_port.synthetic_pairs(B, D, dtype=torch.bfloat16)
Id = I.to(dev).requires_grad_(True); Td = T.to(dev).requires_grad_(True)
theta = torch.tensor(2.6592600369327783, device=dev, requires_grad=True)
fn = flyp_b200.ClipLoss(cache_labels=True)
def step():
    Id.grad = Td.grad = theta.grad = None
    loss = fn(Id, Td, theta.exp())
    loss.mean().backward()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print(f"B={B} D={D}: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us/step wall (host-bound)")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
