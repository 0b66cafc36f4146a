"""Secondary configurations of BASELINE.json on one GPU: ClipLoss fwd+bwd through the drop-in module, inputs resident,
CUDA events, L2 flushed between iterations.  Prints one JSON line per configuration."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import flyp_b200
import _inputs as torch_port

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
configs = [(512, 512, torch.bfloat16), (512, 512, torch.float32), (4096, 768, torch.bfloat16), (8192, 512, torch.bfloat16),
           (16384, 512, torch.bfloat16), (32768, 512, torch.bfloat16), (65536, 512, torch.bfloat16),
           (8192, 1024, torch.bfloat16), (16384, 1024, torch.bfloat16), (32768, 1024, torch.bfloat16),
           (32768, 512, torch.float32)]
if len(sys.argv) > 1:
    configs = [c for c in configs if str(c[0]) in sys.argv[1].split(",")]
for B, D, dt in configs:
    I, T = torch_port.synthetic_pairs(B, D, seed=0, dtype=dt)
    Id = I.to(dev).requires_grad_(True); Td = T.to(dev).requires_grad_(True)
    theta = torch.tensor(2.6592600369327783, device=dev, requires_grad=True)
    fn = flyp_b200.ClipLoss(cache_labels=True)
    def step():
        Id.grad = Td.grad = theta.grad = None
        loss = fn(Id, Td, theta.exp())
        loss.mean().backward()
        return loss
    for _ in range(5): step()
    torch.cuda.synchronize()
    iters = 30 if B <= 8192 else 10
    evs = []
    for _ in range(iters):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    print(json.dumps({"B": B, "D": D, "dtype": str(dt).split(".")[-1], "ms_fwd_bwd": round(ms, 4), "pairs_per_s": round(B / ms * 1e3),
                      "tflops_8B2D": round(8.0 * B * B * D / ms / 1e9, 1)}), flush=True)
    del Id, Td, I, T
    torch.cuda.empty_cache()
