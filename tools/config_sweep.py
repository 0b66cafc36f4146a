"""Secondary configurations of BASELINE.json on one GPU: ClipLoss fwd+bwd through the drop-in module next to the
reference's operator sequence in eager PyTorch on the same GPU (SURVEY 2a names it as the bar), inputs resident.
Per configuration: `ms` = median CUDA-event time of a step with the L2 flushed before it, `ms_b2b` = wall time per step
of 50 back-to-back steps (no flush, no host sync inside): the throughput a training loop sees, host time included.
Prints one JSON line per configuration.    python tools/config_sweep.py [B,B,...] [--no-eager]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F
import flyp_b200
import _inputs as torch_port

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
configs = [(256, 512, torch.bfloat16), (256, 512, torch.float32), (512, 512, torch.bfloat16), (512, 512, torch.float32),
           (4096, 768, torch.bfloat16), (4096, 768, torch.float32), (8192, 512, torch.bfloat16),
           (16384, 512, torch.bfloat16), (32768, 512, torch.bfloat16), (65536, 512, torch.bfloat16),
           (8192, 1024, torch.bfloat16), (16384, 1024, torch.bfloat16), (32768, 1024, torch.bfloat16),
           (65536, 1024, torch.bfloat16), (32768, 512, torch.float32)]
args = [a for a in sys.argv[1:] if not a.startswith("--")]
if args:
    configs = [c for c in configs if str(c[0]) in args[0].split(",")]
with_eager = "--no-eager" not in sys.argv


def eager_loss(I, T, s):
    li = s * I @ T.T
    lt = s * T @ I.T
    lab = torch.arange(li.shape[0], device=I.device, dtype=torch.long)
    return (F.cross_entropy(li, lab, reduction='none') + F.cross_entropy(lt, lab, reduction='none')) / 2


def measure(step, B):
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    iters = 30 if B <= 8192 else 10
    evs = []
    for _ in range(iters):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    n = 50 if B <= 8192 else 10
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    return ms, (time.perf_counter() - t0) / n * 1e3


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

    def step_eager():
        Id.grad = Td.grad = theta.grad = None
        loss = eager_loss(Id, Td, theta.exp())
        loss.mean().backward()
        return loss

    ms, b2b = measure(step, B)
    out = {"B": B, "D": D, "dtype": str(dt).split(".")[-1], "ms": round(ms, 4), "ms_b2b": round(b2b, 4),
           "pairs_per_s": round(B / ms * 1e3), "tflops_8B2D": round(8.0 * B * B * D / ms / 1e9, 1)}
    if with_eager and B * B * (4 if dt == torch.float32 else 2) * 10 < 100e9:
        try:
            ems, eb2b = measure(step_eager, B)
            out.update(eager_ms=round(ems, 4), eager_ms_b2b=round(eb2b, 4), speedup_vs_eager=round(eb2b / b2b, 2))
        except Exception as exc:  # noqa: BLE001
            out["eager_ms"] = f"failed: {type(exc).__name__}"
    print(json.dumps(out), flush=True)
    del Id, Td, I, T
    torch.cuda.empty_cache()
