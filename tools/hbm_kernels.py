"""Achieved HBM bandwidth of the elementwise / normalise kernels (north_star: "achieved HBM GB/s for the elementwise and
normalise kernels") against MEASURED_PEAKS.json: L2-normalise forward (2 n D e bytes) and backward (3 n D e bytes), bf16
and fp32, at n = 32768 .. 262144 rows of D = 512 (>= 2 x the 126 MB L2 for the larger ones), CUDA events, L2 flushed."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from flyp_b200 import ops

dev = torch.device("cuda:0")
peak = 6469.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for dt in (torch.bfloat16, torch.float32):
    e = 2 if dt == torch.bfloat16 else 4
    for n in (32768, 131072, 262144):
        D = 512
        x = torch.randn(n, D, device=dev).to(dt)
        y, inv = ops.l2norm_fwd(x)
        dy = torch.randn(n, D, device=dev).to(dt)
        ms_f = timed(lambda: ops.l2norm_fwd(x))
        ms_b = timed(lambda: ops.l2norm_bwd(y, dy, inv))
        gb_f = 2.0 * n * D * e / ms_f / 1e6
        gb_b = 3.0 * n * D * e / ms_b / 1e6
        print(json.dumps({"kernel": "l2norm", "dtype": str(dt).split(".")[-1], "n": n, "D": D,
                          "fwd_ms": round(ms_f, 4), "fwd_GBps": round(gb_f, 1), "fwd_frac_of_measured_copy": round(gb_f / peak, 3),
                          "bwd_ms": round(ms_b, 4), "bwd_GBps": round(gb_b, 1), "bwd_frac_of_measured_copy": round(gb_b / peak, 3)}),
              flush=True)
        del x, y, inv, dy
