"""A/B timing of the forward statistics pass: multicast cluster kernel vs single-CTA kernel (same process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from flyp_b200 import ops
import _inputs as torch_port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
I, T = torch_port.synthetic_pairs(B, D, dtype=torch.bfloat16)
I, T = I.to(dev), T.to(dev)
sc = torch.tensor([1 / 0.07], device=dev)
res = {}
for rep in range(3):
    for mc in ("1", "0"):
        os.environ["FLYP_FWD_MC"] = mc
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            row_lse, row_nll, col_stat, status = ops.clip_fwd_local(I, T, sc)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[mc] = (row_lse.clone(), col_stat.clone())
        print(f"mc={mc}", " ".join(f"{t:.3f}" for t in ts))
print("max |row_lse diff|", (res["1"][0] - res["0"][0]).abs().max().item(), "col_stat diff", (res["1"][1] - res["0"][1]).abs().max().item())
