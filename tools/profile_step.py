"""One forward + backward of the headline configuration through the C-ABI wrappers (for ncu / launch lists)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from flyp_b200 import ops
import _inputs as torch_port

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
I, T = torch_port.synthetic_pairs(B, D, dtype=torch.bfloat16)
I, T = I.to(dev), T.to(dev)
sc = torch.tensor([1 / 0.07], device=dev)
g = torch.full((B,), 1.0 / B, device=dev)
for _ in range(iters):
    row_lse, row_nll, col_stat, status = ops.clip_fwd_local(I, T, sc)
    col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, B)
    dI, dT, ds = ops.clip_bwd_local(I, T, sc, 0, row_lse, row_nll, col_lse, col_nll, g, g)
torch.cuda.synchronize()
print("loss mean", loss.mean().item(), "status", status.item())
