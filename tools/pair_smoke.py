import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from flyp_b200 import ops
from oracle import clip_oracle as orc
torch.manual_seed(0)
dev = "cuda:0"
def run(n, d, s=14.2857):
    I = torch.nn.functional.normalize(torch.randn(n, d), dim=-1)
    T = torch.nn.functional.normalize(0.5 * I + 0.5 * torch.nn.functional.normalize(torch.randn(n, d), dim=-1), dim=-1)
    I, T = I.bfloat16(), T.bfloat16()
    g = torch.rand(n) / n
    sc = torch.tensor([s], device=dev)
    Ic, Tc, gd = I.to(dev), T.to(dev), g.to(dev)
    row_lse, row_nll, col_stat, st = ops.clip_fwd_local(Ic, Tc, sc)
    col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, n)
    dI, dT, ds = ops.clip_bwd_local(Ic, Tc, sc, 0, row_lse, row_nll, col_lse, col_nll, gd, gd, grad_dtype=torch.float32)
    torch.cuda.synchronize()
    wI, wT, ws = orc.clip_loss_grads(I.double().numpy(), T.double().numpy(), s, g.double().numpy())
    rel = lambda a, b: float(np.abs(a.double().cpu().numpy() - b).max() / np.abs(b).max())
    print(f"n={n} d={d}: dI {rel(dI, wI):.3e} dT {rel(dT, wT):.3e} ds {abs(ds.item()-ws)/abs(ws):.3e}", flush=True)
for n, d in [(256, 512), (128, 128), (300, 256), (1000, 384), (4096, 512)]:
    run(n, d)
