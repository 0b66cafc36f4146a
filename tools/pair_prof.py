"""Wait-cycle breakdown of the CTA-pair backward sweep (cluster 0) - debugging aid, see flyp_debug_profile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from flyp_b200 import ops, _lib
import _inputs as torch_port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = 512
dev = torch.device("cuda:0")
I, T = torch_port.synthetic_pairs(B, D, dtype=torch.bfloat16)
I, T = I.to(dev), T.to(dev)
sc = torch.tensor([1 / 0.07], device=dev)
g = torch.full((B,), 1.0 / B, device=dev)
row_lse, row_nll, col_stat, status = ops.clip_fwd_local(I, T, sc)
col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, B)
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib = _lib.load()
for rep in range(2):
    lib.flyp_debug_profile(buf.data_ptr())
    ops.clip_bwd_local(I, T, sc, 0, row_lse, row_nll, col_lse, col_nll, g, g, need_txt=False, need_scale=False)
    torch.cuda.synchronize()
    lib.flyp_debug_profile(None)
v = buf.cpu().tolist()
tot = v[0]
print(f"MMA thread: total {tot} clk over {v[7]} steps = {tot / max(v[7],1):.0f} clk/step")
for name, x in zip(["sempty", "full_s", "dsfull", "full_t", "accempty", "ifull"], v[1:7]):
    print(f"   wait {name:9s} {x:12d} clk  {100 * x / tot:5.1f}%  ({x / max(v[7],1):.0f}/step)")
print(f"producer cta0: total {v[8]} wait_empty {v[9]} ({100*v[9]/max(v[8],1):.1f}%)   cta1: total {v[10]} wait_empty {v[11]} ({100*v[11]/max(v[10],1):.1f}%)")
print(f"epilogue t128: total {v[12]} wait_sfull {v[13]} ({100*v[13]/max(v[12],1):.1f}%) wait_dsempty {v[14]} ({100*v[14]/max(v[12],1):.1f}%) wait_accfull {v[15]} ({100*v[15]/max(v[12],1):.1f}%)")
