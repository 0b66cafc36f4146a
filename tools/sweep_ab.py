"""A/B timing of one backward sweep (dI + d(scale)) under different FLYP_SCHED_PAIRS settings, same process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from flyp_b200 import ops
import _inputs as torch_port
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
b = int(sys.argv[2]) if len(sys.argv) > 2 else B
settings = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0", "64", "0", "64"]
D = 512
dev = torch.device("cuda:0")
I, T = torch_port.synthetic_pairs(B, D, dtype=torch.bfloat16)
I, T = I.to(dev), T.to(dev)
sc = torch.tensor([1 / 0.07], device=dev)
g = torch.full((B,), 1.0 / B, device=dev)
row_lse, row_nll, col_stat, status = ops.clip_fwd_local(I, T, sc)
col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, B)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for setting in settings:
    os.environ["FLYP_SCHED_PAIRS"] = setting
    ts = []
    for it in range(8):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.clip_bwd_local(I[:b], T, sc, 0, row_lse[:b].contiguous(), row_nll[:b].contiguous(), col_lse, col_nll, g[:b].contiguous(), g, need_txt=False, need_scale=True)
        e1.record(); torch.cuda.synchronize()
        if it >= 2: ts.append(e0.elapsed_time(e1))
    print(f"B={B} rows={b} sched_pairs={setting:>3}: " + " ".join(f"{t:.3f}" for t in ts))
