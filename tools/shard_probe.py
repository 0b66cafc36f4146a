"""Kernel times of ONE rank's share of the row-sharded loss, measured on a single GPU (no exchange): the forward sweep
over the row block [B/W, B] and the backward sweep of the same block, with the library's own events around the tcgen05
launches (flyp_debug_kernel_events).  Ideal = the 1-GPU kernel time / W.
    python tools/shard_probe.py [B] [D] [W,W,...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import flyp_b200
from flyp_b200 import _lib, ops
import _inputs as torch_port

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
D = int(sys.argv[2]) if len(sys.argv) > 2 else 512
worlds = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8]
dev = torch.device("cuda:0")
lib = flyp_b200.load()
I, T = torch_port.synthetic_pairs(B, D, seed=0, dtype=torch.bfloat16)
Id, Td = I.to(dev), T.to(dev)
sc = torch.tensor([1 / 0.07], device=dev)
g = torch.full((B,), 1.0 / B, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
row_lse, row_nll, col_stat, _ = ops.clip_fwd_local(Id, Td, sc)
col_lse, col_nll, _ = ops.clip_fwd_finish(col_stat, 1, row_nll, B)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for e in ev:
    e.record()
torch.cuda.synchronize()
base = {}
for W in worlds:
    b = B // W
    Ib = Id[:b].contiguous()
    rl, rn, gl = row_lse[:b].contiguous(), row_nll[:b].contiguous(), g[:b].contiguous()
    _lib.check(lib.flyp_debug_kernel_events(ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event, ev[3].cuda_event, 0))
    f, s = [], []
    for it in range(8):
        flush.fill_(1)
        ops.clip_fwd_local(Ib, Td, sc, 0)
        ops.clip_bwd_local(Ib, Td, sc, 0, rl, rn, col_lse, col_nll, gl, g, need_txt=False, need_scale=True)
        torch.cuda.synchronize()
        if it >= 3:
            f.append(ev[0].elapsed_time(ev[1])); s.append(ev[2].elapsed_time(ev[3]))
    lib.flyp_debug_kernel_events(None, None, None, None, 0)
    fm, sm = sorted(f)[len(f) // 2], sorted(s)[len(s) // 2]
    if W == 1:
        base = {"fwd": fm, "sweep": sm}
    out = {"B": B, "D": D, "world": W, "fwd_kernel_us": round(fm * 1e3, 1), "sweep_kernel_us": round(sm * 1e3, 1)}
    if base:
        out.update(fwd_vs_ideal=round(fm * W / base["fwd"], 3), sweep_vs_ideal=round(sm * W / base["sweep"], 3))
    print(json.dumps(out), flush=True)
