"""Small end-to-end cases for compute-sanitizer (memcheck): every kernel family once, ragged shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import flyp_b200
from flyp_b200 import ops
import _inputs as torch_port

dev = torch.device("cuda:0")
for (n, d, dt, impl) in [(300, 512, torch.bfloat16, "0"), (257, 640, torch.bfloat16, "2"), (200, 1024, torch.bfloat16, "2"),
                         (130, 768, torch.bfloat16, "1"), (150, 520, torch.bfloat16, "0"), (140, 256, torch.float32, "0")]:
    os.environ["FLYP_BWD_IMPL"] = impl
    I, T = torch_port.synthetic_pairs(n, d, seed=n, dtype=dt)
    Id = I.to(dev).requires_grad_(True); Td = T.to(dev).requires_grad_(True)
    th = torch.tensor(2.659, device=dev, requires_grad=True)
    loss = flyp_b200.ClipLoss(cache_labels=True)(Id, Td, th.exp())
    loss.float().mean().backward()
    torch.cuda.synchronize()
    print(n, d, dt, impl, float(loss.float().mean()), float(Id.grad.float().abs().max()))
os.environ["FLYP_BWD_IMPL"] = "0"
# rectangular CE head, fused argmax, normalise
a, b = torch_port.synthetic_pairs(300, 512, seed=3, dtype=torch.bfloat16)
lab = torch.randint(0, 182, (300,), device=dev)
A = a.to(dev).requires_grad_(True); Bm = b[:182].to(dev).requires_grad_(True)
l = flyp_b200.contrastive_cross_entropy(flyp_b200.l2_normalize(A), flyp_b200.l2_normalize(Bm), torch.tensor(14.0, device=dev), lab, reduction="mean")
l.backward(); torch.cuda.synchronize()
print("ce", float(l), ops.argmax(a.to(dev), b[:182].to(dev))[:4].tolist())
# emulated 3-rank peer step on one GPU
from flyp_b200 import comm as peer
from flyp_b200.comm import PeerComm
W, bb, D = 3, 100, 256
comms = [PeerComm(r, W, bb, D, dev) for r in range(W)]
PeerComm.connect_local(comms)
I, T = torch_port.synthetic_pairs(W * bb, D, seed=5, dtype=torch.bfloat16)
I, T = I.to(dev), T.to(dev)
sc = torch.tensor([14.0], device=dev)
g = torch.rand(W * bb, device=dev)
sts = [peer.fwd_gather(comms[r], I[r * bb:(r + 1) * bb], T[r * bb:(r + 1) * bb], sc) for r in range(W)]
for st in sts: peer.fwd_local(st)
ls = [peer.fwd_finish(st) for st in sts]
gr = [peer.bwd_local(st, g, 1.0, torch.float32, True, True, True) for st in sts]
torch.cuda.synchronize()
print("peer", float(ls[0].mean()), float(gr[1][0].abs().max()))
for c in comms: c.close()
