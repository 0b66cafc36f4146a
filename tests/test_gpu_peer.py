"""Peer-memory exchange path (flyp_b200/comm.py, csrc/comm.cu) on ONE GPU: W emulated ranks, each with its own
communicator and exchange segment wired in-process (flyp_comm_connect_local), run phase by phase so that no kernel ever
waits for a kernel that has not been enqueued.  Checks the gathered ordering bit-exactly (clip/loss.py:66-67), the
loss vector, the gradients of the local rows and d(scale) of every rank against the float64 oracle of the full batch
(reference semantics for local_loss=False, gather_with_grad=False: SURVEY 8c), over two consecutive steps (double
buffering / sequence numbers), for block sizes that are not multiples of the 128-row tiles and for both sweep kernels.
The real multi-process run over NVLink is tests/test_gpu_distributed.py (needs >= 2 GPUs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Raw:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i2", "data": (int(ptr), False), "version": 2}


def _view_bf16(ptr, rows, cols):
    return torch.as_tensor(_Raw(ptr, rows * cols), device="cuda").view(torch.bfloat16).reshape(rows, cols)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("world,b,dim", [(2, 128, 512), (2, 132, 64), (4, 96, 256), (3, 200, 768), (8, 64, 128)])
def test_emulated_ranks_match_oracle(world, b, dim):
    from flyp_b200 import comm as peer
    from flyp_b200.comm import PeerComm
    from oracle import clip_oracle as orc
    from oracle import torch_port
    dev = torch.device("cuda:0")
    B = world * b
    s = 1.0 / 0.07
    sc = torch.tensor([s], device=dev)
    comms = [PeerComm(r, world, b, dim, dev) for r in range(world)]
    PeerComm.connect_local(comms)
    try:
        for step in range(2):
            I, T = torch_port.synthetic_pairs(B, dim, seed=step, dtype=torch.bfloat16)
            Id, Td = I.to(dev), T.to(dev)
            g = torch.rand(B, generator=torch.Generator().manual_seed(7 + step)).to(dev)
            steps = [peer.fwd_gather(comms[r], Id[r * b:(r + 1) * b], Td[r * b:(r + 1) * b], sc) for r in range(world)]
            for st in steps:
                peer.fwd_local(st)
            losses = [peer.fwd_finish(st) for st in steps]
            grads = [peer.bwd_local(st, g, 1.0, torch.float32, True, True, True) for st in steps]
            seq = steps[0].g.seq
            for r in range(world):
                comms[r].push_scalar(seq, grads[r][2])
            ds_tot = [torch.empty(1, device=dev) for _ in range(world)]
            for r in range(world):
                comms[r].sum_scalar(seq, ds_tot[r])
            torch.cuda.synchronize()
            for c in comms:
                c.check_error()

            In, Tn = I.double().numpy(), T.double().numpy()
            want = orc.clip_loss(In, Tn, s)
            wI, wT, ws = orc.clip_loss_grads(In, Tn, s, g.double().cpu().numpy())
            for r, st in enumerate(steps):
                # rank-major gathered ordering, bit exact, in every rank's segment
                assert torch.equal(_view_bf16(st.g.txt_all, B, dim), Td)
                assert torch.equal(_view_bf16(st.g.img_all, B, dim), Id)
                assert rel(losses[r].double().cpu().numpy(), want) < 2e-3
                sl = slice(r * b, (r + 1) * b)
                assert rel(grads[r][0].double().cpu().numpy(), wI[sl]) < 2e-3 * max(1.0, np.abs(wI).max() / np.abs(wI[sl]).max())
                assert rel(grads[r][1].double().cpu().numpy(), wT[sl]) < 2e-3 * max(1.0, np.abs(wT).max() / np.abs(wT[sl]).max())
                assert abs(ds_tot[r].item() - ws) < 2e-3 * abs(ws)
            # identical bits on every rank (fixed summation order)
            assert len({t.item() for t in ds_tot}) == 1
            assert all(torch.equal(losses[0], l) for l in losses[1:])
    finally:
        for c in comms:
            c.close()


def test_stale_step_is_refused():
    from flyp_b200 import comm as peer
    from flyp_b200.comm import PeerComm
    from flyp_b200 import FlypError
    from oracle import torch_port
    dev = torch.device("cuda:0")
    b, dim = 128, 128
    I, T = torch_port.synthetic_pairs(b, dim, seed=0, dtype=torch.bfloat16)
    Id, Td = I.to(dev), T.to(dev)
    sc = torch.tensor([10.0], device=dev)
    c = PeerComm(0, 1, b, dim, dev)
    try:
        first = peer.fwd_gather(c, Id, Td, sc); peer.fwd_local(first); peer.fwd_finish(first)
        for _ in range(2):
            st = peer.fwd_gather(c, Id, Td, sc); peer.fwd_local(st); peer.fwd_finish(st)
        g = torch.ones(b, device=dev)
        with pytest.raises(FlypError):
            peer.bwd_local(first, g, 1.0, torch.float32, True, True, True)
        peer.bwd_local(st, g, 1.0, torch.float32, True, True, True)       # the latest step is still intact
        torch.cuda.synchronize()
    finally:
        c.close()
