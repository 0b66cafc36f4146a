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


def test_emulated_eight_ranks_at_the_benchmark_size():
    """The BASELINE multi-GPU configuration itself - global B = 32768, D = 512, 8 ranks of 4096 rows - emulated on ONE
    GPU (the driver's test box has one), phase by phase: gathered bits, the loss vector, and sampled rows of d image /
    d text of EVERY rank against the float64 reference of tools/sampled_check.py; gather_with_grad=True multiplies the
    feature gradients by the world size."""
    from flyp_b200 import comm as peer
    from flyp_b200.comm import PeerComm
    from oracle import torch_port
    from tools import sampled_check as sck
    dev = torch.device("cuda:0")
    world, b, dim = 8, 4096, 512
    B = world * b
    s = 1.0 / 0.07
    sc = torch.tensor([s], device=dev)
    comms = [PeerComm(r, world, b, dim, dev) for r in range(world)]
    PeerComm.connect_local(comms)
    try:
        I, T = torch_port.synthetic_pairs(B, dim, seed=3, dtype=torch.bfloat16)
        Id, Td = I.to(dev), T.to(dev)
        g = (torch.rand(B, generator=torch.Generator().manual_seed(11)) / B).to(dev)
        steps = [peer.fwd_gather(comms[r], Id[r * b:(r + 1) * b], Td[r * b:(r + 1) * b], sc) for r in range(world)]
        for st in steps:
            peer.fwd_local(st)
        losses = [peer.fwd_finish(st) for st in steps]
        grads = [peer.bwd_local(st, g, float(world) if r % 2 else 1.0, torch.float32, True, True, True)
                 for r, st in enumerate(steps)]
        seq = steps[0].g.seq
        for r in range(world):
            comms[r].push_scalar(seq, grads[r][2])
        ds_tot = [torch.empty(1, device=dev) for _ in range(world)]
        for r in range(world):
            comms[r].sum_scalar(seq, ds_tot[r])
        torch.cuda.synchronize()
        for c in comms:
            c.check_error()
        lse64 = sck.full_lse(Id, Td, s)
        gen = np.random.default_rng(5)
        for r, st in enumerate(steps):
            assert torch.equal(_view_bf16(st.g.txt_all, B, dim), Td)
            assert torch.equal(_view_bf16(st.g.img_all, B, dim), Id)
            assert torch.equal(losses[0], losses[r])
            loc = torch.tensor(gen.choice(b, 32, replace=False), device=dev)
            idx = loc + r * b
            want_loss, want_dI, want_dT = sck.sampled_reference(Id, Td, s, g, idx, lse=lse64)
            assert sck.row_errors(losses[r][idx], want_loss)[0] < 1e-5
            mul = float(world) if r % 2 else 1.0
            for got, want, what in ((grads[r][0][loc], mul * want_dI, "d image"), (grads[r][1][loc], mul * want_dT, "d text")):
                glob, per_row = sck.row_errors(got, want)
                assert glob < 2e-3 and per_row < 6e-3, (r, what, glob, per_row)
        assert len({t.item() for t in ds_tot}) == 1
        # d(scale) = sum over all rows of <dI_i, I_i> / s (ranks with the doubled gradients rescaled)
        tot = sum((grads[r][0].double() * Id[r * b:(r + 1) * b].double()).sum().item() / (float(world) if r % 2 else 1.0)
                  for r in range(world)) / s
        assert abs(ds_tot[0].item() - tot) < 1e-3 * abs(tot)
    finally:
        for c in comms:
            c.close()


@pytest.mark.parametrize("gwg", [False, True])
def test_local_loss_blocks_emulated(golden_dir, gwg):
    """local_loss=True (clip/loss.py:109-111,200-201) for both gather_with_grad settings with two emulated ranks on one
    GPU: the two one-directional cross-entropy blocks of every rank over the gathered matrices, the gathered-side
    gradients reduce-scattered by hand, against the float64 oracle (which is pinned to the reference's gloo runs)."""
    import os
    import flyp_b200
    from oracle import clip_oracle as orc
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(golden_dir, "clip_w2_n264_d64.npz"))
    world = 2
    I = torch.tensor(z["I"]).bfloat16(); T = torch.tensor(z["T"]).bfloat16()
    n = I.shape[0]
    b = n // world
    s = float(z["scale"])
    g = torch.tensor(z["g"][:b], dtype=torch.float32, device=dev)
    I_all = I.to(dev); T_all = T.to(dev)
    Ib = [I[r * b:(r + 1) * b].double().numpy() for r in range(world)]
    Tb = [T[r * b:(r + 1) * b].double().numpy() for r in range(world)]
    leaves, out = [], []
    for r in range(world):
        Il = I_all[r * b:(r + 1) * b].clone().requires_grad_(True)
        Tl = T_all[r * b:(r + 1) * b].clone().requires_grad_(True)
        # what gather_features hands the rank: the gathered matrices carry gradient only with gather_with_grad
        Ia = I_all.clone().requires_grad_(gwg); Ta = T_all.clone().requires_grad_(gwg)
        sc = torch.tensor(s, device=dev, requires_grad=True)
        li = flyp_b200.contrastive_cross_entropy(Il, Ta, sc, None, r * b, grad_dtype=torch.float32)
        lt = flyp_b200.contrastive_cross_entropy(Tl, Ia, sc, None, r * b, grad_dtype=torch.float32)
        loss = (li + lt) / 2
        (loss.float() * g).sum().backward()
        leaves.append((Il, Tl, Ia, Ta, sc)); out.append(loss)
    torch.cuda.synchronize()
    # bf16 storage of the gradients by autograd on top of the 2e-3 bar (several bf16 terms with gather_with_grad)
    tol = (3 if gwg else 1) * 2.0 ** -8 + 2e-3
    for r in range(world):
        Il, Tl, Ia, Ta, sc = leaves[r]
        want = orc.clip_loss_distributed(Ib, Tb, s, r, True)
        assert rel(out[r].detach().double().cpu().numpy(), want) < tol
        wI, wT, ws = orc.clip_loss_distributed_grads(Ib, Tb, s, r, True, gwg, z["g"][:b])
        dI = Il.grad.double(); dT = Tl.grad.double()
        if gwg:                      # reduce-scatter (sum over ranks) of the gradients w.r.t. the gathered matrices
            sl = slice(r * b, (r + 1) * b)
            for q in range(world):
                dI = dI + leaves[q][2].grad.double()[sl]
                dT = dT + leaves[q][3].grad.double()[sl]
        assert rel(dI.cpu().numpy(), wI) < tol and rel(dT.cpu().numpy(), wT) < tol
        assert abs(sc.grad.item() - ws) < tol * abs(ws)


def _view_f32(ptr, rows, cols):
    class _Raw32:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Raw32(ptr, rows * cols), device="cuda").reshape(rows, cols)


@pytest.mark.parametrize("world,b,dim", [(2, 132, 64), (4, 96, 256), (3, 200, 512)])
def test_emulated_ranks_fp32_features(world, b, dim):
    """fp32 features (what FLYP trains in) over the peer-memory exchange: the fp32 rows are pushed as they are and split
    into bf16 / fp16 planes on arrival; 1e-5 / 1e-4 against the float64 oracle, gathered bits exact."""
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
            I, T = torch_port.synthetic_pairs(B, dim, seed=10 + step, dtype=torch.float32)
            Id, Td = I.to(dev), T.to(dev)
            g = torch.rand(B, generator=torch.Generator().manual_seed(3 + step)).to(dev)
            steps = [peer.fwd_gather(comms[r], Id[r * b:(r + 1) * b], Td[r * b:(r + 1) * b], sc) for r in range(world)]
            for st in steps:
                peer.fwd_local(st)
            losses = [peer.fwd_finish(st) for st in steps]
            grads = [peer.bwd_local(st, g, 1.0, torch.float32, True, True, True) for st in steps]
            torch.cuda.synchronize()
            for c in comms:
                c.check_error()
            In, Tn = I.double().numpy(), T.double().numpy()
            want = orc.clip_loss(In, Tn, s)
            wI, wT, ws = orc.clip_loss_grads(In, Tn, s, g.double().cpu().numpy())
            ds = sum(gr[2].item() for gr in grads)
            assert abs(ds - ws) < 1e-4 * abs(ws)
            for r, st in enumerate(steps):
                assert torch.equal(_view_f32(st.g.txt_all, B, dim), Td)
                assert torch.equal(_view_f32(st.g.img_all, B, dim), Id)
                assert rel(losses[r].double().cpu().numpy(), want) < 1e-5
                sl = slice(r * b, (r + 1) * b)
                assert rel(grads[r][0].double().cpu().numpy(), wI[sl]) < 1e-4 * max(1.0, np.abs(wI).max() / np.abs(wI[sl]).max())
                assert rel(grads[r][1].double().cpu().numpy(), wT[sl]) < 1e-4 * max(1.0, np.abs(wT).max() / np.abs(wT[sl]).max())
    finally:
        for c in comms:
            c.close()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_local_loss_over_peer_memory_emulated(golden_dir, dtype):
    """local_loss=True through the peer-memory exchange (flyp_b200/comm.py local_fwd / local_bwd) with two emulated
    ranks: loss of the local rows, gradients of the local operands (gather_with_grad=False semantics) and - summed by
    hand over the ranks - the reduce-scattered gradients of the gathered operands (gather_with_grad=True)."""
    import os
    from flyp_b200 import comm as peer
    from flyp_b200.comm import PeerComm
    from oracle import clip_oracle as orc
    dev = torch.device("cuda:0")
    z = np.load(os.path.join(golden_dir, "clip_w2_n264_d64.npz"))
    world = 2
    I = torch.tensor(z["I"]).to(dtype); T = torch.tensor(z["T"]).to(dtype)
    n, dim = I.shape
    b = n // world
    s = float(z["scale"])
    sc = torch.tensor([s], device=dev)
    g = torch.tensor(z["g"][:b], dtype=torch.float32, device=dev)
    Ib = [I[r * b:(r + 1) * b].double().numpy() for r in range(world)]
    Tb = [T[r * b:(r + 1) * b].double().numpy() for r in range(world)]
    comms = [PeerComm(r, world, b, dim, dev) for r in range(world)]
    PeerComm.connect_local(comms)
    try:
        # phase 1: every rank's gather (pack + pushes); phase 2: the cross-entropy blocks; phase 3: backward
        Id, Td = I.to(dev), T.to(dev)
        gathered = [peer.local_gather(comms[r], Id[r * b:(r + 1) * b].contiguous(), Td[r * b:(r + 1) * b].contiguous(), sc)
                    for r in range(world)]
        fwd = [peer.local_compute(st) for st in gathered]
        bwd = [peer.local_bwd(f[2], g, torch.float32, True) for f in fwd]
        torch.cuda.synchronize()
        for c in comms:
            c.check_error()
        tol_l, tol_g = (1e-5, 1e-4) if dtype == torch.float32 else (2e-3, 2e-3)
        for r in range(world):
            li, lt, _ = fwd[r]
            want = orc.clip_loss_distributed(Ib, Tb, s, r, True)
            assert rel((0.5 * (li + lt)).double().cpu().numpy(), want) < tol_l
            d_img, d_txt, d_s, d_img_all, d_txt_all = bwd[r]
            wI, wT, ws = orc.clip_loss_distributed_grads(Ib, Tb, s, r, True, False, z["g"][:b])
            assert rel(d_img.double().cpu().numpy(), wI) < tol_g and rel(d_txt.double().cpu().numpy(), wT) < tol_g
            assert abs(d_s.item() - ws) < tol_g * abs(ws)
            wI2, wT2, _ = orc.clip_loss_distributed_grads(Ib, Tb, s, r, True, True, z["g"][:b])
            sl = slice(r * b, (r + 1) * b)
            gi = d_img.double() + sum(bwd[q][3].double()[sl] for q in range(world))
            gt = d_txt.double() + sum(bwd[q][4].double()[sl] for q in range(world))
            assert rel(gi.cpu().numpy(), wI2) < tol_g and rel(gt.cpu().numpy(), wT2) < tol_g
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
        st = peer.fwd_gather(c, Id, Td, sc); peer.fwd_local(st); peer.fwd_finish(st)   # a step's backward comes first
        g = torch.ones(b, device=dev)
        with pytest.raises(FlypError):
            peer.bwd_local(first, g, 1.0, torch.float32, True, True, True)
        peer.bwd_local(st, g, 1.0, torch.float32, True, True, True)       # the latest step is still intact
        torch.cuda.synchronize()
    finally:
        c.close()


@pytest.mark.parametrize("world,b,dim", [(2, 1024, 512), (4, 1024, 256), (2, 1100, 384), (8, 4096, 512), (2, 2048, 1024)])
def test_kept_ds_reduce_scatter_emulated(world, b, dim):
    """The multi-GPU kept-dS backward (csrc/clip_dst_gemm.cu + the reduce-scatter buffers of csrc/comm.cu) with W emulated
    ranks on ONE GPU, phase by phase: every rank's first sweep keeps its dS, its product dS_r^T . I_r writes the fp32
    partials of ALL text rows into the owners' buffers, then every rank sums its W slots.  Sampled rows of d image /
    d text of every rank and the all-reduced d(scale) against the float64 reference, over two steps (buffer parity),
    with the gather_with_grad factor on the odd ranks (it scales a rank's OWN rows: SURVEY 8c)."""
    from flyp_b200 import _lib, comm as peer
    from flyp_b200.comm import PeerComm
    from oracle import torch_port
    from tools import sampled_check as sck
    dev = torch.device("cuda:0")
    B = world * b
    assert _lib.load().flyp_clip_keeps_ds(b, B, dim, _lib.FLYP_BF16) == 1
    s = 1.0 / 0.07
    sc = torch.tensor([s], device=dev)
    comms = [PeerComm(r, world, b, dim, dev) for r in range(world)]
    PeerComm.connect_local(comms)
    for c in comms:
        c.set_rs_min_rows(0)          # (the default keeps the transposed sweep below 6144 rows per rank)
    try:
        for step in range(2):
            I, T = torch_port.synthetic_pairs(B, dim, seed=20 + step, dtype=torch.bfloat16)
            Id, Td = I.to(dev), T.to(dev)
            g = (torch.rand(B, generator=torch.Generator().manual_seed(13 + step)) / B).to(dev)
            steps = [peer.fwd_gather(comms[r], Id[r * b:(r + 1) * b], Td[r * b:(r + 1) * b], sc) for r in range(world)]
            for st in steps:
                peer.fwd_local(st)
            for st in steps:
                peer.fwd_finish(st)
            mul = [float(world) if r % 2 else 1.0 for r in range(world)]
            outs = [peer.bwd_step_phase(st, g, mul[r], torch.float32, 1) for r, st in enumerate(steps)]
            for r, st in enumerate(steps):
                peer.bwd_step_phase(st, g, mul[r], torch.float32, 2, outs[r])
            torch.cuda.synchronize()
            for c in comms:
                c.check_error()
            lse64 = sck.full_lse(Id, Td, s)
            gen = np.random.default_rng(5 + step)
            tot = 0.0
            for r in range(world):
                d_img, d_txt, ds = outs[r]
                loc = torch.tensor(gen.choice(b, 32, replace=False), device=dev)
                idx = loc + r * b
                _, want_dI, want_dT = sck.sampled_reference(Id, Td, s, g, idx, lse=lse64)
                for got, want, what in ((d_img[loc], mul[r] * want_dI, "d image"), (d_txt[loc], mul[r] * want_dT, "d text")):
                    glob, per_row = sck.row_errors(got, want)
                    assert glob < 2e-3 and per_row < 6e-3, (step, r, what, glob, per_row)
                tot += (d_txt.double() * Td[r * b:(r + 1) * b].double()).sum().item() / (mul[r] * s)
            # d(scale) = sum over all rows of <dT_j, T_j> / s; identical bits on every rank (fixed summation order)
            assert len({o[2][0].item() for o in outs}) == 1
            assert abs(outs[0][2][0].item() - tot) < 1e-3 * abs(tot)
    finally:
        for c in comms:
            c.close()
