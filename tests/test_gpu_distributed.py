"""Multi-GPU (NCCL) parity of the drop-in module for all four (local_loss, gather_with_grad) combinations against the
golden vectors recorded from the reference under gloo, at world sizes 2, 4 and 8 (each needs that many GPUs; skipped
otherwise).  The default loss runs over NVLink peer memory (multicast when the fabric offers it), local_loss over NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _worker(rank, world, port, name, local_loss, gwg, ret, fdt=torch.bfloat16):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from flyp_b200 import ClipLoss, gather_features
    z = np.load(os.path.join(GOLDEN, name))
    n = z["I"].shape[0]
    b = n // world
    I = torch.tensor(z["I"]).to(fdt); T = torch.tensor(z["T"]).to(fdt)
    Il = I[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
    Tl = T[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
    sc = torch.tensor(float(z["scale"]), device=dev, requires_grad=True)
    fn = ClipLoss(local_loss=local_loss, gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world)
    loss = fn(Il, Tl, sc)
    g = torch.tensor(z["g"][:loss.shape[0]] if local_loss else z["g"], device=dev, dtype=torch.float32)
    (loss.float() * g).sum().backward()
    gi, gt = gather_features(Il.detach(), Tl.detach(), local_loss, False, rank, world, False)
    torch.cuda.synchronize()
    ret[rank] = dict(loss=loss.detach().float().cpu().numpy(), dI=Il.grad.float().cpu().numpy(),
                     dT=Tl.grad.float().cpu().numpy(), ds=sc.grad.float().cpu().numpy(),
                     gathered_I=gi.float().cpu().numpy(), I_bf16=I.float().numpy())
    dist.destroy_process_group()


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


_PORT = [29911]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("local_loss,gwg", [(False, False), (False, True), (True, False), (True, True)])
def test_multi_gpu_semantics(local_loss, gwg, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if world > 2 and local_loss:
        pytest.skip("local_loss combinations are covered at world = 2 (NCCL path, independent of the world size)")
    from oracle import clip_oracle as orc
    name = "clip_w2_n264_d64.npz"               # 264 rows: 132 / 66 / 33 per rank (not multiples of the 128-row tile)
    mgr = mp.Manager()
    ret = mgr.dict()
    _PORT[0] += 1
    mp.spawn(_worker, args=(world, _PORT[0], name, local_loss, gwg, ret), nprocs=world, join=True)
    z = np.load(os.path.join(GOLDEN, name))
    n = z["I"].shape[0]
    b = n // world
    I = torch.tensor(z["I"]).bfloat16().double().numpy(); T = torch.tensor(z["T"]).bfloat16().double().numpy()
    Ib, Tb = [I[r * b:(r + 1) * b] for r in range(world)], [T[r * b:(r + 1) * b] for r in range(world)]
    # bf16 storage of loss / gradients by autograd on top of the 2e-3 bar; with local_loss the gradient of a bf16 leaf
    # is a bf16 sum of several bf16 terms (two cross-entropies, and the reduce-scattered share with gather_with_grad)
    tol = (3 if local_loss else 1) * 2.0 ** -8 + 2e-3
    for r in range(world):
        got = ret[r]
        assert np.array_equal(got["gathered_I"], got["I_bf16"])          # rank-major ordering, bit exact
        want = orc.clip_loss_distributed(Ib, Tb, float(z["scale"]), r, local_loss)
        g = z["g"][:b] if local_loss else z["g"]
        wI, wT, ws = orc.clip_loss_distributed_grads(Ib, Tb, float(z["scale"]), r, local_loss, gwg, g)
        assert got["loss"].shape == want.shape
        assert rel(got["loss"], want) < tol
        assert rel(got["dI"], wI) < tol and rel(got["dT"], wT) < tol
        assert abs(got["ds"] - ws) < tol * abs(ws)


@pytest.mark.parametrize("local_loss,gwg", [(False, False), (True, True)])
def test_multi_gpu_fp32_features(local_loss, gwg):
    """fp32 features (what FLYP trains in) over the peer-memory exchange on two real GPUs: 1e-5 / 1e-4."""
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import clip_oracle as orc
    name = "clip_w2_n264_d64.npz"
    mgr = mp.Manager()
    ret = mgr.dict()
    _PORT[0] += 1
    mp.spawn(_worker, args=(world, _PORT[0], name, local_loss, gwg, ret, torch.float32), nprocs=world, join=True)
    z = np.load(os.path.join(GOLDEN, name))
    n = z["I"].shape[0]
    b = n // world
    I = torch.tensor(z["I"]).float().double().numpy(); T = torch.tensor(z["T"]).float().double().numpy()
    Ib, Tb = [I[r * b:(r + 1) * b] for r in range(world)], [T[r * b:(r + 1) * b] for r in range(world)]
    for r in range(world):
        got = ret[r]
        want = orc.clip_loss_distributed(Ib, Tb, float(z["scale"]), r, local_loss)
        g = z["g"][:b] if local_loss else z["g"]
        wI, wT, ws = orc.clip_loss_distributed_grads(Ib, Tb, float(z["scale"]), r, local_loss, gwg, g)
        assert rel(got["loss"], want) < 1e-5
        assert rel(got["dI"], wI) < 1e-4 and rel(got["dT"], wT) < 1e-4
        assert abs(got["ds"] - ws) < 1e-4 * abs(ws)


def test_soak_changing_inputs_two_gpus():
    """The copy-engine flag ordering under changing inputs (every step pushes NEW bits into the same slots): bench.py's
    soak on two real GPUs - gathered matrices bit-exact after the consumers ran, loss = the rolled base loss."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    _PORT[0] += 1
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]), os.path.join(root, "bench.py"),
                          "--gpus", "2", "--batch", "4096", "--steps", "3", "--warmup", "3", "--soak", "2000"],
                         capture_output=True, text=True, timeout=900, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["check"]["passed"] and line["check"]["soak"]["passed"], line["check"]
    assert line["check"]["soak"]["steps"] == 2000 and line["check"]["soak"]["mismatching_gathered_elements"] == 0


def _kept_worker(rank, world, port, b, dim, gwg, steps, ret):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["FLYP_RS_MIN_ROWS"] = "0"          # (the default keeps the transposed sweep below 6144 rows per rank)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from flyp_b200 import ClipLoss, _lib
    from oracle import torch_port
    assert _lib.load().flyp_clip_keeps_ds(b, b * world, dim, _lib.FLYP_BF16) == 1
    fn = ClipLoss(gather_with_grad=gwg, cache_labels=True, rank=rank, world_size=world, grad_dtype=torch.float32)
    out = {}
    for step in range(steps):                     # several steps: both parities of the reduce-scatter buffers
        I, T = torch_port.synthetic_pairs(b * world, dim, seed=40 + step, dtype=torch.bfloat16)
        Il = I[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
        Tl = T[rank * b:(rank + 1) * b].to(dev).requires_grad_(True)
        sc = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
        g = (torch.rand(b * world, generator=torch.Generator().manual_seed(50 + step)) / (b * world)).to(dev)
        loss = fn(Il, Tl, sc)
        (loss.float() * g).sum().backward()
        torch.cuda.synchronize()
        out[step] = dict(loss=loss.detach().float().cpu().numpy(), dI=Il.grad.float().cpu().numpy(),
                         dT=Tl.grad.float().cpu().numpy(), ds=sc.grad.float().cpu().numpy())
    ret[rank] = out
    dist.destroy_process_group()


@pytest.mark.parametrize("world,b,dim", [(2, 1024, 256), (4, 1024, 128), (8, 1024, 128)])
@pytest.mark.parametrize("gwg", [False, True])
def test_multi_gpu_kept_ds_reduce_scatter(world, b, dim, gwg):
    """The kept-dS backward over real GPUs: the product kernel of every rank writes its fp32 partials of the text
    gradient into the owners' buffers over NVLink, each rank sums its W slots (csrc/clip_dst_gemm.cu, csrc/comm.cu)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import clip_oracle as orc
    from oracle import torch_port
    steps = 3
    mgr = mp.Manager()
    ret = mgr.dict()
    _PORT[0] += 1
    mp.spawn(_kept_worker, args=(world, _PORT[0], b, dim, gwg, steps, ret), nprocs=world, join=True)
    tol = 2.0 ** -8 + 2e-3                        # (autograd stores the gradients of bf16 leaves in bf16)
    for step in range(steps):
        I, T = torch_port.synthetic_pairs(b * world, dim, seed=40 + step, dtype=torch.bfloat16)
        In, Tn = I.double().numpy(), T.double().numpy()
        g = (torch.rand(b * world, generator=torch.Generator().manual_seed(50 + step)) / (b * world)).double().numpy()
        Ib, Tb = [In[r * b:(r + 1) * b] for r in range(world)], [Tn[r * b:(r + 1) * b] for r in range(world)]
        for r in range(world):
            got = ret[r][step]
            want = orc.clip_loss_distributed(Ib, Tb, 1 / 0.07, r, False)
            wI, wT, ws = orc.clip_loss_distributed_grads(Ib, Tb, 1 / 0.07, r, False, gwg, g)
            assert rel(got["loss"], want) < tol
            assert rel(got["dI"], wI) < tol and rel(got["dT"], wT) < tol, (step, r, rel(got["dI"], wI), rel(got["dT"], wT))
            assert abs(got["ds"] - ws) < tol * abs(ws)
