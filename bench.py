#!/usr/bin/env python
"""Headline benchmark: ClipLoss forward + backward pairs/s at global B = 32768, D = 512, bf16 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dim D] [--dtype bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one forward (per-item loss vector) + one backward of loss.mean() through the drop-in module.
`value`      inputs resident in HBM (row block of this rank), CUDA-event timed, max over ranks, L2 flushed between steps.
`e2e`        the same call with HOST inputs: pinned H2D of both feature blocks and D2H of the loss vector inside the
             timed region.
`roofline`   the dominant kernel (the tcgen05 backward sweep `bwd_pair_kernel`), timed ALONE with CUDA events that the
             library records around its launch inside real module steps (flyp_debug_kernel_events); credited with its
             ALGORITHMIC FLOPs and compared with the BURST bf16 peak of MEASURED_PEAKS.json (kernel timed in isolation).
             The kept-dS backward runs ONE sweep (4 b B D: the recompute + dS . T, all algorithmic) and the d-text
             product over the kept dS (`dst_gemm_kernel`, 2 b B D, reported beside it; on several GPUs its epilogue is
             the scatter half of the reduce-scatter of the text gradient).  Shapes that do not keep dS run two sweeps,
             3 b B D credited per launch (half of the one credited recompute + one output product).
`step_frac`  8 B^2 D / ms_per_step / n_gpus / burst peak: the whole step against the tensor-core roofline.
`check`      sampled rows of d image AND d text (32 per rank) and the loss against a float64 reference on the GPU
             (tools/sampled_check.py), at every world size; for n_gpus > 1 also a soak over changing inputs.
`cpu_baseline` / `--impl reference`: the reference's own clip/loss.py (staged unmodified into oracle/_ref by
             oracle/stage_ref.py, kind "reference"; the torch-CPU port when it was never staged) on the host cores, on a
             bounded sample, with the B^2 cost law measured, not assumed.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ClipLoss fwd+bwd pairs/s at B=32k,D=512, 1/2/4/8 B200; % of tensor-core peak"
UNIT = "pairs/s"
THETA0 = 2.6592600369327783            # ln(1/0.07), clip/model.py:299


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(burst=1590.0, sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------------ clocks sampling
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            # median of the upper half = clocks under load (the sampler also sees the idle gaps between steps)
            load = sm[len(sm) // 2:]
            out.update(sm_mhz=load[len(load) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def synthetic_pairs(n, d, seed=0, dtype=None):
    """BASELINE.md section 3 inputs: I = normalize(randn), T = normalize(0.5 I + 0.5 normalize(randn))."""
    import torch
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=gen)
    y = torch.randn(n, d, generator=gen)
    I = F.normalize(x, dim=-1)
    T = F.normalize(0.5 * I + 0.5 * F.normalize(y, dim=-1), dim=-1)
    dtype = dtype or torch.bfloat16
    return I.to(dtype), T.to(dtype)


# ------------------------------------------------------------------------------------------------ reference arm
def _reference_loss_fn():
    """(callable(I, T, scale) -> per-item loss, kind): the staged reference clip/loss.py, else the torch-CPU port."""
    from oracle import stage_ref
    mod = stage_ref.load_reference_module()
    if mod is not None:
        fn = mod.ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=0, world_size=1)
        return fn, "reference"
    from oracle import torch_port
    return torch_port.clip_loss_reference_ops, "port"


def _cpu_step_seconds(loss_fn, n, d, steps, warmup):
    import torch
    I, T = synthetic_pairs(n, d, dtype=torch.float32)
    I.requires_grad_(True); T.requires_grad_(True)
    theta = torch.tensor(THETA0, requires_grad=True)
    times = []
    for it in range(warmup + steps):
        I.grad = T.grad = theta.grad = None
        t0 = time.perf_counter()
        loss = loss_fn(I, T, theta.exp())
        loss.mean().backward()                       # src/models/flyp_loss.py:498-499
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def cpu_reference_measure(batch, dim, steps, warmup, budget_s):
    """Time the reference's CPU ClipLoss fwd+bwd on the host cores within ~budget_s seconds.  The sample batch is the
    largest of {batch, batch/2, ...} whose estimated cost fits (and whose ~10 B x B fp32 temporaries fit in host RAM);
    the cost law used to extrapolate to `batch` is MEASURED from two sample sizes, not assumed."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    loss_fn, kind = _reference_loss_fn()
    try:
        import psutil
        ram = psutil.virtual_memory().available
    except Exception:
        ram = 32 << 30
    probe = min(batch, 2048)
    t_probe = _cpu_step_seconds(loss_fn, probe, dim, 1, 1)
    sample = batch
    while sample > 1024 and (t_probe * (sample / probe) ** 2 * (steps + warmup) > budget_s
                             or 10 * sample * sample * 4 > 0.6 * ram):
        sample //= 2
    sec = _cpu_step_seconds(loss_fn, sample, dim, steps, warmup)
    law = None
    factor = 1.0
    if sample != batch:
        # exponent of the cost law from the sample and half of it (the materialising algorithm: ~2)
        sec_half = _cpu_step_seconds(loss_fn, sample // 2, dim, max(1, steps), 1)
        import math
        law = math.log(sec / sec_half, 2)
        factor = (batch / sample) ** law
    sec_full = sec * factor
    what = "unmodified reference clip/loss.py (oracle/_ref)" if kind == "reference" else "torch-CPU port of clip/loss.py:117-118,208-209"
    desc = f"{what}, ClipLoss(cache_labels=True) fwd + mean().backward(), fp32, B={sample} D={dim}, {steps} steps after {warmup} warm-up"
    if sample != batch:
        desc += f"; extrapolated to B={batch} with the measured cost law t ~ B^{law:.2f} (B={sample // 2}: {sec_half * 1e3:.0f} ms, B={sample}: {sec * 1e3:.0f} ms)"
    return dict(value=batch / sec_full, unit=UNIT, cores=torch.get_num_threads(), kind=kind, sample=desc,
                sec_per_step=sec_full, sample_batch=sample, sample_sec=sec, cost_law_exponent=law)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    cb = cpu_reference_measure(args.batch, args.dim, steps, warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["sec_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ClipLoss fwd+bwd, global B={args.batch}, D={args.dim}, world_size=1 on host cores",
                   "sample_batch": cb["sample_batch"], "same_config": cb["sample_batch"] == args.batch},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ eager side figure
def gpu_eager_ms(dev, shapes, iters=5):
    """The reference's operator sequence (clip/loss.py:117-118,208-209 + mean().backward()) in eager PyTorch on the same
    B200: SURVEY 2a names it as the bar for the shapes FLYP actually runs.  ms per fwd+bwd."""
    import torch
    import torch.nn.functional as F

    def eager(I, T, s):
        li = s * I @ T.T
        lt = s * T @ I.T
        lab = torch.arange(li.shape[0], device=I.device, dtype=torch.long)
        return (F.cross_entropy(li, lab, reduction='none') + F.cross_entropy(lt, lab, reduction='none')) / 2

    out = {}
    for n, d, dt in shapes:
        try:
            I, T = synthetic_pairs(n, d, dtype=dt)
            I = I.to(dev).requires_grad_(True); T = T.to(dev).requires_grad_(True)
            th = torch.tensor(THETA0, device=dev, requires_grad=True)
            for _ in range(3):
                I.grad = T.grad = th.grad = None
                eager(I, T, th.exp()).mean().backward()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                I.grad = T.grad = th.grad = None
                eager(I, T, th.exp()).mean().backward()
            e1.record(); torch.cuda.synchronize()
            out[f"B{n}_D{d}_{'bf16' if dt == torch.bfloat16 else 'fp32'}"] = e0.elapsed_time(e1) / iters
            del I, T
            torch.cuda.empty_cache()
        except Exception as exc:      # noqa: BLE001 - out of memory at the largest shape is an answer, not a failure
            out[f"B{n}_D{d}_{'bf16' if dt == torch.bfloat16 else 'fp32'}"] = f"failed: {type(exc).__name__}"
    return out


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import flyp_b200
    from flyp_b200 import _lib
    from tools import sampled_check as sck

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = flyp_b200.load()
    fdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    esz = 2 if fdt == torch.bfloat16 else 4

    B, D = args.batch, args.dim
    assert B % world == 0
    b = B // world
    I_all, T_all = synthetic_pairs(B, D, seed=0, dtype=fdt)
    I_host = I_all[rank * b:(rank + 1) * b].contiguous().pin_memory()
    T_host = T_all[rank * b:(rank + 1) * b].contiguous().pin_memory()
    I_dev = I_host.to(dev).requires_grad_(True)
    T_dev = T_host.to(dev).requires_grad_(True)
    theta = torch.tensor(THETA0, device=dev, requires_grad=True)
    loss_fn = flyp_b200.ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=rank,
                                 world_size=world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)             # > 126 MB L2
    loss_host = torch.empty(B, dtype=fdt).pin_memory()

    def step_resident():
        I_dev.grad = T_dev.grad = theta.grad = None
        loss = loss_fn(I_dev, T_dev, theta.exp())
        loss.mean().backward()                               # torch.mean + backward, src/models/flyp_loss.py:498-499
        return loss

    # e2e: the step's inputs start in pinned HOST memory.  Copies run on a side stream into double-buffered device
    # tensors so that the H2D of step k+1 overlaps the kernels of step k (what a training loop's prefetcher does);
    # every step still pays its own 2 * b * D * e bytes of H2D and the D2H of its loss vector inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_bufs = [(torch.empty_like(I_host, device=dev), torch.empty_like(T_host, device=dev)) for _ in range(2)]
    copy_done = [torch.cuda.Event() for _ in range(2)]
    compute_done = [torch.cuda.Event() for _ in range(2)]

    def issue_copy(k):
        slot = k & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(compute_done[slot])          # the slot's previous consumer has finished
            dev_bufs[slot][0].copy_(I_host, non_blocking=True)
            dev_bufs[slot][1].copy_(T_host, non_blocking=True)
            copy_done[slot].record(copy_stream)

    def run_e2e(steps):
        main = torch.cuda.current_stream(dev)
        for ev in compute_done:
            ev.record(main)
        issue_copy(0)
        for k in range(steps):
            slot = k & 1
            if k + 1 < steps:
                issue_copy(k + 1)
            main.wait_event(copy_done[slot])
            Ii = dev_bufs[slot][0].detach().requires_grad_(True)
            Ti = dev_bufs[slot][1].detach().requires_grad_(True)
            theta.grad = None
            loss = loss_fn(Ii, Ti, theta.exp())
            loss.mean().backward()
            loss_host.copy_(loss.detach(), non_blocking=True)
            compute_done[slot].record(main)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            # L2 flush, enqueued between the steps and outside their event pairs; no host synchronisation inside the
            # K steps (the region is bracketed by barrier + synchronize, ranks meet in the step's own exchange)
            flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b_) for a, b_ in evs) / steps
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms_res = timed(step_resident, args.steps, args.warmup)
    run_e2e(max(3, args.warmup // 2))
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = t.item()
    clocks = sampler.stop() if sampler else None

    # ---- kernel-only timings inside real module steps: the library records events around the tcgen05 launches ----------
    def kernel_ms(sweep, iters):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for e in ev:
            e.record()                                       # creates the underlying cudaEvent_t
        torch.cuda.synchronize()
        _lib.check(lib.flyp_debug_kernel_events(ev[0].cuda_event, ev[1].cuda_event, ev[2].cuda_event, ev[3].cuda_event, sweep))
        fwd, swp = [], []
        try:
            for it in range(iters + 2):
                flush.fill_(1)
                step_resident()
                barrier()
                if it >= 2:
                    fwd.append(ev[0].elapsed_time(ev[1])); swp.append(ev[2].elapsed_time(ev[3]))
        finally:
            lib.flyp_debug_kernel_events(None, None, None, None, 0)
        return sum(fwd) / len(fwd), sum(swp) / len(swp)

    # backward plan of this shape (include/flyp_clip.h: flyp_clip_backward_plan): 0 two sweeps, 1 sweep + product,
    # 2 (one GPU) dS kernel + two products
    code = _lib.FLYP_BF16 if fdt == torch.bfloat16 else _lib.FLYP_F32
    plan = int(lib.flyp_clip_backward_plan(b, B, D, code))
    if world > 1:
        plan = min(plan, 1)
        # (the library's A/B switch; and the communicator's threshold: the product + NVLink reduce-scatter of the text
        # gradient from 6144 rows per rank on, the transposed sweep below)
        if os.environ.get("FLYP_KEEP_DS_RS", "1") == "0" or b < int(os.environ.get("FLYP_RS_MIN_ROWS", "6144")):
            plan = 0
    n_k = max(3, min(args.steps, 10))
    fwd_ms, sweep0_ms = kernel_ms(0, n_k)                # event pair 0: first sweep / dS kernel
    _, sweep1_ms = kernel_ms(1, n_k)                     # 1: second sweep / product dS^T . image
    gemm_di_ms = kernel_ms(2, n_k)[1] if plan == 2 else 0.0              # 2: product dS . text (unfused backward only)
    kt = torch.tensor([fwd_ms, sweep0_ms, sweep1_ms, gemm_di_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    fwd_ms, sweep0_ms, sweep1_ms, gemm_di_ms = kt.tolist()

    # ---- correctness of the measured configuration, outside the timed region (every rank holds the synthetic batch) ----
    loss_vec = step_resident().detach()
    torch.cuda.synchronize()
    check = {}
    with torch.no_grad():
        Ia, Ta = I_all.to(dev), T_all.to(dev)
        s_val = float(theta.detach().exp().item())
        g = torch.full((B,), 1.0 / B, device=dev)
        lse64 = sck.full_lse(Ia, Ta, s_val)
        gen = torch.Generator().manual_seed(100 + rank)
        loc = torch.randperm(b, generator=gen)[:min(32, b)].to(dev)
        idx = loc + rank * b
        want_loss, want_dI, want_dT = sck.sampled_reference(Ia, Ta, s_val, g, idx, lse=lse64)
        errs = [sck.row_errors(loss_vec[idx], want_loss)[0], *sck.row_errors(I_dev.grad[loc], want_dI),
                *sck.row_errors(T_dev.grad[loc], want_dT)]
        ids = torch.stack([(I_dev.grad.float() * I_dev.detach().float()).sum(),
                           (T_dev.grad.float() * T_dev.detach().float()).sum()]).double()
        et = torch.tensor(errs, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ids)
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dtheta = theta.grad.double().item()
        errs = et.tolist()
        tol = (2.0 ** -8 if fdt == torch.bfloat16 else 0.0) + (2e-3 if fdt == torch.bfloat16 else 1e-4)
        check = {"rows_per_rank": int(loc.numel()), "reference": "float64 on the GPU, tools/sampled_check.py (independent of the kernels)",
                 "loss_rel_err": errs[0], "d_image_rel_err": errs[1], "d_image_worst_row_rel_err": errs[2],
                 "d_text_rel_err": errs[3], "d_text_worst_row_rel_err": errs[4],
                 "sum_dI_I_over_dtheta": ids[0].item() / dtheta, "sum_dT_T_over_dtheta": ids[1].item() / dtheta,
                 "tolerance": tol, "passed": bool(max(errs[0], errs[1], errs[3]) < tol),
                 "note": "through the drop-in module: loss and gradients of bf16 leaves are stored in bf16 (one rounding, 2^-8) on top of the 2e-3 kernel bar"}
        del lse64

    # ---- soak (n_gpus > 1): inputs change every step; what the consumers saw must be the new rows, never stale ones ----
    if world > 1 and args.soak > 0:
        check["soak"] = soak(args.soak, loss_fn, I_all, T_all, theta, rank, world, b, dev, loss_vec)
    del Ia, Ta

    pk = peaks()
    kept = plan >= 1
    f_exec = 4.0 * b * B * D                     # executed by a sweep launch (S recompute + one output GEMM)
    if plan == 2:
        # unfused backward: dS kernel (S recompute + dS, 2 b B D) and two products (2 b B D each), all algorithmic; the
        # roofline entry is the SLOWEST of the three tensor-core kernels
        cands = [("ds_kernel_mc (dS kernel of the unfused backward: S recompute on the forward's pipeline, dS written to "
                  "HBM as fp16)", sweep0_ms),
                 ("dst_gemm_kernel<transposed> (d text = dS^T . image over the kept dS)", sweep1_ms),
                 ("dst_gemm_kernel (d image = dS . text over the kept dS)", gemm_di_ms)]
        kname, sweep_ms = max(cands, key=lambda c: c[1])
        kname += ", per launch; the slowest of the three backward kernels"
        f_sweep = f_exec = 2.0 * b * B * D
    elif kept:
        # kept-dS backward: ONE sweep (S recompute + dS . T, and the dS tiles written out) and the product dS^T . I:
        # every executed FLOP is algorithmic (8 B^2 D per step: forward S, one recompute, dI, dT).  Several GPUs: the
        # product covers the rank's rows of dS against ALL text rows and its epilogue scatters the fp32 partials into the
        # owners' buffers over NVLink (the scatter half of a reduce-scatter), a small kernel sums the W slots
        f_sweep = f_exec
        sweep_ms = sweep0_ms
        kname = ("bwd_pair_kernel (the backward sweep: S recompute + dS . T product, dS tiles kept in HBM for the "
                 "d-text product dst_gemm_kernel), per launch")
    else:
        f_sweep = 3.0 * b * B * D                # ALGORITHMIC FLOPs of one of two sweep launches on this rank
        sweep_ms = 0.5 * (sweep0_ms + sweep1_ms)
        kname = "bwd_pair_kernel (backward sweep: S recompute + dS.B product), per launch, mean of the d-image and d-text launches"
    ach = f_sweep / (sweep_ms * 1e-3) / 1e12
    f_step = 8.0 * B * B * D
    step_tflops = f_step / (ms_res * 1e-3) / 1e12 / world
    traffic = None
    try:                                         # dram bytes per launch of the same kernel from an ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            prof = json.load(f)
        key = ("bwd_pair_kernel_keep" if kept else "bwd_pair_kernel") if world == 1 else f"bwd_pair_kernel_w{world}"
        if plan == 2:
            key = "ds_kernel_mc" if kname.startswith("ds_kernel") else "dst_gemm_kernel"
        if (B, D) == (32768, 512) and fdt == torch.bfloat16 and key in prof:
            traffic = prof[key]["traffic_bytes_per_launch"]
    except Exception:
        traffic = None

    line = {
        "metric": METRIC, "value": B / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": f"ClipLoss fwd+bwd, global B={B}, D={D}, {args.dtype} unit-norm pairs, logit_scale=1/0.07, "
                               f"row-sharded over {world} GPU(s)", "global_batch": B, "dim": D,
                   "parallelism": f"row-shard x{world}", "l2": "flushed (256 MiB write) between timed steps",
                   "timing": "CUDA events around each step (flush outside), no host sync inside the K steps, mean over steps, max over ranks"},
        "clocks": clocks,
        "check": check,
        "step_frac": step_tflops / pk["burst"],
        "e2e": {"value": B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * b * D * esz, "d2h_bytes_per_step": B * esz,
                "how": "pinned host inputs, H2D of step k+1 overlapped with step k on a copy stream, loss vector D2H; "
                       "one timed region over all steps"},
        # own kernels per step: 1 GPU 5 forward (preparation, tcgen05 sweep, finalize, 2 gated robust-path stubs) + 4
        # backward (2 vector kernels, 2 tcgen05 sweeps); peer path 7 forward (pack, sweep, finalize, 2 gated stubs,
        # statistics push, finish) + 5 backward (2 vector kernels, 2 sweeps, d(scale) sum)
        # kept-dS backward on several GPUs: 2 vector kernels, sweep, product, flag release, slot sum, d(scale) sum = 7
        # unfused backward on one GPU: 2 vector kernels, dS kernel, d(scale) sum, 2 products = 6 (11 with the forward's 5)
        "gpu_launches": ((11 if plan == 2 else 9) if world == 1 else (14 if kept else 12)) * args.steps,
        "roofline": {"bound": "tensor", "kernel": kname,
                     "achieved": ach, "peak": pk["burst"], "unit": "TFLOP/s", "frac": ach / pk["burst"],
                     "peak_kind": "burst bf16 (kernel timed alone), " + pk["source"],
                     "frac_of_sustained": ach / pk["sustained"],
                     "ms_per_launch": sweep_ms, "ms_d_image_launch": sweep0_ms, "ms_d_text_launch": sweep1_ms,
                     "backward_plan": plan,
                     "d_text_kernel": ("dst_gemm_kernel (dS^T . I over the kept dS, 2 b B D FLOPs)" if kept
                                       else "bwd_pair_kernel (second sweep)"),
                     "d_text_tflops": (2.0 if kept else 3.0) * b * B * D / (sweep1_ms * 1e-3) / 1e12,
                     **({"ds_kernel_ms": sweep0_ms, "ds_kernel_tflops": 2.0 * b * B * D / (sweep0_ms * 1e-3) / 1e12,
                         "d_image_product_ms": gemm_di_ms,
                         "d_image_product_tflops": 2.0 * b * B * D / (gemm_di_ms * 1e-3) / 1e12} if plan == 2 else {}),
                     "algorithmic_flops_per_launch": f_sweep, "executed_flops_per_launch": f_exec,
                     "executed_tflops": f_exec / (sweep_ms * 1e-3) / 1e12,
                     "how": "cudaEventRecord by the library right before / after the kernel launch, inside real module steps",
                     "traffic": traffic,
                     "traffic_unit": "bytes per launch (dram read + write, profiles/ncu_summary.json)"},
        "step_breakdown": {"fwd_kernel_ms": fwd_ms, "fwd_kernel_tflops": 2.0 * b * B * D / (fwd_ms * 1e-3) / 1e12,
                           "bwd_sweep_ms": [sweep0_ms, sweep1_ms] + ([gemm_di_ms] if plan == 2 else []),
                           "bwd_kernels": (["ds_kernel_mc", "dst_gemm_kernel (d text)", "dst_gemm_kernel (d image)"]
                                           if plan == 2 else ["bwd_pair_kernel", "dst_gemm_kernel (d text)"] if plan == 1
                                           else ["bwd_pair_kernel (d image)", "bwd_pair_kernel (d text)"]),
                           "kernels_ms": fwd_ms + sweep0_ms + sweep1_ms + gemm_di_ms,
                           "other_ms": ms_res - (fwd_ms + sweep0_ms + sweep1_ms + gemm_di_ms),
                           "step_tflops_8B2D_per_gpu": step_tflops, "step_frac_of_burst": step_tflops / pk["burst"],
                           "step_frac_of_sustained": step_tflops / pk["sustained"],
                           "executed_flops_per_step": (8.0 if kept else 10.0) * b * B * D,
                           "step_tflops_6B2D_per_gpu": step_tflops * 0.75},
    }
    if rank == 0:
        if world == 1 and not args.no_side:
            shapes = [(512, 512, torch.bfloat16), (512, 512, torch.float32), (4096, 768, torch.bfloat16),
                      (4096, 768, torch.float32), (32768, 512, torch.bfloat16)]
            line["gpu_eager"] = {"what": "reference operator sequence in eager PyTorch on this GPU, ms per fwd+bwd",
                                 **gpu_eager_ms(dev, shapes)}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_measure(B, D, 2, 1, budget_s=25.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def soak(n_steps, loss_fn, I_all, T_all, theta, rank, world, b, dev, loss_base):
    """Every step every rank feeds the batch ROLLED by one more row, so each push carries new bits into the same
    slots.  After each step, on the stream and without host synchronisation: (1) the gathered matrices in this rank's
    exchange segment must equal the rolled batch bit for bit; (2) the loss vector must be the rolled base loss
    (permutation equivariance) - a kernel that consumed a stale or half-written row block would be off by O(1) in that
    block.  One flag is read at the end."""
    import torch
    import torch.distributed as dist
    B, D = I_all.shape

    class _Raw:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i2", "data": (int(ptr), False), "version": 2}

    Ia, Ta = I_all.to(dev), T_all.to(dev)
    bad_bits = torch.zeros((), dtype=torch.int64, device=dev)
    worst = torch.zeros((), dtype=torch.float32, device=dev)
    base = loss_base.float()
    scale_ref = base.abs().max()
    for k in range(1, n_steps + 1):
        Ik = torch.roll(Ia, shifts=k, dims=0); Tk = torch.roll(Ta, shifts=k, dims=0)
        Il = Ik[rank * b:(rank + 1) * b].contiguous().requires_grad_(True)
        Tl = Tk[rank * b:(rank + 1) * b].contiguous().requires_grad_(True)
        theta.grad = None
        loss = loss_fn(Il, Tl, theta.exp())
        loss.mean().backward()
        st = getattr(loss.grad_fn, "st", None)
        if st is not None and st.comm is not None:
            gg = st.step.gathered
            ti = torch.as_tensor(_Raw(gg.txt_all, B * D), device=dev)
            ii = torch.as_tensor(_Raw(gg.img_all, B * D), device=dev)
            bad_bits += (ti != Tk.view(torch.int16).reshape(-1)).sum() + (ii != Ik.view(torch.int16).reshape(-1)).sum()
        worst = torch.maximum(worst, (loss.detach().float() - torch.roll(base, shifts=k)).abs().max() / scale_ref)
    out = torch.stack([bad_bits.double(), worst.double()])
    dist.all_reduce(out, op=dist.ReduceOp.MAX)
    return {"steps": n_steps, "mismatching_gathered_elements": int(out[0].item()),
            "worst_loss_rel_err_vs_rolled_base": out[1].item(),
            "passed": bool(out[0].item() == 0 and out[1].item() < 2.0 ** -7)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--soak", type=int, default=300, help="soak steps of the in-run check when n_gpus > 1 (0: off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the eager-PyTorch side figure")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
