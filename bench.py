#!/usr/bin/env python
"""Headline benchmark: ClipLoss forward + backward pairs/s at global B = 32768, D = 512, bf16 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dim D]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one forward (per-item loss vector) + one backward of loss.mean() through the drop-in module.
`value`     inputs resident in HBM (row block of this rank), CUDA-event timed, max over ranks, L2 flushed between steps.
`e2e`       the same call with HOST inputs: pinned H2D of both feature blocks and D2H of the loss vector inside the
            timed region.
`roofline`  the dominant kernel (the tcgen05 backward sweep), timed live with CUDA events inside the timed steps.
`cpu_baseline` / `--impl reference`: the reference's CPU operator sequence (oracle/torch_port.py, "port") on the
            host cores, on a bounded sample (smaller batch, extrapolated with the B^2 cost law) - see DESIGN.md.
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
SCALE = 1.0 / 0.07


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


def synthetic_pairs(n, d, seed=0):
    """BASELINE.md section 3 inputs: I = normalize(randn), T = normalize(0.5 I + 0.5 normalize(randn)), bf16."""
    import torch
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=gen)
    y = torch.randn(n, d, generator=gen)
    I = F.normalize(x, dim=-1)
    T = F.normalize(0.5 * I + 0.5 * F.normalize(y, dim=-1), dim=-1)
    return I.bfloat16(), T.bfloat16()


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_port_measure(batch, dim, steps, warmup, budget_s=150.0):
    """Time the CPU port on a bounded sample.  Returns dict(value=pairs/s at `batch`, sample=..., cores=..., sec=...)."""
    import torch
    from oracle import torch_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = min(batch, 8192)
    t1, _ = torch_port.time_cpu_step(min(sample, 2048), dim, 1, 1)       # probe to size the sample
    est = t1 * (sample / min(sample, 2048)) ** 2 * (steps + warmup)
    while est > budget_s and sample > 1024:
        sample //= 2
        est /= 4
    sec, threads = torch_port.time_cpu_step(sample, dim, steps, warmup)
    factor = (batch / sample) ** 2                                         # cost of the materialising algorithm ~ B^2
    sec_full = sec * factor
    desc = (f"torch-CPU port of clip/loss.py:117-118,208-209 + mean().backward(), fp32, B={sample} D={dim}, "
            f"{steps} steps after {warmup} warm-up; time x{factor:.0f} (B^2 law) for B={batch}"
            if sample != batch else f"torch-CPU port, fp32, B={batch} D={dim}, {steps} steps after {warmup} warm-up")
    return dict(value=batch / sec_full, unit=UNIT, cores=threads, kind="port", sample=desc, sec_per_step=sec_full,
                sample_batch=sample, sample_sec=sec)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_port_measure(args.batch, args.dim, max(1, args.steps), max(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["sec_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ClipLoss fwd+bwd, global B={args.batch}, D={args.dim}, world_size=1 on host cores"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import flyp_b200
    from flyp_b200 import ops

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
    flyp_b200.load()

    B, D = args.batch, args.dim
    assert B % world == 0
    b = B // world
    I_all, T_all = synthetic_pairs(B, D, seed=0)                   # the oracle is only used by the cpu_baseline leg
    I_host = I_all[rank * b:(rank + 1) * b].contiguous().pin_memory()
    T_host = T_all[rank * b:(rank + 1) * b].contiguous().pin_memory()
    I_dev = I_host.to(dev).requires_grad_(True)
    T_dev = T_host.to(dev).requires_grad_(True)
    theta = torch.tensor(2.6592600369327783, device=dev, requires_grad=True)     # ln(1/0.07)
    loss_fn = flyp_b200.ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=rank,
                                 world_size=world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)             # > 126 MB L2
    loss_host = torch.empty(B, dtype=torch.bfloat16).pin_memory()

    def step_resident():
        I_dev.grad = T_dev.grad = theta.grad = None
        loss = loss_fn(I_dev, T_dev, theta.exp())
        loss.mean().backward()                               # torch.mean + backward, src/models/flyp_loss.py:498-499
        return loss

    # e2e: the step's inputs start in pinned HOST memory.  Copies run on a side stream into double-buffered device
    # tensors so that the H2D of step k+1 overlaps the kernels of step k (what a training loop's prefetcher does);
    # every step still pays its own 2 * b * D * 2 bytes of H2D and the D2H of its loss vector inside the timed region.
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
            # K steps (the region is bracketed by barrier + synchronize, ranks meet in the step's own collectives)
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

    # correctness of the measured configuration, outside the timed region (every rank holds the synthetic batch):
    # (1) the loss of 64 sampled items against fp32 logsumexp over their full row and column of logits;
    # (2) the identity sum_i <dI_i, I_i> = sum_j <dT_j, T_j> = d loss / d theta over ALL rows of all ranks.
    loss_vec = step_resident().detach()
    torch.cuda.synchronize()
    with torch.no_grad():
        Ia, Ta = I_all.to(dev).float(), T_all.to(dev).float()
        idx = torch.arange(0, B, max(1, B // 64), device=dev)[:64]
        s_val = theta.detach().exp()
        rows, cols = s_val * Ia[idx] @ Ta.T, s_val * Ta[idx] @ Ia.T
        want = 0.5 * (torch.logsumexp(rows, 1) + torch.logsumexp(cols, 1)) - s_val * (Ia[idx] * Ta[idx]).sum(1)
        loss_err = ((loss_vec[idx].float() - want).abs().max() / want.abs().max()).item()
        ids = torch.stack([(I_dev.grad.float() * I_dev.detach().float()).sum(),
                           (T_dev.grad.float() * T_dev.detach().float()).sum()]).double()
        if world > 1:
            dist.all_reduce(ids)
        dtheta = theta.grad.double().item()
        check = {"loss_rel_err_64_sampled_items_vs_fp32_logsumexp": loss_err,
                 "sum_dI_I_over_dtheta": ids[0].item() / dtheta, "sum_dT_T_over_dtheta": ids[1].item() / dtheta,
                 "note": "through the drop-in module: the loss vector and the gradients are stored in bf16 (one rounding each, <= 2^-8 relative), on top of the 2e-3 kernel bar"}
        del Ia, Ta, rows, cols

    # dominant kernel: the tcgen05 backward sweep (dI: S recompute + dS.T product), timed live with CUDA events on the
    # launching stream inside steps of the same sequence (fwd, sweep, sweep), L2 flushed between steps.
    sc = theta.detach().exp().reshape(1)
    Iall_dev = I_all.to(dev) if world > 1 else I_dev.detach()
    Tall_dev = T_all.to(dev) if world > 1 else T_dev.detach()
    off = rank * b
    g = torch.full((B,), 1.0 / B, device=dev)
    ws = ops.clip_workspace(b, B, D, 0, dev)
    sweep_ms, fwd_ms = [], []
    for it in range(args.warmup + args.steps):
        flush.fill_(1)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        row_lse, row_nll, col_stat, _ = ops.clip_fwd_local(I_dev.detach(), Tall_dev, sc, off, workspace=ws)
        e[1].record()
        if world > 1:
            cs = [torch.empty_like(col_stat) for _ in range(world)]
            dist.all_gather(cs, col_stat)
            rn = [torch.empty_like(row_nll) for _ in range(world)]; dist.all_gather(rn, row_nll)
            rl = [torch.empty_like(row_lse) for _ in range(world)]; dist.all_gather(rl, row_lse)
            col_stat_all, row_nll_all, row_lse_all = torch.cat(cs), torch.cat(rn), torch.cat(rl)
        else:
            col_stat_all, row_nll_all, row_lse_all = col_stat, row_nll, row_lse
        col_lse, col_nll, _ = ops.clip_fwd_finish(col_stat_all, world, row_nll_all, B, 0)
        sl = slice(off, off + b)
        e[2].record()
        ops.clip_bwd_local(I_dev.detach(), Tall_dev, sc, off, row_lse_all[sl].contiguous(),
                           row_nll_all[sl].contiguous(), col_lse, col_nll, g[sl].contiguous(), g, need_txt=False,
                           need_scale=True, workspace=ws)
        e[3].record()
        torch.cuda.synchronize()
        if it >= args.warmup:
            fwd_ms.append(e[0].elapsed_time(e[1])); sweep_ms.append(e[2].elapsed_time(e[3]))
    sweep = sum(sweep_ms) / len(sweep_ms)
    fwd = sum(fwd_ms) / len(fwd_ms)
    pk = peaks()
    # algorithmic FLOPs of one backward-sweep launch on this rank (S recompute + one output GEMM)
    f_sweep = 4.0 * b * B * D
    ach = f_sweep / (sweep * 1e-3) / 1e12
    f_step = 8.0 * B * B * D
    traffic = None
    try:                                                     # dram bytes per launch of the same kernel, ncu --set full
        with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
            prof = json.load(f)["bwd_pair_kernel"]
        if world == 1 and (B, D) == (32768, 512):
            traffic = prof["traffic_bytes_per_launch"]
    except Exception:
        traffic = None
    step_tflops = f_step / (ms_res * 1e-3) / 1e12 / world

    line = {
        "metric": METRIC, "value": B / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"ClipLoss fwd+bwd, global B={B}, D={D}, bf16 unit-norm pairs, logit_scale=1/0.07, "
                               f"row-sharded over {world} GPU(s)", "global_batch": B, "dim": D,
                   "parallelism": f"row-shard x{world}", "l2": "flushed (256 MiB write) between timed steps",
                   "timing": "CUDA events around each step (flush outside), no host sync inside the K steps, mean over steps, max over ranks"},
        "clocks": clocks,
        "check": check,
        "e2e": {"value": B / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": 2 * b * D * 2, "d2h_bytes_per_step": B * 2,
                "how": "pinned host inputs, H2D of step k+1 overlapped with step k on a copy stream, loss vector D2H; "
                       "one timed region over all steps"},
        # own kernels per step (profiles/r01c_launches_step_B32768.csv, r01_timeline_*): 1 GPU 8 forward + 10 backward;
        # peer path 10 forward (pack, pair_dot, forward, finalize, 4 gated robust helpers, statistics push, finish) +
        # 9 backward (prep, fast vectors, 2 sweeps, 2 partial reductions, d(scale) sum / push / all-rank sum)
        "gpu_launches": (18 if world == 1 else 19) * args.steps,
        "roofline": {"bound": "tensor", "kernel": "bwd_kernel (dI sweep: S recompute + dS.T), per launch",
                     "achieved": ach, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": ach / pk["sustained"],
                     "peak_kind": "sustained bf16, " + pk["source"], "frac_of_burst": ach / pk["burst"],
                     "ms_per_launch": sweep, "algorithmic_flops_per_launch": f_sweep, "traffic": traffic,
                     "traffic_unit": "bytes per launch (dram read + write, profiles/ncu_summary.json)"},
        "step_breakdown": {"fwd_stats_ms": fwd, "bwd_sweep_ms": sweep,
                           "step_tflops_8B2D_per_gpu": step_tflops, "step_frac_of_burst": step_tflops / pk["burst"],
                           "step_frac_of_sustained": step_tflops / pk["sustained"],
                           "step_tflops_6B2D_per_gpu": step_tflops * 0.75},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_port_measure(B, D, 3, 1, budget_s=25.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
