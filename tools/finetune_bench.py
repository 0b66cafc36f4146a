"""BASELINE.json configuration 5: the end-to-end FLYP finetune step with ViT-B/16-shaped towers (12 x 768 image tower on
224^2 / patch 16, 12 x 512 text tower on 77 tokens, embed 512; random init, synthetic batch), global B = 512, AdamW - the
loop of src/models/flyp_loss.py:365-371,426,495-500 as restated in flyp_b200/finetune.py - timed with

    arm "reference-ops": the reference's loss operator sequence (clip/loss.py:117-118,208-209) and unfused tail
    arm "flyp_b200":     the drop-in ClipLoss and the fused tail (project_normalize)

and the --ce_ablation head step (src/models/ce_ablation.py:104-126).  One JSON line per arm.

    python tools/finetune_bench.py [--batch 512] [--steps 8] [--autocast]
    python -m torch.distributed.run --nproc-per-node N ... tools/finetune_bench.py     (world_size N actually used)
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--autocast", action="store_true", help="bf16 autocast for the towers (the reference trains in fp32)")
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--classes", type=int, default=1000)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
lr_ = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import flyp_b200
from flyp_b200.finetune import StepLog, build_for_rank, ce_ablation_step, finetune_step

b = args.batch // world
g = torch.Generator().manual_seed(rank)
image = torch.randn(b, 3, 224, 224, generator=g).to(dev)
text = torch.randint(1, 49407, (b, 77), generator=g).to(dev)
text[:, -1] = 49407                                       # eot = highest token id (clip/model.py:359 takes argmax)
ids = torch.arange(b, device=dev) + rank * b


def reference_loss(fi, ft, s):                            # clip/loss.py:117-118,195-198,208-209 (world_size = 1 only)
    li = s * fi @ ft.T
    lt = s * ft @ fi.T
    lab = torch.arange(li.shape[0], device=fi.device, dtype=torch.long)
    return (F.cross_entropy(li, lab, reduction='none') + F.cross_entropy(lt, lab, reduction='none')) / 2


def run(arm):
    torch.manual_seed(0)
    fused = arm == "flyp_b200"
    model, loss_fn, opt = build_for_rank(dev, fused_tail=fused, vision_layers=args.layers, text_layers=args.layers)
    if not fused:
        if world > 1:
            return None                                   # the reference loop cannot use more than one rank (flyp_loss.py:365)
        loss_fn = reference_loss
    log = StepLog()

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.autocast):
            return finetune_step(model, loss_fn, opt, image, text, image_ids=ids, log=log)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        loss, _ = step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps * 1e3
    pairs, mean = log.fetch()
    # the loss call alone (forward + backward of the operator on the step's features), same features for both arms
    with torch.no_grad():
        fi, ft, s = model(image[:b], text[:b])
    fi = fi.detach().requires_grad_(True); ft = ft.detach().requires_grad_(True); s = s.detach().requires_grad_(True)
    for _ in range(3):
        loss_fn(fi, ft, s).mean().backward()
    torch.cuda.synchronize()
    l0 = torch.cuda.Event(enable_timing=True); l1 = torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(20):
        fi.grad = ft.grad = s.grad = None
        loss_fn(fi, ft, s).mean().backward()
    l1.record(); torch.cuda.synchronize()
    return {"arm": arm, "world": world, "global_batch": args.batch, "towers": f"ViT-B/16 shapes, {args.layers} layers, "
            f"{'bf16 autocast' if args.autocast else 'fp32 (as the reference)'}", "feature_dtype": str(fi.dtype).split(".")[-1],
            "step_ms": e0.elapsed_time(e1) / args.steps, "step_wall_ms": wall, "loss_fwd_bwd_ms": l0.elapsed_time(l1) / 20,
            "mean_loss": mean, "items_logged": len(pairs), "d2h_transfers": 1}


for arm in ("reference-ops", "flyp_b200"):
    out = run(arm)
    if out is not None and rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.empty_cache()

# --ce_ablation head step (src/models/ce_ablation.py:104-126), single rank like the reference
if world == 1:
    for fused in (False, True):
        torch.manual_seed(0)
        model, _, opt = build_for_rank(dev, vision_layers=args.layers, text_layers=args.layers)
        prompts = torch.randint(1, 49407, (args.classes, 4, 77), generator=g).to(dev)
        prompts[:, :, -1] = 49407
        labels = torch.randint(0, args.classes, (b,), generator=g).to(dev)
        nb = min(b, 128)
        for _ in range(2):
            ce_ablation_step(model, opt, image[:nb], prompts, labels[:nb], fused=fused)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(max(2, args.steps // 2)):
            loss = ce_ablation_step(model, opt, image[:nb], prompts, labels[:nb], fused=fused)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"arm": "ce_ablation " + ("flyp_b200" if fused else "reference-ops"), "batch": nb,
                          "classes": args.classes, "step_ms": e0.elapsed_time(e1) / max(2, args.steps // 2),
                          "loss": loss.item()}), flush=True)
        del model, opt
        torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
