"""ORACLE / CPU BASELINE - test and benchmark infrastructure only (never imported by flyp_b200/).

A PyTorch-CPU port of the reference's default ClipLoss branch, i.e. the same ATen operator sequence the reference
executes on host cores (it cannot be imported on the GPU box, where /root/reference does not exist):

    clip/loss.py:117-118   logits_per_image = logit_scale * image_features @ text_features.T
                           logits_per_text  = logit_scale * text_features @ image_features.T      (second GEMM)
    clip/loss.py:195-198   labels = arange(num_logits)
    clip/loss.py:208-209   (F.cross_entropy(li, labels, 'none') + F.cross_entropy(lt, labels, 'none')) / 2

followed by the caller's reduction and backward (src/models/flyp_loss.py:498-499).  Used by bench.py for the
`cpu_baseline` object and the `--impl reference` arm ("kind": "port"), and by tests as a second opinion on the numpy
oracle.
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F


def clip_loss_reference_ops(image_features, text_features, logit_scale):
    logits_per_image = logit_scale * image_features @ text_features.T
    logits_per_text = logit_scale * text_features @ image_features.T
    labels = torch.arange(logits_per_image.shape[0], device=image_features.device, dtype=torch.long)
    return (F.cross_entropy(logits_per_image, labels, reduction='none') +
            F.cross_entropy(logits_per_text, labels, reduction='none')) / 2


def synthetic_pairs(n, d, seed=0, dtype=torch.float32):
    """BASELINE.md section 3 inputs: I = normalize(randn), T = normalize(0.5 I + 0.5 normalize(randn))."""
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=gen)
    y = torch.randn(n, d, generator=gen)
    I = F.normalize(x, dim=-1)
    T = F.normalize(0.5 * I + 0.5 * F.normalize(y, dim=-1), dim=-1)
    return I.to(dtype), T.to(dtype)


def time_cpu_step(n, d, steps, warmup, threads=None):
    """Seconds per fwd+bwd (loss.mean().backward()) of the port at batch n on the host cores.  Returns (sec, threads)."""
    if threads:
        torch.set_num_threads(threads)
    I, T = synthetic_pairs(n, d)
    I.requires_grad_(True); T.requires_grad_(True)
    theta = torch.tensor(2.6592600369327783, requires_grad=True)   # ln(1/0.07), clip/model.py:299
    times = []
    for it in range(warmup + steps):
        I.grad = T.grad = theta.grad = None
        t0 = time.perf_counter()
        loss = clip_loss_reference_ops(I, T, theta.exp())
        loss.mean().backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times), torch.get_num_threads()
