"""The FLYP finetune step around the operator - host-side mirror of the reference's training loops for the hot path.

    src/models/flyp_loss.py:365-371   ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, ...), AdamW
    src/models/flyp_loss.py:426       optimizer.zero_grad()
    src/models/flyp_loss.py:495-500   features, scale = model(image, text); loss = mean(clip_loss_fn(...)); backward; step
    src/models/ce_ablation.py:104-126 one sampled prompt per class, normalise, logits, F.cross_entropy, backward, step

What is different from the reference loop, on purpose ("train-loop hygiene"):
  * no device-to-host copy per step: the reference ships the per-item loss vector (.cpu().tolist(), :503) and the scalar
    loss (.item(), :513) to the host every step, which serialises host and device; here the per-item losses accumulate on
    the device and are fetched only when the caller asks (`StepLog.fetch`, every `log_every` steps);
  * `world_size > 1` is real: one process per GPU (torchrun), towers under DistributedDataParallel, and the loss built with
    the rank / world size of the process group, so the row-sharded ClipLoss (NVLink peer memory) is what runs - the
    reference hard-codes world_size=1 under nn.DataParallel (:335,365) and its multi-rank loss path is unreachable;
  * the two NameErrors of the shipped loops (`ft_imgid` without --cluster=loss, flyp_loss.py:504; `templates` in
    ce_ablation.py:32) have no counterpart: image ids and prompt tokens are explicit arguments.

The towers are NOT part of the hot path (out of scope): `TwoTowerEncoder` is a plain torch.nn
stand-in with the ViT-B/16 CLIP shapes (12 x 768 image tower on 224^2 / patch 16, 12 x 512 text tower on 77 tokens,
embed 512; random init) so that BASELINE configuration 5 can be run and timed without the reference's model code.  Its
tail - final projection + L2 normalisation, clip/model.py:239-243,359,375-376 - is the fused `project_normalize`
(flyp_b200/tail.py) when `fused_tail=True`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------- stand-in towers
class _Block(nn.Module):
    def __init__(self, width: int, heads: int):
        super().__init__()
        self.heads = heads
        self.ln_1 = nn.LayerNorm(width)
        self.qkv = nn.Linear(width, 3 * width)
        self.out = nn.Linear(width, width)
        self.ln_2 = nn.LayerNorm(width)
        self.fc = nn.Linear(width, 4 * width)
        self.proj = nn.Linear(4 * width, width)

    def forward(self, x, causal: bool):
        b, n, w = x.shape
        q, k, v = self.qkv(self.ln_1(x)).view(b, n, 3, self.heads, w // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
        x = x + self.out(a.transpose(1, 2).reshape(b, n, w))
        h = self.fc(self.ln_2(x))
        return x + self.proj(h * torch.sigmoid(1.702 * h))          # QuickGELU


class _Tower(nn.Module):
    """Transformer trunk returning the pooled, layer-normed feature BEFORE the final projection."""

    def __init__(self, width: int, layers: int, heads: int, seq: int, causal: bool, pre_ln: bool):
        super().__init__()
        self.causal = causal
        self.pos = nn.Parameter(width ** -0.5 * torch.randn(seq, width))
        self.ln_pre = nn.LayerNorm(width) if pre_ln else nn.Identity()       # clip/model.py:234 (vision tower only)
        self.blocks = nn.ModuleList([_Block(width, heads) for _ in range(layers)])
        self.ln_out = nn.LayerNorm(width)

    def trunk(self, x):
        x = self.ln_pre(x + self.pos)
        for blk in self.blocks:
            x = blk(x, self.causal)
        return x


class TwoTowerEncoder(nn.Module):
    """CLIP-shaped two-tower encoder: forward(image, text) -> (image_features, text_features, logit_scale.exp()) like
    clip/model.py:363-378; forward(image, None) / forward(None, text) return the un-normalised features of one tower
    (clip/model.py:364-369, what the ce_ablation loop uses)."""

    def __init__(self, embed_dim=512, image_size=224, patch=16, vision_width=768, vision_layers=12, vision_heads=12,
                 context=77, vocab=49408, text_width=512, text_layers=12, text_heads=8, fused_tail=False):
        super().__init__()
        self.fused_tail = fused_tail
        self.patch = nn.Conv2d(3, vision_width, patch, patch, bias=False)
        self.cls = nn.Parameter(vision_width ** -0.5 * torch.randn(vision_width))
        self.visual = _Tower(vision_width, vision_layers, vision_heads, (image_size // patch) ** 2 + 1, causal=False,
                             pre_ln=True)
        self.visual_proj = nn.Parameter(vision_width ** -0.5 * torch.randn(vision_width, embed_dim))
        self.tok = nn.Embedding(vocab, text_width)
        self.text = _Tower(text_width, text_layers, text_heads, context, causal=True, pre_ln=False)
        self.text_proj = nn.Parameter(text_width ** -0.5 * torch.randn(text_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))         # clip/model.py:299

    def image_trunk(self, image):
        x = self.patch(image).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls.expand(x.shape[0], 1, -1).to(x.dtype), x], dim=1)
        return self.visual.ln_out(self.visual.trunk(x)[:, 0])                       # class token, clip/model.py:239

    def text_trunk(self, text):
        x = self.text.trunk(self.tok(text))
        x = self.text.ln_out(x)
        return x[torch.arange(x.shape[0], device=x.device), text.argmax(dim=-1)]      # eot token, clip/model.py:359

    def _tail(self, pooled, proj, normalize: bool):
        if self.fused_tail and normalize:
            from .tail import project_normalize
            return project_normalize(pooled, proj)
        y = pooled @ proj
        return y / y.norm(dim=-1, keepdim=True) if normalize else y

    def forward(self, image, text):
        if image is None:
            return self._tail(self.text_trunk(text), self.text_proj, False)
        if text is None:
            return self._tail(self.image_trunk(image), self.visual_proj, False)
        return (self._tail(self.image_trunk(image), self.visual_proj, True),
                self._tail(self.text_trunk(text), self.text_proj, True), self.logit_scale.exp())


# ---------------------------------------------------------------------------------------------------- the steps
@dataclass
class StepLog:
    """Per-item losses (with their image ids) and the running loss sum, kept ON THE DEVICE between fetches."""
    items: List[torch.Tensor] = field(default_factory=list)
    ids: List[torch.Tensor] = field(default_factory=list)
    loss_sum: Optional[torch.Tensor] = None
    steps: int = 0

    def add(self, peritem: torch.Tensor, loss: torch.Tensor, ids: Optional[torch.Tensor]):
        self.items.append(peritem.detach())
        if ids is not None:
            self.ids.append(ids)
        self.loss_sum = loss.detach().float() if self.loss_sum is None else self.loss_sum + loss.detach().float()
        self.steps += 1

    def fetch(self):
        """One synchronising device-to-host transfer for everything logged so far: (list of (id, loss), mean loss) -
        what flyp_loss.py:503-513 builds every step."""
        if not self.items:
            return [], float("nan")
        losses = torch.cat([t.float() for t in self.items]).cpu().tolist()
        ids = torch.cat(self.ids).cpu().tolist() if self.ids else list(range(len(losses)))
        mean = (self.loss_sum / self.steps).item()
        self.items.clear(); self.ids.clear(); self.loss_sum = None; self.steps = 0
        return list(zip(ids, losses)), mean


def finetune_step(model, clip_loss_fn, optimizer, image, text, image_ids=None, log: Optional[StepLog] = None,
                  feature_dtype: Optional[torch.dtype] = None, scheduler=None, step: int = 0):
    """One FLYP step (src/models/flyp_loss.py:426,495-500).  Returns (loss, per-item losses), both on the device."""
    if scheduler is not None:
        scheduler(step)
    optimizer.zero_grad(set_to_none=True)
    image_features, text_features, logit_scale = model(image, text)
    if feature_dtype is not None:
        image_features, text_features = image_features.to(feature_dtype), text_features.to(feature_dtype)
    peritem = clip_loss_fn(image_features, text_features, logit_scale)
    loss = torch.mean(peritem)
    loss.backward()
    optimizer.step()
    if log is not None:
        log.add(peritem, loss, image_ids)
    return loss.detach(), peritem.detach()


def ce_ablation_step(model, optimizer, images, class_prompts, labels, fused: bool = True, generator=None):
    """One --ce_ablation step (src/models/ce_ablation.py:104-126).  class_prompts: [C, n_templates, context] tokens; one
    template per class is sampled like :104-109.  `model` maps (images, None) / (None, tokens) to un-normalised features
    and exposes `logit_scale` (the reference reads model.module.model.logit_scale, :119)."""
    from .loss import contrastive_cross_entropy, l2_normalize
    optimizer.zero_grad(set_to_none=True)
    c = class_prompts.shape[0]
    pick = torch.randint(0, class_prompts.shape[1], (c,), generator=generator, device=class_prompts.device)
    current = class_prompts[torch.arange(c, device=class_prompts.device), pick]
    image_features = model(images, None)
    text_features = model(None, current)
    core = model.module if hasattr(model, "module") else model
    scale = core.logit_scale.exp()
    if fused:
        loss = contrastive_cross_entropy(l2_normalize(image_features), l2_normalize(text_features), scale, labels,
                                         reduction="mean")
    else:                                     # the reference's operator sequence (:115-123)
        image_features = image_features / image_features.norm(dim=-1, keepdim=True)
        text_features = text_features / text_features.norm(dim=-1, keepdim=True)
        loss = F.cross_entropy(scale * image_features @ text_features.T, labels)
    loss.backward()
    optimizer.step()
    return loss.detach()


def build_for_rank(device, fused_tail=False, ddp=None, lr=1e-5, wd=0.1, **encoder_kw):
    """(model, ClipLoss, AdamW) for this process: rank / world size come from torch.distributed when it is initialised
    (`ddp=None`: wrap in DistributedDataParallel exactly then)."""
    import torch.distributed as dist
    from .loss import ClipLoss
    model = TwoTowerEncoder(fused_tail=fused_tail, **encoder_kw).to(device)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank, world = (dist.get_rank(), dist.get_world_size()) if multi else (0, 1)
    if (ddp if ddp is not None else multi):
        model = nn.parallel.DistributedDataParallel(model, device_ids=[device.index])
    loss_fn = ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=rank, world_size=world,
                       use_horovod=False)                                                # flyp_loss.py:365-366
    params = [p for p in model.parameters() if p.requires_grad]
    optimizer = torch.optim.AdamW(params, lr=lr, weight_decay=wd)                          # flyp_loss.py:368-371
    return model, loss_fn, optimizer
