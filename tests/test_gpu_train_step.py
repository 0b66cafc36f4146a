"""BASELINE.json configuration 5: an end-to-end FLYP finetune step with the new ClipLoss swapped in, plus the
--ce_ablation head path.

The reference's entry points do not import here (open_clip / wilds / webdataset are absent), and
/root/reference does not exist on the GPU box, so the loop is RESTATED following
    src/models/flyp_loss.py:365-371   ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, ...), AdamW
    src/models/flyp_loss.py:495-500   features, logit_scale = model(image, text); loss = mean(clip_loss_fn(...)); backward; step
    clip/model.py:365-378             encode, x / x.norm(dim=-1, keepdim=True), logit_scale.exp()
    src/models/ce_ablation.py:113-126 normalise, logits = logit_scale * img @ txt.T, F.cross_entropy, backward, step
over a small random-init two-tower transformer (the towers are out of scope - any differentiable producer of [B, D]
features exercises the drop-in boundary: autograd through the module into every upstream parameter).
Arm A runs the reference's operator sequence (oracle/torch_port.py) on the GPU in fp32, arm B the drop-in module; both
start from the same weights and batch.  Checked: per-item loss, gradients of EVERY model parameter (incl. logit_scale)
and the parameters after one AdamW step (where the gradient is not negligible).  fp32 features (what FLYP trains in):
1e-5 / 1e-4; bf16 features: 2e-3 plus the bf16 roundings autograd itself applies to the loss vector's gradient and to the
gradient that enters the towers.
"""
import copy
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class Tower(nn.Module):
    def __init__(self, in_dim, width, embed, seq, heads=4, layers=2):
        super().__init__()
        self.inp = nn.Linear(in_dim, width)
        self.pos = nn.Parameter(0.02 * torch.randn(seq, width))
        layer = nn.TransformerEncoderLayer(width, heads, 2 * width, dropout=0.0, batch_first=True, norm_first=True)
        self.blocks = nn.TransformerEncoder(layer, layers)
        self.ln = nn.LayerNorm(width)
        self.proj = nn.Parameter(width ** -0.5 * torch.randn(width, embed))

    def forward(self, x):                       # x: [B, seq, in_dim]
        h = self.blocks(self.inp(x) + self.pos)
        return self.ln(h[:, 0]) @ self.proj     # class-token style pooling, then projection (clip/model.py:239-243)


class MiniCLIP(nn.Module):
    def __init__(self, embed=64):
        super().__init__()
        self.visual = Tower(48, 96, embed, 17)
        self.text = Tower(32, 64, embed, 9)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))     # clip/model.py:299

    def forward(self, image, text):             # clip/model.py:365-378
        i, t = self.visual(image), self.text(text)
        i = i / i.norm(dim=-1, keepdim=True)
        t = t / t.norm(dim=-1, keepdim=True)
        return i, t, self.logit_scale.exp()


def _batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 17, 48, generator=g).to(DEV), torch.randn(n, 9, 32, generator=g).to(DEV)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


def _one_step(model, loss_fn, image, text, feat_dtype):
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.1)
    fi, ft, s = model(image, text)
    peritem = loss_fn(fi.to(feat_dtype), ft.to(feat_dtype), s)
    loss = torch.mean(peritem)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    return peritem.detach().float(), grads


@pytest.mark.parametrize("n", [96, 512])
@pytest.mark.parametrize("feat_dtype", [torch.float32, torch.bfloat16])
def test_finetune_step_matches_reference_ops(n, feat_dtype):
    from flyp_b200 import ClipLoss
    from oracle import torch_port
    torch.manual_seed(0)
    model_a = MiniCLIP().to(DEV)
    model_b = copy.deepcopy(model_a)
    image, text = _batch(n, 1)

    def reference_loss(fi, ft, s):              # the reference's ATen sequence, fp32 math on the (rounded) features
        return torch_port.clip_loss_reference_ops(fi.float(), ft.float(), s)

    ours = ClipLoss(local_loss=False, gather_with_grad=False, cache_labels=True, rank=0, world_size=1, use_horovod=False)
    la, ga = _one_step(model_a, reference_loss, image, text, feat_dtype)
    lb, gb = _one_step(model_b, ours, image, text, feat_dtype)
    torch.cuda.synchronize()
    if feat_dtype == torch.float32:
        tol_loss, tol_grad = 1e-5, 1e-4
    else:
        # bf16 features make the loss vector bf16 (as in the reference): autograd then rounds the upstream gradient
        # 1/n of mean() to bf16 (a common factor, up to 2^-9; exact for n = 512) and the gradient entering the towers
        tol_loss, tol_grad = 2e-3 + 2.0 ** -8, 2e-3 + 2.0 ** -7
    assert lb.shape == la.shape == (n,)
    assert _rel(lb, la) < tol_loss
    for k in ga:
        assert _rel(gb[k], ga[k]) < tol_grad, k
    # one AdamW step later: the first Adam update is lr * g / (|g| + eps), so the models agree wherever |g| >> eps
    pa, pb = dict(model_a.named_parameters()), dict(model_b.named_parameters())
    for k in pa:
        big = ga[k].abs() > 20 * tol_grad * ga[k].abs().max()       # well clear of the comparison's own noise
        if big.any():
            assert (pa[k] - pb[k])[big].abs().max().item() < 0.1 * 1e-3, k
    assert ours.prev_num_logits == n and len(ours.labels) == 1      # cache_labels bookkeeping (clip/loss.py:195-206)


@pytest.mark.parametrize("n,classes", [(128, 182), (64, 1000)])
def test_ce_ablation_step_matches_reference_ops(n, classes):
    """src/models/ce_ablation.py:113-126 with iWildCam (182) / ImageNet (1000) class counts."""
    from flyp_b200 import contrastive_cross_entropy, l2_normalize
    torch.manual_seed(1)
    model_a = MiniCLIP().to(DEV)
    model_b = copy.deepcopy(model_a)
    g = torch.Generator().manual_seed(5)
    inputs = torch.randn(n, 17, 48, generator=g).to(DEV)
    prompts = torch.randn(classes, 9, 32, generator=g).to(DEV)          # one sampled prompt per class (:104-111)
    labels = torch.randint(0, classes, (n,), generator=g).to(DEV)

    def arm(model, fused):
        fi, ft = model.visual(inputs), model.text(prompts)
        s = model.logit_scale.exp()
        if fused:
            loss = contrastive_cross_entropy(l2_normalize(fi), l2_normalize(ft), s, labels, reduction="mean")
        else:
            fi = fi / fi.norm(dim=-1, keepdim=True)
            ft = ft / ft.norm(dim=-1, keepdim=True)
            loss = F.cross_entropy(s * fi @ ft.T, labels)
        loss.backward()
        return loss.detach(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    la, ga = arm(model_a, False)
    lb, gb = arm(model_b, True)
    torch.cuda.synchronize()
    assert abs(lb.item() - la.item()) < 1e-5 * abs(la.item())
    for k in ga:
        assert _rel(gb[k], ga[k]) < 1e-4, k
