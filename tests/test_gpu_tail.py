"""Fused encoder tail: project_normalize(x, proj) against the reference's operators
(clip/model.py:242-243,359 projection, :375-376 normalisation) - forward and the gradients w.r.t. the pooled features and
the projection matrix - for the ViT-B/16 (768 -> 512, 512 -> 512) and ViT-L/14 (1024 -> 768: two-CTA cluster) tails,
ragged row counts, fp32 (1e-5 / 1e-4) and bf16 (2e-3 + storage rounding); and the FLYP step with the fused tail against
the same step with the reference operators."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-300))


def ref_tail(x, w):
    y = x @ w
    return y / y.norm(dim=-1, keepdim=True)


@pytest.mark.parametrize("n,k,n_out", [(512, 768, 512), (512, 512, 512), (37, 768, 512), (300, 1024, 768), (1000, 512, 1024),
                                       (128, 256, 64), (4096, 768, 512)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_project_normalize_matches_reference_ops(n, k, n_out, dtype):
    import flyp_b200
    g = torch.Generator().manual_seed(n + k + n_out)
    x = torch.randn(n, k, generator=g).to(dtype)
    w = (k ** -0.5 * torch.randn(k, n_out, generator=g)).to(dtype)
    dy = torch.randn(n, n_out, generator=g).to(dtype)
    xa = x.to(DEV).requires_grad_(True); wa = w.to(DEV).requires_grad_(True)
    y = flyp_b200.project_normalize(xa, wa)
    (y.float() * dy.to(DEV).float()).sum().backward()
    # reference operators in float64 on the same (rounded) inputs
    xr = x.to(DEV).double().requires_grad_(True); wr = w.to(DEV).double().requires_grad_(True)
    yr = ref_tail(xr, wr)
    (yr * dy.to(DEV).double()).sum().backward()
    torch.cuda.synchronize()
    assert y.dtype == dtype and y.shape == (n, n_out)
    if dtype == torch.float32:
        assert rel(y, yr) < 1e-5
        assert rel(xa.grad, xr.grad) < 1e-4 and rel(wa.grad, wr.grad) < 1e-4
    else:
        # bf16 storage of y (2^-9 relative) on top of the 2e-3 bar; the gradients pass through bf16 GEMMs of the
        # rounded dz (library calls, fp32 accumulation)
        assert rel(y, yr) < 2.0 ** -8 + 2e-3
        assert rel(xa.grad, xr.grad) < 2.0 ** -6 and rel(wa.grad, wr.grad) < 2.0 ** -6
    # rows are unit vectors
    assert float((y.float().norm(dim=-1) - 1).abs().max()) < (1e-5 if dtype == torch.float32 else 1e-2)


def test_bf16_output_from_fp32_towers_and_fp16_copy():
    from flyp_b200 import tail
    g = torch.Generator().manual_seed(3)
    x = torch.randn(640, 768, generator=g).to(DEV); w = (768 ** -0.5 * torch.randn(768, 512, generator=g)).to(DEV)
    y, inv, _ = tail.project_normalize_fwd(x, w, out_dtype=torch.bfloat16)
    want = ref_tail(x.double(), w.double())
    assert y.dtype == torch.bfloat16 and rel(y, want) < 2.0 ** -8
    z = x.double() @ w.double()
    assert rel(inv, 1.0 / z.norm(dim=-1)) < 1e-5
    xb, wb = x.bfloat16(), w.bfloat16()
    y2, _, y16 = tail.project_normalize_fwd(xb, wb, want_f16=True)
    # the fp16 copy holds the ROUNDED bf16 features: exact wherever fp16 has the range (|v| >= 2^-14), else to within
    # half an fp16 subnormal step
    diff = (y16.float() - y2.float()).abs()
    assert float(diff.max()) <= 2.0 ** -25 and float(diff[y2.float().abs() >= 2.0 ** -14].max()) == 0.0


def test_finetune_step_with_fused_tail_matches_reference_ops():
    """BASELINE configuration 5, reduced depth (the towers are out of scope): src/models/flyp_loss.py:426,495-500 with the
    fused tail + fused loss against the same step with the reference operator sequence; every parameter gradient."""
    from flyp_b200 import ClipLoss
    from flyp_b200.finetune import StepLog, TwoTowerEncoder, finetune_step
    from oracle import torch_port
    torch.manual_seed(0)
    kw = dict(vision_layers=2, text_layers=2, image_size=64, patch=16, context=16, vocab=1000)
    model_a = TwoTowerEncoder(fused_tail=False, **kw).to(DEV)
    model_b = copy.deepcopy(model_a)
    model_b.fused_tail = True
    g = torch.Generator().manual_seed(1)
    n = 256
    image = torch.randn(n, 3, 64, 64, generator=g).to(DEV)
    text = torch.randint(1, 999, (n, 16), generator=g).to(DEV)
    ids = torch.arange(n, device=DEV)

    def reference_loss(fi, ft, s):
        return torch_port.clip_loss_reference_ops(fi, ft, s)

    opt_a = torch.optim.AdamW(model_a.parameters(), lr=1e-5, weight_decay=0.1)
    opt_b = torch.optim.AdamW(model_b.parameters(), lr=1e-5, weight_decay=0.1)
    log = StepLog()
    la, pa = finetune_step(model_a, reference_loss, opt_a, image, text)
    ga = {k: p.grad.detach().clone() for k, p in model_a.named_parameters()}
    lb, pb = finetune_step(model_b, ClipLoss(cache_labels=True), opt_b, image, text, image_ids=ids, log=log)
    gb = {k: p.grad.detach().clone() for k, p in model_b.named_parameters()}
    torch.cuda.synchronize()
    assert rel(pb, pa) < 1e-5 and abs(lb.item() - la.item()) < 1e-5 * abs(la.item())
    # 1e-4 is the bar on the operator's own gradients (checked in test_project_normalize_* and the loss parity tests);
    # what reaches the tower parameters has additionally gone through the towers' fp32 backward (atomics in the
    # layer-norm / embedding gradients make it run-to-run noisy at the 1e-5 level): twice the bar there
    for k in ga:
        assert rel(gb[k], ga[k]) < 2e-4, k
    pairs, mean = log.fetch()                       # the ONE device-to-host transfer (flyp_loss.py:503-513 does it per step)
    assert len(pairs) == n and pairs[5][0] == 5 and abs(mean - la.item()) < 1e-5 * abs(la.item())
    assert log.fetch() == ([], float("nan")) or log.steps == 0
