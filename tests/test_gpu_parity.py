"""Parity of the CUDA path (through the drop-in module / C ABI) against the oracle on a real B200.

Tolerances (BASELINE.json north_star): bf16 inputs with fp32 accumulation: loss and gradients within 2e-3; gathered
ordering and argmax predictions bit-exact.  The oracle is evaluated in float64 on the bf16-ROUNDED inputs (SURVEY.md
section 9.4).  Metric: max|got - want| / max|want| over the tensor.  The 2e-3 check is made on the kernels' fp32
outputs (C-ABI wrappers, `grad_dtype=torch.float32`); through the module, autograd stores loss and gradients of bf16
leaves in bf16, which adds one rounding (<= 2^-8 relative), so those are checked at 2^-8 + 2e-3.
"""
import os

import numpy as np
import pytest
import torch

import flyp_b200
from flyp_b200 import ClipLoss, ops
from oracle import clip_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 2e-3
DEV = "cuda:0"


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def rel_rows(a, b, floor=1e-3):
    """Worst PER-ROW relative error max_j|err_rj| / max_j|want_rj| over the rows whose magnitude is at least
    floor * max|want|: a global max-abs ratio alone would let small-magnitude gradient rows be arbitrarily wrong."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if b.ndim == 1:
        a, b = a[:, None], b[:, None]
    rmax = np.max(np.abs(b), axis=1)
    live = rmax >= floor * max(np.max(np.abs(b)), 1e-300)
    if not live.any():
        return 0.0
    return float(np.max(np.max(np.abs(a - b), axis=1)[live] / rmax[live]))


ROW_TOL = 3 * TOL      # per-row bar (rows above 1e-3 of the largest): the same order as the global one


def to_np(t):
    return t.detach().double().cpu().numpy()


def make_inputs(n, d, seed, mix=0.5, dtype=torch.bfloat16):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=gen)
    y = torch.randn(n, d, generator=gen)
    I = torch.nn.functional.normalize(x, dim=-1)
    T = torch.nn.functional.normalize(mix * I + (1 - mix) * torch.nn.functional.normalize(y, dim=-1), dim=-1)
    g = torch.rand(n, generator=gen) / n
    return I.to(dtype), T.to(dtype), g


def run_module(I, T, s, g, **kw):
    Ic = I.to(DEV).requires_grad_(True)
    Tc = T.to(DEV).requires_grad_(True)
    sc = torch.tensor(float(s), device=DEV, requires_grad=True)
    fn = ClipLoss(cache_labels=True, **kw)
    loss = fn(Ic, Tc, sc)
    (loss.float() * g.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    return loss, Ic.grad, Tc.grad, sc.grad


def run_abi_fp32(I, T, s, g):
    """Forward + backward through the C ABI wrappers with fp32 gradient outputs (autograd would round the gradients of
    bf16 leaves to bf16, hiding the kernel's own accuracy)."""
    Ic, Tc, gd = I.to(DEV), T.to(DEV), g.to(DEV)
    sc = torch.tensor([float(s)], device=DEV)
    row_lse, row_nll, col_stat, status = ops.clip_fwd_local(Ic, Tc, sc)
    col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, I.shape[0])
    dI, dT, ds = ops.clip_bwd_local(Ic, Tc, sc, 0, row_lse, row_nll, col_lse, col_nll, gd, gd,
                                    grad_dtype=torch.float32)
    torch.cuda.synchronize()
    return loss, dI, dT, ds


def run_step_fp32(I, T, s, g):
    """flyp_clip_fwd_step / flyp_clip_bwd_step with comm = NULL: the single-GPU loss as the module runs it."""
    from flyp_b200 import step
    Ic, Tc, gd = I.to(DEV), T.to(DEV), g.to(DEV)
    sc = torch.tensor([float(s)], device=DEV)
    loss, st = step.step_forward(None, Ic, Tc, sc, torch.float32)
    dI, dT, ds = step.step_backward(st, gd, 1.0, torch.float32, True, True, True)
    torch.cuda.synchronize()
    return loss, dI, dT, ds


BF16_EPS = 2.0 ** -8       # one round-to-nearest bf16 storage step


def check_against_oracle(I, T, s, g, tol=TOL):
    In, Tn, gn = to_np(I), to_np(T), to_np(g)
    want = orc.clip_loss(In, Tn, s)
    wdI, wdT, wds = orc.clip_loss_grads(In, Tn, s, gn)
    # (1) kernels, fp32 outputs: the north-star tolerance
    loss, dI, dT, ds = run_abi_fp32(I, T, s, g)
    assert rel(to_np(loss), want) < tol, "loss"
    assert rel(to_np(dI), wdI) < tol, "d image_features"
    assert rel(to_np(dT), wdT) < tol, "d text_features"
    assert abs(ds.item() - wds) <= tol * max(abs(wds), 1e-30), "d logit_scale"
    # ... and row by row: small-magnitude gradient rows and small losses are right in relative terms too
    assert rel_rows(to_np(loss), want) < 3 * tol, "loss, element-wise"
    assert rel_rows(to_np(dI), wdI) < 3 * tol, "d image_features, per row"
    assert rel_rows(to_np(dT), wdT) < 3 * tol, "d text_features, per row"
    # (1b) the whole-step entry points the module uses (one C call per direction), fp32 gradient outputs
    loss, dI, dT, ds = run_step_fp32(I, T, s, g)
    assert rel(to_np(loss), want) < tol and rel_rows(to_np(loss), want) < 3 * tol, "loss (step)"
    assert rel(to_np(dI), wdI) < tol and rel_rows(to_np(dI), wdI) < 3 * tol, "d image_features (step)"
    assert rel(to_np(dT), wdT) < tol and rel_rows(to_np(dT), wdT) < 3 * tol, "d text_features (step)"
    assert abs(ds.item() - wds) <= tol * max(abs(wds), 1e-30), "d logit_scale (step)"
    # (2) the drop-in module: loss and feature gradients come back in the feature dtype like the reference's
    #     (autograd casts); bf16 storage costs one extra rounding
    loss, dI, dT, ds = run_module(I, T, s, g)
    assert loss.shape == (I.shape[0],) and loss.dtype == I.dtype and dI.dtype == I.dtype
    assert rel(to_np(loss), want) < BF16_EPS + tol
    assert rel(to_np(dI), wdI) < BF16_EPS + tol
    assert rel(to_np(dT), wdT) < BF16_EPS + tol
    # the upstream gradient itself reaches backward rounded to bf16 (it is the gradient of a bf16 tensor)
    assert abs(ds.item() - wds) <= (BF16_EPS + tol) * max(abs(wds), 1e-30)


# ---------------------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", ["clip_w1_n24_d16_f64.npz", "clip_w1_n37_d64_s100_f64.npz",
                                  "clip_w1_n130_d72_f64.npz", "clip_w1_n1_d8_f64.npz"])
def test_golden_fixture_inputs(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    I = torch.tensor(z["I"]).bfloat16()
    T = torch.tensor(z["T"]).bfloat16()
    g = torch.tensor(z["g"]).float()
    s = float(z["scale"])
    # fp32 statistics straight from the C ABI: loss within 2e-3 of the oracle on the rounded inputs ...
    sc = torch.tensor([s], device=DEV)
    row_lse, row_nll, col_stat, status = ops.clip_fwd_local(I.to(DEV), T.to(DEV), sc)
    col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, I.shape[0])
    want = orc.clip_loss(to_np(I), to_np(T), s)
    assert rel(to_np(loss), want) < 1e-5          # fp32 statistics are far better than the bf16 bar
    # ... and of the reference's own output on the unrounded inputs up to the input rounding itself
    assert rel(to_np(loss), z["loss"]) < 5e-2
    check_against_oracle(I, T, s, g)


# ---------------------------------------------------------------------------------------------- shapes / edge cases
@pytest.mark.parametrize("n,d", [(1, 512), (8, 512), (37, 512), (255, 640), (512, 512), (640, 768), (1024, 1024),
                                 (4096, 768)])
@pytest.mark.parametrize("s", [1 / 0.07, 100.0])
def test_shapes_and_scales(n, d, s):
    I, T, g = make_inputs(n, d, seed=n + d, mix=0.5 if s < 50 else 0.2)
    check_against_oracle(I, T, s, g)


def test_independent_pairs_loss_is_log_n():
    n, d = 2048, 512
    I, T, g = make_inputs(n, d, seed=3, mix=0.0)
    loss, *_ = run_module(I, T, 1 / 0.07, g)
    assert abs(to_np(loss).mean() - np.log(n)) < 0.35      # log n + var(logit)/2
    check_against_oracle(I, T, 1 / 0.07, g)


def test_mean_reduction_like_the_flyp_loop():
    # src/models/flyp_loss.py:496-499: loss = torch.mean(clip_loss_fn(...)); loss.backward()
    n, d = 512, 512
    I, T, _ = make_inputs(n, d, seed=11)
    g = torch.full((n,), 1.0 / n)
    check_against_oracle(I, T, 1 / 0.07, g)


def test_default_bf16_gradients_are_the_rounded_oracle():
    n, d = 512, 512
    I, T, g = make_inputs(n, d, seed=5)
    Ic = I.to(DEV).requires_grad_(True); Tc = T.to(DEV).requires_grad_(True)
    sc = torch.tensor(1 / 0.07, device=DEV, requires_grad=True)
    loss = ClipLoss()(Ic, Tc, sc)
    (loss.float() * g.to(DEV)).sum().backward()
    assert Ic.grad.dtype == torch.bfloat16 and Tc.grad.dtype == torch.bfloat16
    wdI, wdT, _ = orc.clip_loss_grads(to_np(I), to_np(T), 1 / 0.07, to_np(g))
    # one bf16 rounding (2^-8) on top of the fp32-gradient tolerance
    assert rel(to_np(Ic.grad), wdI) < BF16_EPS + TOL
    assert rel(to_np(Tc.grad), wdT) < BF16_EPS + TOL


def test_unnormalised_inputs_take_the_robust_path():
    # row norms vary by 30x and logit_scale = 100: the fixed exponent window under/overflows, the kernels must notice
    # (status = 1) and recompute with exact per-tile maxima
    n, d = 300, 512
    gen = torch.Generator().manual_seed(9)
    I = (torch.randn(n, d, generator=gen) * torch.logspace(-1.5, 0.0, n)[:, None] * 0.2).bfloat16()
    T = (torch.randn(n, d, generator=gen) * 0.2).bfloat16()
    g = torch.rand(n, generator=gen) / n
    s = 100.0
    sc = torch.tensor([s], device=DEV)
    row_lse, row_nll, col_stat, status = ops.clip_fwd_local(I.to(DEV), T.to(DEV), sc)
    assert status.item() == 1
    col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, n)
    S = orc.logits(to_np(I), to_np(T), s)
    assert rel(to_np(row_lse), orc.logsumexp(S, 1)) < 1e-5
    assert rel(to_np(col_lse), orc.logsumexp(S, 0)) < 1e-5
    assert rel(to_np(loss), orc.clip_loss(to_np(I), to_np(T), s)) < 1e-5
    check_against_oracle(I, T, s, g)


def test_step_entry_point_reports_the_robust_path():
    from flyp_b200 import step
    n, d = 300, 512
    gen = torch.Generator().manual_seed(9)
    I = (torch.randn(n, d, generator=gen) * torch.logspace(-1.5, 0.0, n)[:, None] * 0.2).bfloat16()
    T = (torch.randn(n, d, generator=gen) * 0.2).bfloat16()
    sc = torch.tensor([100.0], device=DEV)
    loss, st, status = step.step_forward(None, I.to(DEV), T.to(DEV), sc, torch.float32, want_status=True)
    assert status.item() == 1
    assert rel(to_np(loss), orc.clip_loss(to_np(I), to_np(T), 100.0)) < 1e-5
    I2, T2, _ = make_inputs(n, d, seed=1)
    loss, st, status = step.step_forward(None, I2.to(DEV), T2.to(DEV), torch.tensor([1 / 0.07], device=DEV), torch.float32,
                                         want_status=True)
    assert status.item() == 0


def test_ce_head_ignore_index_like_torch():
    """F.cross_entropy (src/models/ce_ablation.py:123) skips targets equal to ignore_index = -100: loss 0, no gradient,
    and reduction='mean' averages over the remaining rows."""
    n, c, d = 200, 37, 256
    gen = torch.Generator().manual_seed(5)
    a = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(c, d, generator=gen), dim=-1).bfloat16()
    labels = torch.randint(0, c, (n,), generator=gen)
    labels[::7] = -100
    s = 1 / 0.07
    ac = a.to(DEV).requires_grad_(True); bc = b.to(DEV).requires_grad_(True)
    loss = flyp_b200.contrastive_cross_entropy(ac, bc, torch.tensor(s, device=DEV), labels.to(DEV), reduction="mean",
                                               grad_dtype=torch.float32)
    loss.backward()
    af = a.float().requires_grad_(True); bf = b.float().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(s * af @ bf.T, labels)
    want.backward()
    assert abs(loss.item() - want.item()) < (BF16_EPS + TOL) * abs(want.item())
    assert rel(to_np(ac.grad), to_np(af.grad)) < BF16_EPS + TOL and rel(to_np(bc.grad), to_np(bf.grad)) < BF16_EPS + TOL
    assert float(ac.grad[::7].abs().max()) == 0.0
    per_item = flyp_b200.contrastive_cross_entropy(a.to(DEV), b.to(DEV), torch.tensor(s, device=DEV), labels.to(DEV))
    assert float(per_item[::7].abs().max()) == 0.0


def test_tiny_losses_keep_relative_accuracy():
    # well separated positives: loss ~ 1e-3 while the logits are ~ 14; lse - diag must not be computed by subtraction
    n, d = 256, 512
    I, T, g = make_inputs(n, d, seed=21, mix=0.9)
    sc = torch.tensor([1 / 0.07], device=DEV)
    row_lse, row_nll, col_stat, _ = ops.clip_fwd_local(I.to(DEV), T.to(DEV), sc)
    _, _, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, n)
    want = orc.clip_loss(to_np(I), to_np(T), 1 / 0.07)
    assert want.max() < 0.05
    assert np.max(np.abs(to_np(loss) - want) / want) < 1e-4      # element-wise relative


# ---------------------------------------------------------------------------------------------- row sharding (emulated)
@pytest.mark.parametrize("world", [2, 4])
def test_row_sharded_statistics_merge_exactly(world):
    """The multi-GPU scheme on one GPU: every 'rank' processes its row block against all columns; only the O(B) column
    statistics are combined.  Result must equal the single-rank result and the oracle."""
    n, d, s = 1024, 512, 1 / 0.07
    I, T, g = make_inputs(n, d, seed=17)
    Ic, Tc = I.to(DEV), T.to(DEV)
    sc = torch.tensor([s], device=DEV)
    b = n // world
    stats, lses, nlls = [], [], []
    for r in range(world):
        row_lse, row_nll, col_stat, _ = ops.clip_fwd_local(Ic[r * b:(r + 1) * b], Tc, sc, r * b)
        stats.append(col_stat); lses.append(row_lse); nlls.append(row_nll)
    col_lse, col_nll, loss = ops.clip_fwd_finish(torch.cat(stats), world, torch.cat(nlls), n, 0)
    want = orc.clip_loss(to_np(I), to_np(T), s)
    assert rel(to_np(loss), want) < 1e-5
    row_lse_all, row_nll_all = torch.cat(lses), torch.cat(nlls)
    gd = g.to(DEV)
    wdI, wdT, wds = orc.clip_loss_grads(to_np(I), to_np(T), s, to_np(g))
    ds_total = 0.0
    for r in range(world):
        sl = slice(r * b, (r + 1) * b)
        d_img, _, d_s = ops.clip_bwd_local(Ic[sl], Tc, sc, r * b, row_lse_all[sl].contiguous(),
                                           row_nll_all[sl].contiguous(), col_lse, col_nll, gd[sl].contiguous(), gd,
                                           grad_dtype=torch.float32, need_txt=False)
        d_txt, _, _ = ops.clip_bwd_local(Tc[sl], Ic, sc, r * b, col_lse[sl].contiguous(), col_nll[sl].contiguous(),
                                         row_lse_all, row_nll_all, gd[sl].contiguous(), gd, grad_dtype=torch.float32,
                                         need_txt=False, need_scale=False)
        assert rel(to_np(d_img), wdI[sl]) < TOL
        assert rel(to_np(d_txt), wdT[sl]) < TOL
        ds_total += d_s.item()
    assert abs(ds_total - wds) < TOL * abs(wds)


# ---------------------------------------------------------------------------------------------- ce head
@pytest.mark.parametrize("name", ["ce_n20_c7_d16.npz", "ce_n150_c182_d64.npz"])
def test_ce_head_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    s = float(z["scale"])
    img = torch.tensor(z["img"]).bfloat16(); txt = torch.tensor(z["txt"]).bfloat16()
    labels = torch.tensor(z["labels"]); g = torch.tensor(z["g"]).float()
    ic = img.to(DEV).requires_grad_(True); tc = txt.to(DEV).requires_grad_(True)
    sc = torch.tensor(s, device=DEV, requires_grad=True)
    # src/models/ce_ablation.py:115-123 with the fused ops
    loss = flyp_b200.contrastive_cross_entropy(flyp_b200.l2_normalize(ic), flyp_b200.l2_normalize(tc), sc,
                                               labels.to(DEV), grad_dtype=torch.float32)
    (loss.float() * g.to(DEV)).sum().backward()
    # oracle on what the kernels saw: bf16 inputs, normalised and re-rounded to bf16
    imgn = to_np(torch.tensor(orc.l2_normalize(to_np(img))).bfloat16())
    txtn = to_np(torch.tensor(orc.l2_normalize(to_np(txt))).bfloat16())
    want = orc.cross_entropy(imgn, txtn, s, z["labels"])
    assert rel(to_np(loss), want) < 2.0 ** -8 + TOL
    dA, dB, ds = orc.cross_entropy_grads(imgn, txtn, s, z["labels"], z["g"])
    assert abs(sc.grad.item() - ds) < TOL * abs(ds)
    # through the normalisation backward (bf16 storage of the intermediate gradient: one extra rounding)
    assert rel(to_np(ic.grad), orc.l2_normalize_bwd(to_np(img), dA)) < 2.0 ** -7
    assert rel(to_np(tc.grad), orc.l2_normalize_bwd(to_np(txt), dB)) < 2.0 ** -7


# (10000, 182) and (9600, 1000): more row blocks than CTA pairs with a flat tail that has FEWER units than pairs (pairs with
# an empty range own no partial slot; found by tests/test_sched_host.py)
@pytest.mark.parametrize("n,c,d", [(512, 1000, 512), (256, 182, 512), (100, 37, 768), (10000, 182, 512), (9600, 1000, 256)])
def test_ce_head_shapes(n, c, d):
    gen = torch.Generator().manual_seed(n + c)
    a = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1).bfloat16()
    b = torch.nn.functional.normalize(torch.randn(c, d, generator=gen), dim=-1).bfloat16()
    labels = torch.randint(0, c, (n,), generator=gen)
    g = torch.rand(n, generator=gen) / n
    s = 1 / 0.07
    sc = torch.tensor([s], device=DEV)
    loss, lse = ops.ce_fwd(a.to(DEV), b.to(DEV), sc, labels.to(DEV))
    want = orc.cross_entropy(to_np(a), to_np(b), s, labels.numpy())
    assert rel(to_np(loss), want) < 1e-5
    d_a, d_b, d_s = ops.ce_bwd(a.to(DEV), b.to(DEV), sc, labels.to(DEV), 0, lse, loss, g.to(DEV),
                               grad_dtype=torch.float32)
    wA, wB, ws = orc.cross_entropy_grads(to_np(a), to_np(b), s, labels.numpy(), to_np(g))
    assert rel(to_np(d_a), wA) < TOL and rel(to_np(d_b), wB) < TOL and abs(d_s.item() - ws) < TOL * abs(ws)
    # mean reduction as in ce_ablation.py:123
    m = flyp_b200.contrastive_cross_entropy(a.to(DEV), b.to(DEV), sc, labels.to(DEV), reduction="mean")
    assert abs(m.item() - want.mean()) < 2.0 ** -8 * abs(want.mean()) + 1e-3


# ---------------------------------------------------------------------------------------------- normalise
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_l2_normalise(golden_dir, dtype):
    z = np.load(os.path.join(golden_dir, "l2norm.npz"))
    x = torch.tensor(z["x"]).to(dtype); dy = torch.tensor(z["dy"]).to(dtype)
    xc = x.to(DEV).requires_grad_(True)
    y = flyp_b200.l2_normalize(xc)
    (y.float() * dy.to(DEV).float()).sum().backward()
    tol = 2.0 ** -8 if dtype == torch.bfloat16 else 1e-6
    assert rel(to_np(y), orc.l2_normalize(to_np(x))) < tol
    assert rel(to_np(xc.grad), orc.l2_normalize_bwd(to_np(x), to_np(dy))) < 2 * tol
    big = torch.randn(4096, 768, generator=torch.Generator().manual_seed(1)).to(dtype)
    yb, inv = ops.l2norm_fwd(big.to(DEV))
    assert rel(to_np(yb), orc.l2_normalize(to_np(big))) < tol


def test_normalize_flag_of_the_module():
    n, d = 384, 512
    gen = torch.Generator().manual_seed(2)
    I = (2.0 * torch.randn(n, d, generator=gen)).bfloat16(); T = (0.5 * torch.randn(n, d, generator=gen)).bfloat16()
    sc = torch.tensor(1 / 0.07, device=DEV)
    got = ClipLoss(normalize=True)(I.to(DEV), T.to(DEV), sc)
    In = to_np(torch.tensor(orc.l2_normalize(to_np(I))).bfloat16())
    Tn = to_np(torch.tensor(orc.l2_normalize(to_np(T))).bfloat16())
    assert rel(to_np(got), orc.clip_loss(In, Tn, 1 / 0.07)) < 2.0 ** -8 + TOL


@pytest.mark.parametrize("n,d", [(300, 640), (640, 768), (257, 896), (1024, 1024)])
@pytest.mark.parametrize("impl", ["1", "2"])
def test_wide_features_both_sweep_kernels(n, d, impl, monkeypatch):
    """D > 512: the single-CTA sweep (FLYP_BWD_IMPL=1) and the two-pass CTA-pair sweep with streamed A chunks (=2; chosen
    automatically only for long sweeps) must both meet the bar, incl. the bubble-slot cases D = 640 / 896."""
    monkeypatch.setenv("FLYP_BWD_IMPL", impl)
    I, T, g = make_inputs(n, d, seed=n + d)
    check_against_oracle(I, T, 1 / 0.07, g)


# ---------------------------------------------------------------------------------------------- argmax (bit-exact)
def test_argmax_predictions_are_bit_exact():
    gen = torch.Generator().manual_seed(4)
    img = torch.nn.functional.normalize(torch.randn(4096, 512, generator=gen), dim=-1).bfloat16()
    cls = torch.nn.functional.normalize(torch.randn(1000, 512, generator=gen), dim=-1).bfloat16()
    cls[17] = cls[3]                      # an exact tie: the lowest index must win, as torch.argmax
    img[5] = cls[3]
    pred = flyp_b200.zero_shot_argmax(img.to(DEV), cls.to(DEV)).cpu().numpy()
    want = orc.argmax_predictions(to_np(img), to_np(cls))
    assert np.array_equal(pred, want)
    assert pred[5] == 3
    # symmetric-loss logits: row and column argmax of the square problem
    I, T, _ = make_inputs(1024, 512, seed=8)
    L = ops.debug_logits(I.to(DEV), T.to(DEV))
    S = to_np(I) @ to_np(T).T
    assert np.array_equal(L.argmax(1).cpu().numpy(), S.argmax(1))
    assert np.array_equal(L.argmax(0).cpu().numpy(), S.argmax(0))


@pytest.mark.parametrize("n,c,d,dtype", [(37, 182, 512, torch.bfloat16), (1000, 1000, 768, torch.bfloat16),
                                         (300, 129, 64, torch.bfloat16), (513, 1000, 512, torch.float32),
                                         (8192, 1000, 512, torch.bfloat16)])
def test_fused_argmax_equals_argmax_of_logits(n, c, d, dtype):
    """flyp_argmax (epilogue argmax, logits never written) == torch.argmax of the same kernel's dot products and of the
    float64 oracle, including exact ties (lowest index) and ragged shapes; src/models/eval.py:150-158."""
    gen = torch.Generator().manual_seed(n + c)
    img = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1).to(dtype)
    cls = torch.nn.functional.normalize(torch.randn(c, d, generator=gen), dim=-1).to(dtype)
    cls[c - 1] = cls[2]                   # ties across different column blocks / halves
    cls[70 % c] = cls[2]
    img[n // 2] = cls[2]
    idx, mx = ops.argmax(img.to(DEV), cls.to(DEV), return_max=True)
    L = ops.debug_logits(img.to(DEV), cls.to(DEV))
    assert torch.equal(idx, L.argmax(dim=1))
    assert torch.equal(mx, L.max(dim=1).values)
    assert idx[n // 2].item() == min(2, 70 % c)
    if dtype == torch.bfloat16:
        want = orc.argmax_predictions(to_np(img), to_np(cls))
        assert np.array_equal(idx.cpu().numpy(), want)


# ---------------------------------------------------------------------------------------------- full size properties
def _sampled_full_size_check(n, d, s, seed, n_samples=48):
    """Size-independent parity at full size: the float64 reference of tools/sampled_check.py (chunked logsumexp over the
    whole matrix on the GPU, independent of the kernels) for sampled items - loss, d image_features AND d text_features
    row by row - plus the d(scale) identities over all rows."""
    from tools import sampled_check as sck
    I, T, g = make_inputs(n, d, seed=seed)
    Ic, Tc, gd = I.to(DEV), T.to(DEV), g.to(DEV)
    sc = torch.tensor([s], device=DEV)
    row_lse, row_nll, col_stat, status = ops.clip_fwd_local(Ic, Tc, sc)
    col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, n)
    assert status.item() == 0
    idx = torch.tensor(np.random.default_rng(0).choice(n, n_samples, replace=False), device=DEV)
    lse64 = sck.full_lse(Ic, Tc, s)
    want_loss, want_dI, want_dT = sck.sampled_reference(Ic, Tc, s, gd, idx, lse=lse64)
    assert sck.row_errors(row_lse.double(), lse64[0])[0] < 1e-5
    assert sck.row_errors(col_lse.double(), lse64[1])[0] < 1e-5
    assert sck.row_errors(loss[idx], want_loss)[0] < 1e-5
    d_img, d_txt, d_s = ops.clip_bwd_local(Ic, Tc, sc, 0, row_lse, row_nll, col_lse, col_nll, gd, gd,
                                           grad_dtype=torch.float32)
    for got, want, what in ((d_img[idx], want_dI, "d image_features"), (d_txt[idx], want_dT, "d text_features")):
        glob, per_row = sck.row_errors(got, want)
        assert glob < TOL and per_row < ROW_TOL, (what, glob, per_row)
    # the whole-step entry points (what the module runs) at the same size
    from flyp_b200 import step
    loss2, st = step.step_forward(None, Ic, Tc, sc, torch.float32)
    dI2, dT2, ds2 = step.step_backward(st, gd, 1.0, torch.float32, True, True, True)
    assert sck.row_errors(loss2[idx], want_loss)[0] < 1e-5
    for got, want, what in ((dI2[idx], want_dI, "d image_features (step)"), (dT2[idx], want_dT, "d text_features (step)")):
        glob, per_row = sck.row_errors(got, want)
        assert glob < TOL and per_row < ROW_TOL, (what, glob, per_row)
    # d(scale) identities: sum_i <dI_i, I_i> / s = sum_j <dT_j, T_j> / s = ds
    a = (d_img.double() * Ic.double()).sum().item() / s
    b = (d_txt.double() * Tc.double()).sum().item() / s
    assert abs(a - d_s.item()) < 1e-3 * abs(a) and abs(b - d_s.item()) < 1e-3 * abs(a)
    assert abs(ds2.item() - d_s.item()) < 1e-4 * abs(a)
    return Ic, Tc, sc, loss


def test_full_size_properties():
    """BASELINE config B = 32768, D = 512 bf16 (the oracle cannot form 32768^2 logits)."""
    n = 32768
    Ic, Tc, sc, loss = _sampled_full_size_check(n, 512, 1 / 0.07, seed=0)
    # the loss is permutation-equivariant: permuting the pairs permutes the per-item losses
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1)).to(DEV)
    r2, n2, c2, _ = ops.clip_fwd_local(Ic[perm].contiguous(), Tc[perm].contiguous(), sc)
    _, _, loss_p = ops.clip_fwd_finish(c2, 1, n2, n)
    assert rel(to_np(loss_p), to_np(loss[perm])) < 1e-5


@pytest.mark.parametrize("n,d", [(65536, 512), (32768, 1024), (16384, 768)])
def test_sweep_shapes_sampled_rows(n, d):
    """The other points of the BASELINE sweep (B up to 64k, D = 1024: the two-pass pair sweep), row by row."""
    _sampled_full_size_check(n, d, 1 / 0.07, seed=n // 1024 + d, n_samples=32)


# ---------------------------------------------------------------------------------------------- fp32 features
# north star: fp32 inputs -> loss within 1e-5 relative, gradients within 1e-4 relative.  fp32 features are evaluated
# as split bf16 / fp16 planes on the tensor cores (six-term products, fp32 accumulation).
def check_fp32(I, T, s, g):
    In, Tn, gn = to_np(I), to_np(T), to_np(g)
    want = orc.clip_loss(In, Tn, s)
    wdI, wdT, wds = orc.clip_loss_grads(In, Tn, s, gn)
    loss, dI, dT, ds = run_module(I, T, s, g)
    assert loss.dtype == torch.float32 and dI.dtype == torch.float32
    assert rel(to_np(loss), want) < 1e-5, "loss"
    assert rel(to_np(dI), wdI) < 1e-4, "d image_features"
    assert rel(to_np(dT), wdT) < 1e-4, "d text_features"
    assert abs(ds.item() - wds) <= 1e-4 * abs(wds), "d logit_scale"


@pytest.mark.parametrize("name", ["clip_w1_n24_d16_f64.npz", "clip_w1_n37_d64_s100_f64.npz",
                                  "clip_w1_n130_d72_f64.npz", "clip_w1_n24_d16_f32.npz"])
def test_fp32_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    I = torch.tensor(z["I"]).float(); T = torch.tensor(z["T"]).float(); g = torch.tensor(z["g"]).float()
    check_fp32(I, T, float(z["scale"]), g)
    # and directly against what the reference itself produced on (essentially) these inputs
    loss, dI, dT, ds = run_module(I, T, float(z["scale"]), g)
    assert rel(to_np(loss), z["loss"]) < 2e-5
    assert rel(to_np(dI), z["dI"]) < 2e-4 and rel(to_np(dT), z["dT"]) < 2e-4


@pytest.mark.parametrize("n,d,s", [(512, 512, 1 / 0.07), (300, 768, 100.0), (1024, 512, 1 / 0.07), (37, 520, 1 / 0.07)])
def test_fp32_shapes(n, d, s):
    # (at s = 100 the positives are kept weak so that the losses are O(1): a loss of 1e-5 on logits of magnitude 100 is
    #  below what ANY fp32 evaluation, the reference's included, can resolve to 1e-5 relative)
    I, T, g = make_inputs(n, d, seed=n + d + 1, mix=0.5 if s < 50 else 0.08, dtype=torch.float32)
    check_fp32(I, T, s, g)


def test_fp32_ce_head(golden_dir):
    z = np.load(os.path.join(golden_dir, "ce_n150_c182_d64.npz"))
    s = float(z["scale"])
    a = torch.tensor(z["imgn"]).float(); b = torch.tensor(z["txtn"]).float()
    labels = torch.tensor(z["labels"]); g = torch.tensor(z["g"]).float()
    ac = a.to(DEV).requires_grad_(True); bc = b.to(DEV).requires_grad_(True)
    sc = torch.tensor(s, device=DEV, requires_grad=True)
    loss = flyp_b200.contrastive_cross_entropy(ac, bc, sc, labels.to(DEV))
    (loss * g.to(DEV)).sum().backward()
    assert rel(to_np(loss), z["loss"]) < 1e-5
    assert rel(to_np(ac.grad), z["d_imgn"]) < 1e-4 and rel(to_np(bc.grad), z["d_txtn"]) < 1e-4
    assert abs(sc.grad.item() - float(z["ds"])) < 1e-4 * abs(float(z["ds"]))
