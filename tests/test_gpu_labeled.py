"""Label-aware ClipLoss variants (clip/loss.py:123-192) on a real B200 against the oracle and the
fixtures recorded from the reference.  Tolerances as in test_gpu_parity.py: bf16 features 2e-3 (+ 2^-8 where autograd
stores a bf16 result), fp32 features 2e-5 on the loss and 1e-4 on the gradients."""
import os

import numpy as np
import pytest
import torch

from flyp_b200 import ClipLoss
from oracle import clip_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KW = {"soft": {}, "ignore": dict(ignore=True), "google": dict(google_sup_loss=True)}
NAMES = ["labeled_n24_d16_c5.npz", "labeled_n150_d64_c9.npz", "labeled_n40_d32_c40_s30.npz"]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def run(I, T, s, y, variant, grad_dtype=None, upstream=1.0):
    Ic = I.to(DEV).requires_grad_(True)
    Tc = T.to(DEV).requires_grad_(True)
    sc = torch.tensor(float(s), device=DEV, requires_grad=True)
    loss = ClipLoss(grad_dtype=grad_dtype)(Ic, Tc, sc, ground_labels=y.to(DEV), **KW[variant])
    assert loss.dim() == 0                                   # the reference returns a scalar for these variants
    (loss.float() * upstream).backward()
    torch.cuda.synchronize()
    f = lambda t: t.detach().double().cpu().numpy()
    return float(loss), f(Ic.grad), f(Tc.grad), float(sc.grad)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("variant", ["soft", "ignore", "google"])
def test_fp32_features_against_reference_fixture(golden_dir, name, variant):
    z = np.load(os.path.join(golden_dir, name))
    I, T = torch.from_numpy(z["I"]).float(), torch.from_numpy(z["T"]).float()
    loss, dI, dT, ds = run(I, T, float(z["scale"]), torch.from_numpy(z["y"]), variant)
    assert abs(loss - z[f"{variant}_loss"]) < 2e-5 * abs(z[f"{variant}_loss"])
    assert rel(dI, z[f"{variant}_dI"]) < 1e-4 and rel(dT, z[f"{variant}_dT"]) < 1e-4
    assert abs(ds - z[f"{variant}_ds"]) < 1e-4 * abs(z[f"{variant}_ds"])


def _inputs(n, d, n_cls, seed, mix=0.5):
    gen = torch.Generator().manual_seed(seed)
    I = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1)
    T = torch.nn.functional.normalize(mix * I + (1 - mix) * torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1), dim=-1)
    y = torch.randint(0, n_cls, (n,), generator=gen)
    return I, T, y


@pytest.mark.parametrize("n,d,n_cls", [(8, 512, 3), (37, 512, 4), (512, 512, 100), (640, 768, 7), (1024, 1024, 1000),
                                        (2048, 512, 2), (4096, 768, 300)])
@pytest.mark.parametrize("variant", ["soft", "ignore", "google"])
def test_bf16_shapes(n, d, n_cls, variant):
    I, T, y = _inputs(n, d, n_cls, seed=n + d)
    I, T = I.bfloat16(), T.bfloat16()
    s = 1 / 0.07
    loss, dI, dT, ds = run(I, T, s, y, variant, grad_dtype=torch.float32, upstream=0.7)
    In, Tn = I.double().numpy(), T.double().numpy()
    want = orc.labeled_clip_loss(In, Tn, s, y.numpy(), variant)
    wI, wT, wds = orc.labeled_clip_loss_grads(In, Tn, s, y.numpy(), variant)
    # the loss is returned in the feature dtype and autograd stores the gradients of bf16 leaves in bf16: one rounding
    # (2^-8) on top of the 2e-3 kernel bar, as for the default loss (test_gpu_parity.py)
    tol = 2e-3 + 2.0 ** -8
    assert abs(loss - want) < tol * abs(want)
    assert rel(dI, 0.7 * wI) < tol and rel(dT, 0.7 * wT) < tol
    assert abs(ds - 0.7 * wds) < tol * abs(wds)


@pytest.mark.parametrize("variant", ["soft", "ignore"])
def test_distinct_labels_give_the_default_loss(variant):
    I, T, _ = _inputs(300, 512, 5, seed=9)
    I, T = I.bfloat16(), T.bfloat16()
    y = torch.randperm(300)
    loss, dI, dT, ds = run(I, T, 14.0, y, variant, grad_dtype=torch.float32)
    Ic = I.to(DEV).requires_grad_(True); Tc = T.to(DEV).requires_grad_(True)
    sc = torch.tensor(14.0, device=DEV, requires_grad=True)
    ref = ClipLoss(grad_dtype=torch.float32)(Ic, Tc, sc).float().mean()
    ref.backward()
    assert abs(loss - float(ref)) < 2.0 ** -7 * abs(float(ref))
    # (both sides were stored in bf16 by autograd: they may differ by one bf16 step)
    assert rel(dI, Ic.grad.double().cpu().numpy()) < 2.0 ** -7 and rel(dT, Tc.grad.double().cpu().numpy()) < 2.0 ** -7


def test_world_size_above_one_raises():
    m = ClipLoss(world_size=2, rank=0)
    x = torch.randn(8, 64, device=DEV)
    with pytest.raises(NotImplementedError):
        m(x, x, torch.tensor(10.0, device=DEV), ground_labels=torch.arange(8, device=DEV))
