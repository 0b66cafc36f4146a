"""The kept-dS backward (single rank, >= 1024 pairs, bf16, D a multiple of 128): the first sweep writes its staged fp16
dS tiles to the workspace and d_text is the product dS^T . I over them (csrc/clip_dst_gemm.cu) instead of a second sweep.
Parity against the fp64 oracle on the bf16-rounded inputs (2e-3, as test_gpu_parity.py), per row as well, and agreement
with the two-sweep path (FLYP_KEEP_DS=0, run in a subprocess: the switch is read once per process)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from flyp_b200 import ClipLoss, ops
from oracle import clip_oracle as orc
from test_gpu_parity import DEV, ROW_TOL, TOL, make_inputs, rel, rel_rows, run_abi_fp32, to_np

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,d", [(1024, 512), (1100, 512), (2048, 256), (1536, 768), (2048, 1024), (1024, 128),
                                  (4096, 512), (3000, 384)])
@pytest.mark.parametrize("s", [1 / 0.07, 100.0])
def test_kept_ds_backward_matches_oracle(n, d, s):
    # (at scale 100 well-matched pairs saturate: gradients of 1e-27; weakly matched ones keep the problem non-trivial)
    I, T, g = make_inputs(n, d, seed=n + d, mix=0.5 if s < 50 else 0.15)
    loss, dI, dT, ds = run_abi_fp32(I, T, s, g)
    In, Tn, gn = I.double().numpy(), T.double().numpy(), g.double().numpy()
    wI, wT, wds = orc.clip_loss_grads(In, Tn, s, gn)
    assert rel(to_np(dI), wI) < TOL and rel(to_np(dT), wT) < TOL
    assert rel_rows(to_np(dI), wI) < ROW_TOL and rel_rows(to_np(dT), wT) < ROW_TOL
    assert abs(float(ds) - wds) < TOL * abs(wds)


def test_kept_ds_through_the_module_bf16_gradients():
    n, d, s = 2048, 512, 1 / 0.07
    I, T, g = make_inputs(n, d, seed=5)
    Ic = I.to(DEV).requires_grad_(True); Tc = T.to(DEV).requires_grad_(True)
    sc = torch.tensor(float(s), device=DEV, requires_grad=True)
    loss = ClipLoss(cache_labels=True)(Ic, Tc, sc)
    (loss.float() * g.to(DEV)).sum().backward()
    wI, wT, wds = orc.clip_loss_grads(I.double().numpy(), T.double().numpy(), s, g.double().numpy())
    tol = TOL + 2.0 ** -8
    assert rel(to_np(Ic.grad), wI) < tol and rel(to_np(Tc.grad), wT) < tol
    assert abs(float(sc.grad) - wds) < tol * abs(wds)


_CHILD = """
import sys, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from test_gpu_parity import make_inputs, run_abi_fp32
I, T, g = make_inputs(2048, 512, seed=11)
loss, dI, dT, ds = run_abi_fp32(I, T, 1 / 0.07, g)
np.save(sys.argv[1], dT.detach().float().cpu().numpy())
"""


def test_kept_ds_agrees_with_the_two_sweep_path(tmp_path):
    outs = []
    # default (sweep + product), two sweeps, and the unfused plan (dS kernel + two products; off by default)
    for tag, extra in (("keep", {}), ("sweeps", {"FLYP_KEEP_DS": "0"}), ("unfused", {"FLYP_UNFUSED": "1"})):
        f = str(tmp_path / f"dt_{tag}.npy")
        env = dict(os.environ, **extra)
        subprocess.run([sys.executable, "-c", _CHILD.format(root=ROOT), f], check=True, env=env, timeout=600)
        outs.append(np.load(f))
    # the same fp16-staged dS values enter all products: the results differ by accumulation order only
    assert rel(outs[0], outs[1]) < 2e-4 and rel(outs[0], outs[2]) < 2e-4
    assert not np.array_equal(outs[0], np.zeros_like(outs[0]))


_CHILD_FULL = """
import sys, numpy as np, torch
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
from flyp_b200 import _lib
from test_gpu_parity import make_inputs, check_against_oracle
assert _lib.load().flyp_clip_backward_plan(1536, 1536, 512, _lib.FLYP_BF16) == 2
for n, d, s in ((1536, 512, 1 / 0.07), (1100, 384, 100.0), (4096, 256, 1 / 0.07)):
    I, T, g = make_inputs(n, d, seed=n, mix=0.5 if s < 50 else 0.15)
    check_against_oracle(I, T, s, g)
print("ok")
"""


def test_unfused_plan_matches_oracle():
    """FLYP_UNFUSED=1: the dS kernel (the forward's pipeline with the dS epilogue) + two products, against the oracle through
    the C-ABI wrappers, the whole-step entry points and the module (loss, d image, d text, d scale, per row)."""
    env = dict(os.environ, FLYP_UNFUSED="1")
    res = subprocess.run([sys.executable, "-c", _CHILD_FULL.format(root=ROOT)], env=env, timeout=600, capture_output=True,
                         text=True)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-1500:] + res.stderr[-1500:]


def test_kept_ds_with_unnormalised_inputs_two_exponential_form():
    """Row norms vary by 30x and logit_scale = 100: the forward takes its robust path and the logsumexps spread so far
    that the backward cannot use the one-exponential form - the kept dS then holds the two-exponential values, and the
    product over it must still match the oracle."""
    n, d, s = 1280, 512, 100.0
    gen = torch.Generator().manual_seed(9)
    I = (torch.randn(n, d, generator=gen) * torch.logspace(-1.5, 0.0, n)[:, None] * 0.2).bfloat16()
    T = (torch.randn(n, d, generator=gen) * 0.2).bfloat16()
    g = torch.rand(n, generator=gen) / n
    loss, dI, dT, ds = run_abi_fp32(I, T, s, g)
    In, Tn = I.double().numpy(), T.double().numpy()
    assert rel(to_np(loss), orc.clip_loss(In, Tn, s)) < 1e-5
    wI, wT, wds = orc.clip_loss_grads(In, Tn, s, g.double().numpy())
    assert rel(to_np(dI), wI) < TOL and rel(to_np(dT), wT) < TOL
    assert abs(float(ds) - wds) < TOL * abs(wds)


def test_kept_ds_only_one_gradient_wanted_runs_a_plain_sweep():
    n, d, s = 2048, 512, 1 / 0.07
    I, T, g = make_inputs(n, d, seed=21)
    Ic, Tc, gd = I.to(DEV), T.to(DEV), g.to(DEV)
    sc = torch.tensor([float(s)], device=DEV)
    row_lse, row_nll, col_stat, status = ops.clip_fwd_local(Ic, Tc, sc)
    col_lse, col_nll, loss = ops.clip_fwd_finish(col_stat, 1, row_nll, n)
    wI, wT, wds = orc.clip_loss_grads(I.double().numpy(), T.double().numpy(), s, g.double().numpy())
    dI, none, ds = ops.clip_bwd_local(Ic, Tc, sc, 0, row_lse, row_nll, col_lse, col_nll, gd, gd, grad_dtype=torch.float32,
                                      need_txt=False)
    assert none is None and rel(to_np(dI), wI) < TOL and abs(float(ds) - wds) < TOL * abs(wds)
    none, dT, _ = ops.clip_bwd_local(Ic, Tc, sc, 0, row_lse, row_nll, col_lse, col_nll, gd, gd, grad_dtype=torch.float32,
                                     need_img=False, need_scale=False)
    assert none is None and rel(to_np(dT), wT) < TOL
