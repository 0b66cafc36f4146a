"""TEST-ONLY stand-in for flyp_b200.ops with the same partial-statistics contract (include/flyp_clip.h), computed on
the CPU in float64 from the oracle's formulas.  It exists so that the multi-rank orchestration in flyp_b200/loss.py
(sharding by rows, the O(B) statistics exchange, gradient routing for the four flag combinations) can be exercised
under gloo with world_size 2 on a machine without a GPU.  The product never imports it."""
import numpy as np
import torch

LOG2E = 1.4426950408889634


def _scale_tensor(scale, device):
    if not torch.is_tensor(scale):
        return torch.tensor([float(scale)], dtype=torch.float32, device=device)
    return scale.detach().reshape(-1)[:1].to(torch.float32)


def _np(t):
    return t.detach().double().cpu().numpy()


def _lse(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    return (np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m).squeeze(axis)


def clip_fwd_local(img, txt, scale, row_offset=0, workspace=None):
    I, T, s = _np(img), _np(txt), float(scale.reshape(-1)[0])
    S = s * (I @ T.T)
    n_rows, n_cols = S.shape
    pos = np.arange(n_rows) + row_offset
    diag = S[np.arange(n_rows), pos]
    row_lse = _lse(S, 1)
    row_nll = row_lse - diag
    X = S * LOG2E
    Xm = X.copy()
    Xm[np.arange(n_rows), pos] = -np.inf           # positives are excluded from the partial column sums
    m = np.max(np.where(np.isfinite(Xm), Xm, -1e300), axis=0)
    ssum = np.sum(np.exp2(Xm - m[None, :]), axis=0)
    t = np.full(n_cols, -np.inf)
    t[pos] = diag * LOG2E
    col_stat = np.concatenate([m, ssum, t])
    f = lambda a: torch.tensor(a, dtype=torch.float32)
    return f(row_lse), f(row_nll), f(col_stat), torch.zeros(1, dtype=torch.int32)


def clip_fwd_finish(col_stat_all, world, row_nll, n_cols, row_offset=0, loss_dtype=torch.float32):
    cs = _np(col_stat_all).reshape(world, 3, n_cols)
    m = np.max(cs[:, 0], axis=0)
    ssum = np.sum(cs[:, 1] * np.exp2(cs[:, 0] - m[None, :]), axis=0)
    t = np.max(cs[:, 2], axis=0)
    a = np.log2(ssum) + m
    lse2 = np.logaddexp2(a, t)
    col_lse = lse2 / LOG2E
    col_nll = (lse2 - t) / LOG2E
    n_rows = row_nll.numel()
    loss = 0.5 * (_np(row_nll) + col_nll[row_offset:row_offset + n_rows])
    f = lambda a: torch.tensor(a, dtype=torch.float32)
    return f(col_lse), f(col_nll), f(loss).to(loss_dtype)


def clip_bwd_local(img, txt, scale, row_offset, row_lse, row_nll, col_lse, col_nll, g_row, g_col, grad_mul=1.0,
                   grad_dtype=None, need_img=True, need_txt=True, need_scale=True, workspace=None):
    I, T, s = _np(img), _np(txt), float(scale.reshape(-1)[0])
    S = s * (I @ T.T)
    n_rows, n_cols = S.shape
    Pr = np.exp(S - _np(row_lse)[:, None])
    Pc = np.exp(S - _np(col_lse)[None, :])
    onehot = np.zeros_like(S)
    onehot[np.arange(n_rows), np.arange(n_rows) + row_offset] = 1.0
    dS = 0.5 * _np(g_row)[:, None] * (Pr - onehot) + 0.5 * _np(g_col)[None, :] * (Pc - onehot)
    gdt = img.dtype if grad_dtype is None else grad_dtype
    d_img = torch.tensor(grad_mul * s * (dS @ T)).to(gdt)
    d_txt = torch.tensor(grad_mul * s * (dS.T @ I)).to(gdt)
    d_s = torch.tensor([np.sum(dS * (I @ T.T))], dtype=torch.float32)
    return d_img, d_txt, d_s


def ce_fwd(a, b, scale, labels, label_offset=0, workspace=None):
    A, B, s = _np(a), _np(b), float(scale.reshape(-1)[0])
    S = s * (A @ B.T)
    lab = labels.cpu().numpy() if labels is not None else np.arange(S.shape[0]) + label_offset
    lse = _lse(S, 1)
    loss = lse - S[np.arange(S.shape[0]), lab]
    return torch.tensor(loss, dtype=torch.float32), torch.tensor(lse, dtype=torch.float32)


def ce_bwd(a, b, scale, labels, label_offset, lse, loss, g, grad_dtype=None, need_a=True, need_b=True,
           need_scale=True, workspace=None):
    A, B, s = _np(a), _np(b), float(scale.reshape(-1)[0])
    S = s * (A @ B.T)
    lab = labels.cpu().numpy() if labels is not None else np.arange(S.shape[0]) + label_offset
    P = np.exp(S - _np(lse)[:, None])
    onehot = np.zeros_like(S)
    onehot[np.arange(S.shape[0]), lab] = 1.0
    dS = _np(g)[:, None] * (P - onehot)
    gdt = a.dtype if grad_dtype is None else grad_dtype
    return (torch.tensor(s * (dS @ B)).to(gdt), torch.tensor(s * (dS.T @ A)).to(gdt),
            torch.tensor([np.sum(dS * (A @ B.T))], dtype=torch.float32))


def install(monkeypatch_target):
    """Replace the CUDA-backed functions of flyp_b200.ops by the CPU stand-ins (tests only)."""
    import flyp_b200.ops as ops
    for name in ("_scale_tensor", "clip_fwd_local", "clip_fwd_finish", "clip_bwd_local", "ce_fwd", "ce_bwd"):
        monkeypatch_target(ops, name, globals()[name])
