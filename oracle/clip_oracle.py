"""ORACLE - test infrastructure only.  A CPU (numpy, float64) restatement of the reference algorithm of the FLYP
contrastive-loss hot path.  Nothing under flyp_b200/ may import this module: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg use it, and only as the checker.

Pinned against the reference itself: tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the
unmodified /root/reference/clip/loss.py (torch CPU, fp64/fp32 autograd, and 2-rank gloo runs for gather semantics);
tests/test_oracle.py checks every function here against those fixtures.  The reference repository ships no tests or
golden vectors of its own, so these fixtures are the pin.

Each function cites the reference lines it restates (paths relative to joliang17/FLYP).
"""
from __future__ import annotations

import numpy as np


def _f64(x):
    return np.asarray(x, dtype=np.float64)


def logsumexp(x, axis):
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return (np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m).squeeze(axis)


def l2_normalize(x):
    """clip/model.py:375-376, src/models/ce_ablation.py:115-118: x / x.norm(dim=-1, keepdim=True), no epsilon."""
    x = _f64(x)
    return x / np.sqrt(np.sum(x * x, axis=-1, keepdims=True))


def l2_normalize_bwd(x, dy):
    """Gradient of l2_normalize: (dy - y <y, dy>) / ||x||."""
    x, dy = _f64(x), _f64(dy)
    nrm = np.sqrt(np.sum(x * x, axis=-1, keepdims=True))
    y = x / nrm
    return (dy - y * np.sum(y * dy, axis=-1, keepdims=True)) / nrm


def gather_features(blocks):
    """clip/loss.py:66-67: rank-major concatenation; global row = rank * b + local row."""
    return np.concatenate([np.asarray(b) for b in blocks], axis=0)


def logits(image, text, scale):
    """clip/loss.py:117 (world_size == 1) / :113 (gathered): logit_scale * image_features @ text_features.T"""
    return float(scale) * (_f64(image) @ _f64(text).T)


def clip_loss(image, text, scale):
    """clip/loss.py:117-118,195-209 with world_size == 1: per-item
    (CE(S, arange, 'none') + CE(S.T, arange, 'none')) / 2, S = scale * I @ T.T."""
    S = logits(image, text, scale)
    d = np.diagonal(S)
    return 0.5 * ((logsumexp(S, 1) - d) + (logsumexp(S, 0) - d))


def clip_loss_grads(image, text, scale, g):
    """Closed-form autograd of clip_loss for an upstream gradient vector g:
    dS = g_i/2 (P_row - I) + g_j/2 (P_col - I);  dI = s dS T;  dT = s dS^T I;  ds = sum(dS * S) / s."""
    I, T, g = _f64(image), _f64(text), _f64(g)
    s = float(scale)
    S = s * (I @ T.T)
    n = S.shape[0]
    Pr = np.exp(S - logsumexp(S, 1)[:, None])
    Pc = np.exp(S - logsumexp(S, 0)[None, :])
    eye = np.eye(n)
    dS = 0.5 * g[:, None] * (Pr - eye) + 0.5 * g[None, :] * (Pc - eye)
    return s * (dS @ T), s * (dS.T @ I), float(np.sum(dS * (I @ T.T)))


def cross_entropy(a, b, scale, labels):
    """src/models/ce_ablation.py:122-123 (reduction='none' form) and the local_loss blocks clip/loss.py:109-111:
    CE(scale * a @ b.T, labels) per row."""
    S = logits(a, b, scale)
    labels = np.asarray(labels, dtype=np.int64)
    return logsumexp(S, 1) - S[np.arange(S.shape[0]), labels]


def cross_entropy_grads(a, b, scale, labels, g):
    A, B, g = _f64(a), _f64(b), _f64(g)
    s = float(scale)
    S = s * (A @ B.T)
    P = np.exp(S - logsumexp(S, 1)[:, None])
    onehot = np.zeros_like(S)
    onehot[np.arange(S.shape[0]), np.asarray(labels, dtype=np.int64)] = 1.0
    dS = g[:, None] * (P - onehot)
    return s * (dS @ B), s * (dS.T @ A), float(np.sum(dS * (A @ B.T)))


def _soft_targets(E):
    """clip/loss.py:126-127,189-190: equal_labels is cast `.type(torch.float)`, so the soft targets E / c are float32
    numbers whatever the feature dtype (their rows sum to 1 only to ~1e-8)."""
    E32 = E.astype(np.float32)
    return (E32 / E32.sum(1, keepdims=True)).astype(np.float64)


def labeled_clip_loss(image, text, scale, labels, variant="soft"):
    """clip/loss.py:123-192 (world_size == 1, `ground_labels` given): the scalar label-aware losses.
    E_ij = [y_i == y_j] (:124-127), c_i = sum_j E_ij.
      "soft"    (:188-192) (CE(S, E / c) + CE(S.T, E / c)) / 2 with probability targets (mean over rows);
      "ignore"  (:132-159) -log(e_ii / sum_{j: E_ij = 0 or j = i} e_ij), both directions, means, / 2;
      "google"  (:160-187) mean_i 1/c_i sum_j E_ij (-log(e_ij / (sum_k e_ik - e_ij))), both directions, / 2.
    The reference evaluates "google" as written (exp, subtract, divide): where P_ii rounds to 1 it returns inf; this
    restatement uses log1p(-P) and stays finite a little longer - compare only where the reference is finite."""
    S = logits(image, text, scale)
    y = np.asarray(labels).reshape(-1)
    E = (y[None, :] == y[:, None]).astype(np.float64)
    c = E.sum(1)
    n = S.shape[0]
    eye = np.eye(n)
    total = 0.0
    for L in (S, S.T):
        if variant == "soft":
            t = _soft_targets(E)
            total += np.mean(t.sum(1) * logsumexp(L, 1) - (t * L).sum(1))
        elif variant == "ignore":
            keep = (E == 0) | (eye == 1)
            total += np.mean(logsumexp(np.where(keep, L, -np.inf), 1) - np.diagonal(L))
        elif variant == "google":
            lse = logsumexp(L, 1)[:, None]
            total += np.mean((E * (lse - L + np.log1p(-np.exp(L - lse)))).sum(1) / c)
        else:
            raise ValueError(variant)
    return 0.5 * total


def labeled_clip_loss_grads(image, text, scale, labels, variant="soft"):
    """Closed-form gradient of labeled_clip_loss (upstream gradient 1):  dI = s dS T, dT = s dS^T I, ds = sum(dS * S) / s.
      soft:    dS = [(P_row + P_col) - E (1/c_i + 1/c_j)] / 2n   (with the reference's float32 targets, _soft_targets)
      ignore:  dS = [(P'_row - 1) + (P'_col - 1)] / 2n on the diagonal, (P'_row + P'_col) / 2n where E = 0, 0 elsewhere
               (P' = softmax over the kept entries)
      google:  dS = [(1 + R_i / c_i) P_row + (1 + R'_j / c_j) P_col - E (1 / (c_i (1 - P_row)) + 1 / (c_j (1 - P_col)))] / 2n,
               R_i = sum_j E_ij P_row_ij / (1 - P_row_ij), R'_j the same over the column."""
    I, T = _f64(image), _f64(text)
    s = float(scale)
    S = s * (I @ T.T)
    y = np.asarray(labels).reshape(-1)
    E = (y[None, :] == y[:, None]).astype(np.float64)
    c = E.sum(1)
    n = S.shape[0]
    eye = np.eye(n)
    w = 1.0 / (2.0 * n)
    if variant == "ignore":
        keep = (E == 0) | (eye == 1)
        Sm = np.where(keep, S, -np.inf)
        Pr = np.exp(Sm - logsumexp(Sm, 1)[:, None])
        Pc = np.exp(Sm - logsumexp(Sm, 0)[None, :])
        dS = w * ((Pr - eye) + (Pc - eye))
    else:
        Pr = np.exp(S - logsumexp(S, 1)[:, None])
        Pc = np.exp(S - logsumexp(S, 0)[None, :])
        if variant == "soft":
            t = _soft_targets(E)
            tr = t.sum(1)
            dS = w * (tr[:, None] * Pr + tr[None, :] * Pc) - w * (t + t.T)
        elif variant == "google":
            R = (E * Pr / (1.0 - Pr)).sum(1)
            Rc = (E * Pc / (1.0 - Pc)).sum(0)
            dS = (w * (1.0 + R / c)[:, None] * Pr + w * (1.0 + Rc / c)[None, :] * Pc
                  - w * E * (1.0 / (c[:, None] * (1.0 - Pr)) + 1.0 / (c[None, :] * (1.0 - Pc))))
        else:
            raise ValueError(variant)
    return s * (dS @ T), s * (dS.T @ I), float(np.sum(dS * (I @ T.T)))


def clip_loss_distributed(image_blocks, text_blocks, scale, rank, local_loss):
    """clip/loss.py:103-114,195-209 for world_size > 1 as seen by `rank`.
    local_loss=False: full [B] vector (identical on every rank); True: the local [b] slice with labels offset by
    num_logits * rank (:200-201)."""
    I_all, T_all = gather_features(image_blocks), gather_features(text_blocks)
    if not local_loss:
        return clip_loss(I_all, T_all, scale)
    b = np.asarray(image_blocks[rank]).shape[0]
    lab = np.arange(b) + b * rank
    return 0.5 * (cross_entropy(image_blocks[rank], T_all, scale, lab) +
                  cross_entropy(text_blocks[rank], I_all, scale, lab))


def clip_loss_distributed_grads(image_blocks, text_blocks, scale, rank, local_loss, gather_with_grad, g):
    """Gradients w.r.t. the LOCAL image/text blocks and the scale for rank `rank` differentiating sum(g * loss) on
    every rank with the same g (the four semantics probed in SURVEY.md section 8c):
      (F, F): rows of the full-batch gradient, ds full          (F, T): world x those rows, ds full
      (T, F): only the paths through the local operands          (T, T): plus reduce-scattered paths through the
                                                                          gathered operands of every rank."""
    W = len(image_blocks)
    b = np.asarray(image_blocks[0]).shape[0]
    I_all, T_all = gather_features(image_blocks), gather_features(text_blocks)
    g = _f64(g)
    if not local_loss:
        dI, dT, ds = clip_loss_grads(I_all, T_all, scale, g)
        mul = float(W) if gather_with_grad else 1.0
        sl = slice(rank * b, (rank + 1) * b)
        return mul * dI[sl], mul * dT[sl], ds
    # local loss: rank r contributes loss_r = 0.5 (CE(I_r T_all^T) + CE(T_r I_all^T)) weighted by g (length b)
    def rank_terms(r):
        lab = np.arange(b) + b * r
        dIr, dTall, ds1 = cross_entropy_grads(image_blocks[r], T_all, scale, lab, 0.5 * g)
        dTr, dIall, ds2 = cross_entropy_grads(text_blocks[r], I_all, scale, lab, 0.5 * g)
        return dIr, dTr, dIall, dTall, ds1 + ds2
    dIr, dTr, dIall, dTall, ds = rank_terms(rank)
    sl = slice(rank * b, (rank + 1) * b)
    if gather_with_grad:
        dI_loc, dT_loc = dIr.copy(), dTr.copy()
        for r in range(W):           # reduce-scatter of every rank's gradient w.r.t. the gathered matrices
            _, _, dIall_r, dTall_r, _ = rank_terms(r)
            dI_loc += dIall_r[sl]
            dT_loc += dTall_r[sl]
        return dI_loc, dT_loc, ds
    return dIr, dTr, ds


def argmax_predictions(image, text):
    """src/models/eval.py:158: logits.argmax(dim=1) (ties -> lowest index, as torch.argmax / np.argmax)."""
    return np.argmax(_f64(image) @ _f64(text).T, axis=1)
