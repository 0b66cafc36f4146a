"""Float64 reference of the symmetric contrastive loss for SAMPLED rows of a batch that is too large for the numpy oracle
(B = 32768: the logit matrix has 10^9 entries).  Plain torch float64 on the GPU, independent of every flyp_b200 kernel:
the full row / column logsumexp vectors are accumulated chunk by chunk, then loss, d(image) and d(text) of the sampled
items follow from the closed form (clip/loss.py:117-118,208-209):

    S = s I T^T,  loss_i = ((lse_row_i - S_ii) + (lse_col_i - S_ii)) / 2
    dS_ij = g_i/2 (P^r_ij - d_ij) + g_j/2 (P^c_ij - d_ij),  dI = s dS T,  dT = s dS^T I

Used by tests/ (full-size parity) and by bench.py's in-run check at every world size.  Checker only - never part of the
product path.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def full_lse(I: torch.Tensor, T: torch.Tensor, s: float, chunk: int = 2048):
    """(row_lse[B], col_lse[B]) of S = s I T^T in float64 (natural log), without holding S."""
    I64, T64 = I.double(), T.double()
    B = I64.shape[0]
    row = torch.empty(B, dtype=torch.float64, device=I.device)
    col_m = torch.full((T64.shape[0],), -float("inf"), dtype=torch.float64, device=I.device)
    col_s = torch.zeros(T64.shape[0], dtype=torch.float64, device=I.device)
    for a in range(0, B, chunk):
        S = s * (I64[a:a + chunk] @ T64.T)
        row[a:a + chunk] = torch.logsumexp(S, dim=1)
        m = torch.maximum(col_m, S.max(dim=0).values)
        col_s = col_s * torch.exp(col_m - m) + torch.exp(S - m[None, :]).sum(dim=0)
        col_m = m
    return row, col_m + torch.log(col_s)


@torch.no_grad()
def sampled_reference(I, T, s: float, g, idx, lse=None):
    """loss[idx], dI[idx], dT[idx] (float64) for the global item indices idx; g = upstream gradient on the loss vector.
    I, T: [B, D] on the GPU (any float dtype: evaluated on exactly these values)."""
    I64, T64, g64 = I.double(), T.double(), g.double()
    row_lse, col_lse = lse if lse is not None else full_lse(I, T, s)
    idx = idx.to(I.device)
    k = idx.numel()
    ar = torch.arange(k, device=I.device)
    Sr = s * (I64[idx] @ T64.T)                                  # sampled rows of S      [k, B]
    Sc = s * (I64 @ T64[idx].T)                                  # sampled columns of S   [B, k]
    diag = Sr[ar, idx]
    loss = 0.5 * ((row_lse[idx] - diag) + (col_lse[idx] - diag))
    dSr = 0.5 * g64[idx][:, None] * torch.exp(Sr - row_lse[idx][:, None]) + 0.5 * g64[None, :] * torch.exp(Sr - col_lse[None, :])
    dSr[ar, idx] -= g64[idx]
    dSc = 0.5 * g64[:, None] * torch.exp(Sc - row_lse[:, None]) + 0.5 * g64[idx][None, :] * torch.exp(Sc - col_lse[idx][None, :])
    dSc[idx, ar] -= g64[idx]
    return loss, s * (dSr @ T64), s * (dSc.T @ I64)


def row_errors(got: torch.Tensor, want: torch.Tensor, floor: float = 1e-3):
    """(global, worst_row): max|err| / max|want| over everything, and the worst PER-ROW max|err_r| / max|want_r| over the
    rows whose magnitude is at least floor * max|want| (small rows must be right in relative terms as well)."""
    got, want = got.double(), want.double()
    err = (got - want).abs()
    top = want.abs().max().clamp_min(1e-300)
    glob = (err.max() / top).item()
    if want.dim() == 1:
        return glob, glob
    rmax = want.abs().amax(dim=1)
    live = rmax >= floor * top
    if not bool(live.any()):
        return glob, 0.0
    per_row = (err.amax(dim=1)[live] / rmax[live]).max().item()
    return glob, per_row
