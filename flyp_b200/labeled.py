"""Label-aware ClipLoss variants of clip/loss.py:123-192 (`ground_labels` given): the soft-label loss over items of equal
ground label (:188-192), `ignore=True` (:132-159) and `google_sup_loss=True` (:160-187).  Like the reference they return
a SCALAR (mean over the batch).  No caller in FLYP passes these arguments; they are implemented for drop-in completeness
 on the same tensor-core tiles as the default loss: a class-equality mask in the kernels' epilogues
(flyp_label_stats / flyp_label_sweep, include/flyp_clip.h).  Only O(B) vector arithmetic happens here.

With S = s I T^T, E_ij = [y_i = y_j], c_i = sum_j E_ij, P^r_ij = softmax_j S_ij, P^c_ij = softmax_i S_ij:

  soft labels   L = 1/2 mean_i [lse_row_i - 1/c_i sum_j E_ij S_ij] + 1/2 mean_j [lse_col_j - 1/c_j sum_i E_ij S_ij]
  ignore        L = mean_i 1/2 [(lse'_row_i - S_ii) + (lse'_col_i - S_ii)], lse' over {j: E_ij = 0 or j = i}
  google_sup    L = soft-label L + 1/2 mean_i 1/c_i sum_j E_ij ln(1 - P^r_ij) + 1/2 mean_j 1/c_j sum_i E_ij ln(1 - P^c_ij)

Gradients (dI = s dS T, dT = s dS^T I, ds = sum dS S / s), with w = g / (2B):
  soft labels   dS_ij = w (P^r_ij + P^c_ij) - E_ij w (1/c_i + 1/c_j)
  ignore        dS_ij = w [(P'^r_ij - d_ij) + (P'^c_ij - d_ij)] off the masked entries, 0 on them
  google_sup    dS_ij = w (1 + R_i/c_i) P^r_ij + w (1 + R'_j/c_j) P^c_ij - E_ij w [1/(c_i (1 - P^r_ij)) + 1/(c_j (1 - P^c_ij))]
                with R_i = sum_j E_ij P^r_ij / (1 - P^r_ij), R'_j likewise over the column.
The diagonal entries use exact expressions in the saved cross-entropies (1 - P_ii = -expm1(-nll_i)): no cancellation.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from ._lib import FlypError

SOFT, IGNORE, GOOGLE = "soft", "ignore", "google_sup"


def _stats(a, b, s, cls, mode, lse_rows, ws):
    n, dim = a.shape
    dev = a.device
    out = torch.empty(3, n, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().flyp_label_stats(a.data_ptr(), b.data_ptr(), s.data_ptr(), n, dim, _lib.dtype_code(a),
                                            cls.data_ptr(), cls.data_ptr(), mode, _lib.ptr(lse_rows), out[0].data_ptr(),
                                            out[1].data_ptr(), out[2].data_ptr(), ws.data_ptr(), ws.numel(),
                                            _lib.stream_ptr(dev)))
    return out[0], out[1], out[2]


def _sweep(a, b, s, cls, wr, lr, wc, lc, d_diag, mk_r, mk_c, mode, gmax, gdt, ws, want_ds):
    n, dim = a.shape
    dev = a.device
    out = torch.empty(n, dim, dtype=gdt, device=dev)
    ds = torch.empty(1, dtype=torch.float32, device=dev) if want_ds else None
    gcode = _lib.FLYP_BF16 if gdt == torch.bfloat16 else _lib.FLYP_F32
    _lib.check(_lib.load().flyp_label_sweep(a.data_ptr(), b.data_ptr(), s.data_ptr(), n, dim, _lib.dtype_code(a),
                                            wr.data_ptr(), lr.data_ptr(), wc.data_ptr(), lc.data_ptr(), _lib.ptr(d_diag),
                                            _lib.ptr(cls), _lib.ptr(cls), _lib.ptr(mk_r), _lib.ptr(mk_c), mode,
                                            gmax.data_ptr(), gcode, out.data_ptr(), _lib.ptr(ds), ws.data_ptr(), ws.numel(),
                                            _lib.stream_ptr(dev)))
    return out, ds


class _LabeledLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, scale, labels, variant, grad_dtype):
        ops._check_features(img, txt)
        if img.shape[0] != txt.shape[0] or labels.numel() != img.shape[0]:
            raise FlypError("ground_labels must have one entry per (image, text) pair")
        dev = img.device
        img, txt = img.contiguous(), txt.contiguous()
        n, dim = img.shape
        s = ops._scale_tensor(scale, dev)
        # class ids as dense int32 (index plumbing only) and the class sizes c_i
        _, cls = torch.unique(labels.to(dev).reshape(-1), return_inverse=True)
        cnt = torch.bincount(cls).to(torch.float32)[cls]
        cls = cls.to(torch.int32).contiguous()
        with _lib.device_guard(dev):
            ws = ops.cached_clip_workspace(n, n, dim, _lib.dtype_code(img), dev)
            if variant == IGNORE:
                row_lse, row_nll, _ = _stats(img, txt, s, cls, 1, None, ws)
                col_lse, col_nll, _ = _stats(txt, img, s, cls, 1, None, ws)
                loss = 0.5 * (row_nll + col_nll).mean()
                extra = ()
            else:
                # full softmax statistics (positives handled exactly), then the same-class sums
                row_lse, row_nll, col_stat, _ = ops.clip_fwd_local(img, txt, s, 0, workspace=ws)
                col_lse, col_nll, _ = ops.clip_fwd_finish(col_stat, 1, row_nll, n, 0)
                sx_r, sl_r, sr_r = _stats(img, txt, s, cls, 2, row_lse, ws)
                sx_c, sl_c, sr_c = _stats(txt, img, s, cls, 2, col_lse, ws)
                diag = row_lse - row_nll                                  # S_ii
                loss = 0.5 * ((row_lse - (sx_r + diag) / cnt).mean() + (col_lse - (sx_c + diag) / cnt).mean())
                extra = ()
                if variant == GOOGLE:
                    # diagonal terms from the saved cross-entropies: 1 - P_ii = -expm1(-nll_i)
                    om_r, om_c = -torch.expm1(-row_nll), -torch.expm1(-col_nll)
                    u_r = (sl_r + torch.log(om_r)) / cnt
                    u_c = (sl_c + torch.log(om_c)) / cnt
                    loss = loss + 0.5 * (u_r.mean() + u_c.mean())
                    # R_i, R'_j, and the same sums without the diagonal term (backward: stable diagonal entry)
                    extra = (sr_r + torch.exp(-row_nll) / om_r, sr_c + torch.exp(-col_nll) / om_c, sr_r, sr_c)
        ctx.save_for_backward(img, txt, s, cls, cnt, row_lse, row_nll, col_lse, col_nll, *extra)
        ctx.variant, ctx.grad_dtype = variant, grad_dtype
        ctx.scale_meta = (torch.is_tensor(scale), scale.shape if torch.is_tensor(scale) else None,
                          scale.dtype if torch.is_tensor(scale) else None)
        return loss.to(img.dtype)

    @staticmethod
    def backward(ctx, g):
        img, txt, s, cls, cnt, row_lse, row_nll, col_lse, col_nll, *extra = ctx.saved_tensors
        variant = ctx.variant
        dev = img.device
        n, dim = img.shape
        gdt = img.dtype if ctx.grad_dtype is None else ctx.grad_dtype
        w = (g.float().reshape(()) / (2.0 * n)).expand(n).contiguous()       # upstream scalar gradient / 2B
        p_r, p_c = torch.exp(-row_nll), torch.exp(-col_nll)                  # P^r_ii, P^c_ii
        om_r, om_c = -torch.expm1(-row_nll), -torch.expm1(-col_nll)         # 1 - P_ii
        with _lib.device_guard(dev):
            ws = ops.cached_clip_workspace(n, n, dim, _lib.dtype_code(img), dev)
            if variant == IGNORE:
                wr = wc = w
                mk = None
                d_diag = -w * (om_r + om_c)                                   # w (P'^r_ii - 1) + w (P'^c_ii - 1)
                mode = 1
                bound = 2.0 * w.abs().max()
            elif variant == SOFT:
                wr = wc = w
                mk = (w / cnt).contiguous()
                d_diag = w * (p_r + p_c) - 2.0 * mk
                mode = 2
                bound = 2.0 * w.abs().max() + 2.0 * mk.abs().max()
            else:
                big_r, big_c, off_r, off_c = extra
                wr = (w * (1.0 + big_r / cnt)).contiguous()
                wc = (w * (1.0 + big_c / cnt)).contiguous()
                mk = (w / cnt).contiguous()
                # wr p - mk / (1 - p) with R = R_off + p / (1 - p): the 1 / (1 - p) terms cancel analytically
                # (w p + mk (p R_off - 1 - p)); evaluating them separately loses log10(1 / (1 - p)) digits
                d_diag = w * (p_r + p_c) + mk * (p_r * off_r + p_c * off_c - 2.0 - p_r - p_c)
                mode = 3
                # |dS| <= wr + wc + mk (1 + R_i) + mk (1 + R'_j) (each 1 / (1 - P) of a row is below 1 + its R)
                bound = (wr.abs() + mk.abs() * (1.0 + big_r)).max() + (wc.abs() + mk.abs() * (1.0 + big_c)).max()
            gmax = torch.maximum(bound, d_diag.abs().max()).reshape(1).float().contiguous()
            d_diag = d_diag.contiguous()
            need_s = ctx.needs_input_grad[2]
            d_img, ds = _sweep(img, txt, s, cls, wr, row_lse, wc, col_lse, d_diag, mk, mk, mode, gmax, gdt, ws, need_s)
            d_txt, _ = _sweep(txt, img, s, cls, wc, col_lse, wr, row_lse, d_diag, mk, mk, mode, gmax, gdt, ws, False)
        gs = None
        is_t, shape, dt = ctx.scale_meta
        if need_s and is_t:
            gs = torch.zeros(shape, dtype=torch.float32, device=dev).reshape(-1)
            gs[:1] = ds
            gs = gs.reshape(shape).to(dt)
        return d_img, d_txt, gs, None, None, None


def labeled_clip_loss(image_features, text_features, logit_scale, ground_labels, ignore=False, google_sup_loss=False,
                      grad_dtype=None):
    """clip/loss.py:123-192 (world_size == 1).  Returns the scalar loss in the feature dtype."""
    assert not (ignore and google_sup_loss), 'please specify only one'
    variant = IGNORE if ignore else GOOGLE if google_sup_loss else SOFT
    return _LabeledLoss.apply(image_features, text_features, logit_scale, ground_labels, variant, grad_dtype)
