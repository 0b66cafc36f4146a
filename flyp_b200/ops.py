"""Thin tensor-level wrappers over the C ABI (include/flyp_clip.h).  No math happens in Python: every function
checks its arguments, allocates outputs / workspace through the PyTorch caching allocator and enqueues the CUDA
kernels on the current stream.  There is no CPU path."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import FlypError


def _check_features(a: torch.Tensor, b: torch.Tensor) -> None:
    if not (a.is_cuda and b.is_cuda):
        raise FlypError("flyp_b200 runs on CUDA (sm_100a) only; got a CPU tensor and there is no CPU fallback")
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise FlypError(f"expected two [n, dim] matrices with equal dim, got {tuple(a.shape)} and {tuple(b.shape)}")
    if a.dtype != b.dtype:
        raise FlypError(f"feature dtypes differ: {a.dtype} vs {b.dtype}")
    if a.device != b.device:
        raise FlypError("features live on different devices")
    if a.shape[1] % 8 != 0:
        raise FlypError(f"dim={a.shape[1]} must be a multiple of 8")


def _scale_tensor(scale, device) -> torch.Tensor:
    """logit_scale as a 1-element fp32 device tensor (a longer vector contributes its first element, as
    src/models/flyp_loss_few_shot.py:156-159 does for DataParallel outputs)."""
    if not torch.is_tensor(scale):
        return torch.tensor([float(scale)], dtype=torch.float32, device=device)
    s = scale.detach()
    if s.dtype == torch.float32 and s.device == device and s.numel() == 1:
        return s.view(1)                                    # the usual case: exp() of the fp32 parameter, no copy
    return s.reshape(-1)[:1].to(device=device, dtype=torch.float32).contiguous()


def _f32(n: int, device) -> torch.Tensor:
    return torch.empty(n, dtype=torch.float32, device=device)


def clip_workspace(n_rows: int, n_cols: int, dim: int, dtype_code: int, device) -> torch.Tensor:
    sz = ctypes.c_size_t()
    _lib.check(_lib.load().flyp_clip_workspace_bytes(n_rows, n_cols, dim, dtype_code, ctypes.byref(sz)))
    return torch.empty(sz.value, dtype=torch.uint8, device=device)


_WS_CACHE = {}


def cached_clip_workspace(n_rows: int, n_cols: int, dim: int, dtype_code: int, device) -> torch.Tensor:
    """Kernel scratch, cached per shape, device and stream.  Nothing in it outlives a C call and every use is
    stream-ordered, so consecutive calls (forward and backward, consecutive steps) share it."""
    key = (n_rows, n_cols, dim, dtype_code, device.index, _lib.stream_ptr(device))
    sz = ctypes.c_size_t()
    _lib.check(_lib.load().flyp_clip_workspace_bytes(n_rows, n_cols, dim, dtype_code, ctypes.byref(sz)))
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < sz.value:        # the size also depends on the kernel choice (FLYP_BWD_IMPL)
        if len(_WS_CACHE) > 8:
            _WS_CACHE.clear()
        ws = _WS_CACHE[key] = torch.empty(sz.value, dtype=torch.uint8, device=device)
    return ws


def ce_workspace(n: int, n_classes: int, dim: int, dtype_code: int, device) -> torch.Tensor:
    sz = ctypes.c_size_t()
    _lib.check(_lib.load().flyp_ce_workspace_bytes(n, n_classes, dim, dtype_code, ctypes.byref(sz)))
    return torch.empty(sz.value, dtype=torch.uint8, device=device)


def clip_fwd_local(img: torch.Tensor, txt: torch.Tensor, scale: torch.Tensor, row_offset: int = 0,
                   workspace: Optional[torch.Tensor] = None):
    """Row block statistics of S = scale * img @ txt.T (img: local rows, txt: all rows).
    Returns (row_lse[n_rows], row_nll[n_rows], col_stat[3 * n_cols], status[1])."""
    _check_features(img, txt)
    img = img.contiguous(); txt = txt.contiguous()
    n_rows, dim = img.shape
    n_cols = txt.shape[0]
    dev = img.device
    code = _lib.dtype_code(img)
    with _lib.device_guard(dev):
        ws = workspace if workspace is not None else cached_clip_workspace(n_rows, n_cols, dim, code, dev)
        row_lse, row_nll, col_stat = _f32(n_rows, dev), _f32(n_rows, dev), _f32(3 * n_cols, dev)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        _lib.check(_lib.load().flyp_clip_fwd_local(
            img.data_ptr(), txt.data_ptr(), scale.data_ptr(), n_rows, n_cols, dim, code, row_offset,
            row_lse.data_ptr(), row_nll.data_ptr(), col_stat.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(),
            _lib.stream_ptr(dev)))
    return row_lse, row_nll, col_stat, status


def clip_fwd_finish(col_stat_all: torch.Tensor, world: int, row_nll: torch.Tensor, n_cols: int, row_offset: int = 0,
                    loss_dtype=torch.float32):
    """Merge the column statistics of all ranks.  Returns (col_lse[n_cols], col_nll[n_cols], loss[n_rows]); the loss is
    written by the kernel in ``loss_dtype`` (fp32 or bf16)."""
    dev = row_nll.device
    n_rows = row_nll.numel()
    col_stat_all = col_stat_all.contiguous()
    if col_stat_all.numel() != world * 3 * n_cols:
        raise FlypError("col_stat_all has the wrong size")
    code = {torch.bfloat16: _lib.FLYP_BF16, torch.float32: _lib.FLYP_F32}[loss_dtype]
    with _lib.device_guard(dev):
        col_lse, col_nll = _f32(n_cols, dev), _f32(n_cols, dev)
        loss = torch.empty(n_rows, dtype=loss_dtype, device=dev)
        _lib.check(_lib.load().flyp_clip_fwd_finish_ex(col_stat_all.data_ptr(), world, row_nll.data_ptr(), n_rows, n_cols,
                                                       row_offset, col_lse.data_ptr(), col_nll.data_ptr(),
                                                       loss.data_ptr(), code, None, _lib.stream_ptr(dev)))
    return col_lse, col_nll, loss


def clip_bwd_local(img, txt, scale, row_offset, row_lse, row_nll, col_lse, col_nll, g_row, g_col, grad_mul=1.0,
                   grad_dtype=None, need_img=True, need_txt=True, need_scale=True, workspace=None):
    """Gradients of the row block: d_img[n_rows, dim] (complete), d_txt[n_cols, dim] (this block's share), d_scale[1]."""
    _check_features(img, txt)
    img = img.contiguous(); txt = txt.contiguous()
    n_rows, dim = img.shape
    n_cols = txt.shape[0]
    dev = img.device
    code = _lib.dtype_code(img)
    gdt = img.dtype if grad_dtype is None else grad_dtype
    gcode = {torch.bfloat16: _lib.FLYP_BF16, torch.float32: _lib.FLYP_F32}[gdt]
    need_img = need_img or need_scale
    with _lib.device_guard(dev):
        ws = workspace if workspace is not None else cached_clip_workspace(n_rows, n_cols, dim, code, dev)
        d_img = torch.empty(n_rows, dim, dtype=gdt, device=dev) if need_img else None
        d_txt = torch.empty(n_cols, dim, dtype=gdt, device=dev) if need_txt else None
        d_scale = _f32(1, dev) if need_scale else None
        g_row = g_row.to(torch.float32).contiguous(); g_col = g_col.to(torch.float32).contiguous()
        _lib.check(_lib.load().flyp_clip_bwd_local(
            img.data_ptr(), txt.data_ptr(), scale.data_ptr(), n_rows, n_cols, dim, code, row_offset,
            row_lse.data_ptr(), row_nll.data_ptr(), col_lse.data_ptr(), col_nll.data_ptr(), g_row.data_ptr(),
            g_col.data_ptr(), float(grad_mul), gcode, _lib.ptr(d_img), _lib.ptr(d_txt), _lib.ptr(d_scale),
            ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return d_img, d_txt, d_scale


def ce_fwd(a: torch.Tensor, b: torch.Tensor, scale: torch.Tensor, labels: Optional[torch.Tensor],
           label_offset: int = 0, workspace=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-sample cross-entropy of logits = scale * a @ b.T against integer targets.  Returns (loss[n], lse[n])."""
    _check_features(a, b)
    a = a.contiguous(); b = b.contiguous()
    n, dim = a.shape
    c = b.shape[0]
    dev = a.device
    code = _lib.dtype_code(a)
    if labels is not None:
        if labels.numel() != n:
            raise FlypError("labels must have one entry per row of a")
        labels = labels.to(device=dev, dtype=torch.int64).contiguous()
    with _lib.device_guard(dev):
        ws = workspace if workspace is not None else ce_workspace(n, c, dim, code, dev)
        loss, lse = _f32(n, dev), _f32(n, dev)
        _lib.check(_lib.load().flyp_ce_fwd(a.data_ptr(), b.data_ptr(), scale.data_ptr(), n, c, dim, code,
                                           _lib.ptr(labels), label_offset, loss.data_ptr(), lse.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return loss, lse


def ce_bwd(a, b, scale, labels, label_offset, lse, loss, g, grad_dtype=None, need_a=True, need_b=True,
           need_scale=True, workspace=None):
    _check_features(a, b)
    a = a.contiguous(); b = b.contiguous()
    n, dim = a.shape
    c = b.shape[0]
    dev = a.device
    code = _lib.dtype_code(a)
    gdt = a.dtype if grad_dtype is None else grad_dtype
    gcode = {torch.bfloat16: _lib.FLYP_BF16, torch.float32: _lib.FLYP_F32}[gdt]
    need_a = need_a or need_scale
    if labels is not None:
        labels = labels.to(device=dev, dtype=torch.int64).contiguous()
    with _lib.device_guard(dev):
        ws = workspace if workspace is not None else ce_workspace(n, c, dim, code, dev)
        d_a = torch.empty(n, dim, dtype=gdt, device=dev) if need_a else None
        d_b = torch.empty(c, dim, dtype=gdt, device=dev) if need_b else None
        d_scale = _f32(1, dev) if need_scale else None
        g = g.to(torch.float32).contiguous()
        _lib.check(_lib.load().flyp_ce_bwd(a.data_ptr(), b.data_ptr(), scale.data_ptr(), n, c, dim, code,
                                           _lib.ptr(labels), label_offset, lse.data_ptr(), loss.data_ptr(),
                                           g.data_ptr(), gcode, _lib.ptr(d_a), _lib.ptr(d_b), _lib.ptr(d_scale),
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return d_a, d_b, d_scale


def l2norm_fwd(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    if not x.is_cuda:
        raise FlypError("flyp_b200 runs on CUDA only (no CPU fallback)")
    if x.dim() != 2 or x.shape[1] % 8 != 0:
        raise FlypError(f"expected [n, dim] with dim % 8 == 0, got {tuple(x.shape)}")
    x = x.contiguous()
    n, dim = x.shape
    dev = x.device
    with _lib.device_guard(dev):
        y = torch.empty_like(x)
        inv = _f32(n, dev)
        _lib.check(_lib.load().flyp_l2norm_fwd(x.data_ptr(), n, dim, _lib.dtype_code(x), y.data_ptr(), inv.data_ptr(),
                                               _lib.stream_ptr(dev)))
    return y, inv


def l2norm_bwd(y: torch.Tensor, dy: torch.Tensor, inv: torch.Tensor) -> torch.Tensor:
    y = y.contiguous(); dy = dy.to(y.dtype).contiguous()
    n, dim = y.shape
    dev = y.device
    with _lib.device_guard(dev):
        dx = torch.empty_like(y)
        _lib.check(_lib.load().flyp_l2norm_bwd(y.data_ptr(), dy.data_ptr(), inv.data_ptr(), n, dim, _lib.dtype_code(y),
                                               dx.data_ptr(), _lib.stream_ptr(dev)))
    return dx


def debug_logits(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Raw dot products <a_i, b_j> (fp32 [n_m, n_n]) through the tensor-core path; evaluation / tests only."""
    _check_features(a, b)
    a = a.contiguous(); b = b.contiguous()
    n_m, dim = a.shape
    n_n = b.shape[0]
    dev = a.device
    code = _lib.dtype_code(a)
    with _lib.device_guard(dev):
        ws = clip_workspace(n_m, n_n, dim, code, dev)
        out = torch.empty(n_m, n_n, dtype=torch.float32, device=dev)
        _lib.check(_lib.load().flyp_debug_logits(a.data_ptr(), b.data_ptr(), n_m, n_n, dim, code, out.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return out


def argmax(a: torch.Tensor, b: torch.Tensor, return_max: bool = False):
    """argmax_j <a_i, b_j> per row of a (int64, ties -> lowest index), fused into the tensor-core kernel's epilogue: the
    [n_m, n_n] logits are never materialised."""
    _check_features(a, b)
    a = a.contiguous(); b = b.contiguous()
    n_m, dim = a.shape
    n_n = b.shape[0]
    dev = a.device
    code = _lib.dtype_code(a)
    with _lib.device_guard(dev):
        ws = clip_workspace(n_m, n_n, dim, code, dev)
        idx = torch.empty(n_m, dtype=torch.int64, device=dev)
        mx = _f32(n_m, dev) if return_max else None
        _lib.check(_lib.load().flyp_argmax(a.data_ptr(), b.data_ptr(), n_m, n_n, dim, code, idx.data_ptr(), _lib.ptr(mx),
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)))
    return (idx, mx) if return_max else idx
