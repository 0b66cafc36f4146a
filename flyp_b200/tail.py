"""Fused encoder tail: final projection + L2 normalisation of a CLIP tower, the step right before the loss
(clip/model.py:242-243 `x @ self.proj`, :359 `x[eot] @ self.text_projection`, :375-376 `x / x.norm(dim=-1, keepdim=True)`).

    y = project_normalize(pooled, proj)          # == (pooled @ proj) / (pooled @ proj).norm(dim=-1, keepdim=True)

Forward: ONE tcgen05 kernel (flyp_project_normalize_fwd): the [128, N] accumulator tile stays in tensor memory and is
normalised in the epilogue - the un-normalised projection never reaches HBM; with ``out_dtype=torch.bfloat16`` fp32
towers hand the loss bf16 features without a separate cast.  Backward: the fused normalisation backward
(flyp_l2norm_bwd) followed by the two plain GEMMs of a linear layer (library calls - they are not part of the path).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import FlypError


def _code(dt) -> int:
    if dt == torch.bfloat16:
        return _lib.FLYP_BF16
    if dt == torch.float32:
        return _lib.FLYP_F32
    raise FlypError(f"unsupported dtype {dt} (bf16 and fp32 only)")


def project_normalize_fwd(x: torch.Tensor, w: torch.Tensor, out_dtype=None, want_f16: bool = False):
    """Returns (y[n, N], inv_norm[n], y16 or None)."""
    if not (x.is_cuda and w.is_cuda):
        raise FlypError("flyp_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    if x.dim() != 2 or w.dim() != 2 or x.shape[1] != w.shape[0]:
        raise FlypError(f"expected x [n, k] and w [k, N], got {tuple(x.shape)} and {tuple(w.shape)}")
    if x.dtype != w.dtype:
        raise FlypError(f"x and w dtypes differ: {x.dtype} vs {w.dtype}")
    x = x.contiguous(); w = w.contiguous()
    n, k = x.shape
    n_out = w.shape[1]
    dev = x.device
    code = _code(x.dtype)
    out_dtype = out_dtype or x.dtype
    lib = _lib.load()
    with _lib.device_guard(dev):
        sz = ctypes.c_size_t()
        _lib.check(lib.flyp_project_normalize_workspace_bytes(n, k, n_out, code, ctypes.byref(sz)))
        ws = torch.empty(sz.value, dtype=torch.uint8, device=dev)
        y = torch.empty(n, n_out, dtype=out_dtype, device=dev)
        inv = torch.empty(n, dtype=torch.float32, device=dev)
        y16 = torch.empty(n, n_out, dtype=torch.float16, device=dev) if (want_f16 and out_dtype == torch.bfloat16) else None
        _lib.check(lib.flyp_project_normalize_fwd(x.data_ptr(), w.data_ptr(), n, k, n_out, code, y.data_ptr(),
                                                  _code(out_dtype), _lib.ptr(y16), inv.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), _lib.stream_ptr(dev)))
    return y, inv, y16


class _ProjectNormalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, out_dtype):
        y, inv, _ = project_normalize_fwd(x, w, out_dtype)
        ctx.save_for_backward(x, w, y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import ops
        x, w, y, inv = ctx.saved_tensors
        dz = ops.l2norm_bwd(y, dy, inv)                      # (dy - y <y, dy>) / ||z||, in the dtype of y
        dz = dz.to(x.dtype)
        dx = dz @ w.t() if ctx.needs_input_grad[0] else None
        dw = x.t() @ dz if ctx.needs_input_grad[1] else None
        return dx, dw, None


def project_normalize(x: torch.Tensor, proj: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """(x @ proj) / (x @ proj).norm(dim=-1, keepdim=True), fused; differentiable w.r.t. x and proj."""
    return _ProjectNormalize.apply(x, proj, out_dtype)
