"""Whole-step calls of the symmetric loss: ONE C call per direction (flyp_clip_fwd_step / flyp_clip_bwd_step,
include/flyp_clip.h) - what ``ClipLoss.forward`` and its autograd function run (clip/loss.py:94-121,194-211).

``comm=None`` is the single-GPU loss of the FLYP loop (src/models/flyp_loss.py:365,496-499); with a ``PeerComm`` the same
two calls run the row-sharded loss of a rank over NVLink peer memory (flyp_b200/comm.py).  Everything the backward needs
lives in one fp32 buffer and (single GPU, bf16) one fp16 copy of the features written by the forward's preparation pass.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import FlypError, Step


class FusedStep:
    """State of one forward (kept for the backward)."""
    __slots__ = ("comm", "step", "img", "txt", "s", "b", "B", "D", "buf", "col_lse", "col_nll", "ws", "seq", "feat16",
                 "code", "rank", "world")


def _dtype_code(dt) -> int:
    if dt == torch.bfloat16:
        return _lib.FLYP_BF16
    if dt == torch.float32:
        return _lib.FLYP_F32
    raise FlypError(f"unsupported dtype {dt} (bf16 and fp32 only)")


def step_forward(comm, img: torch.Tensor, txt: torch.Tensor, scale: torch.Tensor, loss_dtype, need_backward: bool = True,
                 want_status: bool = False):
    """Returns (loss[B] in loss_dtype, FusedStep[, status]).  ``comm``: PeerComm or None (single rank)."""
    from . import ops
    ops._check_features(img, txt)
    dev = img.device
    code = _lib.dtype_code(img)
    st = FusedStep()
    st.comm, st.img, st.txt, st.s, st.code = comm, img.contiguous(), txt.contiguous(), scale, code
    b, D = img.shape
    if txt.shape[0] != b:
        raise FlypError(f"image and text blocks differ in size: {b} vs {txt.shape[0]}")
    st.rank, st.world = (comm.rank, comm.world) if comm is not None else (0, 1)
    B = b * st.world
    st.b, st.B, st.D = b, B, D
    st.step = Step()
    with _lib.device_guard(dev):
        st.ws = ops.cached_clip_workspace(b, B, D, code, dev)
        bp, Bp = (b + 3) & ~3, (B + 3) & ~3
        # one allocation for every fp32 vector of the step (each slice 16-byte aligned):
        # row_lse[b] row_nll[b] | col_stat[3B] (peer path only) | col_lse[B] col_nll[B] | status
        n_stat = 3 * Bp if comm is not None else 0
        st.buf = buf = torch.empty(2 * bp + n_stat + 2 * Bp + 4, dtype=torch.float32, device=dev)
        loss = torch.empty(B, dtype=loss_dtype, device=dev)
        st.feat16 = None
        if comm is None and need_backward and code == _lib.FLYP_BF16:
            st.feat16 = torch.empty(2 * b * D, dtype=torch.float16, device=dev)
        base, f4 = buf.data_ptr(), 4
        o = 2 * bp
        col_stat = base + o * f4 if comm is not None else None
        st.col_lse, st.col_nll = base + (o + n_stat) * f4, base + (o + n_stat + Bp) * f4
        status = base + (o + n_stat + 2 * Bp) * f4 if want_status else None
        _lib.check(_lib.load().flyp_clip_fwd_step(
            comm._h if comm is not None else None, st.img.data_ptr(), st.txt.data_ptr(), scale.data_ptr(), b, D, code,
            st.rank, st.world, base, base + bp * f4, col_stat, st.col_lse, st.col_nll, loss.data_ptr(),
            _dtype_code(loss_dtype), _lib.ptr(st.feat16), status, st.ws.data_ptr(), st.ws.numel(), ctypes.byref(st.step),
            _lib.stream_ptr(dev)))
    st.seq = 0
    if comm is not None:
        st.seq = comm.seq = int(st.step.gathered.seq)
        comm.check_error()
    if want_status:
        return loss, st, buf[o + n_stat + 2 * Bp:o + n_stat + 2 * Bp + 1].view(torch.int32)
    return loss, st


def step_backward(st: FusedStep, g: torch.Tensor, grad_mul: float, grad_dtype, need_img: bool, need_txt: bool,
                  need_scale: bool):
    """Returns (d_img, d_txt, d_scale): complete gradients of the local rows and the (all-reduced) d(logit_scale)."""
    comm = st.comm
    if comm is not None and not comm.alive(st.seq):
        raise FlypError("the gathered features of this step were overwritten by a later forward: with the peer-memory "
                        "exchange a step's backward must be issued before the next forward of the same ClipLoss")
    dev = st.img.device
    gdt = st.img.dtype if grad_dtype is None else grad_dtype
    gcode = _dtype_code(gdt)
    if st.code == _lib.FLYP_F32:
        if gdt != torch.float32:
            raise FlypError("fp32 features need fp32 gradients")
        g = g.to(torch.float32)
    elif g.dtype not in (torch.float32, torch.bfloat16):
        g = g.to(torch.float32)
    g = g.contiguous()
    if g.numel() != st.B:
        raise FlypError(f"upstream gradient has {g.numel()} entries, the loss vector {st.B}")
    g_code = _lib.FLYP_BF16 if g.dtype == torch.bfloat16 else _lib.FLYP_F32
    need_img = need_img or need_scale
    with _lib.device_guard(dev):
        d_img = torch.empty(st.b, st.D, dtype=gdt, device=dev) if need_img else None
        d_txt = torch.empty(st.b, st.D, dtype=gdt, device=dev) if need_txt else None
        ds = torch.empty(2, dtype=torch.float32, device=dev) if need_scale else None      # [total, partial]
        part = (ds.data_ptr() + 4) if (need_scale and comm is not None) else None
        _lib.check(_lib.load().flyp_clip_bwd_step(
            comm._h if comm is not None else None, ctypes.byref(st.step), st.img.data_ptr(), st.txt.data_ptr(),
            st.s.data_ptr(), st.b, st.D, st.code, st.rank, st.world, st.col_lse, st.col_nll, g.data_ptr(), g_code,
            float(grad_mul), gcode, _lib.ptr(d_img), _lib.ptr(d_txt), part, ds.data_ptr() if need_scale else None,
            st.ws.data_ptr(), st.ws.numel(), _lib.stream_ptr(dev)))
    if comm is not None:
        comm.check_error()
    return d_img, d_txt, (ds[:1] if need_scale else None)
