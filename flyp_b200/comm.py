"""Peer-memory exchange between the GPUs of one node - the host side of ``flyp_comm_*`` (include/flyp_clip.h).

For the row-sharded loss this replaces the collectives of the reference's ``gather_features`` (clip/loss.py:19-69) and
the statistics exchange: every rank owns an exchange segment that all peers map (CUDA IPC); features are pushed with the
copy engines in ring order while the forward kernel already works on the local column block, and the tensor-core
kernels poll per-rank flag words before they touch a peer's rows.  ``torch.distributed`` is used only once, to hand the
64-byte IPC handles around.
"""
from __future__ import annotations

import ctypes
import socket
from typing import Optional

import torch

from . import _lib
from ._lib import FlypError, Gathered, Ready, Stats, Step


class PeerComm:
    """One rank's communicator: segment sized for blocks of up to ``max_rows x dim`` bf16 features."""

    def __init__(self, rank: int, world: int, max_rows: int, dim: int, device: torch.device, segments=None,
                 multicast: int = 0, keepalive=None):
        """Without ``segments``: the library allocates the segment (wire it with connect_ipc / connect_local).
        With ``segments`` (addresses of every rank's zeroed segment, own one included) and optionally the address of a
        multicast mapping of them: nothing is allocated; ``keepalive`` holds whatever owns that memory."""
        self.rank, self.world, self.max_rows, self.dim, self.device = rank, world, max_rows, dim, device
        self._h = ctypes.c_void_p()
        self._keepalive = keepalive
        lib = _lib.load()
        with _lib.device_guard(device):
            if segments is None:
                _lib.check(lib.flyp_comm_create(rank, world, max_rows, dim, ctypes.byref(self._h)))
            else:
                arr = (ctypes.c_void_p * world)(*[int(p) for p in segments])
                _lib.check(lib.flyp_comm_create_external(rank, world, max_rows, dim, arr, int(multicast) or None,
                                                         ctypes.byref(self._h)))
        self.seq = 0                      # sequence number of the latest gather
        self.multicast = bool(lib.flyp_comm_has_multicast(self._h))

    # ---- wiring -----------------------------------------------------------------------------------------------
    @staticmethod
    def layout_bytes(world: int, max_rows: int, dim: int) -> int:
        sz = ctypes.c_size_t()
        _lib.check(_lib.load().flyp_comm_layout_bytes(world, max_rows, dim, ctypes.byref(sz)))
        return sz.value

    def ipc_handle(self) -> bytes:
        buf = ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(_lib.load().flyp_comm_ipc_handle(self._h, buf))
        return buf.raw

    def connect_ipc(self, handles) -> None:
        blob = b"".join(handles)
        if len(blob) != self.world * _lib.IPC_HANDLE_BYTES:
            raise FlypError("expected one IPC handle per rank")
        with _lib.device_guard(self.device):
            _lib.check(_lib.load().flyp_comm_connect_ipc(self._h, blob))

    @staticmethod
    def connect_local(comms) -> None:
        """Wire communicators that live in this process (tests: several emulated ranks on one GPU)."""
        arr = (ctypes.c_void_p * len(comms))(*[c._h for c in comms])
        for c in comms:
            _lib.check(_lib.load().flyp_comm_connect_local(c._h, arr))

    @classmethod
    def _from_symmetric_memory(cls, rank, world, max_rows, dim, device, group):
        """Segments from torch.distributed._symmetric_memory: peer-mapped by torch, with an NVSwitch multicast mapping
        when the fabric offers one (then every push is one copy that the switch replicates)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        nbytes = cls.layout_bytes(world, max_rows, dim)
        with _lib.device_guard(device):
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
        off = int(getattr(hdl, "offset", 0) or 0)          # position of `buf` inside the rendezvoused allocation
        ptrs = [int(p) + off for p in hdl.buffer_ptrs]
        if len(ptrs) != world or ptrs[rank] != buf.data_ptr():
            raise FlypError("unexpected symmetric-memory layout")
        mc = int(hdl.multicast_ptr or 0)
        return cls(rank, world, max_rows, dim, device, segments=ptrs, multicast=(mc + off) if mc else 0,
                   keepalive=(buf, hdl))

    @classmethod
    def from_process_group(cls, rank: int, world: int, max_rows: int, dim: int, device: torch.device,
                           group=None) -> Optional["PeerComm"]:
        """Create + connect over ``torch.distributed``: symmetric memory (multicast) first, CUDA IPC segments second.
        Returns None (on every rank alike) when the ranks do not all sit on one node or no segment could be mapped;
        the caller then uses the NCCL path."""
        import os
        import torch.distributed as dist

        def agree(ok: bool) -> bool:
            flags = [None] * world
            dist.all_gather_object(flags, bool(ok), group=group)
            return all(flags)

        hosts = [None] * world
        dist.all_gather_object(hosts, socket.gethostname(), group=group)
        if len(set(hosts)) != 1:
            return None
        comm = None
        if os.environ.get("FLYP_COMM_SYMM", "1") != "0":
            try:
                comm = cls._from_symmetric_memory(rank, world, max_rows, dim, device, group)
            except Exception:  # noqa: BLE001 - any failure means "no symmetric memory on this rank"
                comm = None
            if agree(comm is not None):
                torch.cuda.synchronize(device)
                dist.barrier(group=group)      # every segment is zeroed and mapped before anyone pushes
                return comm
            if comm is not None:
                comm.close()
        comm, handle = None, b""
        try:
            comm = cls(rank, world, max_rows, dim, device)
            handle = comm.ipc_handle()
        except Exception:  # noqa: BLE001
            comm = None
        infos = [None] * world
        dist.all_gather_object(infos, handle, group=group)
        ok = all(infos)
        if ok:
            try:
                comm.connect_ipc(infos)
            except Exception:  # noqa: BLE001
                ok = False
        if not agree(ok):
            if comm is not None:
                comm.close()
            return None
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        return comm

    def close(self) -> None:
        if self._h:
            _lib.load().flyp_comm_destroy(self._h)
            self._h = ctypes.c_void_p()
        self._keepalive = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    # ---- steps ------------------------------------------------------------------------------------------------
    def check_error(self) -> None:
        e = _lib.load().flyp_comm_error(self._h)
        if e:
            raise FlypError(f"a kernel gave up waiting for the rows of rank {e - 1} (peer-memory exchange) and trapped; "
                            "the CUDA context of this process is lost")

    def alive(self, seq: int) -> bool:
        """A step's gathered buffers may be consumed until the next gather is issued.  (The segments are double
        buffered by step parity so that a FAST rank's next push cannot overwrite what a slow rank still reads: a rank
        receives the rows of step s + 1 only after their owner has finished - in stream order - everything it issued for
        step s.  That argument needs every rank to issue the backward of step s before its forward of step s + 1, so an
        older step is refused rather than raced with.)"""
        return self.seq == seq

    def reset_error(self) -> None:
        _lib.load().flyp_comm_reset_error(self._h)

    def set_rs_min_rows(self, rows: int) -> None:
        """Rows per rank from which the text gradient goes through the kept-dS product + NVLink reduce-scatter (default
        6144; 0 = whenever the shape keeps dS).  Must be the same on every rank."""
        _lib.check(_lib.load().flyp_comm_set_rs_min_rows(self._h, int(rows)))

    def set_timeout_ms(self, ms: int) -> None:
        _lib.check(_lib.load().flyp_comm_set_timeout_ms(self._h, int(ms)))

    def gather(self, img: torch.Tensor, txt: torch.Tensor) -> Gathered:
        n, d = img.shape
        out = Gathered()
        with _lib.device_guard(self.device):
            _lib.check(_lib.load().flyp_comm_gather_features(self._h, img.data_ptr(), txt.data_ptr(), n, d,
                                                             _lib.dtype_code(img), ctypes.byref(out),
                                                             _lib.stream_ptr(self.device)))
        self.seq = int(out.seq)
        return out

    def push_stats(self, seq: int, col_stat: torch.Tensor, row_lse: torch.Tensor, row_nll: torch.Tensor) -> Stats:
        out = Stats()
        n_rows = row_lse.numel()
        with _lib.device_guard(self.device):
            _lib.check(_lib.load().flyp_comm_push_stats(self._h, seq, col_stat.data_ptr(), row_lse.data_ptr(),
                                                        row_nll.data_ptr(), n_rows, col_stat.numel() // 3,
                                                        ctypes.byref(out), _lib.stream_ptr(self.device)))
        return out

    def all_reduce_scalar(self, seq: int, value: torch.Tensor, out: torch.Tensor) -> None:
        self.push_scalar(seq, value)
        self.sum_scalar(seq, out)

    def push_scalar(self, seq: int, value: torch.Tensor) -> None:
        with _lib.device_guard(self.device):
            _lib.check(_lib.load().flyp_comm_push_scalar(self._h, seq, value.data_ptr(), _lib.stream_ptr(self.device)))

    def sum_scalar(self, seq: int, out: torch.Tensor) -> None:
        with _lib.device_guard(self.device):
            _lib.check(_lib.load().flyp_comm_sum_scalar(self._h, seq, out.data_ptr(), _lib.stream_ptr(self.device)))


# ---------------------------------------------------------------------------------------------------- loss phases
# The row-sharded symmetric loss over a communicator, split into the phases between which ranks exchange data; the
# autograd function in loss.py runs them back to back, the single-GPU emulation test runs each phase for all ranks.

class PeerStep:
    """State of one forward (kept for the backward)."""
    __slots__ = ("comm", "g", "st", "img", "txt", "s", "b", "B", "D", "off", "buf", "row_lse", "row_nll", "col_stat",
                 "col_lse", "col_nll", "loss", "ws")


def _workspace(b: int, B: int, D: int, dev, code: int = _lib.FLYP_BF16) -> torch.Tensor:
    from . import ops
    return ops.cached_clip_workspace(b, B, D, code, dev)


def fwd_gather(comm: PeerComm, img: torch.Tensor, txt: torch.Tensor, scale: torch.Tensor) -> PeerStep:
    from . import ops
    ops._check_features(img, txt)
    st = PeerStep()
    st.comm, st.img, st.txt, st.s = comm, img.contiguous(), txt.contiguous(), scale
    st.b, st.D = img.shape
    st.B, st.off = st.b * comm.world, comm.rank * st.b
    st.g = comm.gather(st.img, st.txt)
    return st


def fwd_local(st: PeerStep) -> None:
    dev = st.img.device
    lib = _lib.load()
    b, B = st.b, st.B
    with _lib.device_guard(dev):
        st.ws = _workspace(b, B, st.D, dev, _lib.dtype_code(st.img))
        # one allocation for every fp32 vector of the step (each slice 16-byte aligned)
        bp = (b + 3) & ~3
        Bp = (B + 3) & ~3
        st.buf = torch.empty(2 * bp + 5 * Bp, dtype=torch.float32, device=dev)
        st.row_lse, st.row_nll = st.buf[0:b], st.buf[bp:bp + b]
        o = 2 * bp
        st.col_stat = st.buf[o:o + 3 * B]
        st.col_lse, st.col_nll = st.buf[o + 3 * Bp:o + 3 * Bp + B], st.buf[o + 4 * Bp:o + 4 * Bp + B]
        _lib.check(lib.flyp_clip_fwd_local_ex(
            st.img.data_ptr(), st.g.txt_all, st.s.data_ptr(), b, B, st.D, _lib.dtype_code(st.img), st.off,
            st.row_lse.data_ptr(), st.row_nll.data_ptr(), st.col_stat.data_ptr(), None,
            st.ws.data_ptr(), st.ws.numel(), ctypes.byref(st.g.txt_ready), _lib.stream_ptr(dev)))
    st.st = st.comm.push_stats(st.g.seq, st.col_stat, st.row_lse, st.row_nll)


def fwd_finish(st: PeerStep, loss_dtype=torch.float32) -> torch.Tensor:
    """Loss of every global row (the reference returns the full vector on every rank, clip/loss.py:113-114,208)."""
    dev = st.img.device
    code = {torch.bfloat16: _lib.FLYP_BF16, torch.float32: _lib.FLYP_F32}[loss_dtype]
    with _lib.device_guard(dev):
        st.loss = torch.empty(st.B, dtype=loss_dtype, device=dev)
        _lib.check(_lib.load().flyp_clip_fwd_finish_ex(
            st.st.col_stat_all, st.comm.world, st.st.row_nll_all, st.B, st.B, 0, st.col_lse.data_ptr(),
            st.col_nll.data_ptr(), st.loss.data_ptr(), code, ctypes.byref(st.st.ready), _lib.stream_ptr(dev)))
    return st.loss


def bwd_local(st: PeerStep, g: torch.Tensor, grad_mul: float, grad_dtype, need_img: bool, need_txt: bool,
              need_scale: bool):
    """Returns (d_img, d_txt, d_scale_partial): complete gradients of the local rows, this rank's share of d(scale)."""
    if not st.comm.alive(st.g.seq):
        raise FlypError("the gathered features of this step were overwritten by a later forward: with the peer-memory "
                        "exchange a step's backward must be issued before the next forward")
    dev = st.img.device
    lib = _lib.load()
    gdt = st.img.dtype if grad_dtype is None else grad_dtype
    gcode = {torch.bfloat16: _lib.FLYP_BF16, torch.float32: _lib.FLYP_F32}[gdt]
    if g.dtype not in (torch.float32, torch.bfloat16):
        g = g.to(torch.float32)
    g = g.contiguous()
    g_code = _lib.FLYP_BF16 if g.dtype == torch.bfloat16 else _lib.FLYP_F32
    need_img = need_img or need_scale
    with _lib.device_guard(dev):
        d_img = torch.empty(st.b, st.D, dtype=gdt, device=dev) if need_img else None
        d_txt = torch.empty(st.b, st.D, dtype=gdt, device=dev) if need_txt else None
        d_s = torch.empty(1, dtype=torch.float32, device=dev) if need_scale else None
        gg = st.g
        _lib.check(lib.flyp_clip_bwd_sharded(
            st.img.data_ptr(), st.txt.data_ptr(), gg.img_all, gg.txt_all, gg.img16_all, gg.txt16_all, st.s.data_ptr(),
            st.b, st.B, st.D, _lib.dtype_code(st.img), st.off, st.st.row_lse_all, st.st.row_nll_all, st.col_lse.data_ptr(),
            st.col_nll.data_ptr(), g.data_ptr(), g_code, float(grad_mul), gcode, _lib.ptr(d_img), _lib.ptr(d_txt),
            _lib.ptr(d_s), st.ws.data_ptr(), st.ws.numel(), ctypes.byref(gg.img_ready), ctypes.byref(gg.txt_ready),
            ctypes.byref(gg.img16_ready), ctypes.byref(gg.txt16_ready), _lib.stream_ptr(dev)))
    return d_img, d_txt, d_s


def bwd_step_phase(st: PeerStep, g: torch.Tensor, grad_mul: float, grad_dtype, phase: int, outs=None):
    """flyp_clip_bwd_step in two phases over the state of the phase-wise forward above (what ``ClipLoss.backward`` runs
    as ONE call, flyp_b200/step.py): phase 1 enqueues what the rank computes and publishes - the first sweep and, with
    the kept-dS backward, the product dS^T . image whose partials go straight into the owners' reduce-scatter buffers -
    phase 2 what waits for the other ranks (sum of the W partials of d_txt, sum of the d(scale) partials).
    Phase 1 returns ``outs`` = (d_img, d_txt, ds[2] = [total, partial]) to pass to phase 2."""
    if not st.comm.alive(st.g.seq):
        raise FlypError("the gathered features of this step were overwritten by a later forward")
    dev = st.img.device
    gdt = st.img.dtype if grad_dtype is None else grad_dtype
    gcode = {torch.bfloat16: _lib.FLYP_BF16, torch.float32: _lib.FLYP_F32}[gdt]
    if g.dtype not in (torch.float32, torch.bfloat16):
        g = g.to(torch.float32)
    g = g.contiguous()
    g_code = _lib.FLYP_BF16 if g.dtype == torch.bfloat16 else _lib.FLYP_F32
    with _lib.device_guard(dev):
        if outs is None:
            outs = (torch.empty(st.b, st.D, dtype=gdt, device=dev), torch.empty(st.b, st.D, dtype=gdt, device=dev),
                    torch.empty(2, dtype=torch.float32, device=dev))
        d_img, d_txt, ds = outs
        step = _lib.Step()
        step.gathered, step.stats = st.g, st.st
        _lib.check(_lib.load().flyp_clip_bwd_step_phase(
            st.comm._h, ctypes.byref(step), st.img.data_ptr(), st.txt.data_ptr(), st.s.data_ptr(), st.b, st.D,
            _lib.dtype_code(st.img), st.comm.rank, st.comm.world, st.col_lse.data_ptr(), st.col_nll.data_ptr(),
            g.data_ptr(), g_code, float(grad_mul), gcode, d_img.data_ptr(), d_txt.data_ptr(), ds.data_ptr() + 4,
            ds.data_ptr(), st.ws.data_ptr(), st.ws.numel(), int(phase), _lib.stream_ptr(dev)))
    return outs


# ---------------------------------------------------------------------------------------------------- local_loss
# clip/loss.py:109-111,200-201: with local_loss=True a rank's loss is two one-directional cross-entropies of its OWN rows
# against the gathered matrices.  The gather is the peer-memory exchange above; the class matrix of each block lives in
# this rank's segment and the kernels wait for its rows as they arrive.

class LocalStep:
    __slots__ = ("comm", "g", "img", "txt", "s", "b", "B", "D", "off", "seq", "ws_i", "ws_t", "lse_i", "lse_t", "loss_i",
                 "loss_t", "code")


def _ce_ws(n, c, d, code, dev):
    from . import ops
    return ops.ce_workspace(n, c, d, code, dev)


def local_fwd(comm: PeerComm, img: torch.Tensor, txt: torch.Tensor, scale: torch.Tensor):
    """Returns (loss_image[b], loss_text[b]) in fp32 and the state for local_bwd."""
    return local_compute(local_gather(comm, img, txt, scale))


def local_gather(comm: PeerComm, img: torch.Tensor, txt: torch.Tensor, scale: torch.Tensor) -> "LocalStep":
    """Phase 1 (pack + pushes).  Split from the compute phase so that several emulated ranks on ONE GPU can be stepped
    phase by phase (a kernel must never wait for a kernel that has not been enqueued)."""
    from . import ops
    ops._check_features(img, txt)
    dev = img.device
    st = LocalStep()
    st.comm, st.img, st.txt, st.s = comm, img.contiguous(), txt.contiguous(), scale
    st.b, st.D = img.shape
    st.B, st.off = st.b * comm.world, comm.rank * st.b
    st.code = _lib.dtype_code(img)
    st.g = comm.gather(st.img, st.txt)
    st.seq = int(st.g.seq)
    return st


def local_compute(st: "LocalStep"):
    """Phase 2: the two cross-entropy blocks over the gathered matrices (their kernels wait for the rows as they arrive)."""
    comm, scale, dev = st.comm, st.s, st.img.device
    lib = _lib.load()
    with _lib.device_guard(dev):
        st.ws_i = _ce_ws(st.b, st.B, st.D, st.code, dev)
        st.ws_t = _ce_ws(st.b, st.B, st.D, st.code, dev)
        buf = torch.empty(4, st.b, dtype=torch.float32, device=dev)
        st.loss_i, st.lse_i, st.loss_t, st.lse_t = buf[0], buf[1], buf[2], buf[3]
        stream = _lib.stream_ptr(dev)
        _lib.check(lib.flyp_ce_fwd_ex(st.img.data_ptr(), st.g.txt_all, scale.data_ptr(), st.b, st.B, st.D, st.code, None,
                                      st.off, st.loss_i.data_ptr(), st.lse_i.data_ptr(), st.ws_i.data_ptr(),
                                      st.ws_i.numel(), ctypes.byref(st.g.txt_ready), stream))
        _lib.check(lib.flyp_ce_fwd_ex(st.txt.data_ptr(), st.g.img_all, scale.data_ptr(), st.b, st.B, st.D, st.code, None,
                                      st.off, st.loss_t.data_ptr(), st.lse_t.data_ptr(), st.ws_t.data_ptr(),
                                      st.ws_t.numel(), ctypes.byref(st.g.img_ready), stream))
    comm.check_error()
    return st.loss_i, st.loss_t, st


def local_bwd(st: LocalStep, g: torch.Tensor, grad_dtype, need_gathered: bool):
    """g[b]: upstream gradient on (loss_image + loss_text) / 2.  Returns (d_img[b, D], d_txt[b, D], d_scale[1]) of the
    local operands and, with need_gathered (gather_with_grad=True), the gradients w.r.t. the GATHERED matrices
    (d_img_all[B, D], d_txt_all[B, D]) that the caller reduce-scatters."""
    comm = st.comm
    if not comm.alive(st.seq):
        raise FlypError("the gathered features of this step were overwritten by a later forward: with the peer-memory "
                        "exchange a step's backward must be issued before the next forward")
    dev = st.img.device
    gdt = st.img.dtype if grad_dtype is None else grad_dtype
    gcode = _lib.FLYP_BF16 if gdt == torch.bfloat16 else _lib.FLYP_F32
    lib = _lib.load()
    half = (0.5 * g.to(torch.float32)).contiguous()
    with _lib.device_guard(dev):
        d_img = torch.empty(st.b, st.D, dtype=gdt, device=dev)
        d_txt = torch.empty(st.b, st.D, dtype=gdt, device=dev)
        d_txt_all = torch.empty(st.B, st.D, dtype=gdt, device=dev) if need_gathered else None
        d_img_all = torch.empty(st.B, st.D, dtype=gdt, device=dev) if need_gathered else None
        ds = torch.empty(2, dtype=torch.float32, device=dev)
        stream = _lib.stream_ptr(dev)
        gg = st.g
        bf16 = st.code == _lib.FLYP_BF16
        _lib.check(lib.flyp_ce_bwd_ex(st.img.data_ptr(), gg.txt_all, st.s.data_ptr(), st.b, st.B, st.D, st.code, None, st.off,
                                      st.lse_i.data_ptr(), st.loss_i.data_ptr(), half.data_ptr(), gcode, d_img.data_ptr(),
                                      _lib.ptr(d_txt_all), ds.data_ptr(), st.ws_i.data_ptr(), st.ws_i.numel(),
                                      gg.txt16_all if bf16 else None, ctypes.byref(gg.txt_ready),
                                      ctypes.byref(gg.txt16_ready), stream))
        _lib.check(lib.flyp_ce_bwd_ex(st.txt.data_ptr(), gg.img_all, st.s.data_ptr(), st.b, st.B, st.D, st.code, None, st.off,
                                      st.lse_t.data_ptr(), st.loss_t.data_ptr(), half.data_ptr(), gcode, d_txt.data_ptr(),
                                      _lib.ptr(d_img_all), ds.data_ptr() + 4, st.ws_t.data_ptr(), st.ws_t.numel(),
                                      gg.img16_all if bf16 else None, ctypes.byref(gg.img_ready),
                                      ctypes.byref(gg.img16_ready), stream))
    comm.check_error()
    return d_img, d_txt, ds.sum().reshape(1), d_img_all, d_txt_all


# ---------------------------------------------------------------------------------------------------- whole steps
# What the drop-in module calls: one C call per direction (flyp_b200/step.py), shared with the single-GPU loss.
from .step import FusedStep, step_backward, step_forward  # noqa: E402,F401
