"""ctypes binding of libflypclip.so (C ABI in include/flyp_clip.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing, or a call is made without a CUDA
device, this module raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C flyp_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint32, c_void_p, POINTER, Structure

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflypclip.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

FLYP_BF16 = 0
FLYP_F32 = 1

IPC_HANDLE_BYTES = 64
COMM_MAX_WORLD = 16


class Ready(Structure):
    """flyp_ready_t: device flag words that say which ranks' rows have arrived."""
    _fields_ = [("flags", c_void_p), ("seq", c_uint32), ("n_flags", c_int), ("rows_per_flag", c_int), ("sub", c_int),
                ("stride", c_int), ("timeout_ms", c_uint32), ("err", c_void_p)]


class Gathered(Structure):
    """flyp_gathered_t"""
    _fields_ = [("img_all", c_void_p), ("txt_all", c_void_p), ("img16_all", c_void_p), ("txt16_all", c_void_p),
                ("img_ready", Ready), ("txt_ready", Ready), ("img16_ready", Ready), ("txt16_ready", Ready),
                ("seq", c_uint32)]


class Stats(Structure):
    """flyp_stats_t"""
    _fields_ = [("col_stat_all", c_void_p), ("row_lse_all", c_void_p), ("row_nll_all", c_void_p), ("ready", Ready)]


class Step(Structure):
    """flyp_step_t"""
    _fields_ = [("gathered", Gathered), ("stats", Stats)]


# name -> (restype, argtypes); must list every symbol declared in include/flyp_clip.h
SIGNATURES = {
    "flyp_last_error": (c_char_p, []),
    "flyp_version": (c_int, []),
    "flyp_clip_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "flyp_clip_keeps_ds": (c_int, [c_int, c_int, c_int, c_int]),
    "flyp_clip_backward_plan": (c_int, [c_int, c_int, c_int, c_int]),
    "flyp_clip_fwd_local": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "flyp_clip_fwd_finish": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "flyp_clip_bwd_local": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "flyp_clip_fwd_local_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(Ready), c_void_p]),
    "flyp_clip_fwd_finish_ex": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        c_int, POINTER(Ready), c_void_p]),
    "flyp_clip_bwd_local_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, POINTER(Ready),
                                       POINTER(Ready), c_void_p]),
    "flyp_clip_bwd_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, POINTER(Ready),
                                      POINTER(Ready), POINTER(Ready), POINTER(Ready), c_void_p]),
    "flyp_clip_fwd_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                   c_size_t, POINTER(Step), c_void_p]),
    "flyp_clip_bwd_step": (c_int, [c_void_p, POINTER(Step), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "flyp_clip_bwd_step_phase": (c_int, [c_void_p, POINTER(Step), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                         c_int, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "flyp_comm_create": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "flyp_comm_layout_bytes": (c_int, [c_int, c_int, c_int, POINTER(c_size_t)]),
    "flyp_comm_create_external": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_void_p), c_void_p, POINTER(c_void_p)]),
    "flyp_comm_set_rs_min_rows": (c_int, [c_void_p, c_int]),
    "flyp_comm_has_multicast": (c_int, [c_void_p]),
    "flyp_comm_segment_bytes": (c_int, [c_void_p, POINTER(c_size_t)]),
    "flyp_comm_ipc_handle": (c_int, [c_void_p, c_void_p]),
    "flyp_comm_connect_ipc": (c_int, [c_void_p, c_void_p]),
    "flyp_comm_connect_local": (c_int, [c_void_p, POINTER(c_void_p)]),
    "flyp_comm_error": (c_int, [c_void_p]),
    "flyp_comm_reset_error": (c_int, [c_void_p]),
    "flyp_comm_set_timeout_ms": (c_int, [c_void_p, c_uint32]),
    "flyp_comm_destroy": (c_int, [c_void_p]),
    "flyp_comm_gather_features": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(Gathered),
                                          c_void_p]),
    "flyp_comm_push_stats": (c_int, [c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_int, c_int, POINTER(Stats),
                                     c_void_p]),
    "flyp_comm_push_scalar": (c_int, [c_void_p, c_uint32, c_void_p, c_void_p]),
    "flyp_comm_sum_scalar": (c_int, [c_void_p, c_uint32, c_void_p, c_void_p]),
    "flyp_ce_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "flyp_ce_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                            c_void_p, c_void_p, c_size_t, c_void_p]),
    "flyp_ce_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                            c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "flyp_ce_fwd_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_size_t, POINTER(Ready), c_void_p]),
    "flyp_ce_bwd_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
                               POINTER(Ready), POINTER(Ready), c_void_p]),
    "flyp_l2norm_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "flyp_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "flyp_label_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "flyp_label_sweep": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "flyp_project_normalize_workspace_bytes": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "flyp_project_normalize_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                           c_void_p, c_void_p, c_size_t, c_void_p]),
    "flyp_argmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                            c_void_p]),
    "flyp_debug_profile": (c_int, [c_void_p]),
    "flyp_debug_kernel_events": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "flyp_debug_logits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
}

_lib = None


class FlypError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libflypclip.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC_DIR], capture_output=True, text=True)
    if res.returncode != 0:
        raise FlypError("building libflypclip.so failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library (never falls back to anything else)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlypError(f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()); "
                        "flyp_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().flyp_last_error()
        raise FlypError(f"libflypclip error {rc}: {msg.decode() if msg else '?'}")


def dtype_code(t) -> int:
    import torch
    if t.dtype == torch.bfloat16:
        return FLYP_BF16
    if t.dtype == torch.float32:
        return FLYP_F32
    raise FlypError(f"unsupported feature dtype {t.dtype} (bf16 and fp32 only)")


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())
    except AttributeError:  # pragma: no cover - older torch
        return torch.cuda.current_stream(device).cuda_stream


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """``torch.cuda.device(device)`` only when ``device`` is not already current (the context manager costs ~5 us)."""
    import torch
    if device.index is None or torch.cuda.current_device() == device.index:
        return _NO_GUARD
    return torch.cuda.device(device)
