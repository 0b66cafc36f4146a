"""The C-ABI library: it builds, loads, exports every symbol include/flyp_clip.h declares, and its argument checking
works.  No compute calls (runs without a GPU)."""
import ctypes
import os
import re

import pytest

from flyp_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "flyp_clip.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flyp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/flyp_clip.h but not exported by libflypclip.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in flyp_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_error_string(lib):
    assert lib.flyp_version() >= 100
    assert isinstance(lib.flyp_last_error(), bytes)


def test_workspace_query_and_argument_errors(lib):
    sz = ctypes.c_size_t()
    assert lib.flyp_clip_workspace_bytes(512, 512, 512, _lib.FLYP_BF16, ctypes.byref(sz)) == 0
    small = sz.value
    assert small > 0
    assert lib.flyp_clip_workspace_bytes(4096, 4096, 768, _lib.FLYP_BF16, ctypes.byref(sz)) == 0
    assert sz.value > small
    assert lib.flyp_ce_workspace_bytes(512, 1000, 512, _lib.FLYP_BF16, ctypes.byref(sz)) == 0
    # dim must be a multiple of 8, sizes positive, dtype known: negative return + message, never an exception/abort
    assert lib.flyp_clip_workspace_bytes(512, 512, 510, _lib.FLYP_BF16, ctypes.byref(sz)) < 0
    assert b"multiple of 8" in lib.flyp_last_error()
    assert lib.flyp_clip_workspace_bytes(0, 512, 512, _lib.FLYP_BF16, ctypes.byref(sz)) < 0
    assert lib.flyp_clip_workspace_bytes(512, 512, 512, 7, ctypes.byref(sz)) < 0
    with pytest.raises(_lib.FlypError):
        _lib.check(lib.flyp_clip_workspace_bytes(512, 512, 512, 7, ctypes.byref(sz)))
    # null pointers are rejected before any CUDA call
    assert lib.flyp_clip_fwd_finish(None, 1, None, 8, 8, 0, None, None, None, None) < 0
    assert lib.flyp_l2norm_fwd(None, 8, 8, 0, None, None, None) < 0


def test_sources_are_sm100a_only():
    mk = open(os.path.join(ROOT, "flyp_b200", "csrc", "Makefile")).read()
    assert "arch=compute_100a,code=sm_100a" in mk
    src = open(os.path.join(ROOT, "flyp_b200", "csrc", "clip_kernels.cu")).read() + \
        open(os.path.join(ROOT, "flyp_b200", "csrc", "sm100.cuh")).read()
    for needle in ("tcgen05.mma", "tcgen05.ld", "cp.async.bulk.tensor", "tcgen05.alloc"):
        assert needle in src
