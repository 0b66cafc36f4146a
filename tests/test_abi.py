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


def test_backward_plan_and_workspace_of_the_kept_ds_backward():
    """flyp_clip_backward_plan / flyp_clip_keeps_ds (host logic, no GPU work): which shapes keep dS, and that their
    workspace holds the n_rows x n_cols fp16 matrix."""
    import ctypes
    from flyp_b200 import _lib
    lib = _lib.load()
    bf16, f32 = _lib.FLYP_BF16, _lib.FLYP_F32

    def ws(n, m, d, dt):
        sz = ctypes.c_size_t()
        assert lib.flyp_clip_workspace_bytes(n, m, d, dt, ctypes.byref(sz)) == 0
        return sz.value

    # one rank, bf16, D a multiple of 128, >= 1024 pairs: one sweep + the product over the kept dS
    assert lib.flyp_clip_keeps_ds(32768, 32768, 512, bf16) == 1 and lib.flyp_clip_backward_plan(32768, 32768, 512, bf16) == 1
    assert lib.flyp_clip_keeps_ds(2048, 2048, 1024, bf16) == 1
    # a rank's row block of a row-sharded problem (8 ranks x 4096 rows)
    assert lib.flyp_clip_keeps_ds(4096, 32768, 512, bf16) == 1
    # not kept: small batches, fp32 features, D not a multiple of 128, rows that do not divide the columns
    for n, m, d, dt in ((512, 512, 512, bf16), (4096, 4096, 512, f32), (2048, 2048, 520, bf16), (3000, 7000, 512, bf16)):
        assert lib.flyp_clip_keeps_ds(n, m, d, dt) == 0 and lib.flyp_clip_backward_plan(n, m, d, dt) == 0
    # the workspace of a kept shape holds the fp16 dS matrix (2 bytes per logit) on top of the O(n D) scratch
    assert ws(8192, 8192, 512, bf16) >= 8192 * 8192 * 2
    assert ws(8192, 8192, 520, bf16) < 8192 * 8192 * 2
    assert ws(4096, 32768, 512, bf16) >= 4096 * 32768 * 2
