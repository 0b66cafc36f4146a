"""Host-side logic of the peer-memory communicator that needs no GPU: segment layout sizes, argument errors, the
ctypes structures matching the C header (sizes are asserted inside libflypclip via a self-check entry point-free path:
the Python Structures must have the C sizes or the step entry points would corrupt memory)."""
import ctypes

import pytest

from flyp_b200 import _lib


def _layout(world, rows, dim):
    sz = ctypes.c_size_t()
    rc = _lib.load().flyp_comm_layout_bytes(world, rows, dim, ctypes.byref(sz))
    return rc, sz.value


def test_layout_bytes_scales_with_shape():
    rc, a = _layout(8, 4096, 512)
    assert rc == 0
    # 2 parities x (4 gathered matrices of [world * rows, dim] 16-bit elements + the fp32 reduce-scatter buffer of the
    # text gradient, [world][rows][dim]) dominate; statistics and flags are small
    feat = 2 * (4 * 8 * 4096 * 512 * 2 + 8 * 4096 * 512 * 4)
    assert feat <= a <= feat * 1.05
    rc, b = _layout(8, 4096, 1024)
    assert rc == 0 and b > 1.9 * a - (8 << 20)
    rc, c = _layout(2, 132, 64)
    assert rc == 0 and c % (1 << 20) == 0 and c >= 2 * 4 * 2 * 132 * 64 * 2


@pytest.mark.parametrize("world,rows,dim", [(0, 128, 512), (17, 128, 512), (2, 0, 512), (2, 128, 12), (2, 128, 0)])
def test_layout_rejects_bad_shapes(world, rows, dim):
    rc, _ = _layout(world, rows, dim)
    assert rc == -1
    assert _lib.load().flyp_last_error()


def test_null_communicator_is_an_argument_error():
    lib = _lib.load()
    out = _lib.Gathered()
    assert lib.flyp_comm_gather_features(None, None, None, 128, 512, 0, ctypes.byref(out), None) == -1
    st = _lib.Stats()
    assert lib.flyp_comm_push_stats(None, 1, None, None, None, 128, 256, ctypes.byref(st), None) == -1
    assert lib.flyp_comm_push_scalar(None, 1, None, None) == -1
    assert lib.flyp_comm_set_rs_min_rows(None, 0) == -1
    assert lib.flyp_comm_sum_scalar(None, 1, None, None) == -1
    assert lib.flyp_comm_error(None) == 0
    assert lib.flyp_comm_has_multicast(None) == 0
    assert lib.flyp_comm_destroy(None) == 0
    step = _lib.Step()
    # world 2 without a communicator; and world 1 (communicator-free single-GPU loss) with null features
    assert lib.flyp_clip_fwd_step(None, None, None, None, 128, 512, 0, 0, 2, None, None, None, None, None, None, 0, None,
                                  None, None, 0, ctypes.byref(step), None) == -1
    assert lib.flyp_clip_fwd_step(None, None, None, None, 128, 512, 0, 0, 1, None, None, None, None, None, None, 0, None,
                                  None, None, 0, ctypes.byref(step), None) == -1
    assert lib.flyp_comm_reset_error(None) == 0
    assert lib.flyp_comm_set_timeout_ms(None, 5) == -1


def test_structures_have_the_header_layout():
    # flyp_ready_t: pointer, u32, 5 ints, pointer -> 8 + 4 + 20 (+ pad) + 8; the C side asserts nothing, so pin it here
    assert ctypes.sizeof(_lib.Ready) == 40
    assert ctypes.sizeof(_lib.Gathered) == 4 * 8 + 4 * 40 + 8
    assert ctypes.sizeof(_lib.Stats) == 3 * 8 + 40
    assert ctypes.sizeof(_lib.Step) == ctypes.sizeof(_lib.Gathered) + ctypes.sizeof(_lib.Stats)
