"""No GPU needed: the C-ABI library builds, loads, exports every symbol include/rightlane.h declares, and its host-side
plan logic (layer schedule, parameter layout, error reporting) agrees with the Python spec."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry
from sim2real_lane_segment_b200 import _lib, spec
from tests.util import layer_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "rightlane.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = re.findall(r"\b(rl_[a-z0-9_]+)\s*\(", hdr)
    assert len(set(names)) >= 18
    for n in set(names):
        assert hasattr(lib, n), f"{n} declared in rightlane.h but not exported"
    assert lib.rl_abi_version() >= 1


def _plan(lib, B, H, W, C, training=1):
    h = ctypes.c_void_p()
    lib.rl_plan_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int]
    rc = lib.rl_plan_create(ctypes.byref(h), B, H, W, C, training)
    return rc, h


def _q(lib, h, what, idx=0):
    lib.rl_plan_query.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
    out = ctypes.c_long()
    assert lib.rl_plan_query(h, what, idx, ctypes.byref(out)) == 0
    return out.value


def test_plan_layout_matches_python_spec(lib):
    for C in (2, 3, 4):
        rc, h = _plan(lib, 2, 120, 160, C)
        assert rc == 0
        lib.rl_plan_num_params.restype = ctypes.c_long
        lib.rl_plan_num_params.argtypes = [ctypes.c_void_p]
        lib.rl_plan_num_bn_buffers.restype = ctypes.c_long
        lib.rl_plan_num_bn_buffers.argtypes = [ctypes.c_void_p]
        pl, bl, counts, npar, nbn = spec.layout(C)
        assert lib.rl_plan_num_params(h) == npar
        assert lib.rl_plan_num_bn_buffers(h) == nbn
        names = layer_names()
        assert _q(lib, h, 0) == len(names) == 66
        for i, nm in enumerate(names):
            key = nm + (".convTrans.weight" if "transUp" in nm else (".weight" if nm.endswith("firstconv") else ".conv.weight"))
            assert _q(lib, h, 11, i) == pl[key][0], (nm, key)
        # level geometry: floor pooling, concat widths
        assert [_q(lib, h, 4, l) for l in range(6)] == [120, 60, 30, 15, 7, 3]
        assert [_q(lib, h, 5, l) for l in range(6)] == [160, 80, 40, 20, 10, 5]
        assert [_q(lib, h, 3, l) for l in range(6)] == [288, 368, 448, 528, 608, 528]
        lib.rl_plan_destroy.argtypes = [ctypes.c_void_p]
        lib.rl_plan_destroy(h)
    assert npar == 3461220


def test_workspace_size_is_dominated_by_the_arenas(lib):
    rc, h = _plan(lib, 64, 120, 160, 4)
    assert rc == 0
    lib.rl_plan_workspace_bytes.restype = ctypes.c_long
    lib.rl_plan_workspace_bytes.argtypes = [ctypes.c_void_p]
    ws = lib.rl_plan_workspace_bytes(h)
    px = [64 * (120 >> l) * (160 >> l) for l in range(6)]
    ct = [288, 368, 448, 528, 608, 528]
    arenas = sum(2 * p * c for p, c in zip(px, ct))
    # activations + gradients + < 2.25 GiB of scratch (dz planes, double-buffered TD/TU scratch, compact x16 / dX16 copies)
    assert 2 * arenas < ws < 2 * arenas + (9 << 28)
    rc2, h2 = _plan(lib, 64, 120, 160, 4, training=0)
    assert lib.rl_plan_workspace_bytes(h2) < ws // 2 + (1 << 29)
    lib.rl_plan_destroy(h)
    lib.rl_plan_destroy(h2)


def test_errors_are_reported_not_thrown(lib):
    lib.rl_last_error.restype = ctypes.c_char_p
    rc, h = _plan(lib, 1, 16, 160, 4)      # too small for five 2x poolings
    assert rc == -2 and b"unsupported shape" in lib.rl_last_error()
    rc, h = _plan(lib, 1, 64, 64, 7)
    assert rc == -2
    rc, h = _plan(lib, 1, 64, 64, 4)
    out = ctypes.c_long()
    lib.rl_plan_query.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_long)]
    assert lib.rl_plan_query(h, 1, 99, ctypes.byref(out)) == -1
    # compute entry points refuse to run on an unbound plan (and there is no CPU fallback behind them)
    lib.rl_forward.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_void_p]
    assert lib.rl_forward(h, None, None, None, 0, None) == -4
    lib.rl_plan_destroy(h)


def test_argument_errors_of_the_stateless_entry_points(lib):
    """null pointers / impossible shapes are rejected with an error code and a message before any device work"""
    lib.rl_last_error.restype = ctypes.c_char_p
    vp = ctypes.c_void_p
    lib.rl_preprocess.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  vp, vp, vp]
    assert lib.rl_preprocess(None, None, 1, 4, 4, 4, 4, 0, None, None, None) == -1
    assert b"rl_preprocess" in lib.rl_last_error()
    assert lib.rl_preprocess(vp(16), None, 1, 4, 4, 4, 4, 0, vp(16), vp(16), None) == -1      # y_out without a label
    assert lib.rl_preprocess(vp(16), None, 0, 4, 4, 4, 4, 0, vp(16), None, None) == -2        # empty batch
    lib.rl_overlay.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp, vp]
    assert lib.rl_overlay(None, 1, 4, 4, 4, 4, None, 0, None, None) == -1
    assert lib.rl_overlay(vp(16), 1, 4, 4, 0, 4, vp(16), 0, vp(16), None) == -2
    rc, h = _plan(lib, 1, 64, 64, 4)
    lib.rl_plan_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_int]
    assert lib.rl_plan_set_option(h, b"no_such_option", 1) == -1
    assert b"no_such_option" in lib.rl_last_error()
    for name in (b"overlap", b"fwd_tma", b"xp", b"bwd_fused"):
        assert lib.rl_plan_set_option(h, name, 1) == 0
    lib.rl_forward.argtypes = [vp, vp, vp, vp, ctypes.c_int, vp]
    assert lib.rl_forward(h, vp(16), vp(16), vp(16), 1, None) != 0                              # plan not bound
    assert b"not bound" in lib.rl_last_error()
    lib.rl_plan_destroy.argtypes = [vp]
    lib.rl_plan_destroy(h)
