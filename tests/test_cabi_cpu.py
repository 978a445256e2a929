"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/maskplanner_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from maskplanner_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "maskplanner_b200.h")).read()
    return re.findall(r"MPB_API\s+[\w\s\*]+?\b(mpb_\w+)\s*\(", src)


def test_header_declares_the_hot_path():
    names = set(header_symbols())
    for n in ("mpb_fps_f32", "mpb_ball_query_f32", "mpb_index_points_f32", "mpb_group_points_f32", "mpb_chamfer_nn_f32",
              "mpb_chamfer_nn_bwd_f32", "mpb_padded_lengths_f32", "mpb_knn_group_f32", "mpb_version", "mpb_last_error_string"):
        assert n in names


def test_library_builds_loads_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in header_symbols():
        assert hasattr(lib, name), "library does not export %s" % name
    lib.mpb_version.restype = ctypes.c_int
    assert lib.mpb_version() >= 100


def test_python_binding_covers_exactly_the_header():
    assert set(_cabi.SIGNATURES) == set(header_symbols())


def test_signatures_have_no_torch_types():
    """extern "C", plain pointers and sizes only: nm shows unmangled T symbols for each entry point."""
    out = subprocess.run(["nm", "-D", "--defined-only", build.LIB], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(header_symbols()) <= exported
    assert not [s for s in exported if s.startswith("_Z") and "mpb" in s and " T " in s]


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_invalid_arguments_are_reported_without_a_gpu():
    lib = _cabi.load()
    rc = lib.mpb_ball_query_f32(None, 0, 0, 0, None, 0, 0, 0, 1, 10, 10, 0.04, 4, None, None)
    assert rc == -1 and b"null pointer" in lib.mpb_last_error_string()
    rc = lib.mpb_chamfer_nn_f32(None, None, 1, 4, 4, 65, None, None, None, None, None, None, None)
    assert rc == -1 and b"D > 64" in lib.mpb_last_error_string()
    with pytest.raises(_cabi.MpbError):
        _cabi.check(rc, "mpb_chamfer_nn_f32")
