"""ctypes view of oracle/c/mp_oracle.c -- TEST INFRASTRUCTURE (see oracle/__init__.py).

All functions take and return numpy arrays / CPU torch tensors converted to numpy; nothing here
touches CUDA.  ``build()`` compiles the library with the Makefile next to the C source.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmp_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/c/mp_oracle.c (gcc, -ffp-contract=off) into oracle/_build/."""
    src = os.path.join(_HERE, "c", "mp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def set_threads(n):
    """OpenMP threads used by the C loops (returns the value in effect)."""
    return int(lib().orc_set_threads(ctypes.c_int(int(n))))


def _np(a, dtype):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fps(xyz, npoint, seed):
    """models/pointnet2_utils.py:65-86 with the seed indices (:77) supplied by the caller."""
    xyz = _np(xyz, np.float32)
    seed = _np(seed, np.int64)
    B, N, _ = xyz.shape
    out = np.zeros((B, npoint), dtype=np.int64)
    lib().orc_fps_f32(_p(xyz), ctypes.c_int64(N * 3), ctypes.c_int64(3), ctypes.c_int64(1),
                      ctypes.c_int(B), ctypes.c_int(N), _p(seed), ctypes.c_int(npoint), _p(out))
    return out


def square_distance(src, dst):
    """models/pointnet2_utils.py:21-42 (3-d points)."""
    src, dst = _np(src, np.float32), _np(dst, np.float32)
    B, N, _ = src.shape
    M = dst.shape[1]
    out = np.empty((B, N, M), dtype=np.float32)
    lib().orc_square_distance_f32(_p(src), _p(dst), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(M), _p(out))
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    """models/pointnet2_utils.py:89-109; the threshold is fp32(radius**2) as torch casts it."""
    xyz, new_xyz = _np(xyz, np.float32), _np(new_xyz, np.float32)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, nsample), dtype=np.int64)
    r2 = np.float32(radius ** 2)
    lib().orc_ball_query_f32(_p(xyz), _p(new_xyz), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(S),
                             ctypes.c_float(float(r2)), ctypes.c_int(nsample), _p(out))
    return out


def knn(p1, p2, len1=None, len2=None, K=1, use_fma=True):
    """pytorch3d knn_points semantics (see mp_oracle.c).  Returns (dists [N,P1,K], idx [N,P1,K])."""
    p1, p2 = _np(p1, np.float32), _np(p2, np.float32)
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    l1 = _np(len1, np.int64) if len1 is not None else None
    l2 = _np(len2, np.int64) if len2 is not None else None
    d = np.empty((N, P1, K), dtype=np.float32)
    i = np.empty((N, P1, K), dtype=np.int64)
    lib().orc_knn_f32(_p(p1), _p(p2), ctypes.c_int(N), ctypes.c_int(P1), ctypes.c_int(P2), ctypes.c_int(D),
                      _p(l1) if l1 is not None else None, _p(l2) if l2 is not None else None,
                      ctypes.c_int(K), ctypes.c_int(1 if use_fma else 0), _p(d), _p(i))
    return d, i


def knn_group(xyz, new_xyz, K):
    """Stress-config kNN grouping oracle: expanded-form distances + k smallest (lowest index on ties)."""
    xyz, new_xyz = _np(xyz, np.float32), _np(new_xyz, np.float32)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    idx = np.empty((B, S, K), dtype=np.int64)
    d = np.empty((B, S, K), dtype=np.float32)
    lib().orc_knn_group_f32(_p(xyz), _p(new_xyz), ctypes.c_int(B), ctypes.c_int(N), ctypes.c_int(S),
                            ctypes.c_int(K), _p(idx), _p(d))
    return idx, d
