"""TEST INFRASTRUCTURE — ctypes access to oracle/_ref/libkpconv_ref.so (the unmodified reference C++ cores of the
KPConv operator set, see oracle/build_ref.py).  Used by tests and bench.py's CPU legs only."""
from __future__ import annotations

import ctypes

import numpy as np

from . import build_ref

_lib = None


def available() -> bool:
    return build_ref.build() is not None


def lib():
    global _lib
    if _lib is None:
        path = build_ref.build()
        if path is None:
            raise RuntimeError("oracle/_ref/libkpconv_ref.so is not built and /root/reference is absent")
        _lib = ctypes.CDLL(path)
    return _lib


def _f(a):
    return np.ascontiguousarray(a, np.float32)


def _i(a):
    return np.ascontiguousarray(a, np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def batch_neighbors(queries, supports, q_batches, s_batches, radius, nanoflann=True):
    """tf_batch_neighbors.cpp:93 (`batch_nanoflann_neighbors`) or the brute-force `batch_ordered_neighbors` -> int32 [Nq, W]."""
    q, s, qb, sb = _f(queries), _f(supports), _i(q_batches), _i(s_batches)
    L = lib()
    L.ref_batch_neighbors.restype = ctypes.c_int
    W = L.ref_batch_neighbors(_p(q), ctypes.c_int(len(q)), _p(s), ctypes.c_int(len(s)), _p(qb), _p(sb), ctypes.c_int(len(qb)),
                              ctypes.c_float(radius), ctypes.c_int(int(nanoflann)))
    out = np.empty((len(q), W), np.int32)
    L.ref_neighbors_copy(_p(out))
    return out


def batch_grid_subsampling(points, batches, dl):
    """tf_batch_subsampling.cpp -> (sub_points [M,3], sub_batches [B]) in the reference's hash-map iteration order."""
    p, b = _f(points), _i(batches)
    L = lib()
    L.ref_batch_grid_subsampling.restype = ctypes.c_int
    ob = np.empty(len(b), np.int32)
    M = L.ref_batch_grid_subsampling(_p(p), ctypes.c_int(len(p)), _p(b), ctypes.c_int(len(b)), ctypes.c_float(dl), _p(ob))
    out = np.empty((M, 3), np.float32)
    L.ref_subsampling_copy(_p(out), None, None)
    return out, ob


def grid_subsampling(points, features=None, classes=None, dl=0.1):
    """cpp_wrappers grid_subsampling.cpp (what `grid_subsampling.compute` runs) -> (points [, features][, classes [M,ld]])."""
    p = _f(points)
    f = _f(features) if features is not None else None
    c = _i(classes) if classes is not None else None
    if c is not None and c.ndim == 1:
        c = c[:, None]
    fdim = f.shape[1] if f is not None else 0
    ldim = c.shape[1] if c is not None else 0
    L = lib()
    L.ref_grid_subsampling.restype = ctypes.c_int
    M = L.ref_grid_subsampling(_p(p), ctypes.c_int(len(p)), _p(f), ctypes.c_int(fdim), _p(c), ctypes.c_int(ldim), ctypes.c_float(dl))
    op = np.empty((M, 3), np.float32)
    of = np.empty((M, fdim), np.float32) if fdim else None
    oc = np.empty((M, ldim), np.int32) if ldim else None
    L.ref_subsampling_copy(_p(op), _p(of), _p(oc))
    return op, of, oc
