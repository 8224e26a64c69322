"""ctypes wrapper around oracle/walk.c.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp oracle/walk.c -> oracle/_build/liboracle.so (git-ignored, travels with gpurun)."""
    src = os.path.join(_HERE, "walk.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_uniform.restype = ctypes.c_double
        L.oracle_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32]
        L.oracle_build_cdf.restype = None
        L.oracle_build_cdf.argtypes = [ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_walk.restype = ctypes.c_int64
        L.oracle_walk.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                                          ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p,
                                                          ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def build_cdf(C, mode: int) -> np.ndarray:
    indptr = np.ascontiguousarray(C.indptr, dtype=np.int64)
    data = np.ascontiguousarray(C.data, dtype=np.float64)
    if mode == 1:
        out = np.empty(data.shape[0], dtype=np.float64)
        lib().oracle_build_cdf(C.shape[0], _p(indptr), _p(data), 1, None, _p(out))
    else:
        out = np.empty(data.shape[0], dtype=np.float32)
        lib().oracle_build_cdf(C.shape[0], _p(indptr), _p(data), 0, _p(out), None)
    return out


def walk(C, cdf: np.ndarray, start_cells, n_transitions: int, seed: int = 0, round_idx: int = 0,
         walker_begin: int = 0, uniforms=None, nthreads: int = 0, strict: bool = True) -> np.ndarray:
    indptr = np.ascontiguousarray(C.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(C.indices, dtype=np.int32)
    starts = np.ascontiguousarray(start_cells, dtype=np.int32)
    W = starts.shape[0]
    out = np.empty((W, n_transitions + 1), dtype=np.int32)
    c32 = cdf if cdf.dtype == np.float32 else None
    c64 = cdf if cdf.dtype == np.float64 else None
    assert (c32 is None) != (c64 is None)
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        assert uniforms.shape == (W, n_transitions)
    miss = lib().oracle_walk(_p(indptr), _p(indices), _p(c32), _p(c64), _p(starts), W, walker_begin,
                             n_transitions, seed, round_idx, _p(uniforms), _p(out), nthreads)
    if miss and strict:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    return out


def uniform(seed, round_idx, walker, t) -> float:
    return lib().oracle_uniform(seed, round_idx, walker, t)
