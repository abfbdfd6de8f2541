"""ctypes loader for the C restatement (oracle/ragfin_oracle.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libragfin_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ragfin_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libragfin_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        f32p, i64p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        L.oracle_synth_rows.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.c_int32, f32p]
        L.oracle_synth_rows.restype = None
        L.oracle_normalize_rows.argtypes = [f32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, f32p]
        L.oracle_normalize_rows.restype = None
        L.oracle_exact_scores.argtypes = [f32p, ctypes.c_int64, ctypes.c_int32, f32p, f32p]
        L.oracle_exact_scores.restype = None
        L.oracle_topk.argtypes = [f32p, ctypes.c_int64, ctypes.c_int32, f32p, ctypes.c_int32,
                                  ctypes.c_int32, ctypes.c_int64, i64p, f32p]
        L.oracle_topk.restype = ctypes.c_int
        _lib = L
    return _lib


def _f32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def synth_rows(seed, row0, n, dim, dup_every=0, zero_every=0):
    out = np.empty((n, dim), dtype=np.float32)
    lib().oracle_synth_rows(seed, row0, n, dim, dup_every, zero_every, _f32(out))
    return out


def normalize_rows(x, dtype="f32"):
    code = {"f32": 0, "bf16": 1, "f16": 2}[dtype]
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    out = np.empty_like(x)
    lib().oracle_normalize_rows(_f32(x), x.shape[0], x.shape[1], code, _f32(out))
    return out


def exact_scores(stored, qhat):
    stored = np.ascontiguousarray(stored, dtype=np.float32)
    qhat = np.ascontiguousarray(qhat, dtype=np.float32)
    s = np.empty(stored.shape[0], dtype=np.float32)
    lib().oracle_exact_scores(_f32(stored), stored.shape[0], stored.shape[1], _f32(qhat), _f32(s))
    return s


def cosine_topk(queries, stored, k, id_base=0):
    stored = np.ascontiguousarray(stored, dtype=np.float32)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    ids = np.empty((q.shape[0], k), dtype=np.int64)
    sc = np.empty((q.shape[0], k), dtype=np.float32)
    rc = lib().oracle_topk(_f32(stored), stored.shape[0], stored.shape[1], _f32(q), q.shape[0], k,
                           id_base, ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _f32(sc))
    if rc != 0:
        raise MemoryError("oracle_topk")
    return ids, sc
