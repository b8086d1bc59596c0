"""ctypes binding of oracle/ctc_prefix_oracle.c (TEST INFRASTRUCTURE)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libctc_prefix_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_i64p = ctypes.POINTER(ctypes.c_longlong)


def build(force=False):
    src = os.path.join(_HERE, "ctc_prefix_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_logaddexpf.restype = ctypes.c_float
        L.oracle_logaddexpf.argtypes = [ctypes.c_float, ctypes.c_float]
        L.oracle_blank_state.restype = None
        L.oracle_blank_state.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, _f32p]
        L.oracle_extend.restype = ctypes.c_int
        L.oracle_extend.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, _f32p, _i32p, ctypes.c_int,
                                    ctypes.c_int, _f32p, _f32p]
        L.oracle_extend_many.restype = ctypes.c_int
        L.oracle_extend_many.argtypes = [ctypes.c_int, _f32p, _i64p, _i32p, ctypes.c_int, ctypes.c_int,
                                         _i32p, _i32p, _f32p, _i64p, _i32p, ctypes.c_int,
                                         _f32p, _f32p, _i64p, ctypes.c_int]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def blank_state(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    r = np.empty((x.shape[0], 2), dtype=np.float32)
    lib().oracle_blank_state(_p(x, _f32p), x.shape[0], x.shape[1], _p(r, _f32p))
    return r


def extend(x, prefix_len, last_tok, r_prev, cands, mode="cheap"):
    x = np.ascontiguousarray(x, dtype=np.float32)
    r_prev = np.ascontiguousarray(r_prev, dtype=np.float32)
    T, V = x.shape
    full = mode == "full"
    cs = np.arange(V, dtype=np.int32) if full else np.ascontiguousarray(cands, dtype=np.int32)
    C = len(cs)
    psi = np.empty(C, dtype=np.float32)
    r = np.empty((C, T, 2), dtype=np.float32)
    rc = lib().oracle_extend(_p(x, _f32p), T, V, V, int(prefix_len), int(last_tok), _p(r_prev, _f32p),
                             _p(cs, _i32p), C, int(full), _p(psi, _f32p), _p(r, _f32p))
    if rc:
        raise IndexError("prefix longer than the encoder output (ctc.py:85)")
    return psi, r


def extend_many(x, x_off, T, V, ldx, prefix_len, last_tok, r_prev, rp_off, cands, C, psi, r, r_off, threads=0):
    """Thin pass-through; all arrays must already be contiguous numpy arrays of the
    right dtype (float32 / int32 / int64 offsets in elements)."""
    n = len(T)
    return lib().oracle_extend_many(n, _p(x, _f32p), _p(x_off, _i64p), _p(T, _i32p), int(V), int(ldx),
                                    _p(prefix_len, _i32p), _p(last_tok, _i32p), _p(r_prev, _f32p),
                                    _p(rp_off, _i64p), _p(cands, _i32p), int(C),
                                    _p(psi, _f32p), _p(r, _f32p), _p(r_off, _i64p), int(threads))


def max_threads():
    return lib().oracle_max_threads()
