"""ctypes front end for oracle/libntru_oracle.so (C restatement of index.js).

TEST INFRASTRUCTURE ONLY -- see oracle/ntru_ref_port.c.  Used by tests/ as a
second checker and by bench.py as the CPU baseline ("port").
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_I64P = ctypes.POINTER(ctypes.c_int64)


def build() -> str:
    """Compile the C oracle in place (gcc only) and return the .so path."""
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return os.path.join(_HERE, "libntru_oracle.so")


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libntru_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_multiply.restype = ctypes.c_int
        _LIB.oracle_divide.restype = ctypes.c_int
        _LIB.oracle_encrypt_batch.restype = ctypes.c_int
        _LIB.oracle_decrypt_batch.restype = ctypes.c_int
        _LIB.oracle_sum.restype = ctypes.c_int
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(_I64P) if a is not None else None


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def multiply(a, b, p):
    a, b = _c64(a), _c64(b)
    out = np.zeros(max(1, len(a) + len(b) - 1), dtype=np.int64)
    margin = ctypes.c_double(0.0)
    n = lib().oracle_multiply(_p(a), len(a), _p(b), len(b), ctypes.c_int64(p), _p(out), ctypes.byref(margin))
    return [int(x) for x in out[:n]], margin.value


def divide(a, b, p):
    a, b = _c64(a), _c64(b)
    quo = np.zeros(max(1, len(a)), dtype=np.int64)
    rem = np.zeros(len(a) + len(b), dtype=np.int64)
    lq, lr = ctypes.c_int(0), ctypes.c_int(0)
    rc = lib().oracle_divide(_p(a), len(a), _p(b), len(b), ctypes.c_int64(p), _p(quo), ctypes.byref(lq),
                             _p(rem), ctypes.byref(lr))
    if rc == -1:
        raise ValueError("Cannot divide by zero polynomial.")
    if rc == -2:
        raise ValueError("No inverse exists for division.")
    return {"quotient": [int(x) for x in quo[:lq.value]], "remainder": [int(x) for x in rem[:lr.value]]}


def encrypt_batch(h, r, m, q, *, nthreads=0, witness=True):
    """h: (N,) shared or (B,N) per-row; r, m: (B,N).  Returns dict of int64 arrays + fft margin."""
    r, m, h = _c64(r), _c64(m), _c64(h)
    B, N = r.shape
    stride = 0 if h.ndim == 1 else N
    value = np.zeros((B, N), dtype=np.int64)
    quo = np.zeros((B, N + 1), dtype=np.int64) if witness else None
    rem = np.zeros((B, N + 1), dtype=np.int64) if witness else None
    margin = ctypes.c_double(0.0)
    rc = lib().oracle_encrypt_batch(N, ctypes.c_int64(q), _p(h), stride, ctypes.c_long(B), _p(r), _p(m),
                                    _p(value), _p(quo), _p(rem), nthreads, ctypes.byref(margin))
    if rc:
        raise RuntimeError(f"oracle_encrypt_batch rc={rc}")
    return {"value": value, "quotientE": quo, "remainderE": rem, "fft_margin": margin.value}


def decrypt_batch(f, fp, e, q, p=3, *, nthreads=0, witness=True):
    f, fp, e = _c64(f), _c64(fp), _c64(e)
    B, N = e.shape
    stride = 0 if f.ndim == 1 else N
    value = np.zeros((B, N), dtype=np.int64)
    outs = [np.zeros((B, N + 1), dtype=np.int64) if witness else None for _ in range(4)]
    margin = ctypes.c_double(0.0)
    rc = lib().oracle_decrypt_batch(N, ctypes.c_int64(q), ctypes.c_int64(p), _p(f), _p(fp), stride,
                                    ctypes.c_long(B), _p(e), _p(value), *[_p(o) for o in outs], nthreads,
                                    ctypes.byref(margin))
    if rc:
        raise RuntimeError(f"oracle_decrypt_batch rc={rc}")
    return {"value": value, "quotient1": outs[0], "remainder1": outs[1], "quotient2": outs[2],
            "remainder2": outs[3], "fft_margin": margin.value}


def sum_batch(e, q):
    e = _c64(e)
    B, N = e.shape
    out = np.zeros(N, dtype=np.int64)
    lib().oracle_sum(N, ctypes.c_int64(q), ctypes.c_long(B), _p(e), _p(out))
    return out
