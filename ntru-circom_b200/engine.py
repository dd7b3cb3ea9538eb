"""Thin Python owner of one ``ntru_ctx`` (one GPU): numpy host buffers and raw device pointers.

All arithmetic happens in libntru_b200.so; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from ._lib import NtruError


def _ptr(a) -> Optional[int]:
    """Address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(f"cannot take the address of {type(a)}")


def _host(a, dtype, shape) -> np.ndarray:
    """C-contiguous numpy view of the exact dtype/shape the ABI expects (copies only if needed)."""
    if hasattr(a, "numpy") and not isinstance(a, np.ndarray):
        a = a.numpy()          # CPU torch tensor (possibly pinned): zero-copy
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.shape != tuple(shape):
        raise IndexError(f"RangeError: expected shape {tuple(shape)}, got {a.shape}")
    return a


class Engine:
    """One context of the CUDA engine bound to one device."""

    def __init__(self, N: int, p: int, q: int, device: int = 0):
        self.lib = _lib.load()
        self.N, self.p, self.q = int(N), int(p), int(q)
        h = ctypes.c_void_p()
        rc = self.lib.ntru_create(ctypes.byref(h), self.N, self.p, self.q, int(device))
        if rc != _lib.NTRU_OK:
            raise NtruError(rc, {_lib.NTRU_E_CUDA: "no usable CUDA device (this engine has no CPU fallback)",
                                 _lib.NTRU_E_PARAM: "unsupported parameters N/p/q"}.get(
                                     rc, self.lib.ntru_strerror(rc).decode()))
        self._h = h
        self.device = int(device)
        self.pitch = int(self.lib.ntru_pitch(h))

    # -- plumbing ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.ntru_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != _lib.NTRU_OK:
            raise NtruError(rc, self.lib.ntru_last_error(self._h).decode() or self.lib.ntru_strerror(rc).decode())

    def set_option(self, key: int, value: int):
        self._check(self.lib.ntru_set_option(self._h, key, value))

    def set_path(self, path: int):
        self.set_option(_lib.NTRU_OPT_PATH, path)

    @property
    def launch_count(self) -> int:
        return int(self.lib.ntru_launch_count(self._h))

    @property
    def last_path(self) -> int:
        return int(self.lib.ntru_last_path(self._h))

    def set_timing(self, on: bool):
        self.set_option(_lib.NTRU_OPT_TIMING, 1 if on else 0)

    def timing_reset(self):
        self._check(self.lib.ntru_timing_reset(self._h))

    def timing_read(self) -> dict:
        """{kind: (total_ms, launches)} of the kernels launched since timing_reset (synchronises)."""
        out = {}
        for i, name in enumerate(_lib.KERNEL_KINDS):
            ms, n = ctypes.c_double(0), ctypes.c_uint64(0)
            self._check(self.lib.ntru_timing_read(self._h, i, ctypes.byref(ms), ctypes.byref(n)))
            if n.value:
                out[name] = (ms.value, int(n.value))
        return out

    def stream(self) -> int:
        return int(self.lib.ntru_stream(self._h) or 0)

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.ntru_set_stream(self._h, cuda_stream))

    def sync(self):
        self._check(self.lib.ntru_sync(self._h))

    # -- keys ------------------------------------------------------------------------
    def set_public_key(self, h):
        h = _host(h, np.uint16, (self.N,))
        self._check(self.lib.ntru_set_public_key(self._h, _ptr(h)))

    def set_private_key(self, f, fp):
        f = _host(f, np.int8, (self.N,))
        fp = _host(fp, np.uint8, (self.N,))
        self._check(self.lib.ntru_set_private_key(self._h, _ptr(f), _ptr(fp)))

    # -- host-buffer batches (packed rows) ------------------------------------------------
    def encrypt_batch(self, r, m, *, h=None, witness=True, out=None):
        """r, m: (B,N).  h: None (context key) or (B,N) distinct keys.  Returns dict of numpy arrays."""
        N = self.N
        r = np.ascontiguousarray(r, dtype=np.uint8)
        B = r.shape[0]
        r = _host(r, np.uint8, (B, N))
        m_arr = np.asarray(m)
        wide = m_arr.dtype.itemsize > 1 and m_arr.size and int(m_arr.max(initial=0)) > 255
        m_arr = _host(m_arr, np.uint16 if wide else np.uint8, (B, N))
        out = out or {}
        value = out.get("value", np.empty((B, N), dtype=np.uint16))
        quo = out.get("quotientE", np.empty((B, N + 1), dtype=np.uint16)) if witness else None
        rem = out.get("remainderE", np.empty((B, N + 1), dtype=np.uint16)) if witness else None
        if h is not None:
            if wide:
                raise NtruError(_lib.NTRU_E_UNSUPPORTED, "wide messages with per-row keys")
            h = _host(h, np.uint16, (B, N))
            rc = self.lib.ntru_encrypt_batch_keys(self._h, B, _ptr(h), _ptr(r), _ptr(m_arr), _ptr(value),
                                                  _ptr(quo), _ptr(rem))
        elif wide:
            rc = self.lib.ntru_encrypt_batch_wide(self._h, B, _ptr(r), _ptr(m_arr), _ptr(value), _ptr(quo),
                                                  _ptr(rem))
        else:
            rc = self.lib.ntru_encrypt_batch(self._h, B, _ptr(r), _ptr(m_arr), _ptr(value), _ptr(quo), _ptr(rem))
        self._check(rc)
        return {"value": value, "quotientE": quo, "remainderE": rem}

    def decrypt_batch(self, e, *, f=None, fp=None, witness=True, out=None):
        N = self.N
        e = np.ascontiguousarray(e, dtype=np.uint16)
        B = e.shape[0]
        e = _host(e, np.uint16, (B, N))
        out = out or {}
        value = out.get("value", np.empty((B, N), dtype=np.uint8))
        q1 = out.get("quotient1", np.empty((B, N + 1), dtype=np.uint16)) if witness else None
        r1 = out.get("remainder1", np.empty((B, N + 1), dtype=np.uint16)) if witness else None
        q2 = out.get("quotient2", np.empty((B, N + 1), dtype=np.uint8)) if witness else None
        r2 = out.get("remainder2", np.empty((B, N + 1), dtype=np.uint8)) if witness else None
        if f is not None or fp is not None:
            f = _host(f, np.int8, (B, N))
            fp = _host(fp, np.uint8, (B, N))
            rc = self.lib.ntru_decrypt_batch_keys(self._h, B, _ptr(f), _ptr(fp), _ptr(e), _ptr(value), _ptr(q1),
                                                  _ptr(r1), _ptr(q2), _ptr(r2))
        else:
            rc = self.lib.ntru_decrypt_batch(self._h, B, _ptr(e), _ptr(value), _ptr(q1), _ptr(r1), _ptr(q2),
                                             _ptr(r2))
        self._check(rc)
        return {"value": value, "quotient1": q1, "remainder1": r1, "quotient2": q2, "remainder2": r2}

    def verify_keys_batch(self, f, fq, fp, g):
        """verifyKeysInputs (index.js:141-197) for B keys: the quotientI / remainderI of the three
        multiply + divide-by-(1 - x^N) pairs (fq case and h case mod q, fp case mod p)."""
        N = self.N
        f = np.ascontiguousarray(f, dtype=np.int8)
        B = f.shape[0]
        f = _host(f, np.int8, (B, N))
        fq = _host(fq, np.uint16, (B, N))
        fp = _host(fp, np.uint8, (B, N))
        g = _host(g, np.int8, (B, N))
        out = {k: np.empty((B, N + 1), dtype=np.uint8 if k.endswith("_fp") else np.uint16)
               for k in ("quotient_fq", "remainder_fq", "quotient_fp", "remainder_fp", "quotient_h", "remainder_h")}
        self._check(self.lib.ntru_verify_keys_batch(self._h, B, _ptr(f), _ptr(fq), _ptr(fp), _ptr(g),
                                                    *[_ptr(out[k]) for k in ("quotient_fq", "remainder_fq", "quotient_fp",
                                                                             "remainder_fp", "quotient_h", "remainder_h")]))
        return out

    def keygen_batch(self, f, g):
        """loadPrivateKeyF + generatePublicKeyH for B keys (index.js:30-79): f, g ternary (B, N).  Returns fixed-length
        fq, fp, h and the validity mask (f invertible modulo 2 and modulo p)."""
        N = self.N
        f = np.ascontiguousarray(f, dtype=np.int8)
        B = f.shape[0]
        f = _host(f, np.int8, (B, N))
        g = _host(g, np.int8, (B, N))
        fq = np.empty((B, N), dtype=np.uint16)
        fp = np.empty((B, N), dtype=np.uint8)
        h = np.empty((B, N), dtype=np.uint16)
        valid = np.empty(B, dtype=np.uint8)
        self._check(self.lib.ntru_keygen_batch(self._h, B, _ptr(f), _ptr(g), _ptr(fq), _ptr(fp), _ptr(h), _ptr(valid)))
        return {"fq": fq, "fp": fp, "h": h, "valid": valid.astype(bool)}

    def muldiv_dev(self, B, x, y, mod_p, quotient=None, remainder=None):
        """multiplyPolynomials(x, y, mod) + dividePolynomials(., 1 - x^N, mod), device-resident rows."""
        self._check(self.lib.ntru_muldiv_dev(self._h, B, _ptr(x), _ptr(y), int(bool(mod_p)), _ptr(quotient), _ptr(remainder)))

    # -- packOutput / unpackInput (index.js:572-620) on device rows --------------------------------------
    def pack_geometry(self, max_val: int, data_len: int):
        v = [ctypes.c_int() for _ in range(4)]
        self._check(self.lib.ntru_pack_geometry(int(max_val), int(data_len), *[ctypes.byref(x) for x in v]))
        return tuple(x.value for x in v)     # maxInputBits, inputsPerOutput, arrLen, outputSize

    def pack_output_dev(self, B, data, elem_bytes, data_len, pitch, max_val, out):
        self._check(self.lib.ntru_pack_output_dev(self._h, B, _ptr(data), int(elem_bytes), int(data_len), int(pitch), int(max_val),
                                                  _ptr(out)))

    def unpack_input_dev(self, B, data, n_elems, max_val, packed_bits, out, elem_bytes, pitch):
        self._check(self.lib.ntru_unpack_input_dev(self._h, B, _ptr(data), int(n_elems), int(max_val), int(packed_bits), _ptr(out),
                                                   int(elem_bytes), int(pitch)))

    def sum(self, e):
        e = np.ascontiguousarray(e, dtype=np.uint16)
        B = e.shape[0]
        e = _host(e, np.uint16, (B, self.N))
        out = np.empty(self.N, dtype=np.uint16)
        self._check(self.lib.ntru_sum(self._h, B, _ptr(e), _ptr(out)))
        return out

    # -- device-resident batches (pitched rows; arguments: torch CUDA tensors or raw pointers) --------
    def encrypt_dev(self, B, r, m, value=None, quotientE=None, remainderE=None, h_rows=None):
        self._check(self.lib.ntru_encrypt_dev(self._h, B, _ptr(h_rows), _ptr(r), _ptr(m), _ptr(value),
                                              _ptr(quotientE), _ptr(remainderE)))

    def decrypt_dev(self, B, e, value=None, quotient1=None, remainder1=None, quotient2=None, remainder2=None,
                    f_rows=None, fp_rows=None):
        self._check(self.lib.ntru_decrypt_dev(self._h, B, _ptr(f_rows), _ptr(fp_rows), _ptr(e), _ptr(value),
                                              _ptr(quotient1), _ptr(remainder1), _ptr(quotient2),
                                              _ptr(remainder2)))

    def sum_partial_dev(self, B, e, partial):
        self._check(self.lib.ntru_sum_partial_dev(self._h, B, _ptr(e), _ptr(partial)))

    def sum_finalize_dev(self, partial, out):
        self._check(self.lib.ntru_sum_finalize_dev(self._h, _ptr(partial), _ptr(out)))

    # ---- cross-GPU sum over peer memory (include/ntru_b200.h: ntru_xchg_*, ntru_sum_allreduce_dev) ----
    def xchg_create(self, world: int, rank: int) -> bytes:
        """Allocates this rank's exchange window; returns its 64-byte CUDA IPC handle (all-gather it)."""
        buf = (ctypes.c_ubyte * 64)()
        self._check(self.lib.ntru_xchg_create(self._h, int(world), int(rank), buf))
        return bytes(buf)

    def xchg_connect(self, handles: bytes):
        """handles: world x 64 bytes, rank-major."""
        buf = (ctypes.c_ubyte * len(handles)).from_buffer_copy(handles)
        self._check(self.lib.ntru_xchg_connect(self._h, buf))

    def sum_allreduce_dev(self, B, e, out):
        """out[k] = column sums mod q over every rank's rows (local rows only when no exchange is connected)."""
        self._check(self.lib.ntru_sum_allreduce_dev(self._h, B, _ptr(e) if B else None, _ptr(out)))

    def sample_r_dev(self, B, dr, seed, row0, r):
        self._check(self.lib.ntru_sample_r_dev(self._h, B, int(dr), int(seed), int(row0), _ptr(r)))


def sampler_rand32(seed: int, row: int, i: int) -> int:
    """Host copy of the device sampler's counter-based generator (csrc/generic_kernels.cu:sampler_rand32)."""
    M = (1 << 64) - 1
    z = (seed + 0x9E3779B97F4A7C15 * (row * 2048 + i)) & M
    z = (z + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    z ^= z >> 31
    return z >> 32
