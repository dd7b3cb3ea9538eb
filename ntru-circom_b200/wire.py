"""Host-side view of the field-element wire format of Engine.encrypt_batch_packed / decrypt_batch_packed.

A row of `width` coefficients below 2^bits crosses the host link as packOutput(maxVal, width, row).expected
(index.js:572-596): n = floor(252 / bits) coefficients per BN254 field element, coefficient j of an element at bit
j * bits, at least three elements per row; an element is eight little-endian uint32 words.  These two functions are
the vectorised numpy forms of packOutput / unpackInput for whole batches -- for callers that hold plain coefficient
rows on the host (the engine packs and unpacks on the device; tests/test_host.py pins both against the BigInt forms).
"""
from __future__ import annotations

import numpy as np


def geometry(max_val: int, width: int):
    """(bits, coefficients per element, elements per row) -- packOutput's maxInputBits, numInputsPerOutput, outputSize."""
    bits = int(max_val).bit_length()          # floor(log2(maxVal) + 1), index.js:573
    if bits < 1:
        raise ValueError("maxVal must be at least 1")
    n = 252 // bits
    return bits, n, max(-(-int(width) // n), 3)


def pack_rows(max_val: int, rows) -> np.ndarray:
    """rows (B, width) of integers in [0, 2^bits) -> (B, elems, 8) uint32."""
    rows = np.ascontiguousarray(rows)
    B, width = rows.shape
    bits, n, elems = geometry(max_val, width)
    if rows.size and (int(rows.max()) >> bits or int(rows.min()) < 0):
        raise ValueError("a coefficient does not fit maxVal's bit length")
    v = np.zeros((B, elems * n), dtype=np.uint32)
    v[:, :width] = rows
    v = v.reshape(B, elems, n)
    out = np.zeros((B, elems, 9), dtype=np.uint64)           # word 8 only catches the shifted-out zeros of word 7
    for j in range(n):
        w, sh = divmod(j * bits, 32)
        x = v[:, :, j].astype(np.uint64) << np.uint64(sh)
        out[:, :, w] |= x & np.uint64(0xFFFFFFFF)
        out[:, :, w + 1] |= x >> np.uint64(32)
    return out[:, :, :8].astype(np.uint32)


def unpack_rows(max_val: int, width: int, packed, dtype=np.uint16) -> np.ndarray:
    """(B, elems, 8) uint32 -> rows (B, width): unpackInput(maxVal, n * bits, elements).unpacked, un-trimmed, cut to width."""
    packed = np.ascontiguousarray(packed, dtype=np.uint32)
    bits, n, elems = geometry(max_val, width)
    B = packed.shape[0]
    if packed.shape[1:] != (elems, 8):
        raise IndexError(f"RangeError: expected {elems} field elements of 8 words per row, got {packed.shape[1:]}")
    ext = np.zeros((B, elems, 9), dtype=np.uint64)
    ext[:, :, :8] = packed
    out = np.empty((B, elems, n), dtype=dtype)
    mask = np.uint64((1 << bits) - 1)
    for j in range(n):
        w, sh = divmod(j * bits, 32)
        two = ext[:, :, w] | (ext[:, :, w + 1] << np.uint64(32))
        out[:, :, j] = ((two >> np.uint64(sh)) & mask).astype(dtype)
    return out.reshape(B, elems * n)[:, :width]
