"""Host-side polynomial helpers with the reference's exported names and behaviour.

north_star keeps key generation and polynomial inversion on the host: they are
sequential, per key, and outside the hot path.  This module mirrors the named
exports of the reference's ``index.js`` (cited per function) so that code
written against the reference finds the same functions; products are computed
with exact integer arithmetic (what the reference's float64 FFT + Math.round
evaluates to) instead of an FFT.

Nothing here runs on the encrypt/decrypt hot path -- that is ``engine.py``.
"""
from __future__ import annotations

import math
import secrets
from typing import Callable, List, Optional, Sequence

import numpy as np


def _rem(a: int, p: int) -> int:
    """JavaScript's truncating ``%``."""
    r = abs(a) % abs(p)
    return -r if a < 0 else r


def degree(poly: Sequence[int]) -> int:
    """index.js:210-215."""
    for i in range(len(poly) - 1, -1, -1):
        if poly[i] != 0:
            return i
    return -1


def trimPolynomial(poly: Sequence[int]) -> List[int]:
    """index.js:218-221."""
    d = degree(poly)
    return [int(x) for x in poly[: d + 1]] if d >= 0 else [0]


def modInverse(a: int, p: int) -> Optional[int]:
    """index.js:224-232 (returns None where the reference returns null)."""
    a = a % p
    try:
        return pow(a, -1, p) if a else None
    except ValueError:
        return None


def addPolynomials(a: Sequence[int], b: Sequence[int], p: int) -> List[int]:
    """index.js:235-244."""
    n = max(len(a), len(b))
    out = np.zeros(n, dtype=np.int64)
    out[: len(a)] += np.asarray(a, dtype=np.int64)
    out[: len(b)] += np.asarray(b, dtype=np.int64)
    return trimPolynomial(np.mod(out, p).tolist())


def subtractPolynomials(a: Sequence[int], b: Sequence[int], p: int) -> List[int]:
    """index.js:247-256."""
    n = max(len(a), len(b))
    out = np.zeros(n, dtype=np.int64)
    out[: len(a)] += np.asarray(a, dtype=np.int64)
    out[: len(b)] -= np.asarray(b, dtype=np.int64)
    return trimPolynomial(np.mod(out, p).tolist())


def multiplyPolynomials(a: Sequence[int], b: Sequence[int], p: int) -> List[int]:
    """index.js:319-355: trim(lin(a,b) mod p) into [0,p); [0] when either input is empty."""
    if len(a) == 0 or len(b) == 0:
        return [0]
    c = np.convolve(np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64))
    return trimPolynomial(np.mod(c, p).tolist())


def dividePolynomials(a: Sequence[int], b: Sequence[int], p: int):
    """index.js:358-401: schoolbook long division mod p, same errors as the reference."""
    deg_b = degree(b)
    if deg_b == -1:
        raise ValueError("Cannot divide by zero polynomial.")
    dividend = [int(x) for x in a]
    divisor = [int(x) for x in b]
    quotient = [0] * max(0, degree(a) - deg_b + 1)
    inv = modInverse(divisor[deg_b], p)
    deg_d = degree(dividend)
    while deg_d >= deg_b:
        if inv is None:
            raise ValueError("No inverse exists for division.")
        coeff = _rem(dividend[deg_d] * inv, p)
        shift = deg_d - deg_b
        quotient[shift] = coeff
        for i in range(deg_b + 1):
            v = _rem(dividend[i + shift] - coeff * divisor[i], p)
            dividend[i + shift] = v + p if v < 0 else v
        deg_d = degree(dividend)
    return {"quotient": trimPolynomial(quotient), "remainder": trimPolynomial(dividend)}


def multiplyPolynomialsByScalar(poly: Sequence[int], scalar: int, p: int) -> List[int]:
    """index.js:404-406."""
    return [_rem(int(c) * scalar, p) for c in poly]


def extendedEuclideanAlgorithm(a: Sequence[int], b: Sequence[int], p: int):
    """index.js:425-459."""
    r0, r1 = [int(x) for x in a], [int(x) for x in b]
    s0, s1 = [1], [0]
    while degree(r1) >= 0:
        qr = dividePolynomials(r0, r1, p)
        r0, r1 = r1, qr["remainder"]
        s0, s1 = s1, subtractPolynomials(s0, multiplyPolynomials(qr["quotient"], s1, p), p)
    inv = modInverse(r0[degree(r0)], p)
    if inv is not None and inv != 1:
        r0 = multiplyPolynomialsByScalar(r0, inv, p)
        s0 = multiplyPolynomialsByScalar(s0, inv, p)
    if len(r0) != 1 and r0[0] != 1:
        raise ValueError("invalid_gcd")
    return {"gcd": r0, "inverse": s0}


def generateCustomArray(length: int, numOnes: int, numNegOnes: int,
                        rand32: Optional[Callable[[], int]] = None) -> List[int]:
    """index.js:461-488; ``rand32`` replaces ``crypto.getRandomValues`` (default: the OS CSPRNG)."""
    if numOnes + numNegOnes > length:
        raise ValueError("The total of 1s and -1s cannot exceed the array length.")
    rand32 = rand32 or (lambda: secrets.randbits(32))
    arr = [0] * length
    arr[:numOnes] = [1] * numOnes
    arr[numOnes:numOnes + numNegOnes] = [-1] * numNegOnes
    for i in range(length - 1, 0, -1):
        j = rand32() % (i + 1)
        arr[i], arr[j] = arr[j], arr[i]
    return arr


def _fold(poly: Sequence[int], N: int, p: int) -> List[int]:
    """Remainder of division by 1 - x^N mod p (what dividePolynomials(., I, p).remainder equals)."""
    out = np.zeros(N, dtype=np.int64)
    src = np.asarray(poly, dtype=np.int64)
    for s in range(0, len(src), N):
        part = src[s:s + N]
        out[: len(part)] += part
    return trimPolynomial(np.mod(out, p).tolist())


def polyInv(polyIn: Sequence[int], polyI: Sequence[int], polyMod: int) -> List[int]:
    """index.js:491-514: EEA mod 2 then Newton lifting for powers of two, EEA for a prime."""
    exponent = math.log2(polyMod)
    if round(exponent) == exponent:
        N = len(polyI) - 1
        is_I = list(polyI) == [1] + [0] * (N - 1) + [-1]
        inverse = extendedEuclideanAlgorithm(polyIn, polyI, 2)["inverse"]
        for _ in range(1, int(exponent)):
            twice = multiplyPolynomialsByScalar(inverse, 2, polyMod)
            ptsi = multiplyPolynomials(polyIn, multiplyPolynomials(inverse, inverse, polyMod), polyMod)
            upd = subtractPolynomials(twice, ptsi, polyMod)
            if is_I:
                inverse = _fold(upd, N, polyMod)
            else:
                inverse = trimPolynomial(dividePolynomials(upd, polyI, polyMod)["remainder"])
        return inverse
    return extendedEuclideanAlgorithm(polyIn, polyI, polyMod)["inverse"]


def expandArrayToMultiple(array: List[int], multiple: int) -> List[int]:
    """index.js:516-532 (mutates its argument, like the reference)."""
    if not isinstance(array, list):
        raise ValueError("First argument must be an array.")
    if not isinstance(multiple, int) or isinstance(multiple, bool) or multiple <= 0:
        raise ValueError("Multiple must be a positive integer.")
    target = -(-len(array) // multiple) * multiple
    while len(array) < target:
        array.append(0)
    return array


def expandArray(arr: Sequence[int], length: int, fill: int = 0) -> List[int]:
    """index.js:534-536; the reference's RangeError becomes IndexError('RangeError: ...')."""
    if length < len(arr):
        raise IndexError("RangeError: Invalid array length")
    return [int(x) for x in arr] + [fill] * (length - len(arr))


def stringToBits(s: str) -> List[int]:
    """index.js:538-546."""
    bits: List[int] = []
    for ch in s:
        bits.extend(int(c) for c in format(ord(ch), "b").rjust(8, "0"))
    return bits


def bitsToString(bits: Sequence[int]) -> str:
    """index.js:548-556 (parseInt semantics: stops at the first non-binary digit)."""
    out = []
    for i in range(0, len(bits), 8):
        good = ""
        for b in bits[i:i + 8]:
            if b in (0, 1):
                good += str(b)
            else:
                break
        out.append(chr(int(good, 2)) if good else "\0")
    return "".join(out)


def bigintToBits(value: int) -> List[int]:
    """index.js:558-566 (LSB first)."""
    bits = []
    while value > 0:
        bits.append(value & 1)
        value >>= 1
    return bits


def bitsToBigInt(bits: Sequence[int]) -> int:
    """index.js:568-570 (MSB first)."""
    return int("".join(str(b) for b in bits), 2)


def packOutput(maxVal: int, dataLen: int, data: Sequence[int]):
    """index.js:572-596."""
    maxInputBits = math.floor(math.log2(maxVal) + 1)
    per = 252 // maxInputBits
    arrLen = max(math.ceil(dataLen / per) * per, per * 3)
    outputSize = max(math.ceil(arrLen / per), 3)
    inArr = expandArray(data, arrLen, 0)
    expected = [0] * outputSize
    for i, v in enumerate(inArr):
        expected[i // per] += int(v) << ((i % per) * maxInputBits)
    return {"maxInputBits": maxInputBits, "maxOutputBits": per * maxInputBits, "outputSize": outputSize,
            "arrLen": arrLen, "expected": expected}


def unpackInput(maxVal: int, packedBits: int, data: Sequence[int]):
    """index.js:598-620."""
    maxInputBits = math.floor(math.log2(maxVal) + 1)
    per = packedBits // maxInputBits
    mask = (1 << maxInputBits) - 1
    unpacked = [0] * (per * len(data))
    for i, d in enumerate(data):
        for j in range(per):
            unpacked[i * per + j] = (int(d) >> (j * maxInputBits)) & mask
    return {"maxInputBits": maxInputBits, "packedBits": packedBits, "packedSize": len(data),
            "unpackedSize": len(unpacked), "unpacked": trimPolynomial(unpacked)}
