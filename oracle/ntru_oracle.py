"""CPU oracle for the numtel/ntru-circom hot path.  TEST INFRASTRUCTURE ONLY.

This module is the *checker* for the CUDA engine: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product package (``ntru-circom_b200``)
never imports anything from ``oracle/``.

It restates, function by function, the algorithm of the reference's single
source file ``/root/reference/index.js`` (cited as ``index.js:LINE``), in two
flavours:

* **literal** -- the same operations in the same order as the JavaScript
  (float64 radix-2 FFT + ``Math.round``, schoolbook long division with the
  linear-search ``modInverse`` and ``degree`` rescans, truncating ``%``).
  Pure-Python loops: use at small N (17, 167) and for the upstream KATs.
* **fast** -- exact integer arithmetic with numpy (``lin_conv``) and the closed
  form of ``dividePolynomials(., 1 - x^N, p)``.  Tests
  pin fast == literal at small N, then use fast at N = 509 ... 821.

Parity pinning: the reference cannot run in this image
(no node), and upstream ships no fixed-key ciphertext vectors.  The oracle is
pinned against every upstream KAT that exists for this path
(tests/test_oracle_kats.py): the multiply worked example
(circuits/ntru.circom:33-71), the five multiply vector pairs
(test/circuits.test.js:61-65), the four division cases
(test/circuits.test.js:166-169), the EEA worked example (index.js:408-423),
the string round trips (test/reference.test.js:6-44), the homomorphic sum
vectors (test/reference.test.js:48-61) and the VerifyEncrypt / VerifyDecrypt
constraint systems (circuits/ntru.circom:155-238) restated in
``verify_encrypt`` / ``verify_decrypt`` below, including the tamper case.
End-to-end ciphertext goldens are produced by this oracle with recorded seeds.
"""
from __future__ import annotations

import math
import random
from typing import List, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------
# JavaScript arithmetic helpers
# ----------------------------------------------------------------------------


def js_rem(a: int, p: int) -> int:
    """JavaScript ``a % p`` (truncating remainder, sign follows the dividend)."""
    r = abs(a) % abs(p)
    return -r if a < 0 else r


def js_round(x: float) -> int:
    """``Math.round``: round half towards +infinity (index.js:351)."""
    return int(math.floor(x + 0.5))


# ----------------------------------------------------------------------------
# L1 polynomial arithmetic -- literal restatements
# ----------------------------------------------------------------------------


def degree(poly: Sequence[int]) -> int:
    """index.js:210-215."""
    for i in range(len(poly) - 1, -1, -1):
        if poly[i] != 0:
            return i
    return -1


def trim_polynomial(poly: Sequence[int]) -> List[int]:
    """index.js:218-221 -- drop trailing zeros; the zero polynomial is [0]."""
    d = degree(poly)
    return list(poly[: d + 1]) if d >= 0 else [0]


def mod_inverse(a: int, p: int) -> Optional[int]:
    """index.js:224-232 -- linear search, returns None when no inverse."""
    a = js_rem(js_rem(a, p) + p, p)
    for x in range(1, p):
        if js_rem(a * x, p) == 1:
            return x
    return None


def add_polynomials(a: Sequence[int], b: Sequence[int], p: int) -> List[int]:
    """index.js:235-244."""
    n = max(len(a), len(b))
    out = []
    for i in range(n):
        ca = a[i] if i < len(a) else 0
        cb = b[i] if i < len(b) else 0
        out.append(js_rem(js_rem(ca + cb, p) + p, p))
    return trim_polynomial(out)


def subtract_polynomials(a: Sequence[int], b: Sequence[int], p: int) -> List[int]:
    """index.js:247-256."""
    n = max(len(a), len(b))
    out = []
    for i in range(n):
        ca = a[i] if i < len(a) else 0
        cb = b[i] if i < len(b) else 0
        out.append(js_rem(js_rem(ca - cb, p) + p, p))
    return trim_polynomial(out)


def _fft(re: List[float], im: List[float], invert: bool) -> None:
    """index.js:277-316 -- in-place radix-2 Cooley-Tukey with running twiddle."""
    n = len(re)
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j -= bit
            bit >>= 1
        j += bit
        if i < j:
            re[i], re[j] = re[j], re[i]
            im[i], im[j] = im[j], im[i]
    length = 2
    while length <= n:
        angle = (2 * math.pi / length) * (-1 if invert else 1)
        wl_re, wl_im = math.cos(angle), math.sin(angle)
        half = length // 2
        for i in range(0, n, length):
            w_re, w_im = 1.0, 0.0
            for k in range(half):
                u_re, u_im = re[i + k], im[i + k]
                x_re, x_im = re[i + k + half], im[i + k + half]
                v_re = x_re * w_re - x_im * w_im
                v_im = x_re * w_im + x_im * w_re
                re[i + k], im[i + k] = u_re + v_re, u_im + v_im
                re[i + k + half], im[i + k + half] = u_re - v_re, u_im - v_im
                w_re, w_im = w_re * wl_re - w_im * wl_im, w_re * wl_im + w_im * wl_re
        length <<= 1
    if invert:
        for i in range(n):
            re[i] /= n
            im[i] /= n


#: worst |Re - round(Re)| seen by multiply_polynomials_fft since import (tests read it)
fft_margin = {"max": 0.0}


def multiply_polynomials_fft(a: Sequence[int], b: Sequence[int], p: int) -> List[int]:
    """index.js:319-355, operation for operation (float64 FFT + Math.round)."""
    if len(a) == 0 or len(b) == 0:
        return [0]
    n = 1
    while n < len(a) + len(b) - 1:
        n <<= 1
    a_re = [float(a[i]) if i < len(a) else 0.0 for i in range(n)]
    a_im = [0.0] * n
    b_re = [float(b[i]) if i < len(b) else 0.0 for i in range(n)]
    b_im = [0.0] * n
    _fft(a_re, a_im, False)
    _fft(b_re, b_im, False)
    for i in range(n):
        a_re[i], a_im[i] = (a_re[i] * b_re[i] - a_im[i] * b_im[i],
                            a_re[i] * b_im[i] + a_im[i] * b_re[i])
    _fft(a_re, a_im, True)
    out = []
    for i in range(len(a) + len(b) - 1):
        r = js_round(a_re[i])
        fft_margin["max"] = max(fft_margin["max"], abs(a_re[i] - r))
        out.append(js_rem(js_rem(r, p) + p, p))
    return trim_polynomial(out)


def lin_conv(a: Sequence[int], b: Sequence[int]) -> np.ndarray:
    """Exact integer linear convolution (what the FFT + round equals)."""
    a64 = np.asarray(a, dtype=np.int64)
    b64 = np.asarray(b, dtype=np.int64)
    # worst case N*(q-1)^2 = 701*8191^2 < 2^36: int64 is exact
    return np.convolve(a64, b64)


def multiply_polynomials_exact(a: Sequence[int], b: Sequence[int], p: int) -> List[int]:
    """trim(lin(a,b) mod p) into [0,p) -- the exact-integer meaning of index.js:319-355."""
    if len(a) == 0 or len(b) == 0:
        return [0]
    return trim_polynomial([int(x) for x in np.mod(lin_conv(a, b), p)])


def divide_polynomials(a: Sequence[int], b: Sequence[int], p: int):
    """index.js:358-401 -- literal long division; raises like the reference."""
    if degree(b) == -1:
        raise ValueError("Cannot divide by zero polynomial.")
    dividend = list(a)
    divisor = list(b)
    deg_divisor = degree(divisor)
    quotient = [0] * max(0, degree(a) - deg_divisor + 1)
    while degree(dividend) >= deg_divisor:
        deg_dividend = degree(dividend)
        lead_dividend = dividend[deg_dividend]
        lead_divisor = divisor[deg_divisor]
        inv = mod_inverse(lead_divisor, p)
        if inv is None:
            raise ValueError("No inverse exists for division.")
        coeff = js_rem(lead_dividend * inv, p)
        deg_diff = deg_dividend - deg_divisor
        quotient[deg_diff] = coeff
        for i in range(deg_divisor + 1):
            idx = i + deg_diff
            cur = dividend[idx] if idx < len(dividend) else 0
            v = js_rem(cur - coeff * divisor[i], p)
            if v < 0:
                v += p
            while idx >= len(dividend):
                dividend.append(0)
            dividend[idx] = v
    return {"quotient": trim_polynomial(quotient), "remainder": trim_polynomial(dividend)}


def divide_by_I_closed(a: Sequence[int], N: int, p: int):
    """Closed form of divide_polynomials(a, [1,0,..,0,-1], p) for len(a) <= 2N-1.

    SURVEY.md section 0: quotient[k] = (-a[N+k]) mod p, remainder[k] = (a[k] + a[N+k]) mod p.
    Inputs must already be reduced into [0,p) (they always are on the hot path).
    """
    c = np.zeros(2 * N, dtype=np.int64)
    c[: len(a)] = np.asarray(a, dtype=np.int64)
    hi = c[N:2 * N]
    quotient = np.mod(-hi, p)
    remainder = np.mod(c[:N] + hi, p)
    return {"quotient": trim_polynomial([int(x) for x in quotient]),
            "remainder": trim_polynomial([int(x) for x in remainder])}


def multiply_polynomials_by_scalar(poly: Sequence[int], scalar: int, p: int) -> List[int]:
    """index.js:404-406 (truncating %, no re-centering)."""
    return [js_rem(c * scalar, p) for c in poly]


def extended_euclidean_algorithm(a, b, p, *, mul=None, div=None):
    """index.js:425-459."""
    mul = mul or multiply_polynomials_fft
    div = div or divide_polynomials
    r0, r1 = list(a), list(b)
    s0, s1 = [1], [0]
    while degree(r1) >= 0:
        qr = div(r0, r1, p)
        r0, r1 = r1, qr["remainder"]
        s0, s1 = s1, subtract_polynomials(s0, mul(qr["quotient"], s1, p), p)
    lead = r0[degree(r0)]
    inv = mod_inverse(lead, p)
    if inv is not None and inv != 1:
        r0 = multiply_polynomials_by_scalar(r0, inv, p)
        s0 = multiply_polynomials_by_scalar(s0, inv, p)
    if len(r0) != 1 and r0[0] != 1:
        raise ValueError("invalid_gcd")
    return {"gcd": r0, "inverse": s0}


def generate_custom_array(length: int, num_ones: int, num_neg_ones: int, rng: random.Random) -> List[int]:
    """index.js:461-488 with the WebCrypto draw replaced by ``rng.getrandbits(32)``."""
    if num_ones + num_neg_ones > length:
        raise ValueError("The total of 1s and -1s cannot exceed the array length.")
    arr = [0] * length
    for i in range(num_ones):
        arr[i] = 1
    for i in range(num_ones, num_ones + num_neg_ones):
        arr[i] = -1
    for i in range(length - 1, 0, -1):
        j = rng.getrandbits(32) % (i + 1)
        arr[i], arr[j] = arr[j], arr[i]
    return arr


def poly_inv(poly_in, poly_I, poly_mod, *, mul=None, div=None, div_I=None):
    """index.js:491-514 (EEA mod 2 + Newton lifting for powers of two, EEA otherwise)."""
    mul = mul or multiply_polynomials_fft
    div_I = div_I or divide_polynomials
    exponent = math.log2(poly_mod)
    if round(exponent) == exponent:
        inverse = extended_euclidean_algorithm(poly_in, poly_I, 2, mul=mul, div=div)["inverse"]
        for _ in range(1, int(exponent)):
            twice = multiply_polynomials_by_scalar(inverse, 2, poly_mod)
            ptsi = mul(poly_in, mul(inverse, inverse, poly_mod), poly_mod)
            upd = subtract_polynomials(twice, ptsi, poly_mod)
            inverse = trim_polynomial(div_I(upd, poly_I, poly_mod)["remainder"])
        return inverse
    return extended_euclidean_algorithm(poly_in, poly_I, poly_mod, mul=mul, div=div)["inverse"]


# ----------------------------------------------------------------------------
# L0 format helpers
# ----------------------------------------------------------------------------


def expand_array(arr: Sequence[int], length: int, fill: int = 0) -> List[int]:
    """index.js:534-536 -- JS throws RangeError when arr is longer than length."""
    if length - len(arr) < 0:
        raise IndexError("RangeError: Invalid array length")
    return list(arr) + [fill] * (length - len(arr))


def expand_array_to_multiple(array: List[int], multiple: int) -> List[int]:
    """index.js:516-532 (mutates and returns its argument)."""
    if not isinstance(array, list):
        raise ValueError("First argument must be an array.")
    if not isinstance(multiple, int) or multiple <= 0:
        raise ValueError("Multiple must be a positive integer.")
    target = -(-len(array) // multiple) * multiple
    while len(array) < target:
        array.append(0)
    return array


def string_to_bits(s: str) -> List[int]:
    """index.js:538-546 -- 8 bits per UTF-16 code unit, MSB first (more if > 0xFF)."""
    bits: List[int] = []
    for ch in s:
        b = format(ord(ch), "b").rjust(8, "0")
        bits.extend(int(c) for c in b)
    return bits


def bits_to_string(bits: Sequence[int]) -> str:
    """index.js:548-556 (parseInt(byte.join(''), 2) stops at the first non-binary digit)."""
    out = []
    for i in range(0, len(bits), 8):
        txt = "".join(str(b) for b in bits[i:i + 8])
        good = ""
        for c in txt:
            if c in "01":
                good += c
            else:
                break
        code = int(good, 2) if good else 0  # JS: NaN -> String.fromCharCode(NaN) == '\0'
        out.append(chr(code))
    return "".join(out)


# ----------------------------------------------------------------------------
# L2 scheme -- class NTRU (index.js:7-207)
# ----------------------------------------------------------------------------


class NTRU:
    """Restatement of ``class NTRU`` (index.js:7-207).

    ``literal=True`` follows the JavaScript operation for operation (FFT
    multiply, long division); ``literal=False`` uses exact integer convolution
    and the closed-form division by ``I`` -- identical results (tests pin it).
    ``rng`` replaces WebCrypto so that runs are reproducible.
    """

    def __init__(self, options: Optional[dict] = None, *, literal: bool = False,
                 rng: Optional[random.Random] = None):
        self.N, self.p, self.q = 167, 3, 128
        self.df, self.dg, self.dr = 61, 20, 18
        self.f = self.fp = self.fq = self.g = self.h = None
        for k, v in (options or {}).items():
            setattr(self, k, v)
        self.I = [0] * (self.N + 1)          # index.js:25-27
        self.I[0] = 1
        self.I[-1] = -1
        self.literal = literal
        self.rng = rng or random.Random(0)

    # -- arithmetic back ends ------------------------------------------------
    def _mul(self, a, b, p):
        return multiply_polynomials_fft(a, b, p) if self.literal else multiply_polynomials_exact(a, b, p)

    def _div_I(self, a, I, p):
        if self.literal:
            return divide_polynomials(a, I, p)
        return divide_by_I_closed(a, self.N, p) if len(a) <= 2 * self.N else divide_polynomials(a, I, p)

    # -- keys (index.js:30-79) -------------------------------------------------
    def loadPrivateKeyF(self, f_arr):
        p, q = self.p, self.q
        self.f = list(f_arr)
        self.fq = poly_inv(self.f, self.I, q, mul=self._mul, div_I=self._div_I_generic)
        self.fp = poly_inv(self.f, self.I, p, mul=self._mul, div_I=self._div_I_generic)
        fmodq = [q - 1 if x == -1 else x for x in self.f]
        fmodp = [p - 1 if x == -1 else x for x in self.f]
        fq_div = self._div_I(self._mul(self.fq, fmodq, q), self.I, q)
        if len(fq_div["remainder"]) != 1 and fq_div["remainder"][0] != 1:
            raise ValueError("invalid fq")
        fp_div = self._div_I(self._mul(self.fp, fmodp, p), self.I, p)
        if len(fp_div["remainder"]) != 1 and fp_div["remainder"][0] != 1:
            raise ValueError("invalid fp")
        return True

    def _div_I_generic(self, a, I, p):
        """Division by I for dividends of any length (polyInv's lifting step)."""
        if self.literal:
            return divide_polynomials(a, I, p)
        # fold exactly like the long division does: x^N == 1
        N = self.N
        a = [int(x) for x in a]
        rem = [0] * N
        for k, v in enumerate(a):
            rem[k % N] += v
        return {"quotient": [0], "remainder": trim_polynomial([js_rem(js_rem(v, p) + p, p) for v in rem])}

    def generatePrivateKeyF(self):
        tries = 0
        ok = False
        while (not ok or not (self.fq and self.fp)) and tries < 100:
            tries += 1
            try:
                ok = self.loadPrivateKeyF(generate_custom_array(self.N, self.df, self.df - 1, self.rng))
            except ValueError:
                pass
        if not self.fq or not self.fp:
            raise ValueError("Could not find invertible f")

    def generateNewPublicKeyGH(self):
        self.g = generate_custom_array(self.N, self.dg, self.dg, self.rng)
        self.generatePublicKeyH()

    def generatePublicKeyH(self):
        if not self.f:
            raise ValueError("missing private key F")
        if not self.g:
            raise ValueError("missing private key G")
        p_fq = multiply_polynomials_by_scalar(self.fq, self.p, self.q)
        p_fq_g = self._mul(p_fq, self.g, self.q)
        self.h = trim_polynomial(self._div_I(p_fq_g, self.I, self.q)["remainder"])

    def verifyKeysInputs(self) -> dict:
        """index.js:141-197: the VerifyInverse witness for the fq, fp and h cases."""
        for name, label in (("f", "private key F"), ("fq", "private key Fq"), ("fp", "private key Fp"),
                            ("g", "private key G")):
            if not getattr(self, name):
                raise ValueError(f"missing {label}")
        if not self.h:
            raise ValueError("missing public key H")
        q, p, N = self.q, self.p, self.N
        fmodq = [q - 1 if x == -1 else x for x in self.f]
        fmodp = [p - 1 if x == -1 else x for x in self.f]
        fqp = [x * p for x in self.fq]                       # index.js:155: NOT reduced mod q
        g = [q - 1 if x == -1 else x for x in self.g]
        fq_div = self._div_I(self._mul(self.fq, fmodq, q), self.I, q)
        fp_div = self._div_I(self._mul(self.fp, fmodp, p), self.I, p)
        h_div = self._div_I(self._mul(fqp, g, q), self.I, q)
        case = lambda mod, nbits, a, b, d: {"params": [mod, nbits, N], "inputs": {       # noqa: E731
            "f": expand_array(a, N, 0), "fq": expand_array(b, N, 0),
            "quotientI": expand_array(d["quotient"], N + 1, 0), "remainderI": expand_array(d["remainder"], N + 1, 0)}}
        return {"fq": case(q, self.calculateNq(), fmodq, self.fq, fq_div),
                "fp": case(p, self.calculateNp(), fmodp, self.fp, fp_div),
                "h": case(q, self.calculateNq(), g, fqp, h_div)}

    def key_is_valid(self) -> bool:
        """f*fq == 1 (mod q, x^N-1) and f*fp == 1 (mod p, x^N-1) -- what the reference *meant* to check."""
        N = self.N
        def cyc(a, b, mod):
            c = lin_conv(a, b)
            out = np.zeros(N, dtype=np.int64)
            for k in range(len(c)):
                out[k % N] += c[k]
            return np.mod(out, mod)
        one = np.zeros(N, dtype=np.int64)
        one[0] = 1
        return bool(np.array_equal(cyc(self.f, self.fq, self.q), one)
                    and np.array_equal(cyc(self.f, self.fp, self.p), one))

    # -- hot path (index.js:80-140) -----------------------------------------------
    def sample_r(self):
        """index.js:89 -- exact-weight ternary with -1 sent to p-1."""
        return [self.p - 1 if x == -1 else x
                for x in generate_custom_array(self.N, self.dr, self.dr, self.rng)]

    def encryptStr(self, s: str):
        return self.encryptBits(string_to_bits(s))["value"]

    def decryptStr(self, e):
        return bits_to_string(expand_array_to_multiple(self.decryptBits(e)["value"], 8))

    def encryptBits(self, m, r=None):
        """index.js:87-110; ``r`` is the injection seam the reference lacks."""
        if r is None:
            r = self.sample_r()
        rhq = self._mul(r, self.h, self.q)
        rhqm = add_polynomials(m, rhq, self.q)
        d = self._div_I(rhqm, self.I, self.q)
        return {
            "value": trim_polynomial(d["remainder"]),
            "inputs": {
                "r": list(r),
                "m": expand_array(m, self.N, 0),
                "h": expand_array(self.h, self.N, 0),
                "quotientE": expand_array([js_rem(x, self.q) for x in d["quotient"]], self.N + 1, 0),
                "remainderE": expand_array(d["remainder"], self.N + 1, 0),
            },
            "params": [self.q, self.calculateNq(), self.N],
        }

    def decryptBits(self, e):
        """index.js:111-140."""
        q, p = self.q, self.p
        f = [q - 1 if x == -1 else x for x in self.f]
        a = self._mul(f, e, q)
        a_div = self._div_I(a, self.I, q)
        a_div_p = [js_rem(x + 1, p) if x > q / 2 else js_rem(x, p) for x in a_div["remainder"]]
        c = self._mul(self.fp, a_div_p, p)
        c_div = self._div_I(c, self.I, p)
        return {
            "value": trim_polynomial(c_div["remainder"]),
            "inputs": {
                "f": expand_array(f, self.N, 0),
                "fp": expand_array(self.fp, self.N, 0),
                "e": expand_array(e, self.N, 0),
                "quotient1": expand_array(a_div["quotient"], self.N + 1, 0),
                "remainder1": expand_array(a_div["remainder"], self.N + 1, 0),
                "quotient2": expand_array(c_div["quotient"], self.N + 1, 0),
                "remainder2": expand_array(c_div["remainder"], self.N + 1, 0),
            },
            "params": [q, self.calculateNq(), p, self.calculateNp(), self.N],
        }

    def calculateNq(self):
        return math.ceil(math.log2(self.q * self.q * self.N))   # index.js:201-203

    def calculateNp(self):
        return math.ceil(math.log2(self.p * self.p * self.N))   # index.js:204-206


def sum_ciphertexts(es: Sequence[Sequence[int]], q: int) -> List[int]:
    """Left fold of addPolynomials (index.js:235-244; test/reference.test.js:58)."""
    acc: List[int] = [0]
    for e in es:
        acc = add_polynomials(acc, e, q)
    return acc


# ----------------------------------------------------------------------------
# Batched closed forms (numpy) -- what the CUDA kernels are compared against
# ----------------------------------------------------------------------------


def _lin_rows(a: np.ndarray, b: np.ndarray, N: int) -> np.ndarray:
    """Row-wise exact linear convolution, zero padded to 2N columns."""
    B = a.shape[0]
    out = np.zeros((B, 2 * N), dtype=np.int64)
    for i in range(B):
        out[i, : 2 * N - 1] = np.convolve(a[i].astype(np.int64), b[i].astype(np.int64))
    return out


def encrypt_batch(h: np.ndarray, r: np.ndarray, m: np.ndarray, q: int):
    """Fixed-length (un-trimmed) encrypt of B rows.  h: (N,) or (B,N); r,m: (B,N).

    Returns dict(value (B,N), quotientE (B,N+1), remainderE (B,N+1)) as int64.
    Closed form of index.js:87-110.
    """
    B, N = r.shape
    hh = np.broadcast_to(h, (B, N)) if h.ndim == 1 else h
    c = _lin_rows(r, hh, N)
    c[:, :N] += m.astype(np.int64)
    c = np.mod(c, q)
    rem = np.zeros((B, N + 1), dtype=np.int64)
    quo = np.zeros((B, N + 1), dtype=np.int64)
    rem[:, :N] = np.mod(c[:, :N] + c[:, N:], q)
    quo[:, :N] = np.mod(-c[:, N:], q)
    return {"value": rem[:, :N].copy(), "quotientE": quo, "remainderE": rem}


def lift_mod_p(x: np.ndarray, q: int, p: int = 3) -> np.ndarray:
    """index.js:117 -- ``x > q/2 ? (x+1)%p : x%p`` (strict >; NOT the centred lift when q%3==1)."""
    return np.where(x > q / 2, (x + 1) % p, x % p)


def decrypt_batch(f: np.ndarray, fp: np.ndarray, e: np.ndarray, q: int, p: int = 3):
    """Fixed-length decrypt of B rows.  f: (N,) or (B,N) in {-1,0,1}; fp likewise in [0,p); e: (B,N).

    Closed form of index.js:111-140.
    """
    B, N = e.shape
    ff = np.broadcast_to(f, (B, N)) if f.ndim == 1 else f
    fpp = np.broadcast_to(fp, (B, N)) if fp.ndim == 1 else fp
    a = np.mod(_lin_rows(ff, e, N), q)
    r1 = np.zeros((B, N + 1), dtype=np.int64)
    q1 = np.zeros((B, N + 1), dtype=np.int64)
    r1[:, :N] = np.mod(a[:, :N] + a[:, N:], q)
    q1[:, :N] = np.mod(-a[:, N:], q)
    b = lift_mod_p(r1[:, :N], q, p)
    c = np.mod(_lin_rows(fpp, b, N), p)
    r2 = np.zeros((B, N + 1), dtype=np.int64)
    q2 = np.zeros((B, N + 1), dtype=np.int64)
    r2[:, :N] = np.mod(c[:, :N] + c[:, N:], p)
    q2[:, :N] = np.mod(-c[:, N:], p)
    return {"value": r2[:, :N].copy(), "quotient1": q1, "remainder1": r1,
            "quotient2": q2, "remainder2": r2}


def sum_batch(e: np.ndarray, q: int) -> np.ndarray:
    """(B,N) -> (N,) column sums mod q (fold of addPolynomials, un-trimmed)."""
    return np.mod(e.astype(np.int64).sum(axis=0), q)


def sample_ternary_rows(B: int, N: int, ones: int, twos: int, rng: np.random.Generator,
                        neg_value: int = 2) -> np.ndarray:
    """B exact-weight ternary rows (``ones`` 1s, ``twos`` entries equal to neg_value)."""
    base = np.zeros(N, dtype=np.int64)
    base[:ones] = 1
    base[ones:ones + twos] = neg_value
    return np.stack([rng.permutation(base) for _ in range(B)])


# ----------------------------------------------------------------------------
# Circuit constraint checkers (circuits/ntru.circom:155-238)
# ----------------------------------------------------------------------------


def _verify_divide(a: np.ndarray, I: np.ndarray, quotient, remainder, p: int) -> bool:
    """VerifyDividePolynomials(p, np, Na=len(a), Nb=len(I)) -- ntru.circom:155-186."""
    nb = len(I)
    quotient = np.asarray(quotient, dtype=np.int64)
    remainder = np.asarray(remainder, dtype=np.int64)
    if len(quotient) != nb or len(remainder) != nb:
        return False
    if (quotient < 0).any() or (remainder < 0).any():
        return False
    prod = np.convolve(I.astype(np.int64), quotient)          # length 2*nb-1
    prod[:nb] += remainder
    prod = np.mod(prod, p)
    na = len(a)
    return bool(np.array_equal(prod[:na], a) and not prod[na:].any())


def verify_encrypt(inputs: dict, params: Sequence[int]) -> bool:
    """VerifyEncrypt(q, nq, N) -- ntru.circom:188-208."""
    q, _nq, N = params
    r = np.asarray(inputs["r"], dtype=np.int64)
    m = np.asarray(inputs["m"], dtype=np.int64)
    h = np.asarray(inputs["h"], dtype=np.int64)
    if not (len(r) == len(m) == len(h) == N):
        return False
    rhq = np.convolve(r, h)
    rhq[:N] += m
    rhq = np.mod(rhq, q)
    I = np.zeros(N + 1, dtype=np.int64)
    I[0], I[N] = 1, q - 1
    return _verify_divide(rhq, I, inputs["quotientE"], inputs["remainderE"], q)


def verify_decrypt(inputs: dict, params: Sequence[int]) -> bool:
    """VerifyDecrypt(q, nq, p, np, N) -- ntru.circom:210-238."""
    q, _nq, p, _np, N = params
    f = np.asarray(inputs["f"], dtype=np.int64)
    fp = np.asarray(inputs["fp"], dtype=np.int64)
    e = np.asarray(inputs["e"], dtype=np.int64)
    if not (len(f) == len(fp) == len(e) == N):
        return False
    a = np.mod(np.convolve(f, e), q)
    I = np.zeros(N + 1, dtype=np.int64)
    I[0], I[N] = 1, q - 1
    if not _verify_divide(a, I, inputs["quotient1"], inputs["remainder1"], q):
        return False
    r1 = np.asarray(inputs["remainder1"], dtype=np.int64)[:N]
    b = np.mod(r1 + (r1 > q / 2), p)                          # ntru.circom:227-231
    c = np.mod(np.convolve(fp, b), p)
    I[N] = p - 1
    return _verify_divide(c, I, inputs["quotient2"], inputs["remainder2"], p)


def pack_output(max_val: int, data_len: int, data: Sequence[int]) -> dict:
    """packOutput -- index.js:572-598: n = floor(252 / bits) coefficients per BN254 field element."""
    max_input_bits = math.floor(math.log2(max_val) + 1)
    per = 252 // max_input_bits
    arr_len = max(math.ceil(data_len / per) * per, per * 3)
    output_size = max(math.ceil(arr_len / per), 3)
    in_arr = expand_array(data, arr_len, 0)
    expected = [0] * output_size
    for i, v in enumerate(in_arr):
        expected[i // per] += int(v) << ((i % per) * max_input_bits)
    return {"maxInputBits": max_input_bits, "maxOutputBits": per * max_input_bits, "outputSize": output_size,
            "arrLen": arr_len, "expected": expected}


def unpack_input(max_val: int, packed_bits: int, data: Sequence[int]) -> dict:
    """unpackInput -- index.js:600-620."""
    max_input_bits = math.floor(math.log2(max_val) + 1)
    per = packed_bits // max_input_bits
    mask = (1 << max_input_bits) - 1
    unpacked = [0] * (per * len(data))
    for i, d in enumerate(data):
        for j in range(per):
            unpacked[i * per + j] = (int(d) >> (j * max_input_bits)) & mask
    return {"maxInputBits": max_input_bits, "packedBits": packed_bits, "packedSize": len(data),
            "unpackedSize": len(unpacked), "unpacked": trim_polynomial(unpacked)}


def combine_array(in_arr: Sequence[int], max_input_bits: int, per: int) -> List[int]:
    """CombineArray as the reference's own test computes its expectation (test/circuits.test.js:27-31)."""
    out = [0] * math.ceil(len(in_arr) / per)
    for i, cur in enumerate(in_arr):
        out[i // per] += int(cur) * (2 ** ((i % per) * max_input_bits))
    return out


def verify_inverse(inputs: dict, params: Sequence[int]) -> bool:
    """VerifyInverse(q, nq, N) -- ntru.circom:242-256 (q is the case's modulus: q for fq / h, p for fp)."""
    q, _nq, N = params
    f = np.asarray(inputs["f"], dtype=np.int64)
    fq = np.asarray(inputs["fq"], dtype=np.int64)
    if not (len(f) == len(fq) == N):
        return False
    a = np.mod(np.convolve(f, fq), q)
    I = np.zeros(N + 1, dtype=np.int64)
    I[0], I[N] = 1, q - 1
    return _verify_divide(a, I, inputs["quotientI"], inputs["remainderI"], q)


# ----------------------------------------------------------------------------
# Parameter sets of BASELINE.json
# ----------------------------------------------------------------------------

CONFIGS = {
    "default167": dict(N=167, q=128, p=3, df=61, dg=20, dr=18),
    "hps509": dict(N=509, q=2048, p=3, df=169, dg=169, dr=169),
    "hps677": dict(N=677, q=2048, p=3, df=225, dg=225, dr=225),
    "hps821": dict(N=821, q=4096, p=3, df=273, dg=273, dr=273),
    "hrss701": dict(N=701, q=8192, p=3, df=233, dg=233, dr=233),
    "tiny17": dict(N=17, q=32, p=3, df=3, dg=2, dr=2),
}


def make_key(config: str, seed: int, *, literal: bool = False) -> NTRU:
    """Seeded, valid key for a named parameter set (host keygen, index.js:30-79)."""
    rng = random.Random(seed)
    for _ in range(50):
        k = NTRU(dict(CONFIGS[config]), literal=literal, rng=rng)
        k.generatePrivateKeyF()
        k.generateNewPublicKeyGH()
        if k.key_is_valid():      # the reference's own check is vacuous
            return k
    raise RuntimeError("no valid key found")
