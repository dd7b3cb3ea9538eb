"""Pins the oracle (oracle/ntru_oracle.py, oracle/ntru_ref_port.c) against every upstream KAT of the path."""
import math
import random

import numpy as np
import pytest

import c_oracle
import ntru_oracle as o

MUL_FNS = [o.multiply_polynomials_fft, o.multiply_polynomials_exact, lambda a, b, p: c_oracle.multiply(a, b, p)[0]]
DIV_FNS = [o.divide_polynomials, c_oracle.divide]


@pytest.mark.parametrize("mul", MUL_FNS)
def test_multiply_worked_example(mul):
    # circuits/ntru.circom:33-71
    assert mul([1, 2, 3, 4], [6, 5, 4, 3], 2 ** 20) == [6, 17, 32, 50, 38, 25, 12]


@pytest.mark.parametrize("mul", MUL_FNS)
@pytest.mark.parametrize("a,b,p", [
    ([1, 4], [0, 3], 7), ([1, 2, 3], [4, 3, 2], 7), ([1, 2, 3, 4], [5, 4, 3, 2], 11),
    ([1, 2, 3, 4, 5], [6, 5, 4, 3, 2], 13), ([1, 2, 3, 4, 5, 0], [7, 6, 5, 4, 3, 2], 13)])
def test_multiply_vector_pairs(mul, a, b, p):
    # test/circuits.test.js:61-65 -- the schoolbook circuit must equal multiplyPolynomials
    school = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            school[i + j] += x * y
    assert mul(a, b, 2 ** 20) == o.trim_polynomial(school)
    assert mul(a, b, p) == o.trim_polynomial([v % p for v in school])


@pytest.mark.parametrize("div", DIV_FNS)
def test_division_cases(div):
    # test/circuits.test.js:166-169
    assert div([1, 2], [2, 3], 8) == {"quotient": [6], "remainder": [5]}
    with pytest.raises(ValueError, match="No inverse"):
        div([1, 2], [2, 3], 3)
    assert div([81, 2, 96], [48, 2, 31], 128) == {"quotient": [32], "remainder": [81, 66]}
    assert div([81, 2, 96], [48, 2, 31], 16) == {"quotient": [0], "remainder": [1, 2]}
    with pytest.raises(ValueError, match="zero polynomial"):
        div([1, 2], [0, 0], 7)


def test_multiply_empty_and_negative():
    assert o.multiply_polynomials_fft([], [1], 7) == [0]                 # index.js:320
    assert o.multiply_polynomials_fft([1, -1], [1, 1], 5) == [1, 0, 4]    # negatives (g has -1)
    assert o.multiply_polynomials_exact([1, -1], [1, 1], 5) == [1, 0, 4]


def test_eea_worked_example():
    # index.js:408-423
    assert o.extended_euclidean_algorithm([4, 2, 0, 3], [3, 2, 1], 11)["inverse"] == [5, 8]


def test_string_codec():
    bits = o.string_to_bits("Hi")
    assert bits == [0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0, 1, 0, 0, 1]
    assert o.bits_to_string(bits) == "Hi"
    with pytest.raises(IndexError):
        o.expand_array([1, 2, 3], 2)


def test_literal_equals_fast_default_params():
    lit = o.make_key("default167", 11, literal=True)
    fast = o.make_key("default167", 11, literal=False)
    assert (lit.f, lit.fp, lit.fq, lit.g, lit.h) == (fast.f, fast.fp, fast.fq, fast.g, fast.h)
    m = o.string_to_bits("Hello World")
    r = lit.sample_r()
    a, b = lit.encryptBits(m, r), fast.encryptBits(m, r)
    assert a == b
    assert lit.decryptBits(a["value"]) == fast.decryptBits(a["value"])
    assert o.fft_margin["max"] < 1e-3


def test_reference_roundtrips():
    # test/reference.test.js:6-13 and :15-25
    k = o.make_key("default167", 5)
    assert k.decryptStr(k.encryptStr("Hello World")) == "Hello World"
    other = o.make_key("default167", 6)
    assert other.decryptStr(k.encryptStr("Hello World")) != "Hello World"


def test_reference_roundtrip_large():
    # test/reference.test.js:27-44 (GO_LARGE)
    k = o.make_key("hrss701", 5)
    assert k.decryptStr(k.encryptStr("Big polys")) == "Big polys"


def test_additive_homomorphism():
    # test/reference.test.js:48-61
    k = o.make_key("default167", 9)
    e1 = k.encryptBits([1, 2, 1, 0, 1])["value"]
    e2 = k.encryptBits([0, 1, 1, 1, 0, 1, 0, 1])["value"]
    s = o.add_polynomials(e1, e2, k.q)
    assert k.decryptBits(s)["value"] == [1, 0, 2, 1, 1, 1, 0, 1]
    assert o.sum_ciphertexts([e1, e2], k.q) == s


@pytest.mark.parametrize("cfg", ["tiny17", "default167", "hrss701"])
def test_witness_satisfies_circuits(cfg):
    # test/circuits.test.js:252-302 with the constraint systems of circuits/ntru.circom:155-238
    k = o.make_key(cfg, 3, literal=cfg == "tiny17")
    m = [1, 0] * 8 + [1]
    enc = k.encryptBits(m)
    assert o.verify_encrypt(enc["inputs"], enc["params"])
    dec = k.decryptBits(enc["value"])
    assert o.verify_decrypt(dec["inputs"], dec["params"])
    assert dec["value"] == m
    dec["inputs"]["remainder2"][0] += 1                      # tamper: circuits.test.js:296-301
    assert not o.verify_decrypt(dec["inputs"], dec["params"])
    enc["inputs"]["quotientE"][0] = (enc["inputs"]["quotientE"][0] + 1) % k.q
    assert not o.verify_encrypt(enc["inputs"], enc["params"])


def test_q4096_reference_lift_is_not_the_centred_lift():
    # SURVEY parity trap #1: the reference decrypts wrongly when q % 3 == 1; the oracle must reproduce that
    k = o.make_key("hps821", 2)
    m = [1, 0, 1, 1, 0, 1]
    dec = k.decryptBits(k.encryptBits(m)["value"])
    assert dec["value"] != m
    assert o.verify_decrypt(dec["inputs"], dec["params"])    # ... yet the witness satisfies the circuit


@pytest.mark.parametrize("cfg", ["tiny17", "default167", "hps509", "hps677", "hps821", "hrss701"])
def test_c_oracle_equals_python_oracle_and_golden(cfg, golden):
    g = golden(cfg)
    q, p = int(g["q"]), int(g["p"])
    enc_c = c_oracle.encrypt_batch(g["h"], g["r"], g["m"], q)
    enc_p = o.encrypt_batch(g["h"].astype(np.int64), g["r"], g["m"], q)
    for k in ("value", "quotientE", "remainderE"):
        assert np.array_equal(enc_c[k], g[k]) and np.array_equal(enc_p[k], g[k]), k
    dec_c = c_oracle.decrypt_batch(g["f"], g["fp"], g["value"], q, p)
    dec_p = o.decrypt_batch(g["f"].astype(np.int64), g["fp"].astype(np.int64), g["value"], q, p)
    for k, gk in (("value", "dec_value"), ("quotient1",) * 2, ("remainder1",) * 2, ("quotient2",) * 2,
                  ("remainder2",) * 2):
        assert np.array_equal(dec_c[k], g[gk]) and np.array_equal(dec_p[k], g[gk]), k
    assert max(enc_c["fft_margin"], dec_c["fft_margin"]) < 1e-3   # FFT+round == exact integers
    assert np.array_equal(c_oracle.sum_batch(g["value"], q), g["sum"])
    assert np.array_equal(o.sum_batch(g["value"], q), g["sum"])


def test_generate_custom_array_weights():
    rng = random.Random(1)
    a = o.generate_custom_array(167, 61, 60, rng)
    assert a.count(1) == 61 and a.count(-1) == 60 and len(a) == 167
    with pytest.raises(ValueError):
        o.generate_custom_array(5, 3, 3, rng)


def test_pack_output_matches_upstream_combine_array_formula():
    """test/circuits.test.js:20-58: CombineArray / UnpackArray at maxVal = 8192, N = 701 -- the expectation formula of
    the upstream test pins pack_output (for in-range data and the same bit width) and unpack_input inverts it."""
    rng = random.Random(5)
    max_val, N = 8192, 701
    bits_up = int(math.log2(max_val))                      # the upstream test packs 13-bit values (values < 8192)
    per_up = 252 // bits_up
    arr_len = math.ceil(N / per_up) * per_up
    in_arr = [rng.randrange(max_val) for _ in range(arr_len)]
    want = o.combine_array(in_arr, bits_up, per_up)
    got = o.pack_output(max_val - 1, arr_len, in_arr)      # bit length of 8191 = 13: the same packing
    assert got["maxInputBits"] == bits_up and got["expected"] == want
    assert o.unpack_input(max_val - 1, got["maxOutputBits"], want)["unpacked"] == o.trim_polynomial(in_arr)
    # packOutput(q, ...) itself uses floor(log2(q) + 1) = 14 bits and at least three outputs (index.js:573-580)
    p = o.pack_output(8192, 5, [1, 2, 3, 4, 5])
    assert (p["maxInputBits"], p["outputSize"], p["arrLen"]) == (14, 3, 54)
    assert p["expected"] == [1 + (2 << 14) + (3 << 28) + (4 << 42) + (5 << 56), 0, 0]
