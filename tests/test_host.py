"""CPU-side checks of the product's host code and of the C ABI surface (no compute calls)."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

import ntru_circom_b200 as nb
import ntru_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ntru_circom_b200 import _lib, build
    build.build()
    hdr = open(os.path.join(ROOT, "include", "ntru_b200.h")).read()
    declared = set(re.findall(r"\b(ntru_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("ntru_ctx")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libntru_b200.so does not export {name}"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_abi_rejects_bad_parameters_without_a_gpu():
    from ntru_circom_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.ntru_create(ctypes.byref(h), 167, 5, 128, 0) == _lib.NTRU_E_PARAM      # p != 3
    assert lib.ntru_create(ctypes.byref(h), 167, 3, 100, 0) == _lib.NTRU_E_PARAM      # q not a power of two
    assert lib.ntru_create(ctypes.byref(h), 4, 3, 128, 0) == _lib.NTRU_E_PARAM        # N too small
    assert lib.ntru_create(ctypes.byref(h), 1024, 3, 16384, 0) == _lib.NTRU_E_PARAM   # fp32 exactness bound
    # q > 8192 is refused outright (the tcgen05 schedule's 5-bit high limb and its one-byte lift stop at q = 8192):
    # the two cases that used to be accepted and then decrypted wrongly on the tensor schedule
    assert lib.ntru_create(ctypes.byref(h), 167, 3, 16384, 0) == _lib.NTRU_E_PARAM
    assert lib.ntru_create(ctypes.byref(h), 128, 3, 32768, 0) == _lib.NTRU_E_PARAM
    assert lib.ntru_strerror(_lib.NTRU_E_NOKEY) == b"key not set"


def test_napi_shim_type_checks_and_binds_every_host_buffer_entry_point():
    """node and node_api.h are absent from this image, so the N-API shim cannot run; it is at least compiled
    (gcc -fsyntax-only, warnings as errors) against a stub of node_api.h with the real C-ABI header, and every
    host-buffer entry point of include/ntru_b200.h must be called from it (the *_dev ones need device pointers, which
    JavaScript does not have)."""
    import subprocess
    shim = os.path.join(ROOT, "bindings", "node", "ntru_napi.c")
    res = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "tests", "napi_stub"),
                          "-I", os.path.join(ROOT, "include"), shim], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    hdr = open(os.path.join(ROOT, "include", "ntru_b200.h")).read()
    declared = set(re.findall(r"\b(ntru_[a-z0-9_]+)\s*\(", hdr))
    src = open(shim).read()
    host_only = {n for n in declared if not n.endswith("_dev")} - {
        "ntru_ctx", "ntru_launch_count", "ntru_last_path", "ntru_timing_read", "ntru_timing_reset",      # bench instrumentation
        "ntru_stream", "ntru_set_stream", "ntru_sync", "ntru_host_alloc", "ntru_host_free"}             # device-side plumbing
    missing = sorted(n for n in host_only if not re.search(r"\b" + n + r"\s*\(", src))
    assert not missing, missing
    # the JS wrapper must not narrow messages to bytes (a coefficient >= 256 wrapped modulo 256 in round 1)
    js = open(os.path.join(ROOT, "bindings", "node", "index.mjs")).read()
    assert "Uint16Array.from(red)" in js and "Uint8Array.from(mExp" not in js


def test_wire_helpers_match_pack_output_and_unpack_input():
    """The numpy forms of the field-element wire format (ntru_circom_b200.wire) against the oracle's BigInt restatement
    of packOutput / unpackInput (index.js:572-620) at every bit width the engine uses, widths N and N + 1, including the
    three-element minimum."""
    rng = np.random.default_rng(5)
    for max_val, width in [(127, 167), (127, 168), (2047, 509), (2047, 510), (4095, 821), (8191, 702), (2, 509), (2, 510),
                           (2, 167), (2, 1025), (2047, 20), (1, 509)]:
        bits, n, elems = nb.wire.geometry(max_val, width)
        rows = rng.integers(0, max_val + 1, size=(5, width))
        rows[0] = max_val
        rows[1] = 0
        packed = nb.wire.pack_rows(max_val, rows)
        assert packed.shape == (5, elems, 8)
        for b in range(5):
            want = o.pack_output(max_val, width, [int(x) for x in rows[b]])
            assert want["outputSize"] == elems and want["maxInputBits"] == bits
            got = [sum(int(packed[b, e, w]) << (32 * w) for w in range(8)) for e in range(elems)]
            assert got == want["expected"], (max_val, width, b)
            un = o.unpack_input(max_val, n * bits, want["expected"])
            back = nb.wire.unpack_rows(max_val, width, packed[b:b + 1])[0]
            assert o.trim_polynomial([int(x) for x in back]) == un["unpacked"]
            assert np.array_equal(back, rows[b])
    with pytest.raises(ValueError):
        nb.wire.pack_rows(2, np.array([[3]]) + 1)                  # 4 does not fit two bits
    with pytest.raises(IndexError):
        nb.wire.unpack_rows(2047, 509, np.zeros((1, 23, 8), dtype=np.uint32))


def test_shipped_library_reads_no_debug_environment_switches():
    """The timing experiments (NTRU_DEBUG_*: skip loads / stores, results wrong on purpose) exist in NTRU_TRACE builds
    only; a stray environment variable cannot change what the shipped library computes."""
    from ntru_circom_b200 import _lib
    blob = open(_lib.LIB_PATH, "rb").read()
    assert b"NTRU_DEBUG_" not in blob and b"NTRU_TRACE_" not in blob


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ntru-circom_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "ntru_oracle" not in txt and "c_oracle" not in txt and "libntru_oracle" not in txt, fn


def test_poly_matches_oracle_kats():
    assert nb.multiplyPolynomials([1, 2, 3, 4], [6, 5, 4, 3], 2 ** 20) == [6, 17, 32, 50, 38, 25, 12]
    assert nb.multiplyPolynomials([], [1], 5) == [0]
    assert nb.dividePolynomials([1, 2], [2, 3], 8) == {"quotient": [6], "remainder": [5]}
    assert nb.dividePolynomials([81, 2, 96], [48, 2, 31], 128) == {"quotient": [32], "remainder": [81, 66]}
    assert nb.dividePolynomials([81, 2, 96], [48, 2, 31], 16) == {"quotient": [0], "remainder": [1, 2]}
    with pytest.raises(ValueError, match="No inverse"):
        nb.dividePolynomials([1, 2], [2, 3], 3)
    assert nb.extendedEuclideanAlgorithm([4, 2, 0, 3], [3, 2, 1], 11)["inverse"] == [5, 8]
    assert nb.modInverse(-1, 2048) == 2047 and nb.modInverse(2, 4) is None
    rng = random.Random(3)
    for _ in range(20):
        p = rng.choice([3, 7, 32, 2048])
        a = [rng.randrange(-3, p) for _ in range(rng.randrange(1, 12))]
        b = [rng.randrange(-3, p) for _ in range(rng.randrange(1, 12))]
        assert nb.multiplyPolynomials(a, b, p) == o.multiply_polynomials_fft(a, b, p)
        assert nb.addPolynomials(a, b, p) == o.add_polynomials(a, b, p)
        assert nb.subtractPolynomials(a, b, p) == o.subtract_polynomials(a, b, p)
        if p in (3, 7):
            b2 = b if o.degree([x % p for x in b]) >= 0 else [1]
            aa, bb = [x % p for x in a], [x % p for x in b2]
            assert nb.dividePolynomials(aa, bb, p) == o.divide_polynomials(aa, bb, p)


def test_format_helpers():
    assert nb.stringToBits("Hi") == o.string_to_bits("Hi")
    assert nb.bitsToString(nb.stringToBits("Hello World")) == "Hello World"
    assert nb.expandArrayToMultiple([1, 1, 1], 8) == [1, 1, 1, 0, 0, 0, 0, 0]
    with pytest.raises(IndexError):
        nb.expandArray([1, 2, 3], 2)
    assert nb.trimPolynomial([0, 0]) == [0] and nb.trimPolynomial([1, 0, 2, 0]) == [1, 0, 2]
    # packOutput / unpackInput round trip (index.js:572-620; test/circuits.test.js:20-58)
    data = [random.Random(1).randrange(8192) for _ in range(701)]
    packed = nb.packOutput(8191, 701, data)
    assert packed["maxInputBits"] == 13 and packed["outputSize"] == 37
    un = nb.unpackInput(8191, packed["maxOutputBits"], packed["expected"])
    assert un["unpacked"] == nb.trimPolynomial(data)
    assert nb.bitsToBigInt(list(reversed(nb.bigintToBits(1234567)))) == 1234567


@pytest.mark.parametrize("cfg", ["tiny17", "default167", "hps509"])
def test_host_keygen_matches_oracle(cfg):
    """Same seeded draws -> same f, fq, fp, g, h as the oracle's restatement of index.js:30-79."""
    ref = o.NTRU(dict(o.CONFIGS[cfg]), rng=random.Random(42))
    ref.generatePrivateKeyF()
    ref.generateNewPublicKeyGH()
    rng = random.Random(42)
    mine = nb.NTRU(dict(o.CONFIGS[cfg]), rand32=lambda: rng.getrandbits(32))
    mine.generatePrivateKeyF()
    mine.generateNewPublicKeyGH()
    assert (mine.f, mine.fq, mine.fp, mine.g, mine.h) == (ref.f, ref.fq, ref.fp, ref.g, ref.h)
    assert mine.calculateNq() == ref.calculateNq() and mine.calculateNp() == ref.calculateNp()
    # the oracle's verifyKeysInputs (index.js:141-197) satisfies the VerifyInverse constraints (ntru.circom:242-256)
    w = ref.verifyKeysInputs()
    assert w["fq"]["inputs"]["remainderI"] == [1] + [0] * ref.N
    assert w["fp"]["inputs"]["remainderI"] == [1] + [0] * ref.N
    assert w["h"]["inputs"]["remainderI"][: len(ref.h)] == ref.h
    for case in ("fq", "fp", "h"):
        assert o.verify_inverse(w[case]["inputs"], w[case]["params"]), case
        bad = {k: list(v) for k, v in w[case]["inputs"].items()}
        bad["remainderI"][0] = (bad["remainderI"][0] + 1) % w[case]["params"][0]
        assert not o.verify_inverse(bad, w[case]["params"]), case
    # the product's verifyKeysInputs runs its three products on the GPU: without a device it must fail loudly
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(nb.NtruError):
            mine.verifyKeysInputs()


def test_constructor_defaults_and_errors():
    k = nb.NTRU()
    assert (k.N, k.p, k.q, k.df, k.dg, k.dr) == (167, 3, 128, 61, 20, 18)
    assert k.I == [1] + [0] * 166 + [-1]
    with pytest.raises(TypeError):
        k.encryptBits([1, 0, 1])          # h === null -> TypeError in the reference (index.js:90)
    with pytest.raises(TypeError):
        k.decryptBits([1, 2, 3])
    with pytest.raises(ValueError, match="missing private key F"):
        k.generatePublicKeyH()
    with pytest.raises(ValueError):
        nb.generateCustomArray(5, 3, 3)


class _Draws:
    """rng stand-in for the oracle's generate_custom_array: hands out a fixed list of 32-bit draws."""

    def __init__(self, words):
        self.it = iter(words)

    def getrandbits(self, n):
        assert n == 32
        return next(self.it)


def test_host_copy_of_the_device_generator_is_chacha20():
    """The device draws r from ChaCha20 (csrc/generic_kernels.cu); its host copy is pinned on the block-function
    vector of RFC 8439 section 2.3.2 (key 00..1f, counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00 -- in the
    64-bit counter / 64-bit nonce layout used here: counter = 1 | 0x09000000 << 32, nonce = 0x4a000000)."""
    ks = nb.chacha20_block(bytes(range(32)), 1 | (0x09000000 << 32), 0x4A000000)
    assert ks == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
                  0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    # a row's draws are the keystream words of blocks 0, 1, ... under nonce = row number
    d = nb.sampler_draws(nb.seed_key(5), 7, 40)
    assert d[:16] == nb.chacha20_block(nb.seed_key(5), 0, 7) and d[16:32] == nb.chacha20_block(nb.seed_key(5), 1, 7)
    assert d != nb.sampler_draws(nb.seed_key(5), 8, 40) and d != nb.sampler_draws(nb.seed_key(6), 7, 40)
    # the product's generateCustomArray and the oracle's consume draws identically (index.js:476-485)
    for N, dr in ((17, 3), (167, 18), (509, 169)):
        words = nb.sampler_draws(nb.seed_key(1), 3, N - 1)
        it = iter(words)
        assert nb.generateCustomArray(N, dr, dr, rand32=lambda: next(it)) == o.generate_custom_array(N, dr, dr, _Draws(words))


def test_seeds_of_the_gpu_class_tests_decrypt_in_the_oracle(golden):
    """tests/test_gpu_parity.py pins the draws of every test that lets ``class NTRU`` sample its own randomness
    (NTRU decryption fails for unlucky r: ~1 in 3000 at the default parameters, ~1 in 20 for a two-ciphertext sum).
    The GPU path is bit-exact with the oracle, so the same seeds replayed here prove those tests deterministic."""
    g = golden("default167")
    k = o.NTRU(dict(o.CONFIGS["default167"], f=g["f"].tolist(), fp=o.trim_polynomial(g["fp"].tolist()),
                    h=o.trim_polynomial(g["h"].tolist())), rng=random.Random(1))
    e1 = k.encryptBits([1, 2, 1, 0, 1])["value"]
    e2 = k.encryptBits([0, 1, 1, 1, 0, 1, 0, 1])["value"]
    assert k.decryptBits(o.sum_ciphertexts([e1, e2], k.q))["value"] == [1, 0, 2, 1, 1, 1, 0, 1]
    k = o.NTRU(rng=random.Random(21))
    k.generatePrivateKeyF()
    k.generateNewPublicKeyGH()
    e = k.encryptStr("Hello World")
    assert k.decryptStr(e) == "Hello World"
    other = o.NTRU(rng=random.Random(22))
    other.generatePrivateKeyF()
    assert other.decryptStr(e) != "Hello World"
    k = o.make_key("default167", 300)
    k.rng = random.Random(11)
    for t in ["Hello World", "", "a", "x" * 20, "NTRU on B200!", "\xe9\xff"]:
        assert k.decryptStr(k.encryptStr(t)) == (t if t else "\x00")
    for t in ["The quick brown fox jumps over the lazy dog. " * 3, "short", "z" * 40]:
        for i in range(0, len(t), 20):
            assert k.decryptStr(k.encryptStr(t[i:i + 20])) == t[i:i + 20]
    d = 701 // 3
    k = o.NTRU({"N": 701, "q": 8192, "df": d, "dg": d, "dr": d}, rng=random.Random(31))
    k.generatePrivateKeyF()
    k.generateNewPublicKeyGH()
    assert k.decryptStr(k.encryptStr("Big polys")) == "Big polys"


def test_packed_mod3_identity_of_the_imma_output_passes():
    """`mod3_2x16` (csrc/imma_kernels.cu): residues mod 3 of two packed 16-bit coefficients from two `umulhi` by 0x5556.
    The kernels feed it lanes below 2^14 (sums of at most two product coefficients <= 4 N, or remainder1 + 1 <= q <= 8192);
    the identity is checked here for every low lane against a spread of high lanes, including the extremes."""
    lo = np.arange(0, 1 << 14, dtype=np.uint64)
    his = sorted(set(list(range(0, 64)) + list(range((1 << 14) - 64, 1 << 14)) + list(range(0, 1 << 14, 37))))
    m32 = np.uint64(0xFFFFFFFF)
    for hi in his:
        v = (np.uint64(hi) << np.uint64(16)) | lo
        qh = (v * np.uint64(0x5556)) >> np.uint64(32)
        ql = (((v << np.uint64(16)) & m32) * np.uint64(0x5556)) >> np.uint64(32)
        r = (v - np.uint64(3) * (ql + (qh << np.uint64(16)))) & m32
        assert np.array_equal(r & np.uint64(0xFFFF), lo % np.uint64(3)), hi
        assert np.all((r >> np.uint64(16)) == np.uint64(hi % 3)), hi
