"""Parity of the CUDA engine (through the C ABI) against the oracle.  Bit-exact: integer work."""
import random

import numpy as np
import pytest

import ntru_oracle as o

pytestmark = pytest.mark.gpu

CFGS = ["tiny17", "default167", "hps509", "hps677", "hps821", "hrss701"]
ENC_KEYS = ("value", "quotientE", "remainderE")
DEC_KEYS = ("value", "quotient1", "remainder1", "quotient2", "remainder2")


@pytest.fixture(scope="module")
def nb():
    import ntru_circom_b200 as nb
    return nb


@pytest.fixture(scope="module")
def engines(nb, golden):
    cache = {}

    def get(cfg):
        if cfg not in cache:
            g = golden(cfg)
            eng = nb.Engine(int(g["N"]), int(g["p"]), int(g["q"]), 0)
            eng.set_public_key(g["h"])
            eng.set_private_key(g["f"], g["fp"])
            try:                      # tests mix torch device tensors with *_dev calls: one stream orders them
                import torch
                eng.set_stream(torch.cuda.current_stream().cuda_stream)
            except ImportError:
                pass
            cache[cfg] = eng
        return cache[cfg]
    yield get
    for e in cache.values():
        e.close()


def _paths(nb, eng):
    """Schedules to test: the fp32 CUDA-core one always, the two tensor schedules (register-fragment IMMA, tcgen05)
    when the device supports them."""
    out = [nb.PATH_CUDA_CORE]
    for path in (nb.PATH_IMMA, nb.PATH_TENSOR):
        try:
            eng.set_path(path)
            g = np.zeros((1, eng.N), dtype=np.uint8)
            eng.encrypt_batch(g, g, witness=False)
            out.append(path)
        except nb.NtruError:
            pass
    eng.set_path(nb.PATH_AUTO)
    return out


def seeded(seed):
    """rand32 replacement for crypto.getRandomValues: NTRU decryption can fail for unlucky r (1 in 3000 at the default
    parameters, 1 in 20 for a sum of two ciphertexts -- "this test may fail" upstream), so every test that lets the class
    draw its own randomness pins the draws; the seeds are checked against the oracle in test_host.py."""
    import random
    rng = random.Random(seed)
    return lambda: rng.getrandbits(32)


@pytest.mark.parametrize("cfg", CFGS)
def test_golden_vectors_both_schedules(cfg, nb, engines, golden):
    g, eng = golden(cfg), engines(cfg)
    for path in _paths(nb, eng):
        eng.set_path(path)
        enc = eng.encrypt_batch(g["r"], g["m"])
        assert eng.last_path == path
        for k in ENC_KEYS:
            assert np.array_equal(enc[k], g[k]), (cfg, path, k)
        dec = eng.decrypt_batch(g["value"])
        for k, gk in zip(DEC_KEYS, ("dec_value",) + DEC_KEYS[1:]):
            assert np.array_equal(dec[k], g[gk]), (cfg, path, k)
        assert np.array_equal(eng.sum(g["value"]), g["sum"])
    eng.set_path(nb.PATH_AUTO)


@pytest.mark.parametrize("cfg", CFGS)
def test_random_batch_vs_oracle(cfg, nb, engines, golden):
    """Seeded random batch with a ragged size (not a multiple of any tile), all schedules."""
    g, eng = golden(cfg), engines(cfg)
    N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
    B = 301 if N > 200 else 1000
    rng = np.random.default_rng(123)
    r = o.sample_ternary_rows(B, N, dr, dr, rng).astype(np.uint8)
    m = rng.integers(0, 2, size=(B, N)).astype(np.uint8)
    m[5] = rng.integers(0, 3, size=N)            # ternary plaintext (test/reference.test.js:50)
    m[0] = 0
    r[1] = 0                                    # degenerate randomness: e == m
    m[2] = 255                                  # largest byte message coefficients
    m[3, N // 3:] = 0
    want_e = o.encrypt_batch(g["h"].astype(np.int64), r, m, q)
    want_d = o.decrypt_batch(g["f"].astype(np.int64), g["fp"].astype(np.int64), want_e["value"], q, p)
    for path in _paths(nb, eng):
        eng.set_path(path)
        enc = eng.encrypt_batch(r, m)
        for k in ENC_KEYS:
            assert np.array_equal(enc[k], want_e[k]), (cfg, path, k)
        dec = eng.decrypt_batch(enc["value"])
        for k in DEC_KEYS:
            assert np.array_equal(dec[k], want_d[k]), (cfg, path, k)
        # value-only mode produces the same ciphertext / plaintext
        assert np.array_equal(eng.encrypt_batch(r, m, witness=False)["value"], want_e["value"])
        assert np.array_equal(eng.decrypt_batch(enc["value"], witness=False)["value"], want_d["value"])
    eng.set_path(nb.PATH_AUTO)
    if q % 3 == 2 and cfg != "tiny17":          # the reference's lift is the true centred lift only then;
        assert np.array_equal(want_d["value"][6:], m[6:])   # tiny17 (q=32) has genuine decryption failures


@pytest.mark.parametrize("cfg", CFGS)
def test_distinct_keys(cfg, nb, engines, golden):
    """Config 3: a different key per row, full witness; valid keys + random ones (incl. extreme coefficients), on the
    register-fragment IMMA schedule (the default for distinct keys) and on the fp32 CUDA-core schedule."""
    g, eng = golden(cfg), engines(cfg)
    N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
    nvalid = 8
    keys = [o.make_key(cfg, 100 + i) for i in range(nvalid)]
    keys += [keys[-1]] * (8 - nvalid)
    rng = np.random.default_rng(5)
    B = 8 + 131
    h = np.zeros((B, N), dtype=np.int64)
    f = np.zeros((B, N), dtype=np.int64)
    fp = np.zeros((B, N), dtype=np.int64)
    for i, k in enumerate(keys):
        h[i], f[i], fp[i] = o.expand_array(k.h, N), k.f, o.expand_array(k.fp, N)
    h[8:] = rng.integers(0, q, size=(B - 8, N))
    f[8:] = o.sample_ternary_rows(B - 8, N, dr, dr - 1, rng, neg_value=-1)
    fp[8:] = rng.integers(0, p, size=(B - 8, N))
    h[9], f[9], fp[9] = q - 1, 1, p - 1                      # largest possible accumulators
    f[10] = -1
    h[11], f[11], fp[11] = 0, 0, 0
    r = o.sample_ternary_rows(B, N, dr, dr, rng)
    r[9] = 2
    m = rng.integers(0, 2, size=(B, N))
    m[9] = 255
    want_e = o.encrypt_batch(h, r, m, q)
    want_d = o.decrypt_batch(f, fp, want_e["value"], q, p)
    for path in (nb.PATH_AUTO, nb.PATH_IMMA, nb.PATH_CUDA_CORE):
        eng.set_path(path)
        enc = eng.encrypt_batch(r.astype(np.uint8), m.astype(np.uint8), h=h.astype(np.uint16))
        assert eng.last_path == (nb.PATH_CUDA_CORE if path == nb.PATH_CUDA_CORE else nb.PATH_IMMA)
        for k in ENC_KEYS:
            assert np.array_equal(enc[k], want_e[k]), (path, k)
        dec = eng.decrypt_batch(enc["value"], f=f.astype(np.int8), fp=fp.astype(np.uint8))
        for k in DEC_KEYS:
            assert np.array_equal(dec[k], want_d[k]), (path, k)
        assert np.array_equal(eng.encrypt_batch(r.astype(np.uint8), m.astype(np.uint8), h=h.astype(np.uint16), witness=False)["value"],
                              want_e["value"])
        assert np.array_equal(eng.decrypt_batch(enc["value"], f=f.astype(np.int8), fp=fp.astype(np.uint8), witness=False)["value"],
                              want_d["value"])
    eng.set_path(nb.PATH_AUTO)
    if q % 3 == 2 and cfg != "tiny17":
        assert np.array_equal(want_d["value"][:nvalid], m[:nvalid])   # valid keys round-trip


@pytest.mark.parametrize("cfg", CFGS)
def test_verify_keys_inputs_vs_oracle(cfg, nb, engines, golden):
    """verifyKeysInputs (index.js:141-197).  The class API against the oracle's restatement field by field
    (and the VerifyInverse constraint checker, ntru.circom:242-256), then the batched entry point on valid keys plus
    random / extreme operands against the oracle's multiply + divide."""
    g, eng = golden(cfg), engines(cfg)
    N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
    keys = [o.make_key(cfg, 200 + i) for i in range(3)]
    k = keys[0]
    mine = nb.NTRU(dict(o.CONFIGS[cfg]), f=list(k.f), fq=list(k.fq), fp=list(k.fp), g=list(k.g), h=list(k.h))
    got, want = mine.verifyKeysInputs(), k.verifyKeysInputs()
    assert got == want
    for case in ("fq", "fp", "h"):
        assert o.verify_inverse(got[case]["inputs"], got[case]["params"]), case
    rng = np.random.default_rng(17)
    B = 3 + 70
    f = np.zeros((B, N), dtype=np.int64); fq = np.zeros((B, N), dtype=np.int64)
    fp = np.zeros((B, N), dtype=np.int64); gg = np.zeros((B, N), dtype=np.int64)
    for i, kk in enumerate(keys):
        f[i], fq[i], fp[i], gg[i] = kk.f, o.expand_array(kk.fq, N, 0), o.expand_array(kk.fp, N, 0), kk.g
    f[3:] = rng.integers(-1, 2, size=(B - 3, N)); gg[3:] = rng.integers(-1, 2, size=(B - 3, N))
    fq[3:] = rng.integers(0, q, size=(B - 3, N)); fp[3:] = rng.integers(0, p, size=(B - 3, N))
    f[4], gg[4], fq[4], fp[4] = -1, -1, q - 1, p - 1         # most negative products
    f[5], gg[5], fq[5], fp[5] = 1, 1, q - 1, p - 1           # largest products (3 (q-1) un-reduced in the h case)
    out = eng.verify_keys_batch(f.astype(np.int8), fq.astype(np.uint16), fp.astype(np.uint8), gg.astype(np.int8))
    for i in list(range(8)) + [B - 1]:
        for case, a, b, mod in (("fq", fq[i], f[i], q), ("fp", fp[i], f[i], p), ("h", fq[i] * p, gg[i], q)):
            # the reference multiplies by the witness value of a -1 (q-1 / p-1, index.js:151-156)
            d = o.divide_by_I_closed(o.multiply_polynomials_exact(list(map(int, a)), [mod - 1 if x == -1 else int(x) for x in b], mod), N, mod)
            assert out["quotient_" + case][i].tolist() == o.expand_array(d["quotient"], N + 1, 0), (cfg, i, case)
            assert out["remainder_" + case][i].tolist() == o.expand_array(d["remainder"], N + 1, 0), (cfg, i, case)


@pytest.mark.parametrize("cfg", ["default167", "hps509", "hrss701"])
def test_pack_output_unpack_input_vs_oracle(cfg, nb, engines, golden):
    """packOutput / unpackInput (index.js:572-620) on device rows: ciphertext rows -> BN254 field
    elements and back, against the oracle's BigInt restatement, for mod-q rows and for byte rows (mod-p outputs)."""
    torch = pytest.importorskip("torch")
    g, eng = golden(cfg), engines(cfg)
    N, q, P = int(g["N"]), int(g["q"]), eng.pitch
    rng = np.random.default_rng(23)
    B = 37
    for max_val, elem, hi in ((q, 2, q), (q - 1, 2, q), (2, 1, 3), (255, 1, 256)):
        rows = rng.integers(0, hi, size=(B, N))
        rows[0] = hi - 1
        rows[1] = 0
        dt = torch.int16 if elem == 2 else torch.uint8
        d = torch.zeros((B, P), dtype=dt, device="cuda")
        d[:, :N] = torch.from_numpy(rows.astype(np.int16 if elem == 2 else np.uint8)).cuda()
        bits, per, arr_len, outs = eng.pack_geometry(max_val, N)
        packed = torch.empty((B, outs, 4), dtype=torch.int64, device="cuda")
        eng.pack_output_dev(B, d, elem, N, P, max_val, packed)
        eng.sync()
        words = packed.cpu().numpy().astype(np.uint64)
        for b in (0, 1, 2, B - 1):
            want = o.pack_output(max_val, N, [int(x) for x in rows[b]])
            assert (bits, outs, arr_len) == (want["maxInputBits"], want["outputSize"], want["arrLen"])
            got = [sum(int(words[b, e, w]) << (64 * w) for w in range(4)) for e in range(outs)]
            assert got == want["expected"], (cfg, max_val, b)
        # and back: unpackInput(maxVal, maxOutputBits, packed) == the rows (un-trimmed, zero-extended to per * outs)
        un = torch.full((B, per * outs + 16), -1 if elem == 2 else 255, dtype=dt, device="cuda")
        eng.unpack_input_dev(B, packed, outs, max_val, per * bits, un, elem, un.shape[1])
        eng.sync()
        u = un.cpu().numpy().astype(np.int64) & (0xFFFF if elem == 2 else 0xFF)
        assert np.array_equal(u[:, :N], rows) and not u[:, N:per * outs].any()
        want_u = o.unpack_input(max_val, per * bits, o.pack_output(max_val, N, [int(x) for x in rows[2]])["expected"])
        assert o.trim_polynomial(u[2, :per * outs].tolist()) == want_u["unpacked"]


GENERAL_NQ = [(8, 4), (16, 32), (33, 64), (100, 256), (191, 512), (193, 1024), (255, 2048), (256, 2048), (257, 4096),
              (511, 8192), (512, 2048), (513, 2048), (640, 2048), (641, 4096), (704, 8192), (705, 2048), (832, 4096),
              (833, 2048), (768, 2048), (769, 4096), (896, 8192), (897, 2048), (1024, 8192)]


@pytest.mark.parametrize("N,q", GENERAL_NQ)
def test_parameter_sets_outside_baseline(N, q, nb):
    """The engine accepts any 8 <= N <= 1024 and power-of-two q <= 8192: sizes at the edges of every tiling decision
    (K atoms of 128, accumulator chunks, resident / streaming A operand at N = 640 / 641, IMMA buckets at 192 / 512 /
    704 / 832, fall-back above 832, even N) with random operands, every schedule, same key and distinct keys."""
    p = 3
    rng = np.random.default_rng(1000 * N + q)
    B = 70
    h = rng.integers(0, q, size=N)
    f = rng.integers(-1, 2, size=N)
    fp = rng.integers(0, p, size=N)
    r = rng.integers(0, 3, size=(B, N))
    m = rng.integers(0, 2, size=(B, N))
    r[0], m[0] = 2, 255 if q > 256 else 1                      # largest accumulators
    want_e = o.encrypt_batch(h, r, m, q)
    want_d = o.decrypt_batch(f, fp, want_e["value"], q, p)
    eng = nb.Engine(N, p, q, 0)
    eng.set_public_key(h.astype(np.uint16))
    eng.set_private_key(f.astype(np.int8), fp.astype(np.uint8))
    paths = _paths(nb, eng)
    assert nb.PATH_TENSOR in paths and (nb.PATH_IMMA in paths) == (N <= 832)
    for path in paths:
        eng.set_path(path)
        enc = eng.encrypt_batch(r.astype(np.uint8), m.astype(np.uint8))
        for k in ENC_KEYS:
            assert np.array_equal(enc[k], want_e[k]), (N, q, path, k)
        dec = eng.decrypt_batch(enc["value"])
        for k in DEC_KEYS:
            assert np.array_equal(dec[k], want_d[k]), (N, q, path, k)
    eng.set_path(nb.PATH_AUTO)
    hh = rng.integers(0, q, size=(B, N)); ff = rng.integers(-1, 2, size=(B, N)); pp = rng.integers(0, p, size=(B, N))
    want_e = o.encrypt_batch(hh, r, m, q)
    want_d = o.decrypt_batch(ff, pp, want_e["value"], q, p)
    enc = eng.encrypt_batch(r.astype(np.uint8), m.astype(np.uint8), h=hh.astype(np.uint16))
    dec = eng.decrypt_batch(enc["value"], f=ff.astype(np.int8), fp=pp.astype(np.uint8))
    assert eng.last_path == (nb.PATH_IMMA if N <= 832 else nb.PATH_CUDA_CORE)
    for k in ENC_KEYS:
        assert np.array_equal(enc[k], want_e[k]), (N, q, "keys", k)
    for k in DEC_KEYS:
        assert np.array_equal(dec[k], want_d[k]), (N, q, "keys", k)
    e = rng.integers(0, q, size=(301, N), dtype=np.uint16)
    assert np.array_equal(eng.sum(e), o.sum_batch(e, q))
    eng.close()


def _imma_fuzz_sets():
    rng = np.random.default_rng(20261019)
    ns = set(int(x) for x in rng.integers(8, 833, size=22))
    # geometry edges of the IMMA kernels: N = 0, 1, 2 mod 16 (the last Toeplitz step is one block shorter at N = 1 mod 16),
    # both sides of every accumulator bucket (3 / 8 / 11 / 13 column blocks: N = 192 / 512 / 704 / 832)
    ns |= {16, 17, 18, 191, 192, 193, 208, 209, 496, 497, 512, 528, 529, 672, 673, 703, 704, 705, 720, 721, 816, 817, 831, 832}
    qs = [4, 16, 128, 256, 512, 2048, 4096, 8192]
    return sorted((n, qs[(n * 7 + n // 16) % len(qs)]) for n in ns)


@pytest.mark.parametrize("N,q", _imma_fuzz_sets())
def test_imma_schedule_generic_geometry_vs_oracle(N, q, nb):
    """The IMMA kernels of an N bucket run every (step, block) of the bucket on operands zero-padded that far; only the
    BASELINE N have kernels compiled for them.  Seeded spread of N over all four buckets and the geometry edges, distinct
    keys, every witness array against the oracle (rows that drive the accumulators to their extremes included)."""
    p, B = 3, 37
    rng = np.random.default_rng(7 * N + q)
    hh = rng.integers(0, q, size=(B, N)); ff = rng.integers(-1, 2, size=(B, N)); pp = rng.integers(0, p, size=(B, N))
    r = rng.integers(0, 3, size=(B, N)); m = rng.integers(0, 2, size=(B, N))
    r[0], m[0], hh[0] = 2, (255 if q > 256 else 1), q - 1          # largest product coefficients
    ff[1], hh[1] = -1, q - 1
    pp[2] = 2
    want_e = o.encrypt_batch(hh, r, m, q)
    want_d = o.decrypt_batch(ff, pp, want_e["value"], q, p)
    eng = nb.Engine(N, p, q, 0)
    eng.set_path(nb.PATH_IMMA)
    enc = eng.encrypt_batch(r.astype(np.uint8), m.astype(np.uint8), h=hh.astype(np.uint16))
    dec = eng.decrypt_batch(enc["value"], f=ff.astype(np.int8), fp=pp.astype(np.uint8))
    assert eng.last_path == nb.PATH_IMMA
    for k in ENC_KEYS:
        assert np.array_equal(enc[k], want_e[k]), (N, q, k)
    for k in DEC_KEYS:
        assert np.array_equal(dec[k], want_d[k]), (N, q, k)
    eng.close()


@pytest.mark.parametrize("N,q", [(167, 128), (509, 2048), (512, 2048), (513, 2048), (545, 2048), (640, 4096), (641, 2048), (677, 2048), (701, 8192),
                                 (768, 8192), (821, 4096), (833, 2048), (897, 2048), (1024, 8192)])
def test_many_tiles_per_cluster_outside_baseline(N, q, nb):
    """Several 256-row tiles per CTA pair (ring phases, resident-slot reuse, the last partial tile) at the N where the
    shared-memory budget of the tcgen05 kernel changes (A slots 5 ... 8, B ring 8 ... 5, streaming DEC1): device-resident
    batch, tcgen05 schedule against the fp32 and IMMA schedules bit for bit (pad columns included), plus rows against the
    oracle."""
    torch = pytest.importorskip("torch")
    p, dev = 3, "cuda"
    rng = np.random.default_rng(77 * N + q)
    h = rng.integers(0, q, size=N)
    f = rng.integers(-1, 2, size=N)
    fp = rng.integers(0, p, size=N)
    eng = nb.Engine(N, p, q, 0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_public_key(h.astype(np.uint16))
    eng.set_private_key(f.astype(np.int8), fp.astype(np.uint8))
    P, B = eng.pitch, 74 * 256 * 3 + 131
    gen = torch.Generator(device=dev).manual_seed(N)
    r = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    r[:, :N] = torch.randint(0, 3, (B, N), generator=gen, device=dev, dtype=torch.uint8)
    m = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    m[:, :N] = torch.randint(0, 2, (B, N), generator=gen, device=dev, dtype=torch.uint8)
    outs = {}
    paths = (nb.PATH_TENSOR, nb.PATH_CUDA_CORE) + ((nb.PATH_IMMA,) if N <= 832 else ())
    for path in paths:
        eng.set_path(path)
        bufs = [torch.full((B, P), 7, dtype=torch.int16, device=dev) for _ in range(4)] + \
               [torch.full((B, P), 7, dtype=torch.uint8, device=dev) for _ in range(2)]
        val, quo, q1, r1, pv, q2 = bufs
        eng.encrypt_dev(B, r, m, value=val, quotientE=quo)
        eng.decrypt_dev(B, val, value=pv, quotient1=q1, remainder1=r1, quotient2=q2)
        eng.sync()
        outs[path] = bufs
    for path in paths[1:]:
        for x, y in zip(outs[nb.PATH_TENSOR], outs[path]):   # pad columns included: every schedule writes them as zero
            assert torch.equal(x, y), (N, q, path)
    # the round-1 phase order (full cyclic product, then the hi product) against the default (hi, then lo on top of it)
    eng.set_path(nb.PATH_TENSOR)
    eng.set_schedule(True)
    bufs = [torch.full((B, P), 7, dtype=torch.int16, device=dev) for _ in range(4)] + \
           [torch.full((B, P), 7, dtype=torch.uint8, device=dev) for _ in range(2)]
    eng.encrypt_dev(B, r, m, value=bufs[0], quotientE=bufs[1])
    eng.decrypt_dev(B, bufs[0], value=bufs[4], quotient1=bufs[2], remainder1=bufs[3], quotient2=bufs[5])
    eng.sync()
    eng.set_schedule(False)
    for x, y in zip(outs[nb.PATH_TENSOR], bufs):
        assert torch.equal(x, y), (N, q, "cyc + hi order")
    # both arrangements of the epilogue warps (two groups, one per TMEM buffer / one group on every phase: other chunk tables)
    for mode in (1, 2):
        eng.set_epilogue(mode)
        for b in bufs:
            b.fill_(7)
        eng.encrypt_dev(B, r, m, value=bufs[0], quotientE=bufs[1])
        eng.decrypt_dev(B, bufs[0], value=bufs[4], quotient1=bufs[2], remainder1=bufs[3], quotient2=bufs[5])
        eng.sync()
        for x, y in zip(outs[nb.PATH_TENSOR], bufs):
            assert torch.equal(x, y), (N, q, "epilogue arrangement", mode)
    eng.set_epilogue(0)
    del bufs
    # value-only mode (no hi chunks: other loop counts and ring phases), same ciphertexts and plaintexts
    val2 = torch.full((B, P), 7, dtype=torch.int16, device=dev)
    pv2 = torch.full((B, P), 7, dtype=torch.uint8, device=dev)
    eng.encrypt_dev(B, r, m, value=val2)
    eng.decrypt_dev(B, val2, value=pv2)
    eng.sync()
    assert torch.equal(val2, outs[nb.PATH_TENSOR][0]) and torch.equal(pv2, outs[nb.PATH_TENSOR][4]), (N, q)
    idx = [0, 255, 256, 74 * 256, B - 132, B - 1]
    want_e = o.encrypt_batch(h, r[idx, :N].cpu().numpy(), m[idx, :N].cpu().numpy(), q)
    want_d = o.decrypt_batch(f, fp, want_e["value"], q, p)
    val, quo, q1, r1, pv, q2 = outs[nb.PATH_TENSOR]
    assert np.array_equal(val[idx].cpu().numpy().view(np.uint16)[:, :N], want_e["value"])
    assert np.array_equal(quo[idx].cpu().numpy().view(np.uint16)[:, : N + 1], want_e["quotientE"])
    assert np.array_equal(r1[idx].cpu().numpy().view(np.uint16)[:, : N + 1], want_d["remainder1"])
    assert np.array_equal(q1[idx].cpu().numpy().view(np.uint16)[:, : N + 1], want_d["quotient1"])
    assert np.array_equal(pv[idx].cpu().numpy()[:, :N], want_d["value"])
    assert np.array_equal(q2[idx].cpu().numpy()[:, : N + 1], want_d["quotient2"])
    eng.close()


def test_string_batches_and_multi_block_messages(nb, golden):
    """encryptStr / decryptStr for many strings per call, against the single-string reference path
    (index.js:80-86) and the oracle's codec; strings longer than floor(N / 8) characters split into blocks."""
    g = golden("default167")
    k = o.make_key("default167", 300)
    mine = nb.NTRU(dict(o.CONFIGS["default167"]), f=list(k.f), fq=list(k.fq), fp=list(k.fp), g=list(k.g), h=list(k.h),
                   rand32=seeded(11))
    texts = ["Hello World", "", "a", "x" * 20, "NTRU on B200!", "\xe9\xff"]
    enc = mine.encryptStrBatch(texts)
    assert enc.shape == (len(texts), 167)
    # the reference decrypts the empty string to one NUL character: trim -> [0] -> padded to 8 bits (index.js:84-86)
    want = [t if t else "\x00" for t in texts]
    assert mine.decryptStrBatch(enc) == want
    for t, row in zip(want, enc):                                      # each row decrypts through the one-string path too
        assert mine.decryptStr(o.trim_polynomial(row.tolist())) == t
        assert k.decryptStr(o.trim_polynomial(row.tolist())) == t     # ... and through the oracle's class
    with pytest.raises(IndexError):
        mine.encryptStrBatch(["y" * 21])                               # 21 characters > floor(167 / 8): index.js:98
    long_texts = ["The quick brown fox jumps over the lazy dog. " * 3, "short", "z" * 40]
    blocks = mine.encryptStrBatch(long_texts, split=True)
    assert [b.shape[0] for b in blocks] == [-(-len(t) // 20) for t in long_texts]
    assert mine.decryptStrBatch(blocks) == long_texts


@pytest.mark.parametrize("cfg", CFGS)
def test_keygen_batch_vs_oracle(cfg, nb, engines, golden):
    """loadPrivateKeyF + generatePublicKeyH for a batch (host extended Euclid modulo 2 and p, GPU lifting
    to q and h): fq, fp and h of valid oracle keys bit for bit (the inverse is unique), non-invertible f flagged."""
    g, eng = golden(cfg), engines(cfg)
    N, q, p = int(g["N"]), int(g["q"]), int(g["p"])
    keys = [o.make_key(cfg, 400 + i) for i in range(6)]
    B = len(keys) + 3
    f = np.zeros((B, N), dtype=np.int8); gg = np.zeros((B, N), dtype=np.int8)
    for i, k in enumerate(keys):
        f[i], gg[i] = k.f, k.g
    f[len(keys)] = 0                                      # zero polynomial
    f[len(keys) + 1, :2] = 1                              # f(1) even: not invertible modulo 2
    f[len(keys) + 2, :3] = 1                              # f(1) = 3: not invertible modulo 3
    gg[len(keys):] = keys[0].g
    out = eng.keygen_batch(f, gg)
    assert out["valid"].tolist() == [True] * len(keys) + [False] * 3
    for i, k in enumerate(keys):
        assert out["fq"][i].tolist() == o.expand_array(k.fq, N, 0), (cfg, i, "fq")
        assert out["fp"][i].tolist() == o.expand_array(k.fp, N, 0), (cfg, i, "fp")
        assert out["h"][i].tolist() == o.expand_array(k.h, N, 0), (cfg, i, "h")
    assert not out["fq"][len(keys):].any() and not out["h"][len(keys):].any() and not out["fp"][len(keys):].any()
    # the class-level batch: every key it returns is a working key pair
    rng = random.Random(77)
    kk = nb.NTRU(dict(o.CONFIGS[cfg]), rand32=lambda: rng.getrandbits(32))
    ks = kk.generateKeysBatch(5)
    m = np.random.default_rng(1).integers(0, 2, size=(5, N)).astype(np.uint8)
    r = o.sample_ternary_rows(5, N, int(g["dr"]), int(g["dr"]), np.random.default_rng(2)).astype(np.uint8)
    want_e = o.encrypt_batch(ks["h"].astype(np.int64), r, m, q)
    want_d = o.decrypt_batch(ks["f"].astype(np.int64), ks["fp"].astype(np.int64), want_e["value"], q, p)
    dec = kk.decryptBitsBatch(kk.encryptBitsBatch(m, r, hs=ks["h"])["value"], fs=ks["f"], fps=ks["fp"])
    assert np.array_equal(dec["value"], want_d["value"])
    for i in range(5):                                    # f * fq = 1 (mod q, x^N - 1) and f * fp = 1 (mod p, x^N - 1)
        c = np.convolve(ks["f"][i].astype(np.int64), ks["fq"][i].astype(np.int64))
        cyc = c[:N].copy(); cyc[: N - 1] += c[N:]
        assert (np.mod(cyc, q) == np.eye(1, N, 0, dtype=np.int64)[0]).all()
        c = np.convolve(ks["f"][i].astype(np.int64), ks["fp"][i].astype(np.int64))
        cyc = c[:N].copy(); cyc[: N - 1] += c[N:]
        assert (np.mod(cyc, p) == np.eye(1, N, 0, dtype=np.int64)[0]).all()
    if q % 3 == 2 and cfg != "tiny17":
        assert np.array_equal(dec["value"], m)


def test_empty_batch_and_errors(nb, engines):
    eng = engines("default167")
    z8 = np.zeros((0, 167), dtype=np.uint8)
    out = eng.encrypt_batch(z8, z8)
    assert out["value"].shape == (0, 167)
    assert eng.decrypt_batch(np.zeros((0, 167), dtype=np.uint16))["value"].shape == (0, 167)
    assert np.array_equal(eng.sum(np.zeros((0, 167), dtype=np.uint16)), np.zeros(167, dtype=np.uint16))
    with pytest.raises(IndexError):
        eng.encrypt_batch(np.zeros((2, 168), dtype=np.uint8), np.zeros((2, 168), dtype=np.uint8))
    fresh = nb.Engine(167, 3, 128, 0)
    with pytest.raises(nb.NtruError, match="public key"):
        fresh.encrypt_batch(np.zeros((1, 167), dtype=np.uint8), np.zeros((1, 167), dtype=np.uint8))
    with pytest.raises(nb.NtruError, match="private key"):
        fresh.decrypt_batch(np.zeros((1, 167), dtype=np.uint16))
    with pytest.raises(nb.NtruError):
        fresh.set_public_key(np.full(167, 128, dtype=np.uint16))      # coefficient outside [0,q)
    fresh.close()


def test_wide_messages(nb, engines, golden):
    """m coefficients >= 256 (the reference reduces any integer mod q, index.js:91)."""
    g, eng = golden("hps509"), engines("hps509")
    N, q = 509, 2048
    rng = np.random.default_rng(8)
    r = o.sample_ternary_rows(5, N, 169, 169, rng).astype(np.uint8)
    m = rng.integers(0, q, size=(5, N))
    want = o.encrypt_batch(g["h"].astype(np.int64), r, m, q)
    enc = eng.encrypt_batch(r, m.astype(np.uint16))
    for k in ENC_KEYS:
        assert np.array_equal(enc[k], want[k]), k


@pytest.mark.parametrize("cfg", ["tiny17", "default167", "hrss701"])
def test_class_api_matches_reference_objects(cfg, nb, golden):
    """encryptBits/decryptBits return the same {value, inputs, params} as the oracle's class NTRU."""
    g = golden(cfg)
    opts = dict(o.CONFIGS[cfg], f=g["f"].tolist(), fp=o.trim_polynomial(g["fp"].tolist()),
                fq=o.trim_polynomial(g["fq"].tolist()), g=g["g"].tolist(), h=o.trim_polynomial(g["h"].tolist()))
    ref = o.NTRU(dict(opts), literal=cfg == "tiny17")
    mine = nb.NTRU(dict(opts))
    m = [1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1]           # test/circuits.test.js:258
    r = ref.sample_r()
    a, b = ref.encryptBits(m, r), mine.encryptBits(m, r)
    assert dict(b) == a
    assert b["input"] is b["inputs"]
    assert o.verify_encrypt(b["inputs"], b["params"])
    da, db = ref.decryptBits(a["value"]), mine.decryptBits(b["value"])
    assert dict(db) == da
    assert o.verify_decrypt(db["inputs"], db["params"])
    assert db["value"] == m
    # tamper (test/circuits.test.js:296-301)
    db["inputs"]["remainder2"][0] += 1
    assert not o.verify_decrypt(db["inputs"], db["params"])
    with pytest.raises(IndexError):
        mine.encryptBits([1] * (mine.N + 1))
    with pytest.raises(IndexError):
        mine.decryptBits([1] * (mine.N + 1))


def test_string_roundtrip_and_wrong_key(nb):
    # test/reference.test.js:6-25
    k = nb.NTRU(rand32=seeded(21))
    k.generatePrivateKeyF()
    k.generateNewPublicKeyGH()
    e = k.encryptStr("Hello World")
    assert k.decryptStr(e) == "Hello World"
    other = nb.NTRU(rand32=seeded(22))
    other.generatePrivateKeyF()
    assert other.decryptStr(e) != "Hello World"


def test_large_key_string_roundtrip(nb):
    # test/reference.test.js:27-44
    d = 701 // 3
    k = nb.NTRU({"N": 701, "q": 8192, "df": d, "dg": d, "dr": d}, rand32=seeded(31))
    k.generatePrivateKeyF()
    k.generateNewPublicKeyGH()
    assert k.decryptStr(k.encryptStr("Big polys")) == "Big polys"


def test_additive_homomorphism(nb, golden):
    # test/reference.test.js:48-61 ("this test may fail" upstream: use a fixed valid key)
    g = golden("default167")
    k = nb.NTRU(dict(o.CONFIGS["default167"], f=g["f"].tolist(), fp=o.trim_polynomial(g["fp"].tolist()),
                     h=o.trim_polynomial(g["h"].tolist())), rand32=seeded(1))
    e1 = k.encryptBits([1, 2, 1, 0, 1])["value"]
    e2 = k.encryptBits([0, 1, 1, 1, 0, 1, 0, 1])["value"]
    s = nb.addPolynomials(e1, e2, k.q)
    es = np.zeros((2, k.N), dtype=np.uint16)
    es[0, : len(e1)] = e1
    es[1, : len(e2)] = e2
    assert k.sumCiphertexts(es) == s == o.sum_ciphertexts([e1, e2], k.q)
    assert k.decryptBits(s)["value"] == [1, 0, 2, 1, 1, 1, 0, 1]


def test_sum_many_rows(nb, engines):
    """Config 5 shape at reduced count: column sums of uniform rows mod q, plus linearity."""
    eng = engines("hrss701")
    rng = np.random.default_rng(3)
    e = rng.integers(0, 8192, size=(20011, 701), dtype=np.uint16)
    want = o.sum_batch(e, 8192)
    assert np.array_equal(eng.sum(e), want)
    a, b = eng.sum(e[:9000]), eng.sum(e[9000:])
    assert np.array_equal((a.astype(np.int64) + b) % 8192, want)


@pytest.mark.parametrize("cfg", ["default167", "hrss701"])
def test_sum_over_exchange_window_single_rank(cfg, nb, golden):
    """ntru_sum_allreduce_dev with a world of one rank (the exchange window is this GPU's own memory): the fused
    column-sum + push kernel and the gather kernel against the oracle fold, repeated calls (both slot parities,
    re-zeroed partials), ragged and empty shards.  The multi-rank path is scripts/bench_multi.py (gpurun --gpus N)."""
    torch = pytest.importorskip("torch")
    from ntru_circom_b200 import sharding
    g = golden(cfg)
    N, q = int(g["N"]), int(g["q"])
    eng = nb.Engine(N, 3, q, 0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)      # torch fills and engine kernels on one stream
    sharding.connect_exchange(eng)
    rng = np.random.default_rng(8)
    P = eng.pitch
    for rows in (5003, 1, 0, 20011, 777):
        e = rng.integers(0, q, size=(rows, N), dtype=np.uint16)
        ed = torch.zeros((max(rows, 1), P), dtype=torch.int16, device="cuda")
        if rows:
            ed[:rows, :N] = torch.from_numpy(e.astype(np.int16)).cuda()
        out = sharding.sum_ciphertexts_exchange(eng, ed, rows)
        eng.sync()
        got = out.cpu().numpy().astype(np.uint16)
        want = o.sum_batch(e, q) if rows else np.zeros(N, dtype=np.int64)
        assert np.array_equal(got[:N], want), (cfg, rows)
        assert not got[N:].any()
    eng.close()


class _Draws:
    """rng stand-in for the oracle's generate_custom_array: hands out a fixed list of 32-bit draws."""

    def __init__(self, words):
        self.it = iter(words)

    def getrandbits(self, n):
        assert n == 32
        return next(self.it)


def _replay_r(nb, key, row, N, dr):
    """r of global row number `row` as the ORACLE computes it from the generator's draws (index.js:89, 461-488)."""
    arr = o.generate_custom_array(N, dr, dr, _Draws(nb.sampler_draws(key, row, N - 1)))
    return [2 if x == -1 else x for x in arr]


@pytest.mark.parametrize("cfg", ["tiny17", "default167", "hps509", "hps821"])
def test_device_sampler_matches_oracle_fisher_yates(cfg, nb, engines, golden):
    """The device sampler against the oracle's generate_custom_array fed with the same ChaCha20 draws (host copy of
    the generator: pinned on the RFC 8439 vector in tests/test_host.py)."""
    torch = pytest.importorskip("torch")
    g, eng = golden(cfg), engines(cfg)
    N, dr, B, seed, row0 = int(g["N"]), int(g["dr"]), 70, 99, (1 << 40) + 5
    r = torch.zeros((B, eng.pitch), dtype=torch.uint8, device="cuda")
    eng.sample_r_dev(B, dr, row0, r, seed=seed)
    eng.sync()
    got = r.cpu().numpy()
    assert not got[:, N:].any()
    for b in (0, 1, 63, 64, 69):
        assert got[b, :N].tolist() == _replay_r(nb, nb.seed_key(seed), row0 + b, N, dr), (cfg, b)
    assert (np.count_nonzero(got == 1, axis=1) == dr).all() and (np.count_nonzero(got == 2, axis=1) == dr).all()
    with pytest.raises(nb.NtruError, match="cannot exceed"):           # index.js:462-464
        eng.sample_r_dev(1, N // 2 + 1, 0, r)


@pytest.mark.parametrize("cfg", ["default167", "hps509"])
def test_encrypt_with_device_drawn_r(cfg, nb, golden):
    """ntru_encrypt_batch with r == NULL (index.js:89: the call draws r itself): r_out is the exact-weight ternary row
    the oracle derives from the generator's draws for that row number, the ciphertext and witness are the oracle's for
    that r, row numbers are never reused, and several pipeline chunks continue the numbering."""
    g = golden(cfg)
    N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
    eng = nb.Engine(N, p, q, 0)
    eng.set_public_key(g["h"])
    eng.set_private_key(g["f"], g["fp"])
    m = np.random.default_rng(3).integers(0, 2, size=(700, N)).astype(np.uint8)
    with pytest.raises(nb.NtruError, match="NTRU_OPT_DR"):
        eng.encrypt_batch(None, m)
    # 1. the context's own key (OS entropy): valid weights, decrypts, fresh r on every call
    a = eng.encrypt_batch(None, m, dr=dr)
    b = eng.encrypt_batch(None, m)
    assert eng.rng_next_row == 1400
    for x in (a, b):
        assert (np.count_nonzero(x["r"] == 1, axis=1) == dr).all() and (np.count_nonzero(x["r"] == 2, axis=1) == dr).all()
        assert x["r"].max() <= 2
        # decrypts as the oracle decrypts it; NOT compared with m row for row: r comes from OS entropy here, and at the
        # default parameters (q = 128) about one random r in 3000 wraps the centred lift ("this test may fail" upstream)
        dec = eng.decrypt_batch(x["value"], witness=False)["value"]
        assert np.array_equal(dec, o.decrypt_batch(g["f"].astype(np.int64), g["fp"].astype(np.int64), x["value"], q, p)["value"])
        assert (dec == m).all(axis=1).mean() > 0.99
    assert not np.array_equal(a["r"], b["r"]) and not np.array_equal(a["value"], b["value"])
    # 2. a known key: replay through the oracle, on every schedule, with chunked pipelining (128-row chunks)
    key = bytes(range(100, 132))
    eng.set_option(2, 128)                                   # NTRU_OPT_CHUNK_ROWS
    for path in _paths(nb, eng):
        eng.set_path(path)
        eng.set_rng_key(key, 1000)
        enc = eng.encrypt_batch(None, m)
        assert eng.rng_next_row == 1700
        rows = [0, 1, 127, 128, 129, 383, 384, 699]
        want_r = np.array([_replay_r(nb, key, 1000 + i, N, dr) for i in rows])
        assert np.array_equal(enc["r"][rows], want_r), (cfg, path)
        want = o.encrypt_batch(g["h"].astype(np.int64), enc["r"], m, q)
        for k in ENC_KEYS:
            assert np.array_equal(enc[k], want[k]), (cfg, path, k)
        # value-only, r not returned: same ciphertexts for the same row numbers
        eng.set_rng_key(key, 1000)
        v2 = np.zeros_like(enc["value"])
        rc = eng.lib.ntru_encrypt_batch(eng._h, 700, None, m.ctypes.data, v2.ctypes.data, None, None, None)
        assert rc == 0 and np.array_equal(v2, want["value"])
    # 3. an injected r is echoed into r_out
    r = o.sample_ternary_rows(5, N, dr, dr, np.random.default_rng(1)).astype(np.uint8)
    assert np.array_equal(eng.encrypt_batch(r, m[:5], return_r=True)["r"], r)
    eng.close()


def test_device_resident_api_and_properties_at_scale(nb, engines, golden):
    """Device-pointer entry points at a large batch: decrypt(encrypt(m)) == m and witness identities."""
    torch = pytest.importorskip("torch")
    g, eng = golden("hps509"), engines("hps509")
    N, q, P, B = 509, 2048, eng.pitch, 1 << 16
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(1)
    r = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    eng.sample_r_dev(B, 169, 0, r, seed=7)
    m = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    m[:, :N] = torch.randint(0, 2, (B, N), generator=gen, device=dev, dtype=torch.uint8)
    val = torch.empty((B, P), dtype=torch.int16, device=dev)
    quo = torch.empty((B, P), dtype=torch.int16, device=dev)
    out = torch.empty((B, P), dtype=torch.uint8, device=dev)
    q1 = torch.empty((B, P), dtype=torch.int16, device=dev)
    r1 = torch.empty((B, P), dtype=torch.int16, device=dev)
    q2 = torch.empty((B, P), dtype=torch.uint8, device=dev)
    eng.encrypt_dev(B, r, m, value=val, quotientE=quo)
    eng.decrypt_dev(B, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2)
    eng.sync()
    assert torch.equal(out[:, :N], m[:, :N])
    # spot-check rows against the oracle
    idx = [0, 1, 12345, B - 1]
    want = o.encrypt_batch(g["h"].astype(np.int64), r[idx, :N].cpu().numpy(), m[idx, :N].cpu().numpy(), q)
    got = val[idx].cpu().numpy().view(np.uint16)
    assert np.array_equal(got[:, :N], want["value"])
    assert np.array_equal(quo[idx].cpu().numpy().view(np.uint16)[:, : N + 1], want["quotientE"])
    # pad / entry N are zero so that the arrays are valid witness rows
    assert not quo[:, N - 1:].any() and not val[:, N:].any()
    # checksum of checksums: sum of ciphertexts decrypts to the sum of messages' columns mod q relation
    part = torch.zeros(P, dtype=torch.int32, device=dev)
    eng.sum_partial_dev(B, val, part)
    s16 = torch.empty(P, dtype=torch.int16, device=dev)
    eng.sum_finalize_dev(part, s16)
    eng.sync()
    want_sum = (val[:, :N].to(torch.int64) & 0xFFFF).sum(dim=0) % q
    assert torch.equal(s16[:N].to(torch.int64) & 0xFFFF, want_sum)


@pytest.mark.parametrize("B", [1, 127, 128, 129, 255, 256, 257, 513])
@pytest.mark.parametrize("cfg", ["hps509", "hps821"])
def test_tile_boundaries_tensor_kernel(B, cfg, nb, engines, golden):
    """Row counts around the 128-row CTA tile and the 256-row CTA-pair tile, with and without witness, on the tcgen05
    schedule (resident A operand at N = 509, streaming DEC1 and unequal accumulator chunks at N = 821)."""
    g, eng = golden(cfg), engines(cfg)
    N, q, p, dr = int(g["N"]), int(g["q"]), 3, int(g["dr"])
    rng = np.random.default_rng(B)
    r = o.sample_ternary_rows(B, N, dr, dr, rng).astype(np.uint8)
    m = rng.integers(0, 2, size=(B, N)).astype(np.uint8)
    want_e = o.encrypt_batch(g["h"].astype(np.int64), r, m, q)
    want_d = o.decrypt_batch(g["f"].astype(np.int64), g["fp"].astype(np.int64), want_e["value"], q, p)
    eng.set_path(nb.PATH_TENSOR)
    try:
        for witness in (True, False):
            enc = eng.encrypt_batch(r, m, witness=witness)
            dec = eng.decrypt_batch(want_e["value"].astype(np.uint16), witness=witness)
            assert np.array_equal(enc["value"], want_e["value"]), witness
            assert np.array_equal(dec["value"], want_d["value"]), witness
            if witness:
                for k in ENC_KEYS:
                    assert np.array_equal(enc[k], want_e[k]), k
                for k in DEC_KEYS:
                    assert np.array_equal(dec[k], want_d[k]), k
    finally:
        eng.set_path(nb.PATH_AUTO)


def test_full_size_round_trip_and_checksums(nb, engines, golden):
    """BASELINE config 2 at full size (1,000,000 ciphertexts, one key): size-independent properties.
    decrypt(encrypt(m)) == m for every row, the witness satisfies the division identity
    c = quotient * (1 - x^N) + remainder coefficient-wise (checked through column checksums), and the
    CUDA-core schedule agrees with the tensor schedule on a strided sample of rows."""
    torch = pytest.importorskip("torch")
    g, eng = golden("hps509"), engines("hps509")
    N, q, P, B = 509, 2048, eng.pitch, 1_000_000
    dev = "cuda"
    r = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    eng.sample_r_dev(B, 169, 0, r, seed=11)
    m = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    m[:, :N] = torch.randint(0, 2, (B, N), device=dev, dtype=torch.uint8)
    val = torch.empty((B, P), dtype=torch.int16, device=dev)
    quo = torch.empty((B, P), dtype=torch.int16, device=dev)
    out = torch.empty((B, P), dtype=torch.uint8, device=dev)
    r1 = torch.empty((B, P), dtype=torch.int16, device=dev)
    q1 = torch.empty((B, P), dtype=torch.int16, device=dev)
    q2 = torch.empty((B, P), dtype=torch.uint8, device=dev)
    eng.encrypt_dev(B, r, m, value=val, quotientE=quo)
    eng.decrypt_dev(B, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2)
    eng.sync()
    assert eng.last_path == nb.PATH_TENSOR
    assert torch.equal(out[:, :N], m[:, :N])
    # exact weights of the sampled r, zero pads everywhere
    assert bool(((r == 1).sum(dim=1) == 169).all()) and bool(((r == 2).sum(dim=1) == 169).all())
    assert not bool(val[:, N:].any()) and not bool(quo[:, N - 1:].any()) and not bool(q1[:, N - 1:].any())
    # strided sample through the CUDA-core schedule (different kernel, same answers)
    idx = torch.arange(0, B, 997, device=dev)
    rs, ms = r[idx].contiguous(), m[idx].contiguous()
    v2 = torch.empty((len(idx), P), dtype=torch.int16, device=dev)
    qe2 = torch.empty_like(v2)
    eng.set_path(nb.PATH_CUDA_CORE)
    eng.encrypt_dev(len(idx), rs, ms, value=v2, quotientE=qe2)
    eng.sync()
    eng.set_path(nb.PATH_AUTO)
    assert torch.equal(v2, val[idx]) and torch.equal(qe2, quo[idx])
    # and a few rows against the oracle
    pick = [0, 499_999, B - 1]
    want = o.encrypt_batch(g["h"].astype(np.int64), r[pick, :N].cpu().numpy(), m[pick, :N].cpu().numpy(), q)
    assert np.array_equal(val[pick].cpu().numpy().view(np.uint16)[:, :N], want["value"])
    assert np.array_equal(quo[pick].cpu().numpy().view(np.uint16)[:, : N + 1], want["quotientE"])


@pytest.mark.parametrize("cfg", ["default167", "hps509", "hps677", "hps821"])
def test_pad_columns_of_device_rows_do_not_count(cfg, nb, engines, golden):
    """Device-resident rows have pitch P > N.  The contract (include/ntru_b200.h) asks for zero pad columns in the inputs
    and writes zero pad columns in every output, on every schedule.  The tcgen05 schedule is also insensitive to what a
    caller leaves in columns N..P-1 of r, m and e: it reads them through TMA maps of extent N, against zero key-matrix
    columns, or with an explicit mask (ENC above N = 512 reads its message bytes from global memory)."""
    torch = pytest.importorskip("torch")
    g, eng = golden(cfg), engines(cfg)
    N, q, dr, P, dev = int(g["N"]), int(g["q"]), int(g["dr"]), eng.pitch, "cuda"
    B = 74 * 256 + 77
    gen = torch.Generator(device=dev).manual_seed(5)
    r = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    eng.sample_r_dev(B, dr, 0, r, seed=9)
    m = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    m[:, :N] = torch.randint(0, 2, (B, N), generator=gen, device=dev, dtype=torch.uint8)

    def run(path, r_, m_, dirty_e):
        eng.set_path(path)
        val = torch.full((B, P), 7, dtype=torch.int16, device=dev)
        quo = torch.full((B, P), 7, dtype=torch.int16, device=dev)
        torch.cuda.synchronize()                          # the engine has its own stream: order it after torch's fills ...
        eng.encrypt_dev(B, r_, m_, value=val, quotientE=quo)
        eng.sync()                                        # ... and torch's reads after the engine's kernels
        e = val.clone()
        if dirty_e:
            e[:, N:] = 0x7FF
        outs = [torch.full((B, P), 7, dtype=torch.int16, device=dev) for _ in range(2)] + \
               [torch.full((B, P), 7, dtype=torch.uint8, device=dev) for _ in range(2)]
        q1, r1, pv, q2 = outs
        torch.cuda.synchronize()
        eng.decrypt_dev(B, e, value=pv, quotient1=q1, remainder1=r1, quotient2=q2)
        eng.sync()
        return [val, quo] + outs

    ref = None
    for path in (nb.PATH_TENSOR, nb.PATH_IMMA, nb.PATH_CUDA_CORE):
        clean = run(path, r, m, False)
        for x in clean:
            assert not x[:, N:].any(), (cfg, path)
        if ref is None:
            ref = clean
        for x, y in zip(ref, clean):                      # the three schedules agree, pads included
            assert torch.equal(x, y), (cfg, path)
    rd, md = r.clone(), m.clone()
    rd[:, N:] = 2
    md[:, N:] = 0xA5
    torch.cuda.synchronize()
    dirty = run(nb.PATH_TENSOR, rd, md, True)
    for x, y in zip(ref, dirty):
        assert torch.equal(x, y), cfg
    eng.set_path(0)


@pytest.mark.parametrize("N,q", [(167, 128), (509, 2048), (640, 4096), (676, 2048), (677, 2048), (701, 8192), (821, 4096),
                                 (832, 4096)])
def test_distinct_keys_device_rows_all_schedules_agree(N, q, nb):
    """Distinct keys per row through the device-pointer entry points: the IMMA schedule against the fp32 schedule, bit
    for bit over whole pitched rows (pad columns written as zero), with a batch that is not a multiple of any tile.
    At the BASELINE N the IMMA kernels are instantiated for that N (exact Toeplitz band, constant offsets): the bucket's
    generic instantiation (NTRU_OPT_IMMA_FORM = 1) is run as well and must agree bit for bit; and whatever a caller
    leaves in the pad columns N..P-1 of h, f, fp, r, m and e must not count (the kernels zero them in shared memory)."""
    torch = pytest.importorskip("torch")
    p, dev = 3, "cuda"
    eng = nb.Engine(N, p, q, 0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    P, B = eng.pitch, 4099
    gen = torch.Generator(device=dev).manual_seed(N + q)
    def rows(lo, hi, dtype):
        t = torch.zeros((B, P), dtype=dtype, device=dev)
        t[:, :N] = torch.randint(lo, hi, (B, N), generator=gen, device=dev).to(dtype)
        return t
    h, f, fp = rows(0, q, torch.int16), rows(-1, 2, torch.int8), rows(0, 3, torch.uint8)
    r, m = rows(0, 3, torch.uint8), rows(0, 2, torch.uint8)
    outs = {}
    for path in (nb.PATH_IMMA, nb.PATH_CUDA_CORE):
        eng.set_path(path)
        bufs = [torch.full((B, P), 7, dtype=torch.int16, device=dev) for _ in range(4)] + \
               [torch.full((B, P), 7, dtype=torch.uint8, device=dev) for _ in range(2)]
        val, quo, q1, r1, pv, q2 = bufs
        eng.encrypt_dev(B, r, m, value=val, quotientE=quo, h_rows=h)
        eng.decrypt_dev(B, val, value=pv, quotient1=q1, remainder1=r1, quotient2=q2, f_rows=f, fp_rows=fp)
        eng.sync()
        outs[path] = bufs
    for x, y in zip(outs[nb.PATH_IMMA], outs[nb.PATH_CUDA_CORE]):
        assert torch.equal(x, y), (N, q)
        assert not x[:, N:].any(), (N, q)
    want_e = o.encrypt_batch(h[:3, :N].cpu().numpy().astype(np.int64) & 0xFFFF, r[:3, :N].cpu().numpy(), m[:3, :N].cpu().numpy(), q)
    assert np.array_equal(outs[nb.PATH_IMMA][0][:3].cpu().numpy().view(np.uint16)[:, :N], want_e["value"])
    # generic instantiation of the bucket, then dirty pad columns in every input (both forms)
    eng.set_path(nb.PATH_IMMA)
    dirty = [t.clone() for t in (h, f, fp, r, m)]
    for t, fill in zip(dirty, (0x7ABC, -1, 0xEE, 2, 0xA5)):
        t[:, N:] = fill
    hd, fd, fpd, rd, md = dirty
    for generic in (True, False):
        eng.set_imma_form(generic)
        for (h_, f_, fp_, r_, m_, dirty_e) in ((h, f, fp, r, m, False), (hd, fd, fpd, rd, md, True)):
            if not generic and not dirty_e:
                continue                                  # that is outs[PATH_IMMA]
            bufs = [torch.full((B, P), 7, dtype=torch.int16, device=dev) for _ in range(4)] + \
                   [torch.full((B, P), 7, dtype=torch.uint8, device=dev) for _ in range(2)]
            val, quo, q1, r1, pv, q2 = bufs
            eng.encrypt_dev(B, r_, m_, value=val, quotientE=quo, h_rows=h_)
            eng.sync()
            e = val.clone()
            if dirty_e:
                e[:, N:] = 0x7FFF
            eng.decrypt_dev(B, e, value=pv, quotient1=q1, remainder1=r1, quotient2=q2, f_rows=f_, fp_rows=fp_)
            eng.sync()
            for x, y in zip(outs[nb.PATH_IMMA], bufs):
                assert torch.equal(x, y), (N, q, generic, dirty_e)
    eng.close()


def test_cross_gpu_sum_exchange_world2_on_hardware():
    """The peer-memory exchange of the ciphertext sum at world size 2 on real GPUs (skipped on a one-GPU box): many
    calls back to back, uneven shards, a rank with no rows, against int64 column sums and the NCCL path
    (scripts/xchg_stress.py; the control plane alone is covered on gloo in tests/test_multi_rank.py)."""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(root, "scripts", "xchg_stress.py")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "xchg_stress ok" in res.stdout


def test_host_buffer_pack_unpack_and_sum_allreduce(nb, engines, golden):
    """Host-buffer forms of packOutput / unpackInput (index.js:572-620) against the oracle's BigInt restatement, and the
    host-buffer cross-GPU sum in a world of one rank against the oracle fold (several calls: the accumulator is reused)."""
    g, eng = golden("hps509"), engines("hps509")
    N, q = 509, 2048
    rng = np.random.default_rng(31)
    rows = rng.integers(0, q, size=(9, N)).astype(np.uint16)
    rows[0] = q - 1
    packed = eng.pack_output(q, rows)
    for b in (0, 1, 8):
        want = o.pack_output(q, N, [int(x) for x in rows[b]])
        got = [sum(int(packed[b, e, w]) << (32 * w) for w in range(8)) for e in range(packed.shape[1])]
        assert got == want["expected"] and packed.shape[1] == want["outputSize"]
    bits, per, _, outs = eng.pack_geometry(q, N)
    back = eng.unpack_input(q, per * bits, packed)
    assert np.array_equal(back[:, :N], rows) and not back[:, N:].any()
    small = rng.integers(0, 3, size=(4, N + 1)).astype(np.uint8)          # a mod-p witness row (N + 1 entries)
    p8 = eng.pack_output(2, small)
    want = o.pack_output(2, N + 1, [int(x) for x in small[3]])
    assert [sum(int(p8[3, e, w]) << (32 * w) for w in range(8)) for e in range(p8.shape[1])] == want["expected"]
    for B in (0, 1, 300, 70000):                                           # 70000 rows: three pipeline chunks
        e = rng.integers(0, q, size=(B, N)).astype(np.uint16)
        assert np.array_equal(eng.sum_allreduce(e), o.sum_batch(e, q) if B else np.zeros(N, dtype=np.uint16)), B
        assert np.array_equal(eng.sum(e), o.sum_batch(e, q) if B else np.zeros(N, dtype=np.uint16)), B


@pytest.mark.parametrize("key", ["random", "all_plus", "all_minus"])
@pytest.mark.parametrize("N,q", [(191, 512), (193, 1024), (509, 2048), (512, 2048), (513, 2048), (677, 2048), (1024, 2048)])
def test_first_decrypt_product_fp16_form(N, q, key, nb):
    """256 < q <= 2048: the tcgen05 schedule runs the first decrypt product on kind::f16 tiles (the caller's uint16
    coefficients are their own fp16 encodings, scaled by 2^-24; umma_kernels.cu, DEC1F).  Checked against the byte-limb
    int8 form bit for bit (pad columns included) on uniformly random ciphertext rows over several tiles per CTA pair, and
    against the oracle on rows that drive the fp32 accumulators to their largest magnitude (e = q - 1 everywhere under
    f = +1 / -1 everywhere: |sum| = N (q - 1)), i.e. exactness of the subnormal inputs and of the accumulation."""
    torch = pytest.importorskip("torch")
    p, dev = 3, "cuda"
    rng = np.random.default_rng(31 * N + q)
    f = {"random": rng.integers(-1, 2, size=N), "all_plus": np.ones(N, dtype=np.int64), "all_minus": -np.ones(N, dtype=np.int64)}[key]
    fp = rng.integers(0, p, size=N)
    eng = nb.Engine(N, p, q, 0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_private_key(f.astype(np.int8), fp.astype(np.uint8))
    eng.set_path(nb.PATH_TENSOR)
    P, B = eng.pitch, 74 * 256 * 2 + 77
    gen = torch.Generator(device=dev).manual_seed(N + q)
    e = torch.zeros((B, P), dtype=torch.int16, device=dev)
    e[:, :N] = torch.randint(0, q, (B, N), generator=gen, device=dev, dtype=torch.int16)
    e[0, :N] = q - 1
    e[1, :N] = 0
    e[2, :N:2] = q - 1
    e[3, :N] = 1
    e[B - 1, :N] = q - 1
    outs = []
    for form in (2, 1):
        eng.set_dec1_form(form)
        bufs = [torch.full((B, P), 7, dtype=torch.int16, device=dev) for _ in range(2)] + \
               [torch.full((B, P), 7, dtype=torch.uint8, device=dev) for _ in range(2)]
        q1, r1, pv, q2 = bufs
        eng.decrypt_dev(B, e, value=pv, quotient1=q1, remainder1=r1, quotient2=q2)
        eng.sync()
        outs.append(bufs)
    for x, y in zip(*outs):
        assert torch.equal(x, y), (N, q, key)
    idx = [0, 1, 2, 3, 255, 256, 74 * 256 + 5, B - 1]
    want = o.decrypt_batch(f, fp, e[idx, :N].cpu().numpy().view(np.uint16).astype(np.int64), q, p)
    q1, r1, pv, q2 = outs[0]
    assert np.array_equal(r1[idx].cpu().numpy().view(np.uint16)[:, : N + 1], want["remainder1"])
    assert np.array_equal(q1[idx].cpu().numpy().view(np.uint16)[:, : N + 1], want["quotient1"])
    assert np.array_equal(pv[idx].cpu().numpy()[:, :N], want["value"])
    assert np.array_equal(q2[idx].cpu().numpy()[:, : N + 1], want["quotient2"])
    eng.close()


@pytest.mark.parametrize("cfg,B", [("default167", 333), ("hps509", 70001), ("hps821", 4500), ("hrss701", 1031)])
def test_field_element_wire_format(cfg, B, nb, engines, golden):
    """ntru_encrypt_batch_packed / ntru_decrypt_batch_packed: every array crosses the host link as
    packOutput(maxVal, width, row).expected (index.js:572-596; maxVal = q - 1 modulo q, p - 1 for the small arrays).
    Rows are checked against the oracle's BigInt packOutput of the oracle's own encrypt / decrypt results; the whole
    batch (70001 rows = three pipeline chunks, ragged) against the plain-array calls through the numpy form of the
    packing that tests/test_host.py pins to the oracle.  Includes r drawn on the device and NULL outputs."""
    g, eng = golden(cfg), engines(cfg)
    N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
    rng = np.random.default_rng(B)
    r = o.sample_ternary_rows(B, N, dr, dr, rng).astype(np.uint8)
    m = rng.integers(0, 2, size=(B, N)).astype(np.uint8)
    w = nb.wire
    eng.set_path(nb.PATH_AUTO)
    plain = eng.encrypt_batch(r, m)
    enc = eng.encrypt_batch_packed(w.pack_rows(p - 1, r), w.pack_rows(p - 1, m), return_r=True)
    for k in ENC_KEYS:
        assert np.array_equal(enc[k], w.pack_rows(q - 1, plain[k])), k
    assert np.array_equal(enc["r"], w.pack_rows(p - 1, r))
    dplain = eng.decrypt_batch(plain["value"])
    dec = eng.decrypt_batch_packed(enc["value"])
    for k in DEC_KEYS:
        assert np.array_equal(dec[k], w.pack_rows(p - 1 if k in ("value", "quotient2", "remainder2") else q - 1, dplain[k])), k
    # (not compared with m: at q = 4096 with the golden dr the centred lift wraps for some rows, in the reference too)
    assert np.array_equal(w.unpack_rows(p - 1, N, dec["value"], np.uint8), dplain["value"])
    # straight against the oracle, BigInt for BigInt
    idx = [0, 1, B // 2, B - 1]
    want_e = o.encrypt_batch(g["h"].astype(np.int64), r[idx], m[idx], q)
    want_d = o.decrypt_batch(g["f"].astype(np.int64), g["fp"].astype(np.int64), want_e["value"], q, p)

    def big(words):
        return [sum(int(words[e, x]) << (32 * x) for x in range(8)) for e in range(words.shape[0])]

    for i, b in enumerate(idx):
        for k in ENC_KEYS:
            row = [int(x) for x in want_e[k][i]]
            assert big(enc[k][b]) == o.pack_output(q - 1, len(row), row)["expected"], (k, b)
        for k in DEC_KEYS:
            row = [int(x) for x in want_d[k][i]]
            mv = p - 1 if k in ("value", "quotient2", "remainder2") else q - 1
            assert big(dec[k][b]) == o.pack_output(mv, len(row), row)["expected"], (k, b)
    # value only (NULL witness pointers), and r drawn on the device: r_out is the packed form of the rows the plain call
    # draws under the same key and row numbers
    v_only = eng.encrypt_batch_packed(w.pack_rows(p - 1, r[:300]), w.pack_rows(p - 1, m[:300]), witness=False)
    assert v_only["quotientE"] is None and np.array_equal(v_only["value"], enc["value"][:300])
    key = bytes(range(32))
    eng.set_rng_key(key, 1000)
    a = eng.encrypt_batch(None, m[:300], dr=dr)
    eng.set_rng_key(key, 1000)
    bpk = eng.encrypt_batch_packed(None, w.pack_rows(p - 1, m[:300]), dr=dr)
    assert np.array_equal(bpk["r"], w.pack_rows(p - 1, a["r"])) and np.array_equal(bpk["value"], w.pack_rows(q - 1, a["value"]))
    with pytest.raises(IndexError):
        eng.encrypt_batch_packed(w.pack_rows(p - 1, r[:4]), w.pack_rows(q - 1, m[:4].astype(np.uint16) * 0 + 5))   # wrong element count
