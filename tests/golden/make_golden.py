"""Generates tests/golden/*.npz with the oracle (oracle/ntru_oracle.py), seeds recorded in the files.

The reference (JavaScript) cannot run in this image and ships no fixed-key
vectors, so these goldens are produced by the restated oracle after it has been
pinned against the upstream KATs (tests/test_oracle_kats.py).  Every vector is
generated through the *literal* path (float64 FFT + long division) for N <= 167
and through the C restatement (same literal algorithm) for the larger sets, and
must also satisfy the restated VerifyEncrypt / VerifyDecrypt constraints.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import c_oracle  # noqa: E402
import ntru_oracle as o  # noqa: E402

ROWS = {"tiny17": 8, "default167": 6, "hps509": 4, "hps677": 3, "hps821": 3, "hrss701": 3}
KEY_SEED, DATA_SEED = 20261018, 7


def main():
    for cfg, B in ROWS.items():
        key = o.make_key(cfg, KEY_SEED, literal=cfg in ("tiny17", "default167"))
        N, q, p = key.N, key.q, key.p
        rng = np.random.default_rng(DATA_SEED)
        r = o.sample_ternary_rows(B, N, key.dr, key.dr, rng)
        m = rng.integers(0, 2, size=(B, N))
        m[0, :] = 0                       # all-zero message
        m[1, N // 2:] = 0                 # ragged (short) message
        if B > 2:
            m[2] = rng.integers(0, 3, size=N)   # ternary plaintext (test/reference.test.js:50)
        h = np.array(o.expand_array(key.h, N), dtype=np.int64)
        f = np.array(key.f, dtype=np.int64)
        fp = np.array(o.expand_array(key.fp, N), dtype=np.int64)
        enc = c_oracle.encrypt_batch(h, r, m, q)
        dec = c_oracle.decrypt_batch(f, fp, enc["value"], q, p)
        # cross-check: closed form, literal python (small N), circuit constraints
        enc2 = o.encrypt_batch(h, r, m, q)
        dec2 = o.decrypt_batch(f, fp, enc2["value"], q, p)
        for k in ("value", "quotientE", "remainderE"):
            assert np.array_equal(enc[k], enc2[k]), (cfg, k)
        for k in ("value", "quotient1", "remainder1", "quotient2", "remainder2"):
            assert np.array_equal(dec[k], dec2[k]), (cfg, k)
        lit = o.NTRU(dict(o.CONFIGS[cfg], f=key.f, fp=key.fp, fq=key.fq, g=key.g, h=key.h),
                     literal=N <= 167)
        for b in range(B):
            eb = lit.encryptBits(o.trim_polynomial(m[b].tolist()), r[b].tolist())
            assert eb["inputs"]["remainderE"] == enc["remainderE"][b].tolist()
            assert eb["inputs"]["quotientE"] == enc["quotientE"][b].tolist()
            assert o.verify_encrypt(eb["inputs"], eb["params"])
            db = lit.decryptBits(eb["value"])
            assert db["inputs"]["remainder2"] == dec["remainder2"][b].tolist()
            assert o.verify_decrypt(db["inputs"], db["params"])
        np.savez_compressed(
            os.path.join(HERE, f"{cfg}.npz"),
            N=N, q=q, p=p, df=key.df, dg=key.dg, dr=key.dr, key_seed=KEY_SEED, data_seed=DATA_SEED,
            f=f.astype(np.int8), fp=fp.astype(np.uint8), fq=np.array(o.expand_array(key.fq, N), dtype=np.uint16),
            g=np.array(key.g, dtype=np.int8), h=h.astype(np.uint16),
            r=r.astype(np.uint8), m=m.astype(np.uint8),
            value=enc["value"].astype(np.uint16), quotientE=enc["quotientE"].astype(np.uint16),
            remainderE=enc["remainderE"].astype(np.uint16),
            dec_value=dec["value"].astype(np.uint8), quotient1=dec["quotient1"].astype(np.uint16),
            remainder1=dec["remainder1"].astype(np.uint16), quotient2=dec["quotient2"].astype(np.uint8),
            remainder2=dec["remainder2"].astype(np.uint8),
            sum=o.sum_batch(enc["value"], q).astype(np.uint16),
            roundtrip_ok=np.array_equal(dec["value"], m),
        )
        print(cfg, "rows", B, "roundtrip", bool(np.array_equal(dec["value"], m)))


if __name__ == "__main__":
    main()
