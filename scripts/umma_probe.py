"""Debug probe for the tcgen05 schedule: one small batch per mode vs the oracle, with a mismatch map."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ntru_circom_b200 as nb  # noqa: E402
import ntru_oracle as o  # noqa: E402


def report(name, got, want):
    got = np.asarray(got).astype(np.int64)
    want = np.asarray(want).astype(np.int64)
    bad = got != want
    if not bad.any():
        print(f"  {name}: OK")
        return True
    rows, cols = np.nonzero(bad)
    print(f"  {name}: {bad.sum()} mismatches of {bad.size}; rows {rows.min()}..{rows.max()} ({len(set(rows))} rows), "
          f"cols {cols.min()}..{cols.max()} ({len(set(cols))} cols)")
    for rr, cc in list(zip(rows, cols))[:6]:
        print(f"     [{rr},{cc}] got {got[rr, cc]} want {want[rr, cc]}")
    return False


def main():
    cfgs = sys.argv[1:] or ["hps509"]
    ok = True
    for cfg in cfgs:
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{cfg}.npz")))
        N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
        eng = nb.Engine(N, p, q, 0)
        eng.set_public_key(g["h"])
        eng.set_private_key(g["f"], g["fp"])
        eng.set_path(nb.PATH_TENSOR)
        eng.set_option(4, int(os.environ.get('NTRU_VARIANT', '0')))
        rng = np.random.default_rng(1)
        B = 300
        r = o.sample_ternary_rows(B, N, dr, dr, rng).astype(np.uint8)
        m = rng.integers(0, 2, size=(B, N)).astype(np.uint8)
        want_e = o.encrypt_batch(g["h"].astype(np.int64), r, m, q)
        want_d = o.decrypt_batch(g["f"].astype(np.int64), g["fp"].astype(np.int64), want_e["value"], q, p)
        print(cfg, "encrypt")
        enc = eng.encrypt_batch(r, m)
        for k in ("value", "quotientE", "remainderE"):
            ok &= report(k, enc[k], want_e[k])
        print(cfg, "decrypt")
        dec = eng.decrypt_batch(want_e["value"].astype(np.uint16))
        for k in ("remainder1", "quotient1", "value", "remainder2", "quotient2"):
            ok &= report(k, dec[k], want_d[k])
        eng.close()
    print("ALL OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
