"""Latency of the one-ciphertext calls of the drop-in API (encryptBits / decryptBits = batch of one row through the host ABI)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntru_circom_b200 as nb

for cfg in ("default167", "hps509", "hps821"):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", cfg + ".npz")))
    N, q = int(g["N"]), int(g["q"])
    eng = nb.Engine(N, 3, q, 0)
    eng.set_public_key(g["h"]); eng.set_private_key(g["f"], g["fp"])
    r, m = g["r"][:1].astype(np.uint8), g["m"][:1].astype(np.uint8)
    for B in (1, 64, 1024):
        rr, mm = np.repeat(r, B, axis=0), np.repeat(m, B, axis=0)
        for path, name in ((0, "auto"), (3, "imma"), (1, "fp32")):
            eng.set_path(path)
            for _ in range(3):                      # first-use costs (allocation growth, lazy module load) out of the timing
                enc = eng.encrypt_batch(rr, mm)
                dec = eng.decrypt_batch(enc["value"])
            t0 = time.perf_counter()
            for _ in range(50):
                enc = eng.encrypt_batch(rr, mm)
            t1 = time.perf_counter()
            for _ in range(50):
                dec = eng.decrypt_batch(enc["value"])
            t2 = time.perf_counter()
            print(f"{cfg} B={B:5d} {name:5s}: encrypt {(t1 - t0) / 50 * 1e6:8.1f} us  decrypt {(t2 - t1) / 50 * 1e6:8.1f} us", flush=True)
    eng.close()
