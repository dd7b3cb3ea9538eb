"""Small device-resident distinct-key enc+dec run for ncu: python scripts/profile_distinct.py [rows] [iters] [cfg] [path]."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntru_circom_b200 as nb  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16 * 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = sys.argv[3] if len(sys.argv) > 3 else "hps677"
path = int(sys.argv[4]) if len(sys.argv) > 4 else 0
g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{cfg}.npz")))
N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
eng = nb.Engine(N, p, q, 0)
eng.set_path(path)
P = eng.pitch
dev = "cuda"
h = torch.zeros((rows, P), dtype=torch.int16, device=dev)
h[:, :N] = torch.randint(0, q, (rows, N), device=dev, dtype=torch.int16)
f = torch.zeros((rows, P), dtype=torch.int8, device=dev)
f[:, :N] = torch.randint(-1, 2, (rows, N), device=dev, dtype=torch.int8)
fp = torch.zeros((rows, P), dtype=torch.uint8, device=dev)
fp[:, :N] = torch.randint(0, 3, (rows, N), device=dev, dtype=torch.uint8)
r = torch.zeros((rows, P), dtype=torch.uint8, device=dev)
eng.sample_r_dev(rows, dr, 0, r, seed=1)
m = torch.zeros((rows, P), dtype=torch.uint8, device=dev)
m[:, :N] = torch.randint(0, 2, (rows, N), device=dev, dtype=torch.uint8)
val = torch.empty((rows, P), dtype=torch.int16, device=dev)
quo = torch.empty((rows, P), dtype=torch.int16, device=dev)
out = torch.empty((rows, P), dtype=torch.uint8, device=dev)
q1 = torch.empty((rows, P), dtype=torch.int16, device=dev)
r1 = torch.empty((rows, P), dtype=torch.int16, device=dev)
q2 = torch.empty((rows, P), dtype=torch.uint8, device=dev)
for _ in range(iters):
    eng.encrypt_dev(rows, r, m, value=val, quotientE=quo, h_rows=h)
    eng.decrypt_dev(rows, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2, f_rows=f, fp_rows=fp)
eng.sync()
torch.cuda.synchronize()
print("ok", rows, iters, "path", eng.last_path)
