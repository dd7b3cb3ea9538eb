"""Per-slice timeline of the tcgen05 pair kernel (trace build, NTRU_TRACE_SLICES=1): when the MMA issuer got the operands of
slice s and issued it, when the key-matrix producer of the same CTA started waiting for the stage of slice s and got it.
python scripts/trace_slices.py <enc|dec2> [first slice] [count];  TRACE_CFG=hps821"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NTRU_TRACE_SLICES"] = "1"
import ntru_circom_b200 as nb
from ntru_circom_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "ntru-circom_b200", "libntru_trace.so")
mode = sys.argv[1] if len(sys.argv) > 1 else "enc"
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cnt = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cfg = os.environ.get("TRACE_CFG", "hps821")
g = dict(np.load(os.path.join(ROOT, "tests", "golden", cfg + ".npz")))
NN, QQ = int(g["N"]), int(g["q"])
eng = nb.Engine(NN, 3, QQ, 0)
eng.set_public_key(g["h"]); eng.set_private_key(g["f"], g["fp"])
rows = 74 * 256 * 8
P = eng.pitch
r = torch.zeros((rows, P), dtype=torch.uint8, device="cuda"); eng.sample_r_dev(rows, int(g["dr"]), 0, r, seed=1)
m = torch.zeros((rows, P), dtype=torch.uint8, device="cuda"); m[:, :NN] = torch.randint(0, 2, (rows, NN), device="cuda", dtype=torch.uint8)
val = torch.empty((rows, P), dtype=torch.int16, device="cuda"); quo = torch.empty_like(val)
out = torch.empty((rows, P), dtype=torch.uint8, device="cuda"); q1 = torch.empty_like(val); r1 = torch.empty_like(val); q2 = torch.empty_like(out)
LANES, CAP = 4, 2048
buf = np.zeros(LANES * CAP, dtype=np.uint64)
dump = eng.lib.ntru_debug_trace_dump
dump.argtypes = [ctypes.c_void_p, ctypes.c_uint]
for _ in range(2):
    eng.encrypt_dev(rows, r, m, value=val, quotientE=quo); eng.sync()
    if mode != "enc":
        eng.decrypt_dev(rows, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2); eng.sync()
    dump(buf.ctypes.data, buf.size)
ev = {}
for role in (0, 2):
    for w in buf[role * CAP:(role + 1) * CAP]:
        w = int(w)
        if w == 0:
            break
        e, idx, t = w >> 52, (w >> 40) & 0xfff, w & 0xffffffffff
        ev.setdefault((role, e, idx), t)
t0 = min(v for (ro, e, i), v in ev.items() if i == s0)
print(f"{cfg} {mode}: slice | issuer got operands, issued (+cycles busy) | producer waits for the stage, gets it (+cycles waited) | issue -> stage free again")
prev_issue = None
for s in range(s0, s0 + cnt):
    g3, g5 = ev.get((0, 3, s)), ev.get((0, 5, s))
    p0, p1 = ev.get((2, 0, s)), ev.get((2, 1, s))
    if None in (g3, g5):
        continue
    line = f"{s:5d} | got {g3 - t0:7d} issued {g5 - t0:7d} (+{g5 - g3:4d})" + (f" since last issue {g5 - prev_issue:5d}" if prev_issue else " " * 23)
    if p0 is not None and p1 is not None:
        line += f" | prod wait {p0 - t0:7d} got {p1 - t0:7d} (+{p1 - p0:5d})"
    print(line)
    prev_issue = g5
eng.close()
