"""Debug: build with -DNTRU_TRACE, run ENC on a small batch, print the per-chunk timeline of cluster 0."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "ntru-circom_b200", "csrc")
lib_path = os.path.join(ROOT, "ntru-circom_b200", "libntru_trace.so")   # *.so: git-ignored, travels with gpurun
if not os.path.exists(lib_path) or "--build" in sys.argv:
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DNTRU_TRACE", "-Xcompiler", "-fPIC",
                    "-shared", "-o", lib_path] + [os.path.join(csrc, f) for f in ("api.cu", "generic_kernels.cu", "imma_kernels.cu", "keygen.cu", "umma_kernels.cu")], check=True)
import ntru_circom_b200 as nb
from ntru_circom_b200 import _lib
_lib.LIB_PATH = lib_path
if "--build" in sys.argv:
    sys.exit(0)
mode = sys.argv[1] if len(sys.argv) > 1 else "enc"
cfg = os.environ.get("TRACE_CFG", "hps509")
g = dict(np.load(os.path.join(ROOT, "tests", "golden", cfg + ".npz")))
NN, QQ = int(g["N"]), int(g["q"])
eng = nb.Engine(NN, 3, QQ, 0)
eng.set_public_key(g["h"]); eng.set_private_key(g["f"], g["fp"])
rows = int(os.environ.get('TRACE_ROWS', 74 * 256 * 6))
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
    n = dump(buf.ctypes.data, buf.size)
    if mode != "enc":
        eng.decrypt_dev(rows, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2); eng.sync()
        n = dump(buf.ctypes.data, buf.size)
recs = []
for role in range(LANES):
    for w in buf[role * CAP:(role + 1) * CAP]:
        w = int(w)
        if w == 0:
            break
        tag, clk = w >> 40, w & 0xffffffffff
        recs.append((role + 1, tag >> 12, tag & 0xfff, clk))
marks = {e: t for (role, e, idx, t) in recs if role == 3 and e in (10, 11, 12, 13)}
if len(marks) == 4:
    cyc, ns = marks[12] - marks[10], marks[13] - marks[11]
    print(f"producer warp lifetime: {cyc} SM cycles in {ns} ns -> {cyc / ns * 1000:.0f} MHz")
recs = [x for x in recs if not (x[0] == 3 and x[1] >= 10)]
rec = np.array(recs, dtype=np.int64)
rec[:, 3] -= rec[:, 3].min()
rec = rec[np.argsort(rec[:, 3], kind="stable")]
nm = {(4,0):"xform loads issued / waits raw",(4,1):"xform got a_empty / raw_full",(4,5):"xform arrived",(1,0):"mma chunk start",(1,1):"mma got tempty",(1,4):"mma wait b_full",(1,2):"mma got b_full",(1,3):"mma got a_full",(1,5):"mma issued",
      (2,0):"epi start",(2,1):"epi got tfull",(2,2):"epi done",(2,3):"epi got m_full",(2,4):"epi tmem loaded",(2,5):"epi math done",(2,6):"epi staged (after wait_group.read + STS)",(2,7):"epi fenced",(3,0):"prod wait b_empty",(3,1):"prod got b_empty",(3,2):"prod issued"}
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 19)
prev = {}
for role, e, idx, t in rec:
    if "--mma" in sys.argv and role not in (1, 4):
        continue
    if "--ring" in sys.argv and not (role == 3 or (role == 1 and e in (2, 4, 5))):
        continue
    xlo, xhi = int(os.environ.get("TRACE_XLO", -1)), int(os.environ.get("TRACE_XHI", -1))
    if (role != 4 and lo <= idx <= hi) or (role == 4 and xlo <= idx <= xhi):
        d = t - prev.get(role, t)
        prev[role] = t
        print(f"{t:9d} (+{d:5d})  cc={idx:3d}  {nm.get((role,e),(role,e))}")
