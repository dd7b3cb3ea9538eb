"""Timing experiments on the CTA-pair kernel's epilogue (trace build, compile-time DBG variants of k_umma_pair):
python scripts/epi_experiments.py [cfg]   -- prints ENC / DEC ms per 2^20 rows and the SM clock seen by cluster 0
for NTRU_DEBUG_EPI in {0, 1, 2, 4, 8, 12} (see umma_pair.cuh).  Results of the DBG variants are wrong on purpose."""
import ctypes, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib_path = os.path.join(ROOT, "ntru-circom_b200", "libntru_trace.so")
import ntru_circom_b200 as nb
from ntru_circom_b200 import _lib
_lib.LIB_PATH = lib_path
cfg = sys.argv[1] if len(sys.argv) > 1 else "hps509"
g = dict(np.load(os.path.join(ROOT, "tests", "golden", cfg + ".npz")))
N, q = int(g["N"]), int(g["q"])
eng = nb.Engine(N, 3, q, 0)
eng.set_public_key(g["h"]); eng.set_private_key(g["f"], g["fp"])
eng.set_stream(torch.cuda.current_stream().cuda_stream)
B, P = 1 << 20, eng.pitch
dev = "cuda"
r = torch.zeros((B, P), dtype=torch.uint8, device=dev); eng.sample_r_dev(B, int(g["dr"]), 0, r, seed=1)
m = torch.zeros((B, P), dtype=torch.uint8, device=dev); m[:, :N] = torch.randint(0, 2, (B, N), device=dev, dtype=torch.uint8)
val = torch.empty((B, P), dtype=torch.int16, device=dev); quo = torch.empty_like(val)
out = torch.empty((B, P), dtype=torch.uint8, device=dev); q1 = torch.empty_like(val); r1 = torch.empty_like(val); q2 = torch.empty_like(out)
dump = eng.lib.ntru_debug_trace_dump
dump.argtypes = [ctypes.c_void_p, ctypes.c_uint]
buf = np.zeros(4 * 448, dtype=np.uint64)


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def clock_mhz():
    dump(buf.ctypes.data, buf.size)
    marks = {}
    for w in buf[2 * 448:3 * 448]:
        w = int(w)
        if w:
            marks[(w >> 40) >> 12] = w & 0xffffffffff
    if all(k in marks for k in (10, 11, 12, 13)):
        return (marks[12] - marks[10]) / (marks[13] - marks[11]) * 1000
    return float("nan")


for dbg in (0, 1, 8, 16):
    os.environ["NTRU_DEBUG_EPI"] = str(dbg)
    t_enc = timed(lambda: eng.encrypt_dev(B, r, m, value=val, quotientE=quo))
    mhz = clock_mhz()
    t_dec = timed(lambda: eng.decrypt_dev(B, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2))
    print(f"DBG={dbg:2d}  enc {t_enc:.3f} ms ({mhz:.0f} MHz in cluster 0)   dec {t_dec:.3f} ms", flush=True)
