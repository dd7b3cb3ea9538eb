"""Clock-trace summary of the tcgen05 pair kernel (trace build, cluster 0 / CTA 0): where the MMA issuer and epilogue warp 0
spend their cycles, per tile in steady state.  python scripts/trace_summary.py <enc|dec1|dec2> ; TRACE_CFG=hps821"""
import ctypes, os, subprocess, sys
from collections import defaultdict
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "ntru-circom_b200", "csrc")
lib_path = os.path.join(ROOT, "ntru-circom_b200", "libntru_trace.so")   # *.so: git-ignored, travels with gpurun
if not os.path.exists(lib_path) or "--build" in sys.argv:
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DNTRU_TRACE", "-Xcompiler", "-fPIC",
                    "-shared", "-o", lib_path] + [os.path.join(csrc, f) for f in ("api.cu", "generic_kernels.cu", "imma_kernels.cu", "keygen.cu", "umma_kernels.cu")], check=True)
if "--build" in sys.argv:
    sys.exit(0)
import ntru_circom_b200 as nb
from ntru_circom_b200 import _lib
_lib.LIB_PATH = lib_path
mode = sys.argv[1] if len(sys.argv) > 1 else "enc"
if mode == "dec1":
    os.environ["NTRU_TRACE_DEC1_ONLY"] = "1"
cfg = os.environ.get("TRACE_CFG", "hps509")
g = dict(np.load(os.path.join(ROOT, "tests", "golden", cfg + ".npz")))
NN, QQ = int(g["N"]), int(g["q"])
eng = nb.Engine(NN, 3, QQ, 0)
if os.environ.get("SCHEDULE"):
    eng.set_schedule(bool(int(os.environ["SCHEDULE"])))
if os.environ.get("DEC1_FORM"):
    eng.set_dec1_form(int(os.environ["DEC1_FORM"]))
eng.set_public_key(g["h"]); eng.set_private_key(g["f"], g["fp"])
rows = int(os.environ.get("TRACE_ROWS", 74 * 256 * 8))
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
lanes = []
for role in range(LANES):
    ev = []
    for w in buf[role * CAP:(role + 1) * CAP]:
        w = int(w)
        if w == 0:
            break
        ev.append(((w >> 52), (w >> 40) & 0xfff, w & 0xffffffffff))     # (event, idx, clock)
    lanes.append(ev)
mma, epi = lanes[0], lanes[1]
# phases per tile: distance between two "mma chunk start" events whose phase index is a multiple of the count
ccs = sorted({i for e, i, t in mma if e == 0})
starts = {i: t for e, i, t in mma if e == 0}
# the tile length in phases: the engine reports it through the number of phases until the pattern of slice counts repeats
slices = defaultdict(int)
for e, i, t in mma:
    if e == 5:
        slices[i] += 1
pat = [slices[i] for i in ccs]
nph = next(n for n in range(1, len(pat)) if pat[:n] * (len(pat) // n) == pat[:n * (len(pat) // n)])
tiles = len(ccs) // nph
print(f"{cfg} {mode}: {nph} phases per tile, slices per phase {pat[:nph]}, {tiles} tiles traced")
t_tile = [(starts[(k + 1) * nph] - starts[k * nph]) for k in range(1, tiles - 1)]
print(f"cycles per tile (steady state): median {int(np.median(t_tile))}  -> {rows / (74 * 256) * np.median(t_tile) / 1.965e6:.3f} ms at 1965 MHz for {rows} rows")
# issuer breakdown
wait_tmem = defaultdict(int); wait_ops = defaultdict(int); issue = defaultdict(int)
prev = None
for e, i, t in mma:
    if prev is not None:
        pe, pi, pt = prev
        d = t - pt
        if e == 1: wait_tmem[i] += d
        elif e == 3: wait_ops[i] += d
        elif e == 5: issue[i] += d
    prev = (e, i, t)
def per_phase(dct):
    return [int(np.median([dct[k * nph + j] for k in range(1, tiles - 1)])) for j in range(nph)]
print("issuer, per phase of a tile (cycles):")
print("  wait tmem_empty :", per_phase(wait_tmem), " sum", sum(per_phase(wait_tmem)))
print("  wait A/B full   :", per_phase(wait_ops), " sum", sum(per_phase(wait_ops)))
print("  issue           :", per_phase(issue), " sum", sum(per_phase(issue)))
# epilogue warp 0 (group 0): phases with even index
names = {1: "wait tmem_full", 3: "wait m_full", 4: "tmem loads", 5: "arithmetic", 6: "wait_group.read + staging stores", 7: "fence + syncwarp", 2: "store issue / end", 0: "loop head"}
acc = defaultdict(lambda: defaultdict(int))
prev = None
for e, i, t in epi:
    if prev is not None and prev[1] == i:
        acc[i][e] += t - prev[2]
    elif prev is not None and e == 0:
        acc[i][0] += t - prev[2]          # from the end of the previous owned phase to this loop head (store issue of the last pass)
    prev = (e, i, t)
own = sorted(acc)
print("epilogue warp 0, per owned phase of a tile (cycles):")
for e in (1, 3, 4, 5, 6, 7, 2, 0):
    row = []
    for j in range(nph):
        vals = [acc[k * nph + j][e] for k in range(1, tiles - 1) if (k * nph + j) in acc]
        row.append(int(np.median(vals)) if vals else -1)
    print(f"  {names[e]:34s}:", row, " sum", sum(x for x in row if x > 0))
eng.close()
