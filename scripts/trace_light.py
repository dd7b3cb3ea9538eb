"""Light profile of the MMA issuer (trace build, NTRU_TRACE_LIGHT=1): cycle sums kept in registers -- three 32-bit clock
reads per slice instead of per-event stores -- so the numbers are close to the production build's.
python scripts/trace_light.py ; prints per kernel and config: cycles per tile, and per slice / phase where they go."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NTRU_TRACE_LIGHT"] = "1"
os.environ["NTRU_B200_LIB"] = os.path.join(ROOT, "ntru-circom_b200", "libntru_trace.so")
import ntru_circom_b200 as nb
for cfg in sys.argv[1:] or ["hps509", "hps821"]:
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", cfg + ".npz")))
    NN, QQ = int(g["N"]), int(g["q"])
    eng = nb.Engine(NN, 3, QQ, 0)
    eng.set_public_key(g["h"]); eng.set_private_key(g["f"], g["fp"])
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    rows = 74 * 256 * 16
    P = eng.pitch
    r = torch.zeros((rows, P), dtype=torch.uint8, device="cuda"); eng.sample_r_dev(rows, int(g["dr"]), 0, r, seed=1)
    m = torch.zeros((rows, P), dtype=torch.uint8, device="cuda"); m[:, :NN] = torch.randint(0, 2, (rows, NN), device="cuda", dtype=torch.uint8)
    val = torch.empty((rows, P), dtype=torch.int16, device="cuda"); quo = torch.empty_like(val)
    out = torch.empty((rows, P), dtype=torch.uint8, device="cuda"); q1 = torch.empty_like(val); r1 = torch.empty_like(val); q2 = torch.empty_like(out)
    buf = np.zeros(4 * 2048, dtype=np.uint64)
    dump = eng.lib.ntru_debug_trace_dump
    dump.argtypes = [ctypes.c_void_p, ctypes.c_uint]

    def show(name):
        dump(buf.ctypes.data, buf.size)
        if int(buf[0]) >> 48 != 0x4C50:
            print(cfg, name, "no light profile in the buffer"); return
        tot, tmem, ops, issue, other, slices, phases = (int(x) for x in buf[1:8])
        tiles = rows // (74 * 256)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        print(f"{cfg} {name}: {tot / tiles:8.0f} cycles per tile, {slices // tiles} slices + {phases // tiles} phases | per slice: operand wait {ops / slices:5.0f}, "
              f"issue {issue / slices:5.0f}, loop {other / (slices + phases):5.0f} | per phase: tmem_empty wait {tmem / phases:6.0f} | "
              f"issuer busy or polling ready barriers {(ops + issue + other) / tot:.0%}, waiting for TMEM {tmem / tot:.0%}")

    for _ in range(2):
        eng.encrypt_dev(rows, r, m, value=val, quotientE=quo); eng.sync()
    show("ENC")
    os.environ["NTRU_TRACE_DEC1_ONLY"] = "1"
    for _ in range(2):
        eng.decrypt_dev(rows, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2); eng.sync()
    show("DEC1")
    del os.environ["NTRU_TRACE_DEC1_ONLY"]
    for _ in range(2):
        eng.decrypt_dev(rows, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2); eng.sync()
    show("DEC2")
    eng.close()
