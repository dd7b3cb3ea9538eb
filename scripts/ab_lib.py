"""A/B of two builds of libntru_b200.so on the device-resident encrypt / decrypt calls (raw ctypes, only entry points that
exist since round 1, so that an older build loads):  python scripts/ab_lib.py <cfg> <lib.so> [<lib.so> ...]
Prints one JSON line per library: per-kernel times (CUDA events inside the library) at 2^20 rows, median of 5 repeats."""
import ctypes
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KINDS = ["enc_tensor", "dec1_tensor", "dec2_tensor"]


def run(cfg, path, rows=1 << 20):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", cfg + ".npz")))
    N, q, p = int(g["N"]), int(g["q"]), int(g["p"])
    lib = ctypes.CDLL(path)
    V = ctypes.c_void_p
    lib.ntru_create.argtypes = [ctypes.POINTER(V), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.ntru_set_option.argtypes = [V, ctypes.c_int, ctypes.c_long]
    lib.ntru_set_public_key.argtypes = [V, V]
    lib.ntru_set_private_key.argtypes = [V, V, V]
    lib.ntru_set_stream.argtypes = [V, V]
    lib.ntru_pitch.argtypes = [V]
    lib.ntru_encrypt_dev.argtypes = [V, ctypes.c_size_t] + [V] * 6
    lib.ntru_decrypt_dev.argtypes = [V, ctypes.c_size_t] + [V] * 8
    lib.ntru_timing_read.argtypes = [V, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    lib.ntru_timing_reset.argtypes = [V]
    lib.ntru_destroy.argtypes = [V]
    h = V()
    assert lib.ntru_create(ctypes.byref(h), N, p, q, 0) == 0
    hk = np.ascontiguousarray(g["h"], dtype=np.uint16)
    f = np.ascontiguousarray(g["f"], dtype=np.int8)
    fp = np.ascontiguousarray(g["fp"], dtype=np.uint8)
    assert lib.ntru_set_public_key(h, hk.ctypes.data) == 0
    assert lib.ntru_set_private_key(h, f.ctypes.data, fp.ctypes.data) == 0
    assert lib.ntru_set_stream(h, torch.cuda.current_stream().cuda_stream) == 0
    P = lib.ntru_pitch(h)
    dev = "cuda"
    gen = torch.Generator(device=dev).manual_seed(7)
    r = torch.zeros((rows, P), dtype=torch.uint8, device=dev)
    r[:, :N] = torch.randint(0, 3, (rows, N), generator=gen, device=dev, dtype=torch.uint8)
    m = torch.zeros((rows, P), dtype=torch.uint8, device=dev)
    m[:, :N] = torch.randint(0, 2, (rows, N), generator=gen, device=dev, dtype=torch.uint8)
    val, quo, q1, r1 = (torch.empty((rows, P), dtype=torch.int16, device=dev) for _ in range(4))
    out, q2 = (torch.empty((rows, P), dtype=torch.uint8, device=dev) for _ in range(2))

    def step():
        assert lib.ntru_encrypt_dev(h, rows, None, r.data_ptr(), m.data_ptr(), val.data_ptr(), quo.data_ptr(), None) == 0
        assert lib.ntru_decrypt_dev(h, rows, None, None, val.data_ptr(), out.data_ptr(), q1.data_ptr(), r1.data_ptr(), q2.data_ptr(), None) == 0

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    lib.ntru_set_option(h, 3, 1)
    reps = {k: [] for k in KINDS}
    for _ in range(5):
        lib.ntru_timing_reset(h)
        for _ in range(4):
            step()
        for i, k in enumerate(KINDS):
            ms, n = ctypes.c_double(0), ctypes.c_uint64(0)
            lib.ntru_timing_read(h, i, ctypes.byref(ms), ctypes.byref(n))
            if n.value:
                reps[k].append(ms.value / n.value)
    med = {k: round(statistics.median(v), 4) for k, v in reps.items() if v}
    chk = int(out[:, :N].to(torch.int64).sum().item()) ^ int(q1.to(torch.int64).sum().item())
    lib.ntru_destroy(h)
    return {"lib": os.path.relpath(path, ROOT), "config": cfg, "rows": rows, "kernel_ms": med, "step_ms": round(sum(med.values()), 4),
            "ct_per_s": rows / (sum(med.values()) * 1e-3), "checksum": chk}


if __name__ == "__main__":
    cfg = sys.argv[1]
    for path in sys.argv[2:]:
        print(json.dumps(run(cfg, os.path.abspath(path))), flush=True)
