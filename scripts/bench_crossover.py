"""Same-key batches between 1 k and 64 k rows on the IMMA and the tcgen05 schedule (device-resident, CUDA events around
encrypt + decrypt): where does NTRU_OPT_PATH = 0 have to switch?  python scripts/bench_crossover.py [cfg ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntru_circom_b200 as nb  # noqa: E402

dev = "cuda"
for cfg in sys.argv[1:] or ["default167", "hps509", "hps677", "hps821"]:
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{cfg}.npz")))
    N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
    eng = nb.Engine(N, p, q, 0)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_public_key(g["h"])
    eng.set_private_key(g["f"], g["fp"])
    P = eng.pitch
    for B in (1024, 2048, 4096, 8192, 12288, 16384, 24576, 32768, 65536):
        r = torch.zeros((B, P), dtype=torch.uint8, device=dev)
        eng.sample_r_dev(B, dr, 0, r, seed=1)
        m = torch.zeros((B, P), dtype=torch.uint8, device=dev)
        m[:, :N] = torch.randint(0, 2, (B, N), device=dev, dtype=torch.uint8)
        val, quo, q1, r1 = (torch.empty((B, P), dtype=torch.int16, device=dev) for _ in range(4))
        out, q2 = (torch.empty((B, P), dtype=torch.uint8, device=dev) for _ in range(2))
        res = {}
        for path in (nb.PATH_IMMA, nb.PATH_TENSOR):
            eng.set_path(path)

            def step():
                eng.encrypt_dev(B, r, m, value=val, quotientE=quo)
                eng.decrypt_dev(B, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2)

            for _ in range(3):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                step()
            e1.record()
            torch.cuda.synchronize()
            res[path] = e0.elapsed_time(e1) / 10 * 1e3
        print(f"{cfg} B={B:6d}  IMMA {res[nb.PATH_IMMA]:8.1f} us   tcgen05 {res[nb.PATH_TENSOR]:8.1f} us   "
              f"{'IMMA' if res[nb.PATH_IMMA] < res[nb.PATH_TENSOR] else 'tcgen05'}", flush=True)
    eng.close()
