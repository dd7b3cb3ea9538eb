import sys, time, random, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import ntru_circom_b200 as nb, ntru_oracle as o
for cfg, B in (("default167", 4096), ("hps509", 2048), ("hps677", 2048), ("hps821", 1024)):
    c = o.CONFIGS[cfg]; N = c["N"]
    rng = np.random.default_rng(0)
    f = np.stack([rng.permutation(np.array([1] * c["df"] + [-1] * (c["df"] - 1) + [0] * (N - 2 * c["df"] + 1), dtype=np.int8)) for _ in range(B)])
    g = np.stack([rng.permutation(np.array([1] * c["dg"] + [-1] * c["dg"] + [0] * (N - 2 * c["dg"]), dtype=np.int8)) for _ in range(B)])
    eng = nb.Engine(N, 3, c["q"], 0)
    eng.keygen_batch(f[:64], g[:64])
    eng.set_timing(True); eng.timing_reset()
    t0 = time.perf_counter(); out = eng.keygen_batch(f, g); dt = time.perf_counter() - t0
    kt = eng.timing_read()
    gpu_ms = sum(v[0] for v in kt.values())
    print(f"{cfg}: {B} keys in {dt*1e3:.1f} ms = {B/dt:.0f} keys/s (valid {int(out['valid'].sum())}); GPU kernels {gpu_ms:.2f} ms, host Euclid + copies {dt*1e3-gpu_ms:.1f} ms", flush=True)
    eng.close()
