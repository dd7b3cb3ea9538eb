"""Throughput of the other BASELINE configs (device-resident, CUDA events): python scripts/bench_configs.py [names...]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntru_circom_b200 as nb

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"


def timed(fn, iters=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def setup(cfg):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{cfg}.npz")))
    N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
    eng = nb.Engine(N, p, q, 0)
    eng.set_public_key(g["h"]); eng.set_private_key(g["f"], g["fp"])
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    if os.environ.get("DEC1_FORM"):                    # first decrypt product at q <= 2048: 0 auto, 1 byte limbs, 2 fp16 tiles
        eng.set_dec1_form(int(os.environ["DEC1_FORM"]))
    if os.environ.get("EPILOGUE"):                     # 0 auto, 1 two epilogue groups everywhere, 2 one group everywhere
        eng.set_epilogue(int(os.environ["EPILOGUE"]))
    if os.environ.get("SCHEDULE"):                     # 1: the round-1 phase order (cyc + hi) instead of hi, then lo on top
        eng.set_schedule(bool(int(os.environ["SCHEDULE"])))
    return g, eng, N, q, dr


def same_key(cfg, B):
    g, eng, N, q, dr = setup(cfg)
    P = eng.pitch
    r = torch.zeros((B, P), dtype=torch.uint8, device=dev); eng.sample_r_dev(B, dr, 0, r, seed=1)
    m = torch.zeros((B, P), dtype=torch.uint8, device=dev); m[:, :N] = torch.randint(0, 2, (B, N), device=dev, dtype=torch.uint8)
    val = torch.empty((B, P), dtype=torch.int16, device=dev); quo = torch.empty_like(val)
    out = torch.empty((B, P), dtype=torch.uint8, device=dev); q1 = torch.empty_like(val); r1 = torch.empty_like(val); q2 = torch.empty_like(out)
    t_enc = timed(lambda: eng.encrypt_dev(B, r, m, value=val, quotientE=quo))
    t_dec = timed(lambda: eng.decrypt_dev(B, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2))
    ok = bool(torch.equal(out[:, :N], m[:, :N]))
    tot = t_enc + t_dec
    eng.set_timing(True); eng.timing_reset()          # per-kernel CUDA events inside the library
    for _ in range(3):
        eng.encrypt_dev(B, r, m, value=val, quotientE=quo)
        eng.decrypt_dev(B, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2)
    kt = {k: round(v[0] / v[1], 4) for k, v in eng.timing_read().items() if v[1]}
    eng.set_timing(False)
    print(json.dumps({"config": cfg, "mode": "same-key tcgen05", "schedule": "cyc + hi" if os.environ.get("SCHEDULE") == "1" else "hi, then lo on top",
                      "rows": B, "enc_ms": t_enc, "dec_ms": t_dec,
                      "kernel_ms": kt,
                      "ct_per_s": B / (tot * 1e-3), "GBps_14N": 14 * N * B / (tot * 1e-3) / 1e9, "frac_hbm": 14 * N * B / (tot * 1e-3) / 1e9 / HBM,
                      "int8_TOPs_10N2": (10 if q > 256 else 6) * N * N * B / (tot * 1e-3) / 1e12, "roundtrip_equals_message": ok}))
    eng.close()


def distinct_key(cfg, B, path=0, valid_keys=0):
    g, eng, N, q, dr = setup(cfg)
    eng.set_path(path)
    P = eng.pitch
    h = torch.zeros((B, P), dtype=torch.int16, device=dev); f = torch.zeros((B, P), dtype=torch.int8, device=dev)
    fp = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    keygen_s = None
    if valid_keys:
        # VALID key pairs from the batched key generation (ntru_keygen_batch), tiled over the batch
        import random
        rng = random.Random(3)
        k = nb.NTRU({"N": N, "q": q, "p": 3, "df": int(g["df"]), "dg": int(g["dg"]), "dr": dr}, rand32=lambda: rng.getrandbits(32))
        t0 = time.perf_counter(); ks = k.generateKeysBatch(valid_keys); keygen_s = time.perf_counter() - t0
        reps = (B + valid_keys - 1) // valid_keys
        h[:, :N] = torch.from_numpy(np.tile(ks["h"].astype(np.int16), (reps, 1))[:B]).to(dev)
        f[:, :N] = torch.from_numpy(np.tile(ks["f"], (reps, 1))[:B]).to(dev)
        fp[:, :N] = torch.from_numpy(np.tile(ks["fp"], (reps, 1))[:B]).to(dev)
    else:
        h[:, :N] = torch.randint(0, q, (B, N), device=dev, dtype=torch.int16)
        f[:, :N] = torch.randint(-1, 2, (B, N), device=dev, dtype=torch.int8)
        fp[:, :N] = torch.randint(0, 3, (B, N), device=dev, dtype=torch.uint8)
    r = torch.zeros((B, P), dtype=torch.uint8, device=dev); eng.sample_r_dev(B, dr, 0, r, seed=1)
    m = torch.zeros((B, P), dtype=torch.uint8, device=dev); m[:, :N] = torch.randint(0, 2, (B, N), device=dev, dtype=torch.uint8)
    val = torch.empty((B, P), dtype=torch.int16, device=dev); quo = torch.empty_like(val)
    out = torch.empty((B, P), dtype=torch.uint8, device=dev); q1 = torch.empty_like(val); r1 = torch.empty_like(val); q2 = torch.empty_like(out)
    t_enc = timed(lambda: eng.encrypt_dev(B, r, m, value=val, quotientE=quo, h_rows=h), iters=3, warm=1)
    t_dec = timed(lambda: eng.decrypt_dev(B, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2, f_rows=f, fp_rows=fp), iters=3, warm=1)
    tot = t_enc + t_dec
    # the decrypted plaintext is compared with the fp32 CUDA-core schedule on a prefix (oracle parity is tests/)
    nchk = min(B, 4096)
    eng.set_path(1)
    val2 = torch.empty((nchk, P), dtype=torch.int16, device=dev); out2 = torch.empty((nchk, P), dtype=torch.uint8, device=dev)
    eng.encrypt_dev(nchk, r[:nchk], m[:nchk], value=val2, h_rows=h[:nchk])
    eng.decrypt_dev(nchk, val2, value=out2, f_rows=f[:nchk], fp_rows=fp[:nchk])
    torch.cuda.synchronize()
    same = bool(torch.equal(val2[:, :N], val[:nchk, :N]) and torch.equal(out2[:, :N], out[:nchk, :N]))
    roundtrip = bool(torch.equal(out[:, :N], m[:, :N])) if valid_keys else None
    print(json.dumps({"config": cfg, "mode": "distinct-key " + {0: "auto", 1: "CUDA-core fp32", 3: "IMMA"}[path], "rows": B,
                      "keys": f"{valid_keys} valid key pairs (ntru_keygen_batch, {keygen_s:.2f} s incl. host draws), tiled" if valid_keys else "random operands",
                      "roundtrip_equals_message": roundtrip,
                      "matches_cuda_core_schedule": same, "TMAC_per_s_3N2": 3 * N * N * B / (tot * 1e-3) / 1e12, "enc_ms": t_enc, "dec_ms": t_dec,
                      "ct_per_s": B / (tot * 1e-3), "GBps_18N": 18 * N * B / (tot * 1e-3) / 1e9, "frac_hbm": 18 * N * B / (tot * 1e-3) / 1e9 / HBM,
                      "GFMA_per_s": 3 * N * N * B / (tot * 1e-3) / 1e9}))
    eng.close()


def sum_rows(cfg, B):
    g, eng, N, q, dr = setup(cfg)
    P = eng.pitch
    e = torch.randint(0, q, (B, P), device=dev, dtype=torch.int16)
    part = torch.zeros(P, dtype=torch.int32, device=dev); out = torch.empty(P, dtype=torch.int16, device=dev)
    def run():
        part.zero_()
        eng.sum_partial_dev(B, e, part)
        eng.sum_finalize_dev(part, out)
    t = timed(run, iters=5, warm=2)
    want = (e[:, :N].to(torch.int64).sum(dim=0)) % q
    ok = bool(torch.equal(out[:N].to(torch.int64) & 0xFFFF, want))
    print(json.dumps({"config": cfg, "mode": "ciphertext sum", "rows": B, "ms": t, "ct_per_s": B / (t * 1e-3),
                      "GBps_2N": 2 * N * B / (t * 1e-3) / 1e9, "frac_hbm_2N": 2 * N * B / (t * 1e-3) / 1e9 / HBM,
                      "GBps_pitched": 2 * P * B / (t * 1e-3) / 1e9, "matches_torch_sum": ok}))
    eng.close()


which = sys.argv[1:] or ["c1", "c3", "c4", "c5"]
if "c1" in which: same_key("default167", 1 << 20)
if "c2" in which: same_key("hps509", 1 << 20)
if "c3" in which: distinct_key("hps677", 1 << 18)
if "c3v" in which: distinct_key("hps677", 1 << 18, 0, 4096)
if "c3core" in which: distinct_key("hps677", 1 << 16, 1)
if "c3_509" in which: distinct_key("hps509", 1 << 18)
if "c3_821" in which: distinct_key("hps821", 1 << 18)
if "c3_167" in which: distinct_key("default167", 1 << 20)
if "c3s" in which: same_key("hps677", 1 << 20)
if "c4" in which: same_key("hps821", 1 << 20)
if "c5" in which: sum_rows("hrss701", 10_000_000)
if "c5small" in which: sum_rows("hrss701", 1_250_000)
if "c5s" in which: same_key("hrss701", 1 << 20)
