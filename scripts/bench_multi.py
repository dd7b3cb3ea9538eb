"""BASELINE configs 4 and 5 on N GPUs (one process per GPU, torchrun):
  config 4: NTRU-HPS N=821 q=4096, 2^20 ciphertexts under one key sharded contiguously over the ranks (strong scaling,
            no collective on the encrypt/decrypt path);
  config 5: NTRU-HRSS N=701 q=8192, homomorphic sum of 10 000 000 ciphertext rows sharded over the ranks, local column
            sums + ONE all-reduce of N int32 over NCCL/NVLink.
Timing: CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks; rank 0 prints one JSON
line per config.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/bench_multi.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntru_circom_b200 as nb  # noqa: E402
from ntru_circom_b200 import sharding  # noqa: E402

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def engine(cfg):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{cfg}.npz")))
    N, q, p, dr = int(g["N"]), int(g["q"]), int(g["p"]), int(g["dr"])
    eng = nb.Engine(N, p, q, local)
    eng.set_public_key(g["h"])
    eng.set_private_key(g["f"], g["fp"])
    eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    return g, eng, N, q, dr


def config4(total=1 << 20):
    g, eng, N, q, dr = engine("hps821")
    lo, hi = sharding.shard_bounds(total, world, rank)
    B, P = hi - lo, eng.pitch
    r = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    eng.sample_r_dev(B, dr, lo, r, seed=821)                      # counter-based: the same rows at every world size
    gen = torch.Generator(device=dev).manual_seed(99)
    m_all = torch.randint(0, 2, (total, N), generator=gen, device=dev, dtype=torch.uint8)
    m = torch.zeros((B, P), dtype=torch.uint8, device=dev)
    m[:, :N] = m_all[lo:hi]
    del m_all
    val = torch.empty((B, P), dtype=torch.int16, device=dev)
    quo = torch.empty_like(val)
    q1, r1 = torch.empty_like(val), torch.empty_like(val)
    out = torch.empty((B, P), dtype=torch.uint8, device=dev)
    q2 = torch.empty_like(out)

    def step():
        eng.encrypt_dev(B, r, m, value=val, quotientE=quo)
        eng.decrypt_dev(B, val, value=out, quotient1=q1, remainder1=r1, quotient2=q2)

    ms = timed(step, 10)
    # checksum of checksums: identical at every world size (q = 4096: decrypt != message by the reference's lift)
    chk = torch.stack([(val[:, :N].to(torch.int64) & 0xFFFF).sum(), out[:, :N].to(torch.int64).sum()])
    if world > 1:
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    if rank == 0:
        cts = total / (ms * 1e-3)
        print(json.dumps({"config": "hps821 same-key, 2^20 ciphertexts sharded", "n_gpus": world, "scaling": "strong",
                          "rows_per_gpu": B, "ms": ms, "ct_per_s": cts, "GBps_14N_per_gpu": 14 * N * B / (ms * 1e-3) / 1e9,
                          "frac_hbm_per_gpu": 14 * N * B / (ms * 1e-3) / 1e9 / HBM, "collective": "none",
                          "checksum": [int(x) for x in chk.tolist()]}), flush=True)
    eng.close()


def config5(total=10_000_000):
    g, eng, N, q, dr = engine("hrss701")
    lo, hi = sharding.shard_bounds(total, world, rank)
    B, P = hi - lo, eng.pitch
    # uniform rows from a per-row-block seeded generator: the global row set does not depend on the world size
    blk = 250_000
    e = torch.empty((B, P), dtype=torch.int16, device=dev)
    for b0 in range((lo // blk) * blk, hi, blk):
        gen = torch.Generator(device=dev).manual_seed(7_000_000 + b0 // blk)
        rows = torch.randint(0, q, (blk, P), generator=gen, device=dev, dtype=torch.int16)
        s0, s1 = max(b0, lo), min(b0 + blk, hi)
        e[s0 - lo:s1 - lo] = rows[s0 - b0:s1 - b0]
    e[:, N:] = 0
    partial = torch.zeros(P, dtype=torch.int32, device=dev)
    loc = torch.empty(P, dtype=torch.int16, device=dev)
    red = torch.empty(P, dtype=torch.int32, device=dev)
    result = {}

    def step():
        partial.zero_()
        eng.sum_partial_dev(B, e, partial)
        eng.sum_finalize_dev(partial, loc)
        red.copy_(loc)
        red.bitwise_and_(0xFFFF)
        if world > 1:
            dist.all_reduce(red, op=dist.ReduceOp.SUM)           # N int32 over NCCL / NVLink
        result["sum"] = red & (q - 1)

    ms_nccl = timed(step, 10)
    sum_nccl = result["sum"].clone()
    # the product path: column sums fused with the exchange over peer memory (ntru_sum_allreduce_dev)
    sharding.connect_exchange(eng)
    out = torch.empty(P, dtype=torch.int16, device=dev)

    def fused():
        eng.sum_allreduce_dev(B, e, out)
        result["sum"] = out

    ms = timed(fused, 10)
    result["sum"] = (out.to(torch.int32) & 0xFFFF)
    agree = bool(torch.equal(result["sum"][:N], sum_nccl[:N]))
    # local-only time (no collective), to show what the all-reduce costs
    def local_only():
        partial.zero_()
        eng.sum_partial_dev(B, e, partial)
        eng.sum_finalize_dev(partial, loc)
    ms_local = timed(local_only, 10)
    want = torch.zeros(N, dtype=torch.int64, device=dev)
    for b0 in range(0, B, 500_000):
        want += e[b0:b0 + 500_000, :N].to(torch.int64).sum(dim=0)
    if world > 1:
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
    ok = bool(torch.equal(result["sum"][:N].to(torch.int64), want % q))
    if rank == 0:
        print(json.dumps({"config": "hrss701 homomorphic sum of 10M ciphertexts", "n_gpus": world, "scaling": "strong",
                          "rows_per_gpu": B, "ms": ms, "ms_nccl_allreduce_path": ms_nccl, "ms_local_only": ms_local,
                          "peer_exchange_equals_nccl_path": agree, "ct_per_s": total / (ms * 1e-3),
                          "GBps_2N_per_gpu": 2 * N * B / (ms * 1e-3) / 1e9, "frac_hbm_per_gpu": 2 * N * B / (ms * 1e-3) / 1e9 / HBM,
                          "collective": "stores into peer exchange windows over NVLink inside the sum kernel (no NCCL on the data path)"
                          if world > 1 else "none",
                          "matches_int64_column_sums": ok, "checksum": int(result["sum"][:N].sum().item())}), flush=True)
    eng.close()


which = sys.argv[1:] or ["c4", "c5"]
if "c4" in which:
    config4()
if "c5" in which:
    config5()
if world > 1:
    dist.destroy_process_group()
