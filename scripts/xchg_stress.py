"""Stress of the cross-GPU ciphertext sum (ntru_sum_allreduce_dev) on N real GPUs, one process per GPU:
many calls back to back without host synchronisation, uneven shards that change from call to call, ranks with no rows,
two parameter sets; every result against int64 column sums (all-reduced over NCCL) and against the NCCL path.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/xchg_stress.py
Rank 0 prints "xchg_stress ok" when every check passed on every rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntru_circom_b200 as nb  # noqa: E402
from ntru_circom_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bad = 0
for N, q in ((701, 8192), (167, 128)):
    eng = nb.Engine(N, 3, q, local)
    eng.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    P = eng.pitch
    sharding.connect_exchange(eng)
    rows_max = 60_000
    gen = torch.Generator(device=dev).manual_seed(1000 * N + rank)
    e = torch.randint(0, q, (rows_max, P), generator=gen, device=dev, dtype=torch.int16)
    e[:, N:] = 0
    calls = 40
    # shard size of (call, rank): uneven, changing, with empty shards (also on every rank at once in call 7)
    size = lambda c, r: 0 if (c % 5 == r % 5 or c == 7) else (1 + (c * 7919 + r * 104729) % rows_max)   # noqa: E731
    outs = [torch.full((P,), 7, dtype=torch.int16, device=dev) for _ in range(calls)]
    for c in range(calls):                                   # back to back, no host synchronisation in between
        eng.sum_allreduce_dev(size(c, rank), e, outs[c])
    eng.sync()                                               # raises if a peer timed out
    for c in range(calls):
        want = e[: size(c, rank), :N].to(torch.int64).sum(dim=0)
        dist.all_reduce(want, op=dist.ReduceOp.SUM)
        got = outs[c].to(torch.int64) & 0xFFFF
        if not torch.equal(got[:N], want % q) or bool(got[N:].any()):
            bad += 1
            print(f"rank {rank}: N={N} call {c} differs", flush=True)
        nccl = sharding.sum_ciphertexts_sharded(eng, e, size(c, rank)) if size(c, rank) else None
        if nccl is None:                                     # an empty shard still joins the all-reduce
            z = torch.zeros(P, dtype=torch.int32, device=dev)
            nccl = sharding.all_reduce_sum_mod_q(z, q)[:N]
        if not torch.equal(nccl.to(torch.int64), want % q):
            bad += 1
            print(f"rank {rank}: N={N} call {c}: NCCL path differs", flush=True)
    sharding.disconnect_exchange(eng)
    eng.close()
t = torch.tensor([bad], device=dev)
dist.all_reduce(t)
if rank == 0:
    print("xchg_stress ok" if int(t.item()) == 0 else f"xchg_stress FAILED: {int(t.item())} mismatches", flush=True)
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
