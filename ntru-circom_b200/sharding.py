"""Multi-GPU layout of the hot path: one process per GPU, contiguous batch shards, keys replicated.

Encrypt / decrypt need no data-path collective (independent ciphertexts).  The homomorphic
ciphertext sum has one exchange step: every rank reduces its shard to N column sums mod q, and
one all-reduce of those N integers (NCCL over NVLink on GPUs, gloo in the CPU tests) finishes it.
"""
from __future__ import annotations

from typing import Tuple


def shard_bounds(total_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `rank`'s shard; sizes differ by at most one row."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(total_rows, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def all_reduce_sum_mod_q(local_sums, q: int, group=None):
    """local_sums: integer tensor of N column sums already reduced into [0, q) on this rank.

    Returns the global column sums mod q on every rank.  Values stay below world_size * q <= 2^18,
    so int32 addition in the collective cannot overflow.
    """
    import torch
    import torch.distributed as dist

    t = local_sums.to(torch.int32).contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t & (q - 1)


def connect_exchange(engine, group=None) -> None:
    """Sets up the peer-memory exchange of the cross-GPU sum: every rank allocates its window
    (ntru_xchg_create), the 64-byte CUDA IPC handles are all-gathered over the process group (control
    plane only), and every rank maps the windows of its peers (ntru_xchg_connect)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        engine.xchg_create(1, 0)
        return
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = engine.xchg_create(world, rank)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    engine.xchg_connect(b"".join(bytes(g.cpu().tolist()) for g in gathered))
    dist.barrier(group=group)


def disconnect_exchange(engine, group=None) -> None:
    """Collective teardown of the exchange: every rank finishes its last sum (sync), all ranks meet at a barrier --
    peers store into a rank's window until THEIR last call has completed -- and only then the windows are unmapped
    and freed (ntru_xchg_destroy; raises if a peer had timed out inside a sum kernel)."""
    import torch.distributed as dist

    engine.sync()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)
    engine.xchg_destroy()


def sum_ciphertexts_exchange(engine, e_dev, rows: int, out=None):
    """Column sums mod q of a sharded batch over the peer-memory exchange (after connect_exchange): local
    column sums, stores into every peer's window over NVLink, one-CTA gather -- all in the library's kernels."""
    import torch

    if out is None:
        out = torch.empty(engine.pitch, dtype=torch.int16, device=e_dev.device)
    engine.sum_allreduce_dev(rows, e_dev, out)
    return out


def sum_ciphertexts_sharded(engine, e_dev, rows: int, group=None):
    """Column sums mod q of a sharded batch through a library all-reduce (NCCL on GPUs, gloo in the CPU tests);
    the baseline the peer-memory exchange above is measured against.
    e_dev: this rank's (rows, pitch) uint16/int16 CUDA tensor."""
    import torch

    partial = torch.zeros(engine.pitch, dtype=torch.int32, device=e_dev.device)
    local = torch.empty(engine.pitch, dtype=torch.int16, device=e_dev.device)
    engine.sum_partial_dev(rows, e_dev, partial)
    engine.sum_finalize_dev(partial, local)
    engine.sync()
    return all_reduce_sum_mod_q(local.to(torch.int32) & 0xFFFF, engine.q, group)[: engine.N]
