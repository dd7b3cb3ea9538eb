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


def sum_ciphertexts_sharded(engine, e_dev, rows: int, group=None):
    """Column sums mod q of a sharded batch.  e_dev: this rank's (rows, pitch) uint16/int16 CUDA tensor."""
    import torch

    partial = torch.zeros(engine.pitch, dtype=torch.int32, device=e_dev.device)
    local = torch.empty(engine.pitch, dtype=torch.int16, device=e_dev.device)
    engine.sum_partial_dev(rows, e_dev, partial)
    engine.sum_finalize_dev(partial, local)
    engine.sync()
    return all_reduce_sum_mod_q(local.to(torch.int32) & 0xFFFF, engine.q, group)[: engine.N]
