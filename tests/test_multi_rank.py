"""world_size-2 gloo test of the sharding logic (CPU): shard cover + the sum's single exchange step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ntru_oracle as o


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, N, rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ntru_circom_b200.sharding import all_reduce_sum_mod_q, shard_bounds
    e = np.random.default_rng(5).integers(0, q, size=(rows, N))       # same data on every rank
    b, t = shard_bounds(rows, world, rank)
    local = torch.from_numpy(o.sum_batch(e[b:t], q))                  # stands in for the GPU partial
    got = all_reduce_sum_mod_q(local, q)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), got.numpy())
    dist.destroy_process_group()


class _FakeEngine:
    """Stands in for Engine in the control-plane test: records what connect_exchange hands to the C ABI."""

    def __init__(self, rank):
        self.rank, self.created, self.connected = rank, None, None

    def xchg_create(self, world, rank):
        self.created = (world, rank)
        return bytes([rank + 1]) * 64                      # this rank's 64-byte "IPC handle"

    def xchg_connect(self, handles):
        self.connected = handles


def _exchange_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ntru_circom_b200.sharding import connect_exchange
    eng = _FakeEngine(rank)
    connect_exchange(eng)
    assert eng.created == (world, rank)
    with open(os.path.join(out_dir, f"h{rank}.bin"), "wb") as fh:
        fh.write(eng.connected)
    dist.destroy_process_group()


def test_exchange_handles_are_gathered_rank_major(tmp_path):
    """connect_exchange (control plane of the peer-memory sum): every rank receives all 64-byte handles, rank-major."""
    world, port = 2, _free_port()
    mp.spawn(_exchange_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = b"".join(bytes([r + 1]) * 64 for r in range(world))
    for r in range(world):
        assert (tmp_path / f"h{r}.bin").read_bytes() == want


def test_shard_bounds_cover_everything_once():
    from ntru_circom_b200.sharding import shard_bounds
    for total in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_sum_all_reduce_world2(tmp_path):
    q, N, rows, world = 8192, 701, 1001, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, q, N, rows, str(tmp_path)), nprocs=world, join=True)
    e = np.random.default_rng(5).integers(0, q, size=(rows, N))
    want = o.sum_batch(e, q)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"r{r}.npy"), want)
