"""The N>1 path on CPU: two gloo ranks compute their shards of one batch; together they must partition it, with the
same instrument mix on every rank, and the timing reduction bench.py uses (MAX over ranks) must agree on both."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libgooey_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, kinds, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard.shard_indices(kinds, rank, world)
    # what bench.py reduces: per-rank elapsed -> MAX; here a stand-in value proves the plumbing
    t = torch.tensor([float(len(idx)) + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(idx)], dtype=torch.int64))
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), idx=idx, tmax=t.numpy(), sizes=torch.cat(sizes).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [4096, 1001])
def test_two_ranks_partition_the_batch_with_equal_type_mix(tmp_path, n):
    kinds = [i % 4 for i in range(n)] if n == 4096 else list(np.random.default_rng(1).integers(0, 5, n))
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), kinds, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(world)]
    allidx = np.concatenate([g["idx"] for g in got])
    assert sorted(allidx.tolist()) == list(range(n))                       # a partition: nothing lost, nothing twice
    assert np.array_equal(got[0]["sizes"], got[1]["sizes"]) and got[0]["sizes"].sum() == n
    assert got[0]["tmax"] == got[1]["tmax"]
    k = np.asarray(kinds)
    for cls in set(kinds):
        per_rank = [int((k[g["idx"]] == cls).sum()) for g in got]
        assert max(per_rank) - min(per_rank) <= 2, (cls, per_rank)         # same cost mix on every rank


def test_shard_bounds_edge_cases():
    assert shard.shard_bounds(10, 0, 4) == (0, 3) and shard.shard_bounds(10, 3, 4) == (9, 10)
    assert shard.shard_bounds(2, 3, 4) == (2, 2)                           # more ranks than items: empty shard
    assert shard.shard_bounds(0, 0, 8) == (0, 0)
    assert list(shard.balanced_order([0, 0, 0, 1])) in ([0, 1, 2, 3], [0, 3, 1, 2], [0, 1, 3, 2])
