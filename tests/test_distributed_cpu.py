"""CPU: the N>1 host path with world_size 2 over gloo -- sharding, unique-id exchange, max-over-ranks timing, and the
outer-axis-shard + all-reduce composition checked against the unsharded oracle."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, tmp):
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port)})
    sys.path.insert(0, str(ROOT))
    import torch
    import oracle
    from cubecl_b200 import synth
    from cubecl_b200.distributed import exchange_unique_id, init_process_group, max_over_ranks, shard_range

    dist = init_process_group("gloo")
    uid = exchange_unique_id(lambda: bytes(range(128)), dist)
    assert uid == bytes(range(128))

    # reduce-sum: outer-axis slabs + all-reduce(sum) == unsharded sum (exact for the integer pattern)
    n = 1 << 16
    full = (np.arange(n) % 8).astype(np.float32)
    lo, hi = shard_range(n, world, rank)
    part = torch.tensor([float(oracle.sum_serial_f32(full[lo:hi]))], dtype=torch.float32)
    dist.all_reduce(part)
    assert float(part.item()) == float(oracle.sum_f64(full))

    # batched matmul: contiguous batch shards, no collective; gathering the shards reproduces the unsharded result
    B, M, N, K = 4, 8, 8, 16
    a = synth.uniform_f32(6, B * M * K, -1, 1).reshape(B, M, K)
    b = synth.uniform_f32(7, B * K * N, -1, 1).reshape(B, K, N)
    b0, b1 = shard_range(B, world, rank)
    mine = oracle.matmul_f32(a[b0:b1], b[b0:b1])
    gathered = [None] * world
    dist.all_gather_object(gathered, (b0, b1, mine))
    out = np.concatenate([g[2] for g in sorted(gathered, key=lambda t: t[0])])
    assert np.array_equal(out, oracle.matmul_f32(a, b))

    assert max_over_ranks(float(rank + 1), dist) == float(world)

    # the closed forms bench.py's multi-GPU parity object checks against (exact-integer data, runtime_tests/all_reduce.rs model)
    import bench
    n_red = 1 << 16
    assert bench.mod8_prefix_sum(n_red) == int((np.arange(n_red) % 8).sum()) and bench.mod8_prefix_sum(13) == int((np.arange(13) % 8).sum())
    lo, hi = shard_range(n_red, world, rank)
    local = bench.mod8_prefix_sum(hi) - bench.mod8_prefix_sum(lo)
    assert local == int((np.arange(lo, hi) % 8).sum())
    tot = torch.tensor([float(local)], dtype=torch.float64)
    dist.all_reduce(tot)
    assert int(tot.item()) == bench.mod8_prefix_sum(n_red)
    # one rank failing its check fails the whole record
    assert bench.all_ranks_ok(True, dist, None) is True
    assert bench.all_ranks_ok(rank != 1, dist, None) is False
    dist.barrier()
    dist.destroy_process_group()
    Path(tmp, f"ok{rank}").write_text("ok")


@pytest.mark.timeout(180)
def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
