"""N>1 host logic on CPU: world_size-2 gloo processes shard the episodes, time a step and
reduce -- the same code path bench.py runs under torchrun with NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jolineedle_b200.sharding import max_over_ranks, reduce_eval_metrics, shard_bounds, sum_over_ranks


def test_shard_bounds_partition_every_episode_once():
    for n in (0, 1, 7, 256, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(1001, rank, world)
        # every rank "processes" its shard; rank 1 is slower
        ms = 10.0 + 5.0 * rank
        slowest = max_over_ranks(ms, "cpu")
        units = sum_over_ranks(float(hi - lo), "cpu")
        metrics = reduce_eval_metrics({"prop_patches_found": 0.5 * (hi - lo), "stop_used": float(rank)}, hi - lo, "cpu")
        results[rank] = (slowest, units, metrics["prop_patches_found"], metrics["stop_used"])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()  # (fork from a multi-threaded pytest process can deadlock)
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert len(results) == 2
    for rank in range(world):
        slowest, units, prop, stop_used = results[rank]
        assert slowest == 15.0 and units == 1001.0
        assert abs(prop - 0.5) < 1e-12 and abs(stop_used - 1.0 / 1001.0) < 1e-12


def test_single_process_is_a_no_op():
    assert max_over_ranks(3.5, "cpu") == 3.5 and sum_over_ranks(2.0, "cpu") == 2.0
    assert reduce_eval_metrics({"a": 6.0}, 3, "cpu") == {"a": 2.0}
