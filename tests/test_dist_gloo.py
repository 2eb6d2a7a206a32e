"""N > 1 path on CPU: two gloo ranks shard 12 envs, step them with the kernels' host build and
gather to rank 0; the result must equal the single-process run env for env."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from miniworld_b200.dist import shard_range
    for total in (1, 7, 8, 4096, 8191):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1


def _worker(rank, world, port, hostsim, total, steps, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from miniworld_b200 import engine
    from miniworld_b200.dist import ShardedMiniWorld
    engine._override_library_for_tests(hostsim)          # spawned worker: select the host build of the kernels here too
    env = ShardedMiniWorld("MiniWorld-FourRooms-v0", total, dist=dist)
    env.local.engine.seed(np.arange(env.count), np.array(
        [__import__("miniworld_b200.engine", fromlist=["x"]).rng_state_of(1000 + env.start + k) for k in range(env.count)]))
    env.local.engine.reset()
    acts_all = torch.as_tensor(np.random.default_rng(5).integers(0, 3, size=(steps, total), dtype=np.int32))
    outs = []
    out = None
    for t in range(steps):
        mine = env.scatter_actions(acts_all[t] if rank == 0 else None, like=torch.zeros(1))
        out = env.local.step_host(mine.numpy(), out)
        obs = env.gather_to_root(torch.as_tensor(out["obs"]))
        rew = env.gather_to_root(torch.as_tensor(out["reward"]))
        if rank == 0:
            outs.append((obs.numpy().copy(), rew.numpy().copy()))
    if rank == 0:
        q.put(outs)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_single_process(hostsim_path):
    import torch.multiprocessing as mp
    total, steps = 12, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, hostsim_path, total, steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process reference run
    from miniworld_b200.batched import BatchedMiniWorld
    from miniworld_b200.engine import rng_state_of
    env = BatchedMiniWorld("MiniWorld-FourRooms-v0", total)
    env.engine.seed(np.arange(total), np.array([rng_state_of(1000 + k) for k in range(total)]))
    env.engine.reset()
    acts_all = np.random.default_rng(5).integers(0, 3, size=(steps, total), dtype=np.int32)
    out = None
    for t in range(steps):
        out = env.step_host(acts_all[t], out)
        assert np.array_equal(out["obs"], sharded[t][0])
        assert np.array_equal(out["reward"], sharded[t][1])
    assert 0 < out["obs"].mean() < 255
