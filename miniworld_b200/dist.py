"""Multi-GPU sharding of the batched step path.

Environments are independent, so the global env index range is cut into contiguous slices,
one per rank (one process per GPU, `torch.distributed`); the step itself needs no exchange.
What the path does exchange, once per step: actions are scattered from rank 0 and the uint8
observations (+ reward / flags) are gathered to rank 0 -- NCCL on GPUs, gloo in the CPU tests.
Global env i is always seeded `seed + i`, whatever the world size, so a sharded run reproduces
the single-process run env for env.
"""
import numpy as np


def shard_range(total, world_size, rank):
    """Contiguous slice [start, start + count) of `total` envs owned by `rank`."""
    base, extra = divmod(int(total), int(world_size))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


class ShardedMiniWorld:
    """This rank's slice of a `total_envs`-wide BatchedMiniWorld plus the per-step exchange."""

    def __init__(self, level, total_envs, dist=None, device=0, **kwargs):
        from .batched import BatchedMiniWorld
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.total = int(total_envs)
        self.start, self.count = shard_range(self.total, self.world, self.rank)
        self.counts = [shard_range(self.total, self.world, r)[1] for r in range(self.world)]
        self.local = BatchedMiniWorld(level, self.count, device=device, **kwargs)

    def reset(self, seed):
        """Global env i gets `seed + i`."""
        return self.local.reset(seed=[int(seed) + self.start + k for k in range(self.count)])

    # ---- exchange helpers (torch tensors on any device; uneven shards supported)
    def scatter_actions(self, actions_all, like):
        """rank 0 holds int32 [total]; every rank receives its [count] slice."""
        import torch
        if self.dist is None or self.world == 1:
            return actions_all
        mine = torch.empty(self.count, dtype=torch.int32, device=like.device)
        if self.rank == 0:
            offs = np.cumsum([0] + self.counts)
            chunks = [actions_all[offs[r]:offs[r + 1]].contiguous() for r in range(self.world)]
            mine.copy_(chunks[0])
            reqs = [self.dist.isend(chunks[r], dst=r) for r in range(1, self.world)]
            for q in reqs:
                q.wait()
        else:
            self.dist.recv(mine, src=0)
        return mine

    def gather_to_root(self, tensor):
        """Concatenate every rank's leading-dim slice on rank 0 (None elsewhere)."""
        import torch
        if self.dist is None or self.world == 1:
            return tensor
        if self.rank == 0:
            parts = [tensor] + [torch.empty((self.counts[r],) + tuple(tensor.shape[1:]), dtype=tensor.dtype,
                                            device=tensor.device) for r in range(1, self.world)]
            reqs = [self.dist.irecv(parts[r], src=r) for r in range(1, self.world)]
            for q in reqs:
                q.wait()
            return torch.cat(parts, dim=0)
        self.dist.send(tensor.contiguous(), dst=0)
        return None

    # ---- peer-memory observations: K2 of every rank stores straight into rank 0's buffer
    def enable_peer_obs(self):
        """Rank 0 allocates uint8 [total, H, W, 3] and shares it over CUDA IPC; the other ranks map
        it and render into their slice of it (stores cross NVLink inside K2).  Returns True if
        every rank succeeded; otherwise the NCCL-gather path stays in use."""
        import torch
        from .engine import EngineError, SharedDeviceBuffer
        H, W = self.local.obs_height, self.local.obs_width
        shape = (self.total, H, W, 3)
        dev = self.local.device
        ok, self._peer = 1, None
        payload = [None]
        try:
            if self.rank == 0:
                self._peer = SharedDeviceBuffer(dev, shape)
                payload = [self._peer.handle]
        except EngineError:
            ok = 0
        if self.dist is not None and self.world > 1:
            self.dist.broadcast_object_list(payload, src=0)
            if self.rank != 0 and payload[0] is not None:
                try:
                    self._peer = SharedDeviceBuffer(dev, shape, handle=payload[0])
                except EngineError:
                    ok = 0
            flag = torch.tensor([ok if payload[0] is not None else 0], device=torch.device("cuda", dev))
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
            ok = int(flag.item())
        if not ok:
            if self._peer is not None:
                self._peer.close()
            self._peer = None
            return False
        self.obs_all = self._peer.tensor()                         # [total, H, W, 3] in rank 0's HBM
        self.local._bufs_ready = self.local._ensure_torch()
        self.local._bufs["obs"] = self.obs_all[self.start:self.start + self.count]   # this rank's slice
        self._sync = torch.zeros(1, device=torch.device("cuda", dev))
        return True

    def step_peer(self, local_actions):
        """K1 + K2 with observations written into rank 0's buffer; one tiny stream-ordered
        all-reduce tells rank 0 that every slice is complete.  Returns obs_all on rank 0."""
        obs, rew, te, tr, info = self.local.step(local_actions)
        if self.dist is not None and self.world > 1:
            self.dist.all_reduce(self._sync)
        return self.obs_all if self.rank == 0 else None

    def step(self, local_actions):
        """Local K1 + K2, then the gather of obs / reward / flags to rank 0."""
        obs, rew, te, tr, info = self.local.step(local_actions)
        return (self.gather_to_root(obs), self.gather_to_root(rew), self.gather_to_root(te.to(obs.dtype)),
                self.gather_to_root(tr.to(obs.dtype)))

    def close(self):
        if getattr(self, "_peer", None) is not None:
            self._peer.close()
        self.local.close()
