"""Sharding of the environment batch across ranks: one process per GPU, contiguous env blocks,
no exchange inside the physics step; one collective per step gathers what the learner reads.

`shard_range` / `StepGather` are pure torch.distributed host logic and run on gloo (CPU tests) and nccl alike.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(num_envs_global: int, rank: int, world: int):
    """env e -> rank floor(e * world / N): contiguous blocks whose sizes differ by at most one."""
    start = (rank * num_envs_global + world - 1) // world
    stop = ((rank + 1) * num_envs_global + world - 1) // world
    return start, stop


def owner_of(env_id: int, num_envs_global: int, world: int) -> int:
    return (env_id * world) // num_envs_global


class StepGather:
    """All-gather of per-step results (observation rows, rewards, done flags) into rank-ordered global tensors.

    Equal shards use one `all_gather_into_tensor`; ragged shards are padded to the largest shard first.
    Buffers are allocated once and reused every step."""

    def __init__(self, num_envs_global: int, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_global = num_envs_global
        self.ranges = [shard_range(num_envs_global, r, self.world) for r in range(self.world)]
        self.equal = len({b - a for a, b in self.ranges}) == 1
        self._out = {}

    def local_range(self):
        return self.ranges[self.rank]

    def gather(self, name: str, local: torch.Tensor) -> torch.Tensor:
        """local: [n_local, ...] -> [n_global, ...] in global env order."""
        if self.world == 1:
            return local
        local = local.contiguous()
        key = (name, local.dtype, tuple(local.shape[1:]))
        out = self._out.get(key)
        if out is None:
            out = torch.empty((self.n_global,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            self._out[key] = out
        as_u8 = local.dtype == torch.bool                    # NCCL has no bool
        if self.equal:
            src, dst = (local.view(torch.uint8), out.view(torch.uint8)) if as_u8 else (local, out)
            dist.all_gather_into_tensor(dst, src, group=self.group)
        else:                                                # ragged: pad every shard to the largest, gather, unpack
            big = max(b - a for a, b in self.ranges)
            pkey = key + ("pad",)
            bufs = self._out.get(pkey)
            if bufs is None:
                bufs = (torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device),
                        torch.empty((self.world * big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device))
                self._out[pkey] = bufs
            pad, allbuf = bufs
            pad[:local.shape[0]].copy_(local)
            src, dst = (pad.view(torch.uint8), allbuf.view(torch.uint8)) if as_u8 else (pad, allbuf)
            dist.all_gather_into_tensor(dst, src, group=self.group)
            for r, (a, b) in enumerate(self.ranges):
                out[a:b].copy_(allbuf[r * big:r * big + (b - a)])
        return out


class PeerStepExchange:
    """The per-step exchange as part of the engine's step graph (csrc/gather.cu): every rank stores its packed step result (task-wrapper
    observation | reward | done) straight into every peer's receive buffer over NVLink peer memory; no NCCL call, no extra launch from
    the host.  `torch.distributed` is used once, to hand the 64-byte IPC handles around.  Equal shards with num_envs % 16 == 0.

        ex = PeerStepExchange(env.engine)          # after the task wrapper has been set up (first env.reset())
        env.step(actions)                          # exchange happens inside the step
        obs, reward, done = ex.latest()            # GLOBAL tensors [N_global, A, D], [N_global, A], [N_global] (zero-copy views)
    """

    def __init__(self, engine, group=None):
        self.engine = engine
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        handle = engine.gather_init(self.rank, self.world)
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, handle, group=group)
        else:
            handles[0] = handle
        engine.gather_connect(handles)
        from . import engine as E
        self._views = []
        for parity in (0, 1):
            buf, L = engine.gather_view(parity)
            self._views.append(E.Engine.split_result(buf, L))
        self.layout = L

    def latest(self):
        return self._views[self.engine.gather_parity()]

    def timed_out(self) -> bool:
        """True if a peer's flag ever failed to arrive within the kernel's bounded wait (sticky; MQE_BUF_STATS[5])."""
        from . import engine as E
        return int(self.engine.tensor(E.BUF_STATS)[E.STAT_GATHER_TIMEOUT].item()) != 0


def split_gathered_result(buf: torch.Tensor, layout, world: int):
    """NCCL / gloo fallback of the peer exchange: `buf` is the all-gathered packed step result, [world x total_bytes] u8 (every rank's
    MQE_BUF_STEP_RESULT half back to back); returns GLOBAL (obs [N_global, Aw, D], reward [N_global, Aw], done [N_global])."""
    from .engine import Engine
    total = int(layout.total_bytes)
    parts = [Engine.split_result(buf[r * total:(r + 1) * total], layout) for r in range(world)]
    return tuple(torch.cat([p[i] for p in parts], dim=0) for i in range(3))
