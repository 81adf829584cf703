"""Multi-rank host logic on CPU: gloo, world_size 2 (and a ragged 3-way split)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mqe_b200.dist import StepGather, owner_of, shard_range


def test_shard_ranges_cover_and_balance():
    for n in (1, 4, 7, 4096, 32768, 1001):
        for w in (1, 2, 3, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
            for e in (0, n // 2, n - 1):
                r = owner_of(e, n, w)
                assert rs[r][0] <= e < rs[r][1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_global, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = StepGather(n_global)
        a, b = g.local_range()
        obs = (torch.arange(a, b, dtype=torch.float32).view(-1, 1, 1) * torch.ones(1, 2, 5)) + 0.25 * rank
        done = (torch.arange(a, b) % 3 == 0)
        rew = torch.arange(a, b, dtype=torch.float32).view(-1, 1).repeat(1, 2)
        for _ in range(2):                                  # second round reuses the cached output buffers
            G_obs, G_done, G_rew = g.gather("obs", obs), g.gather("done", done), g.gather("rew", rew)
        ok = G_obs.shape == (n_global, 2, 5) and G_done.dtype == torch.bool
        ok &= bool(torch.equal(G_done, torch.arange(n_global) % 3 == 0))
        ok &= bool(torch.equal(G_rew[:, 0], torch.arange(n_global, dtype=torch.float32)))
        for r in range(world):
            ra, rb = shard_range(n_global, r, world)
            ok &= bool(torch.allclose(G_obs[ra:rb, 0, 0], torch.arange(ra, rb, dtype=torch.float32) + 0.25 * r))
        q.put((rank, ok))
    except Exception as exc:  # noqa: BLE001 - report instead of timing the parent out
        q.put((rank, f"{type(exc).__name__}: {exc}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_global", [(2, 64), (3, 10)])
def test_step_gather_gloo(world, n_global):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_global, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok is True for _, ok in res), res
