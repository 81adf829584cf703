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


def _pack_worker(rank, world, port, q):
    """The NCCL / gloo fallback of the per-step exchange: ONE all_gather of the packed step result (obs | reward | done), then split."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mqe_b200 import engine as E
        from mqe_b200.dist import split_gathered_result
        n, Aw, D = 16, 2, 5
        L = E.StepResultLayoutC()
        up16 = lambda v: (v + 15) & ~15
        L.num_envs, L.Aw, L.D = n, Aw, D
        L.obs_off, L.obs_bytes = 0, n * Aw * D * 4
        L.reward_off, L.reward_bytes = up16(L.obs_bytes), n * Aw * 4
        L.done_off, L.done_bytes = up16(L.reward_off + L.reward_bytes), n
        L.total_bytes = up16(L.done_off + L.done_bytes)
        local = torch.zeros(int(L.total_bytes), dtype=torch.uint8)
        obs, rew, done = E.Engine.split_result(local, L)
        obs.copy_(torch.arange(n * Aw * D, dtype=torch.float32).view(n, Aw, D) + 1000 * rank)
        rew.copy_(torch.full((n, Aw), float(rank)))
        done.copy_(torch.arange(n) % (rank + 2) == 0)
        g = StepGather(world)                                # one packed row per rank
        G = g.gather("result", local.view(1, -1)).view(-1)
        go, gr, gd = split_gathered_result(G, L, world)
        ok = go.shape == (n * world, Aw, D) and gr.shape == (n * world, Aw) and gd.dtype == torch.bool
        for r in range(world):
            ok &= bool(torch.equal(go[r * n:(r + 1) * n], torch.arange(n * Aw * D, dtype=torch.float32).view(n, Aw, D) + 1000 * r))
            ok &= bool((gr[r * n:(r + 1) * n] == r).all()) and bool(torch.equal(gd[r * n:(r + 1) * n], torch.arange(n) % (r + 2) == 0))
        q.put((rank, ok))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, f"{type(exc).__name__}: {exc}"))
    finally:
        dist.destroy_process_group()


def test_packed_step_result_gather_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pack_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok is True for _, ok in res), res
