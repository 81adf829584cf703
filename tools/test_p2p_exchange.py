#!/usr/bin/env python
"""Multi-GPU check of the in-graph peer exchange (csrc/gather.cu) against an NCCL all_gather of the same packed step result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/test_p2p_exchange.py

Every rank runs its shard of go1sheep-hard (fused wrapper: obs | reward | done) for a few hundred steps with short episodes; after every
step the peer-exchanged GLOBAL tensors must equal, bit for bit, what NCCL gathers from the ranks' local results.  Prints PASS / FAIL."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from types import SimpleNamespace
from mqe_b200 import engine as E
from mqe_b200.dist import PeerStepExchange, StepGather, shard_range, split_gathered_result
from mqe_b200.envs import make_mqe_env

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device(f"cuda:{lr}")
dist.init_process_group("nccl", device_id=dev)
task = sys.argv[1] if len(sys.argv) > 1 else "go1sheep-hard"
n_per = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
N = n_per * world
a, b = shard_range(N, rank, world)
args = SimpleNamespace(num_envs=N, seed=0, headless=True, record_video=False, sim_device=str(dev))


def cc(cfg):
    cfg.env.num_envs = N
    cfg.env.episode_length_s = 0.6
    return cfg


env, cfg = make_mqe_env(task, args, cc, env_slice=(a, b), policy_mode=E.POLICY_BF16X3)
env.reset()
eng = env.env.engine
ex = PeerStepExchange(eng)
env.reset()
g = StepGather(world)
L = eng.result_layout()
res = eng.tensor(E.BUF_STEP_RESULT)
A = env.env._ctrl_agents
rng = np.random.default_rng(0)
bad = 0
t0 = time.perf_counter()
for s in range(steps):
    act = rng.uniform(-1, 1, size=(N, A, 3)).astype(np.float32)[a:b]
    env.step(torch.as_tensor(act, device=dev))
    if s % 37 == 5:
        env.reset()                                  # resets exchange too
    G = g.gather("r", res[eng.result_parity()].view(1, -1)).view(-1)
    want = split_gathered_result(G, L, world)
    got = ex.latest()
    for w_, g_ in zip(want, got):
        if w_.shape != g_.shape or not torch.equal(w_, g_):
            bad += 1
torch.cuda.synchronize()
t = torch.tensor([bad, int(ex.timed_out())], device=dev)
dist.all_reduce(t)
if rank == 0:
    print(f"{task} world {world} x {n_per} envs, {steps} steps: mismatches {int(t[0])}, exchange timeouts {int(t[1])}, "
          f"global obs {tuple(ex.latest()[0].shape)} -> {'PASS' if int(t[0]) == 0 and int(t[1]) == 0 else 'FAIL'}  ({time.perf_counter() - t0:.1f} s)")
env.close()
dist.destroy_process_group()
