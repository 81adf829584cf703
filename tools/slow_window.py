"""The slowest k_substeps warp, step by step, on bench.py's own trajectory (same seeds, same pre-roll): which env sets the launch time,
how many pair contacts / rows it has and where its cycles go.  Run with MQE_TRACE=1 for the per-phase cycles.
    MQE_TRACE=1 python tools/slow_window.py [first_step_after_preroll] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
import bench as B
from mqe_b200 import engine as E
from mqe_b200.envs.utils import make_mqe_env, custom_cfg

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = 4096
eargs = SimpleNamespace(num_envs=n, seed=0, headless=True, record_video=False, sim_device="cuda:0")
env, cfg = make_mqe_env("go1gate", eargs, custom_cfg(eargs), policy_mode=E.POLICY_BF16X3)
base, eng = env.env, env.env.engine
acts = torch.as_tensor(B.synth_actions(n, base._ctrl_agents, 64, env_offset=0), device="cuda:0")
env.reset()
B.desynchronise(base, torch, seed=0)
pre = 3 * int(base.max_episode_length) + 1 + first
for i in range(pre):
    env.step(acts[i % 64])
names_p = ["P1", "P2", "P3rows", "P3b", "PGS", "P5", "pro", "epi", "lim", "probes", "setup", "bp+pub", "narrow", "prow", "psweep", "post"]
root = eng.tensor(E.BUF_ROOT_STATES)
for i in range(steps):
    env.step(acts[(pre + i) % 64])
    torch.cuda.synchronize()
    tr = eng.tensor(E.BUF_WARP_TRACE).cpu().numpy().astype(np.int64)
    d = (tr[:, 1] - tr[:, 0]) * 1e-3
    order = np.argsort(-d)[:3]
    k = int(order[0])
    span = (tr[:, 1].max() - tr[:, 0].min()) * 1e-3
    ph = tr[k, 4:20]
    phs = " ".join(f"{nm}={int(v / 1000)}k" for nm, v in zip(names_p, ph) if v > 0.04 * max(1, ph.sum()))
    r = root.view(n, -1, 13)[4 * k:4 * k + 4].cpu().numpy()
    dist = np.linalg.norm(r[:, 0, :2] - r[:, 1, :2], axis=1)
    print(f"step {pre + i:5d} span {span:6.1f} mean {d.mean():6.1f} | slowest warp {k:4d} {d[k]:6.1f} us pairs {tr[k, 2]:3d} rows {tr[k, 3]:2d} | next {d[order[1]]:6.1f} {d[order[2]]:6.1f} | robot distance in its envs {np.round(dist, 2)} z {np.round(r[:, :, 2].min(), 2)} | {phs}", flush=True)
