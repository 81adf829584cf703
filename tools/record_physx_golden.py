#!/usr/bin/env python
"""Record PhysX golden trajectories from the UNMODIFIED reference (SURVEY 8(c) tier 4).  NOT runnable in this repository's container:
it needs NVIDIA Isaac Gym Preview 4 (closed binary, Python <= 3.8) and the reference checkout on PYTHONPATH.

    # on a machine with Isaac Gym:   (reference = ziyanx02/multiagent-quadruped-environment @ 606208bc)
    cd <reference checkout> && python <this repo>/tools/record_physx_golden.py --task go1gate --num_envs 64 --steps 50 \
        --sim_device cpu --out <this repo>/tests/golden/physx_go1gate.npz

It drives `make_mqe_env(task, args)` exactly as `openrl_ws/train.py` does, with the action stream of `bench.synth_actions`
(Philox keyed by (seed, step), U(-1, 1), shape [N, A_ctrl, 3] -- restated below so the tool has no dependency on this repo), and
stores after reset() and after every env.step(): `all_root_states` [N, A+P, 13], `all_dof_states` [N, 12A+D, 2], `reset_buf` [N],
plus the actions.  tests/test_gpu_baseline_parity.py::test_physx_golden_trajectories consumes the file: it starts the CUDA engine from
the recorded post-reset state (the reference's reset draws from torch's RNG stream, ours from a counter RNG) and compares the next
25 policy steps with the tolerances stated there.  Until such a file exists DESIGN.md says "PhysX parity unpinned".
"""
import argparse
import sys

import numpy as np


def synth_actions(n_envs, a_ctrl, steps, seed=0):
    out = np.empty((steps, n_envs, a_ctrl, 3), dtype=np.float32)
    for s in range(steps):
        rng = np.random.Generator(np.random.Philox(key=seed * 1_000_003 + s))
        out[s] = rng.uniform(-1, 1, size=(n_envs, a_ctrl, 3)).astype(np.float32)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="go1gate")
    ap.add_argument("--num_envs", type=int, default=64)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--sim_device", default="cpu", help="'cpu' = the PhysX CPU pipeline BASELINE.json config 1 names")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    try:
        import isaacgym  # noqa: F401  (must be imported before torch)
    except ImportError:
        sys.exit("Isaac Gym Preview 4 is not installed: this recorder runs on the reference's own stack only")
    import torch
    from mqe.envs.utils import custom_cfg, make_mqe_env
    from mqe.utils import get_args

    sys.argv = [sys.argv[0], "--task", a.task, "--num_envs", str(a.num_envs), "--seed", str(a.seed), "--headless", "--sim_device", a.sim_device]
    args = get_args()
    args.record_video = False
    env, cfg = make_mqe_env(a.task, args, custom_cfg(args))
    base = env
    while hasattr(base, "env"):
        base = base.env
    A = base.num_agents
    a_ctrl = A - 1 if type(base).__name__ == "Go1FootballDefender" else A
    acts = synth_actions(a.num_envs, a_ctrl, a.steps, a.seed)
    dev = base.device
    env.reset()
    grab = lambda t: t.detach().cpu().numpy().copy()
    root, dof, reset = [grab(base.all_root_states).reshape(a.num_envs, -1, 13)], [grab(base.all_dof_states).reshape(a.num_envs, -1, 2)], []
    for s in range(a.steps):
        env.step(torch.as_tensor(acts[s], device=dev))
        root.append(grab(base.all_root_states).reshape(a.num_envs, -1, 13))
        dof.append(grab(base.all_dof_states).reshape(a.num_envs, -1, 2))
        reset.append(grab(base.reset_buf).astype(bool))
    np.savez_compressed(a.out, task=a.task, num_envs=a.num_envs, seed=a.seed, actions=acts, root=np.stack(root), dof=np.stack(dof),
                        reset=np.stack(reset), sim_device=a.sim_device, dt=float(cfg.sim.dt), decimation=int(cfg.control.decimation))
    print("wrote", a.out, "root", np.stack(root).shape)


if __name__ == "__main__":
    main()
