"""Checksum of a go1gate trajectory (bench actions): two builds with bit-identical kernels print identical lines.
    MQE_B200_LIB=<lib> python tools/traj_checksum.py [steps]"""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
import bench as B
from mqe_b200 import engine as E
from mqe_b200.envs.utils import make_mqe_env, custom_cfg
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = 4096
eargs = SimpleNamespace(num_envs=n, seed=0, headless=True, record_video=False, sim_device="cuda:0")
env, cfg = make_mqe_env("go1gate", eargs, custom_cfg(eargs), policy_mode=E.POLICY_BF16X3)
eng = env.env.engine
acts = torch.as_tensor(B.synth_actions(n, env.env._ctrl_agents, 64, env_offset=0), device="cuda:0")
env.reset()
for i in range(steps):
    env.step(acts[i % 64])
    if i in (0, 1, 9, 49, steps - 1):
        torch.cuda.synchronize()
        h = hashlib.sha1(eng.tensor(E.BUF_ROOT_STATES).cpu().numpy().tobytes() + eng.tensor(E.BUF_OBS).cpu().numpy().tobytes()).hexdigest()[:16]
        print(f"step {i + 1:4d} sha1(root_states | obs) = {h}", flush=True)
