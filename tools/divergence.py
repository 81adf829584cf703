"""SURVEY 8(c) tier 2: GPU kernels vs the CPU oracle (same algorithm, fp32) on identical seeds and action sequences over a whole episode:
fraction of envs whose root position / orientation / joint state still agrees within the stated tolerances, per policy step.
Contact dynamics are chaotic (a foot that touches down one substep earlier changes everything after), so agreement is expected to
hold tightly for tens of steps and then peel off env by env; the report is that curve.  Usage: python tools/divergence.py [task] [envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from mqe_b200 import engine as E, scene as S
from mqe_b200.envs import configs as C

task = sys.argv[1] if len(sys.argv) > 1 else "go1gate"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 500
cfg = {"go1gate": C.Go1GateCfg, "go1sheep-hard": C.NineSheepCfg, "go1seesaw": C.Go1SeesawCfg, "go1football-defender": C.Go1FootballDefenderCfg}[task]()
cfg.env.num_envs = n
np.random.seed(0)
sc = S.build_scene(cfg, seed=0, policy_mode=E.POLICY_FP32, wrapper_action_scale=(2.0, 0.5, 0.5))
eng, orc = E.Engine(sc.desc, device=0, keepalive=sc), oracle.Oracle(sc, "f32")
eng.reset(); orc.reset()
A = sc.num_agents
a_ctrl = A - 1 if sc.desc.defender else A
TOL_POS, TOL_VEL = 1e-4, 1e-3                         # abs tol pos / quat 1e-4, velocities 1e-3 for the first 50 policy steps
print(f"{task}: {n} envs x {A} agents, {steps} policy steps, fp32 CUDA-core policy, tolerances pos/quat {TOL_POS}, vel {TOL_VEL}")
print("step  within_tol  within_1cm  max|dpos|   median|dpos|  resets_equal")
alive = np.ones(n, dtype=bool)                       # envs that have not yet left the tolerance band
for s in range(steps):
    rng = np.random.default_rng(1000 + s)
    act = rng.uniform(-1, 1, size=(n, a_ctrl, 3)).astype(np.float32)
    eng.step(torch.as_tensor(act, device="cuda:0").data_ptr()); orc.step(act)
    g = eng.tensor(E.BUF_ROOT_STATES).cpu().numpy().reshape(n, -1, 13)[:, :A]
    r = orc.root_states()[:, :A]
    dp = np.abs(g[..., :7] - r[..., :7]).max(axis=(1, 2))
    dv = np.abs(g[..., 7:] - r[..., 7:]).max(axis=(1, 2))
    alive &= (dp < TOL_POS) & (dv < TOL_VEL)
    same_reset = np.array_equal(eng.tensor(E.BUF_RESET).cpu().numpy(), orc.get(E.BUF_RESET))
    if s < 10 or s % 25 == 24:
        print(f"{s + 1:4d}  {alive.mean():9.3f}  {(dp < 1e-2).mean():9.3f}  {dp.max():9.2e}  {np.median(dp):11.2e}  {same_reset}")
eng.close(); orc.close()
