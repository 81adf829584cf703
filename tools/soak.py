"""Soak: thousands of steps with random actions on every task; reports non-finite states, reset rates, speed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mqe_b200 import engine as E, scene as S
from mqe_b200.envs import configs as C
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for task, fn, n in (("go1gate", C.Go1GateCfg, 4096), ("go1sheep-hard", C.NineSheepCfg, 2048), ("go1football-defender", C.Go1FootballDefenderCfg, 2048),
                    ("go1seesaw", C.Go1SeesawCfg, 2048), ("go1football-2vs2", C.Go1Football2vs2Cfg, 1024), ("go1pushbox", C.Go1PushboxCfg, 1024),
                    ("go1revolvingdoor", C.Go1RotationCfg, 1024), ("go1tug", C.Go1TugCfg, 1024), ("go1wrestling", C.Go1WrestlingCfg, 1024),
                    ("go1bridge", C.Go1BridgeCfg, 1024)):
    cfg = fn(); cfg.env.num_envs = n
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0, policy_mode=E.POLICY_BF16X3, wrapper_action_scale=(2.0, 0.5, 0.5))
    eng = E.Engine(sc.desc, device=0, stream=torch.cuda.current_stream().cuda_stream, keepalive=sc)
    eng.reset()
    actrl = sc.num_agents - 1 if sc.desc.defender else sc.num_agents
    root, dof, rst = eng.tensor(E.BUF_ROOT_STATES), eng.tensor(E.BUF_DOF_STATES), eng.tensor(E.BUF_RESET)
    bad = torch.zeros((), device="cuda"); resets = torch.zeros((), device="cuda"); vmax = torch.zeros((), device="cuda")
    act = torch.zeros((n, actrl, 3), device="cuda")
    t0 = time.time()
    for s in range(steps):
        if s % 25 == 0:
            act = torch.rand((n, actrl, 3), device="cuda") * 2 - 1        # commands held for 0.5 s: robots actually travel
        eng.step(act.data_ptr())
        bad += (~torch.isfinite(root)).any(dim=(1, 2)).sum() + (~torch.isfinite(dof)).any(dim=(1, 2)).sum()
        resets += rst.sum()
        vmax = torch.maximum(vmax, root[..., 7:13].abs().max())
    torch.cuda.synchronize()
    print(f"{task:22s} N={n} steps={steps}: non-finite env-steps {int(bad)}  resets/env/1000 steps {float(resets) / n / steps * 1000:.2f}  max |vel| {float(vmax):.1f}  {n * sc.num_agents * steps / (time.time() - t0) / 1e6:.2f} M agent-steps/s (incl. checks)", flush=True)
    eng.close()
