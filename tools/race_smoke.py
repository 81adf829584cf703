import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from mqe_b200 import engine as E, scene as S
from mqe_b200.envs import configs as C
for task, fn, n in (("go1gate", C.Go1GateCfg, 8), ("go1sheep-hard", C.NineSheepCfg, 2), ("go1football-defender", C.Go1FootballDefenderCfg, 3), ("go1tug", C.Go1TugCfg, 4)):
    cfg = fn(); cfg.env.num_envs = n
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0, policy_mode=E.POLICY_BF16X3, wrapper_action_scale=(2.0, 0.5, 0.5))
    eng = E.Engine(sc.desc, device=0, stream=torch.cuda.current_stream().cuda_stream, keepalive=sc)
    eng.reset()
    actrl = sc.num_agents - 1 if sc.desc.defender else sc.num_agents
    act = torch.zeros((n, actrl, 3), device="cuda")
    for s in range(12):
        if s == 10:      # robots settled: put robot 1 next to robot 0 so the pair phase (publish, masks, candidates, rows, pair PGS) runs
            root = eng.tensor(E.BUF_ROOT_STATES)
            root[:, 1, :3] = root[:, 0, :3] + torch.tensor([0.30, 0.15, 0.0], device="cuda")
        eng.step(act.data_ptr())
    torch.cuda.synchronize()
    print(task, "ok", eng.tensor(E.BUF_STATS).cpu().numpy()[:4], flush=True)
    eng.close()
