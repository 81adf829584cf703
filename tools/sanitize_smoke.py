"""Small all-task run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel, few envs, few steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mqe_b200 import engine as E, scene as S
from mqe_b200.envs import configs as C

for task, fn, n in (("go1gate", C.Go1GateCfg, 8), ("go1sheep-hard", C.NineSheepCfg, 3), ("go1football-defender", C.Go1FootballDefenderCfg, 5),
                    ("go1seesaw", C.Go1SeesawCfg, 7), ("go1football-2vs2", C.Go1Football2vs2Cfg, 3), ("go1pushbox", C.Go1PushboxCfg, 5),
                    ("go1tug", C.Go1TugCfg, 6), ("go1wrestling", C.Go1WrestlingCfg, 5), ("go1bridge", C.Go1BridgeCfg, 4)):
    for mode in (E.POLICY_FP32, E.POLICY_BF16X3):
        cfg = fn(); cfg.env.num_envs = n; cfg.env.episode_length_s = 0.1
        np.random.seed(0)
        sc = S.build_scene(cfg, seed=0, policy_mode=mode, wrapper_action_scale=(2.0, 0.5, 0.5))
        eng = E.Engine(sc.desc, device=0, stream=torch.cuda.current_stream().cuda_stream, keepalive=sc)
        eng.reset()
        actrl = sc.num_agents - 1 if sc.desc.defender else sc.num_agents
        # put robots close together so the pair phase runs
        root = eng.tensor(E.BUF_ROOT_STATES)
        root[:, 1, :2] = root[:, 0, :2] + 0.35
        for s in range(8):
            act = torch.rand((n, actrl, 3), device="cuda") * 2 - 1
            eng.step(act.data_ptr())
        torch.cuda.synchronize()
        print(task, mode, "ok", eng.tensor(E.BUF_STATS).cpu().numpy()[:4], flush=True)
        eng.close()
