"""Micro-timing of the step phases with CUDA events, with and without the L2 flush (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mqe_b200 import engine as E, scene as S
from mqe_b200.envs import configs as C

task = sys.argv[1] if len(sys.argv) > 1 else "go1gate"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg = {"go1gate": C.Go1GateCfg, "go1sheep-hard": C.NineSheepCfg, "go1football-defender": C.Go1FootballDefenderCfg, "go1seesaw": C.Go1SeesawCfg}[task]()
cfg.env.num_envs = n
np.random.seed(0)
sc = S.build_scene(cfg, seed=0, policy_mode=int(os.environ.get("MQE_POLICY_MODE", "1")), wrapper_action_scale=(2.0, 0.5, 0.5))
eng = E.Engine(sc.desc, device=0, stream=torch.cuda.current_stream().cuda_stream, keepalive=sc)
eng.reset()
actrl = sc.num_agents - 1 if sc.desc.defender else sc.num_agents
act = (torch.rand((n, actrl, 3), device="cuda") * 2 - 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(30):
    eng.step(act.data_ptr())
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
flush_i32 = flush.view(torch.int32)
for do_flush in (False, True, "read"):
    acc = np.zeros(3)
    R = 30
    for i in range(R):
        if do_flush == "read":
            _ = flush_i32.sum()
        elif do_flush:
            flush.zero_()
        ev[0].record(); eng.policy(act.data_ptr()); ev[1].record(); eng.substeps(4); ev[2].record(); eng.post_physics(); ev[3].record()
        torch.cuda.synchronize()
        acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
    print(task, n, {False: "noflush", True: "flush-write", "read": "flush-read"}[do_flush], "policy/substeps/post ms:", np.round(acc / R, 4), "stats", eng.tensor(E.BUF_STATS).cpu().numpy()[:4])
# whole steps back to back
ev[0].record()
for i in range(50):
    eng.step(act.data_ptr())
ev[1].record(); torch.cuda.synchronize()
print("back-to-back step ms:", ev[0].elapsed_time(ev[1]) / 50)
