"""Profiling workload: EVERY env holds two robots standing entangled (0.3 m apart), so the pair path of k_substeps (broad phase, narrow
phase, pair rows, pair sweeps) dominates the launch and shows up in an ncu source view.  Not a benchmark.
    ncu --set full --import-source on -k regex:k_substeps -s 40 -c 1 -o gpurun_out/pileup python tools/pileup_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mqe_b200 import engine as E, scene as S
from mqe_b200.envs import configs as C
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = C.Go1GateCfg(); cfg.env.num_envs = n
np.random.seed(0)
sc = S.build_scene(cfg, seed=0, policy_mode=E.POLICY_BF16X3, wrapper_action_scale=(2.0, 0.5, 0.5))
eng = E.Engine(sc.desc, device=0, stream=torch.cuda.current_stream().cuda_stream, keepalive=sc)
eng.reset()
act = torch.zeros((n, 2, 3), device="cuda")
root = eng.tensor(E.BUF_ROOT_STATES)
dx, dy = float(os.environ.get("PILE_DX", "0.10")), float(os.environ.get("PILE_DY", "0.28"))
hold = int(os.environ.get("PILE_HOLD", "6"))
for s in range(20 + hold):
    if s == 20:
        root.view(n, -1, 13)[:, 1, :3] = root.view(n, -1, 13)[:, 0, :3] + torch.tensor([dx, dy, 0.0], device="cuda")
    eng.step(act.data_ptr())
    if s >= 20 and os.environ.get("PILE_VERBOSE"):
        torch.cuda.synchronize()
        st = eng.tensor(E.BUF_STATS).cpu().numpy()
        print(s, "pair contacts per env per substep:", st[2] / n / 4, " local:", st[0] / n / 4)
torch.cuda.synchronize()
st = eng.tensor(E.BUF_STATS).cpu().numpy()
print("pair contacts per env per substep in the last step:", st[2] / n / 4, " local contacts:", st[0] / n / 4)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(20):
    eng.substeps(4)
e1.record(); torch.cuda.synchronize()
print("k_substeps with every env piled up: %.3f ms" % (e0.elapsed_time(e1) / 20))
