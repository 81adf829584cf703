"""Debug: distribution of contacts per robot in steady state, and the per-warp (4 consecutive envs) maximum."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mqe_b200 import engine as E, scene as S
from mqe_b200.envs import configs as C
cfg = C.Go1GateCfg(); cfg.env.num_envs = 4096
np.random.seed(0)
sc = S.build_scene(cfg, seed=0, policy_mode=1, wrapper_action_scale=(2.0, 0.5, 0.5))
eng = E.Engine(sc.desc, device=0, stream=torch.cuda.current_stream().cuda_stream, keepalive=sc)
eng.reset()
for step in range(400):
    act = (torch.rand((4096, 2, 3), device="cuda") * 2 - 1)
    eng.step(act.data_ptr())
    if step in (5, 50, 150, 399):
        torch.cuda.synchronize()
        cf = eng.tensor(E.BUF_CONTACT_FORCES).cpu().numpy().reshape(4096, 2, 17, 3)
        nb = (np.abs(cf).sum(-1) > 0).sum(-1)          # bodies in contact per robot (lower bound on contacts)
        per_env = nb.max(1)
        per_warp = per_env.reshape(-1, 4).max(1)
        root = eng.tensor(E.BUF_ROOT_STATES).cpu().numpy()
        d = np.linalg.norm(root[:, 0, :3] - root[:, 1, :3], axis=1)
        print(step, "bodies-in-contact per robot: mean %.2f  hist" % nb.mean(), np.bincount(nb.ravel(), minlength=10)[:10],
              " per-warp max mean %.2f" % per_warp.mean(), " robots within 1.2 m: %.2f  within 0.7: %.2f" % ((d < 1.21).mean(), (d < 0.7).mean()),
              "ep_len mean", eng.tensor(E.BUF_EPISODE_LENGTH).float().mean().item())
