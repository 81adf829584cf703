"""Debug: cold-start cost of k_substeps: time nsub = 1, 2, 4, 8 with and without an L2 flush before the launch."""
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
act = (torch.rand((4096, 2, 3), device="cuda") * 2 - 1)
for i in range(100):
    eng.step(act.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").view(torch.int32)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for mode in ("warm", "cold"):
    for nsub in (1, 2, 4, 8):
        ts = []
        for r in range(12):
            eng.policy(act.data_ptr())
            if mode == "cold":
                _ = flush.sum()
            e0.record(); eng.substeps(nsub); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            eng.post_physics()
        print(mode, "nsub", nsub, "ms %.4f" % np.median(ts))
