"""In-step stage times (steady state, L2 flushed between steps like bench.py): where the step's milliseconds go INSIDE the graph.
    python tools/stage_times.py [task] [num_envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from mqe_b200 import engine as E
from mqe_b200.envs import custom_cfg, make_mqe_env
import bench

task = sys.argv[1] if len(sys.argv) > 1 else "go1gate"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = int(sys.argv[3]) if len(sys.argv) > 3 else 200
args = SimpleNamespace(num_envs=n, seed=0, headless=True, record_video=False, sim_device="cuda:0")
env, cfg = make_mqe_env(task, args, custom_cfg(args), policy_mode=E.POLICY_BF16X3)
base, eng = env.env, env.env.engine
env.reset()
acts = torch.as_tensor(bench.synth_actions(n, base._ctrl_agents, 64), device="cuda")
bench.desynchronise(base, torch, seed=0)
for i in range(int(base.max_episode_length) + 1):
    env.step(acts[i % 64])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for timing in (False, True):
    eng.stage_timing(timing)
    for i in range(5):
        env.step(acts[i % 64])
    rows, tot = [], []
    for i in range(K):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(acts[i % 64]); e1.record()
        if timing:
            rows.append(list(eng.stage_ms().values()))
        e1.synchronize()
        tot.append(e0.elapsed_time(e1))
    print(f"{task} N={n} stage marks {'on ' if timing else 'off'}: step {np.mean(tot):.4f} ms (p50 {np.median(tot):.4f})", end="")
    if timing:
        m = np.mean(np.asarray(rows), axis=0)
        print(f"   policy {m[0]:.4f} | physics {m[1]:.4f} | bookkeeping {m[2]:.4f} | background join {m[3]:.4f} | outside the marks {np.mean(tot) - m.sum():.4f}")
    else:
        print()
