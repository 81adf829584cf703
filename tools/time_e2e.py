#!/usr/bin/env python
"""Where does the host path spend its time?  Breaks mqe_openrl_wrapper.step (fused host path) into its pieces on one task."""
import os, sys, time
os.environ.setdefault("MQE_POLICY_MODE", "1")          # tensor-core policy (bf16x3), as bench.py; must be set before mqe_b200.engine is imported
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from types import SimpleNamespace
from mqe_b200 import engine as E
from mqe_b200.openrl_adapter import make_env

task = sys.argv[1] if len(sys.argv) > 1 else "go1sheep-hard"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
args = SimpleNamespace(task=task, num_envs=n, seed=0, headless=True, record_video=False, sim_device="cuda:0")
ad, cfg = make_env(args)
ad.reset()
A = ad._task.env._ctrl_agents
acts = np.random.default_rng(0).uniform(-2, 2, size=(16, n, A, 3)).astype(np.float32)
for i in range(300):
    ad.step(acts[i % 16])
eng = ad._task.env.engine
h = ad._host
def T(f, k=100):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(k): f(i)
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / k
d_act = torch.as_tensor(0.5 * acts, device="cuda:0")
print(task, n, "adapter.step            %.3f ms" % T(lambda i: ad.step(acts[i % 16])))
print(task, n, "engine.step_host_result %.3f ms" % T(lambda i: eng.step_host_result(h["act"], h["res"][i & 1])))
print(task, n, "engine.step (device)    %.3f ms" % T(lambda i: eng.step(d_act[i % 16].data_ptr())))
print(task, n, "engine.step + sync each %.3f ms" % T(lambda i: (eng.step(d_act[i % 16].data_ptr()), eng.synchronize())))
print(task, n, "np.multiply only        %.3f ms" % T(lambda i: np.multiply(acts[i % 16].reshape(h["act"].shape), 0.5, out=h["act"])))
print(task, n, "infos list of dicts     %.3f ms" % T(lambda i: [{} for _ in range(n)]))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(50): ad.step(acts[i % 16])
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
res = torch.empty(int(h["L"].total_bytes), dtype=torch.uint8, pin_memory=True)
src = eng.tensor(E.BUF_STEP_RESULT)[0]
print(task, n, "D2H result (torch pinned) %.3f ms" % T(lambda i: (res.copy_(src, non_blocking=True), torch.cuda.synchronize())))
