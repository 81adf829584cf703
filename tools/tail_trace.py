"""Phase timing of k_policy_tail (MQE_TRACE=1): global-timer stamps of the first epilogue thread of every CTA.
    MQE_TRACE=1 python tools/tail_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
import bench as B
from mqe_b200 import engine as E
from mqe_b200.envs.utils import make_mqe_env, custom_cfg
n = 4096
eargs = SimpleNamespace(num_envs=n, seed=0, headless=True, record_video=False, sim_device="cuda:0")
env, cfg = make_mqe_env("go1gate", eargs, custom_cfg(eargs), policy_mode={"bf16x3": E.POLICY_BF16X3, "bf16": E.POLICY_BF16}[os.environ.get("TAIL_MODE", "bf16x3")])
base, eng = env.env, env.env.engine
acts = torch.as_tensor(B.synth_actions(n, base._ctrl_agents, 64, env_offset=0), device="cuda:0")
env.reset()
for i in range(60):
    env.step(acts[i % 64])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
base._set_scale("wrapper")
rows = []
for i in range(20):
    flush.zero_()
    eng.policy(acts[i % 64].data_ptr())
    torch.cuda.synchronize()
    tr = eng.tensor(E.BUF_WARP_TRACE).cpu().numpy().astype(np.int64)[:64, :9]
    rows.append(tr - tr[:, :1].min())
    eng.substeps(4); eng.post_physics()
r = np.mean(np.asarray(rows[2:], dtype=np.float64), axis=0) * 1e-3          # [64 CTAs][9 marks] us since the first CTA started
names = ["start", "setup done (TMEM, constants, pdl_wait)", "adapt.2 accumulator ready", "latent ready", "body.2 A operand produced (8 chunks)",
         "body.2 accumulator ready", "body.4 A operand produced (4 chunks)", "body.4 accumulator ready", "actions written"]
prev = None
for k, nm in enumerate(names):
    col = r[:, k]
    d = "" if prev is None else f"   phase {np.mean(col - prev):6.2f} us"
    print(f"  mark {k} {nm:46s} mean {col.mean():7.2f}  max {col.max():7.2f}{d}")
    prev = col
