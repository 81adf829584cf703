"""Per-warp duration distribution of k_substeps on the bench workload (MQE_BUF_WARP_TRACE): is the launch bound by the
average warp or by its slowest one?  Usage: python tools/warp_trace.py [task] [num_envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np, torch
import bench as B
from mqe_b200 import engine as E
from mqe_b200.envs.utils import make_mqe_env, custom_cfg

task = sys.argv[1] if len(sys.argv) > 1 else "go1gate"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
eargs = SimpleNamespace(num_envs=n, seed=0, headless=True, record_video=False, sim_device="cuda:0")
env, cfg = make_mqe_env(task, eargs, custom_cfg(eargs), policy_mode=E.POLICY_BF16X3)
base = env.env
eng = base.engine
acts = torch.as_tensor(B.synth_actions(n, base._ctrl_agents, 64, env_offset=0), device="cuda:0")
env.reset()
if os.environ.get("MQE_SYNC_EPISODES", "0") != "1":       # steady state as in bench.py: random episode phases + one episode of pre-roll
    B.desynchronise(base, torch, seed=0)
    for i in range(int(base.max_episode_length) + 1):
        env.step(acts[i % 64])
rows = []
worst = (0.0, None, None)
for i in range(steps):
    env.step(acts[i % 64])
    if i >= 50 and i % 25 == 0:
        torch.cuda.synchronize()
        tr = eng.tensor(E.BUF_WARP_TRACE).cpu().numpy().astype(np.int64)
        d = (tr[:, 1] - tr[:, 0]) * 1e-3
        span = (tr[:, 1].max() - tr[:, 0].min()) * 1e-3
        late = (tr[:, 0].max() - tr[:, 0].min()) * 1e-3
        k = int(np.argmax(d))
        print(f"step {i:4d} span {span:7.1f} mean {d.mean():7.1f} p99 {np.percentile(d, 99):7.1f} max {d.max():7.1f}  warps rows>=13 {np.mean(tr[:, 3] >= 13):.3f} rows>=19 {np.mean(tr[:, 3] >= 19):.3f} pairs {np.mean(tr[:, 2] > 0):.4f}")
        if d.max() > worst[0]:
            worst = (float(d.max()), i, tr[k].copy())
        rows.append((span, d.mean(), np.percentile(d, 50), np.percentile(d, 90), np.percentile(d, 99), d.max(), late, tr[k, 2], tr[k, 3], (tr[:, 2] > 0).mean()))
r = np.array(rows, dtype=np.float64)
names = ["span_us", "mean", "p50", "p90", "p99", "max", "last_start", "pairs@slowest", "rows@slowest", "frac_warps_with_pairs"]
print(task, n, "k_substeps per-warp durations [us], mean over", len(rows), "sampled launches")
for i, nm in enumerate(names):
    print(f"  {nm:22s} {r[:, i].mean():10.2f}   (max {r[:, i].max():.2f})")
# dependence of warp time on its load (last sampled launch)
for lo, hi in [(0, 7), (7, 13), (13, 19), (19, 29)]:
    m = (tr[:, 3] >= lo) & (tr[:, 3] < hi) & (tr[:, 2] == 0)
    if m.any():
        print(f"  warps with widest row count in [{lo},{hi}) and no pairs: n={int(m.sum()):5d} mean {d[m].mean():8.2f} us")
if tr[:, 4:].sum() > 0:
    ph = tr[:, 4:20].astype(np.float64)
    names_p = ["P1 actuator", "P2 dynamics", "P3 pass2 rows", "P3b pairs", "P4 PGS", "P5 integrate", "prologue", "epilogue", "P3 limits", "P3 pass1 probes", "P4 setup", "P3b broadphase+publish+capmask", "P3b narrow", "P3b pair rows", "P4 pair sweeps", "-"]
    tot = ph.sum(1)
    print("  per-phase share of warp cycles (mean over warps; MQE_TRACE=1), cycles per launch:")
    for i, nm in enumerate(names_p):
        print(f"    {nm:14s} {100 * (ph[:, i] / tot).mean():6.2f} %   {ph[:, i].mean():10.0f} cyc   slowest warp {ph[k, i]:10.0f}")
m = tr[:, 2] > 0
if m.any():
    print(f"  warps with pair contacts: n={int(m.sum())} mean {d[m].mean():.2f} us, max {d[m].max():.2f}")

if worst[2] is not None and worst[2][4:].sum() > 0:
    w = worst[2]
    nm = ["P1 actuator", "P2 dynamics", "P3 pass2 rows", "P3b pairs", "P4 PGS", "P5 integrate", "prologue", "epilogue", "P3 limits", "P3 pass1 probes", "P4 setup", "P3b broadphase+publish+capmask", "P3b narrow", "P3b pair rows", "P4 pair sweeps", "-"]
    print(f"slowest warp of the run: step {worst[1]}, {worst[0]:.1f} us, pair contacts {w[2]}, widest rows {w[3]}; cycles per phase:")
    print("   " + ", ".join(f"{n} {int(c)}" for n, c in zip(nm, w[4:20]) if c))
