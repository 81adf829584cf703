"""Debug: one-substep error of the CUDA kernel vs fp64 oracle in regimes (free flight / contacts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from mqe_b200 import engine as E, scene as S
from mqe_b200.envs import configs as C

def run(z, iters=None, label=""):
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 64
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0, policy_mode=E.POLICY_FP32, solver_iters=iters)
    eng = E.Engine(sc.desc, device=0, keepalive=sc)
    o64, o32 = oracle.Oracle(sc, "f64"), oracle.Oracle(sc, "f32")
    eng.reset(); o64.reset(); o32.reset()
    root = o64.get(E.BUF_ROOT_STATES).reshape(64, -1, 13).copy()
    root[:, :2, 2] = z
    o64.set(E.BUF_ROOT_STATES, root)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    errs = {k: [] for k in ("gv", "gq", "fv", "fq")}
    for s in range(8):
        r, d = o64.get(E.BUF_ROOT_STATES), o64.get(E.BUF_DOF_STATES)
        eng.tensor(E.BUF_ROOT_STATES).copy_(dev(r).view_as(eng.tensor(E.BUF_ROOT_STATES)))
        eng.tensor(E.BUF_DOF_STATES).copy_(dev(d).view_as(eng.tensor(E.BUF_DOF_STATES)))
        o32.set(E.BUF_ROOT_STATES, r); o32.set(E.BUF_DOF_STATES, d)
        eng.substeps(1); o32.substeps(1); o64.substeps(1)
        torch.cuda.synchronize()
        t = o64.get(E.BUF_ROOT_STATES).reshape(-1, 13); td = o64.get(E.BUF_DOF_STATES).reshape(-1, 2)
        g = eng.tensor(E.BUF_ROOT_STATES).cpu().numpy().reshape(-1, 13); gd = eng.tensor(E.BUF_DOF_STATES).cpu().numpy().reshape(-1, 2)
        f = o32.get(E.BUF_ROOT_STATES).reshape(-1, 13); fd = o32.get(E.BUF_DOF_STATES).reshape(-1, 2)
        errs["gv"].append(np.abs(g[:, 7:] - t[:, 7:]).ravel()); errs["gq"].append(np.abs(gd[:, 1] - td[:, 1]))
        errs["fv"].append(np.abs(f[:, 7:] - t[:, 7:]).ravel()); errs["fq"].append(np.abs(fd[:, 1] - td[:, 1]))
        if s == 0:
            i = np.argmax(np.abs(gd[:, 1] - td[:, 1]))
            print(label, "stats", eng.tensor(E.BUF_STATS).cpu().numpy()[:4], "worst qd idx", i, gd[i], td[i], fd[i])
            tg, to = eng.tensor(E.BUF_TORQUES).cpu().numpy().ravel(), o64.get(E.BUF_TORQUES)
            print(label, "torque err max", np.abs(tg - to).max(), "f32", np.abs(o32.get(E.BUF_TORQUES) - to).max())
    for k, v in errs.items():
        v = np.concatenate(v)
        print(label, k, "p50 %.2e p99 %.2e max %.2e" % tuple(np.percentile(v, [50, 99, 100])))
    eng.close()

run(1.0, label="free-flight")
run(0.33, iters=0, label="contacts-0-iters")
run(0.33, iters=1, label="contacts-1-iter")
run(0.33, label="contacts-8-iters")
