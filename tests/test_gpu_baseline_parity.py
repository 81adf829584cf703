"""GPU parity at the BASELINE.json sizes, over a whole episode, and across shards (VERDICT r1 "next round" item 1).

* `test_baseline_size_parity`: the CUDA path against the fp32 oracle at the sizes BASELINE.json names -- go1gate 4096, go1sheep-hard 4096,
  go1seesaw 8192, go1football-defender 4096 x 3 agents (config 5's per-GPU share) -- five full policy steps from reset; flags bit-exact.
* `test_teacher_forced_episode_parity`: every physics substep of a whole episode (500 policy steps x 4 substeps, 1024 envs) is started from
  the ORACLE's state, so the kernel is judged on the state distribution of a real episode (landing, walking, falling, wall and robot-robot
  contacts, resets) without the chaotic divergence of a free run; the free-running curve is printed beside it.
* `test_shard_independence_bit_exact`: (0,32)+(32,64) and an 8-way split reproduce the unsharded run BIT FOR BIT (state, flags, wrapper obs).
* `test_physx_golden_trajectories`: consumes tests/golden/physx_*.npz recorded by tools/record_physx_golden.py on the unmodified reference
  under real Isaac Gym; skipped while no such recording exists (Isaac Gym is a closed binary, absent here).

Tolerances are fp32 tolerances stated at each assert; integer / flag bookkeeping is bit-exact.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from mqe_b200 import engine as E  # noqa: E402
from mqe_b200 import scene as S  # noqa: E402
from mqe_b200.envs import configs as C  # noqa: E402

TASKS = {"go1gate": C.Go1GateCfg, "go1sheep-hard": C.NineSheepCfg, "go1seesaw": C.Go1SeesawCfg, "go1football-defender": C.Go1FootballDefenderCfg}
FLAGS = (E.BUF_EPISODE_LENGTH, E.BUF_TIMEOUT, E.BUF_RESET, E.BUF_COLLIDE, E.BUF_ROLL_TERM, E.BUF_PITCH_TERM, E.BUF_ZLOW_TERM, E.BUF_ZHIGH_TERM)


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")


def get(eng, which):
    return eng.tensor(which).cpu().numpy()


def actions_for(n, a_ctrl, step, seed=0):
    rng = np.random.default_rng(seed * 1000 + step)
    return rng.uniform(-1, 1, size=(n, a_ctrl, 3)).astype(np.float32)


def build(task, n, mode=E.POLICY_BF16X3, seed=0, env_slice=None, episode_s=None, threads=None):
    cfg = TASKS[task]()
    cfg.env.num_envs = n
    if episode_s is not None:
        cfg.env.episode_length_s = episode_s
    np.random.seed(seed)
    sc = S.build_scene(cfg, seed=seed, policy_mode=mode, env_slice=env_slice, wrapper_action_scale=(2.0, 0.5, 0.5))
    return cfg, sc


@pytest.mark.parametrize("task,n", [("go1gate", 4096), ("go1sheep-hard", 4096), ("go1seesaw", 8192), ("go1football-defender", 4096)])
def test_baseline_size_parity(task, n):
    """BASELINE.json configs C2 / C3 / C4 / C5-per-GPU: 5 x Go1.step() on identical seeds and actions, tensor-core policy (the shipped mode)."""
    cfg, sc = build(task, n)
    eng, orc = E.Engine(sc.desc, device=0, keepalive=sc), oracle.Oracle(sc, "f32")
    oracle.set_threads(os.cpu_count() or 1, "f32")
    eng.reset(); orc.reset()
    a_ctrl = sc.num_agents - 1 if sc.desc.defender else sc.num_agents
    report = []
    for s in range(5):
        act = actions_for(n, a_ctrl, s)
        eng.step(dev(act).data_ptr()); orc.step(act)
        torch.cuda.synchronize()
        for buf in FLAGS:
            assert np.array_equal(get(eng, buf), orc.get(buf)), (task, s, buf)              # bookkeeping: bit-exact on every env
        G = sc.num_agents + sc.num_npcs
        r_g, r_o = get(eng, E.BUF_ROOT_STATES).reshape(n, G, 13), orc.root_states()
        d_g, d_o = get(eng, E.BUF_DOF_STATES).reshape(n, -1, 2), orc.dof_states()
        e_pos = np.abs(r_g[..., :7] - r_o[..., :7]).reshape(n, -1).max(1)
        e_vel = np.abs(r_g[..., 7:] - r_o[..., 7:]).reshape(n, -1).max(1)
        e_q, e_qd = np.abs(d_g[..., 0] - d_o[..., 0]).max(1), np.abs(d_g[..., 1] - d_o[..., 1]).max(1)
        report.append((s, e_pos.max(), np.quantile(e_pos, 0.999), e_vel.max(), e_q.max(), e_qd.max(), float((e_pos < 1e-4).mean())))
        if s == 0:       # one policy step: pure rounding (measured ~2e-7 pos / 1e-5 qd at small sizes)
            assert e_pos.max() < 1e-5 and e_q.max() < 1e-5, report[-1]         # (1 ulp of a 60 m coordinate is 3.8e-6)
            assert e_vel.max() < 2e-4 and e_qd.max() < 2e-3, report[-1]
    for row in report:
        print(task, n, "step %d: pos max %.2e p99.9 %.2e | root vel max %.2e | q max %.2e | qd max %.2e | envs within 1e-4: %.4f" % row)
    s, pos_max, pos_p999, vel_max, q_max, qd_max, frac = report[-1]
    # after five policy steps (20 substeps with landing contacts) every env is still on the oracle's trajectory: SURVEY 8(c) tier 2
    # asks 1e-4 on positions; a handful of envs out of thousands may have switched a contact one substep apart (fp32 summation order)
    assert pos_p999 < 1e-4 and frac > 0.998, report[-1]
    assert pos_max < 5e-3 and q_max < 2e-2, report[-1]
    eng.close(); orc.close()


def test_teacher_forced_episode_parity():
    """Whole-episode substep parity with teacher forcing, go1gate 1024 envs x 2 agents x 500 policy steps x 4 substeps.

    Every substep: kernel state <- oracle state (root, dof; actions once per policy step), both advance ONE substep, compare.  The
    actuator-net histories are never synced: they are functions of the synced states, so they track to rounding on their own.
    Bounds are those of test_single_substep_parity applied to EVERY substep of the episode (fp32 oracle as the reference here)."""
    n, steps = 1024, 500
    cfg, sc = build("go1gate", n, mode=E.POLICY_FP32, episode_s=6.0)                        # time-outs after 300 policy steps: resets inside the run
    eng, orc = E.Engine(sc.desc, device=0, keepalive=sc), oracle.Oracle(sc, "f32")
    free = E.Engine(sc.desc, device=0, keepalive=sc)                                        # free-running twin for the curve beside it
    oracle.set_threads(os.cpu_count() or 1, "f32")
    eng.reset(); orc.reset(); free.reset()
    t_root, t_dof, t_act = eng.tensor(E.BUF_ROOT_STATES), eng.tensor(E.BUF_DOF_STATES), eng.tensor(E.BUF_ACTIONS)
    dec = int(sc.desc.decimation)
    worst = np.zeros(4)
    q999 = np.zeros(4)
    count_mismatch = 0
    contacts = 0
    curve = {}
    resets = 0
    for s in range(steps):
        act = actions_for(n, 2, s)
        orc.policy(act)
        t_act.copy_(dev(orc.get(E.BUF_ACTIONS)).view_as(t_act))
        for k in range(dec):
            t_root.copy_(dev(orc.get(E.BUF_ROOT_STATES)).view_as(t_root))
            t_dof.copy_(dev(orc.get(E.BUF_DOF_STATES)).view_as(t_dof))
            eng.substeps(1); orc.substeps(1)
            torch.cuda.synchronize()
            st_g, st_o = get(eng, E.BUF_STATS), orc.get(E.BUF_STATS)
            count_mismatch += int(tuple(st_g[:3]) != tuple(st_o[:3]))
            contacts += int(st_o[0])
            r_g, r_o = get(eng, E.BUF_ROOT_STATES).reshape(-1, 13), orc.get(E.BUF_ROOT_STATES).reshape(-1, 13)
            d_g, d_o = get(eng, E.BUF_DOF_STATES).reshape(-1, 2), orc.get(E.BUF_DOF_STATES).reshape(-1, 2)
            errs = (np.abs(r_g[:, :7] - r_o[:, :7]), np.abs(r_g[:, 7:] - r_o[:, 7:]), np.abs(d_g[:, 0] - d_o[:, 0]), np.abs(d_g[:, 1] - d_o[:, 1]))
            for i, e in enumerate(errs):
                worst[i] = max(worst[i], float(e.max()))
                q999[i] = max(q999[i], float(np.quantile(e, 0.999)))
        orc.post_physics()
        resets += int(orc.get(E.BUF_RESET).sum())
        free.step(dev(act).data_ptr())
        if s + 1 in (10, 25, 50, 100, 250, 500):
            torch.cuda.synchronize()
            fr, orr = get(free, E.BUF_ROOT_STATES).reshape(n, 2, 13), orc.root_states()
            ok = (np.abs(fr[..., :7] - orr[..., :7]).reshape(n, -1).max(1) < 1e-4) & (np.abs(fr[..., 7:] - orr[..., 7:]).reshape(n, -1).max(1) < 1e-3)
            same_flags = np.array_equal(get(free, E.BUF_RESET), orc.get(E.BUF_RESET))
            curve[s + 1] = (float(ok.mean()), same_flags)
    names = ("pos/quat", "root vel", "q", "qd")
    print("teacher-forced go1gate %d envs x %d policy steps x %d substeps: %d contacts, %d resets, substeps with differing contact counts: %d / %d"
          % (n, steps, dec, contacts, resets, count_mismatch, steps * dec))
    for i, nm in enumerate(names):
        print("  %-9s worst over all substeps: max %.2e, p99.9 %.2e" % (nm, worst[i], q999[i]))
    print("  free-running twin, fraction of envs inside 1e-4 pos / 1e-3 vel (and reset flags identical): ", curve)
    assert contacts > 100 * steps and resets > n // 2                       # the episode really exercised contacts and resets
    # single-substep bounds (tests/test_gpu_parity.py::test_single_substep_parity) on every substep of the episode, fp32 vs fp32:
    assert q999[1] < 2e-4 and q999[3] < 2e-3, q999                           # velocities: the fp32 conditioning floor (0.06 kg foot on a 5 kg trunk)
    assert q999[0] < 5e-6 and q999[2] < 2e-5, q999                           # positions / joint angles: dt x the velocity error (+ 1 ulp)
    assert worst[0] < 2e-4 and worst[2] < 2.5e-3, worst                      # rare contact on/off disagreements at the offset threshold stay small
    assert worst[1] < 5e-2 and worst[3] < 0.5, worst
    assert count_mismatch <= 0.02 * steps * dec, count_mismatch
    assert curve[10][0] == 1.0 and curve[10][1], curve                      # free run: every env inside the tier-2 band for the first 10 policy steps
    eng.close(); orc.close(); free.close()


@pytest.mark.parametrize("task", ["go1gate", "go1sheep-hard", "go1football-defender"])
def test_shard_independence_bit_exact(task):
    """results do not depend on how the env batch is split over GPUs.  One engine on envs (0,64) against two engines on
    (0,32)+(32,64) and eight on 8-env slices, same global seeds and actions, short episodes so resets (counter RNG keyed by the GLOBAL env
    id, sheep flocking noise) are exercised: root / dof state, flags, observation rows and the fused wrapper obs / reward bit for bit."""
    from types import SimpleNamespace
    from mqe_b200.envs import make_mqe_env
    N = 64

    def make(sl):
        args = SimpleNamespace(num_envs=N, seed=1, headless=True, record_video=False, sim_device="cuda:0")

        def cc(cfg):
            cfg.env.num_envs = N
            cfg.env.episode_length_s = 0.24                                   # 12 policy steps: every env resets several times
            return cfg
        env, _ = make_mqe_env(task, args, cc, env_slice=sl, policy_mode=E.POLICY_BF16X3)
        return env

    for split in ([(0, 32), (32, 64)], [(8 * i, 8 * i + 8) for i in range(8)]):
        full = make((0, N))
        shards = [make(sl) for sl in split]
        full.reset()
        for sh in shards:
            sh.reset()
        A = full.env._ctrl_agents
        rng = np.random.default_rng(3)
        bad = []
        for s in range(30):
            a = rng.uniform(-1.2, 1.2, size=(N, A, 3)).astype(np.float32)
            out_f = full.step(dev(a))
            outs = [sh.step(dev(a[lo:hi])) for sh, (lo, hi) in zip(shards, split)]
            torch.cuda.synchronize()
            for which in (E.BUF_ROOT_STATES, E.BUF_DOF_STATES, E.BUF_OBS, E.BUF_TORQUES, E.BUF_ACTIONS) + FLAGS:
                g = get(full.env.engine, which)
                parts = np.concatenate([get(sh.env.engine, which) for sh in shards], axis=0)
                if not (g.shape == parts.shape and np.array_equal(g.view(np.uint8), parts.view(np.uint8))):
                    bad.append((s, which, float(np.abs(g.astype(np.float64) - parts.astype(np.float64)).max())))
            cf = np.concatenate([get(sh.env.engine, E.BUF_CONTACT_FORCES) for sh in shards], axis=0)       # summed with shared-memory atomics:
            if not np.allclose(get(full.env.engine, E.BUF_CONTACT_FORCES), cf, rtol=1e-5, atol=1e-4):      # order-insensitive up to rounding only
                bad.append((s, "contact forces"))
            if torch.is_tensor(out_f[0]):                                     # fused task-wrapper observation and reward
                if not torch.equal(out_f[0], torch.cat([o[0] for o in outs], dim=0)):
                    bad.append((s, "wrapper obs"))
                if not torch.equal(out_f[1], torch.cat([o[1] for o in outs], dim=0)):
                    bad.append((s, "wrapper reward"))
            if not torch.equal(out_f[2], torch.cat([o[2] for o in outs], dim=0)):
                bad.append((s, "done"))
        assert not bad, (task, len(split), bad[:10])
        assert int(get(full.env.engine, E.BUF_EPISODE_LENGTH).max()) <= 13    # episodes really were short: several resets per env
        for sh in shards:
            sh.close()
        full.close()


def test_physx_golden_trajectories(golden_dir):
    """SURVEY 8(c) tier 4: trajectories recorded from the UNMODIFIED reference under real Isaac Gym (tools/record_physx_golden.py writes
    tests/golden/physx_<task>.npz: seeds, the action stream of bench.synth_actions, root / dof states per policy step).  Compared with the
    tolerances the survey states for PhysX-vs-own-engine: flags identical while both are in the band, 2 cm / 0.05 rad over the first 25
    policy steps.  Until somebody with Isaac Gym records them the physics stays GPU-vs-own-oracle only ("PhysX parity unpinned")."""
    files = sorted(glob.glob(os.path.join(golden_dir, "physx_*.npz")))
    if not files:
        pytest.skip("no tests/golden/physx_*.npz: Isaac Gym (closed binary) is needed to record them -- see tools/record_physx_golden.py")
    for f in files:
        z = np.load(f)
        task, n = str(z["task"]), int(z["num_envs"])
        cfg = {**TASKS, "go1sheep-easy": C.SingleSheepCfg}[task]()
        cfg.env.num_envs = n
        np.random.seed(int(z["seed"]))
        sc = S.build_scene(cfg, seed=int(z["seed"]), policy_mode=E.POLICY_BF16X3, wrapper_action_scale=(2.0, 0.5, 0.5))
        eng = E.Engine(sc.desc, device=0, keepalive=sc)
        eng.reset()
        # the reference's reset draws from torch's RNG stream: start from ITS post-reset state instead of ours
        eng.tensor(E.BUF_ROOT_STATES).copy_(dev(z["root"][0]).view_as(eng.tensor(E.BUF_ROOT_STATES)))
        eng.tensor(E.BUF_DOF_STATES).copy_(dev(z["dof"][0]).view_as(eng.tensor(E.BUF_DOF_STATES)))
        steps = min(25, z["actions"].shape[0])
        for s in range(steps):
            eng.step(dev(z["actions"][s]).data_ptr())
            torch.cuda.synchronize()
            r = get(eng, E.BUF_ROOT_STATES).reshape(z["root"][s + 1].shape)
            alive = ~z["reset"][: s + 1].any(0)
            assert np.abs(r[alive][..., :3] - z["root"][s + 1][alive][..., :3]).max() < 2e-2, (task, s)
            assert np.abs(get(eng, E.BUF_DOF_STATES).reshape(z["dof"][s + 1].shape)[alive][..., 0] - z["dof"][s + 1][alive][..., 0]).max() < 5e-2, (task, s)
        eng.close()


def test_substep_logs_lazy_and_match_oracle():
    """post_decimation_step logs (legged_robot.py:112-115): absent until somebody reads them, then filled every step; values against the oracle."""
    cfg, sc = build("go1gate", 32, mode=E.POLICY_FP32)
    eng, orc = E.Engine(sc.desc, device=0, keepalive=sc), oracle.Oracle(sc, "f32")
    eng.reset(); orc.reset()
    a = actions_for(32, 2, 0)
    eng.step(dev(a).data_ptr()); orc.step(a)
    tau = eng.tensor(E.BUF_SUBSTEP_TORQUES)                                   # first access switches the logs on (zeros, as _init_buffers leaves them)
    assert tuple(tau.shape) == (32, 4, 24) and float(tau.abs().sum()) == 0.0
    for s in range(1, 4):
        a = actions_for(32, 2, s)
        eng.step(dev(a).data_ptr()); orc.step(a)
    torch.cuda.synchronize()
    assert np.allclose(get(eng, E.BUF_SUBSTEP_TORQUES).ravel(), orc.get(E.BUF_SUBSTEP_TORQUES), atol=5e-3)
    assert np.allclose(get(eng, E.BUF_SUBSTEP_DOF_VEL).ravel(), orc.get(E.BUF_SUBSTEP_DOF_VEL), atol=5e-3)
    ex_g, ex_o = get(eng, E.BUF_SUBSTEP_EXCEED).ravel(), orc.get(E.BUF_SUBSTEP_EXCEED)
    assert (ex_g != ex_o).mean() < 0.002                                      # a joint exactly at the soft limit may round either way
    assert np.allclose(get(eng, E.BUF_SUBSTEP_TORQUES)[:, -1].ravel(), get(eng, E.BUF_TORQUES).ravel())     # last substep == `torques`
    eng.close(); orc.close()
    from mqe_b200.envs.go1 import Go1
    env = Go1(cfg, sim_device="cuda:0", seed=0)
    env.reset()
    assert env.substep_torques.shape == (32, 4, 24) and env.substep_exceed_dof_pos_limits.dtype == torch.bool and env.substep_dof_vel.shape == (32, 4, 24)
    env.close()


def test_first_reset_observation_has_spawn_gravity():
    """the very first reset() must already report projected_gravity = R(spawn)^T (0,0,-1) (legged_robot.py:570, 622) -- it is the
    gravity entry of the first walk-policy frame and stays in the 30-frame history for 30 steps."""
    from mqe_b200.envs.go1 import Go1
    for task in ("go1gate", "go1football-defender"):
        cfg = TASKS[task](); cfg.env.num_envs = 8
        np.random.seed(0)
        env = Go1(cfg, sim_device="cuda:0", seed=0)
        ob = env.reset()
        assert torch.allclose(ob.projected_gravity, torch.tensor([0.0, 0.0, -1.0], device="cuda:0").expand_as(ob.projected_gravity), atol=1e-6)
        assert float(ob.lin_vel.abs().max()) == 0.0                           # derived before the reset redraws the root velocity
        assert torch.allclose(ob.base_quat.norm(dim=1), torch.ones(ob.base_quat.shape[0], device="cuda:0"), atol=1e-6)
        a = torch.zeros(8, env._ctrl_agents, 3, device="cuda:0")
        env.step_from_wrapper(a)
        assert torch.allclose(env.locomotion_obs[:, :3], torch.tensor([0.0, 0.0, -1.0], device="cuda:0").expand(env.locomotion_obs.shape[0], 3), atol=1e-6)
        env.close()
    with pytest.raises(E.EngineError):
        Go1(TASKS["go1gate"](), sim_device="cpu")                            # BASELINE C1 has no counterpart: refuse, do not coerce


def test_step_result_double_buffer_and_host_adapter():
    """The packed step result: (1) a returned observation survives the next step() untouched (the engine alternates two halves);
    (2) the numpy adapter's one-copy host path (mqe_sim_step_host_result) returns exactly what the torch path returns;
    (3) the device-resident adapter path returns CUDA tensors; (4) MATWrapper passes plain spaces through."""
    from types import SimpleNamespace
    from mqe_b200.openrl_adapter import MATWrapper, make_env
    args = SimpleNamespace(task="go1sheep-hard", num_envs=48, seed=2, headless=True, record_video=False, sim_device="cuda:0")
    host, _ = make_env(args)
    ref, _ = make_env(args)
    ref._host = False                                                        # force the torch path (pinned staging, three copies)
    ref._host_ready = lambda: False
    o_h, o_r = host.reset(), ref.reset()
    assert isinstance(o_h, np.ndarray) and np.array_equal(o_h, o_r) and o_h.shape == (48, 2, 34)
    rng = np.random.default_rng(0)
    prev = None
    for s in range(12):
        a = rng.uniform(-2, 2, size=(48, 2, 3)).astype(np.float32)
        oh, rh, dh, ih = host.step(a)
        orr, rr, dr, ir = ref.step(a)
        assert host._host is not None and oh.shape == (48, 2, 34) and rh.shape == (48, 2, 1) and dh.shape == (48, 2) and dh.dtype == bool and len(ih) == 48
        assert np.array_equal(oh, orr) and np.array_equal(rh, rr) and np.array_equal(dh, dr), s
        if prev is not None:
            assert np.array_equal(prev[0], prev[1]), "the previous step's observation must survive one more step"
        prev = (oh, oh.copy())
    br_h, br_r = host.batch_rewards(), ref.batch_rewards()
    assert set(br_h) == set(br_r) and all(abs(br_h[k] - br_r[k]) <= 1e-6 * max(1.0, abs(br_r[k])) for k in br_h)
    host.close(); ref.close()
    devenv, _ = make_env(args, device_resident=True)
    o0 = devenv.reset()
    assert torch.is_tensor(o0) and o0.is_cuda
    o1, r1, d1, _ = devenv.step(torch.zeros(48, 2, 3, device="cuda:0"))
    keep = o1.clone()
    o2, r2, d2, _ = devenv.step(np.zeros((48, 2, 3), dtype=np.float32))
    assert o1.is_cuda and r1.shape == (48, 2, 1) and d1.shape == (48, 2) and torch.equal(o1, keep) and o2.data_ptr() != o1.data_ptr()
    m = MATWrapper(devenv)
    assert m.observation_space.shape == devenv.observation_space.shape and m.observation(o2) is o2
    devenv.close()


def test_incremental_policy_across_ring_wrap_and_resets():
    """The tensor-core policy as it runs inside a step -- incremental layer 0 (29 frames contracted behind the previous step's physics, the new
    frame in front of the fused tail), CUDA-graph replay -- against the oracle's plain forward pass, for longer than the 30-slot ring and across
    resets (short episodes: history rows are zeroed, the stale partial sums of those rows must be dropped), with stand-alone policy calls mixed in
    (they must re-prime the partial sums).  Raw policy outputs are compared: |Δ| <= 2e-2 on >= 99 % of the entries (policy outputs are O(1..3);
    the physics of the two sides drifts apart by ~1e-4 over these steps, which the policy amplifies), median <= 2e-4."""
    cfg, sc = build("go1gate", 32, mode=E.POLICY_BF16X3, episode_s=0.36)       # 18 policy steps per episode
    eng, orc = E.Engine(sc.desc, device=0, keepalive=sc), oracle.Oracle(sc, "f32")
    eng.reset(); orc.reset()
    errs, n_reset = [], 0
    for s in range(70):
        act = actions_for(32, 2, s)
        if s in (20, 41):                                # stand-alone preprocess_action + physics + post instead of the fused step
            eng.policy(dev(act).data_ptr()); eng.substeps(int(sc.desc.decimation)); eng.post_physics()
        else:
            eng.step(dev(act).data_ptr())
        orc.step(act)
        torch.cuda.synchronize()
        a_g, a_o = get(eng, E.BUF_LOC_ACTION).ravel(), orc.get(E.BUF_LOC_ACTION)
        errs.append(np.abs(a_g - a_o))
        assert np.array_equal(get(eng, E.BUF_RESET), orc.get(E.BUF_RESET)), s
        n_reset += int(orc.get(E.BUF_RESET).sum())
    e = np.concatenate(errs)
    print("incremental policy vs oracle over 70 steps: median %.2e p99 %.2e max %.2e, resets %d" % (np.median(e), np.quantile(e, 0.99), e.max(), n_reset))
    assert n_reset >= 64 and np.isfinite(e).all()
    assert np.median(e) <= 2e-4 and np.quantile(e, 0.99) <= 2e-2, (np.median(e), np.quantile(e, 0.99))
    eng.close(); orc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("task,n,mode", [("go1sheep-hard", 96, "1"), ("go1gate", 160, "1"), ("go1gate", 160, "2"), ("go1football-defender", 64, "1")])
def test_task_order_does_not_change_results(task, n, mode, monkeypatch):
    """k_balance_tasks re-orders which warp of the k_substeps grid integrates which env group (by the duration of each group's warp in the
    previous launch; MQE_BALANCE=1 grouped, 2 spread).  An env never looks at another one, so state, flags and observation rows must be
    BIT-identical to the identity order, step after step, through resets."""
    def run(balance):
        monkeypatch.setenv("MQE_BALANCE", balance)
        cfg, sc = build(task, n, mode=E.POLICY_BF16X3, episode_s=0.5)
        eng = E.Engine(sc.desc, device=0, keepalive=sc)
        eng.reset()
        a_ctrl = sc.num_agents - 1 if sc.desc.defender else sc.num_agents
        out = []
        for s in range(40):
            eng.step(dev(actions_for(n, a_ctrl, s)).data_ptr())
            torch.cuda.synchronize()
            out.append((get(eng, E.BUF_ROOT_STATES).copy(), get(eng, E.BUF_DOF_STATES).copy(), get(eng, E.BUF_OBS).copy(), get(eng, E.BUF_RESET).copy()))
        eng.close()
        return out
    ref, got = run("0"), run(mode)
    for s, (r, g) in enumerate(zip(ref, got)):
        for x, y in zip(r, g):
            assert np.array_equal(x, y), (task, mode, s)
    assert sum(int(r[3].sum()) for r in ref) > 0                     # resets happened


@pytest.mark.gpu
def test_stage_timing_marks():
    """mqe_sim_stage_timing / mqe_sim_stage_ms (ABI 7): event marks between the stages of a step, recorded as nodes of the step graph.  The four
    stage times are positive, add up to less than the step, and switching the marks on and off does not change results."""
    cfg, sc = build("go1gate", 256, mode=E.POLICY_BF16X3)
    eng = E.Engine(sc.desc, device=0, keepalive=sc)
    eng.reset()
    with pytest.raises(E.EngineError):
        eng.stage_ms()                                               # not enabled yet
    states = []
    for timing in (False, True, False):
        eng.stage_timing(timing)
        for s in range(4):                                           # plain steps, capture, replay
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); eng.step(dev(actions_for(256, 2, s)).data_ptr()); e1.record()
            torch.cuda.synchronize()
            if timing:
                st = eng.stage_ms()
                assert set(st) == {"policy", "physics", "bookkeeping", "background_join"}
                assert all(v >= 0.0 for v in st.values()) and st["policy"] > 0.0 and st["physics"] > 0.0
                assert sum(st.values()) <= e0.elapsed_time(e1) + 1e-3
        states.append(get(eng, E.BUF_ROOT_STATES).copy())
    assert np.isfinite(states[-1]).all()
    eng.close()
