"""GPU parity tests proper: the CUDA path through the C ABI against the CPU oracle / the reference's own golden
vectors.  Run on the B200 box with `-m gpu`.

Tolerances (fp32 path, stated once): network outputs vs TorchScript fp32 goldens rtol 1e-4 / atol 1e-4 for the
CUDA-core path and rtol 2e-4 / atol 3e-4 for the tensor-core bf16x3 path; one physics substep is judged against the
fp64 oracle and must be as accurate as the fp32 oracle build is (p99 within 3x, max within 5x of that floor);
trajectories: 5e-6 (pos, quat, q) / 1e-4 (root vel) / 1e-3 (qd) after one policy step, 1e-4 (pos, q) after five --
short horizons only, because contact switching amplifies rounding differences (the oracle uses a dense mass matrix,
the kernel a block-LDL form).  Bookkeeping (episode counters, time-outs, reset masks, indices) is bit-exact.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

import oracle  # noqa: E402
from mqe_b200 import engine as E  # noqa: E402
from mqe_b200 import scene as S  # noqa: E402
from mqe_b200.envs import configs as C  # noqa: E402

TASKS = {"go1pushbox": C.Go1PushboxCfg, "go1gate": C.Go1GateCfg, "go1sheep-hard": C.NineSheepCfg, "go1sheep-easy": C.SingleSheepCfg,
         "go1revolvingdoor": C.Go1RotationCfg, "go1seesaw": C.Go1SeesawCfg, "go1football-defender": C.Go1FootballDefenderCfg, "go1plane": C.Go1PlaneCfg,
         "go1wrestling": C.Go1WrestlingCfg, "go1bridge": C.Go1BridgeCfg, "go1tug": C.Go1TugCfg}


def make_pair(task, n, mode=E.POLICY_FP32, seed=0, precision="f32"):
    cfg = TASKS[task]()
    cfg.env.num_envs = n
    np.random.seed(seed)
    sc = S.build_scene(cfg, seed=seed, policy_mode=mode, wrapper_action_scale=(2.0, 0.5, 0.5))
    eng = E.Engine(sc.desc, device=0, keepalive=sc)
    orc = oracle.Oracle(sc, precision)
    return sc, eng, orc


def actions_for(sc, step, seed=0):
    a_ctrl = sc.num_agents - 1 if sc.desc.defender else sc.num_agents
    rng = np.random.default_rng(seed * 1000 + step)
    return rng.uniform(-1, 1, size=(sc.num_envs, a_ctrl, 3)).astype(np.float32)


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")


def get(eng, which):
    return eng.tensor(which).cpu().numpy()


@pytest.mark.parametrize("mode,rtol,atol", [(E.POLICY_FP32, 1e-4, 1e-4), (E.POLICY_BF16X3, 2e-4, 3e-4)])
def test_policy_forward_golden(mode, rtol, atol, golden_dir):
    """walk-these-ways adaptation module + body vs the TorchScript goldens (SURVEY 8(c) known answers)."""
    z = np.load(os.path.join(golden_dir, "mlp_kat.npz"))
    sc, eng, orc = make_pair("go1gate", 128, mode)
    for key_x, key_a, key_l in (("x", "action", "latent"), ("xs", "actions", "latents")):
        x = z[key_x]
        lat, act = eng.policy_forward(dev(x))
        torch.cuda.synchronize()
        lat, act = lat.cpu().numpy(), act.cpu().numpy()
        scale = np.abs(z[key_a]).max()
        assert np.allclose(act, z[key_a], rtol=rtol, atol=atol * max(1.0, scale)), np.abs(act - z[key_a]).max()
        assert np.allclose(lat, z[key_l], rtol=rtol, atol=atol * max(1.0, np.abs(z[key_l]).max()))
    lat, act = eng.policy_forward(dev(z["x"]))
    assert np.allclose(act.cpu().numpy()[0, :4], [165.2033, -166.0023, 104.1395, 131.9488], atol=5e-2)
    eng.close()


def test_policy_bf16_single_pass_is_close(golden_dir):
    z = np.load(os.path.join(golden_dir, "mlp_kat.npz"))
    sc, eng, orc = make_pair("go1gate", 128, E.POLICY_BF16)
    _, act = eng.policy_forward(dev(z["xs"]))
    ref = z["actions"]
    rel = np.abs(act.cpu().numpy() - ref).max() / np.abs(ref).max()
    assert rel < 2e-2, rel
    eng.close()


def test_actuator_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "mlp_kat.npz"))
    sc, eng, orc = make_pair("go1gate", 4)
    t = eng.actuator_forward(dev(z["xa"])).cpu().numpy()
    assert np.allclose(t, [19.5816, 22.6831, -19.3611, -4.5704, -15.3245], atol=2e-4)
    t = eng.actuator_forward(dev(z["xas"])).cpu().numpy()
    assert np.allclose(t, z["torques"].ravel(), rtol=1e-5, atol=3e-5)
    eng.close()


@pytest.mark.parametrize("task", ["go1gate", "go1sheep-hard", "go1football-defender", "go1seesaw", "go1plane", "go1wrestling", "go1bridge", "go1tug"])
def test_reset_matches_oracle(task):
    """reset_idx + compute_observations: same counter RNG, same arithmetic."""
    sc, eng, orc = make_pair(task, 64)
    eng.reset(); orc.reset()
    torch.cuda.synchronize()
    for buf in (E.BUF_ROOT_STATES, E.BUF_DOF_STATES, E.BUF_OBS):
        assert np.allclose(get(eng, buf).ravel(), orc.get(buf), rtol=1e-6, atol=2e-6), buf   # 1 ulp: FMA contraction differs
    assert np.array_equal(get(eng, E.BUF_RESET), orc.get(E.BUF_RESET))
    assert np.array_equal(get(eng, E.BUF_EPISODE_LENGTH), orc.get(E.BUF_EPISODE_LENGTH))
    eng.close()


def _sync_state(eng, src, others=()):
    root, dof = src.get(E.BUF_ROOT_STATES), src.get(E.BUF_DOF_STATES)
    eng.tensor(E.BUF_ROOT_STATES).copy_(dev(root).view_as(eng.tensor(E.BUF_ROOT_STATES)))
    eng.tensor(E.BUF_DOF_STATES).copy_(dev(dof).view_as(eng.tensor(E.BUF_DOF_STATES)))
    for o in others:
        o.set(E.BUF_ROOT_STATES, root); o.set(E.BUF_DOF_STATES, dof)


def _state_err(a_root, a_dof, b_root, b_dof):
    return (np.abs(a_root[:, :7] - b_root[:, :7]), np.abs(a_root[:, 7:] - b_root[:, 7:]),
            np.abs(a_dof[:, 0] - b_dof[:, 0]), np.abs(a_dof[:, 1] - b_dof[:, 1]))


@pytest.mark.parametrize("task", ["go1gate", "go1sheep-hard", "go1football-defender", "go1seesaw", "go1pushbox", "go1revolvingdoor", "go1wrestling", "go1bridge", "go1tug"])
def test_single_substep_parity(task):
    """Single physics substeps from IDENTICAL states, robots dropped onto the floor so foot / knee / pair contacts,
    joint limits and the actuator net are all active.  Truth = the fp64 oracle; the fp32 oracle (dense Cholesky) run
    beside it measures the fp32 conditioning floor of the problem (mass matrix condition ~1e5: 0.06 kg feet on a
    5 kg trunk), and the CUDA kernel (block-LDL, fp32) must be as close to fp64 as that restatement is."""
    sc, eng, o32 = make_pair(task, 64)
    o64 = oracle.Oracle(sc, "f64")
    eng.reset(); o32.reset(); o64.reset()
    root = o64.get(E.BUF_ROOT_STATES).reshape(sc.num_envs, -1, 13).copy()
    root[:, :sc.num_agents, 2] = 0.30 + np.linspace(0.0, 0.06, sc.num_envs)[:, None]     # some penetrating, some hovering
    if task == "go1seesaw":       # half of the envs: robots standing on the plank (ramp from x = 3.5 to 7.5, pivot at 5.6 / 0.545)
        n = sc.num_envs
        dofs = o64.get(E.BUF_DOF_STATES).reshape(n, -1, 2).copy()
        dofs[:, 24, 0] = -0.238
        xs = np.linspace(3.7, 7.2, n)
        for a in range(2):
            on = np.arange(n) % 2 == 0
            xa = xs + 0.3 * a
            zpl = 0.545 + (xa - 5.6) * np.tan(0.238) + 0.015 / np.cos(0.238)
            root[on, a, 0] = sc.env_origins[on, 0] + xa[on]
            root[on, a, 1] = sc.env_origins[on, 1] + (0.25 if a else -0.25)
            root[on, a, 2] = zpl[on] + 0.29 + 0.02 * (np.arange(n)[on] % 3)
        for o in (o32, o64):
            o.set(E.BUF_DOF_STATES, dofs)
        eng.tensor(E.BUF_DOF_STATES).copy_(dev(dofs).view_as(eng.tensor(E.BUF_DOF_STATES)))
    if task == "go1pushbox":      # half of the envs: both robots pressed against the -x / -y faces of the 1 m box
        n = sc.num_envs
        on = np.arange(n) % 2 == 0
        bx = root[:, 2, :3]
        root[on, 0, 0] = bx[on, 0] - 0.5 - 0.28 + 0.02 * (np.arange(n)[on] % 4); root[on, 0, 1] = bx[on, 1] - 0.2
        root[on, 1, 0] = bx[on, 0] + 0.1; root[on, 1, 1] = bx[on, 1] - 0.5 - 0.15
        root[on, 2, 2] = 0.52
    if task == "go1revolvingdoor":   # half of the envs: one robot pressed on each face of the (already turning) door panel
        n = sc.num_envs
        on = np.arange(n) % 2 == 0
        dofs = o64.get(E.BUF_DOF_STATES).reshape(n, -1, 2).copy()
        dofs[:, 24, 0] = np.linspace(-0.3, 0.3, n); dofs[:, 24, 1] = np.linspace(-1.0, 1.0, n)
        hinge = root[:, 2, :3]
        root[on, 0, 0] = hinge[on, 0] - 0.04 - 0.28 + 0.02 * (np.arange(n)[on] % 4); root[on, 0, 1] = hinge[on, 1] - 0.55
        root[on, 1, 0] = hinge[on, 0] + 0.04 + 0.28 - 0.02 * (np.arange(n)[on] % 3); root[on, 1, 1] = hinge[on, 1] + 0.55
        root[on, 1, 3:7] = (0.0, 0.0, 1.0, 0.0)                                       # facing -x, nose on the far face
        for o in (o32, o64):
            o.set(E.BUF_DOF_STATES, dofs)
        eng.tensor(E.BUF_DOF_STATES).copy_(dev(dofs).view_as(eng.tensor(E.BUF_DOF_STATES)))
    o64.set(E.BUF_ROOT_STATES, root)
    a = np.clip(np.random.default_rng(1).normal(0, 1.0, size=(sc.num_envs * sc.num_agents * 12,)), -3, 3).astype(np.float32)
    for o in (o32, o64):
        o.set(E.BUF_ACTIONS, a)
    eng.tensor(E.BUF_ACTIONS).copy_(dev(a).view_as(eng.tensor(E.BUF_ACTIONS)))
    names = ("pos/quat", "root vel", "q", "qd")
    e_gpu, e_f32 = [[] for _ in range(4)], [[] for _ in range(4)]
    total_contacts, cf_err = 0, 0.0
    for s in range(24):
        _sync_state(eng, o64, (o32,))
        eng.substeps(1); o32.substeps(1); o64.substeps(1)
        torch.cuda.synchronize()
        st_g, st_o = get(eng, E.BUF_STATS), o64.get(E.BUF_STATS)
        assert tuple(st_g[:3]) == tuple(st_o[:3]), f"substep {s}: contact/limit/pair counts differ gpu {st_g[:4]} oracle {st_o[:4]}"
        total_contacts += int(st_o[0])
        R = lambda x: x.reshape(-1, 13)
        D = lambda x: x.reshape(-1, 2)
        t_root, t_dof = R(o64.get(E.BUF_ROOT_STATES)), D(o64.get(E.BUF_DOF_STATES))
        for acc, (r, d) in ((e_gpu, (R(get(eng, E.BUF_ROOT_STATES)), D(get(eng, E.BUF_DOF_STATES)))),
                            (e_f32, (R(o32.get(E.BUF_ROOT_STATES)), D(o32.get(E.BUF_DOF_STATES))))):
            for k, e in enumerate(_state_err(r, d, t_root, t_dof)):
                acc[k].append(e.ravel())
        assert np.allclose(get(eng, E.BUF_TORQUES).ravel(), o64.get(E.BUF_TORQUES), atol=2e-3)
        cf_g, cf_o = get(eng, E.BUF_CONTACT_FORCES).ravel(), o64.get(E.BUF_CONTACT_FORCES)
        cf_err = max(cf_err, float(np.abs(cf_g - cf_o).max() / max(1.0, np.abs(cf_o).max())))
    assert total_contacts > 0
    for k, name in enumerate(names):
        g, f = np.concatenate(e_gpu[k]), np.concatenate(e_f32[k])
        pg, pf = np.percentile(g, [50, 99, 100]), np.percentile(f, [50, 99, 100])
        print(f"{task} {name}: |gpu-f64| p50/p99/max = {pg[0]:.2e}/{pg[1]:.2e}/{pg[2]:.2e}   |f32 oracle-f64| = {pf[0]:.2e}/{pf[1]:.2e}/{pf[2]:.2e}")
        assert pg[1] <= max(3.0 * pf[1], 1e-5), (name, pg, pf)           # as accurate as the fp32 restatement
        assert pg[2] <= max(5.0 * pf[2], 1e-4), (name, pg, pf)
    print(task, "contact-force err / max force:", cf_err, "contacts seen", total_contacts, "last stats", st_o[:4])
    if task in ("go1seesaw", "go1pushbox", "go1revolvingdoor"):
        assert st_o[2] > 0, "no robot-on-box contacts were exercised"
    assert cf_err < 3e-2, cf_err
    eng.close()


@pytest.mark.parametrize("mode", [E.POLICY_FP32, E.POLICY_BF16X3], ids=["fp32", "tcgen05-bf16x3"])
@pytest.mark.parametrize("task", ["go1gate", "go1sheep-hard", "go1football-defender", "go1seesaw", "go1pushbox", "go1revolvingdoor", "go1wrestling", "go1bridge", "go1tug"])
def test_short_trajectory_parity(task, mode):
    """Full Go1.step() x 5 from reset on identical seeds and actions (tensor-core mode: steps 3.. are CUDA-graph replays)."""
    sc, eng, orc = make_pair(task, 32, mode)
    eng.reset(); orc.reset()
    worst = {}
    for s in range(5):
        act = actions_for(sc, s)
        eng.step(dev(act).data_ptr()); orc.step(act)
        torch.cuda.synchronize()
        r_g, r_o = get(eng, E.BUF_ROOT_STATES).reshape(-1, 13), orc.get(E.BUF_ROOT_STATES).reshape(-1, 13)
        d_g, d_o = get(eng, E.BUF_DOF_STATES).reshape(-1, 2), orc.get(E.BUF_DOF_STATES).reshape(-1, 2)
        worst[s] = (np.abs(r_g[:, :7] - r_o[:, :7]).max(), np.abs(r_g[:, 7:] - r_o[:, 7:]).max(),
                    np.abs(d_g[:, 0] - d_o[:, 0]).max(), np.abs(d_g[:, 1] - d_o[:, 1]).max())
        for buf in (E.BUF_EPISODE_LENGTH, E.BUF_TIMEOUT, E.BUF_RESET, E.BUF_COLLIDE, E.BUF_ROLL_TERM, E.BUF_PITCH_TERM,
                    E.BUF_ZLOW_TERM, E.BUF_ZHIGH_TERM):
            assert np.array_equal(get(eng, buf), orc.get(buf)), (task, s, buf)          # bookkeeping: bit-exact
        ob_g, ob_o = get(eng, E.BUF_OBS), orc.obs()
        rpy = slice(*E.OBS_SLICES["base_rpy"])
        d_rpy = np.abs(((ob_g[:, rpy] - ob_o[:, rpy] + np.pi) % (2 * np.pi)) - np.pi)   # angles live on the circle
        ob_g[:, rpy] = ob_o[:, rpy]
        assert d_rpy.max() < 1e-4 and np.allclose(ob_g, ob_o, rtol=1e-4, atol=2e-4), (task, s, "obs rows")
        assert np.allclose(get(eng, E.BUF_LOC_OBS), orc.get(E.BUF_LOC_OBS).reshape(-1, 70), rtol=1e-4, atol=2e-4)
        assert np.allclose(get(eng, E.BUF_CLOCK).ravel(), orc.get(E.BUF_CLOCK), atol=1e-5)
    print(task, {k: tuple(f"{x:.2e}" for x in v) for k, v in worst.items()})
    # measured on B200: ~1e-7 (pos, q), ~2e-6 (root vel), ~1e-5 (qd) after the first policy step, ~1e-4 (qd) after five
    assert worst[0][0] < 5e-6 and worst[0][2] < 5e-6, worst[0]
    assert worst[0][1] < 1e-4 and worst[0][3] < 1e-3, worst[0]
    assert worst[4][0] < 1e-4 and worst[4][2] < 1e-4, worst[4]          # still on the same trajectory after 5 policy steps
    act_g, act_o = get(eng, E.BUF_ACTIONS).ravel(), orc.get(E.BUF_ACTIONS)
    assert np.isfinite(act_g).all()
    eng.close()


def test_history_ring_equals_shift_concat():
    """The ring + age-rotated weights reproduce cat(history[:, 70:], obs) (go1.py:102): compare the logical history."""
    sc, eng, orc = make_pair("go1gate", 16)
    eng.reset(); orc.reset()
    for s in range(33):                               # > 30 so the ring wraps
        act = actions_for(sc, s)
        eng.policy(dev(act).data_ptr()); orc.policy(act)
        # keep both on the oracle's trajectory: copy actions so the frames stay comparable
    torch.cuda.synchronize()
    h_g = eng.history().cpu().numpy()
    h_o = orc.get(E.BUF_HISTORY).reshape(h_g.shape)
    # frames depend on previous network outputs (last actions) -> small fp32 differences accumulate
    assert np.allclose(h_g[:, -70:-28], h_o[:, -70:-28], atol=1e-5)   # newest frame: gravity/commands/dof parts identical
    assert np.allclose(h_g, h_o, rtol=1e-3, atol=2e-3)
    eng.close()


def test_episode_bookkeeping_bit_exact():
    """time-outs, episode counters and reset masks over > one episode (max_episode_length shortened to 20)."""
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 64; cfg.env.episode_length_s = 0.4
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0, policy_mode=E.POLICY_FP32, wrapper_action_scale=(2.0, 0.5, 0.5))
    eng = E.Engine(sc.desc, device=0, keepalive=sc)
    eng.reset()
    zero = torch.zeros((64, 2, 3), device="cuda:0")
    ep_prev = np.zeros(64, dtype=np.int64)
    n_timeouts = 0
    for s in range(50):
        eng.step(zero.data_ptr())
        torch.cuda.synchronize()
        ep, to, rs = get(eng, E.BUF_EPISODE_LENGTH), get(eng, E.BUF_TIMEOUT), get(eng, E.BUF_RESET)
        expect_to = (ep_prev + 1) > sc.desc.max_episode_length
        assert np.array_equal(to.astype(bool), expect_to)
        assert (rs[to.astype(bool)] == 1).all()
        assert np.array_equal(ep, np.where(rs.astype(bool), 0, ep_prev + 1))
        n_timeouts += int(to.sum())
        ep_prev = ep
    assert n_timeouts >= 64
    eng.close()


@pytest.mark.parametrize("task,n", [("go1gate", 4096), ("go1sheep-hard", 1024), ("go1football-defender", 1024), ("go1seesaw", 1024)])
def test_full_size_invariants(task, n):
    """BASELINE sizes: size-independent properties -- finite state, unit quaternions, joint positions inside the URDF
    limits (+margin), bounded heights, feet not below the floor, and walking under the frozen policy."""
    sc, eng, orc = make_pair(task, n)
    orc.close()
    eng.reset()
    a_ctrl = sc.num_agents - 1 if sc.desc.defender else sc.num_agents
    act = torch.zeros((n, a_ctrl, 3), device="cuda:0"); act[..., 0] = 0.5       # 1 m/s forward
    x0 = eng.tensor(E.BUF_ROOT_STATES)[:, 0, 0].clone()
    for s in range(40):
        eng.step(act.data_ptr())
    torch.cuda.synchronize()
    root = get(eng, E.BUF_ROOT_STATES)[:, :sc.num_agents]
    dofs = get(eng, E.BUF_DOF_STATES)[:, :12 * sc.num_agents]
    assert np.isfinite(root).all() and np.isfinite(dofs).all()
    qn = np.linalg.norm(root[..., 3:7], axis=-1)
    assert np.abs(qn - 1).max() < 1e-4
    m = sc.model
    q = dofs[..., 0].reshape(n, sc.num_agents, 12)
    assert (q > m.q_lower - 0.1).all() and (q < m.q_upper + 0.1).all()
    z = root[..., 2]
    assert (z > 0.0).all() and (z < 1.6).all()
    ep = get(eng, E.BUF_EPISODE_LENGTH)
    alive = ep == 40
    frac_alive = alive.mean()
    dx = (get(eng, E.BUF_ROOT_STATES)[:, 0, 0] - x0.cpu().numpy())[alive]
    print(task, "alive fraction", frac_alive, "median forward progress of agent 0 in 0.8 s", np.median(dx) if dx.size else None)
    assert frac_alive > 0.5
    if task != "go1football-defender":
        assert np.median(dx) > 0.3                       # commanded 1 m/s for 0.8 s (includes the landing transient)
    eng.close()


def test_env_surface_go1gate():
    """make_mqe_env / reset / step through the reference-facing VecEnv surface."""
    from types import SimpleNamespace
    from mqe_b200.envs import make_mqe_env, custom_cfg
    args = SimpleNamespace(num_envs=8, seed=0, headless=True, record_video=False, sim_device="cuda:0")
    env, cfg = make_mqe_env("go1gate", args, custom_cfg(args))
    assert env.num_envs == 8 and env.num_agents == 2 and cfg.env.num_envs == 8
    assert env.reset() == 0                               # the shipped gate wrapper returns 0 (go1_gate_wrapper.py:68)
    a = torch.zeros((8, 2, 3), device="cuda:0")
    obs, rew, done, info = env.step(a)
    assert obs == 0 and rew == 0 and done.shape == (8,) and done.dtype == torch.bool
    ob = env.obs_buf
    assert ob.base_pos.shape == (16, 3) and ob.base_rpy.shape == (16, 3) and ob.dof_pos.shape == (16, 12)
    assert ob.env_info["gate_deviation"].shape == (8, 2)
    assert env.root_states.shape == (16, 13) and env.all_root_states.shape == (16, 13)
    env.close()


@pytest.mark.parametrize("task,D,A", [("go1sheep-hard", 34, 2), ("go1sheep-easy", 18, 2), ("go1seesaw", 14, 2), ("go1football-defender", 20, 2),
                                      ("go1pushbox", 22, 2), ("go1revolvingdoor", 12, 2), ("go1wrestling", 12, 2), ("go1bridge", 12, 2), ("go1tug", 10, 2)])
def test_env_surface_wrappers(task, D, A):
    from types import SimpleNamespace
    from mqe_b200.envs import make_mqe_env, custom_cfg
    args = SimpleNamespace(num_envs=16, seed=0, headless=True, record_video=False, sim_device="cuda:0")
    env, cfg = make_mqe_env(task, args, custom_cfg(args))
    obs = env.reset()
    assert obs.shape == (16, A, D), obs.shape
    assert env.observation_space.shape == (D,)
    for s in range(3):
        a = torch.rand((16, A, 3), device="cuda:0") * 2 - 1
        obs, rew, done, info = env.step(a)
    assert obs.shape == (16, A, D) and rew.shape[:2] == (16, A) and done.shape == (16,)      # revolving door: [N, A, 1]
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert float(env.reward_buffer["step count"]) == 3
    env.close()


def test_openrl_adapter_numpy_surface():
    """openrl_ws/utils.py surface: numpy in / numpy out, 0.5 action pre-scale, dones repeated per agent, batch_rewards."""
    from types import SimpleNamespace
    from mqe_b200.openrl_adapter import make_env
    args = SimpleNamespace(task="go1sheep-easy", num_envs=8, seed=0, headless=True, record_video=False, sim_device="cuda:0")
    env, cfg = make_env(args)
    assert env.agent_num == 2 and env.parallel_env_num == 8
    obs = env.reset()
    assert isinstance(obs, np.ndarray) and obs.shape == (8, 2, 18)
    for _ in range(3):
        obs, rew, done, infos = env.step(np.random.uniform(-1, 1, size=(8, 2, 3)).astype(np.float32))
    assert obs.shape == (8, 2, 18) and rew.shape == (8, 2, 1) and done.shape == (8, 2) and done.dtype == bool and len(infos) == 8
    r = env.batch_rewards(None)
    assert "average step reward" in r and float(env.env.reward_buffer["step count"]) == 0
    env.close()
    args.task = "go1seesaw"
    env, cfg = make_env(args, single_agent=True)
    obs = env.reset()
    assert obs.shape == (16, 1, 14)
    obs, rew, done, infos = env.step(np.zeros((16, 1, 3), dtype=np.float32))
    assert obs.shape == (16, 1, 14) and rew.shape == (16, 1, 1) and done.shape == (16, 1)
    env.close()


@pytest.mark.parametrize("task,A", [("go1football-1vs1", 2), ("go1football-2vs2", 4)])
def test_football_game_tasks(task, A):
    """go1football-1vs1 / -2vs2: free robots + ball; the reference wrapper returns None observations and zero rewards."""
    from types import SimpleNamespace
    from mqe_b200.envs import make_mqe_env, custom_cfg
    args = SimpleNamespace(num_envs=32, seed=0, headless=True, record_video=False, sim_device="cuda:0")
    env, cfg = make_mqe_env(task, args, custom_cfg(args))
    assert env.reset() is None and env.num_agents == A
    x0 = env.root_states[:, 0].clone().view(32, A)
    for s in range(30):
        a = torch.zeros((32, A, 3), device="cuda:0"); a[..., 0] = 0.5
        obs, rew, done, info = env.step(a)
    assert obs is None and rew.shape == (32, 4) and float(rew.abs().sum()) == 0.0 and done.shape == (32,)
    rs = env.root_states
    assert torch.isfinite(rs).all() and torch.isfinite(env.root_states_npc).all()
    alive = env.episode_length_buf == 30 if task.endswith("2vs2") else torch.ones(32, dtype=torch.bool, device="cuda:0")
    dx = (rs[:, 0].view(32, A) - x0)[alive]
    # first team walks +x, the mirrored team (yaw = pi) walks -x; 0.6 s includes the landing transient
    assert (dx[:, 0] > 0).float().mean() > 0.9 and dx[:, 0].median() > 0.1 and (dx[:, A - 1] < 0).float().mean() > 0.9 and dx[:, A - 1].median() < -0.1
    assert env.ball_pos.shape == (32, 2, 3)
    env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["go1sheep-hard", "go1sheep-easy", "go1seesaw", "go1football-defender", "go1pushbox", "go1wrestling", "go1bridge",
                                  "go1revolvingdoor"])
def test_fused_wrapper_gather_matches_torch_wrapper(task, monkeypatch):
    """csrc/wrapper.cu (one kernel inside the step graph) against the torch wrappers (pinned to the reference's own code by
    tests/test_wrappers_golden.py): two envs built from the same seed, same actions, resets included (short episodes)."""
    from types import SimpleNamespace
    from mqe_b200.envs import make_mqe_env

    def build(fused):
        monkeypatch.setenv("MQE_FUSED_WRAPPERS", "1" if fused else "0")
        args = SimpleNamespace(num_envs=33, seed=3, headless=True, record_video=False, sim_device="cuda:0")

        def cc(cfg):
            cfg.env.num_envs = 33
            cfg.env.episode_length_s = 0.3                        # 15 policy steps: time-outs exercise the reset-dependent terms
            return cfg
        return make_mqe_env(task, args, cc)[0]

    ef, et = build(True), build(False)
    of, ot = ef.reset(), et.reset()
    assert getattr(ef, "_fused", False) and not getattr(et, "_fused", False)
    assert torch.allclose(of, ot, rtol=1e-6, atol=1e-6)
    A = of.shape[1]
    rng = np.random.default_rng(0)
    for s in range(40):
        a = torch.as_tensor(rng.uniform(-1.2, 1.2, size=(33, A, 3)).astype(np.float32), device="cuda:0")
        of, rf, df, _ = ef.step(a.clone())
        ot, rt, dt_, _ = et.step(a.clone())
        assert torch.equal(df, dt_), s
        assert torch.allclose(of, ot, rtol=1e-6, atol=1e-6), (s, (of - ot).abs().max())
        assert rf.shape == rt.shape and torch.allclose(rf, rt, rtol=1e-5, atol=1e-5), (s, (rf - rt).abs().max())
    for k in et.reward_buffer:
        vt, vf = float(et.reward_buffer[k]), float(ef.reward_buffer[k])
        assert abs(vt - vf) <= 1e-4 * max(1.0, abs(vt)), (k, vt, vf)
    ef.reward_buffer["step count"] = 0                            # loggers zero the buffer: assignment re-bases the running sum
    ef.step(a.clone())
    assert float(ef.reward_buffer["step count"]) == 1
    ef.close(); et.close()


@pytest.mark.gpu
def test_domain_rand_push_and_friction_parity():
    """push_robots + randomize_friction + randomize_base_mass switched on: kernels against the oracle over the first pushes."""
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 16
    cfg.domain_rand.push_robots = True; cfg.domain_rand.push_interval_s = 0.06; cfg.domain_rand.max_push_vel_xy = 0.8
    cfg.domain_rand.randomize_friction = True; cfg.domain_rand.friction_range = [0.05, 1.5]
    cfg.domain_rand.randomize_base_mass = True; cfg.domain_rand.added_mass_range = [-1.0, 3.0]
    cfg.domain_rand.randomize_com = True                                  # legged_robot_field.py:321-332 (com_range of go1_config.py:218-221)
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0, policy_mode=E.POLICY_FP32, wrapper_action_scale=(2.0, 0.5, 0.5))
    assert sc.desc.h_base_com_shift and sc.desc.h_base_added_mass and sc.desc.h_env_friction
    eng, orc = E.Engine(sc.desc, device=0, keepalive=sc), oracle.Oracle(sc, "f32")
    eng.reset(); orc.reset()
    for s in range(7):
        a = actions_for(sc, s)
        eng.step(dev(a).data_ptr()); orc.step(a)
        g, r = get(eng, E.BUF_ROOT_STATES).reshape(16, 2, 13), orc.root_states()
        if (s + 1) % 3 == 0:                                     # push step: the redrawn velocities are the same draws on both sides
            assert np.allclose(g[..., 7:9], r[..., 7:9], atol=1e-6) and np.all(np.abs(g[..., 7:9]) <= 0.8 + 1e-6)
        assert np.allclose(g[..., :7], r[..., :7], atol=2e-4), (s, np.abs(g[..., :7] - r[..., :7]).max())
        assert np.allclose(g[..., 7:], r[..., 7:], atol=5e-3), (s, np.abs(g[..., 7:] - r[..., 7:]).max())
    eng.close(); orc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 5, 131])
def test_ragged_env_counts_match_oracle(n):
    """Env counts that do not fill a warp of 4 envs, a CTA of 7 warps or a wave of 148 CTAs: the last warp / CTA is partly empty."""
    sc, eng, orc = make_pair("go1gate", n)
    eng.reset(); orc.reset()
    for s in range(3):
        a = actions_for(sc, s)
        eng.step(dev(a).data_ptr()); orc.step(a)
    g, r = get(eng, E.BUF_ROOT_STATES).reshape(n, 2, 13), orc.root_states()
    assert np.allclose(g[..., :7], r[..., :7], atol=1e-4) and np.allclose(g[..., 7:], r[..., 7:], atol=2e-3)
    assert np.array_equal(get(eng, E.BUF_RESET), orc.get(E.BUF_RESET)) and np.array_equal(get(eng, E.BUF_EPISODE_LENGTH), orc.get(E.BUF_EPISODE_LENGTH))
    eng.close(); orc.close()


@pytest.mark.gpu
def test_error_paths_return_codes_not_crashes():
    """The C ABI reports bad descriptors / arguments through MqeStatus + mqe_last_error (include/mqe_b200.h), it never throws."""
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 4
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0)
    for field, bad, msg in (("num_agents", 5, "out of range"), ("num_envs", 0, "out of range"), ("abi_version", 99, "abi_version"), ("defender", 1, "defender")):
        keep = getattr(sc.desc, field)
        setattr(sc.desc, field, bad)
        with pytest.raises(E.EngineError) as ei:
            E.Engine(sc.desc, device=0, keepalive=sc)
        assert msg in str(ei.value), (field, str(ei.value))
        setattr(sc.desc, field, keep)
    with pytest.raises(E.EngineError):
        E.Engine(sc.desc, device=99, keepalive=sc)
    eng = E.Engine(sc.desc, device=0, keepalive=sc)
    lib = eng.lib
    assert lib.mqe_sim_step(eng.h, None) < 0 and b"null" in lib.mqe_last_error()
    assert lib.mqe_sim_substeps(eng.h, 0) < 0
    assert lib.mqe_sim_get_buffer(eng.h, 999, None, None, None) < 0
    assert lib.mqe_sim_get_buffer(eng.h, E.BUF_WRAP_SUMS, None, None, None) < 0       # no task wrapper set yet
    assert lib.mqe_sim_get_buffer(eng.h, E.BUF_HISTORY, None, None, None) < 0         # tensor-core policy mode: the bf16 planes are the ring
    assert lib.mqe_sim_get_buffer(eng.h, E.BUF_HISTORY_HI, None, None, None) == 0
    assert lib.mqe_sim_step_joint(eng.h, dev(actions_for(sc, 0)).data_ptr()) == -4 and b"control_type" in lib.mqe_last_error()   # 'C' takes commands
    assert lib.mqe_sim_gather_view(eng.h, 0, None, None) < 0 and lib.mqe_sim_gather_parity(eng.h) == -1   # no peer exchange connected
    assert lib.mqe_sim_wrapper_reset(eng.h) < 0
    with pytest.raises(E.EngineError):
        eng.set_wrapper(E.WRAP_SHEEP, [1, 0, 0, 0, 0, 0])                              # no sheep in go1gate
    with pytest.raises(E.EngineError):
        eng.set_wrapper(77, [0])
    buf = np.zeros(16, dtype=np.float32)
    assert lib.mqe_sim_unpin_host(eng.h, buf.ctypes.data_as(__import__("ctypes").c_void_p)) < 0   # never pinned
    eng.pin_host(buf); eng.pin_host(buf)                                               # idempotent
    eng.reset(); eng.step(dev(actions_for(sc, 0)).data_ptr()); eng.synchronize()       # the handle is still healthy
    assert np.isfinite(get(eng, E.BUF_ROOT_STATES)).all()
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ct", ["P", "T", "V"])
def test_joint_action_control_types_parity(ct):
    """cfg.control.control_type 'P' / 'T' / 'V' (legged_robot.py:384-392): Go1.step bypasses the walk policy (go1.py:43-45) and takes
    [N, 12A] joint actions, clipped to +-clip_actions by pre_physics_step.  Kernels (mqe_sim_step_joint) against the oracle, and the
    reference-facing Go1.step() on top of it; a 3-D command caller is refused instead of silently getting another controller."""
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 8
    cfg.control.control_type = ct
    cfg.control.stiffness = {"joint": 40.0 if ct == "P" else 2.0}; cfg.control.damping = {"joint": 1.0 if ct == "P" else 0.002}
    cfg.normalization.clip_actions = 1.5
    cfg.domain_rand.randomize_motor = True; cfg.domain_rand.leg_motor_strength_range = [0.7, 1.3]     # legged_robot_field.py:283-291, 180-183
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0, policy_mode=E.POLICY_FP32, wrapper_action_scale=(2.0, 0.5, 0.5))
    assert sc.desc.h_motor_strength
    eng, orc = E.Engine(sc.desc, device=0, keepalive=sc), oracle.Oracle(sc, "f32")
    eng.reset(); orc.reset()
    assert eng.lib.mqe_sim_step(eng.h, dev(actions_for(sc, 0)).data_ptr()) == -4           # MQE_ERR_UNSUPPORTED: commands need control_type 'C'
    rng = np.random.default_rng(5)
    for s in range(4):
        a = rng.uniform(-2.0, 2.0, size=(8, 24)).astype(np.float32)                       # beyond the clip on purpose
        eng.step_joint(dev(a).data_ptr()); orc.step_joint(a)
        torch.cuda.synchronize()
        assert np.allclose(get(eng, E.BUF_ACTIONS).ravel(), np.clip(a, -1.5, 1.5).ravel())
        g, r = get(eng, E.BUF_ROOT_STATES).reshape(8, 2, 13), orc.root_states()
        assert np.allclose(g[..., :7], r[..., :7], atol=2e-4), (s, np.abs(g[..., :7] - r[..., :7]).max())
        assert np.allclose(get(eng, E.BUF_TORQUES).ravel(), orc.get(E.BUF_TORQUES).ravel(), atol=2e-2), s
        assert np.array_equal(get(eng, E.BUF_RESET), orc.get(E.BUF_RESET))
        assert np.abs(get(eng, E.BUF_CLOCK)).max() == 0.0                                   # no gait clock outside control_type 'C' (go1.py:241)
    eng.close(); orc.close()
    from mqe_b200.envs.go1 import Go1
    env = Go1(cfg, sim_device="cuda:0", seed=0, policy_mode=E.POLICY_FP32)
    env.reset()
    obs, rew, done, _ = env.step(torch.zeros(8 * 2, 12, device="cuda:0"))                 # the reference reshapes to [N, -1] (go1.py:44)
    assert done.shape == (8,) and obs.dof_pos.shape == (16, 12) and torch.isfinite(env.root_states).all()
    with pytest.raises(AssertionError):
        env.step(torch.zeros(8, 2, 3, device="cuda:0"))
    env.close()


@pytest.mark.gpu
def test_action_lag_parity():
    """domain_rand.randomize_lag_timesteps (go1.py:337-339): the lag ring in k_substeps against the oracle, across policy steps (the ring
    position is a device-side counter of torque evaluations, so CUDA-graph replays keep advancing it)."""
    for mode in (E.POLICY_FP32, E.POLICY_BF16X3):
        cfg = C.Go1GateCfg(); cfg.env.num_envs = 8
        cfg.domain_rand.randomize_lag_timesteps = True; cfg.domain_rand.lag_timesteps = 6
        np.random.seed(0)
        sc = S.build_scene(cfg, seed=0, policy_mode=mode, wrapper_action_scale=(2.0, 0.5, 0.5))
        eng, orc = E.Engine(sc.desc, device=0, keepalive=sc), oracle.Oracle(sc, "f32")
        eng.reset(); orc.reset()
        for s in range(6):
            a = actions_for(sc, s)
            eng.step(dev(a).data_ptr()); orc.step(a)
            g, r = get(eng, E.BUF_ROOT_STATES).reshape(8, 2, 13), orc.root_states()
            assert np.allclose(g[..., :7], r[..., :7], atol=3e-4), (mode, s, np.abs(g[..., :7] - r[..., :7]).max())
            assert np.allclose(get(eng, E.BUF_TORQUES).ravel(), orc.get(E.BUF_TORQUES).ravel(), atol=5e-2), (mode, s)
        eng.close(); orc.close()
