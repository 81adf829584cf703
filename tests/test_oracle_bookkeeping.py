"""The oracle's per-step bookkeeping against golden vectors produced by the REFERENCE's own methods
(tools/gen_bookkeeping_golden.py ran Go1.preprocess_action / _compute_torques / _step_contact_targets /
compute_observations, LeggedRobotField.check_termination, Go1FootballDefender._get_defender_action and
Go1Sheep._step_npc unmodified on a constructor-less instance with `isaacgym` stubbed).  This pins the part of the
oracle that restates reference Python -- the GPU tests then pin the kernels to the oracle."""
import os

import numpy as np
import pytest

import oracle
from mqe_b200 import engine as E
from mqe_b200 import scene as S
from mqe_b200.envs import configs as C

TASKS = {"go1gate": C.Go1GateCfg, "go1sheep-easy": C.SingleSheepCfg, "go1football-defender": C.Go1FootballDefenderCfg,
         "go1seesaw": C.Go1SeesawCfg, "go1tug": C.Go1TugCfg, "go1wrestling": C.Go1WrestlingCfg, "go1bridge": C.Go1BridgeCfg}
O = E.OBS_SLICES


@pytest.mark.parametrize("task", sorted(TASKS))
def test_bookkeeping_matches_reference_methods(task, golden_dir):
    z = np.load(os.path.join(golden_dir, f"bookkeeping_{task}.npz"))
    cfg = TASKS[task]()
    N = z["env_origins"].shape[0]
    cfg.env.num_envs = N
    np.random.seed(0)
    sc = S.build_scene(cfg, seed=0, wrapper_action_scale=(1.0, 1.0, 1.0))      # actions arrive already wrapper-scaled
    A, P = sc.num_agents, sc.num_npcs
    sc.env_origins[:] = z["env_origins"]                 # the arrays the descriptor points at
    sc.agent_origins[:] = z["agent_origins"]
    assert sc.desc.max_episode_length == int(z["max_episode_length"])
    o = oracle.Oracle(sc, "f64")
    # LeggedRobot._init_buffers (legged_robot.py:567-570, 620-622), the reference's own lines run on the spawn state: base_quat = spawn
    # quaternion, projected_gravity = R^T (0, 0, -1), base velocities from the spawn velocities.  reset_idx does not recompute them, so
    # the observation of the very first reset() must carry them (they used to start as zeros).
    M0 = N * sc.num_agents
    assert np.allclose(o.get(E.BUF_PROJ_GRAVITY).reshape(M0, 3), z["init_proj_grav"], atol=1e-6)
    assert np.allclose(o.get(E.BUF_BASE_LIN_VEL).reshape(M0, 3), z["init_base_lin_vel"], atol=1e-6)
    assert np.allclose(o.get(E.BUF_BASE_ANG_VEL).reshape(M0, 3), z["init_base_ang_vel"], atol=1e-6)
    o.reset()
    first = o.obs()
    assert np.allclose(first[:, O["projected_gravity"][0]:O["projected_gravity"][1]], z["init_proj_grav"], atol=1e-6), "first reset(): gravity"
    assert np.allclose(first[:, O["lin_vel"][0]:O["lin_vel"][1]], 2.0 * z["init_base_lin_vel"], atol=1e-6)      # pre-reset values, not the redrawn root velocity
    if sc.num_npcs:                                      # P > 0: base_quat is a stale copy of the spawn quaternion until the first physics step
        assert np.allclose(first[:, O["base_quat"][0]:O["base_quat"][1]], z["init_base_quat"], atol=1e-6)
    o.set_reset_state(False)                             # reset_idx without its RNG part, as in the golden run
    o.set(E.BUF_HISTORY, np.zeros((N * A, 2100), dtype=np.float32))
    o.set(E.BUF_ROOT_STATES, z["root0"]); o.set(E.BUF_DOF_STATES, z["dof0"])
    o.set(E.BUF_EPISODE_LENGTH, z["episode_length0"].astype(np.float32))
    o.observe()
    M = N * A
    a_ctrl = A - 1 if sc.desc.defender else A
    close = lambda a, b, tol=2e-5: np.allclose(np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel(), rtol=1e-5, atol=tol)
    for t in range(z["in_actions"].shape[0]):
        o.policy(z["in_actions"][t].reshape(N, a_ctrl, 3))
        if sc.desc.defender:
            cmd = o.get(E.BUF_COMMANDS).reshape(N, A, 3)[:, 2]
            assert close(cmd, z["defender_cmd"][t], 1e-5), (task, t, "defender command")
        assert close(o.get(E.BUF_LOC_OBS), z["loc_obs"][t]), (task, t, "locomotion_obs")
        assert close(o.get(E.BUF_HISTORY), z["history"][t]), (task, t, "history")
        assert close(o.get(E.BUF_ACTIONS), z["actions"][t], 2e-4), (task, t, "actions", np.abs(o.get(E.BUF_ACTIONS) - z["actions"][t].ravel()).max())
        o.torques()
        assert close(o.get(E.BUF_TORQUES), z["torques1"][t], 2e-4), (task, t, "torques 1")
        o.set(E.BUF_DOF_STATES, z["in_dof_mid"][t])
        o.torques()
        assert close(o.get(E.BUF_TORQUES), z["torques2"][t], 2e-4), (task, t, "torques 2")
        o.set(E.BUF_ROOT_STATES, z["in_root"][t]); o.set(E.BUF_DOF_STATES, z["in_dof"][t]); o.set(E.BUF_CONTACT_FORCES, z["in_contact"][t])
        o.post_physics()
        # flags and counters: bit-exact
        assert np.array_equal(o.get(E.BUF_RESET).astype(bool), z["reset"][t]), (task, t, "reset_buf")
        assert np.array_equal(o.get(E.BUF_TIMEOUT).astype(bool), z["timeout"][t]), (task, t)
        if sc.desc.term_mask & 16:
            assert np.array_equal(o.get(E.BUF_COLLIDE).astype(bool), z["collide"][t]), (task, t)
        if sc.desc.term_mask & 1:
            assert np.array_equal(o.get(E.BUF_ROLL_TERM).astype(bool), z["r_term"][t])
        if sc.desc.term_mask & 2:
            assert np.array_equal(o.get(E.BUF_PITCH_TERM).astype(bool), z["p_term"][t])
        assert np.array_equal(o.get(E.BUF_EPISODE_LENGTH), z["ep_len"][t]), (task, t, "episode_length_buf")
        # derived quantities and the observation struct
        assert close(o.get(E.BUF_BASE_LIN_VEL), z["base_lin_vel"][t]) and close(o.get(E.BUF_BASE_ANG_VEL), z["base_ang_vel"][t])
        assert close(o.get(E.BUF_PROJ_GRAVITY), z["proj_grav"][t])
        assert close(o.get(E.BUF_GAIT), z["gait"][t]) and close(o.get(E.BUF_CLOCK), z["clock"][t], 1e-5), (task, t, "gait clock")
        obs = o.obs()
        for name, key in (("base_pos", "obs_base_pos"), ("base_quat", "obs_base_quat"), ("dof_pos", "obs_dof_pos"), ("dof_vel", "obs_dof_vel"),
                          ("lin_vel", "obs_lin_vel"), ("ang_vel", "obs_ang_vel"), ("last_action", "obs_last_action"),
                          ("last_last_action", "obs_last_last_action"), ("projected_gravity", "obs_proj_grav"),
                          ("clock_inputs", "obs_clock"), ("base_rpy", "obs_rpy")):
            a, b = O[name]
            ref = z[key][t].reshape(M, -1)
            got = obs[:, a:b]
            if name == "base_rpy":                         # angles are returned modulo 2 pi: compare on the circle
                d = np.abs(((got - ref + np.pi) % (2 * np.pi)) - np.pi)
                assert d.max() < 2e-5, (task, t, name, d.max())
            else:
                assert close(got, ref, 2e-4 if "action" in name else 2e-5), (task, t, name, np.abs(got - ref).max())
        if z["sheep_root_after"][t].size:                  # Go1Sheep._step_npc with randomness 0
            got = o.root_states()
            assert close(got[:, A:], z["sheep_root_after"][t][:, A:], 1e-5), (task, t, "sheep step")
    o.close()
