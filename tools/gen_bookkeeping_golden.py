#!/usr/bin/env python
"""Golden vectors for the per-step bookkeeping of Go1.step(), produced by the REFERENCE's own methods.

Runs in this container only.  `isaacgym` is stubbed (torch_utils restated from SURVEY.md appendix B) and the methods
    Go1._prepare_locomotion_policy / preprocess_action / _compute_torques / _step_contact_targets / compute_observations
    LeggedRobotField.check_termination (-> LeggedRobot.check_termination), the derived-quantity lines of post_physics_step,
    Go1FootballDefender._get_defender_action, Go1Sheep._step_npc (randomness 0)
are called UNMODIFIED on a fake `self` holding random state tensors (CPU).  Inputs and outputs go to
tests/golden/bookkeeping_<task>.npz; tests/test_oracle_bookkeeping.py replays them through the C oracle.
"""
import inspect
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_terrain_golden import REF, REPO, install_stubs  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def install_torch_utils():
    tu = sys.modules["isaacgym.torch_utils"]

    def quat_rotate_inverse(q, v):
        q_w = q[:, -1]
        q_vec = q[:, :3]
        a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
        b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
        c = q_vec * torch.bmm(q_vec.view(q.shape[0], 1, 3), v.view(q.shape[0], 3, 1)).squeeze(-1) * 2.0
        return a - b + c

    def get_euler_xyz(q):
        qx, qy, qz, qw = 0, 1, 2, 3
        sinr_cosp = 2.0 * (q[:, qw] * q[:, qx] + q[:, qy] * q[:, qz])
        cosr_cosp = q[:, qw] * q[:, qw] - q[:, qx] * q[:, qx] - q[:, qy] * q[:, qy] + q[:, qz] * q[:, qz]
        roll = torch.atan2(sinr_cosp, cosr_cosp)
        sinp = 2.0 * (q[:, qw] * q[:, qy] - q[:, qz] * q[:, qx])
        pitch = torch.where(torch.abs(sinp) >= 1, torch.sign(sinp) * (np.pi / 2.0), torch.asin(sinp))
        siny_cosp = 2.0 * (q[:, qw] * q[:, qz] + q[:, qx] * q[:, qy])
        cosy_cosp = q[:, qw] * q[:, qw] + q[:, qx] * q[:, qx] - q[:, qy] * q[:, qy] - q[:, qz] * q[:, qz]
        yaw = torch.atan2(siny_cosp, cosy_cosp)
        return roll % (2 * np.pi), pitch % (2 * np.pi), yaw % (2 * np.pi)

    def to_torch(x, dtype=torch.float, device="cpu", requires_grad=False):
        return torch.tensor(x, dtype=dtype, device=device, requires_grad=requires_grad)

    def get_axis_params(value, axis_idx, x_value=0.0, dtype=float, n_dims=3):
        params = np.zeros((n_dims,))
        params[axis_idx] = value
        params[0] = x_value if axis_idx != 0 else params[0]
        return list(params.astype(dtype))

    def torch_rand_float(lower, upper, shape, device):
        return (upper - lower) * torch.rand(*shape, device=device) + lower

    fns = dict(quat_rotate_inverse=quat_rotate_inverse, get_euler_xyz=get_euler_xyz, to_torch=to_torch,
               get_axis_params=get_axis_params, torch_rand_float=torch_rand_float)
    tu.__dict__.update(fns)
    tu.__all__ = list(fns)


class Fake:
    """Stand-in for `self`: data attributes are set by the tool, methods resolve to the reference class (unmodified)."""
    _klass = None

    def __getattr__(self, name):
        klass = object.__getattribute__(self, "_klass")
        if klass is not None and hasattr(klass, name):
            attr = getattr(klass, name)
            if callable(attr):
                return types.MethodType(attr, self)
        raise AttributeError(name)


def quat_random(rng, n, tilt):
    ax = rng.normal(size=(n, 3)); ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = rng.uniform(-tilt, tilt, size=(n, 1))
    yaw = rng.uniform(-np.pi, np.pi, size=(n,))
    q1 = np.concatenate([ax * np.sin(ang / 2), np.cos(ang / 2)], axis=1)
    q2 = np.stack([0 * yaw, 0 * yaw, np.sin(yaw / 2), np.cos(yaw / 2)], axis=1)
    x1, y1, z1, w1 = q2.T; x2, y2, z2, w2 = q1.T
    q = np.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                  w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], axis=1)
    return q.astype(np.float32)


def run(task, cfg_mod, cfg_name, cls_mod, cls_name, N, seed):
    import importlib
    from copy import copy
    cfg = getattr(importlib.import_module(cfg_mod), cfg_name)
    Go1 = importlib.import_module("mqe.envs.go1.go1").Go1
    Field = importlib.import_module("mqe.envs.field.legged_robot_field").LeggedRobotField
    Cls = getattr(importlib.import_module(cls_mod), cls_name)
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    A, P = cfg.env.num_agents, getattr(cfg.env, "num_npcs", 0)
    M, NB = N * A, 17
    T = torch.as_tensor
    s = Cls.__new__(Cls)                 # a real instance of the reference class, constructor skipped (it needs Isaac Gym)
    s.cfg, s.device, s.sim_device = cfg, "cpu", "cpu"
    s.num_envs, s.num_agents, s.num_npcs, s.num_bodies, s.num_actuated_dof = N, A, P, NB, 12 * A
    s.dt = cfg.control.decimation * cfg.sim.dt
    s.max_episode_length = int(np.ceil(cfg.env.episode_length_s / s.dt))
    names = ["FL_hip_joint", "FL_thigh_joint", "FL_calf_joint", "FR_hip_joint", "FR_thigh_joint", "FR_calf_joint",
             "RL_hip_joint", "RL_thigh_joint", "RL_calf_joint", "RR_hip_joint", "RR_thigh_joint", "RR_calf_joint"]
    dflt = np.array([cfg.init_state.default_joint_angles[n] for n in names], dtype=np.float32)
    s.default_dof_pos = T(np.tile(dflt, A)).unsqueeze(0)
    s.torque_limits = T(np.tile(np.array(cfg.control.torque_limits, dtype=np.float32), A))
    s.obs_scales = cfg.normalization.obs_scales
    s.termination_contact_indices = T(np.array([0] if len(cfg.asset.terminate_after_contacts_on) else [], dtype=np.int64))
    s.obs_buf = copy(cfg.obs)
    if cls_name == "Go1Sheep":                                              # go1_sheep.py:31-33
        s.sheep_movement_scale, s.sheep_movement_randomness = cfg.asset.sheep_movement_scale, cfg.asset.sheep_movement_randomness
        s.sheep_movement_range = cfg.asset.sheep_movement_range
    # ---- networks exactly as the reference builds them
    os.chdir(REF)
    Go1._prepare_locomotion_policy(s)
    src = textwrap.dedent(inspect.getsource(Go1._init_buffers)).replace("super()._init_buffers()", "pass")
    src = src.replace("self.lag_buffer =", "self.lag_buffer = None #")
    ns = {"torch": torch}
    exec(src, ns)
    ns["_init_buffers"](s)
    s.last_locomotion_action = torch.zeros(M, 12)
    s.last_two_locomotion_action = torch.zeros(M, 12)
    s.gait_indices = torch.zeros(M); s.clock_inputs = torch.zeros(M, 4)
    s.doubletime_clock_inputs = torch.zeros(M, 4); s.halftime_clock_inputs = torch.zeros(M, 4)
    s.gravity_vec = T(np.tile(np.array([[0, 0, -1.0]], dtype=np.float32), (M, 1)))
    # ---- derived base quantities exactly as LeggedRobot._init_buffers leaves them (legged_robot.py:567-570, 620-622): the reference's OWN
    # assignment lines, executed on the spawn root state (cfg.init_state, quaternion normalised as PhysX does when the actor is created).
    # reset_idx does not recompute them, so the first reset()'s observation carries projected_gravity = R(spawn)^T (0, 0, -1).
    ist = cfg.init_state
    if getattr(ist, "multi_init_state", False):
        spawn = np.asarray([st.pos + st.rot + st.lin_vel + st.ang_vel for st in ist.init_states], dtype=np.float32)
    else:
        spawn = np.asarray([ist.pos + ist.rot + ist.lin_vel + ist.ang_vel] * A, dtype=np.float32)
    spawn[:, 3:7] /= np.linalg.norm(spawn[:, 3:7], axis=1, keepdims=True)
    s.root_states = T(np.tile(spawn, (N, 1)))
    LeggedRobot = importlib.import_module("mqe.envs.base.legged_robot").LeggedRobot
    init_src = inspect.getsource(LeggedRobot._init_buffers).splitlines()
    wanted = ("self.base_quat = ", "self.base_lin_vel = ", "self.base_ang_vel = ", "self.projected_gravity = ")
    picked = [ln.strip() for ln in init_src if ln.strip().startswith(wanted)]
    assert len(picked) == 4, picked
    tu0 = sys.modules["isaacgym.torch_utils"]
    exec("\n".join(picked), {"quat_rotate_inverse": tu0.quat_rotate_inverse, "self": s})
    s.base_quat = s.base_quat.clone()
    init_rec = {"init_spawn": spawn.copy(), "init_base_quat": s.base_quat.numpy().copy(), "init_base_lin_vel": s.base_lin_vel.numpy().copy(),
                "init_base_ang_vel": s.base_ang_vel.numpy().copy(), "init_proj_grav": s.projected_gravity.numpy().copy()}
    s.actions = torch.zeros(N, 12 * A); s.last_actions = torch.zeros(N, 12 * A)
    s.env_origins = T(rng.uniform(0, 30, size=(N, 3)).astype(np.float32)); s.env_origins[:, 2] = 0
    s.agent_origins = s.env_origins.unsqueeze(1).repeat(1, A, 1) + T(rng.uniform(-1, 1, size=(N, A, 3)).astype(np.float32))
    s.agent_origins[..., 2] = 0
    s.env_origins_repeat = s.env_origins.unsqueeze(1).repeat(1, A, 1).reshape(-1, 3)
    s.env_info = {"gate_deviation": torch.zeros(N, 2)}
    s.episode_length_buf = T(rng.integers(0, s.max_episode_length + 2, size=N).astype(np.int64))

    steps = 4
    rec = {**init_rec, "env_origins": s.env_origins.numpy().copy(), "agent_origins": s.agent_origins.numpy().copy(),
           "episode_length0": s.episode_length_buf.numpy().copy(), "max_episode_length": np.int64(s.max_episode_length)}
    out = {k: [] for k in ("in_actions", "in_root", "in_dof", "in_contact", "loc_obs", "history", "actions", "torques1", "torques2",
                           "base_lin_vel", "base_ang_vel", "proj_grav", "gait", "clock", "collide", "timeout", "r_term", "p_term", "reset",
                           "ep_len", "obs_base_pos", "obs_base_quat", "obs_dof_pos", "obs_dof_vel", "obs_lin_vel", "obs_ang_vel",
                           "obs_last_action", "obs_last_last_action", "obs_proj_grav", "obs_clock", "obs_rpy", "in_dof_mid",
                           "defender_cmd", "sheep_root_after")}
    # state before the first policy call = what Go1.reset() + compute_observations leaves: fill obs from a first state
    def set_state(root, dof, contact):
        s.all_root_states = T(root.reshape(-1, 13).copy())
        s.root_states = s.all_root_states.view(N, -1, 13)[:, :A, :].reshape(-1, 13)
        s.root_states_npc = s.all_root_states.view(N, -1, 13)[:, A:, :].reshape(-1, 13)
        s.base_pos = s.root_states[:, 0:3]
        d = T(dof.copy())
        s.dof_pos, s.dof_vel = d[:, :12 * A, 0].contiguous(), d[:, :12 * A, 1].contiguous()
        s.contact_forces = T(contact.copy())

    def rand_state():
        root = np.zeros((N, A + P, 13), dtype=np.float32)
        root[:, :, :3] = s.env_origins.numpy()[:, None, :] + rng.uniform(-2, 8, size=(N, A + P, 3))
        root[:, :A, 2] = rng.uniform(0.05, 0.6, size=(N, A)) if "z_low" in cfg.termination.termination_terms else rng.uniform(0.2, 0.5, size=(N, A))
        root[:, :, 3:7] = quat_random(rng, N * (A + P), 1.0).reshape(N, A + P, 4)
        root[:, :, 7:13] = rng.normal(size=(N, A + P, 6))
        dof = np.zeros((N, 12 * A + getattr(cfg.env, "num_actions_npc", 0), 2), dtype=np.float32)
        dof[:, :12 * A, 0] = np.tile(dflt, A)[None] + rng.uniform(-0.4, 0.4, size=(N, 12 * A))
        dof[:, :12 * A, 1] = rng.normal(size=(N, 12 * A)) * 3
        contact = (rng.normal(size=(N, NB * A + P, 3)) * (rng.random((N, NB * A + P, 1)) < 0.3) * 3).astype(np.float32)
        for a in range(A):                       # base body: contact in ~12 % of the cases only (it terminates the episode)
            contact[:, a * NB] *= (rng.random((N, 1)) < 0.12)
        return root, dof, contact

    root, dof, contact = rand_state()
    set_state(root, dof, contact)
    s.base_quat[:] = s.root_states[:, 3:7]
    Go1.compute_observations(s)
    rec["root0"], rec["dof0"] = root, dof
    a_ctrl = A - 1 if cls_name == "Go1FootballDefender" else A
    for t in range(steps):
        lim = 1.0 if cls_name == "Go1FootballDefender" else 1.3
        act = rng.uniform(-lim, lim, size=(N * a_ctrl, 3)).astype(np.float32)
        out["in_actions"].append(act)
        action = T(act) if cls_name == "Go1FootballDefender" else torch.clip(T(act), -1, 1)   # go1.py:38; the defender env does not clip
        if cls_name == "Go1FootballDefender":                               # go1_football_defender.py:25-54
            d_act = Cls._get_defender_action(s)
            out["defender_cmd"].append(d_act.numpy().copy())
            action = torch.cat([action.reshape(N, a_ctrl, 3), d_act.reshape(N, 1, 3)], dim=1).reshape(-1, 3)
        else:
            out["defender_cmd"].append(np.zeros(0, dtype=np.float32))
        la = Go1.preprocess_action(s, action)
        clip = cfg.normalization.clip_actions
        s.actions = torch.clip(la, -clip, clip).reshape(N, -1)
        out["loc_obs"].append(s.locomotion_obs.numpy().copy()); out["history"].append(s.history_locomotion_obs.numpy().copy())
        out["actions"].append(s.actions.numpy().copy())
        out["torques1"].append(Go1._compute_torques(s, s.actions).numpy().copy())
        root, dof_mid, contact = rand_state()                               # "physics": new random state between the calls
        set_state(root, dof_mid, contact)
        out["in_dof_mid"].append(dof_mid)
        out["torques2"].append(Go1._compute_torques(s, s.actions).numpy().copy())
        root, dof, contact = rand_state()
        set_state(root, dof, contact)
        out["in_root"].append(root); out["in_dof"].append(dof); out["in_contact"].append(contact)
        # ---- post_physics_step, legged_robot.py:126-149 without the gym refreshes / reset_idx
        s.episode_length_buf += 1
        s.base_quat[:] = s.root_states[:, 3:7]
        tu = sys.modules["isaacgym.torch_utils"]
        s.base_lin_vel[:] = tu.quat_rotate_inverse(s.base_quat, s.root_states[:, 7:10])
        s.base_ang_vel[:] = tu.quat_rotate_inverse(s.base_quat, s.root_states[:, 10:13])
        s.projected_gravity[:] = tu.quat_rotate_inverse(s.base_quat, s.gravity_vec)
        Go1._step_contact_targets(s)
        Field.check_termination(s)
        if cls_name == "Go1Sheep":
            s.npc_indices = (torch.arange(N, dtype=torch.int32).view(N, 1) * (A + P) + A + torch.arange(P, dtype=torch.int32).view(1, P)).reshape(-1)
            s.gym, s.sim = types.SimpleNamespace(set_actor_root_state_tensor_indexed=lambda *a: None), None
            sys.modules["isaacgym.gymtorch"].unwrap_tensor = lambda x: x
            Cls._step_npc(s)
            out["sheep_root_after"].append(s.all_root_states.view(N, -1, 13).numpy().copy())
        else:
            out["sheep_root_after"].append(np.zeros(0, dtype=np.float32))
        reset = (s.reset_buf if torch.is_tensor(s.reset_buf) else torch.zeros(N, dtype=torch.bool)).clone()
        # reset_idx without its torch-RNG part: _reset_buffers (go1.py:141-145 -> legged_robot.py:647-652)
        s.last_dof_vel = getattr(s, 'last_dof_vel', torch.zeros(N, 12 * A)); s.feet_air_time = torch.zeros(N, 4 * A)
        s.env_agent_indices = torch.arange(N * A).view(N, A)
        if torch.is_tensor(s.reset_buf):
            Go1._reset_buffers(s, reset.nonzero(as_tuple=False).flatten())
        Go1.compute_observations(s)
        for k, v in (("base_lin_vel", s.base_lin_vel), ("base_ang_vel", s.base_ang_vel), ("proj_grav", s.projected_gravity),
                     ("gait", s.gait_indices), ("clock", s.clock_inputs), ("timeout", s.time_out_buf), ("reset", reset),
                     ("ep_len", s.episode_length_buf), ("obs_base_pos", s.obs_buf.base_pos), ("obs_base_quat", s.obs_buf.base_quat),
                     ("obs_dof_pos", s.obs_buf.dof_pos), ("obs_dof_vel", s.obs_buf.dof_vel), ("obs_lin_vel", s.obs_buf.lin_vel),
                     ("obs_ang_vel", s.obs_buf.ang_vel), ("obs_last_action", s.obs_buf.last_action),
                     ("obs_last_last_action", s.obs_buf.last_last_action), ("obs_proj_grav", s.obs_buf.projected_gravity),
                     ("obs_clock", s.obs_buf.clock_inputs), ("obs_rpy", s.obs_buf.base_rpy)):
            out[k].append(v.numpy().copy())
        out["collide"].append(s.collide_buf.numpy().copy() if len(s.termination_contact_indices) else np.zeros(N, dtype=bool))
        out["r_term"].append(s.r_term_buff.numpy().copy() if "roll" in cfg.termination.termination_terms else np.zeros(N, dtype=bool))
        out["p_term"].append(s.p_term_buff.numpy().copy() if "pitch" in cfg.termination.termination_terms else np.zeros(N, dtype=bool))
        # the reference would now reset_idx(env_ids) with torch's RNG; the replay zeroes episode counters of reset envs instead
        s.last_actions[:] = s.actions[:]
    for k, v in out.items():
        rec[k] = np.stack(v)
    np.savez_compressed(os.path.join(OUT, f"bookkeeping_{task}.npz"), **rec)
    print(task, "N", N, "A", A, "P", P, "resets per step", rec["reset"].sum(1), "timeouts", rec["timeout"].sum(1), "collide", rec["collide"].sum(1))


def main():
    install_stubs()
    install_torch_utils()
    sys.path.insert(0, REF)
    run("go1gate", "mqe.envs.configs.go1_gate_config", "Go1GateCfg", "mqe.envs.go1.go1", "Go1", 12, 11)
    run("go1sheep-easy", "mqe.envs.configs.go1_sheep_config", "SingleSheepCfg", "mqe.envs.npc.go1_sheep", "Go1Sheep", 12, 12)
    run("go1football-defender", "mqe.envs.configs.go1_football_config", "Go1FootballDefenderCfg", "mqe.envs.npc.go1_football_defender", "Go1FootballDefender", 12, 13)
    run("go1seesaw", "mqe.envs.configs.go1_seesaw_config", "Go1SeesawCfg", "mqe.envs.npc.go1_object", "Go1Object", 12, 14)
    run("go1tug", "mqe.envs.configs.go1_tug_config", "Go1TugCfg", "mqe.envs.npc.go1_object", "Go1Object", 12, 15)
    run("go1wrestling", "mqe.envs.configs.go1_wrestling_config", "Go1WrestlingCfg", "mqe.envs.npc.go1_object", "Go1Object", 12, 16)
    run("go1bridge", "mqe.envs.configs.go1_bridge_config", "Go1BridgeCfg", "mqe.envs.npc.go1_object", "Go1Object", 12, 17)


if __name__ == "__main__":
    main()
