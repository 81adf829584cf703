"""Scene construction: cfg -> MqeSimDesc (the init-time work of LeggedRobot._create_envs / _get_env_origins).

Restates, without Isaac Gym, what `mqe/envs/base/legged_robot.py:754-923, 972-1011`,
`mqe/envs/go1/go1.py:411-479` and `mqe/envs/npc/*.py:_prepare_npc` compute at construction: terrain,
env / agent origins, initial root states of agents and NPCs, the default walk-these-ways command frame,
termination set and reset-randomisation ranges.  The result is the flat descriptor both the CUDA engine
and the CPU oracle consume.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import engine as E
from .model import Go1Model, load_go1_model
from .terrain.barrier_track import GROUND_SLAB_TOP, BarrierTrack

TASK_NPC = {            # npc asset -> (kind, ctrl, radius, half length of the capsule segment, mass, inertia)
    "sheep": (E.NPC_RIGID, E.NPC_SHEEP, 0.2, 0.05, 5.0, 0.01),       # resources/objects/sheep.urdf (cyl r .2 l .5)
    "ball": (E.NPC_RIGID, E.NPC_PASSIVE, 0.1, 0.0, 0.318, 0.00462),  # resources/objects/ball.urdf
    "seesaw": (E.NPC_SEESAW, E.NPC_PASSIVE, 0.0, 0.0, 100.0, 100.0), # resources/objects/seesaw.urdf
    "box": (E.NPC_BOX, E.NPC_PASSIVE, 0.0, 0.0, 6.0, 0.25),          # resources/objects/box.urdf (1 x 1 x 1 m, 6 kg)
    "rotation": (E.NPC_SEESAW, E.NPC_PASSIVE, 0.0, 0.0, 4.0, 1.232),  # resources/objects/rotation_door.urdf (izz of the panel)
    "circular": (E.NPC_SEESAW, E.NPC_PASSIVE, 0.0, 0.0, 3.0, 0.54),    # resources/objects/cylinder.urdf: 3 kg disc on a prismatic y joint
    "wrestling": (E.NPC_PLATFORM, E.NPC_PASSIVE, 0.0, 0.0, 0.0, 0.0),  # resources/objects/wrestling_field/urdf/wrestling.urdf (fixed)
    "bridge": (E.NPC_PLATFORM, E.NPC_PASSIVE, 0.0, 0.0, 0.0, 0.0),     # resources/objects/bridge/urdf/bridge.urdf (fixed)
}
# cylinder.urdf: joint "rot1" prismatic along y at (0, 0, 0.25), velocity limit 1.0 (limits +-10 never reached inside the track);
# collision cylinder r 1.2, length 0.5 centred 0.05 above the joint.  [13] = 2: prismatic y, [4] radius, [5] half height, [15] centre z.
TUG_GEOM = [0.0, 0.0, 0.25, 0.0, 1.2, 0.25, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0, 2.0, 0.0, 0.05]
# Fixed assets made of boxes (fix_npc_base_link = True; the STL meshes are 12-triangle boxes, extents read from the files):
# [n boxes, then per box: centre x, y, half extents x, y, top z] relative to the NPC root.
# wrestling.urdf: base_link 4.368 x 4.368 x 0.5 m; the eight Empty_Link* parts are 1-4 cm thick floor markings on its top, not modelled.
WRESTLING_GEOM = [1.0, 0.0, 0.0, 2.184, 2.184, 0.5] + [0.0] * 10
# bridge.urdf: deck base_link 4 x 0.7 x 0.3 m; Link1 / Link2 (2.5 x 1 x 1.3 m boxes, joint rpy (pi/2, 0, pi/2)) become end platforms
# x in +-[2.0, 3.3], |y| <= 1.25, z in [-0.7, 0.3]: all three tops at 0.3 above the root.
BRIDGE_GEOM = [3.0, 0.0, 0.0, 2.0, 0.35, 0.3, 2.65, 0.0, 0.65, 1.25, 0.3, -2.65, 0.0, 0.65, 1.25, 0.3]
# rotation_door.urdf: joint "rot1" at the base origin, axis z, velocity limit 28; panel box 0.08 x 1.95 x 0.8 centred 0.4 above it
DOOR_GEOM = [0.0, 0.0, 0.0, 0.0, 0.04, 0.975, 0.4, 0.0, 0.0, 0.0, 0.0, 0.0, 28.0, 1.0, 0.0, 0.4]
BOX_GEOM = [0.0] * 4 + [0.5, 0.5, 0.5] + [0.0] * 9
# resources/objects/seesaw.urdf: joint "link" origin (-2.4, 0, -0.455), axis y, velocity limit 0.2; plank box 4.123 x 1 x 0.03 at
# x = -0.1031 (its COM too); fixed base box 1 x 1 x 0.03; column cylinder r 0.2, length 1 hanging below the base.
# Layout: MqeSimDesc.npc_geom ([13] = hinge axis 0: y / 1: z, [3], [14], [15] = box centre in the hinged frame)
SEESAW_GEOM = [-2.4, 0.0, -0.455, -0.1031, 4.123 / 2, 0.5, 0.015, 0.5, 0.5, 0.015, 0.2, 1.0, 0.2, 0.0, 0.0, 0.0]


@dataclass
class Scene:
    cfg: object
    desc: E.SimDescC
    model: Go1Model
    num_envs: int
    num_agents: int
    num_npcs: int
    env_origins: np.ndarray          # [N,3]
    agent_origins: np.ndarray        # [N,A,3]
    base_init_state: np.ndarray      # [N*A,13]
    npc_init_state: np.ndarray       # [N*P,13]
    terrain_levels: np.ndarray
    terrain_types: np.ndarray
    terrain: object
    env_info: dict
    sdf: np.ndarray
    keep: list = field(default_factory=list)   # arrays the ctypes pointers reference


def default_command_frame(cfg) -> np.ndarray:
    """go1.py:411-479 `_fill_command_obs` for the command switches the tasks use."""
    c, s = cfg.control.default_command, cfg.control.obs_scales
    sw = cfg.command.cfg
    for name in ("body_height", "gait_freq", "gait", "footswing_height", "body_pose", "stance_width", "stance_length", "aux_reward"):
        if getattr(sw, name):
            raise NotImplementedError(f"command.cfg.{name}=True is not used by any registered task")
    f = np.zeros(70, dtype=np.float32)
    if not sw.vel:
        f[3], f[4], f[5] = c.lin_vel_x * s.lin_vel, c.lin_vel_y * s.lin_vel, c.ang_vel * s.ang_vel
    f[6] = c.body_height * s.body_height
    f[7] = c.gait_freq * s.gait_freq
    g = cfg.command.gaits[c.gait]
    f[8], f[9], f[10] = g[0] * s.gait_phase, g[1] * s.gait_phase, g[2] * s.gait_phase
    f[11] = 0.5 * s.gait_phase
    f[12] = c.footswing_height * s.footswing_height
    f[13], f[14] = c.body_pitch * s.body_pitch, c.body_roll * s.body_roll
    f[15] = c.stance_width * s.stance_width
    f[16] = c.stance_length * s.stance_length
    f[17] = c.aux_reward * s.aux_reward
    return f


def _terrain_assignment(num_envs_global, num_rows, num_cols, seed):
    """legged_robot.py:983-984.  The reference draws `terrain_levels` from torch's global (CUDA) stream; a
    dedicated CPU generator over the GLOBAL env range keeps shards consistent."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(int(seed) + 7919)
    levels = torch.randint(0, num_rows, (num_envs_global,), generator=g).numpy()
    types = np.arange(num_envs_global) % num_cols
    return levels, types


def _sheep_init_states(cfg):
    """go1_sheep.py:66-111 (np.random.randn(4) per sheep, same order)."""
    kw = cfg.terrain.BarrierTrack_kwargs
    rows, cols, dis = cfg.asset.num_rows, cfg.asset.num_cols, cfg.asset.dis_sheep
    origin = np.array([kw["init"]["block_length"] + kw["plane"]["block_length"] / 2 - rows // 2 * dis[0],
                       -(cols // 2) * dis[1], 0.3])
    pos = origin.copy()
    out = []
    for _ in range(rows):
        for _ in range(cols):
            rot = np.array([0.0, 0.0, 0.0, 1.0]) + np.random.randn(4) * np.array([0, 0, np.pi, 1])
            rot = rot / np.linalg.norm(rot)          # PhysX normalises the pose quaternion it is handed
            out.append(np.concatenate((pos, rot, np.zeros(3), np.zeros(3))))
            pos[1] += dis[1]
        pos[0] += dis[0]
        pos[1] = origin[1]
    return np.asarray(out, dtype=np.float32)


def build_scene(cfg, seed=0, env_slice=None, policy_mode=E.POLICY_BF16X3, solver_iters=None, model=None,
                wrapper_action_scale=(2.0, 0.5, 0.5), weights=None) -> Scene:
    """env_slice=(start, stop) selects this rank's contiguous block of the GLOBAL env range."""
    model = model or load_go1_model()
    N_global = int(cfg.env.num_envs)
    A, P = int(cfg.env.num_agents), int(getattr(cfg.env, "num_npcs", 0))
    start, stop = env_slice if env_slice is not None else (0, N_global)
    N = stop - start
    keep = []

    # ---- terrain (legged_robot_field.py:274-281) -------------------------------------------------
    if getattr(cfg.terrain, "mesh_type", None) == "trimesh" and isinstance(cfg.terrain.selected, str):
        assert cfg.terrain.selected == "BarrierTrack"
        terrain = BarrierTrack(cfg.terrain, N_global, A)
        terrain.add_terrain_to_sim(None, None, "cpu")
        sdf = terrain.wall_sdf()
        floor_z, wall_top = GROUND_SLAB_TOP, terrain.wall_top()
        levels, types = _terrain_assignment(N_global, cfg.terrain.num_rows, cfg.terrain.num_cols, seed)
        env_origins = terrain.env_origins[levels, types].astype(np.float32)
        agent_origins = terrain.agent_origins[levels, types].astype(np.float32)
        env_info = {k: v[levels, types] for k, v in terrain.env_info_np.items()}
        cell = float(cfg.terrain.horizontal_scale)
    else:                                                  # plane: legged_robot.py:999-1011
        terrain, env_info = None, {}
        sdf = np.full((2, 2), 1.0e3, dtype=np.float32)
        floor_z, wall_top, cell = 0.0, 1.0e3, 1.0e6
        ncols = np.floor(np.sqrt(N_global))
        nrows = np.ceil(N_global / ncols)
        xx, yy = np.meshgrid(np.arange(nrows), np.arange(ncols), indexing="ij")
        env_origins = np.zeros((N_global, 3), dtype=np.float32)
        env_origins[:, 0] = cfg.env.env_spacing * xx.flatten()[:N_global]
        env_origins[:, 1] = cfg.env.env_spacing * yy.flatten()[:N_global]
        agent_origins = np.repeat(env_origins[:, None, :], A, axis=1)
        levels = types = np.zeros(N_global, dtype=np.int64)

    # ---- initial root states (legged_robot.py:815-831) --------------------------------------------
    ist = cfg.init_state
    if getattr(ist, "multi_init_state", False):
        rows = [s.pos + s.rot + s.lin_vel + s.ang_vel for s in ist.init_states]
        assert len(rows) == A, "Mismatch num_agents and init_states"
    else:
        rows = [ist.pos + ist.rot + ist.lin_vel + ist.ang_vel] * A
    base_init = np.tile(np.asarray(rows, dtype=np.float32), (N_global, 1))

    # ---- NPCs --------------------------------------------------------------------------------------
    npc_kind = npc_ctrl = E.NPC_NONE
    npc_r = npc_hl = npc_m = npc_I = 0.0
    npc_init = np.zeros((0, 13), dtype=np.float32)
    npc_dofs = int(getattr(cfg.env, "num_actions_npc", 0)) * P
    npc_dof_default = np.zeros(max(npc_dofs, 1), dtype=np.float32)
    if P:
        name = cfg.asset.name_npc
        npc_kind, npc_ctrl, npc_r, npc_hl, npc_m, npc_I = TASK_NPC[name]
        if name == "sheep":
            one = _sheep_init_states(cfg)
        else:
            one = np.asarray([s.pos + s.rot + s.lin_vel + s.ang_vel for s in ist.init_states_npc], dtype=np.float32)
        assert one.shape[0] == P
        npc_init = np.tile(one, (N_global, 1))
        if hasattr(ist, "default_npc_joint_angles"):
            npc_dof_default[:npc_dofs] = np.asarray(ist.default_npc_joint_angles, dtype=np.float32)

    d = E.SimDescC()
    d.abi_version = E.ABI_VERSION
    d.num_envs, d.num_agents, d.num_npcs = N, A, P
    d.env_id_offset = start
    d.npc_kind, d.npc_ctrl, d.npc_dofs = npc_kind, npc_ctrl, npc_dofs
    d.decimation = int(cfg.control.decimation)
    px = cfg.sim.physx
    d.solver_iters = int(solver_iters if solver_iters is not None else 2 * px.num_position_iterations)
    dt = cfg.control.decimation * cfg.sim.dt
    d.max_episode_length = int(np.ceil(cfg.env.episode_length_s / dt))        # legged_robot.py:1021
    terms = cfg.termination.termination_terms
    mask = sum(bit for name, bit in (("roll", 1), ("pitch", 2), ("z_low", 4), ("z_high", 8)) if name in terms)
    if len(cfg.asset.terminate_after_contacts_on):
        assert cfg.asset.terminate_after_contacts_on == ["base"]
        mask |= 16
    d.term_mask = mask
    d.quat_alias = 1 if P == 0 else 0
    d.policy_mode = int(policy_mode)
    d.defender = 1 if getattr(cfg.env, "env_name", "") == "go1football" and A == 3 else 0
    d.sim_dt = float(cfg.sim.dt)
    d.gravity_z = float(cfg.sim.gravity[2])
    d.friction = 0.5 * (float(cfg.terrain.static_friction) + 1.0)             # PhysX average combine with the asset default 1.0
    d.contact_offset = float(px.contact_offset)
    d.max_depen_vel = float(px.max_depenetration_velocity)
    d.erp, d.cfm = 0.2, 1.0e-6
    d.floor_z, d.wall_top_z = float(floor_z), float(wall_top)
    d.limit_margin = 0.05
    t = cfg.termination
    d.term_roll, d.term_pitch = float(t.roll_kwargs["threshold"]), float(t.pitch_kwargs["threshold"])
    d.term_zlow, d.term_zhigh = float(t.z_low_kwargs["threshold"]), float(t.z_high_kwargs["threshold"])
    d.command_vel = int(bool(cfg.command.cfg.vel))
    d.act_scale[:] = [float(x) for x in wrapper_action_scale]
    s = cfg.control.obs_scales
    d.cmd_scale[:] = [s.lin_vel, s.lin_vel, s.ang_vel]
    d.action_scale = float(cfg.control.action_scale)
    d.hip_scale = float(cfg.control.hip_scale_reduction)
    d.clip_actions = float(cfg.normalization.clip_actions)
    d.loc_obs_default[:] = default_command_frame(cfg).tolist()
    dr = cfg.domain_rand
    ratio = getattr(dr, "init_dof_pos_ratio_range", None) or (1.0, 1.0)
    d.dof_ratio_lo, d.dof_ratio_hi = float(ratio[0]), float(ratio[1])
    vel = getattr(dr, "init_base_vel_range", None) or (-0.5, 0.5)
    d.base_vel_lo, d.base_vel_hi = float(vel[0]), float(vel[1])
    r = getattr(dr, "init_base_pos_range", None)
    d.has_base_pos_range = int(r is not None)
    if r is not None:
        d.base_pos_x[:], d.base_pos_y[:] = [float(v) for v in r["x"]], [float(v) for v in r["y"]]
    r = getattr(dr, "init_npc_base_pos_range", None) if P else None
    d.has_npc_pos_range = int(r is not None)
    if r is not None:
        d.npc_pos_x[:], d.npc_pos_y[:] = [float(v) for v in r["x"]], [float(v) for v in r["y"]]
    r = getattr(dr, "init_npc_base_rpy_range", None) if P else None
    d.has_npc_rpy_range = int(r is not None)
    if r is not None:
        d.npc_rpy_r[:], d.npc_rpy_p[:], d.npc_rpy_y[:] = list(map(float, r["r"])), list(map(float, r["p"])), list(map(float, r["y"]))
    d.npc_mass, d.npc_inertia, d.npc_radius, d.npc_halflen = npc_m, npc_I, npc_r, npc_hl
    # pair-contact budget per env and substep: two robots alone rarely touch in more than a few capsule pairs
    d.max_pair_contacts = 8 if (A <= 2 and npc_kind == E.NPC_NONE) else 16
    geom = {"seesaw": SEESAW_GEOM, "rotation": DOOR_GEOM, "box": BOX_GEOM, "wrestling": WRESTLING_GEOM, "bridge": BRIDGE_GEOM, "circular": TUG_GEOM}.get(cfg.asset.name_npc if P else "", [0.0] * 16)
    d.npc_geom[:] = geom
    d.sheep_scale = float(getattr(cfg.asset, "sheep_movement_scale", 0.0))
    d.sheep_randomness = float(getattr(cfg.asset, "sheep_movement_randomness", 0.0))
    if d.defender:
        kw = cfg.terrain.BarrierTrack_kwargs
        d.gate_x = float(kw["init"]["block_length"] + kw["plane"]["block_length"])   # go1_football_defender.py:61-63
    d.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    ct = str(cfg.control.control_type)
    assert ct in E.CONTROL_TYPES, f"unknown control_type {ct!r} (legged_robot.py:384-392)"
    d.control_type = E.CONTROL_TYPES[ct]
    d.soft_dof_pos_limit = float(getattr(cfg.rewards, "soft_dof_pos_limit", 1.0))          # legged_robot.py:318-321
    d.stiffness, d.damping = float(cfg.control.stiffness.get("joint", 0.0)), float(cfg.control.damping.get("joint", 0.0))
    if getattr(dr, "randomize_lag_timesteps", False):                          # go1.py:337-339, 363
        d.lag_enabled, d.lag_timesteps = 1, int(dr.lag_timesteps)
    # domain randomisation switches of the reference that are off in its task configs
    if getattr(dr, "push_robots", False):                                      # legged_robot.py:1024, go1.py:237-238
        d.push_interval = int(np.ceil(dr.push_interval_s / dt))
        d.max_push_vel_xy = float(dr.max_push_vel_xy)
    base_added_mass = None
    if getattr(dr, "randomize_base_mass", False):                              # legged_robot.py:332-335: props[0].mass += U(added_mass_range)
        mrng = np.random.Generator(np.random.Philox(key=(int(seed) & 0xFFFFFFFF) * 7919 + 29))
        base_added_mass = np.ascontiguousarray(mrng.uniform(dr.added_mass_range[0], dr.added_mass_range[1], size=(N_global, A)),
                                               dtype=np.float32)[start:stop].reshape(-1).copy()
    base_com_shift = None
    if getattr(dr, "randomize_com", False):                                    # legged_robot_field.py:321-332: props[0].com += U(com_range), per robot
        crng = np.random.Generator(np.random.Philox(key=(int(seed) & 0xFFFFFFFF) * 7919 + 31))
        cr = dr.com_range
        lo, hi = np.array([cr.x[0], cr.y[0], cr.z[0]]), np.array([cr.x[1], cr.y[1], cr.z[1]])
        base_com_shift = np.ascontiguousarray(crng.uniform(lo, hi, size=(N_global, A, 3)), dtype=np.float32)[start:stop].reshape(-1, 3).copy()
    motor_strength = None
    if getattr(dr, "randomize_motor", False):                                  # legged_robot_field.py:283-291: U(leg_motor_strength_range) per env and joint
        srng = np.random.Generator(np.random.Philox(key=(int(seed) & 0xFFFFFFFF) * 7919 + 37))
        motor_strength = np.ascontiguousarray(srng.uniform(dr.leg_motor_strength_range[0], dr.leg_motor_strength_range[1], size=(N_global, 12 * A)),
                                              dtype=np.float32)[start:stop].copy()
    env_friction = None
    if getattr(dr, "randomize_friction", False):                               # legged_robot.py:283-294: 64 buckets over friction_range,
        frng = np.random.Generator(np.random.Philox(key=(int(seed) & 0xFFFFFFFF) * 7919 + 13))   # one bucket per env (own generator: shards agree)
        buckets = frng.uniform(dr.friction_range[0], dr.friction_range[1], size=64)
        coeff = buckets[frng.integers(0, 64, size=N_global)]
        env_friction = np.ascontiguousarray(0.5 * (coeff + float(cfg.terrain.static_friction)), dtype=np.float32)[start:stop].copy()

    sdf = np.ascontiguousarray(sdf, dtype=np.float32)
    d.sdf_nx, d.sdf_ny, d.sdf_cell = sdf.shape[0], sdf.shape[1], cell
    eo = np.ascontiguousarray(env_origins[start:stop], dtype=np.float32)
    ao = np.ascontiguousarray(agent_origins[start:stop], dtype=np.float32)
    bi = np.ascontiguousarray(base_init[start * A:stop * A], dtype=np.float32)
    bi_engine = bi.copy()                                  # PhysX normalises the pose quaternion it is handed (go1_wrestling_config.py
    bi_engine[:, 3:7] /= np.linalg.norm(bi_engine[:, 3:7], axis=1, keepdims=True)   # gives rot = [0, 0, -1, 1]); `base_init_state` stays raw
    ni = np.ascontiguousarray(npc_init[start * P:stop * P], dtype=np.float32) if P else np.zeros((1, 13), dtype=np.float32)
    keep += [sdf, eo, ao, bi, bi_engine, ni, npc_dof_default]
    if base_added_mass is not None:
        keep.append(base_added_mass)
        d.h_base_added_mass = E.as_fp(base_added_mass)
    if env_friction is not None:
        keep.append(env_friction)
        d.h_env_friction = E.as_fp(env_friction)
    if base_com_shift is not None:
        keep.append(base_com_shift)
        d.h_base_com_shift = E.as_fp(base_com_shift)
    if motor_strength is not None:
        keep.append(motor_strength)
        d.h_motor_strength = E.as_fp(motor_strength)
    d.h_sdf, d.h_env_origins, d.h_agent_origins = E.as_fp(sdf), E.as_fp(eo), E.as_fp(ao)
    d.h_base_init_state, d.h_npc_init_state, d.h_npc_dof_default = E.as_fp(bi_engine), E.as_fp(ni), E.as_fp(npc_dof_default)
    d.model = model.to_c()
    if weights is None:
        weights = E.load_weights()
    d.weights, warrays = weights
    keep.append(warrays)
    return Scene(cfg=cfg, desc=d, model=model, num_envs=N, num_agents=A, num_npcs=P, env_origins=eo,
                 agent_origins=ao, base_init_state=bi, npc_init_state=ni, terrain_levels=levels[start:stop],
                 terrain_types=types[start:stop], terrain=terrain, env_info={k: v[start:stop] for k, v in env_info.items()},
                 sdf=sdf, keep=keep)
