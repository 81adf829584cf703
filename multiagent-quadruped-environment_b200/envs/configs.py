"""Task configurations with the reference's attribute surface (cfg.env.num_envs, cfg.terrain.BarrierTrack_kwargs, ...).

The reference spells every task as a tower of nested Python classes
(LeggedRobotCfg -> LeggedRobotFieldCfg -> Go1Cfg -> Go1<Task>Cfg; `mqe/envs/base/legged_robot_config.py`,
`mqe/envs/field/legged_robot_field_config.py`, `mqe/envs/go1/go1_config.py`, `mqe/envs/configs/*.py`) and
passes the *class* around as a mutable namespace (`mqe/envs/utils.py:111-134`).  Here the same tree is data:
`Cfg` is an attribute namespace, `go1_base()` builds the inherited defaults once and each task applies its
overrides.  Only values that reach the hot path or the wrappers are carried; the numbers cite the reference.
"""
from __future__ import annotations

import copy
from types import SimpleNamespace


class Cfg(SimpleNamespace):
    """Attribute namespace; dict-valued leaves (BarrierTrack_kwargs, joint angle tables, ...) stay dicts."""

    def clone(self):
        return copy.deepcopy(self)

    def update(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)
        return self

    def to_dict(self):
        out = {}
        for k, v in vars(self).items():
            out[k] = v.to_dict() if isinstance(v, Cfg) else v
        return out

    def __contains__(self, key):
        return hasattr(self, key)


class InitState(Cfg):
    """`LeggedRobotCfg.init_state` doubles as a record type (legged_robot_config.py:65-83)."""

    def __init__(self, pos=(0.0, 0.0, 1.0), rot=(0.0, 0.0, 0.0, 1.0), lin_vel=(0.0, 0.0, 0.0), ang_vel=(0.0, 0.0, 0.0), **kw):
        super().__init__(pos=list(pos), rot=list(rot), lin_vel=list(lin_vel), ang_vel=list(ang_vel), **kw)


def merge_dict(this: dict, other: dict) -> dict:
    """helpers.py:237-243"""
    out = dict(this)
    out.update(other)
    return out


class ObsCfg(Cfg):
    def keys(self):                                       # go1_config.py:283-289
        return [k for k, v in vars(self.cfgs).items() if v is True]


_FIELD_TRACK = dict(                                      # legged_robot_field_config.py:21-59
    options=["init", "gate", "wall", "plane"], wall_thickness=0.04, track_width=2.0,
    wall=dict(block_length=3.0), plane=dict(block_length=3.0),
    init=dict(block_length=3.0, room_size=(1.0, 1.0), border_width=0.0, offset=(0, 0)),
    gate=dict(block_length=1.6, width=0.5, depth=0.1, offset=(0.4, 0), random=(0.0, 0.0)),
    wall_height=0.5, virtual_terrain=False, no_perlin_threshold=0.06, add_perlin_noise=False,
    border_perlin_noise=False, border_height=0.0, engaging_next_threshold=1.2, curriculum_perlin=False,
)

_STAND = dict(pos=(0.0, 0.0, 0.42))


def go1_base() -> Cfg:
    """Go1Cfg with everything it inherits (go1_config.py:34-311)."""
    c = Cfg()
    c.env = Cfg(env_name="go1", num_envs=256, num_agents=1, num_npcs=0, num_actions_npc=0, num_observations=235,
                use_lin_vel=True, num_privileged_obs=None, num_actions=12, env_spacing=3.0, send_timeouts=True,
                episode_length_s=5, record_video=False, record_actor_id=0, recording_width_px=360,
                recording_height_px=240, recording_mode="COLOR")
    c.terrain = Cfg(                                      # legged_robot_config.py:16-48 + field config :7-59
        mesh_type="trimesh", horizontal_scale=0.025, vertical_scale=0.005, border_size=1, curriculum=False,
        static_friction=1.0, dynamic_friction=1.0, restitution=0.0, measure_heights=True, selected="BarrierTrack",
        max_init_terrain_level=0, num_rows=20, num_cols=50, x_init_range=1.0, y_init_range=1.0, yaw_init_range=0.0,
        slope_treshold=100.0, pad_unavailable_info=True, BarrierTrack_kwargs=copy.deepcopy(_FIELD_TRACK),
        TerrainPerlin_kwargs=dict(zScale=0.12, frequency=10))
    c.commands = Cfg(curriculum=False, max_curriculum=1.0, num_commands=4, resampling_time=10.0, heading_command=True,
                     ranges=Cfg(lin_vel_x=[-1.0, 1.0], lin_vel_y=[-1.0, 1.0], ang_vel_yaw=[-1, 1], heading=[-3.14, 3.14]))
    c.command = Cfg(                                      # go1_config.py:157-188
        gaits={"pronking": [0, 0, 0], "trotting": [0.5, 0, 0], "bounding": [0, 0.5, 0], "pacing": [0, 0, 0.5]},
        curriculum=False, max_curriculum=1.0, num_commands=4, resampling_time=10.0, heading_command=True,
        cfg=Cfg(vel=False, body_height=False, body_pose=False, gait_freq=False, gait=False, footswing_height=False,
                stance_width=False, stance_length=False, aux_reward=False),
        ranges=Cfg(lin_vel_x=[-1.0, 1.0], lin_vel_y=[-1.0, 1.0], ang_vel_yaw=[-1, 1], heading=[-3.14, 3.14]))
    c.init_state = InitState(
        pos=(0.0, 0.0, 0.42),
        default_joint_angles={                            # go1_config.py:88-103
            "FL_hip_joint": 0.1, "RL_hip_joint": 0.1, "FR_hip_joint": -0.1, "RR_hip_joint": -0.1,
            "FL_thigh_joint": 0.8, "RL_thigh_joint": 1.0, "FR_thigh_joint": 0.8, "RR_thigh_joint": 1.0,
            "FL_calf_joint": -1.5, "RL_calf_joint": -1.5, "FR_calf_joint": -1.5, "RR_calf_joint": -1.5})
    c.normalization = Cfg(obs_scales=Cfg(lin_vel=2.0, ang_vel=0.25, dof_pos=1.0, dof_vel=0.05, height_measurements=5.0),
                          clip_observations=100.0, clip_actions=10.0)
    c.control = Cfg(                                      # go1_config.py:107-155
        control_type="C", stiffness={"joint": 20.0}, damping={"joint": 0.5}, action_scale=0.25,
        torque_limits=[20.0, 20.0, 25.0] * 4, computer_clip_torque=True, motor_clip_torque=False, decimation=4,
        hip_scale_reduction=0.5,
        locomotion_policy_dir="./mqe/utils/locomotion_checkpoints/walk_these_ways",
        actuator_network_path="./resources/actuator_nets",
        default_command=Cfg(lin_vel_x=1.0, lin_vel_y=-0.0, ang_vel=-0.0, body_height=0.0, gait_freq=3.0, gait="trotting",
                            footswing_height=0.08, body_pitch=0.0, body_roll=0.0, stance_width=0.25,
                            stance_length=0.428, aux_reward=0.0),
        obs_scales=Cfg(lin_vel=2.0, ang_vel=0.25, dof_pos=1.0, dof_vel=0.05, body_height=2.0, gait_phase=1.0,
                       gait_freq=1.0, footswing_height=0.15, body_pitch=0.3, body_roll=0.3, aux_reward=1.0,
                       compliance=1.0, stance_width=1.0, stance_length=1.0))
    c.asset = Cfg(                                        # go1_config.py:55-84
        file="{LEGGED_GYM_ROOT_DIR}/resources/robots/go1/urdf/go1.urdf", file_npc="", name="go1", name_npc="",
        foot_name="foot", penalize_contacts_on=["base", "thigh"], terminate_after_contacts_on=["base"],
        disable_gravity=False, collapse_fixed_joints=True, fix_base_link=False, default_dof_drive_mode=3,
        self_collisions=0, replace_cylinder_with_capsule=True, flip_visual_attachments=False, density=0.001,
        angular_damping=0.0, linear_damping=0.0, max_angular_velocity=1000.0, max_linear_velocity=1000.0,
        armature=0.0, thickness=0.01)
    c.termination = Cfg(                                  # go1_config.py:190-214
        termination_terms=["roll", "pitch", "z_low", "z_high"], roll_kwargs=dict(threshold=0.8),
        pitch_kwargs=dict(threshold=1.6), z_low_kwargs=dict(threshold=0.08), z_high_kwargs=dict(threshold=1.5),
        out_of_track_kwargs=dict(threshold=1.0))
    c.domain_rand = Cfg(                                  # legged_robot_config.py:135-144 + go1_config.py:216-247
        randomize_friction=False, friction_range=[0.05, 4.5], randomize_base_mass=False, added_mass_range=[-1.0, 3.0],
        push_robots=False, push_interval_s=15, max_push_vel_xy=1.0, max_push_vel_ang=0.0, randomize_com=False,
        com_range=Cfg(x=[-0.05, 0.15], y=[-0.1, 0.1], z=[-0.05, 0.05]),
        randomize_motor=False, leg_motor_strength_range=[0.9, 1.1], randomize_lag_timesteps=False, lag_timesteps=6,
        init_base_pos_range=dict(x=[0.1, 0.1], y=[-0.1, 0.1]), init_dof_pos_ratio_range=[0.7, 1.3],
        init_npc_base_pos_range=dict(x=[-0.2, 0.2], y=[-0.2, 0.2]))
    c.obs = ObsCfg(                                       # go1_config.py:249-289
        cfgs=Cfg(base_pos=True, base_quat=True, dof_pos=True, dof_vel=True, lin_vel=True, ang_vel=True,
                 projected_gravity=True, base_rpy=True, contact_states=False, command=True, height_command=False,
                 gait_commands=False, timing_parameter=False, clock_inputs=False, last_action=True,
                 last_last_action=True, imu=False, depth_image=False, rgb_image=False, env_info=True),
        scales=Cfg(base_pos=1.0, base_quat=1.0, segmentation_image=1.0, rgb_image=1.0, depth_image=1.0))
    c.privileged_obs = Cfg(cfgs=Cfg())
    c.rewards = Cfg(only_positive_rewards=True, tracking_sigma=0.25, soft_dof_pos_limit=0.9, soft_dof_vel_limit=1.0,
                    soft_torque_limit=1.0, base_height_target=0.25, max_contact_force=100.0,
                    scales=Cfg(torques=-0.0002, dof_pos_limits=-10.0))
    c.noise = Cfg(add_noise=True, noise_level=1.0)
    c.viewer = Cfg(ref_env=0, pos=[0.0, 11.0, 5.0], lookat=[4.0, 11.0, 0.0])
    c.sim = Cfg(                                          # legged_robot_config.py:211-229
        dt=0.005, substeps=1, gravity=[0.0, 0.0, -9.81], up_axis=1, no_camera=True,
        physx=Cfg(num_threads=10, solver_type=1, num_position_iterations=4, num_velocity_iterations=0,
                  contact_offset=0.01, rest_offset=0.0, bounce_threshold_velocity=0.5, max_depenetration_velocity=1.0,
                  max_gpu_contact_pairs=2 ** 23, default_buffer_size_multiplier=5, contact_collection=2))
    c.sensor = Cfg(forward_camera=Cfg(resolution=[16, 16], position=[0.26, 0.0, 0.03], rotation=[0.0, 0.0, 0.0]))
    return c


def _track(**over):
    return merge_dict(_FIELD_TRACK, over)


def _two_standing(c):
    c.init_state.multi_init_state = True
    c.init_state.init_states = [InitState(**_STAND), InitState(**_STAND)]


def Go1PlaneCfg() -> Cfg:
    """go1_plane_config.py: flat ground plane, grid of env origins, fixed default command (vel=False)."""
    c = go1_base()
    c.env.update(env_name="go1plane", num_envs=25, num_agents=1, episode_length_s=20, num_recording_envs=1)
    c.terrain = Cfg(mesh_type="plane", selected=False, static_friction=1.0, dynamic_friction=1.0, restitution=0.0,
                    curriculum=False, x_init_range=1.0, y_init_range=1.0, yaw_init_range=0.0, x_init_offset=0.0,
                    y_init_offset=0.0, num_rows=1, num_cols=1, horizontal_scale=0.025, vertical_scale=0.005)
    return c


def Go1GateCfg() -> Cfg:
    """go1_gate_config.py:5-130"""
    c = go1_base()
    c.env.update(env_name="go1gate", num_envs=1, num_agents=2, episode_length_s=10)
    c.terrain.update(num_rows=1, num_cols=1, BarrierTrack_kwargs=_track(
        options=["init", "gate", "plane", "wall"], track_width=3.0,
        init=dict(block_length=2.0, room_size=(1.0, 1.5), border_width=0.0, offset=(0, 0)),
        gate=dict(block_length=3.0, width=0.6, depth=0.1, offset=(0, 0), random=(0.5, 0.5)),
        plane=dict(block_length=1.0), wall=dict(block_length=0.1), wall_height=0.5))
    c.command.cfg.vel = True
    _two_standing(c)
    c.termination.update(check_obstacle_conditioned_threshold=False, termination_terms=["roll", "pitch", "z_low", "z_high"])
    c.domain_rand.init_base_pos_range = None
    c.rewards.scales = Cfg(target_reward_scale=1, success_reward_scale=5, lin_vel_x_reward_scale=0,
                           approach_frame_punishment_scale=0, agent_distance_punishment_scale=-0.025,
                           contact_punishment_scale=-2, lin_vel_y_punishment_scale=0, command_value_punishment_scale=0)
    c.viewer.update(pos=[-2.0, 2.5, 4.0], lookat=[4.0, 2.5, 0.0])
    return c


def _sheep(num_npcs, grid, randomness, rows, cols, track, rewards, num_envs):
    c = go1_base()
    c.env.update(env_name="go1sheep", num_envs=num_envs, num_agents=2, num_npcs=num_npcs, episode_length_s=15)
    c.asset.update(file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/sheep.urdf", name_npc="sheep", num_rows=grid,
                   num_cols=grid, dis_sheep=(1.5, 1.5), sheep_movement_scale=0.2, sheep_movement_randomness=randomness,
                   sheep_movement_range=[2.0, 2.0, 0])
    c.terrain.update(num_rows=rows, num_cols=cols, BarrierTrack_kwargs=track)
    c.command.cfg.vel = True
    _two_standing(c)
    c.termination.update(check_obstacle_conditioned_threshold=False, termination_terms=["roll", "pitch"])
    c.domain_rand.init_base_pos_range = dict(x=[-0.1, 0.1], y=[-0.1, 0.1])
    c.domain_rand.init_npc_base_pos_range = dict(x=[-0.3, 0.3], y=[-0.3, 0.3])
    c.rewards.scales = Cfg(**rewards)
    c.viewer.update(pos=[0.0, 3.0, 5.0], lookat=[4.0, 3.0, 0.0])
    return c


def SingleSheepCfg() -> Cfg:
    """go1_sheep_config.py:5-130"""
    return _sheep(1, 1, 0.0, 1, 1, _track(
        options=["init", "plane", "gate", "plane", "wall"], track_width=4.0,
        init=dict(block_length=1.5, room_size=(1.0, 1.95), border_width=0.0, offset=(0.5, 0)),
        gate=dict(block_length=1.0, width=0.8, depth=0.1, offset=(0, 0), random=(0, 0.5)),
        plane=dict(block_length=3.0), wall=dict(block_length=0.1), wall_height=0.5),
        dict(success_reward_scale=1, contact_punishment_scale=0, sheep_movement_reward_scale=2, mixed_sheep_reward_scale=0,
             sheep_pos_var_exp_punishment_scale=0, sheep_pos_var_lin_punishment_scale=0), 1)


def NineSheepCfg() -> Cfg:
    """go1_sheep_config.py:132-256"""
    return _sheep(9, 3, 0.1, 5, 7, _track(
        options=["init", "plane", "gate", "plane", "wall"], track_width=6.0,
        init=dict(block_length=2, room_size=(1.0, 3), border_width=0.0, offset=(0.5, 0)),
        gate=dict(block_length=1.0, width=1.5, depth=0.1, offset=(0, 0), random=(0, 1)),
        plane=dict(block_length=6.0), wall=dict(block_length=0.1), wall_height=0.5),
        dict(success_reward_scale=0, contact_punishment_scale=0, sheep_movement_reward_scale=0, mixed_sheep_reward_scale=1,
             sheep_pos_var_exp_punishment_scale=0, sheep_pos_var_lin_punishment_scale=0), 35)


def Go1SeesawCfg() -> Cfg:
    """go1_seesaw_config.py:5-136"""
    c = go1_base()
    c.env.update(env_name="go1seesaw", num_envs=1, num_agents=2, num_npcs=1, num_actions_npc=1, episode_length_s=10)
    c.asset.update(file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/seesaw.urdf", name_npc="seesaw", npc_collision=True,
                   fix_npc_base_link=True, npc_gravity=True)
    c.terrain.update(num_rows=1, num_cols=1, BarrierTrack_kwargs=_track(
        options=["init", "plane", "wall"], track_width=3.0,
        init=dict(block_length=2.0, room_size=(1.0, 1.5), border_width=0.0, offset=(0, 0)),
        plane=dict(block_length=8.0), wall=dict(block_length=0.1), wall_height=0.5))
    c.command.cfg.vel = True
    _two_standing(c)
    c.init_state.init_states_npc = [InitState(pos=(8.0, 0.0, 1.0))]
    c.init_state.default_npc_joint_angles = [-0.2]
    c.control.default_command.gait = "pacing"
    c.termination.update(check_obstacle_conditioned_threshold=False, termination_terms=["roll", "pitch", "z_low"])
    c.domain_rand.init_base_pos_range = dict(x=[-0.1, 0.1], y=[-0.1, 0.1])
    c.domain_rand.init_npc_base_pos_range = None
    c.obs.cfgs.env_info = False
    c.rewards.scales = Cfg(height_reward_scale=1, success_reward_scale=10, contact_punishment_scale=-2,
                           agent_distance_punishment_scale=-0.25, x_movement_reward_scale=5, fall_punishment_scale=-2,
                           y_punishment_scale=-0.5)
    c.viewer.update(pos=[0.0, -2.0, 4.0], lookat=[4.0, 2.0, 0.0])
    return c


def Go1FootballDefenderCfg() -> Cfg:
    """go1_football_config.py:5-131"""
    c = go1_base()
    c.env.update(env_name="go1football", num_envs=1, num_agents=3, num_npcs=1, episode_length_s=20)
    c.asset.update(file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/ball.urdf", name_npc="ball",
                   terminate_after_contacts_on=[], npc_collision=True, fix_npc_base_link=False, npc_gravity=True)
    c.terrain.update(num_rows=1, num_cols=1, BarrierTrack_kwargs=_track(
        options=["init", "gate", "plane", "gate", "wall"], track_width=9.0,
        init=dict(block_length=1.0, room_size=(0, 3.0), border_width=0.0, offset=(0.5, 0)),
        plane=dict(block_length=10.0),
        gate=dict(block_length=1.0, width=2.0, depth=1.0, offset=(0, 0), random=(0, 0.0)),
        wall=dict(block_length=0.1), wall_height=1.0))
    c.command.cfg.vel = True
    c.init_state.multi_init_state = True
    c.init_state.init_states = [InitState(pos=(3.0, 1.0, 0.42)), InitState(pos=(3.0, 2.0, 0.42)),
                                InitState(pos=(9.0, -3.0, 0.42), rot=(0.0, 0.0, 1.0, 0.0))]
    c.init_state.init_states_npc = [InitState(pos=(5.0, -2.1, 0.3))]
    c.termination.update(check_obstacle_conditioned_threshold=False, termination_terms=["roll", "pitch"])
    c.domain_rand.init_base_pos_range = dict(x=[-0.1, 0.1], y=[-0.1, 0.1])
    c.rewards.scales = Cfg(goal_reward_scale=10, ball_gate_distance_reward_scale=3)
    c.viewer.update(pos=[2.0, 2.0, 2.0], lookat=[6.0, 5.0, 0.0])
    return c


def Go1PushboxCfg() -> Cfg:
    """go1_pushbox_config.py: two robots push a 1 m, 6 kg box along the track."""
    c = go1_base()
    c.env.update(env_name="go1pushbox", num_envs=1, num_agents=2, num_npcs=1, episode_length_s=15)
    c.asset.update(terminate_after_contacts_on=[], file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/box.urdf", name_npc="box",
                   npc_collision=True, fix_npc_base_link=False, npc_gravity=True)
    c.terrain.update(num_rows=1, num_cols=1, BarrierTrack_kwargs=_track(
        options=["init", "gate", "wall"], track_width=5.0,
        init=dict(block_length=2.0, room_size=(1.0, 2.5), border_width=0.0, offset=(0, 0)),
        gate=dict(block_length=5.0, width=1.5, depth=0.1, offset=(0, 0), random=(0, 0.5)),
        wall=dict(block_length=0.1), wall_height=0.5))
    c.command.cfg.vel = True
    c.init_state.multi_init_state = True
    c.init_state.init_states = [InitState(pos=(0.0, 0.0, 0.42)), InitState(pos=(0.0, 0.0, 0.42))]
    c.init_state.init_states_npc = [InitState(pos=(2.5, 0.0, 0.6))]
    c.termination.update(check_obstacle_conditioned_threshold=False, termination_terms=["roll", "pitch"])
    c.domain_rand.init_base_pos_range = dict(x=[-0.1, 0.1], y=[-0.1, 0.1])
    c.domain_rand.init_npc_base_pos_range = dict(x=[-0.5, 0.5], y=[-0.5, 0.5])
    c.rewards.scales = Cfg(box_x_movement_reward_scale=10)
    c.viewer.update(pos=[0.0, 6.0, 5.0], lookat=[4.0, 6.0, 0.0])
    return c


def Go1RotationCfg() -> Cfg:
    """go1_rotation_config.py (task go1revolvingdoor): a door panel on a vertical hinge stands in the gate opening."""
    c = go1_base()
    c.env.update(env_name="go1rotationCfg", num_envs=1, num_agents=2, num_npcs=1, num_actions_npc=1, episode_length_s=5)
    c.asset.update(terminate_after_contacts_on=[], file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/rotation_door.urdf",
                   name_npc="rotation", npc_collision=True, fix_npc_base_link=True)
    c.terrain.update(num_rows=1, num_cols=1, x_limits=[5.0], y_limits=[-1.5, 1.5], BarrierTrack_kwargs=_track(
        options=["init", "wall", "gate", "wall"], randomize_obstacle_order=False, track_width=3.5,
        init=dict(block_length=0, room_size=(0.0, 0.0), border_width=0.0, offset=(0, 0)),
        gate=dict(block_length=5.0, width=2.0, depth=0.1, offset=(0, 0), random=(0, 0)),
        rotation=dict(block_length=5, depth=0.1, offset=(0, 0), wide_px=(0.84, 0.2)),
        wall=dict(block_length=0.1), wall_height=0.85))
    c.command.cfg.vel = True
    c.init_state.multi_init_state = True
    c.init_state.init_states = [InitState(pos=(0.5, -1.0, 0.42)), InitState(pos=(0.5, 1.0, 0.42))]
    c.init_state.init_states_npc = [InitState(pos=(2.59, -0.01, 0.04))]
    c.init_state.default_npc_joint_angles = [0.0]
    c.termination.update(termination_terms=["roll", "pitch", "z_low", "z_high"])
    c.domain_rand.init_base_pos_range = None
    c.domain_rand.init_npc_base_pos_range = None
    c.rewards.scales = Cfg(punishment_scale=1, success_reward_scale=10, distance_reward_scale=1)
    c.viewer.update(pos=[12.0, 20.0, 20.0], lookat=[13.0, 20.0, 0.0])
    return c


def Go1TugCfg() -> Cfg:
    """go1_tug_config.py:5-124: two robots push a 1.2 m disc that slides along y between them (cylinder.urdf, prismatic joint)."""
    c = go1_base()
    c.env.update(env_name="go1tug", num_envs=1, num_agents=2, num_npcs=1, num_actions_npc=1, episode_length_s=15)
    c.asset.update(terminate_after_contacts_on=[], file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/cylinder.urdf",
                   name_npc="circular", fix_npc_base_link=True)
    c.terrain.update(num_rows=1, num_cols=1, BarrierTrack_kwargs=_track(
        options=["init", "wall", "plane", "wall"], randomize_obstacle_order=False, track_width=6.0,
        init=dict(block_length=0.0, room_size=(0.0, 0.0), border_width=0.0, offset=(0, 0)),
        plane=dict(block_length=3.0), wall=dict(block_length=0.1), wall_height=1.0))
    c.command.cfg.vel = True
    c.init_state.multi_init_state = True
    c.init_state.init_states = [InitState(pos=(1.6, 2.5, 0.34), rot=(0.0, 0.0, -1.0, 1.0)), InitState(pos=(1.6, -2.5, 0.34), rot=(0.0, 0.0, 1.0, 1.0))]
    c.init_state.init_states_npc = [InitState(pos=(1.6, 0.0, 0.0))]
    c.termination.update(termination_terms=["roll", "pitch", "z_low", "z_high"])
    c.domain_rand.init_dof_pos_ratio_range = None
    c.domain_rand.init_base_pos_range = dict(x=[-1.0, 1.0], y=[-0.0, 0.0])
    c.domain_rand.init_npc_base_pos_range = None
    c.rewards.scales = Cfg(success_reward_scale=10, punishment_reward_scale=10, pos_reward_scale=2, pos_punishment_scale=2)
    c.viewer.update(pos=[0.0, 11.0, 5.0], lookat=[4.0, 11.0, 0.0])
    return c


def Go1WrestlingCfg() -> Cfg:
    """go1_wrestling_config.py:5-120: two robots on a fixed 4.37 m square platform 0.5 m high (wrestling.urdf)."""
    c = go1_base()
    c.env.update(env_name="go1wrestling", num_envs=1, num_agents=2, num_npcs=1, episode_length_s=15)
    c.asset.update(terminate_after_contacts_on=[], file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/wrestling_field/urdf/wrestling.urdf",
                   name_npc="wrestling", fix_npc_base_link=True)
    c.terrain.update(num_rows=1, num_cols=1, BarrierTrack_kwargs=_track(
        options=["init", "plane"], randomize_obstacle_order=False, track_width=6,
        init=dict(block_length=0.0, room_size=(0.0, 0.0), border_width=0.0, offset=(0, 0)),
        wall=dict(block_length=0.1), plane=dict(block_length=7), wall_height=0.001))
    c.command.cfg.vel = True
    c.init_state.multi_init_state = True
    c.init_state.init_states = [InitState(pos=(3.1, 1.0, 0.74), rot=(0.0, 0.0, -1.0, 1.0)), InitState(pos=(3.1, -1.0, 0.74), rot=(0.0, 0.0, 1.0, 1.0))]
    c.init_state.init_states_npc = [InitState(pos=(3.1, 0.0, 0.0))]
    c.termination.update(termination_terms=["roll", "pitch", "z_low"], z_low_kwargs=dict(threshold=0.3))
    c.domain_rand.init_dof_pos_ratio_range = None
    c.domain_rand.init_base_pos_range = dict(x=[-0.1, 0.1], y=[-0.1, 0.1])
    c.domain_rand.init_npc_base_pos_range = None
    c.rewards.scales = Cfg(punishment_scale=1, success_reward_scale=10)
    c.viewer.update(pos=[0.0, 3.0, 5.0], lookat=[4.0, 3.0, 0.0])
    return c


def Go1BridgeCfg() -> Cfg:
    """go1_bridge_config.py:5-118: two robots facing each other across a 0.7 m wide deck between two platforms (bridge.urdf)."""
    c = go1_base()
    c.env.update(env_name="go1bridge", num_envs=1, num_agents=2, num_npcs=1, episode_length_s=20)
    c.asset.update(terminate_after_contacts_on=[], file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/bridge/urdf/bridge.urdf",
                   name_npc="bridge", fix_npc_base_link=True)
    c.terrain.update(num_rows=1, num_cols=1, BarrierTrack_kwargs=_track(
        options=["init", "wall", "plane", "wall"], randomize_obstacle_order=False, track_width=6,
        init=dict(block_length=0.5, room_size=(0.0, 0.0), border_width=0.0, offset=(0, 0)),
        plane=dict(block_length=10.0), wall=dict(block_length=0.1), wall_height=0.01))
    c.command.cfg.vel = True
    c.init_state.multi_init_state = True
    c.init_state.init_states = [InitState(pos=(2.0, 0.0, 1.4)), InitState(pos=(7.5, 0.0, 1.4), rot=(0.0, 0.0, 1.0, 0.0))]
    c.init_state.init_states_npc = [InitState(pos=(5.0, 0.0, 0.72))]
    c.termination.update(z_low_kwargs=dict(threshold=0.3))
    c.domain_rand.init_dof_pos_ratio_range = None
    c.domain_rand.init_base_pos_range = dict(x=[-0.1, 0.1], y=[-0.1, 0.1])
    c.domain_rand.init_npc_base_pos_range = None
    c.rewards.scales = Cfg(target_reward_scale=1, punishment_scale=1, success_reward_scale=10)
    c.viewer.update(pos=[0.0, 3.0, 5.0], lookat=[4.0, 3.0, 0.0])
    return c


def _football_game(num_agents, init_xy, episode_length_s):
    """go1_football_config.py:133-371 (1 vs 1 and 2 vs 2): free-play football, ball at (7, 0, 0.2)."""
    c = go1_base()
    c.env.update(env_name="go1football", num_envs=1, num_agents=num_agents, num_npcs=1, episode_length_s=episode_length_s)
    c.asset.update(file_npc="{LEGGED_GYM_ROOT_DIR}/resources/objects/ball.urdf", name_npc="ball",
                   terminate_after_contacts_on=[], npc_collision=True, fix_npc_base_link=False, npc_gravity=True)
    c.terrain.update(num_rows=1, num_cols=1, BarrierTrack_kwargs=_track(
        options=["init", "gate", "plane", "gate", "wall"], track_width=9.0,
        init=dict(block_length=1.0, room_size=(0.0, 0.0), border_width=0.0, offset=(0.5, 0)),
        plane=dict(block_length=10.0),
        gate=dict(block_length=1.0, width=2.0, depth=1.0, offset=(0, 0), random=(0, 0.0)),
        wall=dict(block_length=0.1), wall_height=1.0))
    c.command.cfg.vel = True
    c.init_state.multi_init_state = True
    c.init_state.init_states = [InitState(pos=(x, y, 0.42), rot=(0.0, 0.0, 0.0, 1.0) if x < 6 else (0.0, 0.0, 1.0, 0.0)) for x, y in init_xy]
    c.init_state.init_states_npc = [InitState(pos=(7.0, 0.0, 0.2))]
    c.termination.update(check_obstacle_conditioned_threshold=False, termination_terms=["roll", "pitch"])
    c.domain_rand.init_base_pos_range = dict(x=[-0.1, 0.1], y=[-0.1, 0.1])
    c.rewards.scales = Cfg(goal_reward_scale=1)
    c.viewer.update(pos=[2.0, 2.0, 2.0], lookat=[6.0, 5.0, 0.0])
    return c


def Go1Football1vs1Cfg() -> Cfg:
    """go1_football_config.py:133-250 (episode_length_s = 1 there)"""
    return _football_game(2, [(3.0, 0.0), (9.0, 0.0)], 1)


def Go1Football2vs2Cfg() -> Cfg:
    """go1_football_config.py:252-371"""
    return _football_game(4, [(3.0, 2.0), (3.0, -2.0), (9.0, 2.0), (9.0, -2.0)], 20)


def class_to_dict(obj):
    """helpers.py:46-61 for Cfg trees."""
    if isinstance(obj, Cfg):
        return {k: class_to_dict(v) for k, v in vars(obj).items() if not k.startswith("_")}
    if isinstance(obj, list):
        return [class_to_dict(v) for v in obj]
    return obj
