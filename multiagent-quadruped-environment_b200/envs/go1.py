"""The Go1 environment classes behind `make_mqe_env` -- same constructor, methods and tensor attributes as the
reference (`mqe/envs/go1/go1.py`, `mqe/envs/base/legged_robot.py`, `mqe/envs/field/legged_robot_field.py`,
`mqe/envs/npc/*.py`), with every per-step operation delegated to the CUDA engine through the C ABI.

What the reference does in Python per step (policy inference, 4 x actuator net + gym.simulate, post_physics_step's
~100 torch ops, NPC stepping, indexed resets) is four kernel launches here; this class only owns the zero-copy
views and the bookkeeping objects wrappers read (`obs_buf`, `reset_ids`, `extras`, ...).
"""
from __future__ import annotations

import copy
import ctypes

import numpy as np
import torch

from .. import engine as E
from ..scene import build_scene


class ObsBuf:
    """Attribute struct returned by Go1.step()/reset() (go1.py:26, 153-196): views into the engine's obs rows."""

    def __init__(self, obs: torch.Tensor, env_info: dict):
        self._obs = obs
        for name, (a, b) in E.OBS_SLICES.items():
            setattr(self, name, obs[:, a:b])
        self.env_info = env_info

    def as_tensor(self):
        return self._obs


class Go1:
    """Go1(cfg, sim_params, physics_engine, sim_device, headless) -- go1.py:20."""

    npc_is_scripted_agent = False

    def __init__(self, cfg, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True, *, seed=0,
                 env_slice=None, policy_mode=None, solver_iters=None):
        self.cfg = cfg
        self.env_name = cfg.env.env_name
        self.sim_params, self.physics_engine, self.headless = sim_params, physics_engine, headless
        if "cuda" not in str(sim_device):
            # BASELINE config C1 (`--sim_device cpu`, PhysX CPU pipeline) has no counterpart here: refusing is better than silently
            # running the CUDA engine under a "cpu" label.  The CPU restatement lives in oracle/ and is test infrastructure only.
            raise E.EngineError(f"sim_device={sim_device!r}: mqe_b200 has no CPU pipeline (sm_100a kernels only); pass sim_device='cuda:<i>'")
        if "FLEX" in str(physics_engine).upper():                           # gymapi.SIM_FLEX: the reference only ever configures PhysX
            raise E.EngineError(f"physics_engine={physics_engine!r}: only the PhysX-style rigid-body path is restated (cfg.sim.physx parameters)")
        dev = torch.device(sim_device)
        if not torch.cuda.is_available():
            raise E.EngineError("mqe_b200 needs a CUDA device (sm_100a); there is no CPU pipeline")
        self.device = dev
        self.device_index = dev.index or 0
        mode = E.POLICY_MODE_DEFAULT if policy_mode is None else policy_mode
        self.scene = build_scene(cfg, seed=seed, env_slice=env_slice, policy_mode=mode, solver_iters=solver_iters,
                                 wrapper_action_scale=(1.0, 1.0, 1.0))
        sc = self.scene
        self.num_envs, self.num_agents, self.num_npcs = sc.num_envs, sc.num_agents, sc.num_npcs
        self.num_actions = 12 * self.num_agents
        self.num_actions_npc = int(sc.desc.npc_dofs)
        self.dt = cfg.control.decimation * cfg.sim.dt
        self.decimation = cfg.control.decimation
        self.max_episode_length_s = cfg.env.episode_length_s
        self.max_episode_length = int(np.ceil(self.max_episode_length_s / self.dt))
        with torch.cuda.device(dev):
            self.engine = E.Engine(sc.desc, device=self.device_index, stream=torch.cuda.current_stream(dev).cuda_stream,
                                   keepalive=sc)
        self._bind_views()
        self.extras = {}
        self._reset_ids = torch.arange(self.num_envs, device=dev)
        self.gym = _GymFacade(self)
        self.sim = self.engine
        self.record_now = False
        self._ctrl_agents = self.num_agents - 1 if sc.desc.defender else self.num_agents
        self._scale_mode = None
        self.control_type = str(cfg.control.control_type)
        self._joint_control = int(sc.desc.control_type) != 0                # 'P' / 'V' / 'T': Go1.step takes [N, 12A] joint actions (go1.py:43-45)
        if sim_params is not None:                                          # gymapi.SimParams the reference hands over: must agree with cfg.sim
            dt = getattr(sim_params, "dt", None)
            if dt is not None and abs(float(dt) - float(cfg.sim.dt)) > 1e-9:
                raise E.EngineError(f"sim_params.dt = {dt} disagrees with cfg.sim.dt = {cfg.sim.dt}")

    # -- zero-copy views over the engine buffers (legged_robot.py:554-595) ------------------------------------
    def _bind_views(self):
        eng, N, A, P = self.engine, self.num_envs, self.num_agents, self.num_npcs
        t = eng.tensor
        dev = self.device
        self.all_root_states = t(E.BUF_ROOT_STATES).view(-1, 13)
        self.all_dof_states = t(E.BUF_DOF_STATES).view(-1, 2)
        env_root = t(E.BUF_ROOT_STATES)                                   # [N, A+P, 13]
        env_dof = t(E.BUF_DOF_STATES)                                     # [N, 12A+D, 2]
        self._env_root = env_root                                          # root_states / root_states_npc: properties below
        self.dof_state = env_dof[:, :12 * A, :]
        self.dof_pos = env_dof[:, :12 * A, 0]
        self.dof_vel = env_dof[:, :12 * A, 1]
        self.dof_state_npc = env_dof[:, 12 * A:, :]
        self.contact_forces = t(E.BUF_CONTACT_FORCES)
        self.torques = t(E.BUF_TORQUES)
        self.actions = t(E.BUF_ACTIONS)
        self.last_actions = t(E.BUF_LAST_ACTIONS)
        self.base_lin_vel = t(E.BUF_BASE_LIN_VEL)
        self.base_ang_vel = t(E.BUF_BASE_ANG_VEL)
        self.projected_gravity = t(E.BUF_PROJ_GRAVITY)
        self.commands = t(E.BUF_COMMANDS)
        self.locomotion_obs = t(E.BUF_LOC_OBS)
        self.last_locomotion_action = t(E.BUF_LOC_ACTION)
        self.gait_indices = t(E.BUF_GAIT)
        self.episode_length_buf = t(E.BUF_EPISODE_LENGTH)
        self._reset_u8 = t(E.BUF_RESET)
        self.reset_buf = self._reset_u8.view(torch.bool)
        self.time_out_buf = t(E.BUF_TIMEOUT).view(torch.bool)
        self.collide_buf = t(E.BUF_COLLIDE).view(torch.bool)
        self.r_term_buff = t(E.BUF_ROLL_TERM).view(torch.bool)
        self.p_term_buff = t(E.BUF_PITCH_TERM).view(torch.bool)
        self.z_low_term_buff = t(E.BUF_ZLOW_TERM).view(torch.bool)
        self.z_high_term_buff = t(E.BUF_ZHIGH_TERM).view(torch.bool)
        self.rew_buf = torch.zeros(N * A, device=dev)                      # Go1 registers no reward terms (go1.py:198-219)
        sc = self.scene
        self.env_origins = torch.as_tensor(sc.env_origins, device=dev)
        self.agent_origins = torch.as_tensor(sc.agent_origins, device=dev)
        self.env_origins_repeat = self.env_origins.unsqueeze(1).repeat(1, A, 1).reshape(-1, 3)
        self.base_init_state = torch.as_tensor(sc.base_init_state, device=dev)
        self.env_agent_indices = torch.arange(N * A, device=dev).view(N, A)
        self.env_npc_indices = torch.arange(N * P, device=dev).view(N, P) if P else torch.zeros(N, 0, dtype=torch.long, device=dev)
        ids = torch.arange(N * (A + P), device=dev, dtype=torch.int32).view(N, A + P)
        self.agent_indices = ids[:, :A].reshape(-1)
        self.npc_indices = ids[:, A:].reshape(-1)
        self.actor_indices = ids.reshape(-1)
        if P:
            self.npc_env_origins = self.env_origins.unsqueeze(1).repeat(1, P, 1)
            self.base_init_state_npc = torch.as_tensor(sc.npc_init_state, device=dev)
        stats = t(E.BUF_SHEEP_STATS)
        self.sheep_pos_avg = stats[:, :2]
        self.sheep_pos_var = stats[:, 2]
        self.env_info = {k: torch.as_tensor(np.ascontiguousarray(v), device=dev) for k, v in sc.env_info.items()}
        if sc.desc.defender:                                               # go1_football_defender.py:61-63
            self.gate_pos = self.env_origins.clone()
            self.gate_pos[:, 0] += float(sc.desc.gate_x)
        self.obs_buf = ObsBuf(t(E.BUF_OBS), self.env_info if self.cfg.obs.cfgs.env_info else {})
        self.privileged_obs_buf = copy.copy(self.cfg.privileged_obs)
        self.base_quat = t(E.BUF_OBS)[:, 3:7]

    # root_states is a view when P == 0 and a copy otherwise in the reference too (legged_robot.py:130, 136)
    @property
    def root_states(self):
        return self._env_root[:, :self.num_agents, :].reshape(-1, 13)

    @property
    def root_states_npc(self):
        return self._env_root[:, self.num_agents:, :].reshape(-1, 13)

    @property
    def base_pos(self):
        return self.root_states[:, 0:3]

    @property
    def reset_ids(self):
        """`reset_buf.nonzero().flatten()` (legged_robot.py:145), evaluated only when a caller asks: the step
        itself never synchronises the host."""
        if self._reset_ids is None:
            self._reset_ids = self._reset_u8.nonzero(as_tuple=False).flatten()
        return self._reset_ids

    # -- reference API ------------------------------------------------------------------------------------------
    def _set_scale(self, mode):
        if self._scale_mode != mode:
            self.engine.set_action_scale((2.0, 0.5, 0.5) if mode == "wrapper" else (1.0, 1.0, 1.0))
            self._scale_mode = mode

    def reset(self):
        """go1.py:147-151: reset_idx(all) + compute_observations; no physics step."""
        self.engine.reset()
        self._reset_ids = torch.arange(self.num_envs, device=self.device)
        self._fill_extras()
        return self.obs_buf

    def _fill_extras(self):
        """legged_robot.py:1063-1076 as it comes out for Go1: no reward terms are registered (go1.py:198-219), so `extras["episode"]`
        is empty; `extras["time_outs"]` aliases the time-out flags (cfg.env.send_timeouts), which the engine rewrites in place."""
        self.extras["episode"] = {}
        if getattr(self.cfg.env, "send_timeouts", False):
            self.extras["time_outs"] = self.time_out_buf

    def step(self, action):
        """go1.py:35-62.  control_type 'C': action [N*A_ctrl, 3] already scaled by the task wrapper.  'P' / 'V' / 'T' (go1.py:43-45):
        action reshapes to [N, 12A] joint actions, clipped to +-clip_actions by pre_physics_step; the walk policy is not used."""
        if self._joint_control:
            a = action.reshape(self.num_envs, -1)
            if a.dtype != torch.float32 or not a.is_contiguous() or a.device != self.device:
                a = a.to(device=self.device, dtype=torch.float32).contiguous()
            assert a.shape[1] == self.num_actions, f"expected [{self.num_envs}, {self.num_actions}] joint actions, got {tuple(action.shape)}"
            self.engine.step_joint(a.data_ptr())
            self._last_action_ref = a
            self._reset_ids = None
            return self.obs_buf, self.rew_buf, self.reset_buf, self.extras
        self._set_scale("env")
        return self._step(action)

    def step_from_wrapper(self, action):
        """Fused entry for the task wrappers: raw [N, A_ctrl, 3] policy actions; the clip / [2,.5,.5] scale of
        `wrappers/*.py step()` happens inside the frame kernel."""
        self._set_scale("wrapper")
        return self._step(action)

    def _step(self, action):
        a = action
        if a.dtype != torch.float32 or not a.is_contiguous() or a.device != self.device:
            a = a.to(device=self.device, dtype=torch.float32).contiguous()
        assert a.numel() == self.num_envs * self._ctrl_agents * 3, f"expected {self.num_envs}x{self._ctrl_agents}x3 actions, got {tuple(action.shape)}"
        self.engine.step(a.data_ptr())
        self._last_action_ref = a                                          # keep alive until the kernels have read it
        self._reset_ids = None
        return self.obs_buf, self.rew_buf, self.reset_buf, self.extras

    # post_decimation_step logs (legged_robot.py:112-115).  Nothing on the Go1 path reads them, so the engine only starts writing them
    # once somebody asks: the first access allocates them (zeros, as in _init_buffers :623-625) and every following step fills them.
    @property
    def substep_torques(self):
        return self.engine.tensor(E.BUF_SUBSTEP_TORQUES)

    @property
    def substep_dof_vel(self):
        return self.engine.tensor(E.BUF_SUBSTEP_DOF_VEL)

    @property
    def substep_exceed_dof_pos_limits(self):
        return self.engine.tensor(E.BUF_SUBSTEP_EXCEED).view(torch.bool)

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def render(self, sync_frame_time=True):
        return None

    def start_recording(self):
        self.record_now = True

    def pause_recording(self):
        self.record_now = False

    def get_complete_frames(self):
        return []

    def close(self):
        self.engine.close()


class Go1Sheep(Go1):
    """go1_sheep.py: the flocking step runs inside the post-physics kernel."""


class Go1Object(Go1):
    """go1_object.py: passive NPC object (ball / seesaw)."""


class Go1FootballDefender(Go1Object):
    """go1_football_defender.py: the third agent's command comes from the scripted controller in the frame kernel."""


class _GymFacade:
    """The handful of `env.gym.*` calls wrappers make directly (go1_tug_wrapper.py:67-69)."""

    def __init__(self, env):
        self._env = env

    def set_actor_root_state_tensor_indexed(self, sim, root_states, actor_ids, n):
        self._env.engine.set_root_indexed(root_states, actor_ids[:n])

    def set_dof_state_tensor_indexed(self, sim, dof_states, actor_ids, n):
        self._env.engine.set_dof_indexed(dof_states, actor_ids[:n])

    def set_actor_root_state_tensor(self, sim, root_states):
        ids = self._env.actor_indices
        self._env.engine.set_root_indexed(root_states, ids)

    def refresh_actor_root_state_tensor(self, sim):
        return True

    refresh_dof_state_tensor = refresh_net_contact_force_tensor = refresh_rigid_body_state_tensor = refresh_actor_root_state_tensor

    def simulate(self, sim):
        self._env.engine.substeps(1)

    def fetch_results(self, sim, wait=True):
        if wait:
            self._env.engine.synchronize()
