"""Task wrappers with the reference's observation / reward definitions (`mqe/envs/wrappers/*.py`).

Differences from the reference that do not change any returned value:
  * action clip and the [2, .5, .5] scale are applied inside the engine's frame kernel (`Go1.step_from_wrapper`);
  * `reward_buffer[...]` accumulates device scalars instead of calling `.cpu()` per term per step
    (e.g. go1_sheep_wrapper.py:77,83,93,105,112) -- `float(v)` / `v.cpu()` still works for loggers;
  * the device is the env's device instead of a hard-coded "cuda".
"""
from __future__ import annotations

import os

import torch

from .gym_shim import Wrapper, spaces


class EmptyWrapper(Wrapper):
    """empty_wrapper.py:4-17"""

    def __init__(self, env):
        super().__init__(env)
        self.num_envs = self.env.num_envs
        self.num_agents = self.env.num_agents
        if hasattr(env.cfg.terrain, "BarrierTrack_kwargs"):
            self.BarrierTrack_kwargs = env.cfg.terrain.BarrierTrack_kwargs
        for key in dir(self.env.cfg.rewards.scales):
            if key[0] != "_" and "scale" in key:
                setattr(self, key, getattr(self.env.cfg.rewards.scales, key))
        self._fuse_allowed = os.environ.get("MQE_FUSED_WRAPPERS", "1") != "0"
        dev = self.env.device
        self.obs_ids = torch.eye(self.num_agents, dtype=torch.float32, device=dev).repeat(self.num_envs, 1).reshape(
            self.num_envs, self.num_agents, -1)

    def _acc(self, key, value):
        self.reward_buffer[key] = self.reward_buffer[key] + value

    def _base_info(self, obs_buf):
        return torch.cat([obs_buf.base_pos, obs_buf.base_rpy], dim=1).reshape(self.env.num_envs, self.env.num_agents, -1)

    # -- fused gather (csrc/wrapper.cu): obs / reward / reward_buffer sums computed by one kernel inside the step graph ----------
    def _fuse(self, kind, scales, keys, gate=None):
        """Switch this wrapper to the engine's fused gather.  Only on a real env (an `engine` behind it); the torch code below
        stays the definition (pinned against the reference on the CPU, compared with the kernel on the GPU)."""
        eng = getattr(self.env, "engine", None)
        if eng is None or not self._fuse_allowed:
            return False
        from .. import engine as E
        eng.set_wrapper(kind, scales, None if gate is None else gate.detach().cpu().numpy())
        self._halves = eng.result_views()                  # (obs, reward, done) views of the two halves of the packed step result
        self.reward_buffer = _FusedRewardBuffer(keys, eng.tensor(E.BUF_WRAP_SUMS))
        eng.wrapper_reset()
        return True

    @property
    def _wobs(self):
        """Observation of the latest step / reset: a zero-copy view.  The engine alternates between two result buffers, so a returned
        observation stays intact across the NEXT step() (the reference returns fresh tensors; a rollout loop may hold obs_t while
        stepping) and is overwritten by the one after."""
        return self._halves[self.env.engine.result_parity()][0]

    def _fused_step(self, action):
        _, _, termination, info = self.env.step_from_wrapper(action)
        obs, reward, _ = self._halves[self.env.engine.result_parity()]
        return obs, reward, termination, info

    def step_host(self, h_actions, h_result):
        """Host-buffer step for the numpy adapter (openrl_ws/utils.py:53-67): pinned H2D of the raw [N, A, 3] actions, one step, ONE D2H of
        the packed (obs | reward | done) result.  Fused wrappers only."""
        base = self.env
        base._set_scale("wrapper")
        base.engine.step_host_result(h_actions, h_result)
        base._reset_ids = None


class _FusedRewardBuffer(dict):
    """`reward_buffer` backed by the kernel's running sums: reads are lazy 0-dim tensors (`float(v)` / `v.cpu()` work as on the
    reference's values), assigning a key (loggers zero them) re-bases it."""

    def __init__(self, key_to_index, sums):
        super().__init__({k: 0 for k in key_to_index})
        self._idx, self._sums = dict(key_to_index), sums
        self._off = {k: 0.0 for k in key_to_index}

    def __getitem__(self, k):
        return self._sums[self._idx[k]] + self._off[k]

    def __setitem__(self, k, v):
        if k not in self._idx:
            raise KeyError(k)
        self._off[k] = float(v) - float(self._sums[self._idx[k]])

    def get(self, k, default=None):
        return self[k] if k in self._idx else default

    def items(self):
        return [(k, self[k]) for k in self._idx]

    def values(self):
        return [self[k] for k in self._idx]


class Go1GateWrapper(EmptyWrapper):
    """go1_gate_wrapper.py.  The shipped reference returns `obs = 0, reward = 0` (its 16-D observation and 8 reward
    terms are commented out, :68, :155); `legacy_zero_outputs=False` enables the intended definitions (:78-154)."""

    def __init__(self, env, legacy_zero_outputs=True):
        super().__init__(env)
        self.legacy_zero_outputs = legacy_zero_outputs
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(14 + self.num_agents,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.reward_buffer = {"target reward": 0, "success reward": 0, "agent distance punishment": 0,
                              "contact punishment": 0, "step count": 0}
        self.gate_pos = None

    def _init_extras(self, obs):
        if self.legacy_zero_outputs:
            return
        kw = self.BarrierTrack_kwargs
        gate = obs.env_info["gate_deviation"].clone()
        gate[:, 0] += kw["init"]["block_length"] + kw["gate"]["block_length"] / 2
        self.gate_pos = gate.unsqueeze(1).repeat(1, self.num_agents, 1)
        self.gate_distance = self.gate_pos.reshape(-1, 2)[:, 0]
        self.target_pos = torch.zeros_like(self.gate_pos)
        self.target_pos[:, :, 0] = kw["init"]["block_length"] + kw["gate"]["block_length"] + kw["plane"]["block_length"] / 2
        self.target_pos[:, 0, 1] = kw["track_width"] / 4
        self.target_pos[:, 1, 1] = -kw["track_width"] / 4
        self.target_pos = self.target_pos.reshape(-1, 2)

    def _obs(self, obs_buf):
        base_info = self._base_info(obs_buf)
        return torch.cat([self.obs_ids, base_info, torch.flip(base_info, [1]), self.gate_pos], dim=2)

    def reset(self):
        obs_buf = self.env.reset()
        if self.gate_pos is None:
            self._init_extras(obs_buf)
        return 0 if self.legacy_zero_outputs else self._obs(obs_buf)

    def step(self, action):
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        if self.gate_pos is None:
            self._init_extras(obs_buf)
        if self.legacy_zero_outputs:
            return 0, 0, termination, info
        obs = self._obs(obs_buf)
        self._acc("step count", 1)
        reward = torch.zeros([self.env.num_envs, self.env.num_agents], device=self.env.device)
        base_pos = obs_buf.base_pos
        if self.target_reward_scale != 0:                                   # :87-101
            distance_to_target = torch.norm(base_pos[:, :2] - self.target_pos, p=2, dim=1)
            if not hasattr(self, "last_distance_to_target"):
                self.last_distance_to_target = distance_to_target.clone()
            target_reward = (self.last_distance_to_target - distance_to_target).reshape(self.num_envs, -1).sum(dim=1, keepdim=True)
            target_reward[self.env.reset_buf] = 0
            target_reward *= self.target_reward_scale
            reward += target_reward.repeat(1, self.env.num_agents)
            self.last_distance_to_target = distance_to_target.clone()
            self._acc("target reward", torch.sum(target_reward))
        if self.contact_punishment_scale != 0:                              # :103-107
            collide_reward = self.contact_punishment_scale * self.env.collide_buf
            reward += collide_reward.unsqueeze(1).repeat(1, self.num_agents)
            self._acc("contact punishment", torch.sum(collide_reward))
        if self.success_reward_scale != 0:                                  # :109-114
            success_reward = torch.zeros([self.env.num_envs * self.env.num_agents], device=self.env.device)
            success_reward[base_pos[:, 0] > self.gate_distance + 0.25] = self.success_reward_scale
            reward += success_reward.reshape([self.env.num_envs, self.env.num_agents])
            self._acc("success reward", torch.sum(success_reward))
        if self.agent_distance_punishment_scale != 0:                       # :128-134
            agent_dis = (base_pos[:, :2] - torch.flip(base_pos[:, :2].reshape(self.num_envs, self.num_agents, 2), dims=[1]).reshape(-1, 2)) ** 2
            agent_dis = agent_dis.sum(dim=1).reshape(self.num_envs, -1)
            pun = torch.where(agent_dis < 0.25, self.agent_distance_punishment_scale / agent_dis.clamp_min(1e-12), torch.zeros_like(agent_dis))
            reward += pun
            self._acc("agent distance punishment", torch.sum(pun))
        reward = reward.sum(dim=1).unsqueeze(1).repeat(1, self.num_agents)  # :154
        return obs, reward, termination, info


class Go1SheepWrapper(EmptyWrapper):
    """go1_sheep_wrapper.py:8-118"""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"),
                                            shape=(14 + 2 * self.cfg.env.num_npcs + self.num_agents,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.reward_buffer = {"success reward": 0, "contact punishment": 0, "sheep movement reward": 0,
                              "mixed sheep reward": 0, "sheep pos var punishment": 0, "step count": 0}
        self.gate_pos = None
        self.last_sheep_pos_avg = None

    def _init_extras(self, obs):
        kw = self.BarrierTrack_kwargs
        gate_pos = obs.env_info["gate_deviation"]                           # mutated in place, as the reference does (:31-32)
        gate_pos[:, 0] += kw["init"]["block_length"] + kw["plane"]["block_length"] + kw["gate"]["block_length"] / 2
        self.gate_pos = gate_pos.unsqueeze(1)
        self.gate_distance = gate_pos[:, 0].unsqueeze(1).repeat(1, self.num_npcs)

    def _obs(self, obs_buf):
        sheep_pos = self.root_states_npc[:, :3].reshape(self.num_envs, -1, 3) - self.npc_env_origins
        flat = sheep_pos[..., :2].reshape(self.num_envs, 1, -1).repeat(1, self.num_agents, 1)
        base_info = self._base_info(obs_buf)
        obs = torch.cat([self.obs_ids, base_info, torch.flip(base_info, [1]), self.gate_pos.repeat(1, self.num_agents, 1), flat], dim=2)
        return obs, sheep_pos

    def reset(self):
        obs_buf = self.env.reset()
        if self.gate_pos is None:
            self._init_extras(obs_buf)
            s = self
            self._fused = self._fuse(1, [s.success_reward_scale, s.contact_punishment_scale, s.sheep_movement_reward_scale, s.mixed_sheep_reward_scale,
                                         s.sheep_pos_var_lin_punishment_scale, s.sheep_pos_var_exp_punishment_scale],
                                     {"success reward": 0, "contact punishment": 1, "sheep movement reward": 2, "mixed sheep reward": 3,
                                      "sheep pos var punishment": 4, "step count": 8}, gate=self.gate_pos[:, 0, :])
        if getattr(self, "_fused", False):
            return self._wobs
        obs, _ = self._obs(obs_buf)
        self.last_sheep_pos_avg = None
        return obs

    def step(self, action):
        if getattr(self, "_fused", False):
            return self._fused_step(action)
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        if self.gate_pos is None:
            self._init_extras(obs_buf)
        obs, sheep_pos = self._obs(obs_buf)
        self._acc("step count", 1)
        reward = torch.zeros([self.env.num_envs, 1], device=self.env.device)
        if self.success_reward_scale != 0:
            success_reward = ((sheep_pos[:, :, 0] - self.gate_distance) > 0).sum(dim=1)
            reward[:, 0] = success_reward
            self._acc("success reward", torch.sum(success_reward))
        if self.contact_punishment_scale != 0:
            collide_reward = self.contact_punishment_scale * self.env.collide_buf
            reward += collide_reward.unsqueeze(1)
            self._acc("contact punishment", torch.sum(collide_reward))
        if self.sheep_movement_reward_scale != 0:
            if self.last_sheep_pos_avg is not None:
                x_movement = (self.sheep_pos_avg - self.last_sheep_pos_avg)[:, 0]
                x_movement[self.delayed_reset_buf] = 0
                sheep_movement_reward = self.sheep_movement_reward_scale * x_movement
                reward[:, 0] += sheep_movement_reward
                self._acc("sheep movement reward", torch.sum(sheep_movement_reward))
            self.last_sheep_pos_avg = self.sheep_pos_avg.clone()
        if self.mixed_sheep_reward_scale != 0:
            distance_to_gate = torch.norm(sheep_pos[..., :-1] - self.gate_pos.repeat(1, self.num_npcs, 1), dim=-1)
            mixed = torch.exp(-distance_to_gate / 2) * self.mixed_sheep_reward_scale
            mixed[sheep_pos[..., 0] >= self.gate_distance] = self.mixed_sheep_reward_scale
            reward[:, 0] += mixed.sum(dim=-1)
            self._acc("mixed sheep reward", torch.sum(mixed))
        if self.sheep_pos_var_exp_punishment_scale != 0 or self.sheep_pos_var_lin_punishment_scale != 0:
            pun = self.sheep_pos_var_lin_punishment_scale * (self.sheep_pos_var - 1) + \
                self.sheep_pos_var_exp_punishment_scale * torch.exp(self.sheep_pos_var / 2 - 1)
            reward[:, 0] += pun
            self._acc("sheep pos var punishment", torch.sum(pun))
        reward = reward.repeat(1, self.num_agents)
        self.delayed_reset_buf = self.env.reset_buf.clone()                 # mask form of copy(self.env.reset_ids) (:116)
        return obs, reward, termination, info


class Go1SeesawWrapper(EmptyWrapper):
    """go1_seesaw_wrapper.py:8-120"""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(12 + self.num_agents,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.reward_buffer = {"height reward": 0, "contact punishment": 0, "x movement reward": 0, "y punishment": 0,
                              "agent distance punishment": 0, "success reward": 0, "fall punishment": 0, "step count": 0}

    def _obs(self, obs_buf):
        base_info = self._base_info(obs_buf)
        return torch.cat([self.obs_ids, base_info, torch.flip(base_info, [1])], dim=2)

    def reset(self):
        obs_buf = self.env.reset()
        if not hasattr(self, "_fused"):
            s = self
            self._fused = self._fuse(2, [s.x_movement_reward_scale, s.height_reward_scale, s.y_punishment_scale, s.contact_punishment_scale,
                                         s.agent_distance_punishment_scale, s.success_reward_scale, s.fall_punishment_scale],
                                     {"x movement reward": 0, "height reward": 1, "y punishment": 2, "contact punishment": 3,
                                      "agent distance punishment": 4, "success reward": 5, "fall punishment": 6, "step count": 8})
        return self._wobs if self._fused else self._obs(obs_buf)

    def step(self, action):
        if getattr(self, "_fused", False):
            return self._fused_step(action)
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        obs = self._obs(obs_buf)
        base_pos = obs_buf.base_pos
        self._acc("step count", 1)
        reward = torch.zeros([self.env.num_envs, 1], device=self.env.device)
        if self.x_movement_reward_scale != 0:
            x_pos = base_pos[:, 0].reshape(self.num_envs, -1)
            if not hasattr(self, "last_x_pos"):
                self.last_x_pos = x_pos.clone()
            x_reward = (x_pos - self.last_x_pos).sum(dim=1, keepdim=True)
            x_reward[self.env.reset_buf] = 0
            x_reward *= self.x_movement_reward_scale
            reward += x_reward
            self.last_x_pos = x_pos.clone()
            self._acc("x movement reward", torch.sum(x_reward))
        if self.height_reward_scale != 0:
            height_reward = self.height_reward_scale * (base_pos[:, 2].reshape(self.num_envs, -1).sum(dim=1) - 0.56)
            reward[:, 0] += height_reward
            self._acc("height reward", torch.sum(height_reward))
        if self.y_punishment_scale != 0:
            y_punishment = self.y_punishment_scale * ((base_pos[:, 1].reshape(self.num_envs, -1) ** 2).sum(dim=1) - 0.5)
            reward[:, 0] += y_punishment
            self._acc("y punishment", torch.sum(y_punishment))
        if self.contact_punishment_scale != 0:
            collide_reward = self.contact_punishment_scale * self.env.collide_buf
            reward += collide_reward.unsqueeze(1)
            self._acc("contact punishment", torch.sum(collide_reward))
        if self.agent_distance_punishment_scale != 0:
            agent_dis = (base_pos[:, :2] - torch.flip(base_pos[:, :2].reshape(self.num_envs, self.num_agents, 2), dims=[1]).reshape(-1, 2)) ** 2
            agent_dis = agent_dis.sum(dim=1).reshape(self.num_envs, -1)[:, :1]
            close = agent_dis < 0.25
            pun = torch.where(close, self.agent_distance_punishment_scale / agent_dis.clamp_min(1e-12), torch.zeros_like(agent_dis))
            reward += pun
            self._acc("agent distance punishment", torch.sum(pun))
        if self.success_reward_scale != 0:
            success = (base_pos[:, 0] > 7.7) * (base_pos[:, 2] > 1.3)
            success_reward = self.success_reward_scale * success.reshape(self.num_envs, -1).sum(dim=1)
            reward[:, 0] += success_reward
            self._acc("success reward", torch.sum(success_reward))
        if self.fall_punishment_scale != 0:
            fall = self.env.r_term_buff | self.env.p_term_buff
            reward[fall, 0] += self.fall_punishment_scale
            self._acc("fall punishment", self.fall_punishment_scale * torch.sum(fall))
        return obs, reward.repeat(1, self.num_agents), termination, info


class Go1FootballDefenderWrapper(EmptyWrapper):
    """go1_football_wrapper.py:8-91: two controlled agents; the env's third agent is the scripted defender."""

    def __init__(self, env):
        super().__init__(env)
        self.num_agents = 2
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(18 + self.num_agents,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.obs_ids = torch.eye(2, dtype=torch.float32, device=self.env.device).repeat(self.num_envs, 1).reshape(self.num_envs, 2, -1)
        self.reward_buffer = {"goal reward": 0, "ball gate distance reward": 0, "step count": 0}

    def _obs(self, obs_buf):
        npc = self.root_states_npc
        ball_pos = (npc[:, :3].reshape(self.num_envs, 3) - self.env_origins).unsqueeze(1).repeat(1, 2, 1)
        ball_vel = npc[:, 7:10].reshape(self.num_envs, 3).unsqueeze(1).repeat(1, 2, 1)
        base_info = self._base_info(obs_buf)[:, :2, :]
        return torch.cat([self.obs_ids, base_info, torch.flip(base_info, [1]), ball_pos, ball_vel], dim=2), ball_pos

    def reset(self):
        obs_buf = self.env.reset()
        if not hasattr(self, "_fused"):
            self._fused = self._fuse(3, [self.goal_reward_scale, self.ball_gate_distance_reward_scale],
                                     {"goal reward": 0, "ball gate distance reward": 1, "step count": 8}, gate=self.gate_pos)
        return self._wobs if self._fused else self._obs(obs_buf)[0]

    def step(self, action):
        if getattr(self, "_fused", False):
            return self._fused_step(action)
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        obs, ball_pos = self._obs(obs_buf)
        self._acc("step count", 1)
        reward = torch.zeros([self.env.num_envs, 1], device=self.env.device)
        # NOTE the reference compares the env-relative ball x with the WORLD gate x (go1_football_wrapper.py:77)
        if self.goal_reward_scale != 0:
            goal_reward = torch.zeros_like(reward)
            goal_reward[ball_pos[:, 0, 0] > self.gate_pos[:, 0], 0] = self.goal_reward_scale
            reward += goal_reward
            self._acc("goal reward", torch.sum(goal_reward))
        if self.ball_gate_distance_reward_scale != 0:
            d = torch.norm(ball_pos[:, 0, :2] - self.gate_pos[:, :2], dim=1, keepdim=True)
            r = self.ball_gate_distance_reward_scale * torch.exp(-d / 3)
            reward += r
            self._acc("ball gate distance reward", torch.sum(r))
        return obs, reward.repeat(1, 2), termination, info


class Go1FootballGameWrapper(EmptyWrapper):
    """go1_football_wrapper.py:93-157 (1 vs 1 / 2 vs 2).  The reference leaves this wrapper unfinished: `reset()` returns
    None, `step()` returns `None` observations and zero rewards repeated 4 times (:128, :157); reproduced as is, with the
    quantities it computes on the way (`ball_pos`, `ball_vel`, `base_info`) kept as attributes for callers."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(18 + self.num_agents,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.reward_buffer = {"goal reward": 0, "step count": 0}

    def _gather(self, obs_buf):
        npc = self.root_states_npc
        self.ball_pos = (npc[:, :3].reshape(self.num_envs, 3) - self.env_origins).unsqueeze(1).repeat(1, 2, 1)
        self.ball_vel = npc[:, 7:10].reshape(self.num_envs, 3).unsqueeze(1).repeat(1, 2, 1)
        self.base_info = self._base_info(obs_buf)[:, :2, :]

    def reset(self):
        self._gather(self.env.reset())
        return None

    def step(self, action):
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        self._gather(obs_buf)
        self._acc("step count", 1)
        reward = torch.zeros([self.env.num_envs, 1], device=self.env.device)
        return None, reward.repeat(1, 4), termination, info


class Go1PushboxWrapper(EmptyWrapper):
    """go1_pushbox_wrapper.py: obs = ids, (pos, rpy) self / other, gate xy, box xy, box quaternion (20 + A); reward = box x progress."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(20 + self.num_agents,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.box_x_movement_reward_scale = 1                 # the reference overrides the config value (10) here (:17)
        self.reward_buffer = {"box movement reward": 0, "step count": 0}
        self.gate_pos = None
        self.last_box_pos = None

    def _init_extras(self, obs):
        kw = self.BarrierTrack_kwargs
        gate = obs.env_info["gate_deviation"]                # mutated in place, as the reference does
        gate[:, 0] += kw["init"]["block_length"] + kw["gate"]["block_length"] / 2
        self.gate_pos = gate.unsqueeze(1).repeat(1, self.num_agents, 1)
        self.gate_distance = self.gate_pos.reshape(-1, 2)[:, 0]

    def _obs(self, obs_buf):
        npc = self.root_states_npc
        box_pos = npc[:, :3] - self.env.env_origins
        base_info = self._base_info(obs_buf)
        obs = torch.cat([self.obs_ids, base_info, torch.flip(base_info, [1]), self.gate_pos,
                         box_pos[:, :2].unsqueeze(1).repeat(1, self.num_agents, 1),
                         npc[:, 3:7].unsqueeze(1).repeat(1, self.num_agents, 1)], dim=2)
        return obs, box_pos

    def reset(self):
        obs_buf = self.env.reset()
        if self.gate_pos is None:
            self._init_extras(obs_buf)
            self._fused = self._fuse(4, [self.box_x_movement_reward_scale], {"box movement reward": 0, "step count": 8}, gate=self.gate_pos[:, 0, :])
        if getattr(self, "_fused", False):
            return self._wobs
        self.last_box_pos = None
        return self._obs(obs_buf)[0]

    def step(self, action):
        if getattr(self, "_fused", False):
            return self._fused_step(action)
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        if self.gate_pos is None:
            self._init_extras(obs_buf)
        obs, box_pos = self._obs(obs_buf)
        self._acc("step count", 1)
        reward = torch.zeros([self.env.num_envs, 1], device=self.env.device)
        if self.box_x_movement_reward_scale != 0:
            if self.last_box_pos is not None:
                x_movement = (box_pos - self.last_box_pos)[:, 0]
                x_movement[self.env.reset_buf] = 0
                r = self.box_x_movement_reward_scale * x_movement
                reward[:, 0] += r
                self._acc("box movement reward", torch.sum(r))
        self.last_box_pos = box_pos.clone()
        return obs, reward.repeat(1, self.num_agents), termination, info


class Go1RotationWrapper(EmptyWrapper):
    """go1_rotation_wrapper.py:8-103 (revolving door): obs = (pos, rpy) self / other with agent 1 mirrored in y (12);
    reward on agent 0 only: success when agent 0 is past the door, punishment when agent 1 is, +1 whenever agent 0's
    distance to the target shrank.  Returned reward is [N, A, 1] as in the reference.

    Reference quirk kept: step() subtracts the (per-env, all equal) target x from BOTH x and y of every agent
    (`dis[:, :] -= self.target_pos`, :77), which only broadcasts for num_envs <= 2 there; here the same arithmetic is
    applied for any num_envs.  reset() (:38-39) subtracts it from x only."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(12,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.success_reward_scale = 5
        self.distance_reward_scale = 1
        self.punishment_scale = 1
        self.reward_buffer = {"punishment": 0, "success reward": 0, "distance reward": 0, "step count": 0}

    def _init_extras(self, obs):
        kw = self.BarrierTrack_kwargs
        self.target_pos = torch.full((self.env.num_envs,), kw["rotation"]["block_length"] * 0.75 + kw["wall"]["block_length"],
                                     dtype=torch.float, device=self.env.device)
        dis = obs.base_pos.reshape(self.env.num_envs, self.env.num_agents, -1)[:, :, :2].clone()
        dis[:, :, 0] -= self.target_pos.unsqueeze(1)
        self.last_dis = dis.norm(p=2, dim=-1)

    def _obs(self, obs_buf):
        base_info = self._base_info(obs_buf)
        obs = torch.cat([base_info, torch.flip(base_info, [1])], dim=2)
        obs[:, 1, 1::3] = -obs[:, 1, 1::3]                    # y and pitch of both halves, as seen by the mirrored agent
        return obs

    def reset(self):
        obs_buf = self.env.reset()
        self._init_extras(obs_buf)
        if not hasattr(self, "_fused"):
            self._fused = self._fuse(7, [self.success_reward_scale, self.punishment_scale, self.distance_reward_scale, float(self.target_pos[0])],
                                     {"success reward": 0, "punishment": 1, "distance reward": 2, "step count": 8})
        return self._wobs if self._fused else self._obs(obs_buf)

    def step(self, action):
        action[:, 1, 1:] = -action[:, 1, 1:]                  # in place, as the reference does; commutes with the clip
        if getattr(self, "_fused", False):
            obs, reward, termination, info = self._fused_step(action)
            return obs, reward.reshape(self.env.num_envs, self.env.num_agents, 1), termination, info
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        N, A = self.env.num_envs, self.env.num_agents
        base_pos = obs_buf.base_pos.reshape(N, A, -1)
        reward = torch.zeros([N, A], device=self.env.device)
        if self.success_reward_scale != 0:
            success = (base_pos[:, 0, 0] > self.target_pos) * float(self.success_reward_scale)
            reward[:, 0] += success
            self._acc("success reward", torch.sum(success))
        if self.punishment_scale != 0:
            punishment = (base_pos[:, 1, 0] > self.target_pos) * float(self.punishment_scale)
            reward[:, 0] -= punishment
            self._acc("punishment", torch.sum(punishment))
        if self.distance_reward_scale != 0:
            dis = (base_pos[:, :, :2] - self.target_pos.view(N, 1, 1)).norm(p=2, dim=-1)
            dis_reward = (dis[:, 0] < self.last_dis[:, 0]) * float(self.distance_reward_scale)
            reward[:, 0] += dis_reward
            self._acc("distance reward", torch.sum(dis_reward))
            self.last_dis = dis
        self._acc("step count", 1)
        return self._obs(obs_buf), reward.reshape(N, A, 1), termination, info


def get_euler_xyz(q):
    """isaacgym.torch_utils.get_euler_xyz (xyzw; every angle returned in [0, 2 pi)) -- SURVEY appendix B."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    roll = torch.atan2(2.0 * (w * x + y * z), w * w - x * x - y * y + z * z)
    sinp = 2.0 * (w * y - z * x)
    pitch = torch.where(torch.abs(sinp) >= 1, torch.copysign(torch.full_like(sinp, torch.pi / 2.0), sinp), torch.asin(sinp))
    yaw = torch.atan2(2.0 * (w * z + x * y), w * w + x * x - y * y - z * z)
    two_pi = 2 * torch.pi
    return roll % two_pi, pitch % two_pi, yaw % two_pi


class Go1WrestlingWrapper(EmptyWrapper):
    """go1_wrestling_wrapper.py:9-89: obs = (pos, rpy) self / other with agent 1 mirrored in y (12); reward on agent 0 only:
    success when agent 1 is flipped (|pitch| > 0.9 pi or |roll| >= 0.4 pi), punishment when agent 0 is.  Reward is [N, A, 1]."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(12,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.reward_buffer = {"punishment": 0, "success reward": 0, "step count": 0}

    def _init_extras(self, obs):
        agent_ids = self.env_agent_indices.reshape(-1)
        self.target_pos = self.base_init_state[agent_ids][:, 2].reshape(self.env.num_envs * self.env.num_agents, -1) - 0.5

    def _obs(self, obs_buf):
        base_info = self._base_info(obs_buf)
        obs = torch.cat([base_info, torch.flip(base_info, [1])], dim=2)
        obs[:, 1, 1::3] = -obs[:, 1, 1::3]                    # y and pitch of both halves, as seen by the mirrored agent
        return obs

    def reset(self):
        obs_buf = self.env.reset()
        self._init_extras(obs_buf)
        if not hasattr(self, "_fused"):
            self._fused = self._fuse(5, [self.success_reward_scale, self.punishment_scale], {"success reward": 0, "punishment": 1, "step count": 8})
        return self._wobs if self._fused else self._obs(obs_buf)

    def step(self, action):
        action[:, 1, 1:] = -action[:, 1, 1:]                  # in place, as the reference does
        if getattr(self, "_fused", False):
            obs, reward, termination, info = self._fused_step(action)
            return obs, reward.reshape(self.env.num_envs, self.env.num_agents, 1), termination, info
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        N, A = self.env.num_envs, self.env.num_agents
        self._acc("step count", 1)
        reward = torch.zeros([N, A], device=self.env.device, dtype=torch.float)
        r, p, _ = get_euler_xyz(obs_buf.base_quat)
        r = torch.where(r > torch.pi, r - 2 * torch.pi, r).reshape(N, A)       # to (-pi, pi]
        p = torch.where(p > torch.pi, p - 2 * torch.pi, p).reshape(N, A)
        flipped = (p.abs() > torch.pi * 0.9) | (r.abs() >= torch.pi * 0.4)       # [N, A]
        if self.success_reward_scale != 0:
            success = flipped[:, 1] * float(self.success_reward_scale)
            reward[:, 0] += success
            self._acc("success reward", torch.sum(success))
        if self.punishment_scale != 0:
            punishment = flipped[:, 0] * float(self.punishment_scale)
            reward[:, 0] -= punishment
            self._acc("punishment", torch.sum(punishment))
        return self._obs(obs_buf), reward.reshape(N, A, 1), termination, info


class Go1BridgeWrapper(EmptyWrapper):
    """go1_bridge_wrapper.py:8-80: obs = (pos, rpy) self / other, agent 1 sees x reflected about the midpoint of the two start
    positions and pitch negated (12); reward on agent 0 only: success when agent 1 is below z = 0.5 (fell off), punishment
    when agent 0 is, target reward once agent 0 is past agent 1's start x.  Reward is [N, A]."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(12,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.reward_buffer = {"target reward": 0, "success reward": 0, "punishment": 0, "step count": 0}

    def _init_extras(self, obs):
        base_pos = obs.base_pos.reshape(self.env.num_envs, self.env.num_agents, -1)
        self.target_pos = torch.flip(base_pos, [1]).clone()

    def _obs(self, obs_buf):
        base_info = self._base_info(obs_buf)
        obs = torch.cat([base_info, torch.flip(base_info, [1])], dim=2)
        span = (self.target_pos[:, 0, 0] + self.target_pos[:, 1, 0]).abs()
        obs[:, 1, 0] = span - obs[:, 1, 0]
        obs[:, 1, 4] = -obs[:, 1, 4]
        obs[:, 1, 6] = span - obs[:, 1, 6]
        obs[:, 1, 10] = -obs[:, 1, 10]
        return obs

    def reset(self):
        obs_buf = self.env.reset()
        self._init_extras(obs_buf)
        if not hasattr(self, "_fused"):
            self._fused = self._fuse(6, [self.success_reward_scale, self.punishment_scale, self.target_reward_scale],
                                     {"success reward": 0, "punishment": 1, "target reward": 2, "step count": 8})
        return self._wobs if self._fused else self._obs(obs_buf)

    def step(self, action):
        action[:, 1, 1:] = -action[:, 1, 1:]
        if getattr(self, "_fused", False):
            return self._fused_step(action)
        obs_buf, _, termination, info = self.env.step_from_wrapper(action)
        N, A = self.env.num_envs, self.env.num_agents
        self._acc("step count", 1)
        reward = torch.zeros([N, A], device=self.env.device, dtype=torch.float)
        base_pos = obs_buf.base_pos.reshape(N, A, -1)
        if self.success_reward_scale != 0:
            success = (base_pos[:, 1, 2] < 0.5) * float(self.success_reward_scale)
            reward[:, 0] += success
            self._acc("success reward", torch.sum(success))
        if self.punishment_scale != 0:
            punishment = (base_pos[:, 0, 2] < 0.5) * float(self.punishment_scale)
            reward[:, 0] -= punishment
            self._acc("punishment", torch.sum(punishment))
        if self.target_reward_scale != 0:
            target = (base_pos[:, 0, 0] > self.target_pos[:, 0, 0]) * float(self.target_reward_scale)
            reward[:, 0] += target
            self._acc("target reward", torch.sum(target))
        return self._obs(obs_buf), reward, termination, info


class Go1TugWrapper(EmptyWrapper):
    """go1_tug_wrapper.py:9-136 (tug of war over a sliding disc): obs = (pos, rpy) self, disc (pos, vel), distance to the disc,
    disc pos again (10), agent 1 mirrored in y; reward on agent 0 only: success ~ how far the disc is on the opponent's side
    (y < 0), punishment for the own side, +/- for closing in on the disc.  Reward is [N, A, 1].

    Kept from the reference: no clip before the [2, .5, .5] scale (the env clips afterwards, go1.py:38); for three steps after
    an env reset the disc is pinned back to zero through `env.gym.set_dof_state_tensor_indexed` (:61-69, with the mask taken
    after the decrement); the last obs column is the disc position of THIS step (`last_npc_pos` is refreshed before the obs is
    built, :113); "total reward" leaves the position punishment out (:116).  The reference's `reward[:, 0] += x[:, 0]` on a
    [N, A, 1] tensor only broadcasts for num_envs = 1 (its config's value); here it is the same arithmetic for any N."""

    def __init__(self, env):
        super().__init__(env)
        self.observation_space = spaces.Box(low=-float("inf"), high=float("inf"), shape=(10,), dtype=float)
        self.action_space = spaces.Box(low=-1, high=1, shape=(3,), dtype=float)
        self.action_scale = torch.tensor([[[2, 0.5, 0.5]]], device=self.env.device).repeat(self.num_envs, self.num_agents, 1)
        self.reward_buffer = {"success reward": 0, "pos reward": 0, "pos punishment": 0, "step count": 0, "npc pos": 0, "punishment": 0,
                              "total reward": 0, "pos": 0, "pos_y": 0, "opponet pos": 0, "opponet pos_y": 0}
        self.reset_dic = torch.zeros([self.env.num_envs], dtype=torch.float, device=self.env.device)

    def _init_extras(self, obs):
        self.last_dis = torch.clone(obs.base_pos.reshape([self.env.num_envs, self.env.num_agents, -1]))
        self.last_npc_pos = torch.clone(self.env.dof_state_npc[:, :, :1])
        self.env_step = torch.zeros(self.env.num_envs, dtype=torch.float, device=self.env.device)

    def _obs(self, obs_buf):
        N, A = self.env.num_envs, self.env.num_agents
        base_info = self._base_info(obs_buf)
        npc = self.env.dof_state_npc
        dis = base_info[:, :, :2].clone()
        dis[:, :, 0] -= 1.6
        dis[:, :, 1] -= npc[:, :, 0].repeat(1, A)
        dis = torch.norm(dis, p=2, dim=-1, keepdim=True)
        obs = torch.cat([base_info, npc.repeat(1, A, 1), dis, self.last_npc_pos.repeat(1, A, 1)], dim=2)
        obs[:, 1, 1] = -obs[:, 1, 1]
        obs[:, 1, 4] = -obs[:, 1, 4]
        obs[:, 1, 6] = -obs[:, 1, 6]
        obs[:, 1, -1] = -obs[:, 1, -1]
        return obs

    def reset(self):
        obs_buf = self.env.reset()
        self._init_extras(obs_buf)
        obs = self._obs(obs_buf)
        self.env_step[:] += 1
        return obs

    def step(self, action):
        action[:, 1, 1:] = -action[:, 1, 1:]
        if bool((self.reset_dic > 0).any()):
            self.reset_dic[self.reset_dic > 0] -= 1
            pin = self.reset_dic > 0
            self.env.dof_state_npc[pin, 0, 0] = 0.0
            self.env.dof_state_npc[pin, 0, 1] = 0.0
            npc_indices = self.npc_indices.reshape(-1)
            self.env.gym.set_dof_state_tensor_indexed(self.env.sim, self.env.all_dof_states, npc_indices, len(npc_indices))
        obs_buf, _, termination, info = self.env.step((action * self.action_scale).reshape(-1, self.action_space.shape[0]))
        self.reset_dic[self.env.reset_ids] = 3
        self._acc("step count", 1)
        N, A = self.env.num_envs, self.env.num_agents
        base_pos = obs_buf.base_pos.reshape([N, A, -1])
        npc_y = self.env.dof_state_npc[:, 0, 0]
        last_dis = self.last_dis[:, 0, :2].clone()
        last_dis[:, 0] -= 1.6
        last_dis[:, 1] -= npc_y
        last_dis = torch.norm(last_dis, p=2, dim=-1)
        dis = base_pos[:, 0, :2].clone()
        dis[:, 0] -= 1.6
        dis[:, 1] -= npc_y
        dis = torch.norm(dis, p=2, dim=-1)
        reward = torch.zeros([N, A, 1], device=self.env.device, dtype=torch.float)
        last_y = self.last_npc_pos[:, 0, 0]
        zero = torch.zeros_like(npc_y)
        success = torch.where(npc_y < 0, float(self.success_reward_scale) * -npc_y, zero)
        success = torch.where(last_y <= npc_y, success / 2, success)
        punishment = torch.where(npc_y > 0, float(self.punishment_reward_scale) * npc_y, zero)
        punishment = torch.where(last_y > npc_y, punishment / 2, punishment)
        pos_reward = torch.where(dis < last_dis, (last_dis - dis) * float(self.pos_reward_scale), zero)
        pos_punishment = torch.where(dis >= last_dis, torch.pow(2.0, dis) * float(self.pos_punishment_scale), zero)
        reward[:, 0, 0] += success - punishment + pos_reward - pos_punishment
        self._acc("success reward", torch.sum(success))
        self._acc("punishment", torch.sum(punishment))
        self._acc("pos reward", torch.sum(pos_reward))
        self._acc("pos punishment", torch.sum(pos_punishment))
        self.last_dis = torch.clone(base_pos)
        self.last_npc_pos = torch.clone(self.env.dof_state_npc[:, :, :1])
        self._acc("npc pos", torch.sum(npc_y))
        self._acc("total reward", torch.sum(pos_reward) + torch.sum(success) - torch.sum(punishment))
        self._acc("pos", torch.sum(base_pos[:, 0, 0]))
        self._acc("pos_y", torch.sum(base_pos[:, 0, 1]))
        self._acc("opponet pos", torch.sum(base_pos[:, 1, 0]))
        self._acc("opponet pos_y", torch.sum(base_pos[:, 1, 1]))
        obs = self._obs(obs_buf)
        self.env_step[:] += 1
        self.env_step[self.env.reset_ids] = 1
        return obs, reward, termination, info
