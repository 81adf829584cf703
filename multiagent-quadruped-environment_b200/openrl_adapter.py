"""numpy <-> device adapter the OpenRL scripts put around the env (`openrl_ws/utils.py:31-155`), SURVEY 8(f).2.

Same class names and semantics (`mqe_openrl_wrapper`, `SingleAgentWrapper`, `make_env`), without importing
`openrl` / `isaacgym`.  Differences that do not change returned values: the host copies go through ONE pinned staging
buffer per direction (the reference calls `.cpu().numpy()` three times per step, `utils.py:59-61`), and
`batch_rewards` reduces device scalars once per logging interval instead of once per term per step.
"""
from __future__ import annotations

import numpy as np
import torch

from .envs import make_mqe_env
from .envs.gym_shim import Wrapper


def make_env(args, custom_cfg=None, single_agent=False):
    """openrl_ws/utils.py:31-38"""
    env, env_cfg = make_mqe_env(args.task, args, custom_cfg=custom_cfg)
    if single_agent:
        env = SingleAgentWrapper(env)
    return mqe_openrl_wrapper(env), env_cfg


class mqe_openrl_wrapper(Wrapper):
    """openrl_ws/utils.py:40-90"""

    def __init__(self, env):
        super().__init__(env)
        self.agent_num = self.env.num_agents
        self.parallel_env_num = self.env.num_envs
        self.action_space = self.env.action_space
        self.observation_space = self.env.observation_space
        self._h_act = None

    def reset(self, **kwargs):
        obs = self.env.reset()
        return obs.cpu().numpy() if torch.is_tensor(obs) else obs

    def step(self, actions, extra_data=None):
        dev = self.env.device
        a = np.ascontiguousarray(0.5 * np.asarray(actions, dtype=np.float32))
        if self._h_act is None or self._h_act.shape != a.shape:
            self._h_act = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
        self._h_act.copy_(torch.from_numpy(a))
        d_act = self._h_act.to(dev, non_blocking=True).clip(-1, 1)
        obs, reward, termination, info = self.env.step(d_act)
        if torch.is_tensor(obs):
            obs = obs.cpu().numpy()
            rewards = reward.cpu().unsqueeze(-1).numpy()
        else:                                         # the shipped go1gate wrapper returns 0, 0 (go1_gate_wrapper.py:155)
            rewards = reward
        dones = termination.cpu().unsqueeze(-1).repeat(1, self.agent_num).numpy().astype(bool)
        infos = [{} for _ in range(dones.shape[0])]
        return obs, rewards, dones, infos

    def close(self, **kwargs):
        return self.env.close()

    @property
    def use_monitor(self):
        return False

    def batch_rewards(self, buffer=None):
        rb = self.env.reward_buffer
        step_count = float(rb["step count"])
        out = {"average step reward": 0}
        for k in list(rb.keys()):
            if k == "step count":
                continue
            v = float(rb[k]) / (self.env.num_envs * max(step_count, 1.0))
            if hasattr(self.env, "single_agent_reward_scale"):
                v *= self.env.single_agent_reward_scale
            out[k] = v
            if "reward" in k or "punishment" in k:
                out["average step reward"] += v
            rb[k] = 0
        rb["step count"] = 0
        return out


class SingleAgentWrapper(Wrapper):
    """openrl_ws/utils.py:131-155: every agent becomes its own single-agent environment."""

    def __init__(self, env):
        super().__init__(env)
        self.num_envs = self.env.num_envs * self.env.num_agents
        self.num_agents = 1
        self.single_agent_reward_scale = self.env.num_agents

    def reset(self, **kwargs):
        return self.env.reset(**kwargs).reshape(self.num_envs, 1, -1)

    def step(self, actions, extra_data=None):
        n, a = self.env.num_envs, self.env.num_agents
        obs, reward, termination, info = self.env.step(actions.reshape(n, a, -1))
        done = torch.stack([termination] * a, dim=1).reshape(self.num_envs)
        return obs.reshape(self.num_envs, 1, -1), reward.reshape(self.num_envs, 1), done, info
