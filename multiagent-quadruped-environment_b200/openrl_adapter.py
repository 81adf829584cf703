"""numpy <-> device adapter the OpenRL scripts put around the env (`openrl_ws/utils.py:31-155`), SURVEY 8(f).2.

Same class names and semantics (`make_env`, `mqe_openrl_wrapper`, `MATWrapper`, `SingleAgentWrapper`), without importing
`openrl` / `isaacgym`.  What differs from the reference, none of it in the returned values:

* **Host path.**  The reference does `torch.from_numpy(0.5 * actions).cuda()` and three `.cpu().numpy()` calls on pageable memory per
  step (`utils.py:55-61`).  For the tasks whose wrapper gather is fused into the step graph (sheep, seesaw, football-defender) `step()`
  goes through ONE C-ABI call, `mqe_sim_step_host_result`: the scaled actions are written into a page-locked buffer (H2D by DMA), and the
  observation, reward and done flags come back as ONE packed device->host copy into a page-locked buffer.  The arrays handed back are
  views into that buffer; two buffers alternate, so what step t returned stays intact until step t + 2 (the rollout loop of
  `openrl_ws/train.py` copies them into its own storage right away).  Other wrappers take the torch path with pinned staging.
* **Device-resident path** (`device_resident=True`, SURVEY 8(f).2 "keep obs/reward on device"): `step()` takes and returns CUDA tensors
  (DLPack-exportable views of the engine's double-buffered step result); nothing crosses PCIe.
* `batch_rewards` reads the device-side running sums once per logging interval instead of one `.cpu()` per reward term per step.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as E
from .envs import make_mqe_env
from .envs.gym_shim import Wrapper


def make_env(args, custom_cfg=None, single_agent=False, device_resident=False):
    """openrl_ws/utils.py:31-38"""
    env, env_cfg = make_mqe_env(args.task, args, custom_cfg=custom_cfg)
    if single_agent:
        env = SingleAgentWrapper(env)
    return mqe_openrl_wrapper(env, device_resident=device_resident), env_cfg


class mqe_openrl_wrapper(Wrapper):
    """openrl_ws/utils.py:40-90"""

    def __init__(self, env, device_resident=False):
        super().__init__(env)
        self.agent_num = self.env.num_agents
        self.parallel_env_num = self.env.num_envs
        self.action_space = self.env.action_space
        self.observation_space = self.env.observation_space
        self.device_resident = bool(device_resident)
        self._task = env.env if isinstance(env, SingleAgentWrapper) else env       # the task wrapper (mqe.envs.wrappers.*)
        self._host = None                                                           # pinned buffers of the fused host path
        self._h_act = None

    # -- fused host path --------------------------------------------------------------------------------------------
    def _host_ready(self):
        """Page-locked action / result buffers, allocated once the task wrapper has switched to the fused gather (first reset())."""
        if self._host is None and getattr(self._task, "_fused", False) and not self.device_resident:
            eng = self._task.env.engine
            L = eng.result_layout()
            n, a = self._task.env.num_envs, self._task.env._ctrl_agents
            bufs = {"act": eng.pin_host(np.zeros((n, a, 3), dtype=np.float32)),
                    "res": [eng.pin_host(np.zeros(int(L.total_bytes), dtype=np.uint8)) for _ in range(2)], "L": L, "k": 0}
            bufs["views"] = [E.Engine.split_result(r, L) for r in bufs["res"]]
            self._host = bufs
        return self._host is not None

    def reset(self, **kwargs):
        obs = self.env.reset()
        if self.device_resident or not torch.is_tensor(obs):
            return obs
        return obs.cpu().numpy()

    def step(self, actions, extra_data=None):
        if self.device_resident:
            return self._step_device(actions)
        if self._host_ready():
            return self._step_host(actions)
        dev = self.env.device
        a = np.ascontiguousarray(0.5 * np.asarray(actions, dtype=np.float32))
        if self._h_act is None or self._h_act.shape != a.shape:
            self._h_act = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
            self._h_out = {}
        self._h_act.copy_(torch.from_numpy(a))
        d_act = self._h_act.to(dev, non_blocking=True).clip(-1, 1)
        obs, reward, termination, info = self.env.step(d_act)
        if torch.is_tensor(obs):
            obs = self._to_host("obs", obs)
            rewards = self._to_host("rew", reward)[..., None]
        else:                                         # the shipped go1gate wrapper returns 0, 0 (go1_gate_wrapper.py:155)
            rewards = reward
        done = self._to_host("done", termination)
        torch.cuda.current_stream(dev).synchronize()
        dones = self._dones(done)
        return (obs.copy() if isinstance(obs, np.ndarray) else obs), (rewards.copy() if isinstance(rewards, np.ndarray) else rewards), dones, self._empty_infos(dones.shape[0])

    def _to_host(self, key, t):
        """device tensor -> numpy through a pinned staging tensor (asynchronous copy; the caller synchronises once)"""
        buf = self._h_out.get(key)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = self._h_out[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        buf.copy_(t, non_blocking=True)
        return buf.numpy()

    def _step_host(self, actions):
        h = self._host
        np.multiply(np.asarray(actions, dtype=np.float32).reshape(h["act"].shape), 0.5, out=h["act"])    # utils.py:55; the +-1 clip is the frame kernel's
        h["k"] ^= 1
        k = h["k"]
        self._task.step_host(h["act"], h["res"][k])
        obs, rew, done = h["views"][k]
        if isinstance(self.env, SingleAgentWrapper):                                 # utils.py:150-155
            n = self.env.num_envs
            obs, rew = obs.reshape(n, 1, -1), rew.reshape(n, 1)
            done = np.repeat(done, self._task.num_agents)
        rewards = rew[..., None]
        dones = self._dones(done)
        return obs, rewards, dones, self._empty_infos(dones.shape[0])

    def _empty_infos(self, n):
        """`infos = [{} ...]` of utils.py:63-65 without allocating num_envs dicts per step (0.2 ms for 4096 envs) and without looking at
        every one of them either (`any(infos)` alone is 50 us for 4096 envs): one list of distinct dicts is kept and handed out again; the
        dicts note on a shared flag when a caller writes into one, and only then are they cleared before the next hand-out."""
        infos = getattr(self, "_infos", None)
        if infos is None or len(infos) != n:
            self._infos_flag = [False]
            infos = self._infos = [_InfoDict(self._infos_flag) for _ in range(n)]
        elif self._infos_flag[0]:
            for d in infos:
                dict.clear(d)
            self._infos_flag[0] = False
        return infos

    def _dones(self, done):
        """`dones` [N, agent_num] (utils.py:61): the per-env flag repeated per agent into one of two alternating preallocated arrays
        (np.repeat allocates: 20-35 us per step at 4096-8192 envs); like the fused result views, an array stays intact across the NEXT step."""
        n = done.shape[0]
        bufs = getattr(self, "_done_bufs", None)
        if bufs is None or bufs[0].shape != (n, self.agent_num):
            bufs = self._done_bufs = [np.empty((n, self.agent_num), dtype=bool) for _ in range(2)]
            self._done_k = 0
        self._done_k ^= 1
        out = bufs[self._done_k]
        d = np.asarray(done, dtype=bool)
        for a in range(self.agent_num):                   # column stores: 6 us at 4096 envs (a broadcast copy takes 34, np.repeat 16)
            out[:, a] = d
        return out

    def _step_device(self, actions):
        """CUDA tensors in, CUDA tensors out (zero-copy views of the engine's step result; valid until the step after next)."""
        a = torch.as_tensor(actions, device=self.env.device, dtype=torch.float32) * 0.5
        obs, reward, termination, info = self.env.step(a.clip(-1, 1))
        rewards = reward.unsqueeze(-1) if torch.is_tensor(reward) else reward
        dones = termination.unsqueeze(-1).expand(-1, self.agent_num)
        return obs, rewards, dones, info

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


class _InfoDict(dict):
    """An `info` dict that raises a shared flag when somebody writes into it (see mqe_openrl_wrapper._empty_infos)."""
    __slots__ = ("_flag",)

    def __init__(self, flag):
        super().__init__()
        self._flag = flag

    def _touch(name):                                   # noqa: N805 -- class-body helper
        base = getattr(dict, name)

        def method(self, *a, **k):
            self._flag[0] = True
            return base(self, *a, **k)
        method.__name__ = name
        return method

    __setitem__ = _touch("__setitem__")
    __delitem__ = _touch("__delitem__")
    __ior__ = _touch("__ior__")
    update = _touch("update")
    setdefault = _touch("setdefault")
    pop = _touch("pop")
    popitem = _touch("popitem")
    clear = _touch("clear")
    del _touch


class MATWrapper(Wrapper):
    """openrl_ws/utils.py:92-129: for dict observation spaces with "policy" / "critic" entries the multi-agent transformer sees the
    "policy" part; everything else passes through."""

    def __init__(self, env):
        super().__init__(env)
        self._observation_space = None

    @staticmethod
    def _policy_part(space):
        sub = getattr(space, "spaces", None)
        return sub is not None and "critic" in sub.keys() and "policy" in sub.keys()

    @property
    def observation_space(self):
        space = self.env.observation_space if self._observation_space is None else self._observation_space
        return space["policy"] if self._policy_part(space) else space

    @observation_space.setter
    def observation_space(self, value):
        self._observation_space = value

    def observation(self, observation):
        space = self.env.observation_space if self._observation_space is None else self._observation_space
        return observation["policy"] if self._policy_part(space) else observation

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, actions, extra_data=None):
        return self.env.step(actions, extra_data)


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
