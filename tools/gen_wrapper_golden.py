#!/usr/bin/env python
"""Golden vectors for the task wrappers, produced by the REFERENCE's own wrapper code.

Runs in this container only (needs /root/reference).  `mqe/envs/wrappers/*.py` are plain PyTorch on top of a
`gym.Wrapper`; with `gym` stubbed, `torch.tensor(..., device="cuda")` redirected to the CPU and a fake env that feeds
recorded buffers, their `reset()` / `step()` run unmodified.  The fake env's inputs and the wrapper's outputs over a few
steps are stored in tests/golden/wrappers_<task>.npz; tests/test_wrappers_golden.py replays the same inputs through
mqe_b200.envs.wrappers on a fake env of the same shape.
"""
import copy
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def stub_modules():
    gym = types.ModuleType("gym")

    class Wrapper:
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

    class Box:
        def __init__(self, low, high, shape=None, dtype=float):
            self.shape = shape

    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    gym.Wrapper, gym.spaces = Wrapper, spaces
    sys.modules["gym"], sys.modules["gym.spaces"] = gym, spaces
    for name in ("mqe", "mqe.envs", "mqe.envs.wrappers", "isaacgym", "isaacgym.torch_utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["isaacgym.torch_utils"].__all__ = []

    def get_euler_xyz(q):      # restated from the public Isaac Gym Preview 4 torch_utils (xyzw, angles in [0, 2 pi)); SURVEY appendix B
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

    sys.modules["isaacgym.torch_utils"].get_euler_xyz = get_euler_xyz
    sys.modules["isaacgym"].gymtorch = types.SimpleNamespace(unwrap_tensor=lambda t: t)      # go1_tug_wrapper.py:1, :67-69


def load(name):
    path = os.path.join(REF, "mqe", "envs", "wrappers", name + ".py")
    spec = importlib.util.spec_from_file_location("mqe.envs.wrappers." + name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mqe.envs.wrappers." + name] = mod
    spec.loader.exec_module(mod)
    return mod


class Ns:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class FakeEnv:
    """Carries exactly the attributes the wrappers read."""

    def __init__(self, rec, cfg, num_agents, num_npcs):
        self.rec, self.t = rec, 0
        self.cfg = cfg
        self.num_envs = rec["base_pos"].shape[1] // num_agents
        self.num_agents, self.num_npcs = num_agents, num_npcs
        self.device = "cpu"
        self.env_origins = torch.as_tensor(rec["env_origins"])
        if num_npcs:
            self.npc_env_origins = self.env_origins.unsqueeze(1).repeat(1, num_npcs, 1)
        if "gate_pos" in rec:
            self.gate_pos = torch.as_tensor(rec["gate_pos"])
        self.env_agent_indices = torch.arange(self.num_envs * num_agents).view(self.num_envs, num_agents)
        self.base_init_state = torch.as_tensor(rec["base_init_state"])
        self.npc_indices = torch.arange(self.num_envs * max(num_npcs, 1), dtype=torch.int32)
        self.all_dof_states, self.sim = None, None
        self.gym = types.SimpleNamespace(set_dof_state_tensor_indexed=lambda *a: None)
        self._load(0)

    def _load(self, t):
        r = self.rec
        self.obs_buf = Ns(base_pos=torch.as_tensor(r["base_pos"][t]), base_rpy=torch.as_tensor(r["base_rpy"][t]),
                          lin_vel=torch.as_tensor(r["lin_vel"][t]), base_quat=torch.as_tensor(r["base_quat"][t]),
                          env_info={"gate_deviation": torch.as_tensor(r["gate_deviation"]).clone()})
        if self.num_npcs:
            self.root_states_npc = torch.as_tensor(r["root_states_npc"][t])
            self.dof_state_npc = torch.as_tensor(r["dof_state_npc"][t]).clone()
        self.collide_buf = torch.as_tensor(r["collide"][t])
        self.r_term_buff = torch.as_tensor(r["r_term"][t])
        self.p_term_buff = torch.as_tensor(r["p_term"][t])
        self.reset_buf = torch.as_tensor(r["reset"][t])
        self.reset_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        if self.num_npcs > 1 or "sheep_pos_avg" in r:
            self.sheep_pos_avg = torch.as_tensor(r["sheep_pos_avg"][t])
            self.sheep_pos_var = torch.as_tensor(r["sheep_pos_var"][t])

    def reset(self):
        self._load(0)
        return self.obs_buf

    def step(self, action):
        self.t += 1
        self._load(self.t)
        self.last_action = action
        return self.obs_buf, None, self.reset_buf, {}

    step_from_wrapper = step


def make_record(rng, N, A, P, T, with_gate=False):
    M = N * A
    rec = {
        "env_origins": rng.uniform(0, 20, size=(N, 3)).astype(np.float32),
        "base_pos": rng.uniform(-1, 9, size=(T, M, 3)).astype(np.float32),
        "base_rpy": rng.uniform(0, 6.28, size=(T, M, 3)).astype(np.float32),
        "lin_vel": rng.normal(size=(T, M, 3)).astype(np.float32),
        "gate_deviation": rng.uniform(-0.5, 0.5, size=(N, 2)).astype(np.float32),
        "collide": rng.random((T, N)) < 0.2, "r_term": rng.random((T, N)) < 0.1, "p_term": rng.random((T, N)) < 0.1,
        "reset": rng.random((T, N)) < 0.25,
        "actions": rng.uniform(-1.5, 1.5, size=(T, N, min(A, 2) if with_gate else A, 3)).astype(np.float32),
    }
    rec["base_pos"][:, :, 2] = rng.uniform(0.2, 1.5, size=(T, M))
    # drawn last so that the streams of the earlier goldens stay as they were: base orientation (some robots flipped) and
    # the per-agent initial root states the wrestling wrapper reads
    rng2 = np.random.default_rng(int(rng.integers(1 << 30)) if False else 12345 + N * 31 + A * 7 + P)
    rpy = rng2.uniform(-0.5, 0.5, size=(T, M, 3))
    flip = rng2.random((T, M)) < 0.3
    rpy[flip, 0] = rng2.uniform(1.3, 3.1, size=int(flip.sum())) * rng2.choice([-1.0, 1.0], size=int(flip.sum()))
    flip = rng2.random((T, M)) < 0.2
    rpy[flip, 1] = rng2.uniform(-1.5, 1.5, size=int(flip.sum()))
    cr, sr, cp, sp, cy, sy = (f(rpy[..., i] * 0.5) for i in range(3) for f in (np.cos, np.sin))
    rec["base_quat"] = np.stack([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp,
                                 cy * cr * cp + sy * sr * sp], axis=-1).astype(np.float32)
    rec["base_init_state"] = rng2.uniform(-1, 3, size=(M, 13)).astype(np.float32)
    rec["dof_state_npc"] = rng2.uniform(-1.5, 1.5, size=(T, N, 1, 2)).astype(np.float32)
    if A >= 2:                                   # some envs with the two agents within 0.5 m (agent-distance terms)
        bp = rec["base_pos"].reshape(T, N, A, 3)
        close = rng.random((T, N)) < 0.4
        bp[close, 1, :2] = bp[close, 0, :2] + rng.uniform(-0.3, 0.3, size=(int(close.sum()), 2)).astype(np.float32)
    if P:
        rs = rng.normal(size=(T, N * P, 13)).astype(np.float32)
        rs[:, :, :3] = rec["env_origins"].repeat(P, axis=0)[None] + rng.uniform(-1, 12, size=(T, N * P, 3))
        rec["root_states_npc"] = rs
        rec["sheep_pos_avg"] = rng.uniform(0, 10, size=(T, N, 2)).astype(np.float32)
        rec["sheep_pos_var"] = rng.uniform(0, 3, size=(T, N)).astype(np.float32)
    if with_gate:
        gp = rec["env_origins"].copy()
        gp[:, 0] += 11.0
        rec["gate_pos"] = gp
        far = rng.random((T, N)) < 0.3           # ball beyond the gate line in some envs (goal reward)
        rec["root_states_npc"][far, 0] = rec["env_origins"][None].repeat(T, 0)[far, 0] + gp[None].repeat(T, 0)[far, 0] + 1.0
    return rec


def run(task, wrapper_cls, cfg, A, P, seed, with_gate=False, T=6, N=5):
    rng = np.random.default_rng(seed)
    rec = make_record(rng, N, A, P, T + 1, with_gate)
    env = FakeEnv(rec, cfg, A, P)
    w = wrapper_cls(env)
    obs0 = w.reset()
    out = {"obs_reset": obs0.numpy() if torch.is_tensor(obs0) else np.zeros(0)}
    obs_l, rew_l = [], []
    for t in range(T):
        obs, rew, done, info = w.step(torch.as_tensor(rec["actions"][t]).clone())     # the rotation wrapper flips signs in place
        obs_l.append(obs.numpy().copy() if torch.is_tensor(obs) else np.zeros(0))
        rew_l.append(rew.numpy().copy() if torch.is_tensor(rew) else np.zeros(0))
        assert torch.equal(done, torch.as_tensor(rec["reset"][t + 1]))
    out["obs"], out["reward"] = np.stack(obs_l), np.stack(rew_l)
    out["reward_buffer_keys"] = np.array(sorted(w.reward_buffer.keys()))
    out["reward_buffer_vals"] = np.array([float(w.reward_buffer[k]) for k in sorted(w.reward_buffer.keys())], dtype=np.float64)
    out["last_scaled_action"] = env.last_action.numpy()
    np.savez_compressed(os.path.join(OUT, f"wrappers_{task}.npz"), **{"in_" + k: v for k, v in rec.items()}, **out)
    print(task, "obs", out["obs"].shape, "reward", out["reward"].shape, dict(zip(out["reward_buffer_keys"], np.round(out["reward_buffer_vals"], 3))))


def main():
    stub_modules()
    _orig_tensor = torch.tensor
    torch.tensor = lambda *a, **k: _orig_tensor(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})
    _orig_zeros = torch.zeros
    torch.zeros = lambda *a, **k: _orig_zeros(*a, **{kk: ("cpu" if kk == "device" else vv) for kk, vv in k.items()})
    load("empty_wrapper")
    sheep = load("go1_sheep_wrapper").Go1SheepWrapper
    seesaw = load("go1_seesaw_wrapper").Go1SeesawWrapper
    fb = load("go1_football_wrapper")
    sys.path.insert(0, os.path.dirname(OUT).rsplit("/tests", 1)[0])
    from mqe_b200.envs import configs as C
    run("go1sheep-hard", sheep, C.NineSheepCfg(), 2, 9, 1)
    run("go1sheep-easy", sheep, C.SingleSheepCfg(), 2, 1, 2)
    run("go1seesaw", seesaw, C.Go1SeesawCfg(), 2, 1, 3)
    run("go1football-defender", fb.Go1FootballDefenderWrapper, C.Go1FootballDefenderCfg(), 3, 1, 4, with_gate=True)
    run("go1pushbox", load("go1_pushbox_wrapper").Go1PushboxWrapper, C.Go1PushboxCfg(), 2, 1, 5)
    # the reference's rotation wrapper only broadcasts for num_envs <= 2 (go1_rotation_wrapper.py:77)
    run("go1revolvingdoor", load("go1_rotation_wrapper").Go1RotationWrapper, C.Go1RotationCfg(), 2, 1, 6, N=2, T=12)
    run("go1wrestling", load("go1_wrestling_wrapper").Go1WrestlingWrapper, C.Go1WrestlingCfg(), 2, 1, 7)
    run("go1bridge", load("go1_bridge_wrapper").Go1BridgeWrapper, C.Go1BridgeCfg(), 2, 1, 8)
    # the reference's tug wrapper only broadcasts for num_envs = 1 (go1_tug_wrapper.py:88: [N, 1] += [N])
    run("go1tug", load("go1_tug_wrapper").Go1TugWrapper, C.Go1TugCfg(), 2, 1, 9, N=1, T=16)


if __name__ == "__main__":
    main()
