"""Task-wrapper observation / reward maths against golden vectors produced by the REFERENCE's own wrapper code
(tools/gen_wrapper_golden.py ran mqe/envs/wrappers/*.py unmodified on a fake env).  CPU only: the wrappers are host
torch code above the engine; the same fake env feeds mqe_b200.envs.wrappers here."""
import os

import numpy as np
import pytest
import torch

from mqe_b200.envs import configs as C
from mqe_b200.envs import wrappers as W

CASES = {
    "go1sheep-hard": (W.Go1SheepWrapper, C.NineSheepCfg, 2, 9),
    "go1sheep-easy": (W.Go1SheepWrapper, C.SingleSheepCfg, 2, 1),
    "go1seesaw": (W.Go1SeesawWrapper, C.Go1SeesawCfg, 2, 1),
    "go1football-defender": (W.Go1FootballDefenderWrapper, C.Go1FootballDefenderCfg, 3, 1),
    "go1pushbox": (W.Go1PushboxWrapper, C.Go1PushboxCfg, 2, 1),
    "go1revolvingdoor": (W.Go1RotationWrapper, C.Go1RotationCfg, 2, 1),
    "go1wrestling": (W.Go1WrestlingWrapper, C.Go1WrestlingCfg, 2, 1),
    "go1bridge": (W.Go1BridgeWrapper, C.Go1BridgeCfg, 2, 1),
    "go1tug": (W.Go1TugWrapper, C.Go1TugCfg, 2, 1),
}


class Ns:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class FakeEnv:
    def __init__(self, z, cfg, A, P):
        self.z, self.t, self.cfg = z, 0, cfg
        self.num_agents, self.num_npcs = A, P
        self.num_envs = z["in_base_pos"].shape[1] // A
        self.device = torch.device("cpu")
        self.env_origins = torch.as_tensor(z["in_env_origins"])
        self.npc_env_origins = self.env_origins.unsqueeze(1).repeat(1, max(P, 1), 1)
        if "in_gate_pos" in z.files:
            self.gate_pos = torch.as_tensor(z["in_gate_pos"])
        self.env_agent_indices = torch.arange(self.num_envs * A).view(self.num_envs, A)
        if "in_base_init_state" in z.files:
            self.base_init_state = torch.as_tensor(z["in_base_init_state"])
        self.npc_indices = torch.arange(self.num_envs * max(P, 1), dtype=torch.int32)
        self.all_dof_states, self.sim = None, None
        self.gym = Ns(set_dof_state_tensor_indexed=lambda *a: None)
        self._load(0)

    def _load(self, t):
        z = self.z
        self.obs_buf = Ns(base_pos=torch.as_tensor(z["in_base_pos"][t]), base_rpy=torch.as_tensor(z["in_base_rpy"][t]),
                          lin_vel=torch.as_tensor(z["in_lin_vel"][t]),
                          base_quat=torch.as_tensor(z["in_base_quat"][t]) if "in_base_quat" in z.files else None,
                          env_info={"gate_deviation": torch.as_tensor(z["in_gate_deviation"]).clone()})
        self.root_states_npc = torch.as_tensor(z["in_root_states_npc"][t])
        if "in_dof_state_npc" in z.files:
            self.dof_state_npc = torch.as_tensor(z["in_dof_state_npc"][t]).clone()
        self.collide_buf = torch.as_tensor(z["in_collide"][t])
        self.r_term_buff = torch.as_tensor(z["in_r_term"][t])
        self.p_term_buff = torch.as_tensor(z["in_p_term"][t])
        self.reset_buf = torch.as_tensor(z["in_reset"][t])
        self.reset_ids = self.reset_buf.nonzero(as_tuple=False).flatten()
        self.sheep_pos_avg = torch.as_tensor(z["in_sheep_pos_avg"][t])
        self.sheep_pos_var = torch.as_tensor(z["in_sheep_pos_var"][t])

    def reset(self):
        self._load(0)
        return self.obs_buf

    def step(self, action):
        """Go1.step(): the wrapper has scaled the action itself (go1_tug_wrapper.py:70); the env clips it afterwards."""
        self.last_action = action
        self.t += 1
        self._load(self.t)
        return self.obs_buf, None, self.reset_buf, {}

    def step_from_wrapper(self, action):
        """The engine applies clip(+-1) * [2, .5, .5] inside the frame kernel; record what it would receive."""
        self.last_action = torch.clip(action, -1, 1) * torch.tensor([2.0, 0.5, 0.5])
        self.t += 1
        self._load(self.t)
        return self.obs_buf, None, self.reset_buf, {}


@pytest.mark.parametrize("task", sorted(CASES))
def test_wrapper_matches_reference_code(task, golden_dir):
    cls, cfg_fn, A, P = CASES[task]
    z = np.load(os.path.join(golden_dir, f"wrappers_{task}.npz"))
    env = FakeEnv(z, cfg_fn(), A, P)
    w = cls(env)
    obs0 = w.reset()
    assert np.allclose(obs0.numpy(), z["obs_reset"], rtol=1e-6, atol=1e-6)
    T = z["obs"].shape[0]
    for t in range(T):
        obs, rew, done, info = w.step(torch.as_tensor(z["in_actions"][t]).clone())
        assert np.allclose(obs.numpy(), z["obs"][t], rtol=1e-6, atol=1e-6), (task, t)
        assert np.allclose(rew.numpy(), z["reward"][t], rtol=1e-5, atol=1e-5), (task, t, np.abs(rew.numpy() - z["reward"][t]).max())
        assert torch.equal(done, torch.as_tensor(z["in_reset"][t + 1]))
    # the action the env receives: clip then scale, reshaped to [N * A_ctrl, 3] by the reference (wrappers/*.py step())
    assert np.allclose(env.last_action.reshape(-1, 3).numpy(), z["last_scaled_action"], atol=1e-7)
    keys = sorted(w.reward_buffer.keys())
    assert keys == list(z["reward_buffer_keys"])
    vals = np.array([float(w.reward_buffer[k]) for k in keys])
    assert np.allclose(vals, z["reward_buffer_vals"], rtol=1e-4, atol=1e-3), dict(zip(keys, zip(vals, z["reward_buffer_vals"])))
