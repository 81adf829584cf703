"""compat/: the reference's module paths (`mqe.*`, `isaacgym.*`) resolve to mqe_b200 so caller scripts import unchanged
(openrl_ws/utils.py:5-28, test.py:6-10 of the reference)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]))

IMPORTS = """
import isaacgym
from isaacgym import gymutil, gymapi, gymtorch
from isaacgym.gymutil import parse_device_str
from isaacgym.torch_utils import get_euler_xyz, quat_rotate_inverse, quat_apply, quat_from_euler_xyz, torch_rand_float, to_torch, get_axis_params
from mqe.utils import get_args
from mqe.envs.utils import make_mqe_env, custom_cfg, ENV_DICT
from mqe.envs.go1.go1_config import Go1Cfg
import torch
assert sorted(ENV_DICT) == sorted(["go1plane", "go1gate", "go1sheep-easy", "go1sheep-hard", "go1football-defender", "go1football-1vs1",
    "go1football-2vs2", "go1seesaw", "go1pushbox", "go1tug", "go1wrestling", "go1revolvingdoor", "go1bridge"])       # mqe/envs/utils.py:38-109
assert parse_device_str("cuda:3") == ("cuda", 3) and gymapi.SIM_PHYSX == 0
q = quat_from_euler_xyz(torch.tensor([0.3]), torch.tensor([-0.2]), torch.tensor([1.0]))
r, p, y = get_euler_xyz(q)
assert abs(float(r) - 0.3) < 1e-5 and abs(float(p) - (6.283185307 - 0.2)) < 1e-5 and abs(float(y) - 1.0) < 1e-5
v = torch.tensor([[0.1, -0.4, 0.9]])
assert torch.allclose(quat_rotate_inverse(q, quat_apply(q, v)), v, atol=1e-6)
assert get_axis_params(-1.0, 2) == [0.0, 0.0, -1.0]
assert Go1Cfg().env.num_agents >= 1
print("ok")
"""

LOOP = """
import sys, torch
sys.argv = ["test.py", "--num_envs", "8", "--headless"]
import isaacgym
from mqe.utils import get_args
from mqe.envs.utils import make_mqe_env, custom_cfg
args = get_args()
env, env_cfg = make_mqe_env("go1sheep-easy", args, custom_cfg(args))           # the loop of the reference's test.py:57-70
obs = env.reset()
for i in range(5):
    obs, _, done, _ = env.step(torch.tensor([[[1, 0, 0], [1, 0, 0]]], dtype=torch.float32).repeat(env.num_envs, 1, 1).cuda())
assert obs.shape == (8, 2, 18) and torch.isfinite(obs).all()
print("ok")
"""


def test_reference_module_paths_import():
    out = subprocess.run([sys.executable, "-c", IMPORTS], env=ENV, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


@pytest.mark.gpu
def test_reference_smoke_loop_runs_unchanged():
    out = subprocess.run([sys.executable, "-c", LOOP], env=ENV, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
