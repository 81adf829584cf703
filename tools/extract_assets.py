#!/usr/bin/env python
"""Generate the committed model/weight resources and golden fixtures from /root/reference.

Runs ONLY in the build container (the reference tree is not present on the GPU box).  Outputs:
  multiagent-quadruped-environment_b200/resources/go1_model.json    compiled Go1 tables (model.py)
  multiagent-quadruped-environment_b200/resources/walk_policy.npz   walk-these-ways + actuator-net weights
  tests/golden/mlp_kat.npz                                          TorchScript known-answer vectors
Terrain goldens are produced by tools/gen_terrain_golden.py.
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MQE_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

from mqe_b200.model import compile_go1_urdf  # noqa: E402

RES = os.path.join(REPO, "multiagent-quadruped-environment_b200", "resources")
GOLD = os.path.join(REPO, "tests", "golden")

# go1_config.py:88-103
DEFAULT_ANGLES = {
    "FL_hip_joint": 0.1, "RL_hip_joint": 0.1, "FR_hip_joint": -0.1, "RR_hip_joint": -0.1,
    "FL_thigh_joint": 0.8, "RL_thigh_joint": 1.0, "FR_thigh_joint": 0.8, "RR_thigh_joint": 1.0,
    "FL_calf_joint": -1.5, "RL_calf_joint": -1.5, "FR_calf_joint": -1.5, "RR_calf_joint": -1.5,
}


def main():
    import json
    os.makedirs(RES, exist_ok=True)
    os.makedirs(GOLD, exist_ok=True)
    model = compile_go1_urdf(os.path.join(REF, "resources/robots/go1/urdf/go1.urdf"), DEFAULT_ANGLES)
    with open(os.path.join(RES, "go1_model.json"), "w") as f:
        json.dump(model.to_json(), f, indent=1)
    print("bodies", model.num_bodies, "mass", model.total_mass, "probes", len(model.probes), "caps", len(model.caps))

    wdir = os.path.join(REF, "mqe/utils/locomotion_checkpoints/walk_these_ways")
    body = torch.jit.load(os.path.join(wdir, "body_latest.jit"), map_location="cpu")
    adapt = torch.jit.load(os.path.join(wdir, "adaptation_module_latest.jit"), map_location="cpu")
    act = torch.jit.load(os.path.join(REF, "resources/actuator_nets/unitree_go1.pt"), map_location="cpu")
    w = {}
    for prefix, mod in (("body", body), ("adapt", adapt), ("act", act)):
        for k, v in mod.state_dict().items():
            w[f"{prefix}.{k}"] = v.detach().numpy().astype(np.float32)
    np.savez(os.path.join(RES, "walk_policy.npz"), **w)

    # known-answer vectors straight from the reference's TorchScript modules
    torch.manual_seed(0)
    x = torch.randn(4, 2100)
    with torch.no_grad():
        lat = adapt(x)
        y = body(torch.cat((x, lat), dim=-1))
    torch.manual_seed(0)
    xa = torch.randn(5, 6)
    with torch.no_grad():
        ya = act(xa)
    g = torch.Generator().manual_seed(1234)
    xs = torch.rand(64, 2100, generator=g) * 2 - 1          # obs-history-like magnitudes
    xas = torch.randn(256, 6, generator=g) * torch.tensor([0.3, 0.3, 0.3, 5.0, 5.0, 5.0])
    with torch.no_grad():
        lats = adapt(xs)
        ys = body(torch.cat((xs, lats), dim=-1))
        yas = act(xas)
    np.savez(os.path.join(GOLD, "mlp_kat.npz"), x=x.numpy(), latent=lat.numpy(), action=y.numpy(),
             xa=xa.numpy(), torque=ya.numpy(), xs=xs.numpy(), latents=lats.numpy(), actions=ys.numpy(),
             xas=xas.numpy(), torques=yas.numpy())
    print("KAT body[0,:4]", y[0, :4].tolist())
    print("KAT act", ya.flatten().tolist())


if __name__ == "__main__":
    main()
