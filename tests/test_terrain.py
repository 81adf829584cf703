"""BarrierTrack restatement vs fixtures produced by the reference generator itself
(tools/gen_terrain_golden.py; reference mqe/utils/terrain/barrier_track.py)."""
import hashlib
import os

import numpy as np
import pytest

from mqe_b200.envs import configs as C
from mqe_b200.terrain.barrier_track import BarrierTrack

TASKS = {
    "go1gate": C.Go1GateCfg, "go1sheep-easy": C.SingleSheepCfg, "go1sheep-hard": C.NineSheepCfg,
    "go1seesaw": C.Go1SeesawCfg, "go1football-defender": C.Go1FootballDefenderCfg, "go1revolvingdoor": C.Go1RotationCfg, "go1wrestling": C.Go1WrestlingCfg, "go1tug": C.Go1TugCfg, "go1bridge": C.Go1BridgeCfg,
    "go1pushbox": C.Go1PushboxCfg, "go1football-2vs2": C.Go1Football2vs2Cfg,
}


@pytest.mark.parametrize("task", sorted(TASKS))
@pytest.mark.parametrize("seed", [0, 1])
def test_heightfield_and_origins_bit_exact(task, seed, golden_dir):
    z = np.load(os.path.join(golden_dir, f"terrain_{task}_seed{seed}.npz"))
    cfg = TASKS[task]()
    np.random.seed(seed)
    bt = BarrierTrack(cfg.terrain, 4, cfg.env.num_agents)
    bt.add_terrain_to_sim(None, None, "cpu")
    hf = np.ascontiguousarray(bt.heightfield_raw)
    assert hf.dtype == np.float32 and tuple(z["hf_shape"]) == hf.shape
    assert hashlib.sha256(hf.tobytes()).digest() == z["hf_sha256"].tobytes()
    assert np.array_equal(z["hf_coarse"], hf[::4, ::4].astype(np.float16))
    assert np.array_equal(z["env_origins"], bt.env_origins)
    assert np.array_equal(z["agent_origins"], bt.agent_origins)
    assert np.array_equal(z["track_origins_px"], bt.track_origins_px)
    assert np.array_equal(z["track_width_map"], bt.track_width_map)
    for k in z.files:
        if k.startswith("info_"):
            assert np.array_equal(z[k], bt.env_info_np[k[5:]]), k


def test_survey_known_answers():
    """seed 0 go1gate gate_deviation / origins."""
    cfg = C.Go1GateCfg()
    np.random.seed(0)
    bt = BarrierTrack(cfg.terrain, 4, 2)
    bt.build()
    assert np.allclose(bt.env_info_np["gate_deviation"][0, 0], [0.05, 0.225])
    assert np.allclose(bt.env_origins[0, 0], [1.0, 2.5, 0.0])
    assert np.allclose(bt.agent_origins[0, 0], [[2.0, 1.75, 0.0], [2.0, 3.25, 0.0]])


def test_wall_sdf_sign_and_metric():
    cfg = C.Go1GateCfg()
    np.random.seed(0)
    bt = BarrierTrack(cfg.terrain, 4, 2).build()
    sdf = bt.wall_sdf()
    wall = bt.heightfield_raw > 0
    assert sdf.shape == wall.shape and sdf.dtype == np.float32
    assert (sdf[wall] < 0).all() and (sdf[~wall] > 0).all()
    # spawn points sit in the middle of 1.0 x 1.5 m rooms: the nearest wall is the room's back wall, 0.5 m away
    hs = cfg.terrain.horizontal_scale
    for a in range(2):
        i, j = np.round(bt.agent_origins[0, 0, a, :2] / hs).astype(int)
        assert abs(sdf[i, j] - 0.5) < 2 * hs
    # 1-Lipschitz in grid units
    assert np.abs(np.diff(sdf, axis=0)).max() <= hs * 1.0001 and np.abs(np.diff(sdf, axis=1)).max() <= hs * 1.0001
