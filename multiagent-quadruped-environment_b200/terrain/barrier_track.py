"""BarrierTrack terrain: numpy restatement + the static collision world the kernels use.

Restates `mqe/utils/terrain/barrier_track.py` (reference) for the block types the four BASELINE tasks
use (init / gate / plane / wall), with the SAME numpy RNG call order so that a given `np.random.seed`
reproduces `gate_deviation`, the heightfield and all origins bit for bit (checked against fixtures
generated from the reference itself: tools/gen_terrain_golden.py -> tests/golden/terrain_*.npz).

What differs is what happens to the heightfield afterwards.  The reference converts every track to a
triangle mesh and hands it to PhysX (`barrier_track.py:483-497`) plus one 2 cm ground slab
(`:628-632`).  Here the heightfield becomes a 2-D signed-distance field of the wall footprint
(`wall_sdf`), sampled bilinearly by the contact kernels: walls only take the
values {0, wall_height} and `slope_treshold` is inactive, so floor slab +
extruded footprint describes the same solid.
"""
from __future__ import annotations

import numpy as np

try:  # torch is only needed for the `env_info` tensors the wrappers read
    import torch
except Exception:  # pragma: no cover
    torch = None

GROUND_SLAB_TOP = 0.02      # barrier_track.py:628-632: box [sx, sy, 0.02] centred at size/2 -> top face z=0.02


def _iceil(x):
    return np.ceil(x).astype(int)


class BarrierTrack:
    """Same constructor and public attributes as the reference class (barrier_track.py:55-71)."""

    DEFAULTS = dict(                                   # barrier_track.py:13-53
        options=["gate", "init", "wall", "plane"], track_width=1.6, track_length=None,
        wall_thickness=0.04, wall_height=0.5,
        wall=dict(block_length=3.0), plane=dict(block_length=3.0),
        init=dict(block_length=1.2, room_size=(0.8, 0.8), border_with=0.05, offset=(0, 0)),
        gate=dict(block_length=1.2, width=1.0, depth=1.0, offset=(0, 0)),
        add_perlin_noise=False, border_perlin_noise=False, border_height=0.0, virtual_terrain=False,
        check_skill_combinations=False, engaging_next_threshold=0.0, curriculum_perlin=True,
        no_perlin_threshold=0.02,
    )

    def __init__(self, cfg, num_envs: int, num_agents: int = 1):
        assert cfg.mesh_type == "trimesh", f"BarrierTrack needs mesh_type trimesh, got {cfg.mesh_type}"
        kwargs = getattr(cfg, "BarrierTrack_kwargs", None)
        assert kwargs is not None, "Must provide BarrierTrack_kwargs in cfg.terrain"
        self.cfg, self.num_envs, self.num_agents = cfg, num_envs, num_agents
        self.track_kwargs = dict(self.DEFAULTS)
        self.track_kwargs.update(kwargs)
        if self.track_kwargs["add_perlin_noise"]:
            raise NotImplementedError("perlin noise is off for every task on the hot path (SURVEY.md #13)")
        self.env_origins = np.zeros((cfg.num_rows, cfg.num_cols, 3), dtype=np.float32)
        self.agent_origins = np.zeros((cfg.num_rows, cfg.num_cols, num_agents, 3), dtype=np.float32)
        self.env_info = None
        self.built = False

    # ------------------------------------------------------------------ geometry of one track (barrier_track.py:90-122)
    def _layout(self):
        hs = self.cfg.horizontal_scale
        kw = self.track_kwargs
        width_px = _iceil(kw["track_width"] / hs)
        total = 0.0
        self.env_block_lengths, self.track_block_resolutions = [], []
        for name in kw["options"]:
            total += kw[name]["block_length"]
            self.env_block_lengths.append(kw[name]["block_length"])
            self.track_block_resolutions.append((_iceil(kw[name]["block_length"] / hs), width_px))
        kw["track_length"] = total
        self.track_resolution = (_iceil(total / hs), width_px)
        self.n_blocks_per_track = len(kw["options"])
        self.env_length, self.env_width = total, kw["track_width"]

    def _wall_height_px(self):
        h = self.track_kwargs["wall_height"]
        if isinstance(h, (tuple, list)):
            h = np.random.uniform(*h)
        return h

    # Each painter returns (heights[px], noise mask, info dict, height offset, reset positions or None);
    # statement order inside follows the reference because later statements overwrite earlier ones.
    def _block_wall(self, thick, res):                 # barrier_track.py:157-179
        hf = np.zeros(res, dtype=np.float32)
        hf[:, :] = self._wall_height_px() / self.cfg.vertical_scale
        return hf, np.zeros(res, dtype=np.float32), {}, 0, None

    def _block_plane(self, thick, res):                # barrier_track.py:181-206
        hf, mask = np.zeros(res, dtype=np.float32), np.zeros(res, dtype=np.float32)
        top = self._wall_height_px() / self.cfg.vertical_scale
        t = _iceil(thick / self.cfg.horizontal_scale)
        hf[:, :t] = top
        hf[:, -t:] = top
        mask[:, t:res[1] - t] = 1.0
        return hf, mask, {}, 0, None

    def _block_init(self, thick, res):                 # barrier_track.py:208-262
        hs, kw, A = self.cfg.horizontal_scale, self.track_kwargs["init"], self.num_agents
        hf, mask = np.zeros(res, dtype=np.float32), np.zeros(res, dtype=np.float32)
        spawn = np.zeros((A, 3), dtype=np.float32)
        top = self._wall_height_px() / self.cfg.vertical_scale
        off = (int(kw["offset"][0] / hs), int(kw["offset"][1] / hs))
        room = (int(kw["room_size"][0] / hs), int(kw["room_size"][1] / hs))
        gap = _iceil(kw["border_width"] / hs)
        t = _iceil(thick / hs)
        span_y = room[1] * A + gap * (A - 1)
        x0 = _iceil((res[0] - room[0]) / 2) + off[0]
        y0 = _iceil((res[1] - span_y) / 2) + off[1]
        hf[:x0 + room[0], :] = top
        hf[:, :t] = top
        hf[:, -t:] = top
        mask[x0 + room[0]:, t:res[1] - t] = 1.0
        for i in range(A):
            ya, yb = y0 + i * (room[1] + gap), y0 + (i + 1) * room[1] + i * gap
            hf[x0:x0 + room[0], ya:yb] = 0.0
            mask[x0:x0 + room[0], ya:yb] = 1.0
            spawn[i, 0] = x0 + int(room[0] / 2)
            spawn[i, 1] = ya + int(room[1] / 2)
        hf[:, :t] = top
        hf[:, -t:] = top
        hf[:t, :] = top
        return hf, mask, {}, 0, spawn

    def _block_gate(self, thick, res):                 # barrier_track.py:311-362
        hs, kw = self.cfg.horizontal_scale, self.track_kwargs["gate"]
        hf, mask = np.zeros(res, dtype=np.float32), np.ones(res, dtype=np.float32)
        depth = np.random.uniform(*kw["depth"]) if isinstance(kw["depth"], (tuple, list)) else kw["depth"]
        top = self._wall_height_px() / self.cfg.vertical_scale
        off = np.asarray((_iceil(kw["offset"][0] / hs), _iceil(kw["offset"][1] / hs)))
        rnd = kw.get("random", None)
        if rnd is None:
            raise KeyError("random")                   # the reference indexes ["gate"]["random"] unconditionally
        jitter = np.asarray((rnd[0] / hs, rnd[1] / hs))
        jitter = _iceil(jitter * (np.random.random(2) - 0.5) * 2)
        width = np.random.uniform(*kw["width"]) if isinstance(kw["width"], (tuple, list)) else kw["width"]
        dpx, wpx = int(depth / hs), int(width / hs)
        t = _iceil(thick / hs)
        g = np.asarray([_iceil((res[0] - dpx) / 2), _iceil((res[1] - wpx) / 2)]) + off + jitter
        hf[g[0]:g[0] + dpx, :] = top
        hf[:, :t] = top
        hf[:, -t:] = top
        mask[g[0]:g[0] + dpx, :] = 0.0
        mask[:, :t] = 0.0
        mask[:, -t:] = 0.0
        hf[g[0]:g[0] + dpx, g[1]:g[1] + wpx] = 0.0
        mask[g[0]:g[0] + dpx, g[1]:g[1] + wpx] = 1.0
        info = {"gate_deviation": np.asarray(off + jitter, dtype=np.float32) * np.float32(hs)}
        return hf, mask, info, 0, None

    # ------------------------------------------------------------------ one track (barrier_track.py:412-499)
    def _build_track(self, origin_px, row, col):
        kw = self.track_kwargs
        thick = np.random.uniform(*kw["wall_thickness"]) if isinstance(kw["wall_thickness"], (tuple, list)) else kw["wall_thickness"]
        cursor = origin_px.copy()
        spawn, info = None, {}
        for idx, name in enumerate(kw["options"]):
            res = self.track_block_resolutions[idx]
            hf, mask, binfo, dz, reset = getattr(self, "_block_" + name)(thick, res)
            sl = (slice(cursor[0], cursor[0] + res[0]), slice(cursor[1], cursor[1] + res[1]))
            self.heightfield_raw[sl] = hf + mask * self.heightfield_raw[sl] + cursor[2]
            cursor[0] += res[0]
            cursor[2] += dz
            if reset is not None:
                if spawn is not None:
                    raise RuntimeError("Multiple reset block in the same track. Should be only one.")
                spawn = reset
            info.update(binfo)
        self.track_width_map[row, col] = self.env_width - thick * 2
        return spawn, info

    # ------------------------------------------------------------------ whole map (barrier_track.py:365-393, 501-565)
    def build(self):
        cfg, hs, vs = self.cfg, self.cfg.horizontal_scale, self.cfg.vertical_scale
        self._layout()
        self.border = int(cfg.border_size / hs)
        self.tot_rows = int(cfg.num_rows * self.track_resolution[0]) + 2 * self.border
        self.tot_cols = int(cfg.num_cols * self.track_resolution[1]) + 2 * self.border
        self.heightfield_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.float32)
        self.heightsamples = self.heightfield_raw
        self.track_width_map = np.zeros((cfg.num_rows, cfg.num_cols), dtype=np.float32)
        self.track_origins_px = np.zeros((cfg.num_rows, cfg.num_cols, 3), dtype=int)
        info_maps = None
        for col in range(cfg.num_cols):
            z_px = 0
            for row in range(cfg.num_rows):
                self.track_origins_px[row, col] = [int(row * self.track_resolution[0]) + self.border,
                                                   int(col * self.track_resolution[1]) + self.border, z_px]
                spawn, info = self._build_track(self.track_origins_px[row, col], row, col)
                o = self.track_origins_px[row, col].reshape(1, 3).repeat(self.num_agents, 0)
                self.agent_origins[row, col, :, :2] = (o[:, :2] + spawn[:, :2]) * hs
                self.agent_origins[row, col, :, 2] = (o[:, 2] + spawn[:, 2]) * vs
                if info_maps is None:
                    info_maps = {k: np.tile(np.asarray(v, dtype=np.float32).reshape(1, 1, -1), (cfg.num_rows, cfg.num_cols, 1))
                                 for k, v in info.items()}
                else:
                    for k, v in info.items():
                        info_maps[k][row, col, :] = v
        for i in range(cfg.num_rows):
            for j in range(cfg.num_cols):
                self.env_origins[i, j, 0] = self.track_origins_px[i, j, 0] * hs
                self.env_origins[i, j, 1] = self.track_origins_px[i, j, 1] * hs
                self.env_origins[i, j, 2] = self.track_origins_px[i, j, 2] * vs
                self.env_origins[i, j, 1] += self.track_kwargs["track_width"] / 2
        self.env_info_np = info_maps or {}
        self.env_info = ({k: torch.from_numpy(v.copy()) for k, v in self.env_info_np.items()} if torch is not None else dict(self.env_info_np))
        self.built = True
        return self

    def add_terrain_to_sim(self, gym=None, sim=None, device="cpu"):
        """Reference entry point (barrier_track.py:501); there is no gym to add meshes to."""
        self.device = device
        self.build()
        if torch is not None:
            self.env_info = {k: v.to(device) for k, v in self.env_info.items()}
            self.env_origins_pyt = torch.from_numpy(self.env_origins).to(device)

    # ------------------------------------------------------------------ static collision world for the kernels
    def wall_top(self):
        h = self.track_kwargs["wall_height"]
        return float(h if not isinstance(h, (tuple, list)) else max(h))

    def wall_sdf(self):
        """Signed distance [m] from every heightfield pixel centre to the wall footprint boundary.

        Pixel (i, j) sits at world (i*hs, j*hs) (`barrier_track.py:494-496`: the track mesh is placed at
        origin_px*hs and spans vertex i*hs).  A vertex is 'wall' when its height is above the slab; the
        one-cell ramps of the un-corrected trimesh put the wall face half a cell outside the first high
        vertex, which is exactly the pixel-centre convention of a distance transform.  Positive outside.
        """
        from scipy import ndimage
        hs = self.cfg.horizontal_scale
        wall = (self.heightfield_raw * self.cfg.vertical_scale) > (GROUND_SLAB_TOP + 0.03)
        if not wall.any():
            return np.full(wall.shape, 1.0e3, dtype=np.float32)
        outside = ndimage.distance_transform_edt(~wall) - 0.5
        inside = ndimage.distance_transform_edt(wall) - 0.5
        return (np.where(wall, -inside, outside) * hs).astype(np.float32)
