"""ctypes binding of libmqe_b200.so (include/mqe_b200.h) -- the thin host side of the C ABI.

There is deliberately no CPU implementation behind this class: if the CUDA library is missing or no
sm_100 device is visible, construction raises.  PyTorch is used only to wrap the engine-owned device
buffers zero-copy (the role `gymtorch.wrap_tensor` plays in the reference, legged_robot.py:567-595).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .model import MAX_CAPS, MAX_PROBES, RobotModelC  # noqa: F401

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MQE_B200_LIB") or os.path.join(PKG_DIR, "csrc", "libmqe_b200.so")     # override: A/B runs of two builds on one box

ABI_VERSION = 7
LOC_OBS = 70
OBS_FLOATS = 71

(BUF_ROOT_STATES, BUF_DOF_STATES, BUF_CONTACT_FORCES, BUF_TORQUES, BUF_ACTIONS, BUF_LAST_ACTIONS, BUF_OBS,
 BUF_BASE_LIN_VEL, BUF_BASE_ANG_VEL, BUF_PROJ_GRAVITY, BUF_RESET, BUF_TIMEOUT, BUF_COLLIDE, BUF_ROLL_TERM,
 BUF_PITCH_TERM, BUF_ZLOW_TERM, BUF_ZHIGH_TERM, BUF_EPISODE_LENGTH, BUF_COMMANDS, BUF_LOC_OBS, BUF_LOC_ACTION,
 BUF_GAIT, BUF_HISTORY, BUF_SHEEP_STATS, BUF_STATS, BUF_CLOCK, BUF_WRAP_SUMS, BUF_WARP_TRACE,
 BUF_SUBSTEP_TORQUES, BUF_SUBSTEP_DOF_VEL, BUF_SUBSTEP_EXCEED, BUF_HISTORY_HI, BUF_HISTORY_LO, BUF_STEP_RESULT, BUF_COUNT) = range(35)
STAT_GATHER_TIMEOUT = 5
CONTROL_TYPES = {"C": 0, "control_net": 0, "P": 1, "T": 2, "V": 3}

OBS_SLICES = {                                   # include/mqe_b200.h MQE_OBS_*
    "base_pos": (0, 3), "base_quat": (3, 7), "dof_pos": (7, 19), "dof_vel": (19, 31), "lin_vel": (31, 34),
    "ang_vel": (34, 37), "last_action": (37, 49), "last_last_action": (49, 61), "projected_gravity": (61, 64),
    "clock_inputs": (64, 68), "base_rpy": (68, 71),
}

NPC_NONE, NPC_RIGID, NPC_SEESAW, NPC_BOX, NPC_PLATFORM = 0, 1, 2, 3, 4
WRAP_NONE, WRAP_SHEEP, WRAP_SEESAW, WRAP_FOOTBALL_DEFENDER, WRAP_PUSHBOX, WRAP_WRESTLING, WRAP_BRIDGE, WRAP_ROTATION = range(8)


class StepResultLayoutC(ctypes.Structure):
    """MqeStepResultLayout (include/mqe_b200.h): byte offsets of obs | reward | done inside the packed step result."""
    _fields_ = [("obs_off", ctypes.c_int64), ("obs_bytes", ctypes.c_int64), ("reward_off", ctypes.c_int64), ("reward_bytes", ctypes.c_int64),
                ("done_off", ctypes.c_int64), ("done_bytes", ctypes.c_int64), ("total_bytes", ctypes.c_int64),
                ("num_envs", ctypes.c_int32), ("Aw", ctypes.c_int32), ("D", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class WrapperDescC(ctypes.Structure):
    """MqeWrapperDesc (include/mqe_b200.h)"""
    _fields_ = [("kind", ctypes.c_int32), ("scale", ctypes.c_float * 8), ("h_gate", ctypes.POINTER(ctypes.c_float))]
NPC_PASSIVE, NPC_SHEEP = 0, 1
POLICY_FP32, POLICY_BF16X3, POLICY_BF16 = 0, 1, 2
POLICY_MODE_DEFAULT = int(os.environ.get("MQE_POLICY_MODE", POLICY_FP32))

_fp = ctypes.POINTER(ctypes.c_float)

WEIGHT_FIELDS = [                                 # (field, npz key, shape) -- order of MqeWeights
    ("adapt_w0", "adapt.0.weight", (256, 2100)), ("adapt_b0", "adapt.0.bias", (256,)),
    ("adapt_w1", "adapt.2.weight", (128, 256)), ("adapt_b1", "adapt.2.bias", (128,)),
    ("adapt_w2", "adapt.4.weight", (2, 128)), ("adapt_b2", "adapt.4.bias", (2,)),
    ("body_w0", "body.0.weight", (512, 2102)), ("body_b0", "body.0.bias", (512,)),
    ("body_w1", "body.2.weight", (256, 512)), ("body_b1", "body.2.bias", (256,)),
    ("body_w2", "body.4.weight", (128, 256)), ("body_b2", "body.4.bias", (128,)),
    ("body_w3", "body.6.weight", (12, 128)), ("body_b3", "body.6.bias", (12,)),
    ("act_w0", "act.0.weight", (32, 6)), ("act_b0", "act.0.bias", (32,)),
    ("act_w1", "act.2.weight", (32, 32)), ("act_b1", "act.2.bias", (32,)),
    ("act_w2", "act.4.weight", (1, 32)), ("act_b2", "act.4.bias", (1,)),
]


class WeightsC(ctypes.Structure):
    _fields_ = [(name, _fp) for name, _, _ in WEIGHT_FIELDS]


class SimDescC(ctypes.Structure):
    """Field-for-field mirror of MqeSimDesc."""
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("num_envs", ctypes.c_int32), ("num_agents", ctypes.c_int32), ("num_npcs", ctypes.c_int32),
        ("env_id_offset", ctypes.c_int32),
        ("npc_kind", ctypes.c_int32), ("npc_ctrl", ctypes.c_int32), ("npc_dofs", ctypes.c_int32),
        ("decimation", ctypes.c_int32), ("solver_iters", ctypes.c_int32), ("max_episode_length", ctypes.c_int32),
        ("term_mask", ctypes.c_int32), ("quat_alias", ctypes.c_int32), ("policy_mode", ctypes.c_int32),
        ("defender", ctypes.c_int32), ("command_vel", ctypes.c_int32),
        ("sim_dt", ctypes.c_float), ("gravity_z", ctypes.c_float),
        ("friction", ctypes.c_float), ("contact_offset", ctypes.c_float), ("max_depen_vel", ctypes.c_float),
        ("erp", ctypes.c_float), ("cfm", ctypes.c_float),
        ("floor_z", ctypes.c_float), ("wall_top_z", ctypes.c_float), ("limit_margin", ctypes.c_float),
        ("term_roll", ctypes.c_float), ("term_pitch", ctypes.c_float), ("term_zlow", ctypes.c_float),
        ("term_zhigh", ctypes.c_float),
        ("act_scale", ctypes.c_float * 3), ("cmd_scale", ctypes.c_float * 3),
        ("action_scale", ctypes.c_float), ("hip_scale", ctypes.c_float), ("clip_actions", ctypes.c_float),
        ("loc_obs_default", ctypes.c_float * LOC_OBS),
        ("dof_ratio_lo", ctypes.c_float), ("dof_ratio_hi", ctypes.c_float),
        ("base_vel_lo", ctypes.c_float), ("base_vel_hi", ctypes.c_float),
        ("has_base_pos_range", ctypes.c_int32), ("has_npc_pos_range", ctypes.c_int32),
        ("has_npc_rpy_range", ctypes.c_int32), ("max_pair_contacts", ctypes.c_int32),
        ("base_pos_x", ctypes.c_float * 2), ("base_pos_y", ctypes.c_float * 2),
        ("npc_pos_x", ctypes.c_float * 2), ("npc_pos_y", ctypes.c_float * 2),
        ("npc_rpy_r", ctypes.c_float * 2), ("npc_rpy_p", ctypes.c_float * 2), ("npc_rpy_y", ctypes.c_float * 2),
        ("npc_mass", ctypes.c_float), ("npc_inertia", ctypes.c_float), ("npc_radius", ctypes.c_float),
        ("npc_halflen", ctypes.c_float),
        ("sheep_scale", ctypes.c_float), ("sheep_randomness", ctypes.c_float),
        ("gate_x", ctypes.c_float), ("max_push_vel_xy", ctypes.c_float),
        ("npc_geom", ctypes.c_float * 16),
        ("seed", ctypes.c_uint64),
        ("sdf_nx", ctypes.c_int32), ("sdf_ny", ctypes.c_int32), ("sdf_cell", ctypes.c_float), ("push_interval", ctypes.c_int32),
        ("control_type", ctypes.c_int32), ("stiffness", ctypes.c_float), ("damping", ctypes.c_float),
        ("lag_enabled", ctypes.c_int32), ("lag_timesteps", ctypes.c_int32), ("soft_dof_pos_limit", ctypes.c_float),
        ("h_sdf", _fp), ("h_env_origins", _fp), ("h_agent_origins", _fp), ("h_base_init_state", _fp),
        ("h_npc_init_state", _fp), ("h_npc_dof_default", _fp), ("h_base_added_mass", _fp), ("h_env_friction", _fp),
        ("h_base_com_shift", _fp), ("h_motor_strength", _fp),
        ("model", RobotModelC),
        ("weights", WeightsC),
    ]


def as_fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def load_weights(path=None):
    """walk-these-ways + actuator-net weights (resources/walk_policy.npz, from tools/extract_assets.py)."""
    path = path or os.path.join(PKG_DIR, "resources", "walk_policy.npz")
    z = np.load(path)
    arrays = {}
    wc = WeightsC()
    for field, key, shape in WEIGHT_FIELDS:
        a = np.ascontiguousarray(z[key], dtype=np.float32)
        assert a.shape == shape, (key, a.shape, shape)
        arrays[field] = a
        setattr(wc, field, as_fp(a))
    return wc, arrays


class EngineError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """dlopen libmqe_b200.so and declare prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise EngineError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.mqe_last_error.restype = ctypes.c_char_p
    lib.mqe_abi_version.restype = i32
    lib.mqe_device_count.restype = i32
    lib.mqe_sim_create.argtypes = [ctypes.POINTER(SimDescC), i32, vp, ctypes.POINTER(vp)]
    lib.mqe_sim_destroy.argtypes = [vp]
    lib.mqe_sim_set_stream.argtypes = [vp, vp]
    lib.mqe_sim_set_action_scale.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.mqe_sim_get_buffer.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_int32)]
    lib.mqe_sim_reset.argtypes = [vp]
    lib.mqe_sim_step.argtypes = [vp, vp]
    lib.mqe_sim_step_host.argtypes = [vp, vp, vp, vp]
    lib.mqe_sim_step_joint.argtypes = [vp, vp]
    lib.mqe_sim_step_result_layout.argtypes = [vp, ctypes.POINTER(StepResultLayoutC)]
    lib.mqe_sim_step_host_result.argtypes = [vp, vp, vp]
    lib.mqe_sim_result_parity.argtypes = [vp]
    lib.mqe_sim_stage_timing.argtypes = [vp, i32]
    lib.mqe_sim_stage_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.mqe_sim_gather_init.argtypes = [vp, i32, i32, vp]
    lib.mqe_sim_gather_connect.argtypes = [vp, vp]
    lib.mqe_sim_gather_parity.argtypes = [vp]
    lib.mqe_sim_gather_view.argtypes = [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(StepResultLayoutC)]
    lib.mqe_sim_set_wrapper.argtypes = [vp, vp]
    lib.mqe_sim_wrapper_reset.argtypes = [vp]
    lib.mqe_sim_pin_host.argtypes = [vp, vp, ctypes.c_size_t]
    lib.mqe_sim_unpin_host.argtypes = [vp, vp]
    lib.mqe_sim_policy.argtypes = [vp, vp]
    lib.mqe_sim_substeps.argtypes = [vp, i32]
    lib.mqe_sim_post_physics.argtypes = [vp]
    lib.mqe_sim_set_root_indexed.argtypes = [vp, vp, vp, i32]
    lib.mqe_sim_set_dof_indexed.argtypes = [vp, vp, vp, i32]
    lib.mqe_policy_forward.argtypes = [vp, vp, i32, vp, vp]
    lib.mqe_actuator_forward.argtypes = [vp, vp, i32, vp]
    lib.mqe_sim_history_head.argtypes = [vp]
    lib.mqe_sim_history_head.restype = i32
    lib.mqe_sim_synchronize.argtypes = [vp]
    lib.mqe_sim_launch_count.argtypes = [vp]
    lib.mqe_sim_launch_count.restype = i64
    if path == LIB_PATH:
        _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "mqe_last_error", "mqe_abi_version", "mqe_device_count", "mqe_sim_create", "mqe_sim_destroy", "mqe_sim_set_stream", "mqe_sim_set_action_scale",
    "mqe_sim_set_wrapper", "mqe_sim_wrapper_reset", "mqe_sim_get_buffer", "mqe_sim_reset", "mqe_sim_step", "mqe_sim_step_host", "mqe_sim_pin_host", "mqe_sim_unpin_host", "mqe_sim_policy", "mqe_sim_substeps",
    "mqe_sim_post_physics", "mqe_sim_set_root_indexed", "mqe_sim_set_dof_indexed", "mqe_policy_forward",
    "mqe_actuator_forward", "mqe_sim_history_head", "mqe_sim_synchronize", "mqe_sim_launch_count",
    "mqe_sim_step_joint", "mqe_sim_step_result_layout", "mqe_sim_step_host_result", "mqe_sim_result_parity", "mqe_sim_gather_init", "mqe_sim_gather_connect", "mqe_sim_gather_view", "mqe_sim_gather_parity",
    "mqe_sim_stage_timing", "mqe_sim_stage_ms",
]


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch.as_tensor can alias an engine-owned buffer."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}
        self._owner = owner


_TYPESTR = {("f", 4): "<f4", ("f", 8): "<f8", ("u", 1): "|u1", ("i", 8): "<i8", ("i", 4): "<i4", ("h", 2): "<i2"}
_BUF_KIND = {BUF_SUBSTEP_EXCEED: "u", BUF_HISTORY_HI: "h", BUF_HISTORY_LO: "h", BUF_STEP_RESULT: "u", BUF_RESET: "u", BUF_TIMEOUT: "u", BUF_COLLIDE: "u", BUF_ROLL_TERM: "u", BUF_PITCH_TERM: "u",
             BUF_ZLOW_TERM: "u", BUF_ZHIGH_TERM: "u", BUF_EPISODE_LENGTH: "i", BUF_STATS: "i", BUF_WARP_TRACE: "i"}


class Engine:
    """Owns one MqeSim handle (one process per GPU; `device` is the CUDA ordinal of this rank)."""

    def __init__(self, desc: SimDescC, device: int = 0, stream: int | None = None, keepalive=None):
        self.lib = load_library()
        if self.lib.mqe_abi_version() != ABI_VERSION:
            raise EngineError("libmqe_b200.so ABI version mismatch; rebuild")
        self._keep = keepalive
        self.desc = desc
        self.device = device
        h = ctypes.c_void_p()
        self._check(self.lib.mqe_sim_create(ctypes.byref(desc), device, ctypes.c_void_p(stream or 0), ctypes.byref(h)))
        self.h = h
        self._views = {}

    def _check(self, rc):
        if rc != 0:
            raise EngineError(f"mqe error {rc}: {self.lib.mqe_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.mqe_sim_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- buffers -------------------------------------------------------------------------------------
    def buffer_info(self, which):
        ptr = ctypes.c_void_p()
        shape = (ctypes.c_int64 * 4)()
        es = ctypes.c_int32()
        self._check(self.lib.mqe_sim_get_buffer(self.h, which, ctypes.byref(ptr), shape, ctypes.byref(es)))
        dims = [int(s) for s in shape if s > 0]
        return ptr.value, dims, es.value

    def tensor(self, which):
        """Zero-copy torch view of an engine buffer (gymtorch.wrap_tensor equivalent)."""
        import torch
        if which not in self._views:
            ptr, dims, es = self.buffer_info(which)
            kind = _BUF_KIND.get(which, "f")
            arr = _DevArray(ptr, dims, _TYPESTR[(kind, es)], self)
            self._views[which] = torch.as_tensor(arr, device=f"cuda:{self.device}")
        return self._views[which]

    # -- stepping ------------------------------------------------------------------------------------
    def reset(self):
        self._check(self.lib.mqe_sim_reset(self.h))

    def step(self, actions_ptr: int):
        self._check(self.lib.mqe_sim_step(self.h, ctypes.c_void_p(actions_ptr)))

    def step_joint(self, joint_actions_ptr: int):
        """control_type 'P' / 'V' / 'T': [N][12A] joint actions (go1.py:43-45)."""
        self._check(self.lib.mqe_sim_step_joint(self.h, ctypes.c_void_p(joint_actions_ptr)))

    # -- packed step result (what the learner reads: wrapper obs | reward | done) ---------------------
    def result_layout(self) -> StepResultLayoutC:
        L = StepResultLayoutC()
        self._check(self.lib.mqe_sim_step_result_layout(self.h, ctypes.byref(L)))
        return L

    def result_parity(self) -> int:
        """Half of BUF_STEP_RESULT ([2, total_bytes]) holding the latest step's result."""
        return int(self.lib.mqe_sim_result_parity(self.h))

    def result_views(self):
        """[(obs, reward, done) torch views of half 0, ... of half 1] -- zero-copy; the engine alternates halves every step."""
        buf, L = self.tensor(BUF_STEP_RESULT), self.result_layout()
        return [self.split_result(buf[h], L) for h in (0, 1)]

    @staticmethod
    def split_result(buf, L):
        """Views (obs [N, Aw, D] f32, reward [N, Aw] f32, done [N] bool) into a packed result held in a numpy u8 array or a torch u8 tensor."""
        N, Aw, D = L.num_envs, L.Aw, L.D
        if isinstance(buf, np.ndarray):
            obs = buf[L.obs_off:L.obs_off + L.obs_bytes].view(np.float32).reshape(N, Aw, D)
            rew = buf[L.reward_off:L.reward_off + L.reward_bytes].view(np.float32).reshape(N, Aw)
            done = buf[L.done_off:L.done_off + L.done_bytes].view(np.bool_)
        else:
            import torch
            obs = buf[L.obs_off:L.obs_off + L.obs_bytes].view(torch.float32).view(N, Aw, D)
            rew = buf[L.reward_off:L.reward_off + L.reward_bytes].view(torch.float32).view(N, Aw)
            done = buf[L.done_off:L.done_off + L.done_bytes].view(torch.bool)
        return obs, rew, done

    def step_host_result(self, h_actions: np.ndarray, h_result: np.ndarray):
        """H2D of the actions, one step, ONE D2H of the packed result (mqe_sim_step_host_result); blocks until it has landed."""
        self._check(self.lib.mqe_sim_step_host_result(self.h, h_actions.ctypes.data_as(ctypes.c_void_p), h_result.ctypes.data_as(ctypes.c_void_p)))

    # -- peer exchange of the step result over NVLink (mqe_sim_gather_*) ------------------------------
    def gather_init(self, rank: int, world: int) -> bytes:
        h = (ctypes.c_ubyte * 64)()
        self._check(self.lib.mqe_sim_gather_init(self.h, rank, world, ctypes.cast(h, ctypes.c_void_p)))
        return bytes(h)

    def gather_connect(self, handles):
        blob = b"".join(handles)
        self._check(self.lib.mqe_sim_gather_connect(self.h, ctypes.cast(ctypes.c_char_p(blob), ctypes.c_void_p)))

    def gather_parity(self) -> int:
        return int(self.lib.mqe_sim_gather_parity(self.h))

    def gather_view(self, parity: int):
        """(torch u8 view of the global result half `parity`, its layout)."""
        import torch
        ptr = ctypes.c_void_p()
        L = StepResultLayoutC()
        self._check(self.lib.mqe_sim_gather_view(self.h, parity, ctypes.byref(ptr), ctypes.byref(L)))
        arr = _DevArray(ptr.value, [int(L.total_bytes)], "|u1", self)
        return torch.as_tensor(arr, device=f"cuda:{self.device}"), L

    def set_wrapper(self, kind: int, scales, gate: np.ndarray | None = None):
        """Fuse a task wrapper's obs / reward gather into the step (mqe_sim_set_wrapper)."""
        d = WrapperDescC()
        d.kind = int(kind)
        for i, v in enumerate(scales):
            d.scale[i] = float(v)
        if gate is not None:
            gate = np.ascontiguousarray(gate, dtype=np.float32)
            d.h_gate = gate.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        self._check(self.lib.mqe_sim_set_wrapper(self.h, ctypes.byref(d)))
        for b in (BUF_STEP_RESULT, BUF_WRAP_SUMS):
            self._views.pop(b, None)

    def wrapper_reset(self):
        self._check(self.lib.mqe_sim_wrapper_reset(self.h))

    def pin_host(self, arr: np.ndarray):
        """Page-lock a caller-owned numpy buffer so step_host DMAs straight from / into it; keep `arr` alive until close()."""
        assert arr.flags["C_CONTIGUOUS"]
        self._check(self.lib.mqe_sim_pin_host(self.h, arr.ctypes.data_as(ctypes.c_void_p), arr.nbytes))
        self._pinned = getattr(self, "_pinned", []) + [arr]
        return arr

    def step_host(self, h_actions: np.ndarray, h_obs: np.ndarray | None, h_reset: np.ndarray | None):
        self._check(self.lib.mqe_sim_step_host(
            self.h, h_actions.ctypes.data_as(ctypes.c_void_p),
            h_obs.ctypes.data_as(ctypes.c_void_p) if h_obs is not None else None,
            h_reset.ctypes.data_as(ctypes.c_void_p) if h_reset is not None else None))

    def policy(self, actions_ptr: int):
        self._check(self.lib.mqe_sim_policy(self.h, ctypes.c_void_p(actions_ptr)))

    def substeps(self, count: int):
        self._check(self.lib.mqe_sim_substeps(self.h, count))

    def post_physics(self):
        self._check(self.lib.mqe_sim_post_physics(self.h))

    def synchronize(self):
        self._check(self.lib.mqe_sim_synchronize(self.h))

    def launch_count(self) -> int:
        return int(self.lib.mqe_sim_launch_count(self.h))

    def stage_timing(self, enable=True):
        """Diagnostics: event marks between the stages of step() (they become nodes of the step graph); see stage_ms()."""
        self._check(self.lib.mqe_sim_stage_timing(self.h, 1 if enable else 0))

    def stage_ms(self):
        """Device time [ms] of the last step's stages: policy | physics (+ fused bookkeeping) | bookkeeping / gather / exchange launches | background join."""
        out = (ctypes.c_float * 4)()
        self._check(self.lib.mqe_sim_stage_ms(self.h, out))
        return {"policy": out[0], "physics": out[1], "bookkeeping": out[2], "background_join": out[3]}

    def set_action_scale(self, scale):
        arr = (ctypes.c_float * 3)(*[float(x) for x in scale])
        self._check(self.lib.mqe_sim_set_action_scale(self.h, arr))

    def history_head(self) -> int:
        return int(self.lib.mqe_sim_history_head(self.h))

    def history(self):
        """history_locomotion_obs in the reference layout [M, 2100], oldest frame first (go1.py:102)."""
        import torch
        order = [(self.history_head() + 1 + b) % 30 for b in range(30)]
        if int(self.desc.policy_mode) == POLICY_FP32:
            ring = self.tensor(BUF_HISTORY)                                # [M, 30, 80]
        else:                                                              # tensor-core modes: the bf16 hi / lo planes are the ring
            M = int(self.desc.num_envs) * int(self.desc.num_agents)
            planes = []
            for which in (BUF_HISTORY_HI, BUF_HISTORY_LO):                 # [tiles, 30, 10, 128, 8] -> [tiles*128, 30, 80]
                t = self.tensor(which).view(-1, 30, 10, 128, 8).view(torch.bfloat16).float()
                planes.append(t.permute(0, 3, 1, 2, 4).reshape(-1, 30, 80)[:M])
            ring = planes[0] + planes[1]
        return ring[:, order, :LOC_OBS].reshape(ring.shape[0], -1)

    def policy_forward(self, history, want_latent=True):
        """Stand-alone network op: history [rows, 2100] (cuda fp32) -> (latent [rows, 2], action [rows, 12])."""
        import torch
        h = history.contiguous()
        rows = h.shape[0]
        lat = torch.empty((rows, 2), device=h.device, dtype=torch.float32)
        act = torch.empty((rows, 12), device=h.device, dtype=torch.float32)
        self._check(self.lib.mqe_policy_forward(self.h, ctypes.c_void_p(h.data_ptr()), rows,
                                                ctypes.c_void_p(lat.data_ptr()), ctypes.c_void_p(act.data_ptr())))
        return lat, act

    def actuator_forward(self, x):
        import torch
        x = x.contiguous()
        out = torch.empty((x.shape[0],), device=x.device, dtype=torch.float32)
        self._check(self.lib.mqe_actuator_forward(self.h, ctypes.c_void_p(x.data_ptr()), x.shape[0], ctypes.c_void_p(out.data_ptr())))
        return out

    def set_root_indexed(self, root_states, actor_ids):
        self._check(self.lib.mqe_sim_set_root_indexed(self.h, ctypes.c_void_p(root_states.data_ptr()),
                                                      ctypes.c_void_p(actor_ids.data_ptr()), int(actor_ids.numel())))

    def set_dof_indexed(self, dof_states, actor_ids):
        self._check(self.lib.mqe_sim_set_dof_indexed(self.h, ctypes.c_void_p(dof_states.data_ptr()),
                                                     ctypes.c_void_p(actor_ids.data_ptr()), int(actor_ids.numel())))

    def set_stream(self, stream: int):
        self._check(self.lib.mqe_sim_set_stream(self.h, ctypes.c_void_p(stream)))
