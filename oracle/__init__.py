"""CPU oracle binding (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  It loads oracle/_build/liboracle_{f32,f64}.so (built by oracle/Makefile from mqe_oracle.c) and
drives them with the same MqeSimDesc the product consumes.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build(force=False):
    libs = [os.path.join(BUILD, f"liboracle_{p}.so") for p in ("f32", "f64")]
    src = os.path.join(HERE, "mqe_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "mqe_b200.h")
    stale = force or any((not os.path.exists(l)) or os.path.getmtime(l) < max(os.path.getmtime(src), os.path.getmtime(hdr))
                         for l in libs)
    if stale:
        subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
    return libs


_libs = {}


def load(precision="f64"):
    if precision not in _libs:
        build()
        from mqe_b200 import engine as E
        lib = ctypes.CDLL(os.path.join(BUILD, f"liboracle_{precision}.so"))
        vp = ctypes.c_void_p
        lib.orc_create.argtypes = [ctypes.POINTER(E.SimDescC)]
        lib.orc_create.restype = vp
        lib.orc_destroy.argtypes = [vp]
        lib.orc_reset.argtypes = [vp]
        lib.orc_step.argtypes = [vp, vp]
        lib.orc_step_joint.argtypes = [vp, vp]
        lib.orc_policy.argtypes = [vp, vp]
        lib.orc_substeps.argtypes = [vp, ctypes.c_int]
        lib.orc_post_physics.argtypes = [vp]
        lib.orc_get.argtypes = [vp, ctypes.c_int, vp]
        lib.orc_get.restype = ctypes.c_int64
        lib.orc_set.argtypes = [vp, ctypes.c_int, vp]
        lib.orc_set.restype = ctypes.c_int64
        lib.orc_real_size.restype = ctypes.c_int
        lib.orc_set_threads.argtypes = [ctypes.c_int]
        lib.orc_set_reset_state.argtypes = [vp, ctypes.c_int]
        lib.orc_torques.argtypes = [vp]
        lib.orc_observe.argtypes = [vp]
        lib.orc_policy_forward.argtypes = [ctypes.POINTER(E.WeightsC), vp, ctypes.c_int, vp, vp]
        lib.orc_actuator_forward.argtypes = [ctypes.POINTER(E.WeightsC), vp, ctypes.c_int, vp]
        lib.orc_robot_dynamics.argtypes = [vp, ctypes.c_double if precision == "f64" else ctypes.c_float, vp, vp, vp, vp, vp, vp, vp]
        lib.orc_robot_fk.argtypes = [vp, vp, vp, vp, vp]
        _libs[precision] = lib
    return _libs[precision]


_INT_BUFS = {}


class Oracle:
    """One oracle simulation built from a mqe_b200.scene.Scene."""

    def __init__(self, scene, precision="f64"):
        from mqe_b200 import engine as E
        self.E = E
        self.lib = load(precision)
        self.scene = scene
        self.h = self.lib.orc_create(ctypes.byref(scene.desc))
        self.N, self.A, self.P = scene.num_envs, scene.num_agents, scene.num_npcs

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.lib.orc_reset(self.h)

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        self.lib.orc_step(self.h, a.ctypes.data_as(ctypes.c_void_p))

    def step_joint(self, joint_actions):
        a = np.ascontiguousarray(joint_actions, dtype=np.float32)
        self.lib.orc_step_joint(self.h, a.ctypes.data_as(ctypes.c_void_p))

    def policy(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        self.lib.orc_policy(self.h, a.ctypes.data_as(ctypes.c_void_p))

    def substeps(self, n):
        self.lib.orc_substeps(self.h, n)

    def post_physics(self):
        self.lib.orc_post_physics(self.h)

    # test hooks (golden replays)
    def torques(self):
        self.lib.orc_torques(self.h)

    def observe(self):
        self.lib.orc_observe(self.h)

    def set_reset_state(self, on):
        self.lib.orc_set_reset_state(self.h, int(bool(on)))

    def get(self, which):
        E = self.E
        n = self.lib.orc_get(self.h, which, None)
        if n < 0:
            raise KeyError(which)
        if which in (E.BUF_RESET, E.BUF_TIMEOUT, E.BUF_COLLIDE, E.BUF_ROLL_TERM, E.BUF_PITCH_TERM, E.BUF_ZLOW_TERM, E.BUF_ZHIGH_TERM, E.BUF_SUBSTEP_EXCEED):
            out = np.zeros(n, dtype=np.uint8)
        elif which == E.BUF_EPISODE_LENGTH:
            out = np.zeros(n, dtype=np.int64)
        elif which == E.BUF_STATS:
            out = np.zeros(n, dtype=np.int32)
        else:
            out = np.zeros(n, dtype=np.float32)
        self.lib.orc_get(self.h, which, out.ctypes.data_as(ctypes.c_void_p))
        return out

    def set(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        rc = self.lib.orc_set(self.h, which, a.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise KeyError(which)

    # shaped views -------------------------------------------------------------------------------
    def root_states(self):
        return self.get(self.E.BUF_ROOT_STATES).reshape(self.N, self.A + self.P, 13)

    def dof_states(self):
        return self.get(self.E.BUF_DOF_STATES).reshape(self.N, -1, 2)

    def obs(self):
        return self.get(self.E.BUF_OBS).reshape(self.N * self.A, self.E.OBS_FLOATS)


def set_threads(n, precision="f32"):
    load(precision).orc_set_threads(int(n))


def policy_forward(weights_c, hist, precision="f64"):
    lib = load(precision)
    h = np.ascontiguousarray(hist, dtype=np.float32)
    rows = h.shape[0]
    lat = np.zeros((rows, 2), dtype=np.float32)
    act = np.zeros((rows, 12), dtype=np.float32)
    lib.orc_policy_forward(ctypes.byref(weights_c), h.ctypes.data_as(ctypes.c_void_p), rows,
                           lat.ctypes.data_as(ctypes.c_void_p), act.ctypes.data_as(ctypes.c_void_p))
    return lat, act


def actuator_forward(weights_c, x, precision="f64"):
    lib = load(precision)
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(x.shape[0], dtype=np.float32)
    lib.orc_actuator_forward(ctypes.byref(weights_c), x.ctypes.data_as(ctypes.c_void_p), x.shape[0],
                             out.ctypes.data_as(ctypes.c_void_p))
    return out


def robot_dynamics(model_c, quat, q, v, tau, gz=-9.81, precision="f64"):
    lib = load(precision)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    quat, q, v, tau = f(quat), f(q), f(v), f(tau)
    M, c, acc = np.zeros((18, 18)), np.zeros(18), np.zeros(18)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.orc_robot_dynamics(ctypes.byref(model_c), gz, p(quat), p(q), p(v), p(tau), p(M), p(c), p(acc))
    return M, c, acc


def robot_fk(model_c, quat, q, precision="f64"):
    lib = load(precision)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    quat, q = f(quat), f(q)
    links, feet = np.zeros((13, 3)), np.zeros((4, 3))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.orc_robot_fk(ctypes.byref(model_c), p(quat), p(q), p(links), p(feet))
    return links, feet
