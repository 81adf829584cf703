"""Init-time model compiler: URDF -> flat articulation tables for the CUDA engine.

The reference hands `resources/robots/go1/urdf/go1.urdf` to Isaac Gym's asset
loader (`mqe/envs/base/legged_robot.py:763-790`, options at
`mqe/envs/go1/go1_config.py:55-84`: collapse_fixed_joints=True,
replace_cylinder_with_capsule=True, armature=0, density=0.001).  Isaac Gym is a
closed binary, so this module restates what that loader produces for the Go1:

* fixed joints are collapsed into their parent except the four feet, which carry
  `dont_collapse="true"` (go1.urdf:207,330,453,576) -> 17 rigid bodies, 12 DOF;
* DOF / body order is Isaac Gym's (FL, FR, RL, RR x hip, thigh, calf), which the
  frozen walk-these-ways policy assumes (SURVEY.md "bookkeeping notes");
* for the dynamics the (fixed) foot is merged into the calf -> 13 dynamic links.

Collision geometry is re-expressed as rounded primitives that the kernels test
analytically: sphere "probes" against the static world / boxes, and capsules for
dynamic-vs-dynamic pairs (see DESIGN.md "collision model").
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

LEG_ORDER = ("FL", "FR", "RL", "RR")          # Isaac Gym DOF order
JOINT_KINDS = ("hip", "thigh", "calf")
MAX_PROBES = 32
MAX_CAPS = 20

RESOURCE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources")


# ----------------------------------------------------------------------------- ctypes mirror of include/mqe_b200.h
class RobotModelC(ctypes.Structure):
    _fields_ = [
        ("base_inertial", ctypes.c_float * 10),
        ("leg_offsets", ctypes.c_float * (4 * 4 * 3)),
        ("leg_inertial", ctypes.c_float * (4 * 3 * 10)),
        ("q_lower", ctypes.c_float * 12),
        ("q_upper", ctypes.c_float * 12),
        ("qd_limit", ctypes.c_float * 12),
        ("tau_limit", ctypes.c_float * 12),
        ("q_default", ctypes.c_float * 12),
        ("n_probes", ctypes.c_int),
        ("n_caps", ctypes.c_int),
        ("probes", ctypes.c_float * (MAX_PROBES * 6)),
        ("caps", ctypes.c_float * (MAX_CAPS * 9)),
    ]


def _rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _vec(s, n=3):
    v = [float(x) for x in s.split()]
    assert len(v) == n
    return np.array(v)


@dataclass
class Inertial:
    """mass, COM and inertia tensor about the COM, both in the owning link frame."""
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))
    inertia: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))

    def moved(self, R, t):
        """Re-express in a parent frame: x_parent = R x + t."""
        return Inertial(self.mass, R @ self.com + t, R @ self.inertia @ R.T)

    def merged(self, other: "Inertial"):
        m = self.mass + other.mass
        if m <= 0.0:
            return Inertial()
        c = (self.mass * self.com + other.mass * other.com) / m

        def shift(b):
            d = b.com - c
            return b.inertia + b.mass * (d @ d * np.eye(3) - np.outer(d, d))
        return Inertial(m, c, shift(self) + shift(other))

    def flat10(self):
        I = self.inertia
        return [self.mass, *self.com, I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]]


def _read_inertial(link_el):
    el = link_el.find("inertial")
    if el is None:
        return Inertial()
    org = el.find("origin")
    xyz = _vec(org.get("xyz", "0 0 0")) if org is not None else np.zeros(3)
    rpy = _vec(org.get("rpy", "0 0 0")) if org is not None else np.zeros(3)
    m = float(el.find("mass").get("value"))
    i = el.find("inertia")
    g = lambda k: float(i.get(k, "0"))
    I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
    R = _rpy_to_mat(rpy)
    return Inertial(m, xyz, R @ I @ R.T)


def _read_collisions(link_el):
    out = []
    for c in link_el.findall("collision"):
        org = c.find("origin")
        xyz = _vec(org.get("xyz", "0 0 0")) if org is not None else np.zeros(3)
        rpy = _vec(org.get("rpy", "0 0 0")) if org is not None else np.zeros(3)
        g = c.find("geometry")
        if g.find("box") is not None:
            out.append(("box", xyz, _rpy_to_mat(rpy), _vec(g.find("box").get("size"))))
        elif g.find("sphere") is not None:
            out.append(("sphere", xyz, _rpy_to_mat(rpy), float(g.find("sphere").get("radius"))))
        elif g.find("cylinder") is not None:
            cy = g.find("cylinder")
            out.append(("cylinder", xyz, _rpy_to_mat(rpy), (float(cy.get("radius")), float(cy.get("length")))))
    return out


@dataclass
class Go1Model:
    base: Inertial
    leg_offsets: np.ndarray            # [4 legs][hip_pos, thigh_off, calf_off, foot_off][3]
    leg_links: list                    # [4][3] Inertial (calf includes the fixed foot)
    q_lower: np.ndarray
    q_upper: np.ndarray
    qd_limit: np.ndarray
    q_default: np.ndarray
    tau_limit: np.ndarray
    probes: list                       # (link, body, x, y, z, r)
    caps: list                         # (link, body, p0xyz, p1xyz, r)
    body_names: list
    dof_names: list
    foot_mass: float

    # ------------------------------------------------------------------ derived
    @property
    def total_mass(self):
        return self.base.mass + sum(l.mass for leg in self.leg_links for l in leg)

    @property
    def num_bodies(self):
        return len(self.body_names)

    @property
    def feet_indices(self):
        return [i for i, n in enumerate(self.body_names) if "foot" in n]

    def to_c(self) -> RobotModelC:
        m = RobotModelC()
        m.base_inertial[:] = self.base.flat10()
        m.leg_offsets[:] = [float(x) for x in np.asarray(self.leg_offsets).reshape(-1)]
        flat = []
        for leg in self.leg_links:
            for l in leg:
                flat += l.flat10()
        m.leg_inertial[:] = flat
        m.q_lower[:] = list(self.q_lower)
        m.q_upper[:] = list(self.q_upper)
        m.qd_limit[:] = list(self.qd_limit)
        m.tau_limit[:] = list(self.tau_limit)
        m.q_default[:] = list(self.q_default)
        assert len(self.probes) <= MAX_PROBES and len(self.caps) <= MAX_CAPS
        m.n_probes = len(self.probes)
        m.n_caps = len(self.caps)
        pf = [0.0] * (MAX_PROBES * 6)
        for i, p in enumerate(self.probes):
            pf[i * 6:(i + 1) * 6] = [float(x) for x in p]
        m.probes[:] = pf
        cf = [0.0] * (MAX_CAPS * 9)
        for i, c in enumerate(self.caps):
            cf[i * 9:(i + 1) * 9] = [float(x) for x in c]
        m.caps[:] = cf
        return m

    def to_json(self):
        return {
            "base": self.base.flat10(),
            "leg_offsets": np.asarray(self.leg_offsets).tolist(),
            "leg_links": [[l.flat10() for l in leg] for leg in self.leg_links],
            "q_lower": list(map(float, self.q_lower)), "q_upper": list(map(float, self.q_upper)),
            "qd_limit": list(map(float, self.qd_limit)), "q_default": list(map(float, self.q_default)),
            "tau_limit": list(map(float, self.tau_limit)),
            "probes": [list(map(float, p)) for p in self.probes],
            "caps": [list(map(float, c)) for c in self.caps],
            "body_names": self.body_names, "dof_names": self.dof_names, "foot_mass": self.foot_mass,
        }

    @staticmethod
    def from_json(d):
        def inert(f):
            I = np.array([[f[4], f[5], f[6]], [f[5], f[7], f[8]], [f[6], f[8], f[9]]])
            return Inertial(f[0], np.array(f[1:4]), I)
        return Go1Model(
            base=inert(d["base"]), leg_offsets=np.array(d["leg_offsets"]),
            leg_links=[[inert(l) for l in leg] for leg in d["leg_links"]],
            q_lower=np.array(d["q_lower"]), q_upper=np.array(d["q_upper"]), qd_limit=np.array(d["qd_limit"]),
            q_default=np.array(d["q_default"]), tau_limit=np.array(d["tau_limit"]),
            probes=d["probes"], caps=d["caps"], body_names=d["body_names"], dof_names=d["dof_names"],
            foot_mass=d["foot_mass"])


def compile_go1_urdf(path, default_joint_angles, torque_limits=(20.0, 20.0, 25.0)) -> Go1Model:
    """Parse the Go1 URDF the way the reference's asset options ask Isaac Gym to.

    default_joint_angles: name -> rad (`go1_config.py:88-103`, assigned by joint *name*,
    `legged_robot.py:629-633`).  torque_limits: `go1_config.py:115` overriding the URDF
    effort through `legged_robot_field.py:309-319`.
    """
    root = ET.parse(path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = {}
    for j in root.findall("joint"):
        org = j.find("origin")
        joints[j.find("child").get("link")] = dict(
            name=j.get("name"), type=j.get("type"), parent=j.find("parent").get("link"),
            xyz=_vec(org.get("xyz", "0 0 0")) if org is not None else np.zeros(3),
            rpy=_vec(org.get("rpy", "0 0 0")) if org is not None else np.zeros(3),
            axis=_vec(j.find("axis").get("xyz")) if j.find("axis") is not None else None,
            limit=j.find("limit"), keep=j.get("dont_collapse", "false") == "true")

    def fixed_children(name):
        return [c for c, j in joints.items() if j["parent"] == name and j["type"] == "fixed" and not j["keep"]]

    def collapsed(name):
        """Inertial + colliders of `name` with all collapsible fixed descendants folded in (link frame)."""
        inert = _read_inertial(links[name])
        cols = list(_read_collisions(links[name]))
        for c in fixed_children(name):
            j = joints[c]
            R = _rpy_to_mat(j["rpy"])
            ci, cc = collapsed(c)
            inert = inert.merged(ci.moved(R, j["xyz"]))
            cols += [(k, R @ xyz + j["xyz"], R @ Rc, dim) for (k, xyz, Rc, dim) in cc]
        return inert, cols

    base_inert, base_cols = collapsed("base")
    body_names = ["base"]
    dof_names, q_lower, q_upper, qd_limit, q_default = [], [], [], [], []
    leg_offsets = np.zeros((4, 4, 3))
    leg_links, probes, caps = [], [], []

    # --- base colliders: trunk box -> rounded-box corner probes + one capsule; head box -> 4 probes + sphere
    RB = 0.02                                               # rounding radius of box corner probes
    for kind, xyz, Rc, dim in base_cols:
        assert kind == "box"
        h = np.asarray(dim) / 2.0
        if dim[0] > 0.2:                                     # trunk
            for sx in (1, -1):
                for sy in (1, -1):
                    for sz in (1, -1):
                        p = xyz + Rc @ (np.array([sx, sy, sz]) * (h - RB))
                        probes.append((0, 0, *p, RB))
            r = float(h[2])
            caps.append((0, 0, *(xyz + Rc @ np.array([h[0] - r, 0, 0])), *(xyz + Rc @ np.array([-(h[0] - r), 0, 0])), r))
        else:                                                # head / additional collision box
            for sy in (1, -1):
                for sz in (1, -1):
                    p = xyz + Rc @ (np.array([1, sy, sz]) * (h - RB))
                    probes.append((0, 0, *p, RB))
            r = float(h[1])
            c = xyz + Rc @ np.array([h[0] - r, 0, 0])
            caps.append((0, 0, *c, *c, r))

    foot_mass = 0.0
    for li, leg in enumerate(LEG_ORDER):
        chain = []
        for ki, kind in enumerate(JOINT_KINDS):
            lname = f"{leg}_{kind}"
            j = joints[lname]
            assert j["type"] == "revolute" and np.allclose(j["rpy"], 0)
            expect_axis = [1, 0, 0] if kind == "hip" else [0, 1, 0]
            assert np.allclose(j["axis"], expect_axis), (lname, j["axis"])
            leg_offsets[li, ki] = j["xyz"]
            dof_names.append(j["name"])
            q_lower.append(float(j["limit"].get("lower")))
            q_upper.append(float(j["limit"].get("upper")))
            qd_limit.append(float(j["limit"].get("velocity")))
            q_default.append(float(default_joint_angles[j["name"]]))
            inert, cols = collapsed(lname)
            body_names.append(lname)
            link_id = 1 + 3 * li + ki
            body_id = len(body_names) - 1
            for ckind, xyz, Rc, dim in cols:
                if ckind == "cylinder":                      # hip: capsule along the local z of the collider
                    r, length = dim
                    ax = Rc @ np.array([0, 0, 1.0])
                    p0, p1 = xyz + ax * length / 2, xyz - ax * length / 2
                    caps.append((link_id, body_id, *p0, *p1, r))
                    outer = p0 if abs(p0[1]) > abs(p1[1]) else p1
                    probes.append((link_id, body_id, *outer, r))
                elif ckind == "box":                         # thigh / calf: capsule along the long axis
                    h = np.asarray(dim) / 2.0
                    ax = Rc @ np.array([1.0, 0, 0])
                    r = float(max(h[1], h[2]))
                    p0, p1 = xyz + ax * (h[0] - r), xyz - ax * (h[0] - r)
                    caps.append((link_id, body_id, *p0, *p1, r))
                    if kind == "thigh":                      # knee end probes the world
                        low = p0 if p0[2] < p1[2] else p1
                        probes.append((link_id, body_id, *low, r))
            chain.append(inert)
        # foot: kept as its own rigid body for contact reporting, merged into the calf for dynamics
        fname = f"{leg}_foot"
        fj = joints[fname]
        leg_offsets[li, 3] = fj["xyz"]
        finert, fcols = collapsed(fname)
        foot_mass = finert.mass
        chain[2] = chain[2].merged(finert.moved(_rpy_to_mat(fj["rpy"]), fj["xyz"]))
        body_names.append(fname)
        for ckind, xyz, Rc, dim in fcols:
            assert ckind == "sphere"
            c = fj["xyz"] + xyz
            probes.append((1 + 3 * li + 2, len(body_names) - 1, *c, dim))
            caps.append((1 + 3 * li + 2, len(body_names) - 1, *c, *c, dim))
        leg_links.append(chain)

    # feet first so that capped contact lists always keep them
    probes.sort(key=lambda p: 0 if "foot" in body_names[int(p[1])] else 1)
    return Go1Model(base=base_inert, leg_offsets=leg_offsets, leg_links=leg_links,
                    q_lower=np.array(q_lower), q_upper=np.array(q_upper), qd_limit=np.array(qd_limit),
                    q_default=np.array(q_default), tau_limit=np.array(list(torque_limits) * 4),
                    probes=probes, caps=caps, body_names=body_names, dof_names=dof_names, foot_mass=foot_mass)


def load_go1_model(path=None) -> Go1Model:
    """Load the committed, pre-compiled Go1 tables (generated by tools/extract_assets.py)."""
    path = path or os.path.join(RESOURCE_DIR, "go1_model.json")
    with open(path) as f:
        return Go1Model.from_json(json.load(f))
