"""isaacgym.gymapi: the constants and attribute bags the mqe Python layer touches."""
from types import SimpleNamespace

SIM_PHYSX, SIM_FLEX = 0, 1
DOF_MODE_NONE, DOF_MODE_POS, DOF_MODE_VEL, DOF_MODE_EFFORT = 0, 1, 2, 3
UP_AXIS_Y, UP_AXIS_Z = 0, 1
DOMAIN_SIM, DOMAIN_ENV, DOMAIN_ACTOR = 0, 1, 2


class Vec3(SimpleNamespace):
    def __init__(self, x=0.0, y=0.0, z=0.0):
        super().__init__(x=x, y=y, z=z)


class Quat(SimpleNamespace):
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        super().__init__(x=x, y=y, z=z, w=w)


class SimParams(SimpleNamespace):
    """gymapi.SimParams as an attribute bag (helpers.py:143-166 fills dt, substeps, up_axis, gravity, physx.*)."""

    def __init__(self):
        super().__init__(dt=0.005, substeps=1, up_axis=UP_AXIS_Z, gravity=Vec3(0.0, 0.0, -9.81), use_gpu_pipeline=True,
                         physx=SimpleNamespace(use_gpu=True, num_threads=0, num_subscenes=0, solver_type=1,
                                               num_position_iterations=4, num_velocity_iterations=0, contact_offset=0.01,
                                               rest_offset=0.0, bounce_threshold_velocity=0.5, max_depenetration_velocity=1.0))


def acquire_gym():
    raise RuntimeError("isaacgym shim: there is no Gym object; mqe_b200.envs.Go1 talks to libmqe_b200.so (env.gym is a small facade)")
