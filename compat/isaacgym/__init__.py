"""Names reference-side scripts import from the closed `isaacgym` package (compat/README.md).  No PhysX here: simulation
calls are served by libmqe_b200.so through mqe_b200."""
from . import gymapi, gymtorch, gymutil, torch_utils  # noqa: F401
