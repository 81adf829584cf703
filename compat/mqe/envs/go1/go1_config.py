"""mqe/envs/go1/go1_config.py:34-311 -- the base Go1 config (a factory here: configs are attribute trees, not nested classes)."""
from mqe_b200.envs.configs import go1_base as Go1Cfg  # noqa: F401
