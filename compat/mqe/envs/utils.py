"""mqe/envs/utils.py:38-134"""
from mqe_b200.envs.utils import ENV_DICT, custom_cfg, make_mqe_env  # noqa: F401
