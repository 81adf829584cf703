"""mqe VecEnv surface: `from mqe_b200.envs import make_mqe_env, ENV_DICT, custom_cfg` mirrors `mqe.envs.utils`."""
from .utils import ENV_DICT, custom_cfg, get_args, make_env, make_mqe_env, set_seed  # noqa: F401
