from mqe_b200.envs.configs import class_to_dict, merge_dict  # noqa: F401
from mqe_b200.envs.utils import get_args, make_env, set_seed  # noqa: F401
