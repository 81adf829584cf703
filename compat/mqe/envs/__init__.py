from mqe_b200.envs import configs, wrappers  # noqa: F401
