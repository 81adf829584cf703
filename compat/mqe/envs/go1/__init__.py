from mqe_b200.envs.go1 import Go1, Go1FootballDefender, Go1Object, Go1Sheep  # noqa: F401
