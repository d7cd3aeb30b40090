from tensegrity_rl_b200.envs import tensegrity_env  # noqa: F401
