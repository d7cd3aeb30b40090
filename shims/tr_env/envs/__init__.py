from tensegrity_rl_b200.envs import tr_env  # noqa: F401  (same class name, kwargs and reset / step signatures)
