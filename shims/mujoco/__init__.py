"""Opt-in stand-in for the `mujoco` module as far as the reference's run.py uses it (/root/reference/run.py:4,158).
Only `mj_contactForce` is needed there; the env classes never import this."""
from tensegrity_rl_b200.envs import mj_contactForce  # noqa: F401

__version__ = "2.3.7+tsg.shim"
