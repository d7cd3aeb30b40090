"""tensegrity_rl_b200 -- B200-native batched simulator for the 3-bar tensegrity envs of
drsteinkauz/tensegrity-RL (hot path only: MuJoCo-style step + tr_env / tensegrity_env semantics)."""
from .model import load_model, env_config  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not need CUDA or the built library
    if name in ("TensegrityVecEnv",):
        from .vec_env import TensegrityVecEnv
        return TensegrityVecEnv
    if name in ("tr_env", "tensegrity_env", "make", "mj_contactForce"):
        from . import envs
        return getattr(envs, name)
    if name == "SacActor":
        from .policy import SacActor
        return SacActor
    raise AttributeError(name)
