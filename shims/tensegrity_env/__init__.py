"""Same registration as the reference's tensegrity_env/tensegrity_env/__init__.py."""
try:
    from gym.envs.registration import register
except ImportError:
    from gymnasium.envs.registration import register

register(
    id="tensegrity_env-v0",
    entry_point="tensegrity_env.envs:tensegrity_env",
    max_episode_steps=5000,
)
