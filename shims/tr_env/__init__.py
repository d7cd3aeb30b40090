"""Same registration as the reference's tr_env/tr_env/__init__.py, with the env class served by tensegrity_rl_b200."""
try:
    from gym.envs.registration import register
except ImportError:  # gymnasium-only installs
    from gymnasium.envs.registration import register

register(
    id="tr_env-v0",
    entry_point="tr_env.envs:tr_env",
    max_episode_steps=5000,
)
