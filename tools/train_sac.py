"""SAC training on the device-resident simulator: the training half of the reference's run.py (`--train`,
run.py:23-98) for N envs at once.  Prints one JSON line with collection / update throughput.

  python tools/train_sac.py --envs 4096 --timesteps 2000000 --desired_action straight [--starting_point zip]
                            [--lr_SAC 3e-4] [--gradient_steps 64] [--save_dir models] [--xml flat]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensegrity_rl_b200 import TensegrityVecEnv
from tensegrity_rl_b200.sac import SACLearner

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--timesteps", type=int, default=500_000)
ap.add_argument("--xml", default="flat")
ap.add_argument("--env", default="tr_env")
ap.add_argument("--desired_action", default="straight")
ap.add_argument("--desired_direction", type=float, default=1)
ap.add_argument("--terminate_when_unhealthy", default="yes", choices=["yes", "no"])
ap.add_argument("--starting_point", default=None, help="SB3 zip to resume from (run.py --starting_point)")
ap.add_argument("--lr_SAC", type=float, default=3e-4)
ap.add_argument("--gradient_steps", type=int, default=64, help="gradient steps per vec-env step")
ap.add_argument("--batch_size", type=int, default=256)
ap.add_argument("--buffer_size", type=int, default=4_000_000)
ap.add_argument("--learning_starts", type=int, default=100_000)
ap.add_argument("--save_dir", default=None)
ap.add_argument("--no_graph", action="store_true")
a = ap.parse_args()

env = TensegrityVecEnv(a.envs, xml_file=a.xml, env=a.env, desired_action=a.desired_action,
                       desired_direction=a.desired_direction, terminate_when_unhealthy=a.terminate_when_unhealthy == "yes",
                       auto_reset=True, reset_pool="auto")
lo, hi = float(env.action_space.low[0]), float(env.action_space.high[0])
L = SACLearner(env.obs_dim, 6, action_low=lo, action_high=hi, device="cuda", learning_rate=a.lr_SAC, batch_size=a.batch_size,
               buffer_size=a.buffer_size, learning_starts=a.learning_starts, gradient_steps=a.gradient_steps,
               use_cuda_graph=not a.no_graph)
if a.starting_point:
    L.load_sb3_zip(a.starting_point)
log = []
def cb(learner, vec_steps):
    l = learner.last_losses.tolist()
    log.append({"timesteps": learner.num_timesteps, "critic_loss": l[0], "actor_loss": l[1], "ent_coef": l[3],
                "mean_reward": float(env.reward.mean())})
torch.cuda.synchronize(); t0 = time.perf_counter()
L.learn(env, a.timesteps, log_interval=max(1, a.timesteps // a.envs // 10), callback=cb)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
if a.save_dir:
    os.makedirs(a.save_dir, exist_ok=True)
    L.save(os.path.join(a.save_dir, "SAC_%d.zip" % L.num_timesteps))
print(json.dumps({"envs": a.envs, "timesteps": a.timesteps, "seconds": dt, "env_steps_per_s": a.timesteps / dt,
                  "updates": L.n_updates, "updates_per_s": L.n_updates / dt, "cuda_graph": L._graph is not None, "log": log}))
