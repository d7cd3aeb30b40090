"""Policy rollouts of BASELINE configs 3-5 on the GPU (device resident), optional oracle twin on the host.
  python tools/run_rollouts.py --config forward_flat --envs 4096 --steps 1000 [--oracle-envs 8]
Prints one JSON line per config: throughput, episode / displacement / yaw statistics (all ranks reduced)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

CONFIGS = {
    # name: (xml, env, kwargs, policy)
    "forward_flat": ("flat", "tensegrity_env", dict(desired_action="straight", desired_direction=1), "forward"),
    "forward_uneven": ("uneven", "tensegrity_env", dict(desired_action="straight", desired_direction=1), "forward"),   # BASELINE config 3
    "backward_flat": ("flat", "tensegrity_env", dict(desired_action="straight", desired_direction=-1), "backward"),
    "yaw_ccw_flat": ("flat", "tensegrity_env", dict(desired_action="turn", desired_direction=1, terminate_when_unhealthy=False), "yaw_CCW"),
    "yaw_cw_flat": ("flat", "tensegrity_env", dict(desired_action="turn", desired_direction=-1, terminate_when_unhealthy=False), "yaw_CW"),
    "track_flat": ("flat", "tr_env", dict(desired_action="tracking"), "traj_track"),                                   # BASELINE config 4
    "aim_ccw_flat": ("flat", "tr_env", dict(desired_action="aiming"), "traj_ccw"),
    "aim_cw_flat": ("flat", "tr_env", dict(desired_action="aiming"), "traj_cw"),
    "track_uneven": ("uneven", "tr_env", dict(desired_action="tracking"), "traj_track"),
}


def oracle_rollout(name, n_envs, steps, deterministic, seed=0):
    """the same policy through the numpy + C oracle env on the host (small sample, distribution check)."""
    from oracle.envs import OracleEnv
    from tensegrity_rl_b200.policy import SacActor
    from tensegrity_rl_b200.rollout import EpisodeStats, summarize
    xml, env, kw, pol = CONFIGS[name]
    actor = SacActor(pol, device="cpu", seed=seed)
    rng = np.random.default_rng(seed)
    st = EpisodeStats(n_envs, torch.device("cpu"))
    envs = [OracleEnv(xml, env, **kw) for _ in range(n_envs)]
    obs = np.stack([e.reset(np.concatenate([rng.uniform(0, 1, 2), rng.standard_normal(6), rng.uniform(0, 1, 2)])) for e in envs])
    for k in range(steps):
        a = actor(torch.as_tensor(obs, dtype=torch.float32), deterministic).double().numpy()
        info = torch.zeros(n_envs, 32, dtype=torch.float64)
        rew, done = np.zeros(n_envs), np.zeros(n_envs, np.uint8)
        for i, e in enumerate(envs):
            o, r, term, trunc, inf = e.step(a[i])
            rew[i], done[i] = r, term or trunc
            info[i, 3], info[i, 4], info[i, 5], info[i, 6], info[i, 7], info[i, 31] = inf["x_position"], inf["y_position"], inf["psi"], inf["x_velocity"], inf["y_velocity"], e.reset_psi
            if done[i]:
                o = e.reset(np.concatenate([rng.uniform(0, 1, 2), rng.standard_normal(6), rng.uniform(0, 1, 2)]))
            obs[i] = o
        st.update(torch.as_tensor(rew), torch.as_tensor(done), info, envs[0].dt)
    st.flush_open_episodes(info)
    return summarize(st.totals.numpy())


def gpu_rollout(name, n_envs, steps, deterministic, seed=0, rank=0, world=1, local_rank=0):
    from tensegrity_rl_b200 import TensegrityVecEnv, SacActor
    from tensegrity_rl_b200.rollout import rollout
    xml, env, kw, pol = CONFIGS[name]
    dev = torch.device("cuda", local_rank)
    v = TensegrityVecEnv(n_envs, xml_file=xml, env=env, device=local_rank, seed=seed, env_id_base=rank * n_envs,
                         auto_reset=True, reset_pool="auto", **kw)
    actor = SacActor(pol, device=dev, seed=seed + rank)
    v.reset_tensor()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s = rollout(v, actor, steps, deterministic)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if world > 1:   # device-timed, max over ranks
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    s["env_steps_per_s"] = world * n_envs * steps / ms * 1e3
    s["ncon"], s["niter_per_substep"] = float(v.info[:, 19].mean()), float(v.info[:, 20].mean()) / 20
    s["overflow"], s["bad"] = float(v.info[:, 28].sum()), float(v.info[:, 29].sum())
    v.close()
    return s


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="forward_flat")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--oracle-envs", type=int, default=0)
    ap.add_argument("--oracle-steps", type=int, default=0)
    ap.add_argument("--deterministic", action="store_true")
    args = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(lr)
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)   # NCCL's version banner goes to stdout: park it on stderr
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
            dist.barrier(device_ids=[lr]); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    for name in (CONFIGS if args.config == "all" else args.config.split(",")):
        out = {"config": name, "envs_per_gpu": args.envs, "steps": args.steps, "n_gpus": world,
               "policy": CONFIGS[name][3], "stochastic": not args.deterministic}
        out["gpu"] = gpu_rollout(name, args.envs, args.steps, args.deterministic, rank=rank, world=world, local_rank=lr)
        if args.oracle_envs and rank == 0:
            t0 = time.time()
            out["oracle"] = oracle_rollout(name, args.oracle_envs, args.oracle_steps or args.steps, args.deterministic)
            out["oracle"]["seconds"] = time.time() - t0
        if rank == 0:
            print(json.dumps(out), flush=True)
