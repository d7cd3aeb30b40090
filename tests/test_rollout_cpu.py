"""Host logic of the rollout layer on CPU: episode statistics, the batched waypoint selector and the
world_size-2 statistics reduction over gloo (the N > 1 path of bench.py / rollout.py)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensegrity_rl_b200.rollout import EpisodeStats, STAT_NAMES, WaypointController, summarize, wrap_pi


def _fake_stream(st, n, steps, seed, done_every=7):
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(n, dtype=torch.float64); y = torch.zeros(n, dtype=torch.float64); psi = torch.zeros(n, dtype=torch.float64)
    for k in range(steps):
        dx, dy = 0.01 * torch.rand(n, generator=g, dtype=torch.float64), 0.004 * torch.randn(n, generator=g, dtype=torch.float64)
        x, y, psi = x + dx, y + dy, wrap_pi(psi + 0.05 * torch.randn(n, generator=g, dtype=torch.float64))
        info = torch.zeros(n, 32, dtype=torch.float64)
        info[:, 3], info[:, 4], info[:, 5], info[:, 6], info[:, 7], info[:, 31] = x, y, psi, dx / 0.02, dy / 0.02, 0.3
        done = ((torch.arange(n) + k) % done_every == 0).to(torch.uint8)
        st.update(torch.rand(n, generator=g, dtype=torch.float64), done, info, 0.02)
    st.flush_open_episodes(info)


def test_episode_stats_single_process():
    st = EpisodeStats(64, torch.device("cpu"))
    _fake_stream(st, 64, 40, 0)
    s = summarize(st.totals.numpy())
    assert s["env_steps"] == 64 * 40 and s["episodes"] > 64 * 40 / 7
    assert s["length_sum"] == pytest.approx(64 * 40)          # every env step belongs to exactly one episode
    assert 0 < s["disp_mean"] < 0.1 and s["disp_std"] > 0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 32
    st = EpisodeStats(n, torch.device("cpu"))
    _fake_stream(st, n, 30, seed=100 + rank)
    red = st.reduce()
    if rank == 0:
        torch.save({"reduced": red, "local": st.totals.clone()}, out)
    dist.destroy_process_group()


def test_stats_allreduce_world2_gloo(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)["reduced"]
    expect = torch.zeros(len(STAT_NAMES), dtype=torch.float64)
    for rank in range(2):
        st = EpisodeStats(32, torch.device("cpu"))
        _fake_stream(st, 32, 30, seed=100 + rank)
        expect += st.totals
    for k, v in zip(STAT_NAMES, expect.tolist()):
        assert got[k] == pytest.approx(v, rel=1e-12)
    assert got["env_steps"] == 2 * 32 * 30


class _FakeEnv:
    def __init__(self, n):
        self.num_envs, self.device = n, torch.device("cpu")


def test_waypoint_controller_matches_scalar_logic():
    """the batched mask selects the same policy per env as run.py:257-276's if / elif / else."""
    n = 256
    g = torch.Generator().manual_seed(3)
    env = _FakeEnv(n)
    mk = lambda v: (lambda obs, det=False: torch.full((obs.shape[0], 6), v, dtype=torch.float32))
    ctl = WaypointController(env, mk(1.0), mk(2.0), mk(3.0))
    ctl.hold[:] = False
    ctl.turn_open = torch.rand(n, generator=g) > 0.3
    open_before = ctl.turn_open.clone()
    obs = torch.randn(n, 48, generator=g, dtype=torch.float64)
    a = ctl.action(obs)
    for e in range(n):
        o = obs[e].numpy()
        pos = -o[45:47]
        vec = np.array([0.0, 2.0]) - pos
        tgt = math.atan2(vec[1], vec[0])
        caps = o[:18].reshape(6, 3)
        left, right = caps[0::2].mean(0), caps[1::2].mean(0)
        yaw = math.atan2(right[0] - left[0], left[1] - right[1])
        dy = tgt - yaw
        if dy > math.pi: dy -= 2 * math.pi
        elif dy <= -math.pi: dy += 2 * math.pi
        if dy > math.pi / 15 and open_before[e]: want, still = 2.0, True
        elif dy < 0 and open_before[e]: want, still = 3.0, True
        else: want, still = 1.0, False
        assert a[e, 0].item() == want and bool(ctl.turn_open[e]) == still
    info = torch.zeros(n, 32, dtype=torch.float64)
    info[:8, 3], info[:8, 4] = 0.05, 1.95          # within 0.2 m of waypoint (0, 2)
    reached = ctl.after_step(info)
    assert reached[:8].all() and not reached[8:].any() and (ctl.idx[:8] == 1).all() and ctl.hold[:8].all()
