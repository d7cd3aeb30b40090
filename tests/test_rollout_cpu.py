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


class _ScriptedEnv:
    """a stand-in for TensegrityVecEnv on the CPU (the product has no CPU path): point robots that move 0.05 m per step
    straight towards a per-env goal, report the info columns the writers read and are done on arrival."""

    def __init__(self, n, obs_dim=48, goal=(3.0, 0.5)):
        self.num_envs, self.device, self.obs_dim, self.dt = n, torch.device("cpu"), obs_dim, 0.02
        self.goal = torch.tensor(goal, dtype=torch.float64).repeat(n, 1) + 0.1 * torch.arange(n, dtype=torch.float64)[:, None]
        self.reset_tensor()

    def _fill(self):
        self.obs = torch.zeros(self.num_envs, self.obs_dim, dtype=torch.float64)
        self.obs[:, 0:18] = torch.tensor([0, 0.3, 0, 0, -0.3, 0] * 3, dtype=torch.float64)    # left caps at +y, right at -y: yaw 0
        self.obs[:, 36:45] = 1.0 + 0.01 * torch.arange(9, dtype=torch.float64)
        self.obs[:, 45:47] = -self.xy
        self.obs32 = self.obs.float()
        self.info = torch.zeros(self.num_envs, 32, dtype=torch.float64)
        self.info[:, 3:5], self.info[:, 24:26], self.info[:, 26:28] = self.xy, self.goal, self.ori
        self.info[:, 8:17] = self.obs[:, 36:45]

    def reset_tensor(self):
        self.xy = torch.zeros(self.num_envs, 2, dtype=torch.float64) + 0.25
        self.ori = self.xy.clone()
        self.target = self.goal.clone()
        self._fill()
        return self.obs

    def step_tensor(self, a, want_info=True, auto_reset=False):
        d = self.target - self.xy
        dist_ = d.norm(dim=1, keepdim=True).clamp(min=1e-9)
        self.xy = self.xy + d / dist_ * torch.minimum(dist_, torch.full_like(dist_, 0.05))
        self._fill()
        done = ((self.goal - self.xy).norm(dim=1) < 1e-6).to(torch.uint8)
        return self.obs, torch.zeros(self.num_envs, dtype=torch.float64), done


def test_trace_writers_produce_the_reference_files(tmp_path):
    """run.py's three evaluation loops write .npy files that plot_*.py and the notebooks read (run.py:180-190, 305-308,
    363-365): names, shapes and the geometry of the tracking_test transform."""
    from tensegrity_rl_b200.rollout import run_test, run_test3, run_tracking_test
    actor = lambda obs, det=False: torch.zeros(obs.shape[0], 6, dtype=torch.float32)
    # ---- test: 11 files, one row per step
    env = _ScriptedEnv(4)
    out = run_test(env, actor, str(tmp_path / "t1"), simulation_seconds=1.0)
    names = {"action": (50, 6), "tendon": (50, 9), "observed_tendon": (50, 9), "cap_posi": (50, 18), "observed_cap_posi": (50, 18),
             "total_bar_contact": (50,), "reward_forward": (50,), "reward_ctrl": (50,), "waypt": (50, 2), "x_pos": (50,), "y_pos": (50,)}
    assert sorted(os.listdir(tmp_path / "t1")) == sorted(k + "_data.npy" for k in names)
    for k, shp in names.items():
        assert np.load(tmp_path / "t1" / (k + "_data.npy")).shape == shp, k
    # ---- tracking_test: three files, one row per episode; waypoint on the +x axis, origin at 0
    env = _ScriptedEnv(8)
    out = run_tracking_test(env, actor, str(tmp_path / "t2"), simulation_seconds=30, episode_num=5)
    assert sorted(os.listdir(tmp_path / "t2")) == ["oripoint_data.npy", "waypt_data.npy", "xy_pos_data.npy"]
    way, xy, ori = (np.load(tmp_path / "t2" / f) for f in ("waypt_data.npy", "xy_pos_data.npy", "oripoint_data.npy"))
    assert way.shape == xy.shape == ori.shape == (5, 2) and np.all(ori == 0)
    assert np.abs(way[:, 1]).max() < 1e-12 and np.all(way[:, 0] > 2.5)        # rotated onto the +x axis
    assert np.abs(xy - way).max() < 1e-6                                      # the scripted robots arrive at their goal
    # ---- test3: four files; the robot is steered by the "track" policy towards the four waypoints (the scripted env
    # moves to the controller's current waypoint)
    env = _ScriptedEnv(3)
    ctl_wp = ((0.5, 0.5), (1.0, 0.25))

    class _Follow(_ScriptedEnv):
        pass
    env = _Follow(3)
    env.goal[:] = torch.tensor(ctl_wp[0], dtype=torch.float64)
    env.reset_tensor()
    orig_step = env.step_tensor

    def step(a, want_info=True, auto_reset=False):
        r = orig_step(a, want_info, auto_reset)
        if (env.xy[0] - env.target[0]).norm() < 0.2:
            env.target[:] = torch.tensor(ctl_wp[1], dtype=torch.float64)
        return r
    env.step_tensor = step
    out = run_test3(env, actor, actor, actor, str(tmp_path / "t3"), simulation_seconds=4.0, waypoints=ctl_wp)
    assert sorted(os.listdir(tmp_path / "t3")) == ["del_yaw_data.npy", "waypt_data.npy", "x_pos_data.npy", "y_pos_data.npy"]
    x, y, dyaw, wp = (np.load(tmp_path / "t3" / f) for f in ("x_pos_data.npy", "y_pos_data.npy", "del_yaw_data.npy", "waypt_data.npy"))
    assert wp.shape == (2, 2) and x.shape == y.shape == dyaw.shape and 4 < len(x) < 200
    assert np.hypot(x[-1] - ctl_wp[1][0], y[-1] - ctl_wp[1][1]) < 0.2         # stopped within the threshold of the last waypoint
    assert np.all(np.abs(dyaw) <= np.pi)
