"""Device-resident rollouts: batched SAC actor -> ctrl tensor -> tsg_step, nothing leaves the GPU per step.

Covers SURVEY 8(f) rank 1-2: the evaluation loops of run.py (`test` :103-190, `test3` :192-310,
`tracking_test` :312-365) for N envs at once, the 3-policy waypoint selector of `test3` as a batched mask,
episode / displacement statistics reduced across ranks with ONE small all_reduce (NCCL on GPUs, gloo in
the CPU tests), and trace files with the reference's names so plot_*.py and the notebooks keep working.

Launch overhead note: a step of >= 4096 envs is >= 10 ms of kernel time against ~50 us of launches for
actor + step, so the loop is not captured in a CUDA graph -- there is nothing to win.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import lib as _lib

STAT_NAMES = ("episodes", "return_sum", "length_sum", "disp_sum", "disp_sq_sum", "yaw_sum", "yaw_sq_sum", "env_steps")


def wrap_pi(t):
    """torch tensor version of tr_env._angle_normalize."""
    import torch
    return torch.remainder(t + math.pi, 2 * math.pi) - math.pi


class EpisodeStats:
    """Per-env episode accumulators on the device; `reduce()` sums the 8 statistics over ranks.
    forward displacement = episode COM displacement projected on the reset heading (the `straight` reward
    direction, tr_env.py:407-414); yaw = unwrapped change of psi over the episode."""

    def __init__(self, n, device):
        import torch
        self.t = torch
        z = lambda: torch.zeros(n, dtype=torch.float64, device=device)
        self.ret, self.length, self.x0, self.y0, self.psi_prev, self.yaw = z(), z(), z(), z(), z(), z()
        self.fresh = torch.ones(n, dtype=torch.bool, device=device)   # next step starts an episode
        self.totals = torch.zeros(len(STAT_NAMES), dtype=torch.float64, device=device)

    def update(self, reward, done, info, dt):
        t, I = self.t, _lib.INFO
        x, y, psi = info[:, I["x"]], info[:, I["y"]], info[:, I["psi"]]
        xv, yv = info[:, I["xvel"]], info[:, I["yvel"]]
        # position before this step = after - velocity * dt  (velocities are finite differences of the same COM)
        xs, ys = x - xv * dt, y - yv * dt
        self.x0 = t.where(self.fresh, xs, self.x0)
        self.y0 = t.where(self.fresh, ys, self.y0)
        dpsi = t.where(self.fresh, t.zeros_like(psi), wrap_pi(psi - self.psi_prev))
        self.yaw = t.where(self.fresh, t.zeros_like(psi), self.yaw) + dpsi
        self.ret = t.where(self.fresh, t.zeros_like(reward), self.ret) + reward
        self.length = t.where(self.fresh, t.zeros_like(reward), self.length) + 1
        self.psi_prev = psi.clone()
        d = done.bool()
        rp = info[:, I["reset_psi"]]
        disp = (x - self.x0) * t.cos(rp) + (y - self.y0) * t.sin(rp)
        f = d.to(t.float64)
        self.totals += t.stack([f.sum(), (self.ret * f).sum(), (self.length * f).sum(), (disp * f).sum(),
                                (disp * disp * f).sum(), (self.yaw * f).sum(), (self.yaw * self.yaw * f).sum(),
                                t.tensor(float(reward.numel()), dtype=t.float64, device=reward.device)])
        self.fresh = d

    def flush_open_episodes(self, info):
        """count the still-running episodes as finished (fixed-horizon evaluation)."""
        t, I = self.t, _lib.INFO
        open_ = (~self.fresh).to(t.float64)
        x, y, rp = info[:, I["x"]], info[:, I["y"]], info[:, I["reset_psi"]]
        disp = (x - self.x0) * t.cos(rp) + (y - self.y0) * t.sin(rp)
        self.totals[:7] += t.stack([open_.sum(), (self.ret * open_).sum(), (self.length * open_).sum(), (disp * open_).sum(),
                                    (disp * disp * open_).sum(), (self.yaw * open_).sum(), (self.yaw * self.yaw * open_).sum()])
        self.fresh = t.ones_like(self.fresh)

    def reduce(self, group=None):
        """sum over ranks: the only collective of the path (one 64-byte all_reduce)."""
        import torch.distributed as dist
        tot = self.totals.clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(tot, group=group)
        return summarize(tot.cpu().numpy())


def summarize(tot):
    d = dict(zip(STAT_NAMES, [float(v) for v in tot]))
    n = max(d["episodes"], 1.0)
    d["return_mean"], d["length_mean"] = d["return_sum"] / n, d["length_sum"] / n
    d["disp_mean"] = d["disp_sum"] / n
    d["disp_std"] = math.sqrt(max(d["disp_sq_sum"] / n - d["disp_mean"] ** 2, 0.0))
    d["yaw_mean"] = d["yaw_sum"] / n
    d["yaw_std"] = math.sqrt(max(d["yaw_sq_sum"] / n - d["yaw_mean"] ** 2, 0.0))
    return d


def rollout(env, actor, steps, deterministic=False, stats=None, action_fn=None, flush=True):
    """steps x (actor -> step) on the device.  `actor(obs32) -> ctrl`; `action_fn(env, k)` overrides it
    (e.g. random ctrl).  Returns the reduced statistics dict."""
    import torch
    stats = stats or EpisodeStats(env.num_envs, env.device)
    for k in range(steps):
        ctrl = action_fn(env, k) if action_fn is not None else actor(env.obs32, deterministic)
        obs, rew, done = env.step_tensor(ctrl, want_info=True)
        stats.update(rew, done, env.info, env.dt)
    if flush:
        stats.flush_open_episodes(env.info)
    return stats.reduce()


class WaypointController:
    """Batched form of run.py test3 (:192-310): per env, steer with the ccw / cw policies until the heading
    error is inside (0, pi/15], then track; switch to the next waypoint within 0.2 m.  The env is
    `tr_env` aiming with is_test=True, whose waypoint is the origin so that obs[45:47] = -position."""

    WAYPOINTS = ((0.0, 2.0), (2.0, 0.0), (4.0, 2.0), (4.0, 0.0))

    def __init__(self, env, track, ccw, cw, waypoints=None, threshold=0.2, deterministic=False):
        import torch
        self.t, self.env = torch, env
        self.track, self.ccw, self.cw = track, ccw, cw
        self.wp = torch.tensor(waypoints or self.WAYPOINTS, dtype=torch.float64, device=env.device)
        n = env.num_envs
        self.idx = torch.zeros(n, dtype=torch.long, device=env.device)
        self.turn_open = torch.ones(n, dtype=torch.bool, device=env.device)
        self.hold = torch.ones(n, dtype=torch.bool, device=env.device)   # next action = current tendon lengths
        self.finished = torch.zeros(n, dtype=torch.bool, device=env.device)
        self.threshold, self.deterministic = threshold, deterministic
        self.del_yaw = torch.zeros(n, dtype=torch.float64, device=env.device)

    def action(self, obs):
        t = self.t
        obs = obs.clone()
        wp = self.wp[self.idx.clamp(max=self.wp.shape[0] - 1)]
        pos = -obs[:, 45:47]
        vec = wp - pos
        tgt = t.atan2(vec[:, 1], vec[:, 0])
        caps = obs[:, :18].reshape(-1, 6, 3)
        left, right = caps[:, 0::2].mean(1), caps[:, 1::2].mean(1)
        yaw = t.atan2(right[:, 0] - left[:, 0], left[:, 1] - right[:, 1])
        dy = tgt - yaw
        dy = t.where(dy > math.pi, dy - 2 * math.pi, t.where(dy <= -math.pi, dy + 2 * math.pi, dy))
        self.del_yaw = dy
        use_ccw = (dy > math.pi / 15) & self.turn_open
        use_cw = (dy < 0) & self.turn_open & ~use_ccw
        use_track = ~(use_ccw | use_cw)
        self.turn_open = self.turn_open & ~use_track
        o_turn = obs.clone(); o_turn[:, 45:48] = 0
        o_track = obs.clone()
        o_track[:, 45:47] = vec / vec.norm(dim=1, keepdim=True)
        o_track[:, 47] = tgt
        a = t.where(use_ccw[:, None], self.ccw(o_turn.float(), self.deterministic).double(),
                    t.where(use_cw[:, None], self.cw(o_turn.float(), self.deterministic).double(),
                            self.track(o_track.float(), self.deterministic).double()))
        a = t.where(self.hold[:, None], obs[:, 36:42], a)     # env.step(tendon_loop_init) at each new waypoint
        self.hold = t.zeros_like(self.hold)
        return a

    def after_step(self, info):
        t, I = self.t, _lib.INFO
        xy = t.stack([info[:, I["x"]], info[:, I["y"]]], 1)
        wp = self.wp[self.idx.clamp(max=self.wp.shape[0] - 1)]
        reached = ((xy - wp).norm(dim=1) < self.threshold) & ~self.finished
        self.idx = self.idx + reached.long()
        self.finished = self.idx >= self.wp.shape[0]
        self.hold = reached & ~self.finished
        self.turn_open = self.turn_open | reached
        return reached


# ---------------------------------------------------------------------------------- trace files of run.py
def run_test(env, actor, saved_data_dir, simulation_seconds=30, deterministic=False, env_index=0):
    """run.py `test` (:103-190) for one env of a batch: same loop, same 11 .npy files."""
    import torch
    os.makedirs(saved_data_dir, exist_ok=True)
    I = _lib.INFO
    env.reset_tensor()
    rows = {k: [] for k in ("action", "tendon", "observed_tendon", "cap_posi", "observed_cap_posi", "total_bar_contact",
                            "reward_forward", "reward_ctrl", "waypt", "x_pos", "y_pos")}
    extra = 500
    for _ in range(int(simulation_seconds / env.dt)):
        a = actor(env.obs32, deterministic)
        obs, rew, done = env.step_tensor(a.double(), want_info=True, auto_reset=False)
        o, inf = obs[env_index].cpu().numpy(), env.info[env_index].cpu().numpy()
        rows["action"].append(a[env_index].cpu().numpy()); rows["tendon"].append(inf[I["ten"]:I["ten"] + 9])
        rows["observed_tendon"].append(o[-9:] if env.obs_dim in (27, 45, 39) else o[36:45])
        rows["cap_posi"].append(o[:18]); rows["observed_cap_posi"].append(o[:18])
        rows["total_bar_contact"].append(inf[I["barforce"]]); rows["reward_forward"].append(inf[I["rew_fwd"]])
        rows["reward_ctrl"].append(inf[I["rew_ctrl"]]); rows["waypt"].append(inf[I["waypt"]:I["waypt"] + 2])
        rows["x_pos"].append(inf[I["x"]]); rows["y_pos"].append(inf[I["y"]])
        if bool(done[env_index]):
            extra -= 1
            if extra < 0:
                break
    for k, v in rows.items():
        np.save(os.path.join(saved_data_dir, k + "_data.npy"), np.array(v))
    return {k: np.array(v) for k, v in rows.items()}


def run_test3(env, track, ccw, cw, saved_data_dir, simulation_seconds=30, deterministic=False, env_index=0, waypoints=None):
    """run.py `test3` (:192-310) for one env of a batch (`tr_env` aiming, is_test=True): the 3-policy waypoint loop
    and its four files -- waypt_data.npy (the waypoint list), x_pos_data.npy / y_pos_data.npy (position after every
    policy step), del_yaw_data.npy (heading error before every policy step).  As in the reference the `env.step(
    tendon_loop_init)` taken at each new waypoint is neither counted nor recorded."""
    import torch
    os.makedirs(saved_data_dir, exist_ok=True)
    I, e = _lib.INFO, env_index
    env.reset_tensor()
    ctl = WaypointController(env, track, ccw, cw, waypoints=waypoints, deterministic=deterministic)
    xs, ys, dyaw = [], [], []
    counter, extra, iters = 0, 500, int(simulation_seconds / env.dt)
    while counter < iters and extra >= 0 and not bool(ctl.finished[e]):
        hold = bool(ctl.hold[e])
        a = ctl.action(env.obs)
        obs, rew, done = env.step_tensor(a, want_info=True, auto_reset=False)
        ctl.after_step(env.info)
        if hold:
            continue
        inf = env.info[e].cpu().numpy()
        dyaw.append(float(ctl.del_yaw[e])); xs.append(inf[I["x"]]); ys.append(inf[I["y"]])
        counter += 1
        if bool(done[e]):
            extra -= 1
    out = {"waypt": ctl.wp.cpu().numpy(), "x_pos": np.array(xs), "y_pos": np.array(ys), "del_yaw": np.array(dyaw)}
    for k, v in out.items():
        np.save(os.path.join(saved_data_dir, k + "_data.npy"), v)
    return out


def run_tracking_test(env, actor, saved_data_dir, simulation_seconds=30, episode_num=10, deterministic=False):
    """run.py `tracking_test` (:312-365): `episode_num` tracking episodes, here run side by side on the first
    `episode_num` envs of the batch, and its three files -- waypt_data.npy, xy_pos_data.npy, oripoint_data.npy, all
    relative to the episode's origin point and rotated so that the waypoint lies on the +x axis.  An episode that is
    done keeps stepping for up to 500 more steps without a reset, as in the reference; what is recorded is the info of
    the last step the reference loop would have executed."""
    import torch
    os.makedirs(saved_data_dir, exist_ok=True)
    I, n = _lib.INFO, episode_num
    if env.num_envs < n:
        raise ValueError("run_tracking_test needs num_envs >= episode_num (one episode per env)")
    env.reset_tensor()
    dev = env.device
    extra = torch.full((n,), 500, dtype=torch.long, device=dev)
    running = torch.ones(n, dtype=torch.bool, device=dev)
    last = torch.zeros(n, 6, dtype=torch.float64, device=dev)   # oripoint (2), waypt (2), xy (2)
    for _ in range(int(simulation_seconds / env.dt)):
        a = actor(env.obs32, deterministic)
        obs, rew, done = env.step_tensor(a.double(), want_info=True, auto_reset=False)
        inf = env.info[:n]
        row = torch.cat([inf[:, I["ori"]:I["ori"] + 2], inf[:, I["waypt"]:I["waypt"] + 2], inf[:, I["x"]:I["x"] + 1], inf[:, I["y"]:I["y"] + 1]], 1)
        last = torch.where(running[:, None], row, last)
        extra = extra - (done[:n].bool() & running).long()
        running = running & (extra >= 0)
        if not bool(running.any()):
            break
    last = last.cpu().numpy()
    ori, way, xy = last[:, 0:2], last[:, 2:4] - last[:, 0:2], last[:, 4:6] - last[:, 0:2]
    for i in range(n):
        ang = np.arctan2(way[i, 1], way[i, 0])
        rot = np.array([[np.cos(ang), np.sin(ang)], [np.sin(ang), -np.cos(ang)]])
        way[i], xy[i] = rot @ way[i], rot @ xy[i]
    out = {"waypt": way, "xy_pos": xy, "oripoint": ori - ori}
    for k, v in out.items():
        np.save(os.path.join(saved_data_dir, k + "_data.npy"), v)
    return out
