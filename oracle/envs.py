"""Oracle env semantics (TEST INFRASTRUCTURE): numpy restatement of the reference's two gym envs on top
of the C oracle (`oracle.MjLike`), structured like the reference (Python env over a C engine).

Follows /root/reference/tr_env/tr_env/envs/tr_env.py (step :327-527, _get_obs :529-646, reset_model
:709-872, is_healthy :306-320) and /root/reference/tensegrity_env/tensegrity_env/envs/tensegrity_env.py
(step :291-410, _get_obs :412-430, reset_model :433-512), plus gym 0.26.2 MujocoEnv plumbing
(reset = mj_resetData + reset_model, set_state = write + mj_forward, do_simulation = ctrl write +
mj_step x frame_skip + mj_rnePostConstraint) and the TimeLimit(5000) of the registrations.

The reference draws its reset randomness from unseeded numpy; here the draws are an explicit argument
(`draws[10]`: pose u01, heading u01, 6 standard normals, waypoint length u01, waypoint yaw u01) so the
CUDA path can be checked on identical inputs.  PARITY UNPINNED with respect to MuJoCo itself.
"""
from __future__ import annotations

from collections import deque

import numpy as np
from scipy.spatial.transform import Rotation

from . import oracle as O
from tensegrity_rl_b200 import model as M


def wrap_pi(t):
    while t > np.pi:
        t -= 2 * np.pi
    while t <= -np.pi:
        t += 2 * np.pi
    return t


class OracleEnv:
    def __init__(self, xml_file="flat", env_kind="tr_env", max_episode_steps=5000, **kwargs):
        self.mj = O.MjLike(xml_file)
        self.cfg = M.env_config(self.mj.md, env_kind=env_kind, max_episode_steps=max_episode_steps, **kwargs)
        self.legacy = self.cfg.env_kind == M.ENV_LEGACY
        self.task = {v: k for k, v in M.TASKS.items()}[self.cfg.task]
        self.dt = self.mj.md["timestep"] * self.cfg.frame_skip
        self.heading = deque()
        self.reset_psi = 0.0
        self.waypt = np.zeros(2)
        self.oripoint = np.zeros(2)
        self.xvel = self.yvel = 1.0
        self.step_num = 0
        self.elapsed = 0
        self.poses = np.ctypeslib.as_array(self.cfg.reset_pose).reshape(M.NPOSE, M.NQ)

    # ---- plumbing
    def do_simulation(self, ctrl):
        self.mj.ctrl[:] = ctrl
        self.mj.step(self.cfg.frame_skip)
        self.mj.rne_post_constraint()

    def com_xy(self):
        return self.mj.xpos[:, :2].mean(axis=0)

    def left_right(self):
        left = (self.mj.sphere_pos(0) + self.mj.sphere_pos(2) + self.mj.sphere_pos(4)) / 3
        right = (self.mj.sphere_pos(1) + self.mj.sphere_pos(3) + self.mj.sphere_pos(5)) / 3
        return left, right

    def psi(self):
        left, right = self.left_right()
        v = left - right
        return np.arctan2(-v[0], v[1])

    # ---- observation noise (tr_env.py:552-644); z = the standard-normal draws, one per noisy component, in
    # observation order [cap positions 18, cap velocities 18 (if used), tendon lengths 9]
    def noisy_obs(self, obs, z):
        c = self.cfg
        n = 27 + (18 if c.use_cap_velocity else 0)
        z = np.asarray(z, np.float64)[:n]
        out = np.array(obs, np.float64)
        sp, st = c.obs_noise_cap_pos_stdev, c.obs_noise_tendon_stdev
        out[:n - 9] = z[:n - 9] * sp + out[:n - 9]          # cap positions :554-565, cap velocities :606-617 (same stdev)
        out[n - 9:n] = z[n - 9:] * st + out[n - 9:n]         # tendon lengths :575-576
        if self.task in ("tracking", "aiming"):                # :626-639
            centre_noise = out[:18].reshape(6, 3).sum(axis=0) / 6
            tv = obs[n:n + 2] - centre_noise[:2]
            d = tv / np.linalg.norm(tv)
            out[n:n + 3] = [tv[0], tv[1], np.arctan2(d[1], d[0])]
        return out

    def get_obs(self):
        mj = self.mj
        if self.legacy:
            quats = [Rotation.from_matrix(mj.geom_xmat[b, 0].reshape(3, 3)).as_quat() for b in range(3)]
            return np.concatenate(quats + [mj.qvel.copy(), mj.ten_length.copy()])
        caps = np.array([mj.sphere_pos(k) for k in range(6)])
        centre = caps.sum(axis=0) / 6
        parts = [(caps - centre).reshape(-1)]
        if self.cfg.use_cap_velocity:
            v = []
            for k in range(6):
                b = k // 2
                lin, ang = mj.qvel[6 * b:6 * b + 3], mj.qvel[6 * b + 3:6 * b + 6]
                v.append(lin + np.cross(ang, caps[k] - mj.xpos[b]))
            parts.append(np.concatenate(v))
        parts.append(mj.ten_length.copy())
        if self.task in ("tracking", "aiming"):
            tv = self.waypt - centre[:2]
            d = tv / np.linalg.norm(tv)
            parts.append(np.array([tv[0], tv[1], np.arctan2(d[1], d[0])]))
        elif self.task == "vel_track":
            parts.append(np.array([0.5 * np.cos(self.reset_psi), 0.5 * np.sin(self.reset_psi), 0.0]))
        return np.concatenate(parts)

    def ditch(self, xy):
        c = self.cfg
        pv = self.waypt - self.oripoint
        dp = np.linalg.norm(pv)
        pn = pv / dp
        tv = self.waypt - xy
        along = np.dot(tv, pn)
        bias = np.linalg.norm(tv - along * pn)
        a = c.ditch_reward_max * (1.0 - abs(along) / dp) * np.exp(-bias ** 2 / (2 * c.ditch_reward_stdev ** 2))
        b = c.waypt_reward_amplitude * np.exp(-np.linalg.norm(xy - self.waypt) ** 2 / (2 * c.waypt_reward_stdev ** 2))
        return a + b

    # ---- step
    def step(self, action, internal=False):
        c, mj, dt = self.cfg, self.mj, self.dt
        action = np.asarray(action, np.float64)
        xy0 = self.com_xy()
        psi0 = self.psi()
        if self.legacy:
            ctrl = action
        else:
            last = mj.ctrl.copy()
            ctrl = last + 1 * (action - last) * dt
        self.do_simulation(ctrl)
        xy1 = self.com_xy()
        self.xvel, self.yvel = (xy1 - xy0) / dt
        left, right = self.left_right()
        psi1 = np.arctan2(-(left - right)[0], (left - right)[1])
        if self.legacy and self.task == "turn":
            psi1 = np.arctan2((right - left)[1], (right - left)[0])
        ten6 = mj.ten_length[:6]
        cc = c.ctrl_cost_weight * (np.sum(np.square(action)) if self.legacy else np.sum(np.square(action + 0.5 - ten6)))
        ctrl_cost = cc
        fwd = 0.0
        healthy = c.healthy_reward if c.terminate_when_unhealthy else 0.0
        state = np.concatenate([mj.qpos, mj.qvel])
        finite = np.isfinite(state).all()
        h_turn = finite and np.any(np.abs(mj.qvel) > 0.1)
        h_lin = finite and (abs(self.xvel) > 1e-4 or abs(self.yvel) > 1e-4)
        is_healthy = h_lin
        extra = False
        d = c.reward_delay_steps
        psi_info = psi1
        if self.task == "turn":
            is_healthy = h_turn
            self.heading.append(psi1)
            if len(self.heading) > d:
                old = self.heading.popleft()
                pa = psi1
                if pa < -np.pi / 2 and old > np.pi / 2:
                    pa = 2 * np.pi + pa
                elif pa > np.pi / 2 and old < -np.pi / 2:
                    pa = -2 * np.pi + pa
                psi_info = pa
                fwd = (pa - old) / (dt * d) * c.desired_direction
            else:
                fwd, ctrl_cost = 0.0, 0.0
        elif self.task == "straight":
            dxy = xy1 - xy0
            pd = abs(np.arctan2(dxy[1], dxy[0]) - self.reset_psi)
            fwd = c.desired_direction * (np.sqrt(dxy[0] ** 2 + dxy[1] ** 2) * np.cos(pd) / dt)
        elif self.task == "aiming":
            is_healthy = h_turn
            td = self.waypt - xy0
            td = td / np.linalg.norm(td)
            new = wrap_pi(np.arctan2(td[1], td[0]) - psi1)
            self.heading.append(new)
            if len(self.heading) > d:
                old = self.heading.popleft()
                fwd = -(abs(new) - abs(old)) / (dt * d) * c.yaw_reward_weight
            healthy = 0.0
            extra = self.step_num > 1000
        elif self.task == "tracking":
            fwd = self.ditch(xy1) - self.ditch(xy0)
            healthy = 0.0
            extra = self.step_num > 1000
        else:  # vel_track
            ang = wrap_pi(psi1 - psi0) / dt
            cmd = np.array([0.5 * np.cos(self.reset_psi), 0.5 * np.sin(self.reset_psi), 0.0])
            le = np.linalg.norm(np.array([self.xvel, self.yvel]) - cmd[:2])
            fwd = 1.0 * np.exp(-5.0 * le ** 2) + 0.5 * np.exp(-7.0 * (ang - cmd[2]) ** 2)
        terminated = (not is_healthy) if c.terminate_when_unhealthy else False
        if extra:
            terminated = True
        if np.any(mj.cfrc_ext > c.kill_force) or np.any(mj.cfrc_ext < -c.kill_force):
            terminated = True
        costs, info_ctrl = ctrl_cost, -ctrl_cost
        if c.use_contact_forces:   # tr_env.py:292-304, 513-516
            lo, hi = c.contact_force_range[0], c.contact_force_range[1]
            contact_cost = c.contact_cost_weight * np.sum(np.square(np.clip(mj.cfrc_ext, lo, hi)))
            costs += contact_cost
            info_ctrl = -contact_cost
        reward = fwd + healthy - costs
        self.step_num += 1
        obs = self.get_obs()
        truncated = False
        if not internal:
            self.elapsed += 1
            truncated = c.max_episode_steps > 0 and self.elapsed >= c.max_episode_steps
        barforce = sum(np.linalg.norm(mj.contact_force(i)[:3]) for i, k in enumerate(mj.contacts()) if k.geom1 != 0)
        info = dict(reward_forward=fwd, reward_ctrl=info_ctrl, reward_survive=healthy, x_position=xy1[0],
                    y_position=xy1[1], psi=psi_info, x_velocity=self.xvel, y_velocity=self.yvel,
                    tendon_length=mj.ten_length.copy(), total_bar_contact=barforce,
                    max_cfrc=float(np.abs(mj.cfrc_ext).max()))
        return obs, float(reward), bool(terminated), bool(truncated), info

    # ---- reset
    def reset(self, draws, noise=None):
        """noise: (21 uniform(-1, 1) values, 18 standard normals) for reset_noise_scale (tr_env.py:734-743)"""
        c, mj = self.cfg, self.mj
        u = np.asarray(draws, np.float64)
        mj.reset_data()
        idx = min(max(int(np.floor(u[0] * c.npose)), 0), c.npose - 1)
        qpos = self.poses[idx].copy()
        qvel = np.zeros(18)
        if c.reset_noise_scale > 0:
            qpos = qpos + c.reset_noise_scale * np.asarray(noise[0])
            qvel = c.reset_noise_scale * np.asarray(noise[1])
        if not self.legacy:
            mj.set_state(qpos, qvel)
        if (not self.legacy and self.task in ("turn", "tracking", "aiming")) or (self.legacy and self.task == "turn"):
            mj.set_state(qpos, qvel)
        theta = c.min_reset_heading + u[1] * (c.max_reset_heading - c.min_reset_heading)
        Rz = np.array([[np.cos(theta), -np.sin(theta), 0], [np.sin(theta), np.cos(theta), 0], [0, 0, 1]])
        new = []
        for b in range(3):
            p, q = qpos[7 * b:7 * b + 3], qpos[7 * b + 3:7 * b + 7]
            q = q / np.linalg.norm(q)
            qz = np.array([np.cos(theta / 2), 0, 0, np.sin(theta / 2)])
            new += list(Rz @ p) + list(M.quat_mul(qz, q))
        mj.set_state(np.array(new), qvel)
        tend = np.clip(u[2:8] * c.tendon_reset_stdev + c.tendon_reset_mean, c.tendon_min_length, c.tendon_max_length)
        if self.legacy:
            for _ in range(c.warmup_steps):
                self.step(tend, internal=True)
        else:
            for _ in range(c.warmup_steps):
                self.do_simulation(tend)
        left, right = self.left_right()
        self.reset_psi = np.arctan2(-(left - right)[0], (left - right)[1])
        lo, hi = c.waypt_range[0], c.waypt_range[1]
        if not self.legacy and self.task == "tracking":
            self.oripoint = np.array([(left[0] + right[0]) / 2, (left[1] + right[1]) / 2])
            length = lo + u[8] * (hi - lo)
            yaw = c.waypt_angle_range[0] + u[9] * (c.waypt_angle_range[1] - c.waypt_angle_range[0]) + self.reset_psi
            if c.is_test:
                length = 0.5 * hi + 0.5 * lo
                yaw = (0.5 * c.waypt_angle_range[1] + 0.5 * c.waypt_angle_range[0]) + self.reset_psi
            self.waypt = self.oripoint + length * np.array([np.cos(yaw), np.sin(yaw)])
        elif not self.legacy and self.task == "aiming":
            self.oripoint = np.array([left[0] + right[0] / 2, (left[1] + right[1]) / 2])
            length = lo + u[8] * (hi - lo)
            yaw = -np.pi + u[9] * 2 * np.pi + self.reset_psi
            if c.is_test:
                length = 0.5 * hi + 0.5 * lo
                yaw = (0.75 * np.pi + 0.25 * (-np.pi)) + self.reset_psi
            self.waypt = self.oripoint + length * np.array([np.cos(yaw), np.sin(yaw)])
            if c.is_test:
                self.waypt = np.array([0.0, 0.0])
        self.step_num = 0
        if not self.legacy and self.task in ("turn", "aiming"):
            for _ in range(c.reward_delay_steps):
                self.step(tend, internal=True)
        self.elapsed = 0
        return self.get_obs()
