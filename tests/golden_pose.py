"""Reconstruction of a full simulator state (qpos, qvel, waypoint) from a checkpoint's `_last_obs` (a vector the
REFERENCE's MuJoCo run produced; tests/golden/last_obs.json).  The observation over-determines the pose: the few pose
parameters it does not contain are fitted so that the 9 tendon lengths are reproduced (tests/test_golden_last_obs.py).
Feeding the reconstructed state through set_state -> mj_forward -> _get_obs of an implementation must then return the
golden vector itself: a check of kinematics, site / tendon tables, geom frames and the observation code against
reference-produced numbers that runs through the implementation's own code path (oracle, and CUDA on the GPU)."""
import json
import os

import numpy as np
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

from tensegrity_rl_b200 import model as M

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "last_obs.json")))
TASK_OF = {"traj_track": "tracking", "traj_ccw": "aiming", "traj_cw": "aiming"}
HALF = 0.688   # end-cap centres sit at +-0.688 on the bar axis


def _lengths(md, pos, R):
    s, tb = np.array(md["ten_site"]), np.array(md["ten_body"])
    return np.array([np.linalg.norm((pos[tb[t, 1]] + R[tb[t, 1]] @ s[t, 1]) - (pos[tb[t, 0]] + R[tb[t, 0]] @ s[t, 0]))
                     for t in range(9)])


def _best(fun, sampler, n=60):
    best = None
    for k in range(n):
        r = least_squares(fun, sampler(np.random.default_rng(k)), xtol=1e-15, ftol=1e-15, gtol=1e-15)
        if best is None or r.cost < best.cost:
            best = r
        if np.abs(best.fun).max() < 1e-13:
            break
    return best


def _quat_wxyz(R):
    x, y, z, w = Rotation.from_matrix(R).as_quat()
    return np.array([w, x, y, z])


def state_from_tr_obs(obs, md, centre=(0.3, -0.2, 0.9)):
    """tr_env 48-dim observation -> qpos[21], qvel[18], waypt[2].  The absolute position of the 6-cap centroid is not
    observable (positions are centroid-relative): `centre` is a free choice."""
    obs = np.asarray(obs, np.float64)
    caps = obs[:18].reshape(6, 3) + np.asarray(centre)
    vcap, ten = obs[18:36].reshape(6, 3), obs[36:45]
    cen = [(caps[2 * b] + caps[2 * b + 1]) / 2 for b in range(3)]
    z = [(caps[2 * b] - caps[2 * b + 1]) / np.linalg.norm(caps[2 * b] - caps[2 * b + 1]) for b in range(3)]
    uv = []
    for b in range(3):
        t = np.array([1.0, 0, 0]) if abs(z[b][0]) < 0.9 else np.array([0, 1.0, 0])
        u = np.cross(z[b], t); u /= np.linalg.norm(u)
        uv.append((u, np.cross(z[b], u)))

    def rots(phi):
        R = []
        for b in range(3):
            x = np.cos(phi[b]) * uv[b][0] + np.sin(phi[b]) * uv[b][1]
            R.append(np.stack([x, np.cross(z[b], x), z[b]], 1))
        return R
    best = _best(lambda phi: _lengths(md, cen, rots(phi)) - ten, lambda g: g.uniform(-np.pi, np.pi, 3))
    R = rots(best.x)
    qpos, qvel = np.zeros(21), np.zeros(18)
    for b in range(3):
        qpos[7 * b:7 * b + 3] = cen[b]
        qpos[7 * b + 3:7 * b + 7] = _quat_wxyz(R[b])
        # cap velocity = v + w x r with the BODY-frame angular velocity used as if it were world-frame (tr_env.py:599-604)
        r = caps[2 * b] - cen[b]
        qvel[6 * b:6 * b + 3] = (vcap[2 * b] + vcap[2 * b + 1]) / 2
        d = (vcap[2 * b] - vcap[2 * b + 1]) / 2
        qvel[6 * b + 3:6 * b + 6] = np.cross(r, d) / np.dot(r, r)
    waypt = np.asarray(centre)[:2] + obs[45:47]
    return qpos, qvel, waypt, float(np.abs(best.fun).max())


def state_from_legacy_obs(obs, md):
    """tensegrity_env 39-dim observation -> qpos[21], qvel[18] (bar 0 at the origin: absolute positions are not observed)."""
    obs = np.asarray(obs, np.float64)
    R = [Rotation.from_quat(obs[4 * b:4 * b + 4]).as_matrix() @ np.diag([-1.0, -1.0, 1.0]) for b in range(3)]
    ten = obs[30:39]
    best = _best(lambda p: _lengths(md, [np.zeros(3), p[:3], p[3:]], R) - ten, lambda g: g.uniform(-0.5, 0.5, 6), n=120)
    pos = [np.zeros(3), best.x[:3], best.x[3:]]
    qpos = np.zeros(21)
    for b in range(3):
        qpos[7 * b:7 * b + 3] = pos[b] + np.array([0, 0, 1.0])
        qpos[7 * b + 3:7 * b + 7] = _quat_wxyz(R[b])
    return qpos, obs[12:30].copy(), float(np.abs(best.fun).max())
