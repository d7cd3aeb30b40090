"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE -- never imported by the product package).

`MjLike` gives the slice of the mujoco python API the reference envs touch
(data.qpos/qvel/ctrl/ten_length/cfrc_ext, geom/body xpos, mj_step, mj_forward,
mj_resetData, mj_rnePostConstraint) on top of oracle/libtsg_oracle.so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tensegrity_rl_b200 import model as M  # noqa: E402  (model compiler + struct mirrors only)

MAXCON, MAXEFC = 96, 576
d, i32 = C.c_double, C.c_int32
NB, NG, NT, NV, NQ, NA = M.NBAR, M.NGEOM_BAR, M.NTEN, M.NV, M.NQ, M.NACT


class Contact(C.Structure):
    _fields_ = [("dist", d), ("pos", d * 3), ("frame", d * 9), ("geom1", i32), ("geom2", i32),
                ("body1", i32), ("body2", i32), ("exclude", i32), ("efc_address", i32)]


class Data(C.Structure):
    _fields_ = [
        ("qpos", d * NQ), ("qvel", d * NV), ("act", d * NA), ("ctrl", d * NA), ("qacc_warmstart", d * NV), ("time", d),
        ("xpos", (d * 3) * NB), ("xquat", (d * 4) * NB), ("xmat", (d * 9) * NB),
        ("geom_xpos", ((d * 3) * NG) * NB), ("geom_xmat", ((d * 9) * NG) * NB), ("site_xpos", ((d * 3) * 2) * NT),
        ("com_world", d * 3), ("ten_length", d * NT), ("ten_J", (d * NV) * NT),
        ("ten_velocity", d * NT), ("qfrc_passive", d * NV), ("qfrc_bias", d * NV), ("qfrc_actuator", d * NV),
        ("actuator_force", d * NA), ("act_dot", d * NA), ("qfrc_smooth", d * NV), ("qacc_smooth", d * NV),
        ("ncon", i32), ("nefc", i32), ("contact", Contact * MAXCON),
        ("efc_J", (d * NV) * MAXEFC), ("efc_pos", d * MAXEFC), ("efc_vel", d * MAXEFC), ("efc_aref", d * MAXEFC),
        ("efc_R", d * MAXEFC), ("efc_D", d * MAXEFC), ("efc_force", d * MAXEFC), ("efc_state", i32 * MAXEFC),
        ("qacc", d * NV), ("qfrc_constraint", d * NV), ("solver_cost", d),
        ("solver_iter", i32), ("ls_evals", i32), ("mpr_calls", i32), ("con_overflow", i32),
        ("cfrc_ext", (d * 6) * M.NBODY), ("warning", i32), ("pad_", i32),
    ]


_lib = None


def build(force=False):
    so = os.path.join(HERE, "libtsg_oracle.so")
    src = [os.path.join(HERE, f) for f in ("tsg_oracle.c", "tsg_oracle.h")] + [
        os.path.join(os.path.dirname(HERE), "include", "tsg_model.h")]
    if force or not os.path.isfile(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        assert L.tsgo_sizeof_data() == C.sizeof(Data), (L.tsgo_sizeof_data(), C.sizeof(Data))
        assert L.tsgo_sizeof_model() == C.sizeof(M.TsgModel)
        L.tsgo_primal_cost.restype = d
        L.tsgo_bench.restype = C.c_long
        L.tsgo_bench.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, d, d, C.c_ulonglong, C.c_int, C.POINTER(d)]
        _lib = L
    return _lib


def arr(x):
    return np.ctypeslib.as_array(x)


class MjLike:
    """Single-env oracle instance (model + data)."""

    def __init__(self, xml_file=None, **overrides):
        self.md = dict(M.load_model(xml_file))
        self.md.update(overrides)
        self.model, self._keep = M.model_struct(self.md)
        self.data = Data()
        self.L = lib()
        self.reset_data()
        for name in ("qpos", "qvel", "act", "ctrl", "qacc_warmstart", "ten_length", "ten_velocity", "xpos", "xmat",
                     "geom_xpos", "geom_xmat", "site_xpos", "cfrc_ext", "qacc", "qacc_smooth", "qfrc_smooth",
                     "qfrc_constraint", "ten_J", "efc_J", "efc_force", "efc_D", "efc_R", "efc_aref", "efc_pos",
                     "efc_vel", "efc_state", "actuator_force", "qfrc_passive", "qfrc_bias", "qfrc_actuator", "com_world"):
            setattr(self, name, arr(getattr(self.data, name)))

    # -- mujoco-like entry points
    def reset_data(self):
        self.L.tsgo_reset_data(C.byref(self.model), C.byref(self.data))

    def forward(self):
        self.L.tsgo_forward(C.byref(self.model), C.byref(self.data))

    def step(self, nstep=1):
        self.L.tsgo_step(C.byref(self.model), C.byref(self.data), int(nstep))

    def rne_post_constraint(self):
        self.L.tsgo_rne_post_constraint(C.byref(self.model), C.byref(self.data))

    def contact_force(self, i):
        out = (d * 6)()
        self.L.tsgo_contact_force(C.byref(self.model), C.byref(self.data), int(i), out)
        return np.array(out)

    def primal_cost(self, qacc, want_grad=False):
        q = np.ascontiguousarray(qacc, np.float64)
        g = np.zeros(NV)
        c = self.L.tsgo_primal_cost(C.byref(self.model), C.byref(self.data), q.ctypes.data_as(C.POINTER(d)),
                                    g.ctypes.data_as(C.POINTER(d)) if want_grad else None)
        return (c, g) if want_grad else c

    @property
    def ncon(self):
        return self.data.ncon

    @property
    def nefc(self):
        return self.data.nefc

    def contacts(self):
        return [self.data.contact[i] for i in range(self.data.ncon)]

    # -- state helpers
    def set_state(self, qpos, qvel):
        self.qpos[:] = qpos
        self.qvel[:] = qvel
        self.forward()

    def sphere_pos(self, k):
        """geom s<k> centre (k=0..5): bar k//2, geom 1 (+z end) or 2 (-z end)."""
        return self.geom_xpos[k // 2, 1 + (k % 2)].copy()


def bench(xml_file, n_envs, n_steps, threads=0, frame_skip=20, warm_steps=0, lo=-0.45, hi=-0.15, seed=0):
    """CPU baseline: returns (env_steps, seconds, threads)."""
    import time
    mj = MjLike(xml_file)
    L = lib()
    if threads <= 0:
        threads = L.tsgo_max_threads()
    cs = d(0)
    t0 = time.perf_counter()
    n = L.tsgo_bench(C.addressof(mj.model), n_envs, n_steps, frame_skip, warm_steps, lo, hi, seed, threads, C.byref(cs))
    return n, time.perf_counter() - t0, threads, cs.value


class Batch:
    """persistent batch of oracle envs stepped by a pthread pool (CPU arm of bench.py)."""

    def __init__(self, xml_file, n_envs, seed=0):
        self.mj = MjLike(xml_file)
        self.L = lib()
        self.L.tsgo_batch_create.restype = C.c_void_p
        self.L.tsgo_batch_create.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong]
        self.L.tsgo_batch_step.restype = C.c_long
        self.L.tsgo_batch_step.argtypes = [C.c_void_p, C.c_int, C.c_int, d, d, C.c_int]
        self.L.tsgo_batch_destroy.argtypes = [C.c_void_p]
        self.n_envs = n_envs
        self.b = self.L.tsgo_batch_create(C.addressof(self.mj.model), n_envs, seed)

    def step(self, n_steps=1, frame_skip=20, lo=-0.45, hi=-0.15, threads=0):
        return self.L.tsgo_batch_step(self.b, n_steps, frame_skip, lo, hi, threads)

    def close(self):
        if self.b:
            self.L.tsgo_batch_destroy(self.b)
            self.b = None


def step_states(xml_file, qpos, qvel, act, warm, ctrl, nstep=20, threads=0):
    """every row stepped `nstep` substeps from the given state by the C oracle on all host threads.
    Returns qpos', qvel', ten_length', and per row the (min, max) number of active contacts over the substeps."""
    mj = MjLike(xml_file)
    f = lambda a, w: np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1, w))
    qpos, qvel, act, warm, ctrl = f(qpos, NQ), f(qvel, NV), f(act, NA), f(warm, NV), f(ctrl, NA)
    n = len(qpos)
    oq, ov, ot, mm = np.zeros((n, NQ)), np.zeros((n, NV)), np.zeros((n, NT)), np.zeros((n, 2), np.int32)
    P = lambda a: a.ctypes.data_as(C.POINTER(d))
    lib().tsgo_step_states(C.byref(mj.model), n, int(nstep), P(qpos), P(qvel), P(act), P(warm), P(ctrl), P(oq), P(ov), P(ot),
                           mm.ctypes.data_as(C.POINTER(i32)), int(threads))
    return oq, ov, ot, mm


def mpr(type1, pos1, mat1, size1, type2, pos2, mat2, size2, tol=1e-6, iters=50):
    L = lib()
    f = lambda a: np.ascontiguousarray(a, np.float64)
    p1, m1, s1, p2, m2, s2 = map(f, (pos1, mat1, size1, pos2, mat2, size2))
    depth, dr, ps = d(0), np.zeros(3), np.zeros(3)
    P = lambda a: a.ctypes.data_as(C.POINTER(d))
    hit = L.tsgo_mpr(int(type1), P(p1), P(m1), P(s1), int(type2), P(p2), P(m2), P(s2), d(tol), int(iters),
                     C.byref(depth), P(dr), P(ps))
    return bool(hit), depth.value, dr, ps
