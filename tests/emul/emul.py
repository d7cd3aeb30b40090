"""ctypes driver of the host 'one warp' emulator of the CUDA source (TEST HARNESS ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

from tensegrity_rl_b200 import model as M

HERE = os.path.dirname(os.path.abspath(__file__))
STATE_STRIDE, INFO_DIM, HEADING_SLOTS, NDRAW = 96, 32, 32, 10
_libs = {}


def lib(reverse=False):
    """reverse=True: the variant that runs the items of every phase in reverse order (hazard detector)."""
    if reverse not in _libs:
        so = os.path.join(HERE, "libtsg_emul_rev.so" if reverse else "libtsg_emul.so")
        csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tensegrity_rl_b200", "csrc")
        deps = [os.path.join(HERE, "tsg_emul.cpp")] + [os.path.join(csrc, f) for f in ("tsg_core.cuh", "tsg_env.cuh", "tsg_host.h")]
        if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off",
                                   "-Wno-unknown-pragmas"] + (["-DTSG_EMUL_REVERSE"] if reverse else []) +
                                  ["-o", so, os.path.join(HERE, "tsg_emul.cpp")])
        L = C.CDLL(so)
        L.emul_create.restype = C.c_char_p
        _libs[reverse] = L
    return _libs[reverse]


def P(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Emul:
    def __init__(self, xml_file="flat", reverse=False, **env_kwargs):
        self.md = M.load_model(xml_file)
        self.model, self._keep = M.model_struct(self.md)
        self.cfg = M.env_config(self.md, **env_kwargs)
        self.L = lib(reverse)
        h = C.c_void_p()
        err = self.L.emul_create(C.byref(self.model), C.byref(self.cfg), C.byref(h))
        if err:
            raise RuntimeError(err.decode())
        self.h = h
        self.rec = np.zeros(STATE_STRIDE)
        self.heading = np.zeros(HEADING_SLOTS)
        self.obs = np.zeros(self.cfg.obs_dim)
        self.info = np.zeros(INFO_DIM)
        self.draws = np.zeros(NDRAW)
        self.rec[0:21] = self.md["qpos0"]

    def __del__(self):
        try:
            self.L.emul_destroy(self.h)
        except Exception:
            pass

    # state record views
    @property
    def qpos(self): return self.rec[0:21]
    @property
    def qvel(self): return self.rec[21:39]
    @property
    def warm(self): return self.rec[39:57]
    @property
    def ctrl(self): return self.rec[57:63]
    @property
    def act(self): return self.rec[63:69]

    def mj_step(self, ctrl, nstep=1):
        ten = np.zeros(9); cfrc = np.zeros(24); stats = np.zeros(6, np.int32)
        c = np.ascontiguousarray(ctrl, np.float64)
        self.L.emul_mj_step(self.h, P(self.rec), P(c), int(nstep), P(ten), P(cfrc), P(stats, C.c_int))
        return ten, cfrc.reshape(4, 6), stats

    def step(self, action):
        a = np.ascontiguousarray(action, np.float64)
        rew = np.zeros(1); done = np.zeros(1, np.uint8)
        self.L.emul_step(self.h, P(self.rec), P(self.heading), P(a), P(self.obs), P(rew), P(done, C.c_uint8), P(self.info))
        return self.obs.copy(), float(rew[0]), bool(done[0]), self.info.copy()

    def reset(self, draws=None, seed=0, env_id=0):
        if draws is not None:
            self.draws[:] = draws
        self.L.emul_reset(self.h, P(self.rec), P(self.heading), P(self.draws), int(draws is not None),
                          C.c_ulonglong(seed), C.c_longlong(env_id), P(self.obs))
        return self.obs.copy()

    def set_noise(self, seed, env_id=0):
        """observation-noise keys; the true observation of every step / reset then lands in self.real_obs"""
        self.real_obs = np.zeros(self.cfg.obs_dim)
        self.L.emul_set_noise(self.h, C.c_ulonglong(seed), C.c_longlong(env_id), P(self.real_obs))

    def forward(self):
        self.L.emul_forward(self.h, P(self.rec), P(self.heading), P(self.obs), P(self.info))
        return self.obs.copy(), self.info.copy()


def obs_normals(seed, stream, nreset, step, n, reverse=False):
    """the standard-normal draws the CUDA source uses for the observation of (stream, reset count, episode step)"""
    out = np.zeros(n + 1)
    lib(reverse).emul_obs_normals(C.c_ulonglong(seed), C.c_ulonglong(stream), C.c_ulonglong(nreset), C.c_ulonglong(step),
                                  int(n), P(out))
    return out[:n]
