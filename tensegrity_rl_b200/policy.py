"""Batched inference of the reference's SB3 2.2.1 SAC actors on the simulator's device.

Semantics of `SAC.predict`: obs -> float32 -> Linear-ReLU-Linear-ReLU (`latent_pi`) ->
`mu`, `log_std` clamped to [-20, 2] -> a = tanh(mu + exp(log_std) * eps)  (deterministic: tanh(mu)) ->
unscale with the CHECKPOINT's action bounds: low + 0.5 (a + 1)(high - low).  run.py:137 never passes
`deterministic`, so the reference evaluates stochastically.

Weights come from assets/policies/*.npz (actor tensors extracted by tools/extract_assets.py) or straight
from an SB3 zip (`policy.pth` + `data`), without importing stable_baselines3.
"""
from __future__ import annotations

import io
import json
import os
import re
import zipfile

import numpy as np

from .model import ASSET_DIR


def _parse_repr(s, n):
    s = s.strip()
    vals = [float(x) for x in re.findall(r"[-+0-9.eE]+", s)] if s.startswith("[") else [float(s)] * n
    return np.array(vals, np.float32)


def load_actor_arrays(path_or_name):
    """-> dict(W0,b0,W1,b1,Wmu,bmu,Wls,bls, action_low, action_high, obs_dim)"""
    p = path_or_name
    if not os.path.isfile(p):
        p = os.path.join(ASSET_DIR, "policies", path_or_name + ".npz")
    if p.endswith(".npz"):
        z = np.load(p)
        get = lambda k: z["actor__" + k.replace(".", "__")]
        low, high, obs_dim = z["action_low"], z["action_high"], int(z["obs_dim"])
    else:
        import torch

        zf = zipfile.ZipFile(p)
        sd = torch.load(io.BytesIO(zf.read("policy.pth")), map_location="cpu", weights_only=True)
        data = json.loads(zf.read("data"))
        get = lambda k: sd["actor." + k].numpy()
        n = int(data["action_space"]["_shape"][0])
        low, high = _parse_repr(data["action_space"]["low_repr"], n), _parse_repr(data["action_space"]["high_repr"], n)
        obs_dim = int(data["observation_space"]["_shape"][0])
    return dict(W0=get("latent_pi.0.weight"), b0=get("latent_pi.0.bias"), W1=get("latent_pi.2.weight"),
                b1=get("latent_pi.2.bias"), Wmu=get("mu.weight"), bmu=get("mu.bias"), Wls=get("log_std.weight"),
                bls=get("log_std.bias"), action_low=np.asarray(low, np.float32), action_high=np.asarray(high, np.float32),
                obs_dim=obs_dim)


class SacActor:
    """torch module-free batched actor (weights as device tensors)."""

    def __init__(self, path_or_name, device="cuda", seed=0):
        import torch

        self.torch = torch
        a = load_actor_arrays(path_or_name)
        self.obs_dim = a["obs_dim"]
        dev = torch.device(device)
        f = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32, device=dev)
        self.W0, self.b0, self.W1, self.b1 = f(a["W0"]).t().contiguous(), f(a["b0"]), f(a["W1"]).t().contiguous(), f(a["b1"])
        # fuse the two heads into one GEMM
        self.Wh = torch.cat([f(a["Wmu"]), f(a["Wls"])], 0).t().contiguous()
        self.bh = torch.cat([f(a["bmu"]), f(a["bls"])], 0)
        self.low, self.high = f(a["action_low"]), f(a["action_high"])
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        self.device = dev

    def __call__(self, obs, deterministic=False):
        """obs [N, obs_dim] (any float dtype, on device) -> action [N, 6] float32 in the checkpoint's bounds."""
        t = self.torch
        x = obs.to(t.float32)
        x = t.relu(t.addmm(self.b0, x, self.W0))
        x = t.relu(t.addmm(self.b1, x, self.W1))
        hcat = t.addmm(self.bh, x, self.Wh)
        mu, log_std = hcat[:, :6], hcat[:, 6:].clamp(-20.0, 2.0)
        if deterministic:
            a = t.tanh(mu)
        else:
            eps = t.randn(mu.shape, generator=self.gen, device=self.device, dtype=t.float32)
            a = t.tanh(mu + log_std.exp() * eps)
        return self.low + 0.5 * (a + 1.0) * (self.high - self.low)

    def predict(self, obs, deterministic=False):
        """numpy in / numpy out, SB3-style signature: returns (action, None)."""
        t = self.torch
        o = t.as_tensor(np.asarray(obs, np.float32), device=self.device)
        single = o.ndim == 1
        act = self(o.reshape(-1, self.obs_dim), deterministic).cpu().numpy()
        return (act[0] if single else act), None
