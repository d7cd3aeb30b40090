#!/usr/bin/env python
"""Materialise the DATA the hot path needs from the read-only reference tree into assets/.

Run in the build container (where /root/reference exists); the GPU box only sees assets/.
  * model constants of the two MJCF files (parsed by tensegrity_rl_b200.model.parse_mjcf)
      -> assets/model_flat.json, assets/model_uneven.json (+ _hfield.npy)
  * the reset pose table `rolling_qpos` of tr_env.reset_model (tr_env.py:723-728), read with `ast`
      -> assets/reset_poses.json
  * SB3 SAC checkpoints: actor tensors, saved action bounds, obs dim, `_last_obs`, ep_info stats
      -> assets/policies/<name>.npz     (fp32 actor only; critics/optimisers are out of scope)
  * golden fixtures for tests: `_last_obs` vectors -> tests/golden/last_obs.json
"""
import ast
import base64
import io
import json
import os
import pickle
import re
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("TSG_REFERENCE", "/root/reference")
ASSETS = os.path.join(ROOT, "assets")

from tensegrity_rl_b200 import model as M  # noqa: E402


def extract_models():
    for xml, out in (("3prism_jonathan_steady_side.xml", "model_flat.json"),
                     ("3prism_jonathan_steady_side_uneven_ground.xml", "model_uneven.json")):
        md = M.parse_mjcf(os.path.join(REF, xml))
        md["source"] = xml
        M.save_model_json(md, os.path.join(ASSETS, out))
        print("wrote", out)


def extract_reset_poses():
    src = open(os.path.join(REF, "tr_env/tr_env/envs/tr_env.py")).read()
    tree = ast.parse(src)
    table = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "rolling_qpos" for t in node.targets):
            table = ast.literal_eval(node.value)
    arr = np.array(table, float)
    assert arr.shape == (6, 21)
    json.dump({"source": "tr_env.py reset_model rolling_qpos", "rolling_qpos": arr.tolist()},
              open(os.path.join(ASSETS, "reset_poses.json"), "w"), indent=1)
    print("wrote reset_poses.json", arr.shape)


def _unpickle(blob):
    return pickle.loads(base64.b64decode(blob[":serialized:"]))


def _parse_repr(s, n):
    s = s.strip()
    if s.startswith("["):
        vals = [float(x) for x in re.findall(r"[-+0-9.eE]+", s)]
    else:
        vals = [float(s)] * n
    assert len(vals) == n, (s, n)
    return np.array(vals, np.float32)


CHECKPOINTS = {
    # BASELINE.json configs 3-5 + the remaining pretrained/legacy policies named in the README
    "forward": "best_models_pretrained/forward/SAC_5500000.zip",
    "backward": "best_models_pretrained/backward/SAC_4700000.zip",
    "yaw_CCW": "best_models_pretrained/yaw_CCW/SAC_5000000.zip",
    "yaw_CW": "best_models_pretrained/yaw_CW/SAC_4000000.zip",
    "traj_track": "models_traj/SAC_16525000_track.zip",
    "traj_ccw": "models_traj/SAC_2175000_ccw.zip",
    "traj_cw": "models_traj/SAC_1250000_cw.zip",
}


def extract_policies():
    import torch

    os.makedirs(os.path.join(ASSETS, "policies"), exist_ok=True)
    golden = {}
    for name, rel in CHECKPOINTS.items():
        path = os.path.join(REF, rel)
        if not os.path.isfile(path):
            cands = sorted(f for f in os.listdir(os.path.dirname(path)) if f.endswith(".zip"))
            print("missing", rel, "available:", cands)
            continue
        z = zipfile.ZipFile(path)
        data = json.loads(z.read("data"))
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=True)
        act = data["action_space"]
        n_act = int(act["_shape"][0])
        low = _parse_repr(act["low_repr"], n_act)
        high = _parse_repr(act["high_repr"], n_act)
        obs_dim = int(data["observation_space"]["_shape"][0])
        out = {k.replace(".", "__"): v.numpy().astype(np.float32) for k, v in sd.items() if k.startswith("actor.")}
        last_obs = np.asarray(_unpickle(data["_last_obs"]), np.float64).reshape(-1)
        eps = _unpickle(data["ep_info_buffer"])
        ep = np.array([[e["r"], e["l"], e["t"]] for e in eps], np.float64)
        np.savez(os.path.join(ASSETS, "policies", name + ".npz"), action_low=low, action_high=high,
                 obs_dim=np.int32(obs_dim), last_obs=last_obs, ep_info=ep, source=np.bytes_(rel), **out)
        golden[name] = {"source": rel, "obs_dim": obs_dim, "last_obs": last_obs.tolist(),
                        "action_low": low.tolist(), "action_high": high.tolist(),
                        "ep_return_mean": float(ep[:, 0].mean()), "ep_len_mean": float(ep[:, 1].mean())}
        print("wrote policies/%s.npz obs=%d act=[%g,%g] keys=%d" % (name, obs_dim, low[0], high[0], len(out)))
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    json.dump(golden, open(os.path.join(ROOT, "tests", "golden", "last_obs.json"), "w"), indent=1)


if __name__ == "__main__":
    os.makedirs(ASSETS, exist_ok=True)
    extract_models()
    extract_reset_poses()
    extract_policies()
