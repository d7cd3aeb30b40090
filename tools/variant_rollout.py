"""Forward / yaw policies on a DERIVED model variant: bars, sites, tendons and filter actuators of the uneven-ground XML
on a flat plane at z = 0 (hypothesis for the model the legacy checkpoints were trained on; see tests/test_golden_last_obs.py)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensegrity_rl_b200 import TensegrityVecEnv, SacActor, load_model
from tensegrity_rl_b200.rollout import rollout

md = dict(load_model("uneven"))
md["floor_type"] = 0; md["floor_pos"] = [0.0, 0.0, 0.0]; md["hfield"] = None
n, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
for pol, kw in (("forward", dict(desired_action="straight", desired_direction=1)),
                ("backward", dict(desired_action="straight", desired_direction=-1)),
                ("yaw_CCW", dict(desired_action="turn", desired_direction=1, terminate_when_unhealthy=False)),
                ("yaw_CW", dict(desired_action="turn", desired_direction=-1, terminate_when_unhealthy=False))):
    v = TensegrityVecEnv(n, xml_file=md, env="tensegrity_env", auto_reset=True, reset_pool="auto", **kw)
    v.reset_tensor()
    s = rollout(v, SacActor(pol), steps)
    dt_total = s["length_sum"] * v.dt
    print(json.dumps({"policy": pol, "model": "uneven-XML bars+actuators on a flat plane (derived)", "envs": n, "steps": steps,
                      "episodes": s["episodes"], "length_mean": s["length_mean"], "return_mean": s["return_mean"],
                      "return_per_step": s["return_sum"] / max(s["length_sum"], 1),
                      "forward_speed_m_per_s": s["disp_sum"] / dt_total, "yaw_rate_rad_per_s": s["yaw_sum"] / dt_total,
                      "disp_mean": s["disp_mean"], "disp_std": s["disp_std"], "yaw_mean": s["yaw_mean"], "yaw_std": s["yaw_std"]}), flush=True)
    v.close()
