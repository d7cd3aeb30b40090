"""Dump the inputs / outputs of env steps whose CUDA result differs from the oracle's by more than a threshold, for
offline replay (oracle, emulator).  usage: python tools/find_outliers.py xml steps threshold out.npz [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensegrity_rl_b200 import TensegrityVecEnv
from oracle import oracle as O
xml, steps, thr, out = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
prec = sys.argv[5] if len(sys.argv) > 5 else "f64"
n = 4096
v = TensegrityVecEnv(n, xml_file=xml, env="tr_env", auto_reset=False, terminate_when_unhealthy=False, max_episode_steps=0, precision=prec)
v.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(1)
cases = []
for step in range(steps):
    a = -0.45 + 0.3 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
    before = v.get_state()
    v.step_tensor(a)
    after, info = v.get_state(), v.info.cpu().numpy()
    oq, ov, ot, mm = O.step_states(xml, before["qpos"], before["qvel"], before["act"], before["qacc_warmstart"], after["ctrl"])
    scale = lambda x: np.maximum(1.0, np.abs(x).max(axis=1))
    err = np.maximum(np.abs(after["qpos"] - oq).max(1) / scale(oq), np.abs(after["qvel"] - ov).max(1) / scale(ov))
    for e in np.nonzero(err > thr)[0]:
        cases.append(np.concatenate([[step, e, err[e]], before["qpos"][e], before["qvel"][e], before["act"][e], before["qacc_warmstart"][e],
                                     after["ctrl"][e], after["qpos"][e], after["qvel"][e], info[e], mm[e]]))
print("cases", len(cases), "largest", max([c[2] for c in cases]) if cases else 0)
np.savez(out, cases=np.array(cases))
