"""A/B timing of a libtsg build: env-steps/s at N envs (flat, random ctrl) + an oracle spot check."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensegrity_rl_b200 import TensegrityVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
xml = sys.argv[3] if len(sys.argv) > 3 else "flat"
check = int(sys.argv[4]) if len(sys.argv) > 4 else 1
env = TensegrityVecEnv(n, xml_file=xml, env="tr_env", precision=os.environ.get("TSG_PRECISION", "f64"), auto_reset=bool(int(os.environ.get("TSG_AUTORESET", "1"))), reset_pool=os.environ.get("TSG_POOL", "auto") if os.environ.get("TSG_POOL", "auto") == "auto" else int(os.environ["TSG_POOL"]))
env.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(0)
ctrl = -0.45 + 0.3 * torch.rand(steps + 3, n, 6, generator=g, device="cuda", dtype=torch.float64)
for k in range(2):
    env.step_tensor(ctrl[k], want_info=False)
for k in range(int(os.environ.get("TSG_SETTLE", "0"))):   # untimed random-ctrl steps: the steady-state contact load
    env.step_tensor(-0.45 + 0.3 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64), want_info=False)
worst = -1
if check:
    from oracle import oracle as O
    mj = O.MjLike(xml)
    before = env.get_state()
    env.step_tensor(ctrl[2])
    after = env.get_state()
    worst = 0
    for e in range(0, n, max(1, n // 24)):
        mj.reset_data(); mj.qpos[:] = before["qpos"][e]; mj.qvel[:] = before["qvel"][e]; mj.act[:] = before["act"][e]
        mj.qacc_warmstart[:] = before["qacc_warmstart"][e]; mj.ctrl[:] = after["ctrl"][e]
        mj.step(20)
        worst = max(worst, np.abs(after["qpos"][e] - mj.qpos).max(), np.abs(after["qvel"][e] - mj.qvel).max() / max(1, np.abs(mj.qvel).max()))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for k in range(steps):
    env.step_tensor(ctrl[3 + k], want_info=(k == steps - 1))
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
info = env.info
print("%s N=%d: %.0f env-steps/s (%.1f ms/step) | oracle worst %.2e | ncon %.2f niter/sub %.2f nls/sub %.2f overflow %d bad %d | %s"
      % (os.environ.get("TSG_LIB", "default"), n, n * steps / ms * 1e3, ms / steps, worst, float(info[:, 19].mean()),
         float(info[:, 20].mean()) / 20, float(info[:, 21].mean()) / 20, int(info[:, 28].sum()), int(info[:, 29].sum()), env.kernel_config()))
