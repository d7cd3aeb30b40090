"""Small driver for ncu: N envs, reset, a few env steps with random ctrl (flat XML by default)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensegrity_rl_b200 import TensegrityVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
xml = sys.argv[3] if len(sys.argv) > 3 else "flat"
env = TensegrityVecEnv(n, xml_file=xml, env="tr_env", auto_reset=True)
env.reset_tensor()
g = torch.Generator(device="cuda"); g.manual_seed(0)
for k in range(steps):
    a = -0.45 + 0.3 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
    env.step_tensor(a)
torch.cuda.synchronize()
info = env.info
print("ncon %.2f niter/sub %.2f nls/sub %.2f" % (float(info[:, 19].mean()), float(info[:, 20].mean()) / 20, float(info[:, 21].mean()) / 20))
