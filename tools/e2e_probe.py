"""Where does the host-API step spend its time?  (pinned H2D, step, D2H, sync) -- one-off measurement aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensegrity_rl_b200 import TensegrityVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
env = TensegrityVecEnv(n, xml_file="flat", env="tr_env", auto_reset=True, reset_pool="auto")
env.reset_tensor(); torch.cuda.synchronize()
dev = env.device
hc = torch.empty(n, 6, dtype=torch.float64).pin_memory(); ho = torch.empty(n, env.obs_dim, dtype=torch.float64).pin_memory()
hr = torch.empty(n, dtype=torch.float64).pin_memory(); hd = torch.empty(n, dtype=torch.uint8).pin_memory()
dc = torch.empty(n, 6, dtype=torch.float64, device=dev)
src = -0.45 + 0.3 * torch.rand(8, n, 6, dtype=torch.float64)
T = {k: 0.0 for k in ("hostcopy", "h2d", "step_call", "d2h_call", "sync")}
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
for k in range(8):
    t0 = time.perf_counter(); hc.copy_(src[k]); t1 = time.perf_counter()
    dc.copy_(hc, non_blocking=True); t2 = time.perf_counter()
    ev[k][0].record(); obs, rew, done = env.step_tensor(dc, want_info=False); ev[k][1].record(); t3 = time.perf_counter()
    ho.copy_(obs, non_blocking=True); hr.copy_(rew, non_blocking=True); hd.copy_(done, non_blocking=True); t4 = time.perf_counter()
    torch.cuda.synchronize(); t5 = time.perf_counter()
    if k >= 2:
        for name, d in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)): T[name] += d / 6
print({k: round(v * 1e3, 2) for k, v in T.items()}, "ms; device step ms:", [round(a.elapsed_time(b), 1) for a, b in ev[2:]])
