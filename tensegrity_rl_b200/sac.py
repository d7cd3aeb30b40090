"""Device-resident SAC training around the batched simulator.

The reference trains with stable-baselines3 2.2.1 `SAC('MlpPolicy', env, learning_rate=lr, ...)` (run.py:36-98):
one CPU env, a 1e6-transition numpy replay buffer, one gradient step per env step.  Here the same algorithm runs
where the simulator runs: transitions of all N envs go straight from `TensegrityVecEnv.step_tensor` into a replay
buffer in HBM, and the SAC update (SB3 `SAC.train`, restated below) is captured ONCE in a CUDA graph, so a gradient
step is a single graph launch instead of ~150 small kernels.  Nothing crosses PCIe inside the loop.

What is restated from SB3 2.2.1 (the package is not importable here; checked against the checkpoint
layout of the reference's zips -- `policy.pth`, `pytorch_variables.pth`, `*.optimizer.pth`, `data`):
  * networks: actor `latent_pi` [obs-256-256, ReLU] + `mu` + `log_std` (clamped to [-20, 2]); two critics
    `qf0`, `qf1` [(obs+act)-256-256-1]; `critic_target` a Polyak copy (tau 0.005) -- parameter names identical to
    SB3's state_dict, so the reference's checkpoints load and ours load back into SB3;
  * action distribution: a = tanh(mu + sigma * eps); log pi = sum N(.).log_prob - sum log(1 - a^2 + 1e-6);
  * losses: entropy coefficient `-(log_alpha * (log_pi + target_entropy).detach()).mean()`; critic
    `0.5 * sum_i mse(q_i, r + (1 - done) * gamma * (min_j q'_j(s', a') - alpha * log pi(a'|s')))`; actor
    `(alpha * log_pi - min_j q_j(s, a_pi)).mean()`; Adam(lr 3e-4) x 3; target update every
    `target_update_interval` gradient steps;
  * replay: uniform sampling, `done` stored as terminated-and-not-time-limit (handle_timeout_termination),
    next_obs of a finished episode = its terminal observation, actions stored scaled to [-1, 1];
  * collection: uniform random actions until `learning_starts`, then stochastic actor actions.
torch is used for autograd and the library GEMMs of the 256-wide MLPs (plumbing, per the task framing); the hot
path of this repo stays the simulator kernel.
"""
from __future__ import annotations

import io
import json
import math
import zipfile

import numpy as np

LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0
TANH_EPS = 1e-6


def _nn():
    import torch
    return torch, torch.nn


def build_policy(obs_dim, act_dim, hidden=256):
    """SB3 `SACPolicy` with its parameter names (actor.latent_pi.0.weight, critic.qf0.4.bias, ...)."""
    torch, nn = _nn()

    class Actor(nn.Module):
        def __init__(self):
            super().__init__()
            self.latent_pi = nn.Sequential(nn.Linear(obs_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU())
            self.mu = nn.Linear(hidden, act_dim)
            self.log_std = nn.Linear(hidden, act_dim)

        def dist(self, obs):
            z = self.latent_pi(obs)
            return self.mu(z), self.log_std(z).clamp(LOG_STD_MIN, LOG_STD_MAX)

        def action_log_prob(self, obs, eps=None):
            mu, log_std = self.dist(obs)
            std = log_std.exp()
            if eps is None:
                eps = torch.randn_like(mu)
            g = mu + std * eps
            a = torch.tanh(g)
            # Normal(mu, std).log_prob(g) with (g - mu) / std == eps
            logp = (-0.5 * eps * eps - log_std - 0.5 * math.log(2 * math.pi)).sum(1)
            logp = logp - torch.log(1.0 - a * a + TANH_EPS).sum(1)
            return a, logp

        def forward(self, obs, deterministic=False):
            mu, log_std = self.dist(obs)
            if deterministic:
                return torch.tanh(mu)
            return torch.tanh(mu + log_std.exp() * torch.randn_like(mu))

    class Critic(nn.Module):
        def __init__(self):
            super().__init__()
            mk = lambda: nn.Sequential(nn.Linear(obs_dim + act_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden),
                                       nn.ReLU(), nn.Linear(hidden, 1))
            self.qf0, self.qf1 = mk(), mk()

        def forward(self, obs, act):
            x = torch.cat([obs, act], 1)
            return self.qf0(x), self.qf1(x)

    class SacPolicy(nn.Module):
        def __init__(self):
            super().__init__()
            self.actor, self.critic, self.critic_target = Actor(), Critic(), Critic()
            self.critic_target.load_state_dict(self.critic.state_dict())
            for p in self.critic_target.parameters():
                p.requires_grad_(False)

    return SacPolicy()


class ReplayBuffer:
    """Ring buffer of transitions in device memory; `add` takes the N transitions of one vec-env step."""

    def __init__(self, capacity, obs_dim, act_dim, device):
        torch, _ = _nn()
        self.capacity, self.device = int(capacity), torch.device(device)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.obs, self.next_obs = z(self.capacity, obs_dim), z(self.capacity, obs_dim)
        self.act, self.rew, self.done = z(self.capacity, act_dim), z(self.capacity), z(self.capacity)
        self.pos, self.full = 0, False
        self.size_t = torch.zeros((), dtype=torch.int64, device=self.device)   # device copy of `size` (graph-safe sampling)

    @property
    def size(self):
        return self.capacity if self.full else self.pos

    def add(self, obs, next_obs, act, rew, done):
        """all [n, ...] device tensors; done = 1 only for true terminations (time-limit truncations store 0)."""
        n = obs.shape[0]
        if n > self.capacity:
            raise ValueError("replay buffer smaller than one vec-env step")
        first = min(n, self.capacity - self.pos)
        for dst, src in ((self.obs, obs), (self.next_obs, next_obs), (self.act, act), (self.rew, rew), (self.done, done)):
            dst[self.pos:self.pos + first] = src[:first]
            if first < n:
                dst[:n - first] = src[first:]
        self.pos += n
        if self.pos >= self.capacity:
            self.full, self.pos = True, self.pos - self.capacity
        self.size_t.fill_(self.size)

    def sample_indices(self, batch_size):
        torch, _ = _nn()
        u = torch.rand(batch_size, device=self.device)
        return (u * self.size_t).long().clamp_(max=self.capacity - 1)

    def gather(self, idx):
        return self.obs[idx], self.act[idx], self.next_obs[idx], self.done[idx], self.rew[idx]


class SACLearner:
    """SB3-equivalent SAC on device tensors.  `env` follows the TensegrityVecEnv tensor protocol
    (num_envs, obs_dim, action_space, reset_tensor(), step_tensor(ctrl) -> (obs, reward, done), .info, .term_obs)."""

    def __init__(self, obs_dim, act_dim=6, action_low=-0.45, action_high=0.15, device="cuda", learning_rate=3e-4,
                 gamma=0.99, tau=0.005, batch_size=256, buffer_size=1_000_000, learning_starts=100, train_freq=1,
                 gradient_steps=1, target_update_interval=1, ent_coef="auto", target_entropy="auto", seed=0,
                 use_cuda_graph=None):
        torch, nn = _nn()
        self.torch = torch
        self.device = torch.device(device)
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        torch.manual_seed(seed)
        self.policy = build_policy(self.obs_dim, self.act_dim).to(self.device)
        f = lambda v: torch.as_tensor(np.broadcast_to(np.asarray(v, np.float32), (self.act_dim,)).copy(), device=self.device)
        self.low, self.high = f(action_low), f(action_high)
        self.gamma, self.tau, self.batch_size = float(gamma), float(tau), int(batch_size)
        self.learning_starts, self.train_freq, self.gradient_steps = int(learning_starts), int(train_freq), int(gradient_steps)
        self.target_update_interval = int(target_update_interval)
        self.target_entropy = -float(self.act_dim) if target_entropy == "auto" else float(target_entropy)
        self.auto_ent = isinstance(ent_coef, str) and ent_coef.startswith("auto")
        init = 1.0
        if self.auto_ent and "_" in ent_coef:
            init = float(ent_coef.split("_")[1])
        if not self.auto_ent:
            init = float(ent_coef)
        self.log_ent_coef = torch.full((1,), math.log(init), device=self.device, requires_grad=self.auto_ent)
        cap = self.device.type == "cuda"
        self.use_cuda_graph = cap if use_cuda_graph is None else (bool(use_cuda_graph) and cap)
        kw = dict(lr=learning_rate, capturable=True) if cap else dict(lr=learning_rate)
        self.actor_opt = torch.optim.Adam(self.policy.actor.parameters(), **kw)
        self.critic_opt = torch.optim.Adam(self.policy.critic.parameters(), **kw)
        self.ent_opt = torch.optim.Adam([self.log_ent_coef], **kw) if self.auto_ent else None
        self.buffer = ReplayBuffer(buffer_size, self.obs_dim, self.act_dim, self.device)
        self.num_timesteps, self.n_updates = 0, 0
        self._graph = None
        self._static = None
        self.last_losses = torch.zeros(4, device=self.device)   # critic, actor, ent_coef loss, ent_coef
        self._source_data = None

    # ------------------------------------------------------------------ actions
    def scale_action(self, a):
        return 2.0 * (a - self.low) / (self.high - self.low) - 1.0

    def unscale_action(self, a):
        return self.low + 0.5 * (a + 1.0) * (self.high - self.low)

    def act(self, obs32, deterministic=False, random=False):
        """-> (scaled action in [-1, 1] for the buffer, unscaled ctrl for the env)"""
        torch = self.torch
        with torch.no_grad():
            if random:   # SB3 samples action_space.sample() and scales it: uniform in [-1, 1]
                a = 2.0 * torch.rand(obs32.shape[0], self.act_dim, device=self.device) - 1.0
            else:
                a = self.policy.actor(obs32, deterministic)
        return a, self.unscale_action(a)

    # ------------------------------------------------------------------ one gradient step (SB3 SAC.train body)
    def losses(self, obs, act, next_obs, done, rew, eps_pi=None, eps_next=None):
        """(ent_coef_loss, critic_loss, actor_loss, ent_coef) for a batch; pure function of the current parameters
        (used by update_batch and, with fixed eps, by the formula tests)."""
        torch, P = self.torch, self.policy
        a_pi, logp = P.actor.action_log_prob(obs, eps_pi)
        ent_coef = self.log_ent_coef.detach().exp()
        ent_loss = -(self.log_ent_coef * (logp + self.target_entropy).detach()).mean() if self.auto_ent else logp.new_zeros(())
        with torch.no_grad():
            a_next, logp_next = P.actor.action_log_prob(next_obs, eps_next)
            q0, q1 = P.critic_target(next_obs, a_next)
            next_q = torch.minimum(q0, q1).squeeze(1) - ent_coef * logp_next
            target_q = rew + (1.0 - done) * self.gamma * next_q
        c0, c1 = P.critic(obs, act)
        critic_loss = 0.5 * (((c0.squeeze(1) - target_q) ** 2).mean() + ((c1.squeeze(1) - target_q) ** 2).mean())
        p0, p1 = P.critic(obs, a_pi)
        actor_loss = (ent_coef * logp - torch.minimum(p0, p1).squeeze(1)).mean()
        return ent_loss, critic_loss, actor_loss, ent_coef

    def _update_from_buffer(self, do_target_update=True):
        torch, P = self.torch, self.policy
        idx = self.buffer.sample_indices(self.batch_size)
        obs, act, next_obs, done, rew = self.buffer.gather(idx)
        # SB3 order: entropy coefficient, critic, actor; each optimizer zeroes its own grads first
        a_pi, logp = P.actor.action_log_prob(obs)
        ent_coef = self.log_ent_coef.detach().exp()
        if self.auto_ent:
            ent_loss = -(self.log_ent_coef * (logp + self.target_entropy).detach()).mean()
            self.ent_opt.zero_grad(set_to_none=True)
            ent_loss.backward()
            self.ent_opt.step()
        else:
            ent_loss = logp.new_zeros(())
        with torch.no_grad():
            a_next, logp_next = P.actor.action_log_prob(next_obs)
            q0, q1 = P.critic_target(next_obs, a_next)
            target_q = rew + (1.0 - done) * self.gamma * (torch.minimum(q0, q1).squeeze(1) - ent_coef * logp_next)
        c0, c1 = P.critic(obs, act)
        critic_loss = 0.5 * (((c0.squeeze(1) - target_q) ** 2).mean() + ((c1.squeeze(1) - target_q) ** 2).mean())
        self.critic_opt.zero_grad(set_to_none=True)
        critic_loss.backward()
        self.critic_opt.step()
        p0, p1 = P.critic(obs, a_pi)
        actor_loss = (ent_coef * logp - torch.minimum(p0, p1).squeeze(1)).mean()
        self.actor_opt.zero_grad(set_to_none=True)
        actor_loss.backward()
        self.actor_opt.step()
        if do_target_update:
            with torch.no_grad():
                tp, sp = list(P.critic_target.parameters()), list(P.critic.parameters())
                torch._foreach_mul_(tp, 1.0 - self.tau)
                torch._foreach_add_(tp, sp, alpha=self.tau)
        self.last_losses.copy_(torch.stack([critic_loss.detach(), actor_loss.detach(), ent_loss.detach(), ent_coef.squeeze()]))

    GRAPH_WARMUP = 3

    def _capture(self):
        """whole-update capture (PyTorch's recipe: warm up on a side stream, then capture with capturable Adam).
        The warm-up passes are ordinary gradient steps and are counted as such."""
        torch = self.torch
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(self.GRAPH_WARMUP):
                self._update_from_buffer()
        torch.cuda.current_stream(self.device).wait_stream(s)
        for opt in (self.actor_opt, self.critic_opt, self.ent_opt):
            if opt is not None:
                opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._update_from_buffer()          # recorded, not executed
        self._graph = g
        self.n_updates += self.GRAPH_WARMUP
        return self.GRAPH_WARMUP

    def update(self, gradient_steps=1):
        if self.buffer.size == 0:
            raise RuntimeError("update() on an empty replay buffer")
        graph_ok = self.use_cuda_graph and self.target_update_interval == 1
        if graph_ok and self._graph is None:
            gradient_steps -= self._capture()
        for _ in range(max(0, gradient_steps)):
            if graph_ok:
                self._graph.replay()
            else:
                self._update_from_buffer(do_target_update=(self.n_updates % self.target_update_interval == 0))
            self.n_updates += 1

    # ------------------------------------------------------------------ collection + training loop (SB3 learn)
    def collect_step(self, env, obs32):
        """one vec-env step into the replay buffer; returns the next observation batch (float32)."""
        torch = self.torch
        from . import lib as _lib
        a, ctrl = self.act(obs32, random=self.num_timesteps < self.learning_starts)
        obs2, rew, done = env.step_tensor(ctrl)
        d = done.bool()
        term = env.info[:, _lib.INFO["terminated"]] > 0
        next_obs = torch.where(d.unsqueeze(1), env.term_obs, obs2).to(torch.float32) if env.auto_reset else obs2.to(torch.float32)
        self.buffer.add(obs32, next_obs, a, rew.to(torch.float32), (d & term).to(torch.float32))
        self.num_timesteps += env.num_envs
        return obs2.to(torch.float32)

    def learn(self, env, total_timesteps, log_interval=0, callback=None):
        """total_timesteps counts env transitions (SB3: num_timesteps += n_envs per vec step)."""
        obs32 = env.reset_tensor().to(self.torch.float32)
        vec_steps = 0
        target = self.num_timesteps + int(total_timesteps)
        while self.num_timesteps < target:
            obs32 = self.collect_step(env, obs32)
            vec_steps += 1
            if vec_steps % self.train_freq == 0 and self.num_timesteps > self.learning_starts and self.buffer.size >= 1:
                gs = self.gradient_steps if self.gradient_steps > 0 else self.train_freq * env.num_envs
                self.update(gs)
            if callback is not None and (log_interval and vec_steps % log_interval == 0):
                callback(self, vec_steps)
        return self

    # ------------------------------------------------------------------ checkpoints in the SB3 zip layout
    def load_sb3_zip(self, path, load_optimizers=True):
        """weights (+ optimizers, log_ent_coef, counters) of an SB3 2.2.1 SAC zip, e.g. the reference's
        best_models_pretrained/*/SAC_*.zip -- resumes the reference's training (run.py:44 `SAC.load(starting_point)`)."""
        torch = self.torch
        zf = zipfile.ZipFile(path)
        rd = lambda n: torch.load(io.BytesIO(zf.read(n)), map_location=self.device, weights_only=True)
        self.policy.load_state_dict(rd("policy.pth"))
        pv = rd("pytorch_variables.pth")
        if pv and "log_ent_coef" in pv:
            with torch.no_grad():
                self.log_ent_coef.copy_(pv["log_ent_coef"].reshape(1).to(self.device))
        data = json.loads(zf.read("data"))
        self._source_data = data
        self.num_timesteps = int(data.get("num_timesteps", 0))
        self.n_updates = int(data.get("_n_updates", 0))
        if load_optimizers:
            for name, opt in (("actor.optimizer.pth", self.actor_opt), ("critic.optimizer.pth", self.critic_opt),
                              ("ent_coef_optimizer.pth", self.ent_opt)):
                if opt is not None and name in zf.namelist():
                    sd = rd(name)
                    for g in sd["param_groups"]:
                        g["capturable"] = self.device.type == "cuda"
                    opt.load_state_dict(sd)
                    if self.device.type == "cuda":   # capturable Adam keeps `step` on the device
                        for st in opt.state.values():
                            if "step" in st and not st["step"].is_cuda:
                                st["step"] = st["step"].to(self.device)
        self._graph = None
        return self

    def save(self, path):
        """SB3 zip layout: policy.pth / pytorch_variables.pth / *.optimizer.pth / data / _stable_baselines3_version.
        A learner resumed from an SB3 checkpoint (load_sb3_zip) re-uses that checkpoint's `data` entry with the counters
        updated, so `SAC.load` reads the result.  A learner started from scratch has no serialized SB3 `Space` /
        `policy_class` objects to put there: its `data` is a plain-JSON record of hyper-parameters and shapes, which
        `SacActor` and `load_sb3_zip` read but `SAC.load` rejects (only policy.pth and the optimizer files are
        SB3-compatible); a warning says so."""
        torch = self.torch
        if not self._source_data:
            import warnings
            warnings.warn("SACLearner.save: not resumed from an SB3 checkpoint -- the zip's `data` entry is plain JSON; "
                          "stable_baselines3.SAC.load cannot read it (policy.pth / optimizer files are compatible)")
        def dump(obj):
            b = io.BytesIO()
            torch.save(obj, b)
            return b.getvalue()
        cpu = lambda sd: {k: (v.detach().cpu() if hasattr(v, "detach") else v) for k, v in sd.items()}
        def opt_sd(opt):
            sd = opt.state_dict()
            for st in sd["state"].values():
                for k, v in list(st.items()):
                    if hasattr(v, "detach"):
                        st[k] = v.detach().cpu()
            for g in sd["param_groups"]:
                g["capturable"] = False
            return sd
        data = dict(self._source_data) if self._source_data else {
            "learning_rate": self.actor_opt.param_groups[0]["lr"], "gamma": self.gamma, "tau": self.tau,
            "batch_size": self.batch_size, "buffer_size": self.buffer.capacity, "learning_starts": self.learning_starts,
            "gradient_steps": self.gradient_steps, "target_update_interval": self.target_update_interval,
            "ent_coef": "auto" if self.auto_ent else float(self.log_ent_coef.exp()), "target_entropy": self.target_entropy,
            "use_sde": False, "n_envs": 1,
            "observation_space": {"_shape": [self.obs_dim], "dtype": "float64"},
            "action_space": {"_shape": [self.act_dim], "dtype": "float32",
                             "low_repr": str(self.low.cpu().numpy().tolist()), "high_repr": str(self.high.cpu().numpy().tolist())},
        }
        data["num_timesteps"], data["_n_updates"] = int(self.num_timesteps), int(self.n_updates)
        with zipfile.ZipFile(path, "w") as zf:
            zf.writestr("data", json.dumps(data))
            zf.writestr("policy.pth", dump(cpu(self.policy.state_dict())))
            zf.writestr("pytorch_variables.pth", dump({"log_ent_coef": self.log_ent_coef.detach().cpu()}))
            zf.writestr("actor.optimizer.pth", dump(opt_sd(self.actor_opt)))
            zf.writestr("critic.optimizer.pth", dump(opt_sd(self.critic_opt)))
            if self.ent_opt is not None:
                zf.writestr("ent_coef_optimizer.pth", dump(opt_sd(self.ent_opt)))
            zf.writestr("_stable_baselines3_version", "2.2.1")
        return path
