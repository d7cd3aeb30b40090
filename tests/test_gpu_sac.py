"""SAC learner on the real simulator (GPU): device-resident collection, CUDA-graph update, truncation handling."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _env(n, **kw):
    from tensegrity_rl_b200 import TensegrityVecEnv
    return TensegrityVecEnv(n, xml_file="flat", env="tr_env", auto_reset=True, reset_pool=0, **kw)


def test_collection_stores_terminal_observation_and_truncation_flag():
    from tensegrity_rl_b200.sac import SACLearner
    env = _env(64, max_episode_steps=3)
    L = SACLearner(env.obs_dim, 6, device="cuda", learning_starts=10 ** 9, buffer_size=4096, use_cuda_graph=False)
    obs = env.reset_tensor().float()
    seen = []
    for k in range(3):
        prev = obs
        obs = L.collect_step(env, obs)
        seen.append((prev, env.done.clone(), env.term_obs.float().clone(), obs.clone(), env.info[:, 17].clone()))
    assert L.buffer.size == 192 and L.num_timesteps == 192
    prev, done, term, new_obs, terminated = seen[2]
    assert bool(done.all())                                   # TimeLimit at step 3
    rows = slice(128, 192)
    assert torch.equal(L.buffer.obs[rows], prev)
    assert torch.equal(L.buffer.next_obs[rows], term)        # terminal observation, not the reset observation
    assert not torch.equal(term, new_obs)
    assert torch.equal(L.buffer.done[rows], (terminated > 0).float())   # pure time-limit truncations store done = 0
    assert float(L.buffer.act[:192].abs().max()) <= 1.0
    env.close()


def test_cuda_graph_update_trains_on_simulator_data():
    from tensegrity_rl_b200.sac import SACLearner
    env = _env(512)
    L = SACLearner(env.obs_dim, 6, device="cuda", learning_starts=1024, batch_size=256, buffer_size=65536,
                   gradient_steps=8, seed=0)
    L.learn(env, 512 * 10)
    assert L.use_cuda_graph and L._graph is not None
    assert L.buffer.size == 5120 and L.n_updates == 8 * 8     # vec steps 3..10 train (num_timesteps > learning_starts)
    assert torch.isfinite(L.last_losses).all()
    g = torch.Generator(device="cuda").manual_seed(0)
    idx = torch.randint(0, L.buffer.size, (1024,), device="cuda", generator=g)
    batch = L.buffer.gather(idx)
    e1 = torch.randn(1024, 6, device="cuda", generator=g)
    e2 = torch.randn(1024, 6, device="cuda", generator=g)
    c0 = float(L.losses(*batch, e1, e2)[1].detach())
    before = [p.detach().clone() for p in L.policy.actor.parameters()]
    L.update(300)                                             # 300 graph replays on the same buffer
    c1 = float(L.losses(*batch, e1, e2)[1].detach())
    assert c1 < c0 and L.n_updates == 64 + 300
    assert all(not torch.equal(a, b.detach()) for a, b in zip(before, L.policy.actor.parameters()))
    # eager path agrees with the graph path in distribution: one more eager learner on the same data reaches a
    # comparable critic loss (not bitwise: RNG streams differ)
    E = SACLearner(env.obs_dim, 6, device="cuda", batch_size=256, buffer_size=65536, seed=0, use_cuda_graph=False)
    E.buffer.add(L.buffer.obs[:5120], L.buffer.next_obs[:5120], L.buffer.act[:5120], L.buffer.rew[:5120], L.buffer.done[:5120])
    E.update(364)
    ce = float(E.losses(*batch, e1, e2)[1].detach())
    assert ce < c0 and 0.2 < (c1 + 1e-9) / (ce + 1e-9) < 5.0
    env.close()
