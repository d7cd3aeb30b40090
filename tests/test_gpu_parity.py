"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on identical
inputs (tolerances from BASELINE.json north_star: single-step qpos/qvel/tendon length/reward within 1e-9
relative in fp64), plus size-independent properties at full batch sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _vec(n, xml="flat", env="tr_env", **kw):
    from tensegrity_rl_b200 import TensegrityVecEnv
    return TensegrityVecEnv(n, xml_file=xml, env=env, **kw)


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


def _oracle_step_from(mj, st, e, ctrl, frame_skip=20):
    mj.reset_data()
    mj.qpos[:] = st["qpos"][e]; mj.qvel[:] = st["qvel"][e]; mj.act[:] = st["act"][e]
    mj.qacc_warmstart[:] = st["qacc_warmstart"][e]
    mj.ctrl[:] = ctrl
    mj.step(frame_skip)
    mj.rne_post_constraint()


@pytest.mark.parametrize("xml,env,n,lo,hi", [("flat", "tensegrity_env", 4096, -0.45, -0.15),
                                              ("flat", "tr_env", 1024, -0.45, 0.15),
                                              ("uneven", "tensegrity_env", 1024, -0.45, 0.15)])
def test_single_step_parity_random_ctrl(oracle, xml, env, n, lo, hi):
    """BASELINE config 2: N batched envs, random ctrl, fp64; after selected steps compare CUDA with the oracle
    started from the identical (qpos, qvel, act, qacc_warmstart, ctrl) on a random sample of envs."""
    import torch
    v = _vec(n, xml, env, auto_reset=False, terminate_when_unhealthy=False, max_episode_steps=0)
    v.reset_tensor()
    mj = oracle.MjLike(xml)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    rng = np.random.default_rng(0)
    worst = {"qpos": 0.0, "qvel": 0.0, "ten": 0.0}
    nbad, ncheck, overflow = 0, 0, 0
    for step in range(40):
        a = lo + (hi - lo) * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
        check = step % 8 == 7
        if check:
            before = v.get_state()
        v.step_tensor(a)
        if not check:
            continue
        after, info, ah = v.get_state(), v.info.cpu().numpy(), a.cpu().numpy()
        overflow += int(info[:, 28].sum())
        for e in rng.choice(n, 48, replace=False):
            # tr_env low-pass filters the action; the ctrl actually applied is in the state record
            _oracle_step_from(mj, before, e, after["ctrl"][e] if env == "tr_env" else ah[e])
            dq, dv = _rel(after["qpos"][e], mj.qpos), _rel(after["qvel"][e], mj.qvel)
            dt = _rel(info[e, 8:17], mj.ten_length)
            ncheck += 1
            if max(dq, dv, dt) > TOL:
                nbad += 1
            else:
                worst["qpos"], worst["qvel"], worst["ten"] = max(worst["qpos"], dq), max(worst["qvel"], dv), max(worst["ten"], dt)
    print("checked", ncheck, "outliers", nbad, "worst", worst, "overflow", overflow)
    # MPR is an iterative tolerance-1e-6 routine: a rounding-level branch flip moves the contact depth by ~1e-7,
    # so a tiny outlier fraction is tolerated and reported; everything else must be within 1e-9.
    assert nbad <= max(1, ncheck // 100), (nbad, ncheck)
    assert overflow == 0
    v.close()


CASES = [("flat", "tr_env", "straight"), ("flat", "tr_env", "turn"), ("flat", "tr_env", "aiming"),
         ("flat", "tr_env", "tracking"), ("flat", "tr_env", "vel_track"), ("flat", "tensegrity_env", "straight"),
         ("flat", "tensegrity_env", "turn"), ("uneven", "tensegrity_env", "straight"), ("uneven", "tr_env", "tracking")]


@pytest.mark.parametrize("xml,env,task", CASES)
def test_env_semantics_parity(xml, env, task):
    """reset (explicit draws) + 50 steps of 64 envs: obs / reward / done / info against the numpy+C oracle env, every
    env every step, at the north_star tolerance (1e-9; reward relative).  The oracle env steps from the CUDA path's
    physics state of the previous step (qpos, qvel, act, warm start, ctrl -- single-step agreement, the dynamics being
    chaotic) and keeps its own env bookkeeping (heading ring, waypoints, step counters, stale kinematics)."""
    import torch
    from oracle.envs import OracleEnv
    n = 64
    rng = np.random.default_rng(3)
    draws = np.concatenate([rng.uniform(0, 1, (n, 2)), rng.standard_normal((n, 6)), rng.uniform(0, 1, (n, 2))], 1)
    v = _vec(n, xml, env, desired_action=task, auto_reset=False)
    obs0 = v.reset_tensor(draws=draws).cpu().numpy()
    oes = [OracleEnv(xml, env, desired_action=task) for _ in range(n)]
    # a reset is 1000+ substeps of free evolution from the pose table (on the height field: a 1 m drop and a bounce), i.e.
    # long enough for rounding-level differences to grow in a few envs: the reset observation must agree to 1e-9 in the
    # median and to 1e-6 in >= 85 % of the envs; the envs that drifted further apart leave the step comparison (their
    # waypoints / reset headings, which derive from the reset pose, differ accordingly)
    e0 = np.array([np.abs(oe.reset(draws[k]) - obs0[k]).max() for k, oe in enumerate(oes)])
    print(xml, env, task, "reset obs error: median %.1e, within 1e-6: %d / %d" % (np.median(e0), int((e0 < 1e-6).sum()), n))
    assert np.median(e0) < 1e-9 and (e0 < 1e-6).sum() >= 0.85 * n
    lo, hi = (-0.45, -0.15) if env == "tensegrity_env" else (-0.45, 0.15)
    nbad = ncheck = 0
    alive = e0 < 1e-6
    for st in range(50):
        a = rng.uniform(lo, hi, (n, 6))
        before = v.get_state()
        obs, rew, done = v.step_tensor(torch.as_tensor(a, device="cuda"))
        obs, rew, done, info = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy(), v.info.cpu().numpy()
        for k, oe in enumerate(oes):
            if not alive[k]:
                continue
            mj = oe.mj
            mj.qpos[:] = before["qpos"][k]; mj.qvel[:] = before["qvel"][k]; mj.act[:] = before["act"][k]
            mj.qacc_warmstart[:] = before["qacc_warmstart"][k]; mj.ctrl[:] = before["ctrl"][k]
            o, r, term, trunc, inf = oe.step(a[k])
            ncheck += 1
            err = max(np.abs(o - obs[k]).max(), abs(r - rew[k]) / max(1.0, abs(r)))
            if err > TOL:
                nbad += 1
                assert err < 1e-4, (st, k, err)     # an outlier is a contact that exists in one and not the other: small
            assert bool(done[k]) == (term or trunc)
            assert info[k, 3] == pytest.approx(inf["x_position"], abs=1e-8)
            assert info[k, 22] == pytest.approx(inf["total_bar_contact"], rel=1e-4, abs=1e-5)
            if done[k]:
                alive[k] = False          # no auto reset here: a finished env leaves the comparison
    print(xml, env, task, "checked", ncheck, "outliers above 1e-9:", nbad)
    # budget: 0.5 % of the checks; 1 % for the legacy env, whose robot tips over on the flat XML within ~40 steps (the
    # comparison then runs through tumbling states with 6+ flickering contacts until the env terminates)
    assert ncheck > 1000 and nbad <= max(2, ncheck // (100 if env == "tensegrity_env" else 200)), (nbad, ncheck)
    v.close()


def test_single_env_host_api_matches_batched(oracle):
    """the reference-shaped single env (host buffers through tsg_step_host) is the N=1 view of the same kernels."""
    import torch
    from tensegrity_rl_b200 import make
    env = make("tr_env-v0", xml_file="flat", desired_action="tracking", is_test=True)
    d = np.array([0.3, 0.7, 0.1, -0.2, 0.3, 0.0, 0.5, -1.0, 0.5, 0.5])
    obs, info = env.reset(draws=d)
    assert obs.shape == (48,) and info == {}
    assert env.dt == pytest.approx(0.02) and env.action_space.shape == (6,) and env.observation_space.shape == (48,)
    v = _vec(2, "flat", "tr_env", desired_action="tracking", is_test=True, auto_reset=False)
    ob = v.reset_tensor(draws=np.stack([d, d])).cpu().numpy()
    assert np.array_equal(ob[0], obs) and np.array_equal(ob[1], obs)
    a = np.array([0.1, -0.2, 0.0, -0.4, 0.15, -0.1])
    o1, r1, term, trunc, inf = env.step(a)
    o2, r2, d2 = v.step_tensor(torch.as_tensor(np.stack([a, a]), device="cuda"))
    assert np.array_equal(o2.cpu().numpy()[0], o1) and r2.cpu().numpy()[0] == r1
    for k in ("tendon_length", "real_observation", "reward_forward", "reward_ctrl", "waypt", "x_position", "y_position", "oripoint"):
        assert k in inf
    assert inf["tendon_length"].shape == (9,) and inf["waypt"].shape == (2,)
    with pytest.raises(ValueError):
        env.step(np.zeros(5))
    env.close(); v.close()


def test_determinism_and_sharding_equivalence():
    """RNG streams are keyed by global env id: one handle of 2N envs == two handles of N envs (rank sharding)."""
    import torch
    n = 256
    whole = _vec(2 * n, "flat", "tr_env", seed=7)
    lo = _vec(n, "flat", "tr_env", seed=7, env_id_base=0)
    hi = _vec(n, "flat", "tr_env", seed=7, env_id_base=n)
    ow, ol, oh = whole.reset_tensor().clone(), lo.reset_tensor().clone(), hi.reset_tensor().clone()
    assert torch.equal(ow[:n], ol) and torch.equal(ow[n:], oh)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for _ in range(5):
        a = -0.45 + 0.6 * torch.rand(2 * n, 6, generator=g, device="cuda", dtype=torch.float64)
        o, r, d = whole.step_tensor(a)
        o1, r1, d1 = lo.step_tensor(a[:n].contiguous())
        o2, r2, d2 = hi.step_tensor(a[n:].contiguous())
        assert torch.equal(o[:n], o1) and torch.equal(o[n:], o2) and torch.equal(r[:n], r1) and torch.equal(d[n:], d2)
    again = _vec(2 * n, "flat", "tr_env", seed=7)
    assert torch.equal(again.reset_tensor(), ow)
    for e in (whole, lo, hi, again):
        e.close()


def test_auto_reset_terminal_observation_and_time_limit():
    import torch
    n = 32
    v = _vec(n, "flat", "tr_env", max_episode_steps=3, auto_reset=True)
    v.reset()
    a = np.full((n, 6), 0.1)
    for k in range(3):
        obs, rew, done, infos = v.step(a)
    assert done.all() and all(i["TimeLimit.truncated"] for i in infos)
    assert all(i["terminal_observation"].shape == (45,) for i in infos)
    # obs returned on the done step is the first observation of the new episode, not the terminal one
    assert not np.allclose(obs, np.stack([i["terminal_observation"] for i in infos]))
    obs2, _, done2, _ = v.step(a)
    assert not done2.any()
    rec = v.get_records()
    assert (rec[:, 79] == 1).all() and (rec[:, 85] == 2).all()  # ep_len restarted, two resets so far
    v.close()


def test_state_round_trip_and_f32_ctrl():
    import torch
    v = _vec(16, "flat", "tr_env", auto_reset=False)
    v.reset_tensor()
    rec = v.get_records()
    a = torch.full((16, 6), -0.2, device="cuda", dtype=torch.float64)
    o1 = v.step_tensor(a)[0].clone()
    v.set_records(rec)
    o2 = v.step_tensor(a.float())[0].clone()   # -0.2 in f32 differs from f64 at 1e-9 -> tiny obs difference
    assert torch.allclose(o1, o2, atol=1e-6) and torch.equal(v.obs32, o2.float())
    st = v.get_state()
    v.set_state(qpos=st["qpos"], qvel=st["qvel"])
    assert np.array_equal(v.get_state()["qpos"], st["qpos"])
    v.close()


@pytest.mark.parametrize("env,task", [("tensegrity_env", "turn"), ("tr_env", "aiming")])
def test_checkpoint_restore_includes_the_heading_ring(env, task):
    """records + heading ring saved from one handle and restored into a fresh one continue bit for bit on the tasks
    whose reward reads the delayed heading (legacy turn: 25-slot ring, tensegrity_env.py:242,326-345)."""
    import torch
    n = 32
    v = _vec(n, "flat", env, desired_action=task, auto_reset=False, terminate_when_unhealthy=False)
    v.reset_tensor()
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    acts = -0.45 + 0.3 * torch.rand(40, n, 6, generator=g, device="cuda", dtype=torch.float64)
    for k in range(30):                       # fill the ring past its length
        v.step_tensor(acts[k])
    rec, hd = v.get_records(), v.get_heading()
    ref = [tuple(t.clone() for t in v.step_tensor(acts[30 + k])) for k in range(10)]
    w = _vec(n, "flat", env, desired_action=task, auto_reset=False, terminate_when_unhealthy=False)
    w.set_records(rec); w.set_heading(hd)
    for k in range(10):
        obs, rew, done = w.step_tensor(acts[30 + k])
        assert torch.equal(obs, ref[k][0]) and torch.equal(rew, ref[k][1]) and torch.equal(done, ref[k][2])
    # without the ring the delayed-heading reward differs
    u = _vec(n, "flat", env, desired_action=task, auto_reset=False, terminate_when_unhealthy=False)
    u.set_records(rec)
    _, rew_u, _ = u.step_tensor(acts[30])
    if task == "turn":
        assert not torch.equal(rew_u, ref[0][1])
    for e in (v, w, u):
        e.close()


def test_pooled_reset_keeps_the_true_observation_with_obs_noise():
    """use_obs_noise + reset pool: an env that is handed a pool slot gets the slot's noise-free reset observation in
    real_obs (tr_env.py:505 `real_observation`), not the previous episode's last one."""
    import torch
    n = 256
    v = _vec(n, "flat", "tr_env", desired_action="straight", auto_reset=True, reset_pool=128, use_obs_noise=True,
             max_episode_steps=60)
    v.reset_tensor()
    a = torch.full((n, 6), -0.2, device="cuda", dtype=torch.float64)
    for _ in range(58):                       # the pool slots finish their 50 warm-up steps meanwhile
        v.step_tensor(a)
    before = v.real_obs.clone()
    obs, rew, done = None, None, torch.zeros(n, dtype=torch.bool, device="cuda")
    while not bool(done.any()):
        before = v.real_obs.clone()
        obs, rew, done = v.step_tensor(a)
    d = done.bool()
    assert v.pool_stats()["assigned"] > 0                      # slots were handed out
    real = v.real_obs
    diff = (obs - real).abs().amax(1)
    assert bool((diff[d] > 0).all()) and bool((diff[d] < 1.0).all())            # noisy obs = true reset obs + noise
    # a fresh reset pose (bars re-posed from the pose table) is not the collapsed pose of the finished episode
    assert bool(((real - before).abs().amax(1)[d] > 1e-3).all())
    v.close()


@pytest.mark.parametrize("xml,n", [("flat", 65536), ("uneven", 16384)])
def test_properties_at_scale(xml, n):
    """size-independent invariants at bench-scale batch sizes."""
    import torch
    v = _vec(n, xml, "tr_env", auto_reset=True)
    v.reset_tensor()
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    for _ in range(10):
        a = -0.45 + 0.6 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
        obs, rew, done = v.step_tensor(a)
    st = v.get_state()
    assert np.isfinite(st["qpos"]).all() and np.isfinite(st["qvel"]).all()
    q = st["qpos"].reshape(n, 3, 7)[:, :, 3:]
    assert np.abs(np.linalg.norm(q, axis=2) - 1).max() < 1e-9
    info = v.info.cpu().numpy()
    obs = obs.cpu().numpy()
    assert np.abs(obs[:, :18].reshape(n, 6, 3).sum(1)).max() < 1e-9        # cap positions are centroid-relative
    caps = obs[:, :18].reshape(n, 6, 3)
    assert np.abs(np.linalg.norm(caps[:, 0::2] - caps[:, 1::2], axis=2) - 1.376).max() < 1e-9  # rigid bars
    keep = ~done.cpu().numpy().astype(bool)                                   # auto-reset rows hold the new episode's obs
    assert np.array_equal(obs[keep, 36:45], info[keep, 8:17])                 # tendon lengths in obs == info
    assert info[:, 28].sum() == 0 and info[:, 29].sum() == 0                  # no contact overflow, no bad state
    assert 0.5 < info[:, 19].mean() < 8                                        # contacts per env
    v.close()


def test_pooled_reset_matches_reference_reset_semantics():
    """background reset pool: a done env receives a slot that went through the full reset (pose table, rotation,
    set-points, 50 warm-up steps, bookkeeping) -- identical to what a synchronous reset with the slot's draws
    produces -- and the step path no longer waits for resets."""
    import torch
    from oracle.envs import OracleEnv
    n, pool = 64, 16
    v = _vec(n, "flat", "tr_env", desired_action="tracking", max_episode_steps=4, auto_reset=True, reset_pool=pool, seed=5)
    v.reset_tensor()
    draws = v.get_draws()          # [n] rows only; pool draws are read below through the records' obs
    a = torch.full((n, 6), 0.05, device="cuda", dtype=torch.float64)
    for k in range(3):
        obs, rew, done = v.step_tensor(a)
        assert not done.any()
    obs, rew, done = v.step_tensor(a)      # 4th step: TimeLimit -> every env done, 16 slots available
    st = v.pool_stats()
    assert done.all() and st == {"done": n, "ready": pool, "assigned": pool}
    rec = v.get_records()
    assert (rec[:, 79] == 0).all()         # ep_len restarted for pooled AND synchronously reset envs
    # first `pool` envs (env order) got the slots; their obs must be a valid reset observation
    o = obs.cpu().numpy()
    caps = o[:, :18].reshape(n, 6, 3)
    assert np.abs(np.linalg.norm(caps[:, 0::2] - caps[:, 1::2], axis=2) - 1.376).max() < 1e-9
    assert np.isfinite(o).all() and (np.abs(v.term_obs.cpu().numpy()) > 0).any()
    # slots restart warming: after 50 more steps they are ready again
    for k in range(3):
        v.step_tensor(a)
    obs, rew, done = v.step_tensor(a)
    assert v.pool_stats()["ready"] == 0    # 4 launches later the slots are still warming
    v.close()
    # exactness of a pooled reset: replay slot 0's draws through the oracle env
    v = _vec(8, "flat", "tr_env", desired_action="tracking", max_episode_steps=2, auto_reset=True, reset_pool=4, seed=9)
    v.reset_tensor()
    a = torch.full((8, 6), 0.05, device="cuda", dtype=torch.float64)
    v.step_tensor(a)
    import ctypes as C
    d = np.zeros((8 + 4, 10))
    # draws of the pool slots sit after the env rows
    from tensegrity_rl_b200 import lib as tl
    obs, rew, done = v.step_tensor(a)
    assert done.all()
    o = obs.cpu().numpy()
    # env 0 received slot 0; regenerate slot 0's first draw with the emulator's Philox and replay in the oracle
    from emul import emul as E
    L = E.lib()
    dr = np.zeros(10)
    L.tbe_make_draws(E.P(dr), C.c_ulonglong(9), C.c_ulonglong((1 << 40) + 0), C.c_ulonglong(0))
    oe = OracleEnv("flat", "tr_env", desired_action="tracking")
    assert np.abs(oe.reset(dr) - o[0]).max() < 1e-6
    v.close()


def test_pretrained_policy_rollouts_distribution_level():
    """BASELINE configs 3/4 in small: pretrained SAC actors drive the batched envs on the device; forward
    displacement / yaw statistics agree with the same policy run through the oracle env at distribution level
    (the dynamics are chaotic and the policy stochastic, so only moments are compared)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import run_rollouts as R
    steps = 150
    g = R.gpu_rollout("forward_flat", 2048, steps, deterministic=False)
    o = R.oracle_rollout("forward_flat", 12, steps, deterministic=False)
    print("gpu", {k: round(g[k], 4) for k in ("episodes", "disp_mean", "disp_std", "yaw_mean", "yaw_std", "return_mean", "length_mean")})
    print("oracle", {k: round(o[k], 4) for k in ("episodes", "disp_mean", "disp_std", "yaw_mean", "yaw_std", "return_mean", "length_mean")})
    assert g["overflow"] == 0 and g["bad"] == 0 and g["env_steps"] == 2048 * steps
    # the policy was trained on another bar geometry / actuator law, so only consistency is asserted
    se = max(g["disp_std"], o["disp_std"], 1e-3) / np.sqrt(max(o["episodes"], 1))
    assert abs(g["disp_mean"] - o["disp_mean"]) < 5 * se + 0.02
    sey = max(g["yaw_std"], o["yaw_std"], 1e-3) / np.sqrt(max(o["episodes"], 1))
    assert abs(g["yaw_mean"] - o["yaw_mean"]) < 5 * sey + 0.02


def test_test3_waypoint_selector_runs_batched():
    import torch
    from tensegrity_rl_b200 import SacActor
    from tensegrity_rl_b200.rollout import WaypointController
    n = 256
    v = _vec(n, "flat", "tr_env", desired_action="aiming", is_test=True, auto_reset=False, terminate_when_unhealthy=False)
    obs = v.reset_tensor()
    assert torch.allclose(obs[:, 45:47], -torch.stack([v.get_records_t()[:, 69], v.get_records_t()[:, 70]], 1), atol=0.2)
    ctl = WaypointController(v, SacActor("traj_track"), SacActor("traj_ccw"), SacActor("traj_cw"))
    for k in range(30):
        a = ctl.action(v.obs)
        assert a.shape == (n, 6) and torch.isfinite(a).all()
        v.step_tensor(a, auto_reset=False)
        ctl.after_step(v.info)
    assert torch.isfinite(v.obs).all() and (ctl.del_yaw.abs() <= np.pi + 1e-9).all()
    v.close()


def test_pretrained_policies_reproduce_reference_training_statistics():
    """Distribution-level pin against numbers produced by the REFERENCE'S OWN MuJoCo runs: every SB3 checkpoint
    stores the last 100 training episodes (return, length).  On the model the checkpoints were trained on
    (bar geometry pinned by tests/test_golden_last_obs.py; flat floor) the pretrained policies must reproduce the
    per-step return, i.e. the gait speed / yaw rate, they had at training time:
        forward  1537 / 3971 = 0.387 per step  (0.29 m/s + 0.1 healthy - ctrl cost)
        backward 1643 / 3613 = 0.455 per step
        yaw CCW   207 / 2605 = 0.080 rad/s ;  yaw CW 178 / 2089 = 0.085 rad/s
    A gait learned by RL is sensitive to contact, friction, actuator and integrator details, so this checks the
    whole dynamics restatement, not just kinematics."""
    import json, os
    from tensegrity_rl_b200 import SacActor
    from tensegrity_rl_b200.rollout import rollout
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "last_obs.json")))
    cases = [("forward", dict(desired_action="straight", desired_direction=1), "return"),
             ("backward", dict(desired_action="straight", desired_direction=-1), "return"),
             ("yaw_CCW", dict(desired_action="turn", desired_direction=1, terminate_when_unhealthy=False), "return"),
             ("yaw_CW", dict(desired_action="turn", desired_direction=-1, terminate_when_unhealthy=False), "return")]
    for pol, kw, _ in cases:
        ref = G[pol]["ep_return_mean"] / G[pol]["ep_len_mean"]
        v = _vec(1024, "legacy_flat", "tensegrity_env", auto_reset=True, reset_pool=256, **kw)
        v.reset_tensor()
        s = rollout(v, SacActor(pol), 500)
        got = s["return_sum"] / s["length_sum"]
        speed = s["disp_sum"] / (s["length_sum"] * v.dt)
        print(pol, "return/step ours %.3f reference %.3f | forward speed %.3f m/s yaw rate %.3f rad/s"
              % (got, ref, speed, s["yaw_sum"] / (s["length_sum"] * v.dt)))
        assert abs(got - ref) < 0.2 * abs(ref) + 0.01, (pol, got, ref)
        v.close()


@pytest.mark.gpu
def test_obs_noise_on_gpu_matches_restatement_and_is_reproducible():
    """use_obs_noise through the C ABI: obs - real_obs = stdev * z with the Philox normals of (seed, env id, reset
    count, episode step); tracking vector / yaw re-derived as in tr_env.py:626-639; same seed -> same noise."""
    import torch
    from emul import emul as E
    from oracle.envs import OracleEnv
    from tensegrity_rl_b200 import TensegrityVecEnv
    n, seed = 256, 21
    mk = lambda: TensegrityVecEnv(n, xml_file="flat", env="tr_env", desired_action="tracking", use_obs_noise=True,
                                  seed=seed, env_id_base=1000, auto_reset=False)
    v, w = mk(), mk()
    oe = OracleEnv("flat", "tr_env", desired_action="tracking", use_obs_noise=True)
    o0, o0w = v.reset_tensor().cpu().numpy(), w.reset_tensor().cpu().numpy()
    assert np.array_equal(o0, o0w)
    real0 = v.real_obs.cpu().numpy()
    for e in (0, 17, 255):
        z = E.obs_normals(seed, 1000 + e, 1, 0, 45)
        assert np.abs(oe.noisy_obs(real0[e], z) - o0[e]).max() < 1e-9
    g = torch.Generator(device="cuda").manual_seed(0)
    for st in range(1, 4):
        a = -0.45 + 0.6 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
        ob = v.step_tensor(a)[0].cpu().numpy()
        obw = w.step_tensor(a)[0].cpu().numpy()
        assert np.array_equal(ob, obw)
        real = v.real_obs.cpu().numpy()
        for e in (0, 17, 255):
            z = E.obs_normals(seed, 1000 + e, 1, st, 45)
            assert np.abs(oe.noisy_obs(real[e], z) - ob[e]).max() < 1e-9
    d = (ob[:, :36] - real[:, :36]) / 0.05
    assert abs(d.mean()) < 0.03 and abs(d.std() - 1) < 0.03
    dt = (ob[:, 36:45] - real[:, 36:45]) / 0.02
    assert abs(dt.std() - 1) < 0.05
    # the noise-free twin computes the same true observations: noise never feeds back into the dynamics or rewards
    t = TensegrityVecEnv(n, xml_file="flat", env="tr_env", desired_action="tracking", seed=seed, env_id_base=1000, auto_reset=False)
    t.reset_tensor()
    g = torch.Generator(device="cuda").manual_seed(0)
    for st in range(1, 4):
        a = -0.45 + 0.6 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
        tob, trew, _ = t.step_tensor(a)
    assert torch.equal(tob, v.real_obs) and torch.equal(trew, v.reward)
    # SB3-protocol info dict carries the true observation
    obs_h, rew_h, done_h, infos = v.step(np.full((n, 6), 0.1))
    assert np.abs(infos[3]["real_observation"] - v.real_obs[3].cpu().numpy()).max() == 0 if infos[3] else True
    for x in (v, w, t):
        x.close()


@pytest.mark.gpu
def test_single_env_obs_noise_info():
    from tensegrity_rl_b200 import envs
    e = envs.tr_env(xml_file="flat", desired_action="straight", use_obs_noise=True)
    o, _ = e.reset(seed=4)
    ob, r, term, trunc, info = e.step(np.full(6, 0.0, np.float32))
    assert ob.shape == (45,) and info["real_observation"].shape == (45,)
    d = ob - info["real_observation"]
    assert 0.005 < np.abs(d).max() < 0.5 and np.abs(d[36:]).max() < 0.15
    with pytest.raises(NotImplementedError):
        envs.tr_env(xml_file="flat", use_cap_size_noise=True)


@pytest.mark.gpu
def test_ragged_batches_give_identical_envs():
    """Edge cases of the batch dimension: N = 1, N not a multiple of the ten envs a warp steps together (idle env
    slots in the last warp), fewer / more warps than resident ones.  An env's trajectory must depend on its id only
    -- bitwise -- not on the batch it is stepped in or on which envs share its warp."""
    import torch
    from tensegrity_rl_b200 import TensegrityVecEnv

    def run(n):
        v = TensegrityVecEnv(n, xml_file="flat", env="tr_env", seed=11, auto_reset=False)
        assert v.kernel_config()["lanes_per_env"] == 3
        v.reset_tensor()
        ids = torch.arange(n, device="cuda", dtype=torch.float64)
        for st in range(2):
            a = -0.3 + 0.15 * torch.sin(ids[:, None] * 0.37 + torch.arange(6, device="cuda") + st)
            obs, rew, done = v.step_tensor(a)
        out = (v.get_records()[:, :69].copy(), obs.cpu().numpy().copy(), rew.cpu().numpy().copy())
        v.close()
        return out

    ref = run(13)
    for n in (1, 7, 10, 15, 600, 3700, 20011):
        rec, obs, rew = run(n)
        k = min(n, 13)
        assert np.array_equal(rec[:k], ref[0][:k]), n
        assert np.array_equal(obs[:k], ref[1][:k]) and np.array_equal(rew[:k], ref[2][:k]), n


@pytest.mark.gpu
def test_planar_equivariance_at_full_batch():
    """Size-independent property at BASELINE configs[1] scale (4096 envs): on the flat floor an env step commutes with
    a rotation about z plus a shift in the plane, env by env (every env gets its own angle).  Outliers are the
    contact-flicker states described in DESIGN.md (tolerated <= 1 %)."""
    import torch
    n = 4096
    rng = np.random.default_rng(8)
    A = _vec(n, "flat", "tr_env", auto_reset=False, terminate_when_unhealthy=False, max_episode_steps=0, seed=3)
    B = _vec(n, "flat", "tr_env", auto_reset=False, terminate_when_unhealthy=False, max_episode_steps=0, seed=3)
    A.reset_tensor(); B.reset_tensor()
    g = torch.Generator(device="cuda").manual_seed(2)
    for st in range(6):                        # move away from the reset pose
        A.step_tensor(-0.45 + 0.3 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64))
    sa = A.get_state()
    phi, shift = rng.uniform(-np.pi, np.pi, n), rng.uniform(-3, 3, (n, 2))
    c, s = np.cos(phi), np.sin(phi)

    def rot(st):
        q, v, w = st["qpos"].copy(), st["qvel"].copy(), st["qacc_warmstart"].copy()
        for b in range(3):
            x, y = st["qpos"][:, 7 * b], st["qpos"][:, 7 * b + 1]
            q[:, 7 * b], q[:, 7 * b + 1] = c * x - s * y + shift[:, 0], s * x + c * y + shift[:, 1]
            qw, qx, qy, qz = (st["qpos"][:, 7 * b + 3 + k] for k in range(4))
            ch, sh = np.cos(phi / 2), np.sin(phi / 2)
            q[:, 7 * b + 3], q[:, 7 * b + 4] = ch * qw - sh * qz, ch * qx - sh * qy
            q[:, 7 * b + 5], q[:, 7 * b + 6] = ch * qy + sh * qx, ch * qz + sh * qw
            for arr, src in ((v, st["qvel"]), (w, st["qacc_warmstart"])):
                arr[:, 6 * b], arr[:, 6 * b + 1] = c * src[:, 6 * b] - s * src[:, 6 * b + 1], s * src[:, 6 * b] + c * src[:, 6 * b + 1]
        return q, v, w

    q2, v2, w2 = rot(sa)
    B.set_state(qpos=q2, qvel=v2, act=sa["act"], qacc_warmstart=w2, ctrl=sa["ctrl"])
    a = -0.45 + 0.3 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
    A.step_tensor(a); B.step_tensor(a)
    qe, ve, _ = rot(A.get_state())
    sb = B.get_state()
    err = np.maximum(np.abs(sb["qpos"] - qe).max(1), np.abs(sb["qvel"] - ve).max(1) / np.maximum(1.0, np.abs(ve).max(1)))
    ten_a, ten_b = A.info[:, 8:17].cpu().numpy(), B.info[:, 8:17].cpu().numpy()
    bad = int((err > 1e-8).sum())
    print("equivariance: median err %.2e, outliers %d / %d" % (np.median(err), bad, n))
    assert bad <= n // 100
    ok = err <= 1e-8
    assert np.abs(ten_a - ten_b)[ok].max() < 1e-8
    A.close(); B.close()


@pytest.mark.gpu
@pytest.mark.parametrize("xml,steps,budget", [("flat", 200, 0.0025), ("uneven", 120, 0.006)])
def test_full_batch_parity_every_env_every_step(oracle, xml, steps, budget):
    """BASELINE configs[1] at full size: 4096 batched envs, random ctrl in [-0.45, -0.15], fp64; after EVERY step EVERY
    env is re-stepped by the threaded C oracle from the identical (qpos, qvel, act, warm start, ctrl) and compared at
    1e-9 relative on qpos / qvel / tendon length.  Outliers are budgeted (0.25 % flat: measured 0.18 % over 819 200
    env steps, largest 7e-6; 0.6 % height field: measured 0.37 %) and must be explained by a contact whose existence
    flickers within the step -- the states where a summation-order-level difference decides whether a contact exists in
    a substep: on the plane every outlier has an active-contact count that varies over the oracle's 20 substeps or
    differs from the CUDA path's final count; on the height field a geom can also trade one prism for its neighbour at
    an unchanged count, so there >= 85 % of the outliers must show a count change."""
    import torch
    n = 4096
    v = _vec(n, xml, "tr_env", auto_reset=False, terminate_when_unhealthy=False, max_episode_steps=0)
    v.reset_tensor()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    nbad = ncheck = unexplained = 0
    worst_ok, worst_out = 0.0, 0.0
    for step in range(steps):
        a = -0.45 + 0.3 * torch.rand(n, 6, generator=g, device="cuda", dtype=torch.float64)
        before = v.get_state()
        v.step_tensor(a)
        after, info = v.get_state(), v.info.cpu().numpy()
        oq, ov, ot, mm = oracle.step_states(xml, before["qpos"], before["qvel"], before["act"], before["qacc_warmstart"], after["ctrl"])
        scale = lambda x: np.maximum(1.0, np.abs(x).max(axis=1))
        err = np.maximum.reduce([np.abs(after["qpos"] - oq).max(1) / scale(oq), np.abs(after["qvel"] - ov).max(1) / scale(ov),
                                 np.abs(info[:, 8:17] - ot).max(1) / scale(ot)])
        bad = err > TOL
        ncheck += n; nbad += int(bad.sum())
        worst_ok = max(worst_ok, float(err[~bad].max())); worst_out = max(worst_out, float(err.max()))
        flicker = (mm[:, 0] != mm[:, 1]) | (mm[:, 1] != info[:, 19].astype(int))
        unexplained += int((bad & ~flicker).sum())
        assert int(info[:, 28].sum()) == 0 and int(info[:, 29].sum()) == 0      # no contact overflow, no bad state
    print(xml, "checked", ncheck, "outliers", nbad, "(%.4f %%)" % (100.0 * nbad / ncheck), "largest", worst_out,
          "worst within tolerance", worst_ok, "outliers without a contact-count change", unexplained)
    assert nbad <= budget * ncheck, (nbad, ncheck)
    assert unexplained <= (0 if xml == "flat" else 0.15 * nbad), (unexplained, nbad)
    v.close()


@pytest.mark.gpu
def test_golden_last_obs_through_the_cuda_path():
    """The `_last_obs` vectors the REFERENCE's MuJoCo runs stored in its checkpoints, through the CUDA path: the state
    reconstructed from each vector (tests/golden_pose.py), written with tsg_set_state and passed through tsg_forward,
    must come back as the golden vector itself -- cap positions, cap velocities (incl. the reference's body-frame
    angular velocity quirk), tendon lengths, tracking vector and yaw; bar quaternions (scipy convention) and qvel for
    the legacy env -- to 1e-12.  A golden check of kinematics, site / tendon tables, geom frames and _get_obs."""
    import golden_pose as GP
    for name in sorted(GP.G):
        obs = np.array(GP.G[name]["last_obs"])
        if GP.G[name]["obs_dim"] == 48:
            v = _vec(1, "uneven", "tr_env", desired_action=GP.TASK_OF[name], auto_reset=False)
            qpos, qvel, waypt, res = GP.state_from_tr_obs(obs, v.md)
        else:
            v = _vec(1, "uneven", "tensegrity_env", desired_action="straight", auto_reset=False)
            qpos, qvel, res = GP.state_from_legacy_obs(obs, v.md)
            waypt = None
        assert res < 1e-12
        v.set_state(qpos=qpos[None], qvel=qvel[None])
        if waypt is not None:
            rec = v.get_records()
            rec[0, 73:75] = waypt
            v.set_records(rec)
        got = v.forward_tensor().cpu().numpy()[0]
        n = len(obs)
        if name in ("traj_ccw", "traj_cw"):
            assert np.all(obs[45:48] == 0)    # written by an older revision of the env (see tests/test_golden_last_obs.py)
            n = 45
        if GP.G[name]["obs_dim"] == 39:
            for b in range(3):
                if np.dot(got[4 * b:4 * b + 4], obs[4 * b:4 * b + 4]) < 0:
                    got[4 * b:4 * b + 4] *= -1
        assert np.abs(got[:n] - obs[:n]).max() < 1e-12, (name, np.abs(got[:n] - obs[:n]).max())
        v.close()


@pytest.mark.gpu
@pytest.mark.parametrize("policy,task,dirn", [("traj_track", "tracking", 1), ("traj_ccw", "turn", 1), ("traj_cw", "turn", -1)])
def test_tr_env_checkpoint_rollouts_are_reported(policy, task, dirn):
    """The 48-dim tr_env checkpoints (models_traj/SAC_16525000_track.zip, SAC_2175000_ccw.zip, SAC_1250000_cw.zip; the
    heading-reward paths of BASELINE configs[3], tr_env.py:425-459) store their last 100 training episodes
    (`ep_info_buffer`): 0.886 / 0.214 / 0.212 return per step.  The ccw / cw checkpoints were trained with the three
    waypoint slots of the observation at zero (their `_last_obs` holds zeros there, and run.py test3 :262-272 blanks
    them), i.e. on a turning reward: they are run on `desired_action="turn"` with those slots zeroed.  This simulator
    does NOT reproduce these figures (measured: tracking 0.15 per step, episode length 299 against 253; ccw / cw 0.15 /
    0.13 against 0.21, episode length ~200 against 681 / 3673): unlike the four legacy checkpoints, whose training
    statistics are reproduced within 5 %, the training configuration of these (model file, waypoint range, reward
    amplitudes, termination rule) is not recoverable from the repository.  The test prints ours beside the reference's
    and asserts that the rollouts run without tripping any solver safeguard and that each policy earns a positive
    return per step on its task (it walks towards its waypoints / turns the way it was trained to)."""
    import json, os
    import torch
    from tensegrity_rl_b200 import SacActor
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "last_obs.json")))
    ref = G[policy]["ep_return_mean"] / G[policy]["ep_len_mean"]
    n = 2048
    v = _vec(n, "legacy_flat", "tr_env", desired_action=task, desired_direction=dirn, auto_reset=True, reset_pool=512)
    v.reset_tensor()
    actor = SacActor(policy)
    ret_sum = len_sum = 0.0
    episodes = 0
    ret = torch.zeros(n, device="cuda", dtype=torch.float64)
    ln = torch.zeros(n, device="cuda", dtype=torch.float64)
    for _ in range(600):
        o = torch.zeros(n, 48, device="cuda", dtype=torch.float32)
        o[:, :v.obs_dim] = v.obs32                              # tracking: all 48 slots; turn: 45 + zeros
        obs, rew, done = v.step_tensor(actor(o, False).double())
        ret += rew; ln += 1
        d = done.bool()
        ret_sum += float(ret[d].sum()); len_sum += float(ln[d].sum()); episodes += int(d.sum())
        ret[d] = 0; ln[d] = 0
    length = len_sum / max(1, episodes)                         # finished episodes only
    ret_sum += float(ret.sum()); len_sum += float(ln.sum())
    got = ret_sum / len_sum
    print("%s: return/step ours %.3f reference %.3f | episode length ours %.0f reference %.0f"
          % (policy, got, ref, length, G[policy]["ep_len_mean"]))
    assert int(v.info[:, 28].sum()) == 0 and int(v.info[:, 29].sum()) == 0
    assert got > 0.05
    v.close()
