"""CPU check of the CUDA SOURCE: csrc/tb_*.cuh compiled as plain C++ and run by a 32-lane fibre warp emulator
(tests/emul: shuffles, votes and warp barriers are rendezvous points between the lanes) against the oracle.
The GPU twin of these tests is tests/test_gpu_parity.py."""
import numpy as np
import pytest

from emul import emul as E
from oracle.envs import OracleEnv


def _copy_state(mj, em):
    em.rec[0:21] = mj.qpos
    em.rec[21:39] = mj.qvel
    em.rec[39:57] = mj.qacc_warmstart
    em.rec[63:69] = mj.act


@pytest.mark.parametrize("model,lo,hi", [("flat", -0.45, -0.15), ("flat", -0.45, 0.15), ("uneven", -0.45, 0.15)])
def test_single_step_state_parity(oracle, model, lo, hi):
    """one env step (20 substeps) from identical (qpos, qvel, act, warmstart, ctrl): 1e-9 relative."""
    mj, em = oracle.MjLike(model), E.Emul(model)
    rng = np.random.default_rng(5)
    worst = 0.0
    for st in range(60):
        ctrl = rng.uniform(lo, hi, 6)
        _copy_state(mj, em)
        mj.ctrl[:] = ctrl
        mj.step(20)
        mj.rne_post_constraint()
        ten, cfrc, stats = em.mj_step(ctrl, 20)
        assert stats[4] == 0 and stats[5] == 0
        for a, b in ((em.qpos, mj.qpos), (em.qvel, mj.qvel), (ten, mj.ten_length)):
            worst = max(worst, np.abs(a - b).max() / max(1.0, np.abs(b).max()))
        assert np.abs(cfrc - mj.cfrc_ext).max() <= 1e-7 * max(1.0, np.abs(mj.cfrc_ext).max())
    assert worst < 1e-9, worst


@pytest.mark.parametrize("model", ["flat", "uneven"])
def test_full_warp_of_desynchronised_envs(oracle, model):
    """ten envs in one warp, each at a different point of its trajectory (different contact sets, Newton iteration and
    line-search counts per substep): the predication that lets them share one instruction stream must not leak
    between envs.  Every env is compared with its own oracle instance."""
    n = 10
    mjs = [oracle.MjLike(model) for _ in range(n)]
    w = E.EmulWarp(model, n)
    rng = np.random.default_rng(3)
    for i, mj in enumerate(mjs):
        for _ in range(2 * i + (30 if model == "uneven" else 0)):   # (the uneven model starts 1 m above its floor)
            mj.ctrl[:] = rng.uniform(-0.45, -0.15, 6)
            mj.step(20)
    iters = set()
    for st in range(6):
        ctrl = rng.uniform(-0.45, -0.15, (n, 6))
        for i, mj in enumerate(mjs):
            w.rec[i, 0:21] = mj.qpos; w.rec[i, 21:39] = mj.qvel; w.rec[i, 39:57] = mj.qacc_warmstart; w.rec[i, 63:69] = mj.act
            mj.ctrl[:] = ctrl[i]
            mj.step(20)
            mj.rne_post_constraint()
        ten, cfrc, stats = w.mj_step(ctrl, 20)
        for i, mj in enumerate(mjs):
            assert np.abs(w.rec[i, 0:21] - mj.qpos).max() < 1e-9
            assert np.abs(w.rec[i, 21:39] - mj.qvel).max() <= 1e-9 * max(1.0, np.abs(mj.qvel).max())
            assert np.abs(ten[i] - mj.ten_length).max() < 1e-9
            assert np.abs(cfrc[i] - mj.cfrc_ext).max() <= 1e-7 * max(1.0, np.abs(mj.cfrc_ext).max())
            assert stats[i, 4] == 0 and stats[i, 5] == 0
        iters.update(stats[:, 1].tolist())
    assert len(iters) > 5   # the envs really did different amounts of solver work


@pytest.mark.parametrize("model,pre", [("flat", 0), ("uneven", 30)])
def test_fp32_mode_single_step_within_1e_4(oracle, model, pre):
    """the optional fp32 mode (north_star: single-step qpos / qvel / tendon length within 1e-4 of the fp64 reference):
    one mj_step from identical states against the fp64 oracle.  The mode keeps the MPR narrow phase and the constraint
    solver in fp64 -- with those in fp32 the same test lands at 5e-3 (MPR's 1e-6 tolerance and the 1e6 spread of the
    Newton Hessian are out of fp32's reach)."""
    n = 10
    mjs = [oracle.MjLike(model) for _ in range(n)]
    w = E.EmulWarp(model, n)
    rng = np.random.default_rng(9)
    for i, mj in enumerate(mjs):
        for _ in range(pre + 3 * i):
            mj.ctrl[:] = rng.uniform(-0.45, -0.15, 6)
            mj.step(20)
    errs = []
    for st in range(40):
        ctrl = rng.uniform(-0.45, -0.15, (n, 6))
        for i, mj in enumerate(mjs):
            w.rec[i, 0:21] = mj.qpos; w.rec[i, 21:39] = mj.qvel; w.rec[i, 39:57] = mj.qacc_warmstart; w.rec[i, 63:69] = mj.act
            mj.ctrl[:] = ctrl[i]
            mj.step(1)
        ten, stats = w.mj_step_f32(ctrl, 1)
        for i, mj in enumerate(mjs):
            errs.append(max(np.abs(w.rec[i, 0:21] - mj.qpos).max(), np.abs(w.rec[i, 21:39] - mj.qvel).max() / max(1.0, np.abs(mj.qvel).max()),
                            np.abs(ten[i] - mj.ten_length).max()))
            assert stats[i, 4] == 0 and stats[i, 5] == 0
        if st % 8 == 7:      # let the trajectories move on between the checks
            for mj in mjs:
                mj.step(19)
    errs = np.array(errs)
    print(model, "fp32 mode single-step error: median %.1e max %.1e" % (np.median(errs), errs.max()))
    assert (errs < 1e-4).mean() >= 0.995 and errs.max() < 1e-3 and np.median(errs) < 2e-5


def test_conservative_prefilter_matches_unfiltered_oracle(oracle):
    """the CUDA source filters bar-bar pairs with an analytic capsule bound before MPR; the oracle runs MPR on
    every pair that passes MuJoCo's bounding-sphere test.  Squeeze the bars together and compare."""
    mj, em = oracle.MjLike("flat"), E.Emul("flat")
    rng = np.random.default_rng(7)
    nbar = 0
    for st in range(80):
        ctrl = rng.uniform(-0.45, -0.35, 6)
        _copy_state(mj, em)
        mj.ctrl[:] = ctrl
        mj.step(5)
        ten, cfrc, stats = em.mj_step(ctrl, 5)
        nbar += sum(1 for c in mj.contacts() if c.geom1 != 0 and c.efc_address >= 0)
        assert np.abs(em.qvel - mj.qvel).max() <= 1e-9 * max(1.0, np.abs(mj.qvel).max())
        assert stats[0] == mj.nefc // 6
    assert nbar > 20  # the scenario really exercises bar-bar contact


CASES = [("flat", "tr_env", "straight"), ("flat", "tr_env", "turn"), ("flat", "tr_env", "aiming"),
         ("flat", "tr_env", "tracking"), ("flat", "tr_env", "vel_track"), ("flat", "tensegrity_env", "straight"),
         ("flat", "tensegrity_env", "turn"), ("uneven", "tensegrity_env", "straight")]


@pytest.mark.parametrize("xml,kind,task", CASES)
def test_env_semantics_parity(xml, kind, task):
    rng = np.random.default_rng(11)
    oe = OracleEnv(xml, kind, desired_action=task)
    em = E.Emul(xml, env_kind=kind, desired_action=task)
    draws = np.concatenate([rng.uniform(0, 1, 2), rng.standard_normal(6), rng.uniform(0, 1, 2)])
    o1, o2 = oe.reset(draws), em.reset(draws)
    assert o1.shape == o2.shape == (oe.cfg.obs_dim,)
    assert np.abs(o1 - o2).max() < 1e-7
    lo, hi = (-0.45, -0.15) if kind == "tensegrity_env" else (-0.45, 0.15)
    for st in range(12):
        a = rng.uniform(lo, hi, 6)
        ob1, r1, t1, tr1, i1 = oe.step(a)
        ob2, r2, d2, info = em.step(a)
        assert np.abs(ob1 - ob2).max() < 1e-7
        assert abs(r1 - r2) <= 1e-7 * max(1.0, abs(r1))
        assert (t1 or tr1) == d2
        assert info[3] == pytest.approx(i1["x_position"], abs=1e-8)
        assert info[5] == pytest.approx(i1["psi"], abs=1e-7)


def test_time_limit_and_step_cap():
    em = E.Emul("flat", env_kind="tr_env", desired_action="tracking", max_episode_steps=3)
    em.reset(np.array([0.1, 0.2, 0, 0, 0, 0, 0, 0, 0.5, 0.5]))
    dones = [em.step(np.full(6, 0.1))[2] for _ in range(3)]
    assert dones == [False, False, True]  # TimeLimit.truncated on the 3rd step


def test_philox_draws_are_keyed_by_env_and_reset_count():
    import ctypes as C
    L = E.lib()
    def draws(seed, env, n):
        d = np.zeros(10)
        L.tbe_make_draws(E.P(d), C.c_ulonglong(seed), C.c_ulonglong(env), C.c_ulonglong(n))
        return d
    a, b, c, d = draws(1, 5, 0), draws(1, 5, 0), draws(1, 6, 0), draws(1, 5, 1)
    assert np.array_equal(a, b) and not np.array_equal(a, c) and not np.array_equal(a, d)
    u = np.array([draws(3, e, 0) for e in range(4000)])
    assert (u[:, [0, 1, 8, 9]] >= 0).all() and (u[:, [0, 1, 8, 9]] < 1).all()
    assert abs(u[:, [0, 1, 8, 9]].mean() - 0.5) < 0.02
    assert abs(u[:, 2:8].mean()) < 0.03 and abs(u[:, 2:8].std() - 1) < 0.03


def _quat_mul(a, b):
    w1, x1, y1, z1 = a; w2, x2, y2, z2 = b
    return [w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2]


@pytest.mark.parametrize("yaw,dx,ncon0,nstep", [(0.0, 0.0, 19, 1), (1.1, 0.02, 10, 3)])
def test_many_contacts_per_bar(oracle, yaw, dx, ncon0, nstep):
    """two bars pressed flat into the floor and into each other: up to 19 contacts (9-10 per bar, all owned by the
    bar's lane and spilled past the first; the usual count is 1); results must still match the dense oracle.

    yaw = 0 puts the two bars on ONE axis, overlapping: the deepest-penetration search is degenerate there (every
    direction around the axis is equally good) and the portal refinement stops within its 1e-6 tolerance at a point
    that depends on the last bit of the support points, so only the first step (identical inputs, 1e-13 agreement) is
    compared; the crossed pair (yaw = 1.1) is well conditioned and is followed for three steps."""
    mj, em = oracle.MjLike("flat"), E.Emul("flat")
    lay = [np.cos(np.pi / 4), np.sin(np.pi / 4), 0, 0]
    q = [0.0, 0.0, 0.0375] + lay + [dx, 0.4, 0.0375] + _quat_mul([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)], lay) \
        + [0.0, 0.8, 3.0, 1, 0, 0, 0]
    mj.reset_data(); mj.qpos[:] = q; mj.ctrl[:] = 0.15
    for st in range(nstep):
        em.rec[0:21] = mj.qpos; em.rec[21:39] = mj.qvel; em.rec[39:57] = mj.qacc_warmstart
        mj.step(1)
        ten, cfrc, stats = em.mj_step(np.full(6, 0.15), 1)
        assert stats[0] == mj.nefc // 6 and stats[4] == 0
        if st == 0:
            assert stats[0] == ncon0
        assert np.abs(em.qvel - mj.qvel).max() < 1e-10
        assert np.abs(em.warm - mj.qacc_warmstart).max() <= 1e-9 * np.abs(mj.qacc_warmstart).max()


def _lying_bar(x, y, z, yaw):
    """pose of a bar lying flat (axis horizontal), its axis turned by yaw about z"""
    lay = [np.cos(np.pi / 4), np.sin(np.pi / 4), 0, 0]
    return [x, y, z] + _quat_mul([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)], lay)


@pytest.mark.parametrize("scene", ["triangle", "two_on_one"])
def test_coupled_bars_block_factorisation(oracle, scene):
    """bars lying across one another: the Hessian blocks between bars exist, so the distributed block LDL^T runs its
    coupled stages.  triangle: every pair touches (all three blocks below the diagonal); two_on_one: bars 1 and 2 both
    cross bar 0 but not each other (blocks (1,0), (2,0) present, (2,1) appears as fill-in only)."""
    if scene == "triangle":
        q = []
        for b in range(3):
            phi = np.pi / 2 + 2 * np.pi / 3 * b
            q += _lying_bar(0.08 * np.cos(phi), 0.08 * np.sin(phi), 0.0375 + 0.02 * b, phi)
        want = {(0, 1), (0, 2), (1, 2)}
    else:
        q = _lying_bar(0, 0, 0.0375, np.pi / 2) + _lying_bar(-0.15, 0, 0.09, 0.0) + _lying_bar(0.15, 0, 0.09, 0.05)
        want = {(0, 1), (0, 2)}
    mj, em = oracle.MjLike("flat"), E.Emul("flat")
    mj.reset_data(); mj.qpos[:] = q; mj.ctrl[:] = -0.2
    for st in range(4):
        em.rec[0:21] = mj.qpos; em.rec[21:39] = mj.qvel; em.rec[39:57] = mj.qacc_warmstart
        mj.step(1)
        ten, cfrc, stats = em.mj_step(np.full(6, -0.2), 1)
        pairs = {((c.geom1 - 1) // 5, (c.geom2 - 1) // 5) for c in mj.contacts() if not c.exclude and c.dist < 0 and c.geom1 > 0}
        if st == 0:
            assert pairs == want
        assert stats[0] == mj.nefc // 6 and stats[4] == 0
        assert np.abs(em.qvel - mj.qvel).max() < 1e-10
        assert np.abs(em.warm - mj.qacc_warmstart).max() <= 1e-9 * np.abs(mj.qacc_warmstart).max()


@pytest.mark.parametrize("task", ["straight", "tracking", "aiming", "vel_track"])
def test_obs_noise_parity(task):
    """use_obs_noise (tr_env.py:142, 552-644): the CUDA source's noisy observation against the oracle's restatement on
    the same normal draws; the true observation (info["real_observation"]) and the reward are untouched by the noise."""
    rng = np.random.default_rng(3)
    oe = OracleEnv("flat", "tr_env", desired_action=task, use_obs_noise=True)
    em = E.Emul("flat", env_kind="tr_env", desired_action=task, use_obs_noise=True)
    em.set_noise(77, 5)
    draws = np.concatenate([rng.uniform(0, 1, 2), rng.standard_normal(6), rng.uniform(0, 1, 2)])
    o_true, o_em = oe.reset(draws), em.reset(draws, seed=77, env_id=5)
    assert np.abs(em.real_obs - o_true).max() < 1e-7
    assert np.abs(oe.noisy_obs(o_true, E.obs_normals(77, 5, 1, 0, 45)) - o_em).max() < 1e-7
    for st in range(1, 5):
        a = rng.uniform(-0.45, 0.15, 6)
        ob1, r1, _, _, _ = oe.step(a)
        ob2, r2, _, _ = em.step(a)
        assert np.abs(em.real_obs - ob1).max() < 1e-7 and abs(r1 - r2) <= 1e-7 * max(1.0, abs(r1))
        assert np.abs(oe.noisy_obs(ob1, E.obs_normals(77, 5, 1, st, 45)) - ob2).max() < 1e-7
        assert 0.01 < np.abs(ob2[:36] - ob1[:36]).max() < 0.5          # the noise really is there (stdev 0.05)


def test_obs_noise_draws_are_standard_normal_and_keyed():
    z = np.array([E.obs_normals(9, s, 1, 3, 45) for s in range(3000)])
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.06            # the two outputs of a Box-Muller pair
    a = E.obs_normals(9, 4, 1, 3, 45)
    assert np.array_equal(a, E.obs_normals(9, 4, 1, 3, 45))
    for other in (E.obs_normals(10, 4, 1, 3, 45), E.obs_normals(9, 5, 1, 3, 45), E.obs_normals(9, 4, 2, 3, 45), E.obs_normals(9, 4, 1, 4, 45)):
        assert not np.array_equal(a, other)


def test_obs_noise_is_a_tr_env_option():
    with pytest.raises(ValueError):
        E.Emul("flat", env_kind="tensegrity_env", use_obs_noise=True)


def test_reset_noise_and_contact_cost_parity():
    """reset_noise_scale (tr_env.py:734-743: qpos += U(-s, s), qvel = s N(0, 1), Philox-keyed here) and
    use_contact_forces (:292-304, 513-516: reward -= w sum(clip(cfrc_ext)^2), info["reward_ctrl"] = -contact_cost)
    against the oracle env fed the same draws."""
    rng = np.random.default_rng(21)
    kw = dict(desired_action="tracking", reset_noise_scale=0.02, use_contact_forces=True, contact_cost_weight=5e-4,
              contact_force_range=(-50.0, 50.0))
    oe = OracleEnv("flat", "tr_env", **kw)
    em = E.Emul("flat", env_kind="tr_env", **kw)
    draws = np.concatenate([rng.uniform(0, 1, 2), rng.standard_normal(6), rng.uniform(0, 1, 2)])
    nz = E.reset_noise(41, 6, 0)
    assert np.abs(nz[0]).max() <= 1 and nz[0].std() > 0.3 and abs(nz[1].mean()) < 0.8
    o1, o2 = oe.reset(draws, noise=nz), em.reset(draws, seed=41, env_id=6)
    assert np.abs(o1 - o2).max() < 1e-7
    # the noise moved the reset: the same draws without it give another observation
    assert np.abs(OracleEnv("flat", "tr_env", desired_action="tracking").reset(draws) - o1).max() > 1e-3
    costs = []
    for st in range(10):
        a = rng.uniform(-0.45, 0.15, 6)
        ob1, r1, t1, tr1, i1 = oe.step(a)
        ob2, r2, d2, info = em.step(a)
        assert np.abs(ob1 - ob2).max() < 1e-7 and abs(r1 - r2) <= 1e-7 * max(1.0, abs(r1))
        assert info[1] == pytest.approx(i1["reward_ctrl"], rel=1e-7, abs=1e-9)
        costs.append(-i1["reward_ctrl"])
    assert max(costs) > 0.1      # the contact cost is really in play (wrenches of order 100 N clipped at 50)
