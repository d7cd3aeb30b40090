"""Oracle self-checks: analytic statics, Jacobians by finite differences, solver optimality (KKT),
MPR against closed forms.  These pin the restatement to physics, not to a MuJoCo binary (none exists here)."""
import numpy as np
import pytest


def rand_state(mj, rng, vel=0.5):
    mj.reset_data()
    q = np.array(mj.md["qpos0"])
    q += rng.normal(0, 0.02, 21)
    mj.qpos[:] = q
    mj.qvel[:] = rng.normal(0, vel, 18)
    mj.ctrl[:] = rng.uniform(-0.45, 0.15, 6)


def test_struct_sizes(oracle):
    oracle.lib()


def test_static_weight_carried_by_floor(oracle):
    mj = oracle.MjLike("flat")
    mj.ctrl[:] = 0.1
    for _ in range(150):
        mj.step(20)
    mj.rne_post_constraint()
    assert abs(mj.qvel).max() < 1e-2
    # world row: minus the summed floor reaction = -(total weight) ; 12 kg * 9.81
    assert mj.cfrc_ext[0, 5] == pytest.approx(-117.72, abs=1e-2)
    assert np.allclose(mj.cfrc_ext[1:, 5].sum(), 117.72, atol=1e-2)


def test_free_fall_no_contact(oracle):
    """With the full (cross-bar) damping derivative, internal tendon forces cannot move the COM:
    semi-implicit Euler free fall z = z0 - g h^2 n(n+1)/2.  The believed-2.3.7 default keeps the
    derivative on per-bar blocks only, which acts as a small spurious drag."""
    expect = -9.81 * 1e-6 * 100 * 101 / 2
    drops = {}
    for flags in (1 | 2 | 4, 1 | 4):
        mj = oracle.MjLike("flat", flags=flags)
        mj.reset_data()
        mj.qpos[2::7] += 5.0
        z0 = mj.qpos[2::7].mean()
        mj.ctrl[:] = -0.2
        mj.step(100)
        drops[flags] = mj.qpos[2::7].mean() - z0
    assert drops[7] == pytest.approx(expect, abs=1e-10)
    assert expect < drops[5] < 0.9 * expect


def test_tendon_jacobian_finite_difference(oracle):
    rng = np.random.default_rng(0)
    mj = oracle.MjLike("flat")
    rand_state(mj, rng)
    mj.forward()
    J = mj.ten_J.copy()
    L0 = mj.ten_length.copy()
    q0 = mj.qpos.copy()
    eps = 1e-6
    for k in range(18):
        b, j = divmod(k, 6)
        q = q0.copy()
        if j < 3:
            q[7 * b + j] += eps
        else:  # local-frame rotation: q <- q * exp(eps e_j / 2)
            w = np.zeros(3); w[j - 3] = eps
            a = q[7 * b + 3:7 * b + 7]
            dq = np.array([1.0, *(0.5 * w)])
            q[7 * b + 3:7 * b + 7] = oracle.M.quat_mul(a, dq)
        mj.qpos[:] = q
        mj.forward()
        fd = (mj.ten_length - L0) / eps
        assert np.allclose(fd, J[:, k], atol=2e-5), k


def test_contact_jacobian_matches_point_velocity(oracle):
    rng = np.random.default_rng(1)
    mj = oracle.MjLike("flat")
    rand_state(mj, rng)
    mj.qpos[2::7] -= 0.03
    mj.forward()
    assert mj.ncon >= 3
    for c in mj.contacts():
        if c.efc_address < 0:
            continue
        a0 = c.efc_address
        frame = np.array(c.frame).reshape(3, 3)
        pos = np.array(c.pos)
        vel = np.zeros(6)
        for body, s in ((c.body1, -1.0), (c.body2, 1.0)):
            if body == 0:
                continue
            b = body - 1
            R = mj.xmat[b].reshape(3, 3)
            v = mj.qvel[6 * b:6 * b + 3]
            w = R @ mj.qvel[6 * b + 3:6 * b + 6]
            vp = v + np.cross(w, pos - mj.xpos[b])
            vel[:3] += s * frame @ vp
            vel[3:] += s * frame @ w
        assert np.allclose(mj.efc_vel[a0:a0 + 6], vel, atol=1e-12)


@pytest.mark.parametrize("model", ["flat", "uneven"])
def test_newton_solution_is_stationary(oracle, model):
    """Solver-independent optimality: gradient of the convex primal cost vanishes at qacc,
    and no random perturbation lowers the cost."""
    rng = np.random.default_rng(2)
    mj = oracle.MjLike(model)
    mj.ctrl[:] = rng.uniform(-0.45, 0.0, 6)
    worst = 0.0
    for it in range(60):
        mj.ctrl[:] = rng.uniform(-0.45, 0.15, 6)
        mj.step(20)
        mj.forward()
        if mj.nefc == 0:
            continue
        c0, g = mj.primal_cost(mj.qacc, want_grad=True)
        scale = 1.0 / (mj.md["meaninertia"] * 18)
        worst = max(worst, scale * np.linalg.norm(g))
        for _ in range(5):
            dq = rng.normal(0, 1e-3, 18)
            assert mj.primal_cost(mj.qacc + dq) >= c0 - 1e-9 * max(1.0, abs(c0))
    assert worst < 1e-6, worst


def test_cone_force_feasible(oracle):
    rng = np.random.default_rng(3)
    mj = oracle.MjLike("flat")
    for it in range(40):
        mj.ctrl[:] = rng.uniform(-0.45, 0.15, 6)
        mj.step(20)
        mj.forward()
        fr = np.array(mj.md["friction"])
        for c in mj.contacts():
            if c.efc_address < 0:
                continue
            f = mj.efc_force[c.efc_address:c.efc_address + 6]
            assert f[0] >= -1e-12
            # dual (friction) cone: ||f_j / mu_j|| <= f_n   (mu = 1, impratio 1)
            assert np.linalg.norm(f[1:] / fr) <= f[0] * (1 + 1e-9) + 1e-9


def test_mpr_sphere_sphere_closed_form(oracle):
    I = np.eye(3).reshape(-1)
    hit, depth, dr, pos = oracle.mpr(2, [0, 0, 0], I, [0.5, 0, 0], 2, [0.8, 0.1, 0], I, [0.4, 0, 0])
    dvec = np.array([0.8, 0.1, 0]); dist = np.linalg.norm(dvec)
    assert hit and depth == pytest.approx(0.9 - dist, abs=2e-6)
    assert np.allclose(dr, dvec / dist, atol=1e-3)
    hit, *_ = oracle.mpr(2, [0, 0, 0], I, [0.5, 0, 0], 2, [1.0, 0, 0], I, [0.4, 0, 0])
    assert not hit


def test_mpr_crossed_cylinders(oracle):
    # two long thin cylinders crossing at right angles, axis distance 0.06, radii 0.0381
    I = np.eye(3).reshape(-1)
    Rx = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], float).reshape(-1)  # z-axis -> -y... axis along y
    hit, depth, dr, pos = oracle.mpr(5, [0, 0, 0], I, [0.0381, 0.688, 0], 5, [0.06, 0, 0], Rx, [0.0381, 0.688, 0])
    assert hit and depth == pytest.approx(2 * 0.0381 - 0.06, abs=5e-6)
    assert np.allclose(dr, [1, 0, 0], atol=1e-3)
    assert np.allclose(pos, [0.03, 0, 0], atol=1e-3)
    hit, *_ = oracle.mpr(5, [0, 0, 0], I, [0.0381, 0.688, 0], 5, [0.08, 0, 0], Rx, [0.0381, 0.688, 0])
    assert not hit


def test_quaternions_stay_normalised_and_energy_bounded(oracle):
    mj = oracle.MjLike("flat")
    rng = np.random.default_rng(4)
    for _ in range(100):
        mj.ctrl[:] = rng.uniform(-0.45, 0.15, 6)
        mj.step(20)
    q = mj.qpos.reshape(3, 7)[:, 3:]
    assert np.allclose(np.linalg.norm(q, axis=1), 1, atol=1e-9)
    assert np.isfinite(mj.qvel).all() and abs(mj.qvel).max() < 50
    assert mj.data.warning == 0


def test_heightfield_flat_region_supports_weight(oracle):
    mj = oracle.MjLike("uneven")
    mj.ctrl[:] = 0.0
    for _ in range(150):
        mj.step(20)
    mj.rne_post_constraint()
    assert np.isfinite(mj.qpos).all()
    assert mj.qpos[2::7].min() > -1.2  # resting on the terrain, not fallen through
    assert abs(mj.qvel).max() < 0.5
    assert mj.cfrc_ext[0, 5] == pytest.approx(-117.72, rel=0.2)


def _rotate_state(q, v, w, phi, shift):
    """the same physical state seen from a frame rotated by phi about z and shifted in the plane: free-joint positions
    and linear velocities rotate, quaternions get the z-rotation on the left, body-frame angular velocities are unchanged"""
    c, s = np.cos(phi), np.sin(phi)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    qz = np.array([np.cos(phi / 2), 0, 0, np.sin(phi / 2)])
    q2, v2, w2 = q.copy(), v.copy(), w.copy()
    for b in range(3):
        q2[7 * b:7 * b + 3] = R @ q[7 * b:7 * b + 3] + np.array([shift[0], shift[1], 0.0])
        a, bq = qz, q[7 * b + 3:7 * b + 7]
        q2[7 * b + 3:7 * b + 7] = [a[0] * bq[0] - a[3] * bq[3], a[0] * bq[1] - a[3] * bq[2], a[0] * bq[2] + a[3] * bq[1], a[0] * bq[3] + a[3] * bq[0]]
        v2[6 * b:6 * b + 3] = R @ v[6 * b:6 * b + 3]
        w2[6 * b:6 * b + 3] = R @ w[6 * b:6 * b + 3]
    return q2, v2, w2


def test_flat_floor_dynamics_are_equivariant_under_planar_motions(oracle):
    """Physics pin that needs no MuJoCo binary: on the infinite flat floor the step commutes with rotations about z
    and translations in the plane (contacts, friction cones, tendons, gravity all respect that symmetry).  Any frame
    mix-up in contact Jacobians, cone bases, body-frame angular velocities or the quaternion integrator breaks it."""
    rng = np.random.default_rng(4)
    a, b = oracle.MjLike("flat"), oracle.MjLike("flat")
    a.reset_data(); a.ctrl[:] = -0.3
    worst = 0.0
    for st in range(40):
        ctrl = rng.uniform(-0.45, -0.15, 6)
        phi, shift = rng.uniform(-np.pi, np.pi), rng.uniform(-3, 3, 2)
        q2, v2, w2 = _rotate_state(a.qpos, a.qvel, a.qacc_warmstart, phi, shift)
        b.reset_data(); b.qpos[:] = q2; b.qvel[:] = v2; b.qacc_warmstart[:] = w2; b.act[:] = a.act
        a.ctrl[:] = ctrl; b.ctrl[:] = ctrl
        a.step(5); b.step(5)
        qe, ve, _ = _rotate_state(a.qpos, a.qvel, a.qacc_warmstart, phi, shift)
        worst = max(worst, np.abs(b.qpos - qe).max(), np.abs(b.qvel - ve).max() / max(1.0, np.abs(ve).max()))
        assert np.abs(b.ten_length - a.ten_length).max() < 1e-9
    assert worst < 1e-8, worst
