"""CPU tests that pin the oracle (oracle/zb_oracle.c).

The reference has no tests, golden vectors or importable implementation for this path, so the
restatement is pinned by (1) analytic known answers, (2) cross-algorithm identities of rigid-body dynamics,
(3) constraint-solver properties, (4) an independent numpy transcription of the reference's own servo model
(/root/reference/train.py:1137-1196, 1242-1280), and (5) the committed restatement goldens in tests/golden/.
"""

import math
import os

import numpy as np
import pytest
import zb_oracle as zo
from helpers import quat_from_axis_angle, random_actions, random_states

import ksim_gym_zbot_b200  # noqa: F401
from ksim_gym_zbot_b200 import model as M

G = float(np.float32(9.81))  # the model stores gravity in fp32


def frictionless(m):
    f = m.copy()
    f.field("dof_frictionloss")[:] = 0
    f.field("dof_damping")[:] = 0
    f.field("jnt_limited")[:] = 0
    return f


def airborne(m, seed, n=1, vel=1.0):
    q, v = random_states(m, n, seed, joint_noise=0.3, vel=vel)
    q[:, 2] += 1.0
    q[:, 3:7] /= np.linalg.norm(q[:, 3:7], axis=1, keepdims=True)
    return q.astype(np.float64), v.astype(np.float64)


# ---------------------------------------------------------------------------------------------------
# (2) cross-algorithm identities
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_mass_matrix_crb_vs_jacobians(zmodel, seed):
    """qM from the composite-rigid-body pass == sum_b J_b^T I_b J_b built independently in model.py."""
    q, v = airborne(zmodel, seed)
    o = zo.Oracle(zmodel, M.make_opts(), np.float64)
    d = o.forward_one(q[0], v[0], ctrl=np.zeros(20))
    ref, _ = M._mass_matrix_np(zmodel, d.get("qpos"))
    qM = d.get("qM")
    assert np.abs(qM - ref).max() < 1e-8
    assert np.abs(qM - qM.T).max() == 0
    assert np.linalg.eigvalsh(qM).min() > 0
    L = d.get("qL")
    assert np.abs(L @ L.T - qM).max() < 1e-12
    # tree sparsity: 21 (base) + 2*57 (legs) + 2*34 (arms) = 203 structural non-zeros in the lower triangle
    anc = np.zeros((26, 26), bool)
    for i in range(26):
        j = i
        while j >= 0:
            anc[i, j] = True
            j = zmodel.dof_parent[j]
    assert anc.sum() == 203
    assert np.all(qM[~(anc | anc.T)] == 0)


def test_mass_matrix_from_rne_columns(zmodel):
    """M e_j == RNE(q, 0, e_j) - RNE(q, 0, 0): bias-force pass and CRB pass agree (via cfrc_int of rne_postconstraint)."""
    m = frictionless(zmodel)
    q, _ = airborne(m, 3)
    o = zo.Oracle(m, M.make_opts(gravity=(0.0, 0.0, 0.0)), np.float64)
    d0 = o.forward_one(q[0], np.zeros(26), ctrl=np.zeros(20))
    qM = d0.get("qM")
    # kinetic energy identity for random velocities: 0.5 v^T M v == 0.5 sum_b cvel_b . (cinert_b * cvel_b)
    rng = np.random.default_rng(0)
    for _ in range(3):
        v = rng.uniform(-1, 1, 26)
        d = o.forward_one(q[0], v, ctrl=np.zeros(20))
        cv, ci = d.get("cvel"), d.get("cinert")
        ke = 0.0
        for b in range(26):
            I = np.array([[ci[b, 0], ci[b, 3], ci[b, 4]], [ci[b, 3], ci[b, 1], ci[b, 5]], [ci[b, 4], ci[b, 5], ci[b, 2]]])
            h, mass = ci[b, 6:9], ci[b, 9]
            w, vl = cv[b, :3], cv[b, 3:]
            ke += 0.5 * (w @ I @ w + 2 * w @ np.cross(h, vl) + mass * vl @ vl)
        arm = 0.5 * np.sum(m.field("dof_armature") * v * v)
        assert abs(ke + arm - 0.5 * v @ qM @ v) < 1e-10


def test_gravity_bias_is_potential_gradient(zmodel):
    """qfrc_bias(q, 0) == dV/dq with V = sum_b m_b g z_b  (finite differences on the fp64 oracle)."""
    m = frictionless(zmodel)
    q, _ = airborne(m, 4)
    o = zo.Oracle(m, M.make_opts(), np.float64)

    def V(qq):
        d = o.forward_one(qq, np.zeros(26), ctrl=np.zeros(20))
        return float(np.sum(m.field("body_mass") * d.get("xipos")[:, 2]) * G)

    d = o.forward_one(q[0], np.zeros(26), ctrl=np.zeros(20))
    bias = d.get("qfrc_bias")
    eps = 1e-6
    for dof in [2, 6, 8, 9, 13, 19, 22, 25]:
        qa = dof + 1 if dof >= 6 else dof
        qp, qm_ = q[0].copy(), q[0].copy()
        qp[qa] += eps
        qm_[qa] -= eps
        assert abs((V(qp) - V(qm_)) / (2 * eps) - bias[dof]) < 1e-5, dof
    assert abs(bias[2] - M.total_mass(m) * G) < 1e-6 and abs(bias[0]) < 1e-12


def test_forward_inverse_dynamics_identity(zmodel):
    """No active constraints: qacc == qacc_smooth and M qacc + bias == passive + actuator."""
    m = zmodel.copy()
    m.field("dof_frictionloss")[:] = 0
    q, v = airborne(m, 5)
    o = zo.Oracle(m, M.make_opts(), np.float64)
    ctrl = np.random.default_rng(1).uniform(-1, 1, 20)
    d = o.forward_one(q[0], v[0], ctrl=ctrl)
    assert d.get("con_active").sum() == 0 and (d.get("efc_type")[26:52] == 0).all()
    np.testing.assert_array_equal(d.get("qacc"), d.get("qacc_smooth"))
    lhs = d.get("qM") @ d.get("qacc") + d.get("qfrc_bias")
    rhs = d.get("qfrc_passive") + d.get("qfrc_actuator")
    assert np.abs(lhs - rhs).max() < 1e-10
    # net external wrench on a free-flying tree is zero: root cfrc_int vanishes
    assert np.abs(d.get("cfrc_int")[1]).max() < 1e-9
    np.testing.assert_allclose(d.get("qfrc_passive"), -m.field("dof_damping") * v[0])


# ---------------------------------------------------------------------------------------------------
# (1) analytic known answers
# ---------------------------------------------------------------------------------------------------
def test_free_fall_momentum_known_answer(zmodel):
    """Airborne, no joint motion, rigid pose: the base follows v_{k+1} = v_k - g dt, z_{k+1} = z_k + v_{k+1} dt exactly."""
    m = frictionless(zmodel)
    m.field("dof_armature")[6:] = 1e6  # freeze the joints so the robot is one rigid body
    o = zo.Oracle(m, M.make_opts(dt=0.002, integrator=M.INT_EULER), np.float64)
    st = o.new_state(1)
    q0 = M.default_qpos(m)[None].copy()
    q0[0, 2] += 2.0
    o.reset(st, q0, np.zeros((1, 26)), np.array([1e9]))
    st["act_pos"][...] = q0[:, 7:]
    z, vz = q0[0, 2], 0.0
    for k in range(100):
        o.step(st, q0[:, 7:], 1)
        vz -= G * 0.002
        z += vz * 0.002
        assert abs(st["qvel"][0, 2] - vz) < 1e-7 and abs(st["qpos"][0, 2] - z) < 1e-7
    assert abs(st["time"][0] - 0.2) < 1e-7
    assert abs(np.linalg.norm(st["qpos"][0, 3:7]) - 1) < 1e-12


def test_energy_drift_first_order(zmodel):
    """Conservative free flight (no damping/friction/limits/actuation): energy error shrinks ~linearly with dt."""
    m = frictionless(zmodel)
    q, v = airborne(m, 6, vel=2.0)

    def drift(dt, steps):
        o = zo.Oracle(m, M.make_opts(dt=dt, integrator=M.INT_EULER, eulerdamp=0), np.float64)
        d = o.Data()
        d.set("qpos", q[0]); d.set("qvel", v[0])
        es = []
        for _ in range(steps):
            o.step_one(d)
            f = o.forward_one(d.get("qpos"), d.get("qvel"), ctrl=np.zeros(20))
            vv = f.get("qvel")
            es.append(0.5 * vv @ f.get("qM") @ vv + np.sum(m.field("body_mass") * f.get("xipos")[:, 2]) * G)
        return abs(es[-1] - es[0]) / abs(es[0])

    d1, d2 = drift(2e-4, 250), drift(1e-4, 500)
    assert d1 < 5e-3 and d2 < 0.7 * d1, (d1, d2)


def test_quaternion_integration_identity(zmodel):
    """free-joint position update: q' = normalize(q * exp(omega' dt / 2)) with the NEW angular velocity (A.10)."""
    q, v = airborne(zmodel, 7)
    o = zo.Oracle(zmodel, M.make_opts(dt=0.004), np.float64)
    d = o.Data()
    d.set("qpos", q[0]); d.set("qvel", v[0])
    o.step_one(d)
    w = d.get("qvel")[3:6]
    qa = quat_from_axis_angle(w, np.linalg.norm(w) * 0.004)
    exp = M.quat_mul(q[0, 3:7], qa)
    np.testing.assert_allclose(d.get("qpos")[3:7], exp / np.linalg.norm(exp), atol=1e-9)
    np.testing.assert_allclose(d.get("qpos")[:3], q[0, :3] + 0.004 * d.get("qvel")[:3], atol=1e-14)
    np.testing.assert_allclose(d.get("qpos")[7:], q[0, 7:] + 0.004 * d.get("qvel")[6:], atol=1e-14)
    np.testing.assert_array_equal(d.get("qacc_warmstart"), d.get("qacc"))


# ---------------------------------------------------------------------------------------------------
# (3) constraints / solver
# ---------------------------------------------------------------------------------------------------
def test_standing_contact_force_balances_weight(zmodel):
    o = zo.Oracle(zmodel, M.make_opts(dt=0.002), np.float64)
    st, out = o.new_state(1), o.new_out(1)
    o.reset(st, M.default_qpos(zmodel)[None], np.zeros((1, 26)), np.array([1e9]))
    for _ in range(60):
        o.step(st, M.BIAS_POSE[None], 10, out=out)
    sd = out["sensordata"][0]
    weight = M.total_mass(zmodel) * G
    assert abs(sd[10] + sd[11] - weight) / weight < 0.02  # touch L + touch R
    assert abs(sd[2 + 4] - G) < 0.2  # accelerometer z reads +g at rest
    assert np.abs(sd[7:10]).max() < 0.05  # gyro ~ 0
    assert (out["efc_force"][0, 52:] >= 0).all()
    assert 0.05 < st["qpos"][0, 2] < 0.5
    # force sensors: |z| component carries the foot load (sign = cfrc_int convention)
    assert abs(abs(sd[14]) + abs(sd[17]) - weight) / weight < 0.1


def test_cg_and_newton_agree_at_convergence(zmodel):
    q, v = random_states(zmodel, 6, 8)
    for i in range(6):
        res = {}
        for solver in (M.SOLVER_CG, M.SOLVER_NEWTON):
            o = zo.Oracle(zmodel, M.make_opts(solver=solver, iterations=200, ls_iterations=50, tolerance=1e-14), np.float64)
            res[solver] = o.forward_one(q[i], v[i], ctrl=np.zeros(20))
        a, b = res[0].get("qacc"), res[1].get("qacc")
        assert np.abs(a - b).max() < 1e-5 * max(1.0, np.abs(b).max()), i
        assert abs(res[0].solver_cost - res[1].solver_cost) < 1e-7 * max(1.0, abs(res[1].solver_cost))
        # at the optimum the gradient vanishes: M a - qfrc_smooth - J^T f == 0
        d = res[1]
        grad = d.get("qM") @ d.get("qacc") - d.get("qfrc_smooth") - d.get("qfrc_constraint")
        assert np.abs(grad).max() < 1e-6
        np.testing.assert_allclose(d.get("efc_J").T @ d.get("efc_force"), d.get("qfrc_constraint"), atol=1e-10)


def test_force_law_and_cost_monotone(zmodel):
    """efc_force == -D (J qacc - aref) on the active rows (Huber-clamped on friction rows) at ANY solver output, and the
    solver cost never increases from one iteration to the next (trace[4 i + 1])."""
    q, v = random_states(zmodel, 8, 20)
    for solver in (M.SOLVER_CG, M.SOLVER_NEWTON):
        o = zo.Oracle(zmodel, M.make_opts(solver=solver), np.float64)
        for i in range(8):
            d = o.forward_one(q[i], v[i], ctrl=np.zeros(20))
            jar = d.get("efc_J") @ d.get("qacc") - d.get("efc_aref")
            D, f, t, fl, R = d.get("efc_D"), d.get("efc_force"), d.get("efc_type"), d.get("efc_floss"), d.get("efc_R")
            for r in range(len(t)):
                if t[r] == 0:
                    assert f[r] == 0
                elif t[r] == 1:
                    exp = np.clip(-D[r] * jar[r], -fl[r], fl[r])
                    assert abs(f[r] - exp) < 1e-9 * max(1, abs(exp)), (r, f[r], exp)
                    assert abs(R[r] * D[r] - 1) < 1e-12
                else:
                    exp = -D[r] * jar[r] if jar[r] < 0 else 0.0
                    assert abs(f[r] - exp) < 1e-9 * max(1, abs(exp)), (r, f[r], exp)
            costs = d.get("trace")[1:4 * d.solver_niter:4]
            assert np.all(np.diff(costs) <= 1e-9 * np.abs(costs[:-1]) + 1e-12), costs
            assert d.solver_niter <= 8


def test_constraint_rows(zmodel):
    """Row bookkeeping: friction rows on the 20 hinge dofs, limits only when violated, 4 pyramid rows per active contact."""
    m = zmodel
    q, v = random_states(m, 1, 9)
    q[0, 7 + 3] = m.field("jnt_range")[6 + 3][0] - 0.05  # right knee below its lower limit
    q[0, 7 + 13] = m.field("jnt_range")[6 + 13][1] + 0.02  # left shoulder roll above its upper limit
    o = zo.Oracle(m, M.make_opts(), np.float64)
    d = o.forward_one(q[0], v[0], ctrl=np.zeros(20))
    t = d.get("efc_type")
    assert (t[:6] == 0).all() and (t[6:26] == 1).all()
    lim = np.where(t[26:52] == 2)[0]
    assert set(lim) == {9, 19}
    J = d.get("efc_J")
    assert J[26 + 9, 9] == 1 and J[26 + 19, 19] == -1
    assert abs(d.get("efc_pos")[26 + 9] + 0.05) < 1e-6 and abs(d.get("efc_pos")[26 + 19] + 0.02) < 1e-6
    act = d.get("con_active")
    for c in range(8):
        assert ((t[52 + 4 * c: 56 + 4 * c] == 3).all()) == bool(act[c])
    # contact Jacobian rows == finite-difference of the contact point velocity along n +- mu t
    c = int(np.argmax(act))
    if act[c]:
        mu, fr = d.get("con_mu")[c], d.get("con_frame")[c].reshape(3, 3)
        r0 = J[52 + 4 * c]
        body = d.get("con_body")[c]
        cd, off = d.get("cdof"), d.get("con_pos")[c] - d.get("subtree_com")[1]
        jp = np.zeros((3, 26))
        dof = m.body_dofadr[body - 1] if m.body_jnttype[body] == 0 else m.body_dofadr[body]
        j = dof
        while j >= 0:
            jp[:, j] = np.cross(cd[j, :3], off) + cd[j, 3:]
            j = m.dof_parent[j]
        np.testing.assert_allclose(r0, (fr[0] + mu * fr[1]) @ jp, atol=1e-12)
    # impedance / regularisation: R = invweight (1 - imp) / imp with imp = d0 at pos = 0 (friction rows)
    R = d.get("efc_R")[6:26]
    np.testing.assert_allclose(R, m.field("dof_invweight0")[6:] * (1 - 0.9) / 0.9, rtol=1e-6)


def test_plane_box_manifold(zmodel):
    """plane_convex: deepest vertices become slots, duplicates are parked at dist = 1, pos = vertex - dist/2 n."""
    o = zo.Oracle(zmodel, M.make_opts(), np.float64)
    q = M.default_qpos(zmodel, penetration=2e-3)
    q[3:7] = quat_from_axis_angle([1, 0.3, 0], 0.02)  # tilt so that not all foot vertices are within 1 mm
    d = o.forward_one(q, np.zeros(26), ctrl=np.zeros(20))
    dist, pos, vert = d.get("con_dist"), d.get("con_pos"), d.get("con_vert")
    for g in range(2):
        sl = slice(4 * g, 4 * g + 4)
        uniq = [i for i in range(4) if vert[sl][i] not in vert[sl][:i]]
        for i in range(4):
            if i in uniq:
                assert dist[sl][i] < 1
            else:
                assert dist[sl][i] == 1.0
        # bottom vertices only (z index bit 0 == 0 -> -size_z)
        assert all(int(vv) % 2 == 0 for vv in vert[sl][uniq])
        bx, bm = d.get("box_xpos")[g], d.get("box_xmat")[g].reshape(3, 3)
        sz = zmodel.field("box_size")[g]
        for i in uniq:
            vv = int(vert[sl][i])
            local = np.array([1 if vv & 4 else -1, 1 if vv & 2 else -1, 1 if vv & 1 else -1]) * sz
            w = bx + bm @ local
            assert abs(w[2] - dist[sl][i]) < 1e-12
            np.testing.assert_allclose(pos[sl][i], w - 0.5 * dist[sl][i] * np.array([0, 0, 1.0]), atol=1e-12)
    np.testing.assert_allclose(d.get("con_frame")[0].reshape(3, 3), [[0, 0, 1], [0, 1, 0], [-1, 0, 0]], atol=1e-15)


# ---------------------------------------------------------------------------------------------------
# (4) the reference's own servo model, transcribed independently from train.py
# ---------------------------------------------------------------------------------------------------
def _feetech_numpy(m, action, qpos_j, qvel_j, pos, vel, dt):
    """numpy transcription of trapezoidal_step + get_stateful_ctrl (/root/reference/train.py:1146-1196, 1252-1269)."""
    vmax, amax = m.field("servo_vmax").astype(np.float64), m.field("servo_amax").astype(np.float64)
    err = action - pos
    thr = np.where(err >= 0, m.servo_deadband_pos, m.servo_deadband_neg)
    in_db = np.abs(err) <= thr
    db_v = vel * 0.8
    db_p = pos + db_v * dt
    tdir = np.sign(err)
    stop = np.abs(vel ** 2) / (2 * amax)
    vdir = np.sign(vel)
    towards = vdir * tdir >= 0
    accel = np.logical_and(towards, np.abs(err) > stop)
    acc = np.where(accel, tdir * amax, -vdir * amax)
    acc = np.where(np.abs(vel) < 1e-6, tdir * amax, acc)
    pv = np.clip(vel + acc * dt, -vmax, vmax)
    pp = pos + pv * dt
    nv, np_ = np.where(in_db, db_v, pv), np.where(in_db, db_p, pp)
    f = lambda name: m.field(name).astype(np.float64)  # noqa: E731
    duty = f("servo_kp") * f("servo_error_gain") * (np_ - qpos_j) + f("servo_kd") * (nv - qvel_j)
    duty = np.clip(duty, -f("servo_max_pwm"), f("servo_max_pwm"))
    return np_, nv, duty * f("servo_vin") * f("servo_kt") / f("servo_R")


def test_feetech_actuator_matches_reference_transcription(zmodel):
    m = zmodel
    o = zo.Oracle(m, M.make_opts(dt=0.001), np.float64)
    rng = np.random.default_rng(10)
    n = 64
    q, v = random_states(m, n, 11)
    q, v = q.astype(np.float64), v.astype(np.float64)
    st, out = o.new_state(n), o.new_out(n)
    o.reset(st, q, v, np.full(n, 1e9))
    # exercise every branch: deadband, accelerate, decelerate (overshoot), zero velocity
    st["act_vel"][...] = rng.choice([0.0, 1e-7, 0.5, -0.5, 4.9, -3.0], size=(n, 20))
    act = st["act_pos"] + rng.choice([0.0, 0.002, -0.002, 0.004, 0.3, -0.3, 1.0], size=(n, 20))
    st["act_pos"][::2] += rng.uniform(-0.6, 0.6, (n // 2, 20))  # planner far from the joint -> duty saturates at max_pwm
    p0, v0 = st["act_pos"].copy(), st["act_vel"].copy()
    qn = st["qpos"].copy()
    o.step(st, act, 1, out=out)
    ep, ev, et = _feetech_numpy(m, act, qn[:, 7:], v[:, 6:], p0, v0, 0.001)
    np.testing.assert_allclose(st["act_pos"], ep, rtol=0, atol=1e-9)
    np.testing.assert_allclose(st["act_vel"], ev, rtol=0, atol=1e-9)
    np.testing.assert_allclose(st["act_torque"], et, rtol=0, atol=1e-8)
    np.testing.assert_allclose(out["ctrl"], et, rtol=0, atol=1e-8)
    lim = m.field("act_forcerange")[:, 1]
    np.testing.assert_allclose(out["actuator_force"], np.clip(et, -lim, lim), rtol=0, atol=1e-8)
    assert (np.abs(et) >= lim - 1e-9).any() and (np.abs(et) < lim).any()  # both the clipped and the linear regime


def test_push_event_and_reset_semantics(zmodel):
    m = zmodel
    o = zo.Oracle(m, M.make_opts(dt=0.005), np.float64)
    n = 4
    q, v = random_states(m, n, 12)
    st = o.new_state(n)
    o.reset(st, q, v, np.array([0.012, 0.004, 5.0, 0.019]))
    np.testing.assert_array_equal(st["act_pos"], st["qpos"][:, 7:])
    np.testing.assert_array_equal(st["act_vel"], st["qvel"][:, 6:])
    assert not st["act_torque"].any() and not st["time"].any() and not st["qacc_warmstart"].any()
    pr = np.zeros((n, 7))
    pr[:, 0] = 10.0
    pr[:, 6] = 3.0
    ref = o.new_state(n)
    for k in st:
        ref[k][...] = st[k]
    o.step(st, st["qpos"][:, 7:].copy(), 4, push_rand=pr)
    o.step(ref, ref["qpos"][:, 7:].copy(), 4)
    # timers 0.012/0.004/0.019 expire within 4 x 5 ms (t <= 0 after the decrement), 5.0 does not
    kicked = np.abs(st["qvel"][:, 0] - ref["qvel"][:, 0]) > 1.0
    np.testing.assert_array_equal(kicked, [True, True, False, True])
    np.testing.assert_allclose(st["push_t_rem"], [3.0 - 0.005, 3.0 - 0.015, 5.0 - 0.02, 3.0], atol=1e-12)
    np.testing.assert_allclose(st["time"], 0.02)
    # partial reset
    done = np.array([1, 0, 0, 1], np.uint8)
    before = {k: a.copy() for k, a in st.items()}
    q2, v2 = random_states(m, n, 13)
    o.reset(st, q2, v2, np.full(n, 2.5), done=done)
    for k in st:
        np.testing.assert_array_equal(st[k][1:3], before[k][1:3])
    np.testing.assert_array_equal(st["time"][[0, 3]], 0)
    np.testing.assert_array_equal(st["push_t_rem"][[0, 3]], 2.5)


def test_action_latency_semantics(zmodel):
    """MjxEngine.step latency select: substeps s < latency_steps track prev_action, later ones the new one."""
    n, S = 5, 4
    q, v = random_states(zmodel, n, 30)
    a_new, a_old = random_actions(n, 31), random_actions(n, 32)
    o = zo.Oracle(zmodel, M.make_opts(dt=0.005), np.float64)

    def run(action, params):
        st = o.new_state(n)
        o.reset(st, q, v, np.full(n, 1e9))
        o.step(st, action, S, params=params)
        return st

    base_new, base_old = run(a_new, None), run(a_old, None)
    lat0 = run(a_new, {"prev_action": a_old, "latency_steps": np.zeros(n, np.int32)})
    latS = run(a_new, {"prev_action": a_old, "latency_steps": np.full(n, S, np.int32)})
    mixed = run(a_new, {"prev_action": a_old, "latency_steps": np.array([0, 1, 2, 3, 9], np.int32)})
    for k in ("qpos", "qvel", "act_pos", "act_torque"):
        np.testing.assert_array_equal(lat0[k], base_new[k])
        np.testing.assert_array_equal(latS[k], base_old[k])
        np.testing.assert_array_equal(mixed[k][0], base_new[k][0])
        np.testing.assert_array_equal(mixed[k][4], base_old[k][4])
    assert np.abs(mixed["act_pos"][2] - base_new["act_pos"][2]).max() > 1e-6
    assert np.abs(mixed["act_pos"][2] - base_old["act_pos"][2]).max() > 1e-6


def test_env_param_overrides(zmodel):
    """Randomizer fields (train.py:1439-1455) reach the dynamics: heavier body -> larger M, floor mu -> contact rows."""
    m = zmodel
    n = 3
    q, v = random_states(m, n, 14)
    p = M.randomize_env_params(m, n, 3)
    o = zo.Oracle(m, M.make_opts(), np.float64)
    base = o.forward_one(q[1], v[1], ctrl=np.zeros(20))
    d = o.forward_one(q[1], v[1], ctrl=np.zeros(20), params=p, env=1)
    scale = p["body_mass"][1, 1] / m.field("body_mass")[1]
    np.testing.assert_allclose(d.get("cinert")[:, 9], m.field("body_mass") * scale, rtol=1e-6)
    assert np.all(d.get("con_mu") == np.float64(p["floor_friction"][1]))
    assert not np.allclose(d.get("qM"), base.get("qM"))
    np.testing.assert_allclose(np.diag(d.get("qM"))[6:] - np.diag(M._mass_matrix_np(m, d.get("qpos"))[0])[6:] + m.field("dof_armature")[6:],
                               p["dof_armature"][1, 6:], atol=5e-3)
    np.testing.assert_allclose(d.get("efc_floss")[6:26], p["dof_frictionloss"][1, 6:], rtol=1e-6)
    # qpos0 offset shifts the hinge reference angle
    p2 = {"qpos0": np.repeat(m.field("qpos0")[None], n, 0).copy()}
    p2["qpos0"][:, 7 + 2] += 0.1
    q2 = q.copy()
    q2[:, 7 + 2] += 0.1
    a = o.forward_one(q2[0], v[0], ctrl=np.zeros(20), params=p2, env=0)
    b = o.forward_one(q[0], v[0], ctrl=np.zeros(20))
    np.testing.assert_allclose(a.get("xpos"), b.get("xpos"), atol=1e-6)


# ---------------------------------------------------------------------------------------------------
# (5) goldens + precision
# ---------------------------------------------------------------------------------------------------
def test_golden_trace_and_forward_cases(zmodel):
    here = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(here, "zbot_syn_1000.npz"))
    for name, dt, tol in (("f64", np.float64, 1e-9), ("f32", np.float32, 5e-4)):
        o = zo.Oracle(zmodel, M.make_opts(dt=0.001), dt)
        st = o.new_state(1)
        o.reset(st, M.default_qpos(zmodel)[None].astype(np.float32), np.zeros((1, 26)), np.array([1e9]))
        for t in range(50):
            o.step(st, M.BIAS_POSE[None].astype(np.float32), 20)
            assert np.abs(st["qpos"][0] - g["qpos_" + name][t]).max() < tol, (name, t)
    # robot keeps standing for the whole second
    assert 0.25 < g["qpos_f64"][-1, 2] < 0.30 and np.abs(g["qvel_f64"][-1]).max() < 0.05
    f = np.load(os.path.join(here, "zbot_syn_forward16.npz"))
    params = {k[2:]: f[k] for k in f.files if k.startswith("p_")}
    o = zo.Oracle(zmodel, M.make_opts(dt=0.005), np.float64)
    n = f["in_qpos"].shape[0]
    st, out = o.new_state(n), o.new_out(n)
    o.reset(st, f["in_qpos"], f["in_qvel"], np.full(n, 1e9), params=params, out=out)
    for k in ("xpos", "xquat", "cinert", "cvel", "qacc", "sensordata", "efc_force"):
        np.testing.assert_allclose(out[k], f["fwd_" + k], rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(out["contact_mask"], f["fwd_contact_mask"])
    o.step(st, f["in_action"], 4, params=params, out=out)
    np.testing.assert_allclose(st["qpos"], f["step_qpos"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(st["qvel"], f["step_qvel"], rtol=1e-8, atol=1e-8)


def test_fp32_oracle_tracks_fp64(zmodel):
    """The fp32 oracle (JAX's dtype) stays within the north_star tolerance of the fp64 one for a single step."""
    n = 128
    q, v = random_states(zmodel, n, 15)
    act = random_actions(n, 16)
    res = {}
    for dt in (np.float32, np.float64):
        o = zo.Oracle(zmodel, M.make_opts(dt=0.001), dt, threads=4)
        st, out = o.new_state(n), o.new_out(n)
        o.reset(st, q, v, np.full(n, 1e9))
        o.step(st, act, 1, out=out)
        res[dt] = (st, out)
    a, b = res[np.float32], res[np.float64]
    # median at rounding level; the tail is the iteration-capped solver's sensitivity (see test_parity_gpu._assert_close)
    eq = np.abs(a[0]["qpos"] - b[0]["qpos"]).max(axis=1)
    ev = (np.abs(a[0]["qvel"] - b[0]["qvel"]) / np.maximum(1, np.abs(b[0]["qvel"]))).max(axis=1)
    assert np.median(eq) < 1e-6 and eq.max() < 1e-3, (np.median(eq), eq.max())
    assert np.median(ev) < 2e-4 and ev.max() < 0.2, (np.median(ev), ev.max())
    assert (a[1]["contact_mask"] == b[1]["contact_mask"]).mean() > 0.98


def test_flop_counters(zmodel):
    """algorithmic FLOPs per env-step come from the oracle's counters (count build)."""
    o = zo.Oracle(zmodel, M.make_opts(dt=0.005), np.float32, count=True)
    n = 8
    q, v = random_states(zmodel, n, 17)
    st = o.new_state(n)
    o.reset(st, q, v, np.full(n, 1e9))
    o.flops_reset()
    o.step(st, random_actions(n, 18), 4)
    fl = o.flops()
    per_step = sum(fl.values()) / (n * 4)
    assert 2e4 < per_step < 1e6, per_step
    assert fl["solver"] > 0 and fl["kinematics"] > 0 and fl["rne"] > 0
