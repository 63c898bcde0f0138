"""Collision groups beyond the two foot boxes: plane-sphere, plane-capsule and the
body-body primitives sphere-sphere / sphere-capsule / capsule-capsule (mjx collision_primitive [R]) with the two-body contact
Jacobian jac(body 2) - jac(body 1).

CPU part: the oracle's primitives against analytic answers, a brute-force closest-point search and a finite-difference check of the
two-body Jacobian rows.  GPU part: for every group configuration the kernel against the oracle — contact distance / position /
active set from forward(), then one full physics step (statistically, like every other full-solve comparison)."""
import numpy as np
import pytest

from helpers import random_actions, random_states

import ksim_gym_zbot_b200  # noqa: F401
from ksim_gym_zbot_b200 import model as M

QY90 = M.axis_angle_quat(np.array([0.0, 1.0, 0.0]), np.pi / 2)  # capsule axis (local z) along the foot's x


def variants(zmodel):
    """name -> (model, z offset that puts the new foot geoms about where the boxes' bottoms were)."""
    out = {}
    m = zmodel.copy()
    for g, foot in enumerate(("Right_Foot", "Left_Foot")):
        M.set_collision_group(m, g, body=foot, geom="sphere", size=0.02)
    out["plane-sphere feet"] = (m, 0.015)
    m = zmodel.copy()
    for g, foot in enumerate(("Right_Foot", "Left_Foot")):
        M.set_collision_group(m, g, body=foot, geom="capsule", size=(0.012, 0.04), quat=QY90)
    out["plane-capsule feet"] = (m, 0.007)
    m = zmodel.copy()
    M.set_collision_group(m, 1, body="left_knee_pitch", geom="capsule", size=(0.042, 0.035), pos=(0, 0, -0.04),
                          other_body="right_knee_pitch", other_geom="capsule", other_size=(0.042, 0.035), other_pos=(0, 0, -0.04), friction=(0.6, 0, 0),
                          other_friction=(0.9, 0, 0))
    out["box foot + capsule-capsule shins"] = (m, 0.0)
    m = zmodel.copy()
    M.set_collision_group(m, 1, body="left_knee_pitch", geom="capsule", size=(0.04, 0.035), pos=(0, 0, -0.04),
                          other_body="right_ankle_pitch", other_geom="sphere", other_size=0.05, other_pos=(0, 0, 0.01))
    out["box foot + sphere-capsule"] = (m, 0.0)
    m = zmodel.copy()
    M.set_collision_group(m, 0, body="Right_Foot", geom="capsule", size=(0.012, 0.04), quat=QY90)
    M.set_collision_group(m, 1, body="left_ankle_roll", geom="sphere", size=0.05, other_body="right_ankle_roll", other_geom="sphere", other_size=0.05,
                          priority=1)
    out["capsule foot + sphere-sphere ankles"] = (m, 0.007)
    m = zmodel.copy()  # both geoms on the SAME chain: thigh against the foot of the same leg
    M.set_collision_group(m, 1, body="Left_Foot", geom="sphere", size=0.03, other_body="left_hip_pitch", other_geom="sphere", other_size=0.125,
                          other_pos=(0.0, 0.0, -0.04))
    out["box foot + same-chain sphere-sphere"] = (m, 0.0)
    return out


def _states(model, n, seed, zoff):
    q, v = random_states(model, n, seed, joint_noise=0.3)
    q[:, 2] += zoff
    return q, v


# ---------------------------------------------------------------------------------------------------------------------------
# CPU: the oracle's primitives
# ---------------------------------------------------------------------------------------------------------------------------
def _forward(model, q, v, dtype=np.float64):
    import zb_oracle

    o = zb_oracle.Oracle(model, M.make_opts(dt=0.002), dtype, threads=2)
    n = len(q)
    st, out = o.new_state(n), o.new_out(n)
    o.reset(st, q.astype(dtype), v.astype(dtype), np.full(n, 1e9), out=out)
    return o, st, out


def test_plane_sphere_and_capsule_known_answers(zmodel):
    import zb_oracle

    V = variants(zmodel)
    for name in ("plane-sphere feet", "plane-capsule feet"):
        model, zoff = V[name]
        q, v = _states(model, 40, 3, zoff)
        o = zb_oracle.Oracle(model, M.make_opts(), np.float64)
        for i in range(8):
            d = o.forward_one(q[i].astype(np.float64), v[i].astype(np.float64))
            for g in range(2):
                c, R = d.get("box_xpos")[g], d.get("box_xmat")[g].reshape(3, 3)
                r = float(model.box_size[g][0])
                if "sphere" in name:
                    np.testing.assert_allclose(d.get("con_dist")[4 * g], c[2] - r, atol=1e-12)            # dist = height of the centre - radius
                    np.testing.assert_allclose(d.get("con_pos")[4 * g], [c[0], c[1], 0.5 * (c[2] - r)], atol=1e-12)  # midway between surface and plane
                    assert (d.get("con_dist")[4 * g + 1:4 * g + 4] == 1).all()                             # one contact per sphere
                else:
                    h = float(model.box_size[g][1])
                    for s, sg in enumerate((1.0, -1.0)):
                        e = c + sg * h * R[:, 2]
                        np.testing.assert_allclose(d.get("con_dist")[4 * g + s], e[2] - r, atol=1e-12)
                    fr = d.get("con_frame")[4 * g].reshape(3, 3)
                    np.testing.assert_allclose(fr[0], [0, 0, 1], atol=1e-12)
                    ax = R[:, 2] - np.array([0, 0, 1.0]) * R[2, 2]
                    np.testing.assert_allclose(fr[1], ax / np.linalg.norm(ax), atol=1e-9)                  # frame aligned with the capsule axis
                    np.testing.assert_allclose(fr @ fr.T, np.eye(3), atol=1e-9)
                    assert (d.get("con_dist")[4 * g + 2:4 * g + 4] == 1).all()


def test_body_body_primitives_against_brute_force(zmodel):
    import zb_oracle

    V = variants(zmodel)
    for name in ("box foot + capsule-capsule shins", "box foot + sphere-capsule", "capsule foot + sphere-sphere ankles"):
        model, zoff = V[name]
        q, v = _states(model, 60, 5, zoff)
        o = zb_oracle.Oracle(model, M.make_opts(), np.float64)
        seen_touch = 0
        for i in range(30):
            d = o.forward_one(q[i].astype(np.float64), v[i].astype(np.float64))
            g = 1
            p2, R2 = d.get("box_xpos")[g], d.get("box_xmat")[g].reshape(3, 3)
            p1, R1 = d.get("pair_xpos")[g], d.get("pair_xmat")[g].reshape(3, 3)

            def pts(p, R, typ, size):
                t = np.linspace(-1, 1, 801) if typ == M.GEOM_CAPSULE else np.zeros(1)
                return p[None] + t[:, None] * float(size[1]) * R[:, 2][None]

            A = pts(p1, R1, model.pair_type[g], model.pair_size[g])
            B = pts(p2, R2, model.box_type[g], model.box_size[g])
            dd = np.linalg.norm(A[:, None] - B[None], axis=-1)
            ia, ib = np.unravel_index(dd.argmin(), dd.shape)
            r1, r2 = float(model.pair_size[g][0]), float(model.box_size[g][0])
            dist = d.get("con_dist")[4 * g]
            assert abs(dist - (dd.min() - r1 - r2)) < 2e-5, (name, i)                  # brute-force segment sampling (step 1e-4 m)
            n = d.get("con_frame")[4 * g].reshape(3, 3)[0]
            nb = (B[ib] - A[ia]) / dd.min()
            assert np.dot(n, nb) > 1 - 1e-4, (name, i)                                  # normal from geom 1 to geom 2
            np.testing.assert_allclose(d.get("con_pos")[4 * g], A[ia] + nb * (r1 + 0.5 * dist), atol=2e-4)
            assert d.get("con_body")[4 * g] == model.box_body[g] and d.get("con_body1")[4 * g] == model.pair_body[g]
            seen_touch += dist < 0
        assert 0 < seen_touch < 30, (name, seen_touch)                                  # touching and separated states both occur


def test_two_body_jacobian_rows_by_finite_differences(zmodel):
    """efc_J of a body-body contact row == d/dq of the gap between the two material points at the contact, along the row direction."""
    import zb_oracle

    for name in ("box foot + capsule-capsule shins", "box foot + same-chain sphere-sphere"):
        model, zoff = variants(zmodel)[name]
        q, v = _states(model, 80, 9, zoff)
        o = zb_oracle.Oracle(model, M.make_opts(), np.float64)
        checked = 0
        for i in range(80):
            d = o.forward_one(q[i].astype(np.float64), np.zeros(26))
            if not d.get("con_active")[4]:
                continue
            b2, b1 = int(d.get("con_body")[4]), int(d.get("con_body1")[4])
            pos = d.get("con_pos")[4]
            fr, mu = d.get("con_frame")[4].reshape(3, 3), d.get("con_mu")[4]
            loc2 = d.get("xmat")[b2].reshape(3, 3).T @ (pos - d.get("xpos")[b2])
            loc1 = d.get("xmat")[b1].reshape(3, 3).T @ (pos - d.get("xpos")[b1])
            J = d.get("efc_J")[2 * 26 + 16:2 * 26 + 20]  # the four pyramid rows of contact slot 4
            qq = q[i].astype(np.float64)
            qq[3:7] /= np.linalg.norm(qq[3:7])
            Jfd = np.zeros((3, 26))
            for j in range(6, 26):  # hinge dofs (the base dofs move both bodies alike: their columns must vanish)
                def gap(dq):
                    q2 = qq.copy(); q2[j + 1] += dq
                    dd = o.forward_one(q2, np.zeros(26))
                    a = dd.get("xpos")[b2] + dd.get("xmat")[b2].reshape(3, 3) @ loc2
                    b = dd.get("xpos")[b1] + dd.get("xmat")[b1].reshape(3, 3) @ loc1
                    return a - b
                Jfd[:, j] = (gap(1e-6) - gap(-1e-6)) / 2e-6
            for k in range(4):
                dirk = fr[0] + (1 if k % 2 == 0 else -1) * mu * fr[1 + k // 2]
                np.testing.assert_allclose(J[k, 6:], dirk @ Jfd[:, 6:], atol=2e-6, err_msg=f"{name} env {i} row {k}")
                np.testing.assert_allclose(J[k, :6], 0, atol=1e-12)
            checked += 1
            if checked >= 4:
                break
        assert checked >= 2, name


# ---------------------------------------------------------------------------------------------------------------------------
# GPU: kernel vs oracle for every group configuration
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plane-sphere feet", "plane-capsule feet", "box foot + capsule-capsule shins", "box foot + sphere-capsule",
                                  "capsule foot + sphere-sphere ankles", "box foot + same-chain sphere-sphere"])
def test_collision_group_parity(zmodel, name):
    import os

    import torch
    import zb_oracle
    from ksim_gym_zbot_b200.engine import ZbotCudaEngine

    import parity_lib as P

    model, zoff = variants(zmodel)[name]
    n, seed = 1024, 21
    opts = M.make_opts(dt=0.002)
    q, v = _states(model, n, seed, zoff)
    eng = ZbotCudaEngine(model, n, device=0, opts=opts, ctrl_dt=0.002)
    ps = eng.reset(reset_qpos=torch.tensor(q), reset_qvel=torch.tensor(v), reset_t_rem=torch.full((n,), 1e9))
    torch.cuda.synchronize()
    thr = os.cpu_count() or 1
    o32, o64 = zb_oracle.Oracle(model, opts, np.float32, threads=thr), zb_oracle.Oracle(model, opts, np.float64, threads=thr)
    res = {}
    for k, o in (("f32", o32), ("f64", o64)):
        st, out = o.new_state(n), o.new_out(n)
        o.reset(st, q.astype(o.dtype), v.astype(o.dtype), np.full(n, 1e9), out=out)
        res[k] = (st, out)
    out64 = res["f64"][1]
    cd, cp, cm = ps.data.contact_dist.cpu().numpy(), ps.data.contact_pos.cpu().numpy(), ps.data.contact_mask.cpu().numpy()
    # ---- forward(): distances / positions at rounding level, contact sets identical except at the activation threshold
    # slots that exist (inert slots carry dist = 1).  Slot structure against the fp32 oracle: a box hovering above the floor has an
    # all-masked manifold (-1e6 + d^2 ties in fp32, not in fp64), so WHICH far-away vertices fill the inactive slots is a precision matter
    live = (out64["contact_dist"] < 0.5) & (res["f32"][1]["contact_dist"] < 0.5)
    same = (cd < 0.5) == (res["f32"][1]["contact_dist"] < 0.5)
    for g in range(2):  # sphere / capsule primitives: the slots are fixed; a box manifold may pick a duplicate vertex at a rounding-level tie
        sl = slice(4 * g, 4 * g + 4)
        assert same[:, sl].all() if model.box_type[g] != M.GEOM_BOX or model.pair_body[g] != 0 else same[:, sl].mean() > 0.995, (name, g)
    live &= same
    np.testing.assert_allclose(cd[live], out64["contact_dist"][live], rtol=0, atol=3e-6)
    # the closest points of two nearly parallel capsule axes slide along the axes under rounding (the distance does not): the position
    # bound follows the fp32 oracle's own deviation from fp64 there
    spread = np.abs(res["f32"][1]["contact_pos"].astype(np.float64) - out64["contact_pos"])
    assert (np.abs(cp - out64["contact_pos"])[live] <= 3e-6 + 4 * spread.max(axis=-1, keepdims=True).repeat(3, -1)[live]).all(), name
    assert np.quantile(np.abs(cp - out64["contact_pos"])[live], 0.99) < 3e-6, name
    edge = (np.abs(out64["contact_dist"]) < 3e-6).any(axis=1) | ~same.all(axis=1)
    assert np.array_equal(cm[~edge], out64["contact_mask"][~edge]), name
    frac = [(((out64["contact_mask"] >> (4 * g)) & 15) != 0).mean() for g in range(2)]
    print(f"{name}: environments with an active contact in group 0 / 1: {frac[0]:.2f} / {frac[1]:.2f}")
    assert all(0.03 < f < 0.99 for f in frac), frac  # both branches of every group are exercised
    np.testing.assert_allclose(ps.data.qacc.cpu().numpy(), out64["qacc"], rtol=0, atol=float(np.abs(out64["qacc"]).max()) * 3e-2)
    # ---- the smooth part + constraint rows without the iteration-capped solver in the picture: iterations = 0 -> qacc = qacc_smooth,
    #      and converged solves (Newton is not available with body-body groups: CG with a generous budget)
    # ---- one full physics step from the same state, statistically against fp64 with the fp32 oracle as yard-stick
    act = random_actions(n, 77)
    rep = P.compare_single_step(model, opts, {"qpos": q, "qvel": v, "qacc_warmstart": np.zeros((n, 26), np.float32), "act_pos": q[:, 7:].copy(),
                                              "act_vel": v[:, 6:].copy(), "act_torque": np.zeros((n, 20), np.float32),
                                              "push_t_rem": np.full(n, 1e9, np.float32), "time": np.zeros(n, np.float32)}, act)
    P.check_report(rep, tag=name, check_matched=False)
    eng.close()


@pytest.mark.gpu
def test_newton_is_refused_with_body_body_groups(zmodel):
    from ksim_gym_zbot_b200.engine import ZB_EUNSUP, ZbError, ZbotCudaEngine

    model, _ = variants(zmodel)["box foot + capsule-capsule shins"]
    with pytest.raises(ZbError) as ei:
        ZbotCudaEngine(model, 8, device=0, opts=M.make_opts(dt=0.002, solver=M.SOLVER_NEWTON))
    assert ei.value.code == ZB_EUNSUP
    model, _ = variants(zmodel)["plane-capsule feet"]  # floor groups keep the Newton path
    ZbotCudaEngine(model, 8, device=0, opts=M.make_opts(dt=0.002, solver=M.SOLVER_NEWTON)).close()


def test_mjcf_round_trip_of_every_group_configuration(zmodel):
    """to_mjcf -> compile_mjcf reproduces the collision-group fields (geom types, <contact><pair> elements) of every variant."""
    from ksim_gym_zbot_b200 import mjcf

    for name, (model, _) in variants(zmodel).items():
        xml = mjcf.to_mjcf(model, opts=M.make_opts(dt=0.002))
        c = mjcf.compile_mjcf(xml, servo_params=mjcf.servo_params_of(model))
        for f in ("box_body", "box_type", "box_size", "box_pos", "box_quat", "box_friction", "box_priority", "pair_body", "pair_type", "pair_size",
                  "pair_pos", "pair_quat", "pair_priority"):
            a, b = np.asarray(c.model.field(f)), np.asarray(model.field(f))
            world = np.asarray(model.field("pair_body")) == 0
            if f.startswith("pair_") and f != "pair_body":  # pair_* fields of floor groups are unused
                a, b = a[~world], b[~world]
            assert np.array_equal(a, b), (name, f, a, b)
        assert not c.ignored_geoms, (name, c.ignored_geoms)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["box foot + capsule-capsule shins", "capsule foot + sphere-sphere ankles"])
def test_collision_group_teacher_forced_rollout(zmodel, name):
    """Self-collision inside a developing rollout: 8 control steps (4 substeps each) of the kernel with random actions; at EVERY step the
    kernel's state is handed to the oracles and one physics step from it is compared like the single-step test — contacts of both groups
    open and close along the way."""
    import torch
    from ksim_gym_zbot_b200.engine import ZbotCudaEngine

    import parity_lib as P

    model, zoff = variants(zmodel)[name]
    n = 768
    opts = M.make_opts(dt=0.005)
    q, v = _states(model, n, 31, zoff)
    eng = ZbotCudaEngine(model, n, device=0, opts=opts, ctrl_dt=0.02)
    ps = eng.reset(reset_qpos=torch.tensor(q), reset_qvel=torch.tensor(v), reset_t_rem=torch.full((n,), 1e9))
    seen = np.zeros(2, dtype=bool)
    for t in range(8):
        st = P.cuda_state_np(ps)
        act = random_actions(n, 300 + t)
        rep = P.compare_single_step(model, opts, st, act)
        P.check_report(rep, tag=f"{name} step {t}", check_matched=False, verbose=False, cap_factor=8.0, flip_slack=0.05)  # (violent states: ~93 % of the line-search paths flip between ANY two evaluations, so the flip-rate comparison sits at saturation; maximum of a heavy-tailed error over 768 envs: the 8x of the follow test)
        eng.step(torch.tensor(act, device=eng.device), None, ps)
        torch.cuda.synchronize()
        cm = ps.data.contact_mask.cpu().numpy()
        seen |= np.array([(((cm >> (4 * g)) & 15) != 0).any() for g in range(2)])
        assert np.isfinite(ps.data.qpos.cpu().numpy()).all()
    assert seen.all(), seen  # both collision groups were active somewhere in the rollout
    eng.close()
