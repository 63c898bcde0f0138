"""Observation epilogue (ZbObs): the reference's network input vectors.

CPU part: the oracle's ``zbo_observe`` against an independent numpy transcription of the reference's own code —
``rotate_quat_by_quat`` (/root/reference/train.py:751-787), ``ImuOrientationObservation.observe_stateful``
(train.py:837-866, carry :846-848), ``run_actor`` / ``run_critic`` concatenations (train.py:1624-1681).
GPU part: the kernel epilogue (through the C ABI) against the oracle on the kernel's own state / mjx.Data arrays,
including the carry restart of auto-reset environments, and the guarantee that the epilogue leaves the physics untouched.
"""

import numpy as np
import pytest

from helpers import random_actions, random_states

NU, NB = 20, 26


# ---- independent transcription of train.py (numpy, dtype-generic) ---------------------------------------------------
def rotate_quat_by_quat(quat_to_rotate, rotating_quat, inverse=False, eps=1e-6):
    dt = quat_to_rotate.dtype
    eps = dt.type(eps)
    quat_to_rotate = quat_to_rotate / (np.linalg.norm(quat_to_rotate, axis=-1, keepdims=True).astype(dt) + eps)
    rotating_quat = rotating_quat / (np.linalg.norm(rotating_quat, axis=-1, keepdims=True).astype(dt) + eps)
    if inverse:
        rotating_quat = rotating_quat.copy()
        rotating_quat[..., 1:] *= -1
    w1, x1, y1, z1 = np.split(rotating_quat, 4, axis=-1)
    w2, x2, y2, z2 = np.split(quat_to_rotate, 4, axis=-1)
    w = w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2
    x = w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2
    y = w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2
    z = w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2
    result = np.concatenate([w, x, y, z], axis=-1)
    return result / (np.linalg.norm(result, axis=-1, keepdims=True).astype(dt) + eps)


def euler_to_quat(euler):  # xax.euler_to_quat (roll, pitch, yaw) [R]
    r, p, y = euler[..., 0:1] * 0.5, euler[..., 1:2] * 0.5, euler[..., 2:3] * 0.5
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.concatenate([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
                           cr * cp * sy - sr * sp * cy], axis=-1).astype(euler.dtype)


def imu_orientation_observe(sensordata, command, carry):
    framequat = sensordata[..., 0:4]
    yaw = command[..., 3:4]
    hq = euler_to_quat(np.concatenate([np.zeros_like(yaw), np.zeros_like(yaw), yaw], axis=-1))
    back = rotate_quat_by_quat(framequat, hq, inverse=True)
    back = np.where(back[..., 0:1] < 0, -back, back)
    x, lag = carry
    x = x * lag[..., None] + back * (1 - lag[..., None])
    return x, (x, lag)


def reference_obs(qpos, qvel, cinert, cvel, sensordata, actuator_force, command, carry, noise_quat, noise_acc):
    imu, carry = imu_orientation_observe(sensordata, command, carry)
    imu = imu + noise_quat
    acc = sensordata[..., 4:7] + noise_acc
    n = qpos.shape[0]
    actor = np.concatenate([qpos[:, 7:], qvel[:, 6:], imu, command[:, :2], command[:, 3:4], command[:, 4:]], axis=-1)
    critic = np.concatenate([qpos[:, 7:], qvel[:, 6:] / 10.0, cinert[:, 1:].reshape(n, -1), cvel[:, 1:].reshape(n, -1), acc,
                             sensordata[:, 7:10], imu, command, actuator_force / 100.0, qpos[:, 0:3], qpos[:, 3:7]], axis=-1)
    return actor, critic, carry


def _random_fields(n, seed, dtype):
    r = np.random.default_rng(seed)
    f = dict(qpos=r.normal(size=(n, 27)), qvel=r.normal(size=(n, 26)), cinert=r.normal(size=(n, NB, 10)), cvel=r.normal(size=(n, NB, 6)),
             sensordata=r.normal(size=(n, 18)), actuator_force=r.normal(size=(n, NU)) * 3, command=r.uniform(-2, 2, (n, 7)))
    q = r.normal(size=(n, 4))
    f["sensordata"][:, 0:4] = q / np.linalg.norm(q, axis=1, keepdims=True)
    return {k: v.astype(dtype) for k, v in f.items()}


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-13), (np.float32, 2e-6)])
def test_oracle_observe_matches_reference_transcription(zmodel, dtype, tol):
    import zb_oracle
    from ksim_gym_zbot_b200 import model as M

    assert (50, 484) == (2 * NU + 4 + 2 + 1 + 3, 2 * NU + 250 + 150 + 3 + 3 + 4 + 7 + NU + 3 + 4)  # train.py:31-54
    n = 257
    orc = zb_oracle.Oracle(zmodel, M.make_opts(), dtype)
    r = np.random.default_rng(5)
    x_ref, lag_ref = np.zeros((n, 4), dtype), r.uniform(0.0, 0.1, n).astype(dtype)
    x, lag = x_ref.copy(), lag_ref.copy()
    for t in range(4):
        f = _random_fields(n, 100 + t, dtype)
        if t == 0:
            f["command"][:] = 0  # ConstantZeroCommand (train.py:416-427): heading 0 -> the backspin is the identity
        nq, na = (r.normal(size=(n, 4)) * 0.0175).astype(dtype), (r.normal(size=(n, 3)) * 0.5).astype(dtype)
        reinit = (r.uniform(size=n) < 0.2) if t == 2 else None
        new_lag = r.uniform(0.0, 0.1, n).astype(dtype)
        if reinit is not None:  # ksim: carry <- initial_carry = (zeros(4), U(lag_range)) for reset environments
            x_ref = np.where(reinit[:, None], 0, x_ref).astype(dtype)
            lag_ref = np.where(reinit, new_lag, lag_ref).astype(dtype)
        a_ref, c_ref, (x_ref, lag_ref) = reference_obs(f["qpos"], f["qvel"], f["cinert"], f["cvel"], f["sensordata"],
                                                       f["actuator_force"], f["command"], (x_ref, lag_ref), nq, na)
        st = {"qpos": f["qpos"], "qvel": f["qvel"]}
        out = {k: f[k] for k in ("cinert", "cvel", "sensordata", "actuator_force")}
        a, c = orc.observe(st, out, x, lag, command=f["command"], reinit=reinit, reset_lag=new_lag, noise_quat=nq, noise_acc=na)
        assert a.shape == (n, 50) and c.shape == (n, 484)
        np.testing.assert_allclose(a, a_ref, rtol=0, atol=tol)
        np.testing.assert_allclose(c, c_ref, rtol=0, atol=tol * 10)
        np.testing.assert_allclose(x, x_ref, rtol=0, atol=tol)
        np.testing.assert_array_equal(lag, lag_ref)
        if t == 0:  # zero heading: imu_orientation == (1-lag) * framequat on the w >= 0 hemisphere
            fq = f["sensordata"][:, 0:4] * np.sign(f["sensordata"][:, 0:1])
            np.testing.assert_allclose(x, fq * (1 - lag[:, None]), atol=3e-6)


# ---- GPU: kernel epilogue vs oracle ---------------------------------------------------------------------------------
@pytest.mark.gpu
def test_observation_epilogue_matches_oracle(zmodel):
    import os

    import torch
    import zb_oracle
    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import PushEventConfig, ZbotCudaEngine

    n, S = 300, 4  # 300: the last CTA has spare warps
    opts = M.make_opts(dt=0.005)
    eng = ZbotCudaEngine(zmodel, n, device=0, opts=opts, ctrl_dt=0.02, push=PushEventConfig())
    ref = ZbotCudaEngine(zmodel, n, device=0, opts=opts, ctrl_dt=0.02, push=PushEventConfig())
    orc = zb_oracle.Oracle(zmodel, opts, np.float32, threads=os.cpu_count() or 1)
    q, v = random_states(zmodel, n, 11, tilt=0.6)  # large tilts: some environments terminate within a few steps
    r = np.random.default_rng(3)
    T = lambda a, dt=torch.float32: torch.tensor(np.ascontiguousarray(a), dtype=dt, device=eng.device)  # noqa: E731
    trem = T(np.full(n, 0.03, np.float32))
    obs = eng.new_observations(n)
    lag0 = r.uniform(0.0, 0.1, n).astype(np.float32)
    cmd = r.uniform(-1.5, 1.5, (n, 7)).astype(np.float32)
    ps = eng.reset(reset_qpos=T(q), reset_qvel=T(v), reset_t_rem=trem, obs=obs, command=T(cmd), reset_imu_lag=T(lag0))
    ps_ref = ref.reset(reset_qpos=T(q), reset_qvel=T(v), reset_t_rem=trem)
    x, lag = np.zeros((n, 4), np.float32), lag0.copy()

    def check(ps, reinit, new_lag, nq, na, tag):
        st = {"qpos": ps.data.qpos.cpu().numpy(), "qvel": ps.data.qvel.cpu().numpy()}
        out = {k: getattr(ps.data, k).cpu().numpy() for k in ("cinert", "cvel", "sensordata", "actuator_force")}
        a, c = orc.observe(st, out, x, lag, command=cmd, reinit=reinit, reset_lag=new_lag, noise_quat=nq, noise_acc=na)
        torch.cuda.synchronize()
        np.testing.assert_allclose(obs.actor_obs.cpu().numpy(), a, rtol=0, atol=2e-6, err_msg=tag)
        # critic: cinert / cvel entries are copies (exact); the scaled / filtered ones are within rounding
        np.testing.assert_allclose(obs.critic_obs.cpu().numpy(), c, rtol=1e-6, atol=2e-6, err_msg=tag)
        np.testing.assert_array_equal(obs.critic_obs.cpu().numpy()[:, 40:440], c[:, 40:440])
        np.testing.assert_allclose(obs.imu_x.cpu().numpy(), x, rtol=0, atol=2e-6, err_msg=tag)
        np.testing.assert_array_equal(obs.imu_lag.cpu().numpy(), lag)

    check(ps, np.ones(n, bool), lag0, None, None, "reset")
    rq, rv = random_states(zmodel, n, 12)
    rq[:, 3:7] /= np.linalg.norm(rq[:, 3:7], axis=1, keepdims=True)
    done = torch.zeros(n, dtype=torch.uint8, device=eng.device)
    done_ref = torch.zeros_like(done)
    n_done = 0
    for t in range(12):
        act = T(random_actions(n, 50 + t))
        pr = eng.sample_push(n)
        nq, na = (r.normal(size=(n, 4)) * 0.0175).astype(np.float32), (r.normal(size=(n, 3)) * 0.5).astype(np.float32)
        new_lag = r.uniform(0.0, 0.1, n).astype(np.float32)
        ar = (T(rq), T(rv), T(np.full(n, 0.05, np.float32)))
        eng.step(act, None, ps, push_rand=pr, auto_reset=ar, done_out=done, obs=obs, command=T(cmd), reset_imu_lag=T(new_lag),
                 obs_noise=(T(nq), T(na)))
        ref.step(act, None, ps_ref, push_rand=pr, auto_reset=ar, done_out=done_ref)
        dn = done.cpu().numpy().astype(bool)
        n_done += int(dn.sum())
        # the epilogue must not change the physics: bitwise the same state / outputs / mask as the plain entry point
        assert torch.equal(done, done_ref)
        for k in ("qpos", "qvel", "qacc_warmstart", "cinert", "cvel", "sensordata", "actuator_force", "xpos"):
            assert torch.equal(getattr(ps.data, k), getattr(ps_ref.data, k)), k
        check(ps, dn, new_lag, nq, na, f"step {t}")
    assert n_done > 5, "the scenario is meant to exercise the carry restart of auto-reset environments"


@pytest.mark.gpu
def test_observation_argument_checks(zmodel):
    import ctypes

    import torch
    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import ZB_EINVAL, ZbError, ZbotCudaEngine

    n = 8
    eng = ZbotCudaEngine(zmodel, n, device=0, opts=M.make_opts(dt=0.005), ctrl_dt=0.02)
    q, v = random_states(zmodel, n, 1)
    ps = eng.reset(reset_qpos=torch.tensor(q), reset_qvel=torch.tensor(v))
    obs = eng.new_observations(n)
    ob = obs.c_struct()
    ob.imu_x = None  # actor_obs without the filter carry
    st, out = eng._c_state(ps), eng._c_out(ps)
    act = torch.zeros((n, 20), device=eng.device)
    rc = eng.lib.zb_step_obs_f32(eng._h, eng._stream(), None, ctypes.byref(st), ctypes.c_void_p(act.data_ptr()), None, 0, ctypes.byref(out),
                                 ctypes.byref(ob), n, 4, None)
    assert rc == ZB_EINVAL and b"imu_x" in eng.lib.zb_last_error()
    with pytest.raises(ZbError):
        eng._check(rc)
    # actor only (no critic buffer) is fine
    obs2 = eng.new_observations(n, critic=False)
    eng.step(act, None, ps, obs=obs2)
    torch.cuda.synchronize()
    assert torch.isfinite(obs2.actor_obs).all() and float(obs2.actor_obs[:, 40:44].norm(dim=1).min()) > 0.5


# ---- reward inputs: transcription of the reference's reward / observation classes ---------------------------------------
def quat_to_euler(q):  # xax.quat_to_euler [R]
    q = q / (np.linalg.norm(q, axis=-1, keepdims=True) + q.dtype.type(1e-6))
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    roll = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    sp = 2 * (w * y - z * x)
    pitch = np.where(np.abs(sp) >= 1, np.sign(sp) * np.pi / 2, np.arcsin(np.clip(sp, -1, 1)))
    yaw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return np.stack([roll, pitch, yaw], axis=-1)


def rotate_vector_by_quat_inverse(v, q):  # xax.rotate_vector_by_quat(v, q, inverse=True) [R]: q^-1 v q
    q = q / (np.linalg.norm(q, axis=-1, keepdims=True) + q.dtype.type(1e-6))
    w, u = q[..., 0:1], -q[..., 1:]
    t = 2 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def reference_reward_inputs(qpos, xpos, xquat, site_xpos, sensordata, command, left_body, right_body):
    # FeetPositionObservation.observe, train.py:452-465 (floor_threshold = 0, in_robot_frame)
    fl = rotate_vector_by_quat_inverse(site_xpos[:, 1], qpos[:, 3:7])
    fr = rotate_vector_by_quat_inverse(site_xpos[:, 2], qpos[:, 3:7])
    feet = np.concatenate([fl, fr], axis=-1)
    # SimpleSingleFootContactReward.get_reward, train.py:698-711 with stand_still_threshold=None (train.py:1559)
    lc, rc = sensordata[:, 10] > 0.1, sensordata[:, 11] > 0.1
    single = np.logical_xor(lc, rc)
    # FeetOrientationReward.get_reward, train.py:680-689
    err = np.abs(quat_to_euler(xquat[:, left_body])[:, :2]).sum(-1) + np.abs(quat_to_euler(xquat[:, right_body])[:, :2]).sum(-1)
    orient = np.exp(-err / 0.25)
    # FeetTooClosePenalty.get_reward, train.py:740-748
    close = np.linalg.norm(feet[:, :3] - feet[:, 3:], axis=-1) < 0.12
    # BaseHeightReward.get_reward, train.py:476-483
    height = np.exp(-np.abs(xpos[:, 1, 2] - (command[:, 4] + 0.27)) / 0.25)
    terms = np.stack([single, orient, close, height, lc, rc], axis=-1).astype(qpos.dtype)
    return feet, terms


# (tolerance 6e-6 also in fp64 — the test quaternions have norms down to 0.5: the reference normalises quaternions as q / (|q| + 1e-6); rotating with the rotation matrix of that
#  slightly-short quaternion or with the quaternion-product formula differs at that 1e-6 level — the exact xax formula is [R])
@pytest.mark.parametrize("dtype,tol", [(np.float64, 6e-6), (np.float32, 8e-6)])
def test_oracle_reward_inputs_match_reference_transcription(zmodel, dtype, tol):
    import zb_oracle
    from ksim_gym_zbot_b200 import model as M

    n = 400
    r = np.random.default_rng(17)
    orc = zb_oracle.Oracle(zmodel, M.make_opts(), dtype)
    q = r.normal(size=(n, 27))
    q[:, 3:7] = r.normal(size=(n, 4)) * r.uniform(0.5, 1.5, (n, 1))
    xq = r.normal(size=(n, NB, 4))
    xq /= np.linalg.norm(xq, axis=-1, keepdims=True)
    sx = r.normal(size=(n, 3, 3)) * 0.1
    sx[: n // 2, 2] = sx[: n // 2, 1] + r.normal(size=(n // 2, 3)) * 0.06  # feet around the 0.12 m threshold
    sd = r.normal(size=(n, 18))
    sd[:, 10:12] = r.choice([0.0, 0.05, 0.2, 3.0], size=(n, 2))
    f = dict(qpos=q, xpos=r.normal(size=(n, NB, 3)) * 0.3, xquat=xq, site_xpos=sx, sensordata=sd, command=r.uniform(-0.2, 0.2, (n, 7)))
    f = {k: v.astype(dtype) for k, v in f.items()}
    lb, rb = int(zmodel.site_body[1]), int(zmodel.site_body[2])
    feet_ref, terms_ref = reference_reward_inputs(f["qpos"], f["xpos"], f["xquat"], f["site_xpos"], f["sensordata"], f["command"], lb, rb)
    feet, terms = orc.reward_inputs({"qpos": f["qpos"]}, f, command=f["command"])
    np.testing.assert_allclose(feet, feet_ref, rtol=0, atol=tol)
    edge = np.abs(np.linalg.norm(feet_ref[:, :3] - feet_ref[:, 3:], axis=-1) - 0.12) < 1e-5  # threshold ties may flip at rounding
    np.testing.assert_allclose(terms[~edge], terms_ref[~edge], rtol=0, atol=tol)
    assert 0.2 < terms[:, 0].mean() < 0.8 and 0.05 < terms[:, 2].mean() < 0.95  # both branches exercised


@pytest.mark.gpu
def test_reward_inputs_epilogue_matches_oracle(zmodel):
    import os

    import torch
    import zb_oracle
    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import ZbotCudaEngine

    n = 300
    opts = M.make_opts(dt=0.005)
    eng = ZbotCudaEngine(zmodel, n, device=0, opts=opts, ctrl_dt=0.02)
    orc = zb_oracle.Oracle(zmodel, opts, np.float32, threads=os.cpu_count() or 1)
    q, v = random_states(zmodel, n, 31, tilt=0.3)
    r = np.random.default_rng(4)
    cmd = r.uniform(-0.1, 0.1, (n, 7)).astype(np.float32)
    T = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=eng.device)  # noqa: E731
    obs = eng.new_observations(n, rewards=True)
    ps = eng.reset(reset_qpos=T(q), reset_qvel=T(v), reset_t_rem=torch.full((n,), 1e9), obs=obs, command=T(cmd))

    def check(tag):
        torch.cuda.synchronize()
        st = {"qpos": ps.data.qpos.cpu().numpy()}
        out = {k: getattr(ps.data, k).cpu().numpy() for k in ("xpos", "xquat", "site_xpos", "sensordata")}
        feet, terms = orc.reward_inputs(st, out, command=cmd)
        np.testing.assert_allclose(obs.feet_pos_obs.cpu().numpy(), feet, rtol=0, atol=2e-6, err_msg=tag)
        got = obs.reward_terms.cpu().numpy()
        edge = (np.abs(np.linalg.norm(feet[:, :3] - feet[:, 3:], axis=-1) - 0.12) < 1e-5) | (np.abs(out["sensordata"][:, 10:12] - 0.1) < 1e-6).any(axis=1)
        np.testing.assert_allclose(got[~edge], terms[~edge], rtol=0, atol=5e-6, err_msg=tag)
        return got

    got = check("reset")
    seen = [got]
    for t in range(6):
        eng.step(T(random_actions(n, 80 + t)), None, ps, obs=obs, command=T(cmd))
        seen.append(check(f"step {t}"))
    allv = np.concatenate(seen)
    assert 0 < allv[:, 0].mean() < 1 and allv[:, 4].mean() > 0.1 and (allv[:, 1] < 0.99).any()  # contacts and tilted feet occur
    # reward inputs alone (no network vectors, no filter carry) are a valid request
    from ksim_gym_zbot_b200.engine import Observations

    only = Observations(imu_x=None, imu_lag=None, actor_obs=None, critic_obs=None, feet_pos_obs=torch.zeros((n, 6), device=eng.device),
                        reward_terms=torch.zeros((n, 6), device=eng.device))
    eng.step(T(random_actions(n, 99)), None, ps, obs=only, command=T(cmd))
    torch.cuda.synchronize()
    assert torch.isfinite(only.reward_terms).all() and float(only.feet_pos_obs.abs().max()) > 0


# ---- FeetAirtimeReward over a rollout chunk (train.py:485-546) -----------------------------------------------------------
def feet_airtime_reference_one_env(touch_l, touch_r, done, carry, ctrl_dt=0.02, touchdown_penalty=0.3):
    """Line-by-line numpy transcription of FeetAirtimeReward.get_reward_stateful for ONE environment (the reference vmaps it),
    stand_still_threshold = None (train.py:1560-1565)."""
    f32 = np.float32

    def airtime_sequence(initial, contact_bool, done):  # train.py:503-516
        t, out = f32(initial), []
        for c in np.logical_or(contact_bool, done):
            t = f32(0.0) if c else f32(t + f32(ctrl_dt))
            out.append(t)
        return t, np.array(out, dtype=f32)

    left_contact = np.where(touch_l > 0.1, True, False)[:, 0]   # train.py:519-520
    right_contact = np.where(touch_r > 0.1, True, False)[:, 0]
    left_carry, left_air = airtime_sequence(carry[0], left_contact, done)
    right_carry, right_air = airtime_sequence(carry[1], right_contact, done)

    def touchdown(c):  # train.py:528-531
        prev = np.concatenate([np.array([False]), c[:-1]])
        return np.logical_and(c, np.logical_not(prev))

    td_l, td_r = touchdown(left_contact), touchdown(right_contact)
    left_air_shifted, right_air_shifted = np.roll(left_air, 1), np.roll(right_air, 1)
    reward = (left_air_shifted - f32(touchdown_penalty)) * td_l.astype(f32) + (right_air_shifted - f32(touchdown_penalty)) * td_r.astype(f32)
    return reward.astype(f32), np.array([left_carry, right_carry], dtype=f32)


def _airtime_case(T, n, seed):
    r = np.random.default_rng(seed)
    # contact phases of random length so that swing times and touchdowns both occur
    lc = np.zeros((T, n), bool); rc = np.zeros((T, n), bool)
    for arr in (lc, rc):
        for i in range(n):
            t, s = 0, bool(r.integers(0, 2))
            while t < T:
                k = int(r.integers(1, 12))
                arr[t:t + k, i] = s
                s, t = not s, t + k
    done = r.uniform(size=(T, n)) < 0.03
    carry = r.uniform(0, 0.3, (n, 2)).astype(np.float32)
    return lc, rc, done, carry


def test_feet_airtime_oracle_matches_reference_transcription():
    import zb_oracle

    T, n = 57, 40
    lc, rc, done, carry = _airtime_case(T, n, 3)
    rew, new_carry = zb_oracle.feet_airtime(lc, rc, done, carry, ctrl_dt=0.02, touchdown_penalty=0.3)
    for i in range(n):
        touch_l = np.where(lc[:, i], 2.0, 0.05)[:, None].astype(np.float32)
        touch_r = np.where(rc[:, i], 0.7, 0.0)[:, None].astype(np.float32)
        r_ref, c_ref = feet_airtime_reference_one_env(touch_l, touch_r, done[:, i], carry[i])
        np.testing.assert_array_equal(rew[:, i], r_ref)
        np.testing.assert_array_equal(new_carry[i], c_ref)
    assert (rew != 0).mean() > 0.02 and (rew > 0).any() and (rew < 0).any()


@pytest.mark.gpu
def test_feet_airtime_kernel_is_exact(zmodel):
    import torch
    import zb_oracle
    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import ZbotCudaEngine

    eng = ZbotCudaEngine(zmodel, 4, device=0, opts=M.make_opts(dt=0.005), ctrl_dt=0.02)
    for T, n, seed in ((1, 5, 1), (200, 333, 2), (64, 4096, 3)):
        lc, rc, done, carry = _airtime_case(T, n, seed)
        rew, new_carry = zb_oracle.feet_airtime(lc, rc, done, carry, ctrl_dt=0.02, touchdown_penalty=0.3)
        # contacts as slices 4 / 5 of the stacked per-step reward_terms [T, N, 6] (non-contiguous views), done time-major
        terms = torch.zeros((T, n, 6), device="cuda")
        terms[:, :, 4] = torch.tensor(lc, dtype=torch.float32)
        terms[:, :, 5] = torch.tensor(rc, dtype=torch.float32)
        c = torch.tensor(carry, device="cuda")
        out = eng.feet_airtime_reward(terms[:, :, 4], terms[:, :, 5], torch.tensor(done.astype(np.uint8), device="cuda"), c, touchdown_penalty=0.3)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy(), rew)
        np.testing.assert_array_equal(c.cpu().numpy(), new_carry)
        # env-major layout ([N, T] storage viewed as [T, N]) gives the same
        lcT = torch.tensor(lc.T.copy(), dtype=torch.float32, device="cuda").T
        rcT = torch.tensor(rc.T.copy(), dtype=torch.float32, device="cuda").T
        dT = torch.tensor(done.T.copy().astype(np.uint8), device="cuda").T
        c2 = torch.tensor(carry, device="cuda")
        out2 = torch.empty((n, T), device="cuda").T
        eng.feet_airtime_reward(lcT, rcT, dT, c2, reward=out2, touchdown_penalty=0.3)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out2.cpu().numpy(), rew)
