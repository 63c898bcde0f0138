"""CPU checks of tests/ref_rewards.py (the transcription of the reference's reward list used by the GPU rollout test): it must agree
with the oracle's own restatement of the reward inputs (zbo_reward_inputs, pinned in test_observations.py against train.py) and of
FeetAirtimeReward, and its joint-deviation tables must be the reference's (train.py:59-82, 585-641)."""
import numpy as np

import ref_rewards as RR
from helpers import random_states


def test_step_terms_agree_with_oracle_reward_inputs(zmodel):
    import zb_oracle
    from ksim_gym_zbot_b200 import model as M

    n = 300
    opts = M.make_opts(dt=0.005)
    o = zb_oracle.Oracle(zmodel, opts, np.float64, threads=2)
    q, v = random_states(zmodel, n, 5, tilt=0.3)
    st = o.new_state(n)
    out = o.new_out(n)
    o.reset(st, q.astype(np.float64), v.astype(np.float64), np.full(n, 1e9), out=out)
    act = (M.BIAS_POSE[None] + 0.2 * np.random.default_rng(1).uniform(-1, 1, (n, 20)))
    o.step(st, act, 4, out=out)
    feet, terms = o.reward_inputs(st, out, command=np.zeros((n, 7)))
    f = {"qpos": st["qpos"], "qvel": st["qvel"], "xquat": out["xquat"], "site_xpos": out["site_xpos"], "sensordata": out["sensordata"],
         "done": np.zeros(n, bool)}
    t = RR.step_terms(f, int(zmodel.site_body[1]), int(zmodel.site_body[2]))
    assert np.array_equal(terms[:, 0], t["single_foot_contact"]) and np.array_equal(terms[:, 4] > 0.5, t["left_contact"])
    np.testing.assert_allclose(terms[:, 1], t["feet_orientation"], atol=1e-9)
    edge = np.abs(t["feet_dist"] - 0.12) < 2e-6  # (q / (|q| + 1e-6): quaternion-product and rotation-matrix forms differ at 1e-6)
    assert np.array_equal(terms[~edge, 2], t["feet_too_close"][~edge])
    np.testing.assert_allclose(np.linalg.norm(feet[:, :3] - feet[:, 3:], axis=1), t["feet_dist"], atol=1e-6)
    assert 0 < t["upright"].min() and t["upright"].max() <= 1 + 1e-12 and (t["straight_leg"] >= 0).all()


def test_airtime_and_tables():
    import zb_oracle

    r = np.random.default_rng(3)
    T, n = 40, 50
    lc, rc, dn = r.random((T, n)) < 0.6, r.random((T, n)) < 0.6, r.random((T, n)) < 0.05
    c0 = r.uniform(0, 0.3, (n, 2)).astype(np.float32)
    a, ca = RR.feet_airtime(lc, rc, dn, c0)
    b, cb = zb_oracle.feet_airtime(lc, rc, dn, c0)
    assert np.array_equal(a, b) and np.array_equal(ca, cb)
    from ksim_gym_zbot_b200 import model as M

    assert [x[0] for x in RR.JOINT_BIASES] == M.JOINT_NAMES and np.allclose([x[1] for x in RR.JOINT_BIASES], M.BIAS_POSE)
    assert set(RR.SCALES) == {"stay_alive", "upright", "naive_forward", "naive_forward_orientation", "linvel_y_penalty", "single_foot_contact",
                              "feet_airtime", "feet_orientation", "feet_too_close", "straight_leg", "ankle_knee", "arm_pose"}  # train.py:1548-1585


def test_product_rewards_equal_transcription_on_cpu_tensors(zmodel):
    """ksim-gym-zbot_b200/rewards.py (torch ops, the product's reward list) against tests/ref_rewards.py on random trajectories."""
    import types

    import torch
    from ksim_gym_zbot_b200 import rewards as PR

    r = np.random.default_rng(9)
    T, n = 7, 64
    q = r.normal(size=(T, n, 27)).astype(np.float32) * 0.3
    q[..., 3:7] = r.normal(size=(T, n, 4)) * r.uniform(0.6, 1.4, (T, n, 1))
    v = r.normal(size=(T, n, 26)).astype(np.float32)
    xq = r.normal(size=(T, n, 26, 4)).astype(np.float32)
    rt = (r.random((T + 1, n, 6)) < 0.5).astype(np.float32)
    dn = r.random((T, n)) < 0.1
    air = r.normal(size=(T, n)).astype(np.float32)
    traj = types.SimpleNamespace(qpos=torch.tensor(q), qvel=torch.tensor(v), xquat=torch.tensor(xq), reward_terms=torch.tensor(rt),
                                 done=torch.tensor(dn.astype(np.uint8)))
    lb, rb = int(zmodel.site_body[1]), int(zmodel.site_body[2])
    terms = PR.reward_terms(traj, torch.tensor(air), lb, rb)
    assert set(terms) == set(RR.SCALES) and {k: s for k, (s, _) in RR.SCALES.items()} == PR.SCALES
    for t in range(T):
        f = {"qpos": q[t], "qvel": v[t], "xquat": xq[t], "site_xpos": np.zeros((n, 3, 3), np.float32), "sensordata": np.zeros((n, 18), np.float32), "done": dn[t]}
        ref = RR.step_terms(f, lb, rb)
        for k in ("stay_alive", "upright", "naive_forward", "naive_forward_orientation", "linvel_y_penalty", "feet_orientation", "straight_leg",
                  "ankle_knee", "arm_pose"):
            np.testing.assert_allclose(terms[k][t].numpy(), ref[k], rtol=0, atol=5e-6, err_msg=k)
    assert torch.equal(terms["single_foot_contact"], traj.reward_terms[:T, :, 0]) and torch.equal(terms["feet_too_close"], traj.reward_terms[:T, :, 2])
