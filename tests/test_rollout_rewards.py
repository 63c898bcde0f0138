"""50-control-step rollouts with the reference's REAL terminations and reward list (north_star: "rewards /
terminations equal over 50-step rollouts from the same seed").

Reward list = train.py:1546-1586, terminations = train.py:1588-1593, transcribed in tests/ref_rewards.py.  Two comparisons:

  * teacher-forced (every control step the oracles restart from the kernel's state, so a difference is THIS step's, not the
    accumulated divergence of a chaotic rollout): per reward term, threshold flags must be EQUAL to the fp64 oracle's except where
    the thresholded quantity sits within the fp32 floor of its threshold (the flip rate is reported and bounded by the fp32
    oracles' own), smooth terms agree at the fp32 floor (bounded by what the fp32 oracles show against fp64);
  * the kernel's own reward inputs (ZbObs.reward_terms / feet_pos_obs from the observation epilogue + zb_feet_airtime_f32 over the
    rollout chunk) against the same transcription evaluated on the kernel's exported mjx.Data fields: the values a trainer would
    consume are the reference's formulas, bit for bit on flags.
"""
import os

import numpy as np
import pytest

import ref_rewards as RR

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

T_STEPS = 50


def _fields_cuda(ps, done):
    d = ps.data
    return {"qpos": d.qpos.cpu().numpy(), "qvel": d.qvel.cpu().numpy(), "xquat": d.xquat.cpu().numpy(), "site_xpos": d.site_xpos.cpu().numpy(),
            "sensordata": d.sensordata.cpu().numpy(), "xpos": d.xpos.cpu().numpy(), "done": done}


def test_rollout_50_steps_reference_rewards_and_terminations(zmodel):
    import parity_lib as P
    import zb_oracle
    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import EnvParams, PushEventConfig, ZbotCudaEngine

    n, S, seed = 2048, 4, 5
    opts = M.make_opts(dt=0.005)
    # pushes strong and frequent enough that a good part of the batch falls inside 50 steps: terminations + resets are exercised
    push = PushEventConfig(x_linvel=0.8, y_linvel=0.8, z_linvel=0.3, interval_range=(0.1, 0.4))
    params = M.randomize_env_params(zmodel, n, 77)
    eng = ZbotCudaEngine(zmodel, n, device=0, opts=opts, ctrl_dt=0.02, push=push, seed=seed)
    dev = eng.device
    ep = EnvParams(n, dev, **params)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 17)
    bias = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=dev)
    rq, rv, rt = eng.sample_reset(n)
    rq_n, rv_n, rt_n = rq.cpu().numpy(), rv.cpu().numpy(), rt.cpu().numpy()
    obs = eng.new_observations(n, rewards=True)
    lag = eng.sample_imu_lag(n, (0.0, 0.1))
    ps = eng.reset(ep, reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt, obs=obs, reset_imu_lag=lag)
    done_d = torch.zeros((n,), dtype=torch.uint8, device=dev)
    threads = os.cpu_count() or 1
    orcs = {"f32": zb_oracle.Oracle(zmodel, opts, np.float32, threads=threads), "f64": zb_oracle.Oracle(zmodel, opts, np.float64, threads=threads),
            "f32p": zb_oracle.Oracle(zmodel, opts, np.float32, threads=threads)}
    lb, rb = int(zmodel.site_body[1]), int(zmodel.site_body[2])
    flag_terms = ("stay_alive", "single_foot_contact", "feet_too_close", "left_contact", "right_contact")
    cont_terms = ("upright", "naive_forward", "naive_forward_orientation", "linvel_y_penalty", "feet_orientation", "straight_leg", "ankle_knee", "arm_pose")
    flips = {a: {k: 0 for k in flag_terms + ("done",)} for a in ("cuda", "f32")}
    unexplained = {k: 0 for k in flag_terms + ("done",)}
    err = {a: {k: [] for k in cont_terms + ("total",)} for a in ("cuda", "f32")}
    traj_terms, traj_epi, traj_done = [], [], []
    n_done = 0
    for t in range(T_STEPS):
        act = bias + 0.3 * (torch.rand((n, M.NU), generator=g, device=dev) * 2 - 1)
        pr = eng.sample_push(n)
        pre = P.cuda_state_np(ps)
        eng.step(act, ep, ps, push_rand=pr, auto_reset=(rq, rv, rt), done_out=done_d, obs=obs, reset_imu_lag=lag)
        torch.cuda.synchronize()
        done_c = done_d.cpu().numpy() != 0
        n_done += int(done_c.sum())
        fc = _fields_cuda(ps, done_c)
        tc = RR.step_terms(fc, lb, rb)
        traj_terms.append(tc); traj_done.append(done_c)
        traj_epi.append((obs.reward_terms.cpu().numpy().copy(), obs.feet_pos_obs.cpu().numpy().copy()))
        act_n, pr_n = act.cpu().numpy(), pr.cpu().numpy()
        to = {}
        for name, o in orcs.items():
            st = P.oracle_state_from(pre, o.dtype, 991 + t if name == "f32p" else None)
            out = o.new_out(n)
            o.step(st, act_n.astype(o.dtype), S, params=params, push_rand=pr_n.astype(o.dtype), out=out)
            d = RR.terminations(st["qpos"], st["time"], out["flags"])
            if d.any():
                o.reset(st, rq_n, rv_n, rt_n, done=d.astype(np.uint8), params=params, out=out)
            f = {"qpos": st["qpos"], "qvel": st["qvel"], "xquat": out["xquat"], "site_xpos": out["site_xpos"], "sensordata": out["sensordata"], "done": d}
            to[name] = (RR.step_terms(f, lb, rb), d, f)
        t64, d64, f64 = to["f64"]
        # ---- terminations: equal to fp64 unless the state sits at a threshold
        q64 = f64["qpos"]
        for arm, d_arm in (("cuda", done_c), ("f32", to["f32"][1])):
            flips[arm]["done"] += int((d_arm != d64).sum())
        both = ~done_c & ~d64  # environments stepped (not re-initialised) in both arms: their reward terms are comparable
        # ---- flags: equality with fp64; a flip is "explained" when the thresholded quantity is within the fp32 floor of its threshold
        margin = {"single_foot_contact": np.minimum(np.abs(f64["sensordata"][:, 10] - 0.1), np.abs(f64["sensordata"][:, 11] - 0.1)),
                  "left_contact": np.abs(f64["sensordata"][:, 10] - 0.1), "right_contact": np.abs(f64["sensordata"][:, 11] - 0.1),
                  "feet_too_close": np.abs(t64["feet_dist"] - 0.12), "stay_alive": np.zeros(n)}
        floor = {"single_foot_contact": 0.5, "left_contact": 0.5, "right_contact": 0.5, "feet_too_close": 1e-5, "stay_alive": 0.0}
        # (touch = summed normal force of a stiff contact, tens of newtons moving by ~1 N between fp32 evaluations of the capped solver:
        #  a 0.1 N threshold flips exactly when a foot is just touching down / lifting off)
        for k in flag_terms:
            for arm, ta in (("cuda", tc), ("f32", to["f32"][0])):
                fl = (np.asarray(ta[k], np.float64) != np.asarray(t64[k], np.float64)) & both
                flips[arm][k] += int(fl.sum())
                if arm == "cuda":
                    unexplained[k] += int((fl & (margin[k] > floor[k])).sum())
        # ---- smooth terms + the per-step total (without the stateful airtime term, compared separately below)
        for arm, ta in (("cuda", tc), ("f32", to["f32"][0]), ("f32", to["f32p"][0])):
            for k in cont_terms:
                err[arm][k].append(np.abs(np.asarray(ta[k], np.float64) - t64[k])[both])
            tot_a = sum(s * np.asarray(ta[k], np.float64) for k, (s, kind) in RR.SCALES.items() if kind != "air")
            tot_r = sum(s * np.asarray(t64[k], np.float64) for k, (s, kind) in RR.SCALES.items() if kind != "air")
            err[arm]["total"].append(np.abs(tot_a - tot_r)[both])
    eng_ctrl_dt = eng.ctrl_dt
    # ---- report + assertions
    steps = T_STEPS * n
    print(f"rollout rewards: {n} envs x {T_STEPS} control steps, {n_done} terminations ({n_done / n:.2f} per env)")
    assert n_done > 0.2 * n, "the rollout must exercise terminations + resets"
    bad = []
    for k in flag_terms + ("done",):
        fc_, fo_ = flips["cuda"][k], flips["f32"][k]
        print(f"  flag {k:22s} differs from fp64 on {fc_:5d} env-steps ({fc_ / steps:.2e}); fp32 oracle {fo_:5d}; cuda flips away from the threshold: {unexplained.get(k, 0)}")
        if fc_ > 2 * fo_ + max(5, int(3e-4 * steps)):
            bad.append(f"{k}: {fc_} flips vs fp32 oracle's {fo_}")
    for k in cont_terms + ("total",):
        ec, eo = np.concatenate(err["cuda"][k]), np.concatenate(err["f32"][k])
        pc, po = np.percentile(ec, [50, 99, 100]), np.percentile(eo, [50, 99, 100])
        print(f"  term {k:26s} |cuda - f64| p50 {pc[0]:.2e} p99 {pc[1]:.2e} max {pc[2]:.2e} | fp32 oracles p50 {po[0]:.2e} p99 {po[1]:.2e} max {po[2]:.2e}")
        for i, name in enumerate(("p50", "p99", "max")):
            if pc[i] > max(2e-5, 3 * po[i]):
                bad.append(f"{k}: {name} {pc[i]:.3e} > 3 x fp32 oracles' {po[i]:.3e}")
    assert not bad, bad
    # ---- the kernel's own reward inputs == the reference's formulas on the kernel's exported fields (flags bit for bit)
    for t, (tc, (rt_e, feet_e)) in enumerate(zip(traj_terms, traj_epi)):
        edge = (np.abs(tc["feet_dist"] - 0.12) < 1e-5)
        assert np.array_equal(rt_e[:, 0], tc["single_foot_contact"]), t
        assert np.array_equal(rt_e[:, 4] > 0.5, tc["left_contact"]) and np.array_equal(rt_e[:, 5] > 0.5, tc["right_contact"]), t
        assert np.array_equal(rt_e[~edge, 2], tc["feet_too_close"][~edge]), t
        np.testing.assert_allclose(rt_e[:, 1], tc["feet_orientation"], rtol=0, atol=1e-5)
    # FeetAirtimeReward over the chunk: zb_feet_airtime_f32 on the epilogue's contact flags == the transcription (exact)
    terms_d = torch.tensor(np.stack([e[0] for e in traj_epi]), device=dev)  # [T, n, 6]
    done_tn = torch.tensor(np.stack(traj_done).astype(np.uint8), device=dev)
    carry = torch.zeros((n, 2), dtype=torch.float32, device=dev)
    air = eng.feet_airtime_reward(terms_d[:, :, 4], terms_d[:, :, 5], done_tn, carry, ctrl_dt=eng_ctrl_dt, touchdown_penalty=0.3).cpu().numpy()
    lc = np.stack([x["left_contact"] for x in traj_terms]); rc = np.stack([x["right_contact"] for x in traj_terms])
    air_ref, carry_ref = RR.feet_airtime(lc, rc, np.stack(traj_done), np.zeros((n, 2), np.float32), eng_ctrl_dt, 0.3)
    assert np.array_equal(air, air_ref) and np.array_equal(carry.cpu().numpy(), carry_ref)
    all_terms = {k: np.stack([np.asarray(x[k], np.float64) for x in traj_terms]) for k in RR.SCALES if k != "feet_airtime"}
    all_terms["feet_airtime"] = air.astype(np.float64)
    ret = RR.total_reward(all_terms).sum(0)
    print(f"  return over {T_STEPS} steps (reference scales): mean {ret.mean():.3f} min {ret.min():.3f} max {ret.max():.3f}; airtime term mean {air.sum(0).mean():.4f}")
    assert np.isfinite(ret).all()
    eng.close()
