"""The rollout loop + Trajectory writer (ksim-gym-zbot_b200/rollout.py) on the GPU:

  * the trajectory the fused launches write (state-like fields = the state BEFORE the done -> reset select, obs-like fields one slot
    ahead) equals, bit for bit, what a step-by-step loop of separate calls records: step -> record -> terminations -> reset(done);
  * the whole chunk replayed as ONE CUDA graph gives the same bits as issuing the launches one by one;
  * the reference's reward list evaluated on that trajectory by the product code (rewards.py: torch ops + the kernels' reward inputs +
    zb_feet_airtime_f32) equals the numpy transcription of train.py:1546-1586 (tests/ref_rewards.py).
"""
import numpy as np
import pytest

import ref_rewards as RR

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _setup(zmodel, n, T, push, seed=3):
    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import EnvParams, PushEventConfig, ZbotCudaEngine
    from ksim_gym_zbot_b200.rollout import Rollout

    opts = M.make_opts(dt=0.005)
    cfg = PushEventConfig(x_linvel=1.0, y_linvel=1.0, z_linvel=0.4, interval_range=(0.05, 0.2)) if push else None
    eng = ZbotCudaEngine(zmodel, n, device=0, opts=opts, ctrl_dt=0.02, push=cfg, seed=seed)
    ep = EnvParams(n, eng.device, **M.randomize_env_params(zmodel, n, 11))
    ro = Rollout(eng, T, ep)
    g = torch.Generator(device=eng.device)
    g.manual_seed(seed)
    bias = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=eng.device)
    ro.traj.action.copy_(bias + 0.4 * (torch.rand((T, n, M.NU), generator=g, device=eng.device) * 2 - 1))
    pushes = torch.stack([eng.sample_push(n) for _ in range(T)]) if push else None
    return eng, ep, ro, pushes


def test_trajectory_equals_step_by_step_loop(zmodel):
    from ksim_gym_zbot_b200.engine import ZbotCudaEngine

    n, T = 512, 24
    eng, ep, ro, pushes = _setup(zmodel, n, T, push=True)
    ro.reset()
    tr = ro.unroll(pushes=pushes)
    torch.cuda.synchronize()
    assert int((tr.done != 0).sum()) > n // 10, "the chunk must contain terminations"
    # ---- the same chunk with separate calls: step (no auto-reset) -> record -> terminations -> reset(done)
    eng2 = ZbotCudaEngine(zmodel, n, device=0, opts=eng.opts, ctrl_dt=0.02, push=eng.push, seed=99)
    obs = eng2.new_observations(n, rewards=True)
    ps = eng2.reset(ep, reset_qpos=ro.rq, reset_qvel=ro.rv, reset_t_rem=ro.rt, obs=obs, reset_imu_lag=ro.reset_imu_lag)
    rec = {k: [] for k in ("qpos", "qvel", "xpos", "xquat", "done", "actor_obs", "critic_obs", "feet_pos_obs", "reward_terms", "sensordata")}

    def obs_like():
        for k in ("actor_obs", "critic_obs", "feet_pos_obs", "reward_terms"):
            rec[k].append(getattr(obs, k).clone())
        rec["sensordata"].append(ps.data.sensordata.clone())

    obs_like()
    for t in range(T):
        eng2.step(tr.action[t], ep, ps, push_rand=pushes[t], obs=obs)
        for k in ("qpos", "qvel", "xpos", "xquat"):
            rec[k].append(getattr(ps.data, k).clone())
        done = eng2.terminations(ps)
        rec["done"].append(done.clone())
        eng2.reset(ep, physics_state=ps, done=done, reset_qpos=ro.rq, reset_qvel=ro.rv, reset_t_rem=ro.rt, obs=obs, reset_imu_lag=ro.reset_imu_lag)
        obs_like()
    torch.cuda.synchronize()
    for k, v in rec.items():
        a, b = getattr(tr, k).cpu().numpy(), torch.stack(v).cpu().numpy()
        assert a.shape == b.shape, k
        assert np.array_equal(a, b), f"trajectory.{k} differs from the step-by-step loop (max |d| {np.abs(a.astype(np.float64) - b).max():.3e})"
    eng.close(); eng2.close()


def test_chunk_graph_replay_and_reference_rewards(zmodel):
    n, T = 384, 50
    eng, ep, ro, _ = _setup(zmodel, n, T, push=False, seed=8)
    ro.reset()
    first = {k: getattr(ro.traj, k)[0].clone() for k in ro.traj.OBS_LIKE}
    ro.unroll()  # captures the T launches into one graph (capture also runs nothing: replay to execute)
    ro.reset()
    ro.unroll()  # replay
    torch.cuda.synchronize()
    tr = ro.traj
    snap = {k: getattr(tr, k).clone() for k in ("qpos", "qvel", "xpos", "xquat", "done", "actor_obs", "critic_obs", "reward_terms", "sensordata")}
    for k, v in first.items():
        assert torch.equal(getattr(tr, k)[0], v), k  # the reset launch's slot is untouched by the chunk
    # launches issued one by one (policy path) from the same reset: identical bits
    ro.reset()
    ro.unroll(policy=lambda o, it=iter(range(T)): tr.action[next(it)])
    torch.cuda.synchronize()
    for k, v in snap.items():
        assert torch.equal(getattr(tr, k), v), f"graph replay and step-by-step launches differ in trajectory.{k}"
    assert torch.isfinite(tr.qpos).all() and float(tr.actor_obs.abs().max()) > 0
    # ---- rewards: product code vs the numpy transcription of the reference's list
    total, terms = ro.rewards()
    torch.cuda.synchronize()
    lb, rb = int(zmodel.site_body[1]), int(zmodel.site_body[2])
    q, v, xq = tr.qpos.cpu().numpy(), tr.qvel.cpu().numpy(), tr.xquat.cpu().numpy()
    sd, feet, dn = tr.sensordata.cpu().numpy()[:T], tr.feet_pos_obs.cpu().numpy()[:T], tr.done.cpu().numpy() != 0
    ref = {k: [] for k in RR.SCALES}
    for t in range(T):
        f = {"qpos": q[t], "qvel": v[t], "xquat": xq[t], "site_xpos": np.zeros((n, 3, 3), np.float32), "sensordata": sd[t], "done": dn[t]}
        st = RR.step_terms(f, lb, rb)
        st["feet_too_close"] = (np.linalg.norm(feet[t][:, :3] - feet[t][:, 3:], axis=-1) < 0.12).astype(np.float32)  # obs-like: slot t
        for k in ref:
            if k != "feet_airtime":
                ref[k].append(st[k])
    lc, rc = sd[:, :, 10] > 0.1, sd[:, :, 11] > 0.1
    ref["feet_airtime"], carry = RR.feet_airtime(lc, rc, dn, np.zeros((n, 2), np.float32), 0.02, 0.3)
    for k in RR.SCALES:
        a, b = terms[k].cpu().numpy(), np.asarray(ref[k], np.float64).reshape(T, n)
        if k == "feet_too_close":
            edge = np.abs(np.linalg.norm(feet[:, :, :3] - feet[:, :, 3:], axis=-1) - 0.12) < 1e-6
            assert np.array_equal(a[~edge], b[~edge]), k
        elif k in ("stay_alive", "single_foot_contact", "feet_airtime"):
            assert np.array_equal(a, b.astype(np.float32)), k
        else:
            np.testing.assert_allclose(a, b, rtol=0, atol=3e-6, err_msg=k)
    assert np.array_equal(ro.airtime_carry.cpu().numpy(), carry)
    tot_ref = RR.total_reward({k: np.asarray(ref[k], np.float64).reshape(T, n) for k in RR.SCALES})
    np.testing.assert_allclose(total.cpu().numpy(), tot_ref, rtol=0, atol=3e-5)
    assert set(terms) == set(RR.SCALES) and dn.sum() >= 0
    # next chunk continues from the last observation slot
    last = tr.actor_obs[T].clone()
    ro.next_chunk()
    assert torch.equal(tr.actor_obs[0], last)
    eng.close()


def test_actor_in_the_loop_rollout(zmodel):
    """unroll_actor (zb_actor_f32 + physics per control step, the chunk as ONE CUDA graph) == the same chunk driven step by step with
    the actions the CPU restatement of Actor.forward computes from the observations the kernels wrote (closed loop: every step's action
    feeds the next state), incl. carry restarts after terminations and a second chunk."""
    import zb_policy_oracle as PO
    from ksim_gym_zbot_b200.policy import ActorParams, ActorPolicy

    n, T = 256, 30
    eng, ep, ro, _ = _setup(zmodel, n, T, push=True, seed=5)  # strong pushes: environments fall inside the chunk (ro.pushes, sampled once)
    prm = ActorParams.random(seed=9)
    prm.arrays["output_proj.weight"] *= 4.0  # livelier actions
    pol = ActorPolicy(prm, n, eng.device)
    ro.reset()
    ro.unroll_actor(pol, argmax=True)           # capture + first replay
    torch.cuda.synchronize()
    tr = ro.traj
    g1 = {k: getattr(tr, k).clone() for k in ("qpos", "qvel", "done", "action", "actor_obs")}
    carry1 = pol.carry.clone()
    # ---- the same chunk launch by launch (graph=False): identical bits
    ro.reset(); pol.reset_carry(); ro.prev_done.zero_()
    ro.unroll_actor(pol, argmax=True, graph=False)
    torch.cuda.synchronize()
    for k, v in g1.items():
        assert torch.equal(getattr(tr, k), v), f"graph vs launch-by-launch: trajectory.{k}"
    assert torch.equal(pol.carry, carry1)
    # ---- teacher-forced check of every actor launch against the oracle: action[t] from (actor_obs[t], carry, done[t-1])
    obs, act, done = tr.actor_obs.cpu().numpy(), tr.action.cpu().numpy(), tr.done.cpu().numpy()
    carry = np.zeros((n, prm.depth, 128))
    for t in range(T):
        if t > 0:
            carry[done[t - 1] != 0] = 0.0
        mean, std, logits, carry = PO.actor_forward(prm.oracle_dict(), obs[t], carry)
        np.testing.assert_allclose(act[t], PO.select_action(mean, std, logits), rtol=0, atol=2e-4, err_msg=f"step {t}")
    np.testing.assert_allclose(pol.carry.cpu().numpy(), carry, rtol=0, atol=5e-5)
    assert int((done != 0).sum()) > 0, "no termination inside the chunk: the carry restart is not exercised"
    # ---- second chunk: continues from the last observation slot, prev_done carries the restart across the chunk boundary
    last_done = tr.done[T - 1].clone()
    ro.next_chunk()
    ro.unroll_actor(pol, argmax=True)
    torch.cuda.synchronize()
    assert torch.equal(ro.prev_done, tr.done[T - 1]) and last_done.shape == ro.prev_done.shape
    carry[last_done.cpu().numpy() != 0] = 0.0
    mean, std, logits, _ = PO.actor_forward(prm.oracle_dict(), tr.actor_obs[0].cpu().numpy(), carry)
    np.testing.assert_allclose(tr.action[0].cpu().numpy(), PO.select_action(mean, std, logits), rtol=0, atol=2e-4)
    # ---- sampled actions: the chunk's noise buffers are consumed (different numbers -> different actions), finite states
    ro.reset(); pol.reset_carry(); ro.prev_done.zero_()
    ro.unroll_actor(pol, argmax=False)
    torch.cuda.synchronize()
    a1 = tr.action.clone()
    ro.reset(); pol.reset_carry(); ro.prev_done.zero_(); ro.sample_noise()
    ro.unroll_actor(pol, argmax=False)
    torch.cuda.synchronize()
    assert not torch.equal(a1, tr.action) and torch.isfinite(tr.qpos).all()
    eng.close()
