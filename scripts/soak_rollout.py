#!/usr/bin/env python
"""Soak of the device-resident rollout: tensor-core actor + physics (pushes, randomisation, in-launch terminations / resets), chunks of 200
control steps replayed as one CUDA graph each.  Checks after every chunk that no non-finite value survives in the state, the actor carry or
the trajectory, and that the chunk's reward list is finite."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ksim_gym_zbot_b200  # noqa
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.engine import EnvParams, PushEventConfig, ZbotCudaEngine
from ksim_gym_zbot_b200.policy import ActorParams, ActorPolicy, PRECISION_TF32X3
from ksim_gym_zbot_b200.rollout import Rollout

n, T, chunks = int(os.environ.get("ENVS", 16384)), 200, int(os.environ.get("CHUNKS", 10))
model = M.build_zbot_syn()
eng = ZbotCudaEngine(model, n, device=0, opts=M.make_opts(dt=0.005), ctrl_dt=0.02, push=PushEventConfig(), seed=3)
ep = EnvParams(n, eng.device, **M.randomize_env_params(model, n, 5))
ro = Rollout(eng, T, ep)
prm = ActorParams.random(seed=4)
prm.arrays["output_proj.weight"] *= 3.0
pol = ActorPolicy(prm, n, eng.device, precision=PRECISION_TF32X3)
ro.reset()
t0 = time.time()
resets = 0
for c in range(chunks):
    ro.sample_noise() if ro.noise_u is not None else None
    ro.sample_pushes() if ro.pushes is not None else None
    tr = ro.unroll_actor(pol, argmax=False)
    torch.cuda.synchronize()
    for name in ("qpos", "qvel", "action", "actor_obs", "critic_obs", "reward_terms"):
        assert torch.isfinite(getattr(tr, name)).all(), (c, name)
    assert torch.isfinite(pol.carry).all() and torch.isfinite(ro.ps.data.qpos).all(), c
    resets += int((tr.done != 0).sum())
    total, _ = ro.rewards()
    assert torch.isfinite(total).all()
    ro.next_chunk()
dt = time.time() - t0
print(f"soak: {n} envs x {chunks * T} control steps x 4 substeps = {n * chunks * T * 4 / 1e6:.0f} M env-steps in {dt:.1f} s wall "
      f"({n * chunks * T * 4 / dt:.3e} env-steps/s incl. the checks), {resets} episode ends, mean reward {float(total.mean()):.3f}, all finite")
