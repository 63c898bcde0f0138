"""Soak run (not a pytest): many control steps of the randomised workload with pushes and in-launch resets; checks that the
state stays finite, that non-finite flags never survive a reset and reports throughput + solver statistics."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import ksim_gym_zbot_b200
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.engine import ZbotCudaEngine, EnvParams, PushEventConfig

n, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, int(sys.argv[2]) if len(sys.argv) > 2 else 3000
m = M.build_zbot_syn()
eng = ZbotCudaEngine(m, n, dt=0.005, ctrl_dt=0.02, push=PushEventConfig(x_linvel=1.0, y_linvel=1.0, z_linvel=0.5, interval_range=(0.2, 0.4)), seed=3)
ep = EnvParams(n, eng.device, **M.randomize_env_params(m, n, 5))
rq, rv, rt = eng.sample_reset(n)
ps = eng.reset(ep, reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt)
done = torch.zeros(n, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda"); g.manual_seed(0)
bias = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device="cuda")
tot_done = torch.zeros((), dtype=torch.int64, device="cuda")
bad = torch.zeros((), dtype=torch.int64, device="cuda")
niter = torch.zeros((), dtype=torch.float64, device="cuda")
t0 = time.time()
for t in range(steps):
    act = bias + 0.5 * (torch.rand((n, 20), generator=g, device="cuda") * 2 - 1)
    eng.step(act, ep, ps, auto_reset=(rq, rv, rt), done_out=done)
    tot_done += (done != 0).sum()
    bad += (~torch.isfinite(ps.data.qpos).all(dim=1)).sum() + (ps.data.flags != 0).sum()
    niter += ps.data.solver_niter.float().mean()
torch.cuda.synchronize()
dt = time.time() - t0
print(f"soak: {n} envs x {steps} control steps (S=4): {n*4*steps/dt:.3e} env-steps/s wall (incl. torch action sampling + checks), "
      f"resets/step {tot_done.item()/steps:.1f}, non-finite or flagged states after the in-launch reset: {bad.item()}, "
      f"mean solver iterations {niter.item()/steps:.2f}, max |qvel| {ps.data.qvel.abs().max().item():.2f}")
assert bad.item() == 0
