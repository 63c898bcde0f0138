"""Tiny end-to-end case for compute-sanitizer (memcheck / racecheck): reset, forward, step, auto-reset, Newton, trace."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import ksim_gym_zbot_b200
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.engine import ZbotCudaEngine, EnvParams, PushEventConfig
from helpers import random_states, random_actions

m = M.build_zbot_syn()
n = 37
for solver in (M.SOLVER_CG, M.SOLVER_NEWTON):
    eng = ZbotCudaEngine(m, n, opts=M.make_opts(dt=0.005, solver=solver), ctrl_dt=0.01, push=PushEventConfig(interval_range=(0.0, 0.02)))
    ep = EnvParams(n, eng.device, **M.randomize_env_params(m, n, 1))
    q, v = random_states(m, n, 2, tilt=1.0)
    ps = eng.reset(ep, reset_qpos=torch.tensor(q), reset_qvel=torch.tensor(v), reset_t_rem=torch.full((n,), 0.004))
    eng.forward(ep, ps)
    rq, rv, rt = eng.sample_reset(n)
    done = torch.zeros(n, dtype=torch.uint8, device="cuda")
    eng.trace = torch.zeros((n, 160), device="cuda")
    for t in range(2):
        eng.step(torch.tensor(random_actions(n, t), device="cuda"), ep, ps, auto_reset=(rq, rv, rt), done_out=done)
    eng.trace = None
    eng.step(torch.tensor(random_actions(n, 9), device="cuda"), ep, ps)
    eng.terminations(ps, done)
    eng.reset(ep, physics_state=ps, done=done, reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt)
    torch.cuda.synchronize()
    assert torch.isfinite(ps.data.qpos).all()
    eng.close()
print("sanitize case ok")
