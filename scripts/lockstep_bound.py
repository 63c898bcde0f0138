#!/usr/bin/env python
"""Upper bound of what grouping similar environments into a lockstep group could gain: the step kernel on a batch whose environments
are all DIFFERENT (bench workload) vs. a batch where the 28 environments of every CTA are copies of ONE environment of that batch
(identical control flow inside a CTA: no barrier wait, perfect instruction sharing), same total work."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ksim_gym_zbot_b200  # noqa
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.engine import ZbotCudaEngine

n, S = 4096, 4
model = M.build_zbot_syn()
opts = M.make_opts(dt=0.005)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(17)
bias = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=dev)
K, W = 60, 40
actions = bias + 0.3 * (torch.rand((K + W, n, M.NU), generator=g, device=dev) * 2 - 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for mode in ("distinct", "cta-uniform"):
    eng = ZbotCudaEngine(model, n, device=0, opts=opts, ctrl_dt=0.02, seed=0)
    rq, rv, rt = eng.sample_reset(n)
    idx = torch.arange(n, device=dev)
    if mode == "cta-uniform":
        idx = (idx // 28) * 28  # every CTA (28 consecutive environments) runs 28 copies of its first environment
        rq, rv, rt = rq[idx].contiguous(), rv[idx].contiguous(), rt[idx].contiguous()
    ps = eng.reset(None, reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt)
    done = torch.zeros((n,), dtype=torch.uint8, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    for t in range(K + W):
        a = actions[t][idx].contiguous()
        if t >= W:
            flush.zero_(); ev[t - W][0].record()
        eng.step(a, None, ps, auto_reset=(rq, rv, rt), done_out=done)
        if t >= W:
            ev[t - W][1].record()
    torch.cuda.synchronize()
    ms = np.array([e[0].elapsed_time(e[1]) for e in ev])
    print(mode, "kernel ms mean %.4f median %.4f niter %.2f" % (ms.mean(), np.median(ms), float(ps.data.solver_niter.float().mean())), flush=True)
    eng.close()
