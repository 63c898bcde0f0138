"""Diagnostic (not a test): CUDA vs oracle error statistics for several solver budgets.  python scripts/gpu_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import ksim_gym_zbot_b200
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.engine import ZbotCudaEngine, EnvParams
import zb_oracle
from helpers import random_states, random_actions

m = M.build_zbot_syn()
n = 256
q, v = random_states(m, n, 12)
for iters, ls, tol in [(0, 8, 1e-8), (1, 8, 1e-8), (2, 8, 1e-8), (8, 8, 1e-8), (100, 50, 1e-12)]:
    opts = M.make_opts(dt=0.001, iterations=iters, ls_iterations=ls, tolerance=tol)
    eng = ZbotCudaEngine(m, n, opts=opts)
    ps = eng.reset(reset_qpos=torch.tensor(q), reset_qvel=torch.tensor(v), reset_t_rem=torch.full((n,), 1e9))
    torch.cuda.synchronize()
    c = {k: getattr(ps.data, k).cpu().numpy() for k in ("qacc", "qfrc_constraint", "efc_force", "solver_niter", "sensordata", "contact_mask")}
    res = {}
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        o = zb_oracle.Oracle(m, opts, dt, threads=8)
        st, out = o.new_state(n), o.new_out(n)
        o.reset(st, q, v, np.full(n, 1e9), out=out)
        res[name] = out
    def err(a, b, floor=10.0):
        return np.abs(a - b) / np.maximum(floor, np.abs(b))
    ec = err(c["qacc"], res["f64"]["qacc"]).max(axis=1)
    eo = err(res["f32"]["qacc"], res["f64"]["qacc"]).max(axis=1)
    print(f"iters={iters:3d}: qacc err cuda-f64 max {ec.max():.2e} median {np.median(ec):.2e} p90 {np.percentile(ec,90):.2e} | f32-f64 max {eo.max():.2e} median {np.median(eo):.2e} p90 {np.percentile(eo,90):.2e} | niter cuda {c['solver_niter'].mean():.2f} f32 {res['f32']['solver_niter'].mean():.2f} f64 {res['f64']['solver_niter'].mean():.2f}")
    ef = err(c["efc_force"], res["f64"]["efc_force"], 1.0).max(axis=1)
    print(f"          efc_force err max {ef.max():.2e} median {np.median(ef):.2e}; sensordata err {err(c['sensordata'], res['f64']['sensordata']).max():.2e}; mask equal {(c['contact_mask']==res['f64']['contact_mask']).mean():.3f}")
    eng.close()
