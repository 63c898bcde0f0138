"""Diagnostic: per-iteration solver trace (alpha, cost, ls_iters, |search|) CUDA vs oracle f32/f64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import ksim_gym_zbot_b200
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.engine import ZbotCudaEngine
import zb_oracle
from helpers import random_states
np.set_printoptions(linewidth=200, precision=6, suppress=False)
m = M.build_zbot_syn()
n = 1024
q, v = random_states(m, n, 12)
opts = M.make_opts(dt=0.001, iterations=8)
eng = ZbotCudaEngine(m, n, opts=opts)
eng.trace = torch.zeros((n, 160), device="cuda")
ps = eng.reset(reset_qpos=torch.tensor(q), reset_qvel=torch.tensor(v), reset_t_rem=torch.full((n,), 1e9))
traw = eng.trace.cpu().numpy(); tc = traw[:, :64].reshape(n, 16, 4)
qc = ps.data.qacc.cpu().numpy()
res = {}
for name, dt in (("f32", np.float32), ("f64", np.float64)):
    o = zb_oracle.Oracle(m, opts, dt, threads=8)
    st, out = o.new_state(n), o.new_out(n)
    o.reset(st, q, v, np.full(n, 1e9), out=out)
    res[name] = (out["solver_trace"][:, :64].reshape(n, 16, 4).astype(np.float64), out["qacc"].astype(np.float64), out["solver_trace"].astype(np.float64))
t32, t64 = res["f32"][0], res["f64"][0]
for it in range(8):
    # relative cost gap to f64 after iteration it
    gc = (tc[:, it, 1] - t64[:, it, 1]) / np.maximum(1e-9, np.abs(t64[:, it, 1]))
    go = (t32[:, it, 1] - t64[:, it, 1]) / np.maximum(1e-9, np.abs(t64[:, it, 1]))
    ac = np.abs(tc[:, it, 0] - t64[:, it, 0]) / np.maximum(1e-12, np.abs(t64[:, it, 0]))
    ao = np.abs(t32[:, it, 0] - t64[:, it, 0]) / np.maximum(1e-12, np.abs(t64[:, it, 0]))
    print(f"iter {it}: cost rel gap vs f64: cuda median {np.median(gc):+.2e} p90 {np.percentile(gc,90):+.2e} | f32 median {np.median(go):+.2e} p90 {np.percentile(go,90):+.2e} || alpha rel err cuda median {np.median(ac):.2e} p90 {np.percentile(ac,90):.2e} | f32 median {np.median(ao):.2e} p90 {np.percentile(ao,90):.2e} || ls iters cuda {tc[:,it,2].mean():.2f} f32 {t32[:,it,2].mean():.2f} f64 {t64[:,it,2].mean():.2f}")
e = np.abs(qc - res["f64"][1]).max(axis=1)
worst = np.argsort(-e)[:3]
for i in worst:
    print("env", i, "qacc err", e[i])
    print(" cuda", tc[i, :8].T)
    print(" f32 ", t32[i, :8].T)
    print(" f64 ", t64[i, :8].T)

print("---- flip statistics")
for it in range(3):
    ac = np.abs(tc[:, it, 0] - t64[:, it, 0]) / np.maximum(1e-12, np.abs(t64[:, it, 0]))
    ao = np.abs(t32[:, it, 0] - t64[:, it, 0]) / np.maximum(1e-12, np.abs(t64[:, it, 0]))
    print(f"iter {it}: frac(alpha rel err > 1e-3): cuda {(ac>1e-3).mean():.3f} f32 {(ao>1e-3).mean():.3f}; p75 cuda {np.percentile(ac,75):.2e} f32 {np.percentile(ao,75):.2e}")
same_c = (tc[:, :8, 2] == t64[:, :8, 2]).all(axis=1)
same_o = (t32[:, :8, 2] == t64[:, :8, 2]).all(axis=1)
ec = (np.abs(qc - res["f64"][1]) / np.maximum(10, np.abs(res["f64"][1]))).max(axis=1)
eo = (np.abs(res["f32"][1] - res["f64"][1]) / np.maximum(10, np.abs(res["f64"][1]))).max(axis=1)
print(f"envs with identical ls-iteration sequence as f64: cuda {same_c.mean():.3f}, f32 {same_o.mean():.3f}")
for nm, sm, ee in (("cuda", same_c, ec), ("f32", same_o, eo)):
    if sm.any():
        print(f"  {nm}: qacc err on those: p50 {np.percentile(ee[sm],50):.2e} p90 {np.percentile(ee[sm],90):.2e} max {ee[sm].max():.2e}; on the others p50 {np.percentile(ee[~sm],50):.2e}")

print("---- first line search, envs where cuda deviates at iter 0 but f32 does not")
ac = np.abs(tc[:, 0, 0] - t64[:, 0, 0]) / np.maximum(1e-12, np.abs(t64[:, 0, 0]))
ao = np.abs(t32[:, 0, 0] - t64[:, 0, 0]) / np.maximum(1e-12, np.abs(t64[:, 0, 0]))
idx = np.where((ac > 1e-3) & (ao < 1e-5))[0][:4]
for i in idx:
    print("env", i, "alpha cuda/f32/f64", tc[i,0,0], t32[i,0,0], t64[i,0,0], "ls iters", tc[i,0,2], t32[i,0,2], t64[i,0,2])
    for nm, arr in (("cuda", traw[i]), ("f32", res["f32"][2][i]), ("f64", res["f64"][2][i])):
        pts = arr[32:32+3*11].reshape(-1, 3)
        print("  ", nm, " ".join(f"({a:.6g},{d0:.4g},{d1:.4g})" for a, d0, d1 in pts))
