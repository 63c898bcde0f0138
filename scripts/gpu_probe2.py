"""Diagnostic: which per-env parameter / stage produces deterministic (iterations=0) differences."""
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
q, v = random_states(m, n, 14)
allp = M.randomize_env_params(m, n, 4)
def run(params, iters=0, S=1, label=""):
    opts = M.make_opts(dt=0.001, iterations=iters)
    eng = ZbotCudaEngine(m, n, opts=opts, ctrl_dt=0.001*S)
    ep = EnvParams(n, eng.device, **params) if params else None
    ps = eng.reset(ep, reset_qpos=torch.tensor(q), reset_qvel=torch.tensor(v), reset_t_rem=torch.full((n,), 1e9))
    c0 = {k: getattr(ps.data, k).cpu().numpy().copy() for k in ("qacc", "cinert", "cvel", "xpos", "sensordata", "efc_force", "qfrc_constraint")}
    act = random_actions(n, 5)
    eng.step(torch.tensor(act, device="cuda"), ep, ps)
    c1 = {k: getattr(ps.data, k).cpu().numpy().copy() for k in ("qpos", "qvel", "qacc", "ctrl", "actuator_force")}
    c1["act_torque"] = ps.actuator_state.last_computed_torque.cpu().numpy()
    out = {}
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        o = zb_oracle.Oracle(m, opts, dt, threads=8)
        st, oo = o.new_state(n), o.new_out(n)
        o.reset(st, q, v, np.full(n, 1e9), params=params, out=oo)
        r0 = {k: oo[k].copy() for k in c0}
        o.step(st, act, S, params=params, out=oo)
        r1 = {"qpos": st["qpos"].copy(), "qvel": st["qvel"].copy(), "qacc": oo["qacc"].copy(), "ctrl": oo["ctrl"].copy(), "actuator_force": oo["actuator_force"].copy(), "act_torque": st["act_torque"].copy()}
        out[name] = (r0, r1)
    def e(a, b, fl=1.0): return float(np.max(np.abs(a - b) / np.maximum(fl, np.abs(b))))
    s = f"{label:18s}"
    for k in c0: s += f" {k}:{e(c0[k], out['f64'][0][k], 10.0 if k in ('qacc','sensordata') else 1.0):.1e}/{e(out['f32'][0][k], out['f64'][0][k], 10.0 if k in ('qacc','sensordata') else 1.0):.1e}"
    print(s)
    s = f"{'   after step':18s}"
    for k in c1: s += f" {k}:{e(c1[k], out['f64'][1][k], 10.0 if k=='qacc' else 1.0):.1e}/{e(out['f32'][1][k], out['f64'][1][k], 10.0 if k=='qacc' else 1.0):.1e}"
    print(s)
    eng.close()
run(None, label="no params")
for k in allp:
    run({k: allp[k]}, label=k)
run(allp, label="all")
run(None, iters=8, label="iters8 noparam")
run(allp, iters=8, label="iters8 all")
