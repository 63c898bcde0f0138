"""Regenerates the committed golden fixtures from the CPU oracle.

These are RESTATEMENT goldens (label per SURVEY.md §8c(5)): the reference's MJX cannot be imported here
(jax / mujoco / mujoco-mjx / ksim are not installed, no network), so the vectors pin the oracle against
regressions and give the CUDA path a fixed target; they do not pin the oracle to MJX.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import ksim_gym_zbot_b200  # noqa: E402,F401
import zb_oracle  # noqa: E402
from helpers import random_actions, random_states  # noqa: E402
from ksim_gym_zbot_b200 import model as M  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def trace_1000(m):
    """BASELINE.json config 1: 1 env, 1000 physics steps (dt=0.001) from the default pose, bias-pose action."""
    out = {}
    for name, dt in (("f64", np.float64), ("f32", np.float32)):
        o = zb_oracle.Oracle(m, M.make_opts(dt=0.001), dt)
        st = o.new_state(1)
        o.reset(st, M.default_qpos(m)[None].astype(np.float32), np.zeros((1, 26)), np.array([1e9]))
        qs, vs = [], []
        for t in range(50):
            o.step(st, M.BIAS_POSE[None].astype(np.float32), 20)
            qs.append(st["qpos"][0].copy())
            vs.append(st["qvel"][0].copy())
        out["qpos_" + name], out["qvel_" + name] = np.stack(qs), np.stack(vs)
    np.savez_compressed(os.path.join(HERE, "zbot_syn_1000.npz"), **out)


def forward_cases(m):
    """16 random states: mjx.forward outputs + one mjx.step (fp64 oracle)."""
    n = 16
    q, v = random_states(m, n, 1234)
    act = random_actions(n, 99)
    params = M.randomize_env_params(m, n, 77)
    o = zb_oracle.Oracle(m, M.make_opts(dt=0.005), np.float64)
    st, out = o.new_state(n), o.new_out(n)
    o.reset(st, q, v, np.full(n, 1e9), params=params, out=out)
    g = {"in_qpos": q, "in_qvel": v, "in_action": act}
    g.update({"p_" + k: a for k, a in params.items()})
    g.update({"fwd_" + k: a.copy() for k, a in out.items()})
    o.step(st, act, 4, params=params, out=out)
    g.update({"step_" + k: a.copy() for k, a in st.items()})
    g.update({"stepout_" + k: out[k].copy() for k in ("sensordata", "contact_mask", "qacc", "actuator_force")})
    np.savez_compressed(os.path.join(HERE, "zbot_syn_forward16.npz"), **g)


if __name__ == "__main__":
    m = M.build_zbot_syn()
    trace_1000(m)
    forward_cases(m)
    print("golden fixtures written to", HERE)
