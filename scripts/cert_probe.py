#!/usr/bin/env python
"""Line-search certificate (tests/parity_lib.py) of the default library and of the mutation builds on several state populations:
python scripts/cert_probe.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ksim_gym_zbot_b200  # noqa: E402,F401
import parity_lib as P  # noqa: E402
from helpers import random_actions, random_states  # noqa: E402
from ksim_gym_zbot_b200 import build as B  # noqa: E402
from ksim_gym_zbot_b200 import model as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = M.build_zbot_syn()
libs = {"default": None, **({k: B.variant_path(k) for k in B.MUTANTS} if "--mutants" in sys.argv else {})}
for name, dt, S in (("rollout dt=0.005", 0.005, 4), ("rollout dt=0.001", 0.001, 20), ("limits", 0.002, 1), ("random warm start", 0.005, 4), ("after reset", 0.005, 4), ("flat feet", 0.005, 1)):
    opts = M.make_opts(dt=dt)
    if name == "limits":
        q, v = random_states(m, n, 101)
        rng = np.random.default_rng(102)
        lo, hi = m.field("jnt_range")[6:, 0], m.field("jnt_range")[6:, 1]
        for i in range(n):
            for j in rng.choice(20, size=3, replace=False):
                q[i, 7 + j] = (lo[j] - rng.uniform(0.01, 0.1)) if rng.random() < 0.5 else (hi[j] + rng.uniform(0.01, 0.1))
        st = {"qpos": q, "qvel": v, "qacc_warmstart": np.zeros((n, 26), np.float32), "act_pos": q[:, 7:].copy(), "act_vel": v[:, 6:].copy(),
              "act_torque": np.zeros((n, 20), np.float32), "push_t_rem": np.full(n, 1e9, np.float32), "time": np.zeros(n, np.float32)}
    elif name == "flat feet":
        q, v = random_states(m, n, 103, joint_noise=2e-3, z_jitter=1.5e-3, tilt=2e-3, vel=0.3)
        st = {"qpos": q, "qvel": v, "qacc_warmstart": np.zeros((n, 26), np.float32), "act_pos": q[:, 7:].copy(), "act_vel": v[:, 6:].copy(),
              "act_torque": np.zeros((n, 20), np.float32), "push_t_rem": np.full(n, 1e9, np.float32), "time": np.zeros(n, np.float32)}
    else:
        st = P.rollout_states(m, opts, n, 2 if name == "after reset" else 25, S, seed=3)
        if name.startswith("random"):
            st["qacc_warmstart"] = (st["qacc_warmstart"] + np.random.default_rng(5).uniform(-20, 20, (n, 26))).astype(np.float32)
    act = random_actions(n, 5)
    for lname, lib in libs.items():
        try:
            rep = P.compare_single_step(m, opts, st, act, lib_path=lib, with_oracles=False)
        except AssertionError as e:
            print(f"{name:22s} {lname:8s} -> {e}")
            continue
        r = P.linesearch_certificate(m, opts, st, act, rep["trace"], rep["out"])
        print(f"{name:22s} {lname:8s} line searches {r['n_ls']} points {r['n_points']} ambiguous {r['n_ambiguous']} geometry-ambiguous envs {r['n_geometry']} FAIL {len(r['fail'])}",
              r["fail"][:2])
        if lname == "default":
            print("      contacts per env:", np.bincount(np.array([bin(int(x)).count("1") for x in rep["out"]["contact_mask"]]), minlength=9))
            for k, e in r["dist"].items():
                print(f"      {k}: p50 {np.percentile(e, 50):.2e} p99 {np.percentile(e, 99):.2e} p99.9 {np.percentile(e, 99.9):.2e} max {e.max():.2e}")
