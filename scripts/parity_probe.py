#!/usr/bin/env python
"""Prints the distributions behind tests/test_parity_tail.py (path flip rates, path-conditional maxima, caps) for a batch of
rollout states: python scripts/parity_probe.py [n] [lib_path]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import ksim_gym_zbot_b200  # noqa: E402,F401
import parity_lib as P  # noqa: E402
from ksim_gym_zbot_b200 import model as M  # noqa: E402
from helpers import random_actions  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = sys.argv[2] if len(sys.argv) > 2 else None
m = M.build_zbot_syn()
for dt, S, rand in ((0.005, 4, False), (0.001, 20, False), (0.005, 4, True)):
    opts = M.make_opts(dt=dt)
    params = M.randomize_env_params(m, n, 7) if rand else None
    st = P.rollout_states(m, opts, n, 25, S, seed=3, params=params, lib_path=lib)
    rep = P.compare_single_step(m, opts, st, random_actions(n, 5), params=params, lib_path=lib)
    try:
        P.check_report(rep, tag=f"dt={dt} rand={rand}")
        print("   -> checks pass")
    except AssertionError as e:
        print("   -> FAIL:", e)
    mc = rep.get("path_match_cuda")
    if mc is not None:
        for k in ("qpos", "qvel", "qacc"):
            e = rep["d_cuda_f32_" + k]
            print(f"   {k}: matched p50 {np.percentile(e[mc],50):.2e} p99 {np.percentile(e[mc],99):.2e} max {e[mc].max():.2e}; "
                  f"unmatched ({(~mc).sum()}) p50 {np.percentile(e[~mc],50) if (~mc).any() else 0:.2e} max {e[~mc].max() if (~mc).any() else 0:.2e}")
        print("   niter mean cuda %.2f f32 %.2f" % (rep["niter_cuda"].mean(), rep["niter_f32"].mean()))
