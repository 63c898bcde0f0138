"""Quantiles of the follow-comparison errors (tests/parity_lib.follow_compare) of a library variant: kernel vs the fp32 oracle, both
against the fp64 oracle following the kernel's step sizes.  usage: ZBPHYS_LIB=... python scripts/follow_probe.py [iters]"""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import numpy as np
import ksim_gym_zbot_b200  # noqa
from ksim_gym_zbot_b200 import model as M
import parity_lib as P
import test_parity_tail as T

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
m = M.build_zbot_syn()
for name, opts, st, act, prm in T._populations(m, 4096):
    o = M.make_opts(dt=float(opts.dt), iterations=iters)
    r = P.follow_compare(m, o, st, act, params=prm)
    ok = ~r["geometry_ambiguous"]
    for k in ("qpos", "qvel", "qacc"):
        ec, eo = r["e_" + k][ok], r["e32_" + k][ok]
        q = [50, 90, 99, 99.9, 100]
        print(f"{name:22s} {k:5s} cuda " + " ".join(f"{np.percentile(ec, x):.2e}" for x in q) + " | f32 oracle " + " ".join(f"{np.percentile(eo, x):.2e}" for x in q))
