#!/usr/bin/env python
"""Kernel time of zb_actor_f32 over the batch size (CUDA events, L2 flushed between launches) + its FP32 roofline fraction."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ksim_gym_zbot_b200  # noqa: E402,F401
import zb_policy_oracle as PO  # noqa: E402
from ksim_gym_zbot_b200.policy import ActorParams, ActorPolicy  # noqa: E402

out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for prec, n in [(p_, n_) for p_ in (0, 1) for n_ in (1024, 4096, 8192, 16384, 65536)]:
    pol = ActorPolicy(ActorParams.random(0), n, "cuda:0", precision=prec)
    obs = torch.randn((n, 50), device="cuda")
    for _ in range(5):
        pol(obs)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(30)]
    for a, b in ev:
        flush.zero_()
        a.record(); pol(obs); b.record()
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
    out[("fp32 " if prec == 0 else "tf32x3-tcgen05 ") + str(n)] = {"ms": round(ms, 4), "TFLOP/s": round(n * PO.flops_per_env() / (ms * 1e-3) / 1e12, 2), "actor_evals/s": round(n / (ms * 1e-3))}
print(json.dumps(out))
