#!/usr/bin/env python
"""First contact with the tcgen05 actor kernel: one launch, error against the fp64 restatement, next to the fp32 CUDA-core kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import ksim_gym_zbot_b200  # noqa
import zb_policy_oracle as PO
from ksim_gym_zbot_b200.policy import ActorParams, ActorPolicy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prm = ActorParams.random(seed=3)
g = np.random.default_rng(0)
obs = g.normal(0, 1, (n, 50)).astype(np.float32); carry = g.uniform(-1, 1, (n, 5, 128)).astype(np.float32)
mean, std, logits, nc = PO.actor_forward(prm.oracle_dict(), obs, carry)
for prec in (0, 1):
    pol = ActorPolicy(prm, n, "cuda:0", precision=prec, lib_path=os.environ.get("ZBPHYS_LIB_PROBE"))
    pol.carry.copy_(torch.tensor(carry))
    dist = torch.zeros((n, 3, 20, 5), device="cuda")
    pol.enqueue(torch.tensor(obs, device="cuda"), pol.action, dist_out=dist)
    torch.cuda.synchronize()
    c = pol.carry.cpu().numpy(); d = dist.cpu().numpy()
    print("precision", prec, "carry err per layer", [float(np.abs(c[:, l] - nc[:, l]).max()) for l in range(5)], "mean err", float(np.abs(d[:, 0] - mean).max()), flush=True)
