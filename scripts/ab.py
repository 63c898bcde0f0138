#!/usr/bin/env python
"""A/B harness for tuning variants of libzbphys.so (ksim-gym-zbot_b200/build.py: build_variants).

    python scripts/ab.py base=ksim-gym-zbot_b200/csrc/libzbphys.so v1=ksim-gym-zbot_b200/csrc/variants/libzbphys_v1.so \
        --envs 4096 --substeps 4 --steps 60 --warmup 10 --reps 3 [--c3]

Every library runs the SAME seeded workload (bench.py's c2 / c3: reset, random actions, fused termination + reset), interleaved
`reps` times; per control step the kernel is timed with CUDA events (L2 flushed in between).  Prints mean / median kernel time per
library and a hash of the final state: variants that are meant to be bit-exact must show the same hash as the base.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ksim_gym_zbot_b200  # noqa: E402,F401
from ksim_gym_zbot_b200 import model as M  # noqa: E402
from ksim_gym_zbot_b200.engine import EnvParams, PushEventConfig, ZbotCudaEngine  # noqa: E402


def run(lib, a, model, opts, seed=0):
    dev = torch.device("cuda", 0)
    n, S, K, W = a.envs, a.substeps, a.steps, a.warmup
    push = PushEventConfig() if a.c3 else None
    eng = ZbotCudaEngine(model, n, device=0, opts=opts, ctrl_dt=a.ctrl_dt, push=push, seed=seed, lib_path=lib,
                         outputs=("xpos", "qacc", "sensordata", "contact_mask", "solver_niter", "flags") if a.few_outputs else None)
    ep = EnvParams(n, dev, **M.randomize_env_params(model, n, seed + 100)) if a.c3 else None
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 17)
    bias = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=dev)
    actions = bias + 0.3 * (torch.rand((K + W, n, M.NU), generator=g, device=dev) * 2 - 1)
    pushes = torch.stack([eng.sample_push(n) for _ in range(K + W)]) if a.c3 else None
    rq, rv, rt = eng.sample_reset(n)
    ps = eng.reset(ep, reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt)
    done = torch.zeros((n,), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    ndone = torch.zeros((), dtype=torch.int64, device=dev)
    for t in range(K + W):
        if t >= W:
            flush.zero_()
            ev[t - W][0].record()
        eng.step(actions[t], ep, ps, push_rand=None if pushes is None else pushes[t], auto_reset=(rq, rv, rt), done_out=done)
        if t >= W:
            ev[t - W][1].record()
        ndone += (done != 0).sum()
    torch.cuda.synchronize()
    ms = np.array([e[0].elapsed_time(e[1]) for e in ev])
    h = hashlib.sha1()
    for t in (ps.data.qpos, ps.data.qvel, ps.data.qacc_warmstart, ps.actuator_state.position, ps.data.sensordata, ps.data.xpos):
        h.update(t.cpu().numpy().tobytes())
    niter = float(ps.data.solver_niter.float().mean())
    eng.close()
    return ms, h.hexdigest()[:12], int(ndone), niter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+", help="name=path")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--substeps", type=int, default=4)
    ap.add_argument("--ctrl-dt", type=float, default=0.02)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--c3", action="store_true")
    ap.add_argument("--few-outputs", action="store_true")
    a = ap.parse_args()
    model = M.build_zbot_syn()
    opts = M.make_opts(dt=a.ctrl_dt / a.substeps)
    libs = [x.split("=", 1) for x in a.libs]
    res = {k: [] for k, _ in libs}
    hashes = {}
    for rep in range(a.reps):
        for k, p in libs:
            ms, h, nd, ni = run(os.path.abspath(p), a, model, opts)
            res[k].append(ms)
            hashes[k] = (h, nd, ni)
    base = None
    for k, _ in libs:
        allms = np.concatenate(res[k])
        best = min(m.mean() for m in res[k])
        if base is None:
            base = best
        rate = a.envs * a.substeps / (best * 1e-3)
        print(f"{k:>14s}  envs {a.envs} S {a.substeps}  kernel ms mean(best rep) {best:.4f}  median {np.median(allms):.4f}  "
              f"env-steps/s {rate:.3e}  vs first {base / best:.3f}x  hash {hashes[k][0]} resets {hashes[k][1]} niter {hashes[k][2]:.2f}")


if __name__ == "__main__":
    main()
