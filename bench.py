#!/usr/bin/env python
"""bench.py — Z-Bot physics env-steps/s on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (default N=1)
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on the host cores (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, weak scaling

A "step" is ONE CONTROL STEP of the whole batch: S fused physics substeps (PushEvent -> Feetech servo -> mjx.step), the
termination check and the partial reset of finished environments — the body of ksim's rollout scan around the engine
 — in ONE kernel launch (zb_step_autoreset_f32).  One env-step = one physics substep of one environment.
Default workload = BASELINE.json configs[1]: 4096 envs, random actions, ctrl_dt = 0.02 s with 4 substeps.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "zbot_physics_env_steps_per_sec"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # rollout length of the reference: 4.0 s / 0.02 s (train.py:1775)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c5"], help="BASELINE.json configs[1] / configs[2] / configs[4] (contact-heavy)")
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="0 = 4096 (c2) / 16384 (c3)")
    ap.add_argument("--substeps", type=int, default=4, help="physics substeps per control step (BASELINE config: 4; train.py:1777: 20)")
    ap.add_argument("--ctrl-dt", type=float, default=0.02)
    ap.add_argument("--solver", default="cg", choices=["cg", "newton"], help="ZbOpts.solver (ksim default unverified, see PARITY_RISKS.md)")
    ap.add_argument("--solver-iterations", type=int, default=8)
    ap.add_argument("--ls-iterations", type=int, default=8)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-shards", type=int, default=0, help="environment blocks (streams) the end-to-end loop drives the batch in; "
                    "0 = pick the fastest of 1 / 2 / 4 in a short untimed calibration")
    ap.add_argument("--e2e-blocking", action="store_true", help="end-to-end loop: wait for the WHOLE read-back before the next step "
                    "(default: only for what the next action depends on; the trajectory arrays stream out double-buffered)")
    ap.add_argument("--obs", action="store_true", help="observation epilogue: the launch also writes the actor [50] / critic [484] input vectors "
                    "(train.py:1624-1681); only those + the reward inputs are written / read back instead of every mjx.Data array")
    ap.add_argument("--seed", type=int, default=0)
    return ap.parse_args()


# -------------------------------------------------------------------------------------------------
# clocks sampler (recipe: /opt/skills/guides/B200_PROFILING.md)
# -------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)), "samples": len(sm),
                "reasons": sorted(reasons)}


# -------------------------------------------------------------------------------------------------
# workload pieces shared by both arms
# -------------------------------------------------------------------------------------------------
def workload_config(a):
    n = a.envs_per_gpu or (4096 if a.workload == "c2" else 16384)
    return n


def cpu_oracle_rate(model, opts, state_np, actions_np, S, params, seconds, threads, push_rand=None):
    """env-steps/s of the fp32 oracle (OpenMP over environments) on a bounded sample."""
    import zb_oracle

    o = zb_oracle.Oracle(model, opts, np.float32, threads=threads)
    st = {k: np.ascontiguousarray(v, dtype=np.float32).copy() for k, v in state_np.items()}
    n = st["qpos"].shape[0]
    o.step(st, actions_np[0], S, params=params)  # warm-up (page-in, thread pool)
    t0, steps, i = time.perf_counter(), 0, 1
    while True:
        o.step(st, actions_np[i % len(actions_np)], S, params=params)
        steps += 1
        i += 1
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return n * S * steps / dt, steps, dt


def oracle_flops_per_env_step(model, opts, state_np, actions_np, S, params):
    import zb_oracle

    o = zb_oracle.Oracle(model, opts, np.float32, count=True, threads=1)
    n = min(64, state_np["qpos"].shape[0])
    st = {k: np.ascontiguousarray(v[:n], dtype=np.float32).copy() for k, v in state_np.items()}
    pp = {k: v[:n] for k, v in params.items()} if params else None
    out = o.new_out(n)
    o.flops_reset()
    o.step(st, actions_np[0][:n], S, params=pp, out=out)
    fl = o.flops()
    return sum(fl.values()) / (n * S), {k: v / (n * S) for k, v in fl.items()}, float(out["solver_niter"].mean())


# bytes per control step per env that MUST cross HBM: state in/out + action + per-env model + outputs
def algorithmic_bytes_per_ctrl_step(randomized: bool, push: bool, obs: bool = False) -> int:
    state = 27 + 26 + 26 + 60 + 2  # qpos qvel warmstart planner(3x20) t_rem time
    rd = state + 20 + (7 if push else 0) + (165 if randomized else 0)
    wr = state + (78 + 104 + 260 + 156 + 26 + 26 + 20 + 20 + 18 + 9 + 27 + 8 + 24 + 84 + 3)  # every ZbOut array
    if obs:  # carry (x, lag) in/out + both network vectors + the reward inputs (xpos qacc actuator_force sensordata site_xpos, 3 ints)
        rd += 5
        wr = state + 5 + 50 + 484 + (78 + 26 + 20 + 18 + 9 + 3)
    return 4 * (rd + wr)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    S = a.substeps
    n_env = workload_config(a)
    import ksim_gym_zbot_b200  # noqa: F401
    from ksim_gym_zbot_b200 import model as M

    model = M.build_zbot_syn()
    opts = M.make_opts(dt=a.ctrl_dt / S, iterations=a.solver_iterations, ls_iterations=a.ls_iterations,
                       solver=M.SOLVER_NEWTON if a.solver == "newton" else M.SOLVER_CG)
    randomized = a.workload in ("c3", "c5")
    config = {
        "workload": {"c2": "BASELINE.json configs[1]: Z-Bot 4096 envs random-action rollout, ctrl_dt=0.02, 4 substeps",
                     "c3": "BASELINE.json configs[2]: Z-Bot 16384 envs, push + friction/mass randomization + resets",
                     "c5": "BASELINE.json configs[4]: contact-heavy (falls and pushes): resets tilted up to 40 deg and dropped from +8 cm, "
                           "pushes every 0.2-0.4 s at 10x magnitude, randomization + resets"}[a.workload],
        "model": "zbot_syn v1 (synthetic Z-Bot topology; the real MJCF is not shipped with the reference)",
        "envs_per_gpu": n_env, "substeps": S, "dt": a.ctrl_dt / S, "ctrl_dt": a.ctrl_dt, "solver": a.solver,
        "iterations": a.solver_iterations, "ls_iterations": a.ls_iterations, "integrator": "implicitfast",
        "actions": "bias + 0.3*U(-1,1), torch Philox seed %d" % a.seed, "terminations": "z in [0.05,0.5], tilt<60deg -> in-place reset",
        "l2": "flushed between timed steps (256 MiB memset outside the event pair)", "parallelism": f"dp{world} (envs sharded, no collective in the physics step)",
    }

    if a.impl == "reference":
        if rank != 0:
            return
        run_reference(a, model, opts, S, n_env, randomized, config, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the physics step has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from ksim_gym_zbot_b200.engine import EnvParams, PushEventConfig, ZbotCudaEngine
    from ksim_gym_zbot_b200.sharding import aggregate_throughput, rank_seed

    push_cfg = PushEventConfig() if randomized else None
    if a.workload == "c5":
        push_cfg = PushEventConfig(x_linvel=1.0, y_linvel=1.0, z_linvel=0.5, interval_range=(0.2, 0.4))
    # --obs: the network inputs come from the epilogue; the remaining mjx.Data outputs are the ones the rewards / terminations read
    # (train.py:476, 503-546, 681-733: base / feet positions, touch + force sensors, joint state, contact flags)
    obs_outputs = ("xpos", "qacc", "actuator_force", "sensordata", "site_xpos", "contact_mask", "solver_niter", "flags")
    eng = ZbotCudaEngine(model, n_env, device=local, opts=opts, ctrl_dt=a.ctrl_dt, push=push_cfg, seed=rank_seed(a.seed, rank),
                         outputs=obs_outputs if a.obs else None)
    config["observation_epilogue"] = bool(a.obs)
    params_np = M.randomize_env_params(model, n_env, a.seed + 100 + rank) if randomized else None
    ep = EnvParams(n_env, dev, **params_np) if randomized else None
    K, W = a.steps, a.warmup
    g = torch.Generator(device=dev)
    g.manual_seed(rank_seed(a.seed, rank) + 17)
    bias = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=dev)
    actions = bias + 0.3 * (torch.rand((K + W, n_env, M.NU), generator=g, device=dev) * 2 - 1)
    pushes = torch.stack([eng.sample_push(n_env) for _ in range(K + W)]) if randomized else None
    rq, rv, rt = eng.sample_reset(n_env)
    if a.workload == "c5":  # tilted + dropped reset states
        ax = torch.randn((n_env, 3), generator=g, device=dev)
        ax = ax / ax.norm(dim=1, keepdim=True)
        ang = (torch.rand((n_env,), generator=g, device=dev) * 2 - 1) * math.radians(40)
        rq[:, 3] = torch.cos(ang / 2)
        rq[:, 4:7] = ax * torch.sin(ang / 2)[:, None]
        rq[:, 2] += 0.08
    obs = eng.new_observations(n_env) if a.obs else None
    obs_lag = eng.sample_imu_lag(n_env, (0.0, 0.1)) if a.obs else None  # lag of re-initialised environments (pre-sampled)
    ps = eng.reset(ep, reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt, obs=obs, reset_imu_lag=obs_lag)
    done = torch.zeros((n_env,), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_done = torch.zeros((), dtype=torch.int64, device=dev)

    def ctrl_step(t, act):
        # ONE launch: S fused substeps + termination check + partial reset (zb_step_autoreset_f32)
        eng.step(act, ep, ps, push_rand=None if pushes is None else pushes[t], auto_reset=(rq, rv, rt), done_out=done, obs=obs,
                 reset_imu_lag=obs_lag)

    for t in range(W):
        ctrl_step(t, actions[t])
    torch.cuda.synchronize()
    # snapshot for the CPU baseline: the state distribution the timed region starts from
    snap = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        ns = min(512, n_env)
        snap = {"qpos": ps.data.qpos[:ns].cpu().numpy(), "qvel": ps.data.qvel[:ns].cpu().numpy(),
                "qacc_warmstart": ps.data.qacc_warmstart[:ns].cpu().numpy(), "act_pos": ps.actuator_state.position[:ns].cpu().numpy(),
                "act_vel": ps.actuator_state.velocity[:ns].cpu().numpy(), "act_torque": ps.actuator_state.last_computed_torque[:ns].cpu().numpy(),
                "push_t_rem": np.full(ns, 1e9, np.float32), "time": np.zeros(ns, np.float32)}
        snap_actions = actions[W:W + 8, :ns].cpu().numpy()

    # ---------------- timed region 1: inputs resident in HBM ----------------
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = eng.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    niter_sum, ncon_sum, ndone = 0.0, 0.0, 0
    for i in range(K):
        t = W + i
        flush.zero_()  # L2 flush, outside the event pair
        ev[i][0].record()
        ctrl_step(t, actions[t])
        ev[i][1].record()
        ev[i][2].record()
        n_done += (done != 0).sum()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count() - launches0
    step_ms = sum(e[0].elapsed_time(e[1]) for e in ev)
    total_ms = sum(e[0].elapsed_time(e[2]) for e in ev)
    value, total_ms_max = aggregate_throughput(n_env * S * K, total_ms, dev)  # sum of units / max over ranks of device time
    niter_mean = float(ps.data.solver_niter.float().mean().item())
    ncon_mean = float(sum(((ps.data.contact_mask >> c) & 1).float().mean().item() for c in range(8)))
    resets_per_step = float(n_done.item()) / K

    # ---------------- timed region 2: end to end through the public API with HOST buffers ----------------
    # Every step: H2D of that step's actions from pinned memory -> engine.step (one launch incl. termination + reset) -> D2H of
    # everything the reference's observations, rewards and trajectory read back (train.py:452-461, 476, 681, 854, 882, 1307)
    # + the done mask.  The host waits, before it issues the next step, for what the next action can depend on (the new state,
    # the sensors, the servo state, the done mask [+ the actor observation with --obs]: `gate_fields`); the arrays that only the
    # trajectory / critic / rewards consume (`bulk_fields`) are written into one of two device buffers in turn and stream out on a
    # copy stream while the next step computes (--e2e-blocking waits for them as well).  The batch can additionally be driven as
    # several environment blocks on separate streams (the C ABI takes any [lo, hi) block); every group of arrays is ONE D2H copy.
    gate_fields = ("qpos", "qvel", "sensordata")
    bulk_fields = ("xpos", "xquat", "cinert", "cvel", "qacc", "actuator_force")
    if a.obs:
        bulk_fields = ("xpos", "qacc", "actuator_force", "site_xpos")
    from ksim_gym_zbot_b200.sharding import shard_range

    def build_shards(nsh):
        shards = []
        for k in range(nsh):
            lo, hi = shard_range(n_env, nsh, k)
            sp = eng.new_state(hi - lo)
            for f in ("qpos", "qvel", "qacc_warmstart", "time"):
                getattr(sp.data, f).copy_(getattr(ps.data, f)[lo:hi])
            for f in ("position", "velocity", "last_computed_torque"):
                getattr(sp.actuator_state, f).copy_(getattr(ps.actuator_state, f)[lo:hi])
            sp.event_states["push_t_rem"].copy_(ps.event_states["push_t_rem"][lo:hi])
            sob = None
            if a.obs:
                sob = eng.new_observations(hi - lo)
                sob.imu_x.copy_(obs.imu_x[lo:hi]); sob.imu_lag.copy_(obs.imu_lag[lo:hi])

            def owner(f):  # (object, attribute) holding array f
                if f == "act_torque":
                    return sp.actuator_state, "last_computed_torque"
                return (sob, f) if f.endswith("_obs") else (sp.data, f)

            def pack(fields):  # one flat device buffer with a view per field (initialised from the current arrays)
                sizes = {f: getattr(*owner(f)).numel() for f in fields}
                flat = torch.empty(sum(sizes.values()), dtype=torch.float32, device=dev)
                views, off = {}, 0
                for f, nel in sizes.items():
                    cur = getattr(*owner(f))
                    views[f] = flat[off:off + nel].view(cur.shape)
                    views[f].copy_(cur)
                    off += nel
                return flat, views

            gate = gate_fields + ("act_torque",) + (("actor_obs",) if a.obs else ())
            bulk = bulk_fields + (("critic_obs",) if a.obs else ())
            g_flat, g_views = pack(gate)
            for f, v in g_views.items():
                setattr(*owner(f), v)
            b_flat, b_views = zip(*[pack(bulk) for _ in range(2)])
            done_k = torch.zeros((hi - lo,), dtype=torch.uint8, device=dev)
            ep_k = None if ep is None else ep.slice(lo, hi)
            rq_k, rv_k, rt_k = rq[lo:hi].contiguous(), rv[lo:hi].contiguous(), rt[lo:hi].contiguous()
            lag_k = None if obs_lag is None else obs_lag[lo:hi].contiguous()
            plans = []
            for b in range(2):  # one pre-bound control step per bulk buffer (engine.plan_step: argument structs built once)
                for f, v in b_views[b].items():
                    setattr(*owner(f), v)
                plans.append(eng.plan_step(sp, ep_k, auto_reset=(rq_k, rv_k, rt_k), done_out=done_k, obs=sob, reset_imu_lag=lag_k))
            shards.append(dict(lo=lo, hi=hi, ps=sp, owner=owner, plans=plans, g_flat=g_flat, h_gate=torch.empty_like(g_flat, device="cpu").pin_memory(),
                               b_flat=b_flat, b_views=b_views, h_bulk=[torch.empty_like(b, device="cpu").pin_memory() for b in b_flat],
                               done=done_k, h_done=torch.empty((hi - lo,), dtype=torch.uint8).pin_memory(),
                               d_act=torch.empty((hi - lo, M.NU), dtype=torch.float32, device=dev), stream=torch.cuda.Stream(device=dev), copy_stream=torch.cuda.Stream(device=dev),
                               ev_kernel=[torch.cuda.Event() for _ in range(2)], ev_copy=[torch.cuda.Event() for _ in range(2)]))
        return shards

    h_act = actions[W:W + K].cpu().pin_memory()

    def run_e2e(shards, steps):
        """`steps` closed-loop control steps; returns the device time in ms (events on the main stream around everything)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream()
        e0.record()
        for sh in shards:
            sh["stream"].wait_stream(main)
            sh["copy_stream"].wait_stream(main)
        for i in range(steps):
            t, b = W + i, i & 1
            for sh in shards:
                with torch.cuda.stream(sh["stream"]):
                    if i >= 2:
                        sh["stream"].wait_event(sh["ev_copy"][b])  # bulk buffer b has been read out
                    sh["d_act"].copy_(h_act[i, sh["lo"]:sh["hi"]], non_blocking=True)
                    # this step's trajectory arrays go to bulk buffer b (plan b is bound to it)
                    sh["plans"][b](sh["d_act"], None if pushes is None else pushes[t, sh["lo"]:sh["hi"]])
                    sh["h_gate"].copy_(sh["g_flat"], non_blocking=True)
                    sh["h_done"].copy_(sh["done"], non_blocking=True)
                    # recorded AFTER the gating copies: all D2H copies share one DMA queue, and the bulk copy must not get in front
                    # of the small one the host is waiting for
                    sh["ev_kernel"][b].record(sh["stream"])
                    if a.e2e_blocking:
                        sh["h_bulk"][b].copy_(sh["b_flat"][b], non_blocking=True)
                        sh["ev_copy"][b].record(sh["stream"])
                if not a.e2e_blocking:
                    with torch.cuda.stream(sh["copy_stream"]):
                        sh["copy_stream"].wait_event(sh["ev_kernel"][b])
                        sh["h_bulk"][b].copy_(sh["b_flat"][b], non_blocking=True)
                        sh["ev_copy"][b].record(sh["copy_stream"])
            for sh in shards:
                sh["stream"].synchronize()  # the caller has the gating result of step i before it issues step i+1
        for sh in shards:
            main.wait_stream(sh["stream"])
            main.wait_stream(sh["copy_stream"])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # how many blocks: a launch picks its CTA shape for a GPU of its own, so the best split depends on the batch size
    cands = [max(1, min(a.e2e_shards, n_env // 16))] if a.e2e_shards > 0 else [c for c in (1, 2, 4) if n_env // c >= 16]
    calib = {}
    if len(cands) > 1:
        for c in cands:
            sh_ = build_shards(c)
            run_e2e(sh_, min(3, K))
            calib[c] = run_e2e(sh_, min(10, K)) / min(10, K)
            del sh_
        t_c = torch.tensor([calib[c] for c in cands], device=dev)
        if world > 1:
            dist.all_reduce(t_c, op=dist.ReduceOp.MAX)  # every rank must take the same decision
        nsh = cands[int(torch.argmin(t_c).item())]
    else:
        nsh = cands[0]
    shards = build_shards(nsh)
    h2d = h_act[0].numel() * 4
    d2h_gate = sum(sh["h_gate"].numel() * 4 + sh["h_done"].numel() for sh in shards)
    d2h = d2h_gate + sum(sh["h_bulk"][0].numel() * 4 for sh in shards)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_ms = run_e2e(shards, K)
    e2e_value, _ = aggregate_throughput(n_env * S * K, e2e_ms, dev)
    clk = clocks.stop() if rank == 0 else None

    # ---------------- FP32 peak probe (roofline denominator; not in MEASURED_PEAKS.json) ----------------
    import ctypes

    scratch = torch.empty(148 * 16 * 256, dtype=torch.float32, device=dev)
    best = 0.0
    for _ in range(4):
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        eng.lib.zb_bench_ffma(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.c_void_p(scratch.data_ptr()), 148 * 16, 20000)
        p1.record()
        torch.cuda.synchronize()
        best = max(best, 148 * 16 * 256 * 20000 * 32 / (p0.elapsed_time(p1) * 1e-3) / 1e12)
    fp32_peak = best

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- CPU baseline + algorithmic flops (rank 0, N=1 only) ----------------
    cpu, flops_per, flops_stage, niter_cpu = None, None, None, None
    cores = os.cpu_count() or 1
    if snap is not None:
        sp = {k: v[:snap["qpos"].shape[0]] for k, v in params_np.items()} if params_np else None
        rate, nst, secs = cpu_oracle_rate(model, opts, snap, snap_actions, S, sp, a.cpu_baseline_seconds, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{snap['qpos'].shape[0]} envs x {nst} control steps ({secs:.1f} s) from the bench's post-warm-up state, fp32 oracle, OpenMP",
               "note": "CPU restatement (MJX/JAX not installable in this image)"}
        flops_per, flops_stage, niter_cpu = oracle_flops_per_env_step(model, opts, snap, snap_actions, S, sp)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None  # DRAM bytes per launch of the step kernel from the committed ncu capture of THIS workload, if there is one
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get("c2_s20" if (a.workload == "c2" and S == 20) else a.workload)
        if tr and tr["envs"] == n_env and tr["substeps"] == S and not a.obs and a.solver == "cg":
            traffic = tr["bytes_per_launch"]
    except Exception:
        pass
    bytes_ctrl = algorithmic_bytes_per_ctrl_step(randomized, randomized, a.obs)
    step_s = step_ms * 1e-3 / K  # average duration of the dominant kernel (zb_kernel<STEP>)
    roof = {
        "bound": "fp32", "unit": "TFLOP/s", "kernel": "zb_kernel<ZB_MODE_STEP>",
        "achieved": None if flops_per is None else flops_per * n_env * S / step_s / 1e12,
        "peak": fp32_peak, "peak_source": "measured in this run (FFMA probe zb_bench_ffma; MEASURED_PEAKS.json has no FP32 figure)",
        "frac": None if flops_per is None else flops_per * n_env * S / step_s / 1e12 / fp32_peak,
        "traffic": traffic, "traffic_note": "ncu dram bytes per launch (profiles/r1_traffic.json); algorithmic bytes per launch = %d" % (bytes_ctrl * n_env),
        "flops_per_env_step": flops_per, "flops_source": "oracle op counters (fp32, 64-env sample of this workload)",
        "kernel_ms_per_launch": step_ms / K, "kernel_share_of_step": step_ms / total_ms,
        "hbm": {"bound": "hbm", "unit": "GB/s", "achieved": bytes_ctrl * n_env / step_s / 1e9, "peak": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "frac": bytes_ctrl * n_env / step_s / 1e9 / hbm_peak,
                "bytes_per_env_ctrl_step": bytes_ctrl},
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms_max / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "d2h_gating_bytes_per_step": d2h if a.e2e_blocking else d2h_gate,
                "readback": "blocking" if a.e2e_blocking else "gating arrays blocking, trajectory arrays double-buffered on a copy stream", "streams": nsh,
                "streams_calibration_ms_per_step": calib or None},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        "stats": {"solver_niter_mean_last_step": niter_mean, "active_contacts_mean_last_step": ncon_mean, "resets_per_step": resets_per_step,
                  "ctrl_steps_per_s": value / S, "solver_niter_mean_cpu_sample": niter_cpu, "flops_per_stage": flops_stage},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(a, model, opts, S, n_env, randomized, config, world):
    """The reference arm: the reference's algorithm for this path on the host cores.  MJX/JAX cannot be installed here
    (no wheels, no network), so this times the oracle port (oracle/zb_oracle.c, fp32, OpenMP over environments)."""
    from ksim_gym_zbot_b200 import model as M

    cores = os.cpu_count() or 1
    n = min(512, n_env)  # bounded sample of the workload per step
    import zb_oracle

    o = zb_oracle.Oracle(model, opts, np.float32, threads=cores)
    rng = np.random.default_rng(a.seed)
    st = o.new_state(n)
    q = np.repeat(M.default_qpos(model)[None], n, 0).astype(np.float32)
    v = np.zeros((n, M.NV), np.float32)
    v[:, 6:] = rng.uniform(-0.01, 0.01, (n, M.NU))
    params = {k: x for k, x in M.randomize_env_params(model, n, a.seed + 100).items()} if randomized else None
    if a.workload == "c5":  # tilted + dropped reset states, as in the CUDA arm
        ax = rng.normal(size=(n, 3))
        ax /= np.linalg.norm(ax, axis=1, keepdims=True)
        ang = rng.uniform(-1, 1, n) * math.radians(40)
        q[:, 3] = np.cos(ang / 2)
        q[:, 4:7] = ax * np.sin(ang / 2)[:, None]
        q[:, 2] += 0.08
    # PushEvent numbers (same stand-in sampler as engine.sample_push)
    scale = np.array([0.1, 0.1, 0.05, 0, 0, 0], np.float32) * (10.0 if a.workload == "c5" else 1.0)
    iv = (0.2, 0.4) if a.workload == "c5" else (2.0, 4.0)
    K, W = a.steps, a.warmup
    pushes = None
    t_rem0 = np.full(n, 1e9, np.float32)
    if randomized:
        u = rng.uniform(0, 1, (K + W, n, 8)).astype(np.float32)
        pushes = np.concatenate([(u[..., :6] * 2 - 1) * scale * ((0.05 + 0.10 * u[..., 6:7]) / 0.15), iv[0] + (iv[1] - iv[0]) * u[..., 7:8]], axis=-1).astype(np.float32)
        t_rem0 = rng.uniform(iv[0], iv[1], n).astype(np.float32)
    o.reset(st, q, v, t_rem0, params=params)
    acts = (M.BIAS_POSE[None, None] + 0.3 * rng.uniform(-1, 1, (K + W, n, M.NU))).astype(np.float32)

    def term(s):
        qq = s["qpos"]
        up = qq[:, 3] ** 2 - qq[:, 4] ** 2 - qq[:, 5] ** 2 + qq[:, 6] ** 2
        return ((qq[:, 2] < 0.05) | (qq[:, 2] > 0.5) | (up < 0.5) | ~np.isfinite(qq[:, 2])).astype(np.uint8)

    def ctrl_step(t):
        o.step(st, acts[t], S, params=params, push_rand=None if pushes is None else pushes[t])
        d = term(st)
        if d.any():
            o.reset(st, q, v, t_rem0, done=d, params=params)

    for t in range(W):
        ctrl_step(t)
    t0 = time.perf_counter()
    for t in range(W, W + K):
        ctrl_step(t)
    dt = time.perf_counter() - t0
    value = n * S * K / dt
    sample = f"{n} envs per step (bounded sample of the {n_env}-env workload), {K} control steps, fp32 oracle port, OpenMP {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt * 1e3 / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU restatement of mjx.step + MjxEngine.step (MJX/JAX/ksim not installable: no wheels, no network)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
