#!/usr/bin/env python
"""bench.py — Z-Bot physics env-steps/s on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (default N=1)
    python bench.py --impl reference --steps K --warmup W    # the reference's algorithm on the host cores (oracle port)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, weak scaling

A "step" is ONE CONTROL STEP of the whole batch: S fused physics substeps (PushEvent -> Feetech servo -> mjx.step), the
termination check, the partial reset of finished environments and the observation gathering for the two networks — the body of
ksim's rollout scan around the engine — in ONE kernel launch (zb_step_obs_f32).  One env-step = one physics
substep of one environment.  Default workload = BASELINE.json configs[1]: 4096 envs, random actions, ctrl_dt = 0.02 s with 4
substeps.  The same JSON line carries short runs of the other single-GPU configurations (`other_workloads`: 8192 envs per GPU =
the 65 536-env job of north_star on 8 GPUs, configs[2] at 16 384 envs, and the reference's own S = 20 launch config).
"""
from __future__ import annotations

import argparse
import hashlib
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
WORKLOADS = {
    "c2": "BASELINE.json configs[1]: Z-Bot 4096 envs random-action rollout, ctrl_dt=0.02, 4 substeps",
    "c3": "BASELINE.json configs[2]: Z-Bot 16384 envs, push + friction/mass randomization + resets",
    "c5": "BASELINE.json configs[4]: contact-heavy (falls and pushes): resets tilted up to 40 deg and dropped from +8 cm, "
          "pushes every 0.2-0.4 s at 10x magnitude, randomization + resets",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # rollout length of the reference: 4.0 s / 0.02 s (train.py:1775)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS), help="BASELINE.json configs[1] / configs[2] / configs[4] (contact-heavy)")
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="0 = 4096 (c2) / 16384 (c3)")
    ap.add_argument("--substeps", type=int, default=4, help="physics substeps per control step (BASELINE config: 4; train.py:1777: 20)")
    ap.add_argument("--ctrl-dt", type=float, default=0.02)
    ap.add_argument("--solver", default="cg", choices=["cg", "newton"], help="ZbOpts.solver (ksim default unverified, see PARITY_RISKS.md)")
    ap.add_argument("--solver-iterations", type=int, default=8)
    ap.add_argument("--ls-iterations", type=int, default=8)
    ap.add_argument("--cpu-baseline-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the short runs of the other configurations (`other_workloads`)")
    ap.add_argument("--e2e-shards", type=int, default=1, help="environment blocks (streams) the end-to-end loop drives the batch in")
    ap.add_argument("--e2e-blocking", action="store_true", help="end-to-end loop: wait for the WHOLE read-back before the next step "
                    "(default: only for what the next action depends on; the rest streams out double-buffered)")
    ap.add_argument("--e2e-no-bulk", action="store_true", help=argparse.SUPPRESS)  # diagnosis: skip the streamed bulk read-back
    ap.add_argument("--e2e-no-graph", action="store_true", help="end-to-end loop: plain enqueues instead of one CUDA graph per control step")
    ap.add_argument("--no-obs", action="store_true", help="without the observation epilogue: the launch writes every mjx.Data array the task reads "
                    "(xpos xquat cinert cvel ...) instead of the actor [50] / critic [484] input vectors (train.py:1624-1681) + the reward inputs")
    ap.add_argument("--obs", action="store_true", help=argparse.SUPPRESS)  # (the default since round 2; kept so that old command lines still run)
    ap.add_argument("--flush", default="write", choices=["write", "write+read"],
                    help="L2 flush between timed steps: memset of 256 MiB; write+read adds a read pass over it, so that the step starts on a "
                         "cold AND clean L2 (after a bare memset the L2 holds ~126 MB of dirty lines whose write-back ncu books on the step kernel)")
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
# workload description shared by both arms
# -------------------------------------------------------------------------------------------------
class Spec:
    """One benchmark configuration."""

    def __init__(self, a, workload=None, n_env=None, substeps=None):
        self.workload = workload or a.workload
        self.S = int(substeps or a.substeps)
        self.n_env = int(n_env or a.envs_per_gpu or (4096 if self.workload == "c2" else 16384))
        self.ctrl_dt = a.ctrl_dt
        self.solver, self.iterations, self.ls_iterations = a.solver, a.solver_iterations, a.ls_iterations
        self.obs = not a.no_obs
        self.seed = a.seed
        self.flush = a.flush
        self.randomized = self.workload in ("c3", "c5")

    def opts(self):
        from ksim_gym_zbot_b200 import model as M

        return M.make_opts(dt=self.ctrl_dt / self.S, iterations=self.iterations, ls_iterations=self.ls_iterations,
                           solver=M.SOLVER_NEWTON if self.solver == "newton" else M.SOLVER_CG)

    def config(self, world: int) -> dict:
        """The `config` object of the JSON line — identical for both arms (the driver compares them)."""
        return {
            "workload": WORKLOADS[self.workload],
            "model": "zbot_syn v1 (synthetic Z-Bot topology; the real MJCF is not shipped with the reference)",
            "envs_per_gpu": self.n_env, "substeps": self.S, "dt": self.ctrl_dt / self.S, "ctrl_dt": self.ctrl_dt, "solver": self.solver,
            "iterations": self.iterations, "ls_iterations": self.ls_iterations, "integrator": "implicitfast",
            "actions": "bias + 0.3*U(-1,1), seed %d" % self.seed,
            "terminations": "z in [0.05,0.5], tilt<60deg, episode length 80 s -> in-place reset",
            "observations": "actor [50] + critic [484] input vectors (train.py:1624-1681) + reward inputs per control step" if self.obs
                            else "every mjx.Data array the task reads",
            "l2": "flushed between timed steps (256 MiB memset" + (" + read-back of it" if self.flush == "write+read" else "")
                  + ", outside the event pair)",
            "parallelism": f"dp{world} (envs sharded, no collective in the physics step)",
        }

    def short(self) -> str:
        return f"{self.workload} {self.n_env} envs/GPU S={self.S}"


class CpuArm:
    """The reference's algorithm for this path on the host cores: the oracle port (oracle/zb_oracle.c, fp32, OpenMP over environments,
    built -O3 -march=native on this host) driving the SAME workload as the CUDA arm: reset, random actions, pushes, terminations,
    partial resets and (with observations on) the observation vectors.  Used by `--impl reference` and by `cpu_baseline`."""

    def __init__(self, spec: Spec, model, threads: int, table_steps: int = 32):
        import zb_oracle
        from ksim_gym_zbot_b200 import model as M

        self.spec, self.n, self.S, self.threads = spec, spec.n_env, spec.S, threads
        n = self.n
        self.o = zb_oracle.Oracle(model, spec.opts(), np.float32, threads=threads, fast=True)
        rng = np.random.default_rng(spec.seed)
        self.st = self.o.new_state(n)
        self.q = np.repeat(M.default_qpos(model)[None], n, 0).astype(np.float32)
        self.v = np.zeros((n, M.NV), np.float32)
        self.v[:, 6:] = rng.uniform(-0.01, 0.01, (n, M.NU))
        self.params = dict(M.randomize_env_params(model, n, spec.seed + 100)) if spec.randomized else None
        if spec.workload == "c5":  # tilted + dropped reset states, as in the CUDA arm
            ax = rng.normal(size=(n, 3))
            ax /= np.linalg.norm(ax, axis=1, keepdims=True)
            ang = rng.uniform(-1, 1, n) * math.radians(40)
            self.q[:, 3] = np.cos(ang / 2)
            self.q[:, 4:7] = ax * np.sin(ang / 2)[:, None]
            self.q[:, 2] += 0.08
        # PushEvent numbers (same stand-in sampler as engine.sample_push)
        scale = np.array([0.1, 0.1, 0.05, 0, 0, 0], np.float32) * (10.0 if spec.workload == "c5" else 1.0)
        iv = (0.2, 0.4) if spec.workload == "c5" else (2.0, 4.0)
        self.T = table_steps
        self.pushes = None
        self.t_rem0 = np.full(n, 1e9, np.float32)
        if spec.randomized:
            u = rng.uniform(0, 1, (self.T, n, 8)).astype(np.float32)
            self.pushes = np.concatenate([(u[..., :6] * 2 - 1) * scale * ((0.05 + 0.10 * u[..., 6:7]) / 0.15), iv[0] + (iv[1] - iv[0]) * u[..., 7:8]], axis=-1).astype(np.float32)
            self.t_rem0 = rng.uniform(iv[0], iv[1], n).astype(np.float32)
        self.acts = (M.BIAS_POSE[None, None] + 0.3 * rng.uniform(-1, 1, (self.T, n, M.NU))).astype(np.float32)
        self.out = self.o.new_out(n) if spec.obs else None
        self.imu_x, self.imu_lag = np.zeros((n, 4), np.float32), rng.uniform(0, 0.1, n).astype(np.float32)
        self.o.reset(self.st, self.q, self.v, self.t_rem0, params=self.params, out=self.out)
        self.t = 0

    def ctrl_step(self):
        t, st = self.t % self.T, self.st
        self.o.step(st, self.acts[t], self.S, params=self.params, push_rand=None if self.pushes is None else self.pushes[t], out=self.out)
        qq = st["qpos"]
        up = qq[:, 3] ** 2 - qq[:, 4] ** 2 - qq[:, 5] ** 2 + qq[:, 6] ** 2
        d = ((qq[:, 2] < 0.05) | (qq[:, 2] > 0.5) | (up < 0.5) | ~np.isfinite(qq[:, 2]) | (st["time"] > 80.0)).astype(np.uint8)
        if d.any():
            self.o.reset(st, self.q, self.v, self.t_rem0, done=d, params=self.params, out=self.out)
        if self.spec.obs:
            self.o.observe(st, self.out, self.imu_x, self.imu_lag, reinit=d, reset_lag=self.imu_lag)
            self.o.reward_inputs(st, self.out)
        self.t += 1

    def run_steps(self, k: int) -> float:
        t0 = time.perf_counter()
        for _ in range(k):
            self.ctrl_step()
        return time.perf_counter() - t0

    def run_seconds(self, seconds: float):
        t0, k = time.perf_counter(), 0
        while True:
            self.ctrl_step()
            k += 1
            if time.perf_counter() - t0 >= seconds:
                break
        return k, time.perf_counter() - t0

    def describe(self, k: int, secs: float) -> dict:
        return {"cores": self.threads, "kind": "port",
                "sample": f"{self.n} envs (the full per-GPU batch of this workload) x {k} control steps ({secs:.1f} s), closed loop with terminations + resets"
                          + (" + observation vectors" if self.spec.obs else "") + f", fp32 oracle port built -O3 -march=native, OpenMP {self.threads} threads",
                "note": "CPU restatement of mjx.step + MjxEngine.step (MJX/JAX/ksim not installable: no wheels, no network)"}


def oracle_flops_per_env_step(model, opts, state_np, actions_np, S, params):
    """Algorithmic FLOPs per env-step by stage from the oracle's op counters (64-env sample of the bench's own state);
    returns (per-stage dict, mean solver iterations of the sample)."""
    import zb_oracle

    o = zb_oracle.Oracle(model, opts, np.float32, count=True, threads=1)
    n = min(64, state_np["qpos"].shape[0])
    st = {k: np.ascontiguousarray(v[:n], dtype=np.float32).copy() for k, v in state_np.items()}
    pp = {k: v[:n] for k, v in params.items()} if params else None
    out = o.new_out(n)
    o.flops_reset()
    o.step(st, actions_np[:n], S, params=pp, out=out)
    fl = o.flops()
    return {k: v / (n * S) for k, v in fl.items()}, float(out["solver_niter"].mean())


# bytes per control step per env that MUST cross HBM: state in/out + action + per-env model + outputs
def algorithmic_bytes_per_ctrl_step(randomized: bool, push: bool, obs: bool = False) -> int:
    state = 27 + 26 + 26 + 60 + 2  # qpos qvel warmstart planner(3x20) t_rem time
    rd = state + 20 + (7 if push else 0) + (165 if randomized else 0)
    wr = state + (78 + 104 + 260 + 156 + 26 + 26 + 20 + 20 + 18 + 9 + 27 + 8 + 24 + 84 + 3)  # every ZbOut array
    if obs:  # carry (x, lag) in/out + both network vectors + reward inputs (6 + 6) + (xpos xquat qacc actuator_force sensordata site_xpos, 3 ints)
        rd += 5
        wr = state + 5 + 50 + 484 + 12 + (78 + 104 + 26 + 20 + 18 + 9 + 3)
    return 4 * (rd + wr)


def kernel_source_sha() -> str:
    h = hashlib.sha1()
    for f in ("ksim-gym-zbot_b200/csrc/zbphys.cu", "ksim-gym-zbot_b200/csrc/zb_kernel.cuh", "ksim-gym-zbot_b200/csrc/zb_math.cuh", "include/zbphys.h"):
        h.update(open(os.path.join(ROOT, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(spec: Spec):
    """DRAM bytes per launch of the step kernel from the committed ncu capture (profiles/traffic.json, written by
    scripts/refresh_traffic.sh) — only if that capture was taken from the kernel sources as they are NOW, else None."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tr.get("kernel_source_sha") != kernel_source_sha():
            return None, "profiles/traffic.json is from other kernel sources (run scripts/refresh_traffic.sh)"
        key = f"{spec.workload}_n{spec.n_env}_s{spec.S}_{'obs' if spec.obs else 'noobs'}_{spec.solver}"
        e = tr["captures"].get(key)
        if e:
            return int(e["dram_bytes_per_launch"]), f"ncu dram__bytes_read+write per launch, {tr.get('when', '')} ({key})"
        return None, f"no capture for {key}"
    except Exception as ex:  # noqa: BLE001
        return None, f"profiles/traffic.json unreadable: {ex}"


# -------------------------------------------------------------------------------------------------
# the CUDA arm: one workload
# -------------------------------------------------------------------------------------------------
def measure(a, spec: Spec, model, K: int, W: int, rank: int, world: int, local: int, primary: bool):
    import torch
    import torch.distributed as dist

    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import EnvParams, PushEventConfig, ZbotCudaEngine
    from ksim_gym_zbot_b200.sharding import aggregate_throughput, rank_seed, shard_range

    dev = torch.device("cuda", local)
    S, n_env, opts = spec.S, spec.n_env, spec.opts()
    push_cfg = PushEventConfig() if spec.randomized else None
    if spec.workload == "c5":
        push_cfg = PushEventConfig(x_linvel=1.0, y_linvel=1.0, z_linvel=0.5, interval_range=(0.2, 0.4))
    # with the observation epilogue the network inputs come from the kernel; the remaining mjx.Data outputs are the ones the rewards /
    # terminations / Trajectory read (train.py:297, 476, 503-546, 681-733: base / body poses, touch + force sensors, contact flags)
    obs_outputs = ("xpos", "xquat", "qacc", "actuator_force", "sensordata", "site_xpos", "contact_mask", "solver_niter", "flags")
    eng = ZbotCudaEngine(model, n_env, device=local, opts=opts, ctrl_dt=spec.ctrl_dt, push=push_cfg, seed=rank_seed(spec.seed, rank),
                         outputs=obs_outputs if spec.obs else None)
    params_np = M.randomize_env_params(model, n_env, spec.seed + 100 + rank) if spec.randomized else None
    ep = EnvParams(n_env, dev, **params_np) if spec.randomized else None
    g = torch.Generator(device=dev)
    g.manual_seed(rank_seed(spec.seed, rank) + 17)
    bias = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=dev)
    actions = bias + 0.3 * (torch.rand((K + W, n_env, M.NU), generator=g, device=dev) * 2 - 1)
    pushes = torch.stack([eng.sample_push(n_env) for _ in range(K + W)]) if spec.randomized else None
    rq, rv, rt = eng.sample_reset(n_env)
    if spec.workload == "c5":  # tilted + dropped reset states
        ax = torch.randn((n_env, 3), generator=g, device=dev)
        ax = ax / ax.norm(dim=1, keepdim=True)
        ang = (torch.rand((n_env,), generator=g, device=dev) * 2 - 1) * math.radians(40)
        rq[:, 3] = torch.cos(ang / 2)
        rq[:, 4:7] = ax * torch.sin(ang / 2)[:, None]
        rq[:, 2] += 0.08
    obs = eng.new_observations(n_env, rewards=True) if spec.obs else None
    obs_lag = eng.sample_imu_lag(n_env, (0.0, 0.1)) if spec.obs else None  # lag of re-initialised environments (pre-sampled)
    ps = eng.reset(ep, reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt, obs=obs, reset_imu_lag=obs_lag)
    done = torch.zeros((n_env,), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_sink = torch.zeros((), dtype=torch.int64, device=dev)
    n_done = torch.zeros((), dtype=torch.int64, device=dev)
    niter_acc = torch.zeros((), dtype=torch.float64, device=dev)

    def ctrl_step(t, act):
        # ONE launch: S fused substeps + termination check + partial reset (+ observation epilogue)
        eng.step(act, ep, ps, push_rand=None if pushes is None else pushes[t], auto_reset=(rq, rv, rt), done_out=done, obs=obs,
                 reset_imu_lag=obs_lag)

    for t in range(W):
        ctrl_step(t, actions[t])
    torch.cuda.synchronize()

    # ---------------- timed region 1: inputs resident in HBM ----------------
    clocks = Clocks(local)
    if rank == 0 and primary:
        clocks.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    launches0 = eng.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for i in range(K):
        t = W + i
        flush.zero_()  # L2 flush, outside the event pair
        if spec.flush == "write+read":
            flush_sink.add_(flush.view(torch.int64).sum())  # read pass: the memset's dirty lines are written back HERE, not during the step
        ev[i][0].record()
        ctrl_step(t, actions[t])
        ev[i][1].record()
        ev[i][2].record()
        n_done += (done != 0).sum()
        niter_acc += ps.data.solver_niter.sum()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count() - launches0
    step_ms = sum(e[0].elapsed_time(e[1]) for e in ev)
    total_ms = sum(e[0].elapsed_time(e[2]) for e in ev)
    value, total_ms_max = aggregate_throughput(n_env * S * K, total_ms, dev)  # sum of units / max over ranks of device time
    niter_gpu = float(niter_acc.item()) / (K * n_env)  # solver iterations of the LAST substep of every timed control step
    ncon_mean = float(sum(((ps.data.contact_mask >> c) & 1).float().mean().item() for c in range(8)))
    resets_per_step = float(n_done.item()) / K
    # state + action the FLOP sample is taken from: the end of the timed region (a developed rollout, not the just-reset pose)
    ns = min(64, n_env)
    snap = {"qpos": ps.data.qpos[:ns].cpu().numpy(), "qvel": ps.data.qvel[:ns].cpu().numpy(),
            "qacc_warmstart": ps.data.qacc_warmstart[:ns].cpu().numpy(), "act_pos": ps.actuator_state.position[:ns].cpu().numpy(),
            "act_vel": ps.actuator_state.velocity[:ns].cpu().numpy(), "act_torque": ps.actuator_state.last_computed_torque[:ns].cpu().numpy(),
            "push_t_rem": np.full(ns, 1e9, np.float32), "time": np.zeros(ns, np.float32)}
    snap_action = actions[W + K - 1, :ns].cpu().numpy()

    # ---------------- timed region 2: end to end through the public API with HOST buffers ----------------
    # Every control step: H2D of that step's actions from pinned host memory -> engine step (one launch incl. termination + reset +
    # observations) -> D2H of the step's results into pinned host memory.  The host waits, before it issues the next step, for what
    # the next action depends on (`gate`: the actor's input vector + the done mask, packed into ONE buffer = one copy); what only the
    # critic / rewards / Trajectory consume (`bulk`: critic input vector, reward inputs, qpos / qvel / body poses / sensors) goes to
    # one of two device buffers in turn and streams out on a copy stream while the next step computes (--e2e-blocking waits for it
    # as well).  {launch -> state bookkeeping -> gating D2H} is ONE CUDA graph per bulk buffer, so a control step costs the host one
    # cudaMemcpyAsync (actions) + one cudaGraphLaunch + one stream synchronize.
    if spec.obs:
        gate_fields = ("actor_obs",)
        bulk_fields = ("critic_obs", "reward_terms", "feet_pos_obs", "qpos", "qvel", "xpos", "xquat", "sensordata")
    else:
        gate_fields = ("qpos", "qvel", "sensordata", "act_torque")
        bulk_fields = ("xpos", "xquat", "cinert", "cvel", "qacc", "actuator_force")

    def build_shards(nsh):
        shards = []
        for k in range(nsh):
            lo, hi = shard_range(n_env, nsh, k)
            m_ = hi - lo
            sp = eng.new_state(m_)
            for f in ("qpos", "qvel", "qacc_warmstart", "time"):
                getattr(sp.data, f).copy_(getattr(ps.data, f)[lo:hi])
            for f in ("position", "velocity", "last_computed_torque"):
                getattr(sp.actuator_state, f).copy_(getattr(ps.actuator_state, f)[lo:hi])
            sp.event_states["push_t_rem"].copy_(ps.event_states["push_t_rem"][lo:hi])
            sob = None
            if spec.obs:
                sob = eng.new_observations(m_, rewards=True)
                sob.imu_x.copy_(obs.imu_x[lo:hi]); sob.imu_lag.copy_(obs.imu_lag[lo:hi])

            def owner(f, sp=sp, sob=sob):  # (object, attribute) holding array f
                if f == "act_torque":
                    return sp.actuator_state, "last_computed_torque"
                return (sob, f) if (f.endswith("_obs") or f == "reward_terms") else (sp.data, f)

            def pack(fields, extra_bytes=0):  # one flat device buffer with a view per field (initialised from the current arrays)
                sizes = {f: getattr(*owner(f)).numel() for f in fields}
                flat = torch.empty(sum(sizes.values()) * 4 + extra_bytes, dtype=torch.uint8, device=dev)
                views, off = {}, 0
                for f, nel in sizes.items():
                    cur = getattr(*owner(f))
                    views[f] = flat[off:off + 4 * nel].view(torch.float32).view(cur.shape)
                    views[f].copy_(cur)
                    off += 4 * nel
                return flat, views, off

            g_flat, g_views, g_off = pack(gate_fields, extra_bytes=m_)  # ... + the done mask (uint8 [m]) at the end
            done_k = g_flat[g_off:g_off + m_]
            done_k.zero_()
            for f, v in g_views.items():
                setattr(*owner(f), v)
            packs = [pack(bulk_fields) for _ in range(2)]
            b_flat, b_views = [p[0] for p in packs], [p[1] for p in packs]
            ep_k = None if ep is None else ep.slice(lo, hi)
            rq_k, rv_k, rt_k = rq[lo:hi].contiguous(), rv[lo:hi].contiguous(), rt[lo:hi].contiguous()
            lag_k = None if obs_lag is None else obs_lag[lo:hi].contiguous()
            d_act = torch.empty((m_, M.NU), dtype=torch.float32, device=dev)
            d_push = None if pushes is None else torch.empty((m_, 7), dtype=torch.float32, device=dev)
            h_gate = torch.empty_like(g_flat, device="cpu").pin_memory()
            stream = torch.cuda.Stream(device=dev)
            plans, graphs = [], []
            for b in range(2):  # one pre-bound control step per bulk buffer (engine.plan_step: argument structs built once)
                for f, v in b_views[b].items():
                    setattr(*owner(f), v)
                plans.append(eng.plan_step(sp, ep_k, auto_reset=(rq_k, rv_k, rt_k), done_out=done_k, obs=sob, reset_imu_lag=lag_k))
            if not a.e2e_no_graph:
                for b in range(2):
                    with torch.cuda.stream(stream):  # warm-up outside capture, then capture {launch, bookkeeping, gating D2H}
                        gr = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gr, stream=stream):
                            plans[b](d_act, d_push)
                            h_gate.copy_(g_flat, non_blocking=True)
                    graphs.append(gr)
            shards.append(dict(lo=lo, hi=hi, ps=sp, plans=plans, graphs=graphs, g_flat=g_flat, h_gate=h_gate, b_flat=b_flat,
                               h_bulk=[torch.empty_like(b, device="cpu").pin_memory() for b in b_flat], d_act=d_act, d_push=d_push,
                               stream=stream, copy_stream=torch.cuda.Stream(device=dev),
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
        def bulk_out(sh, b):  # stream bulk buffer b to the host once its kernel has finished
            with torch.cuda.stream(sh["copy_stream"]):
                sh["copy_stream"].wait_event(sh["ev_kernel"][b])
                sh["h_bulk"][b].copy_(sh["b_flat"][b], non_blocking=True)
                sh["ev_copy"][b].record(sh["copy_stream"])

        for i in range(steps):
            t, b = W + i, i & 1
            for sh in shards:
                with torch.cuda.stream(sh["stream"]):
                    if i >= 2 and not a.e2e_no_bulk:
                        sh["stream"].wait_event(sh["ev_copy"][b])  # bulk buffer b has been read out
                    sh["d_act"].copy_(h_act[i, sh["lo"]:sh["hi"]], non_blocking=True)
                    if pushes is not None:
                        sh["d_push"].copy_(pushes[t, sh["lo"]:sh["hi"]], non_blocking=True)
                    if sh["graphs"]:
                        sh["graphs"][b].replay()
                    else:
                        sh["plans"][b](sh["d_act"], sh["d_push"])  # this step's bulk arrays go to buffer b (plan b is bound to it)
                        sh["h_gate"].copy_(sh["g_flat"], non_blocking=True)
                    sh["ev_kernel"][b].record(sh["stream"])
                    if a.e2e_blocking:
                        sh["h_bulk"][b].copy_(sh["b_flat"][b], non_blocking=True)
                        sh["ev_copy"][b].record(sh["stream"])
            # The bulk read-back of step i-1 is enqueued only NOW, behind the launch of step i: measured (scripts/e2e_probe2.py), a
            # stream synchronize does not return before a D2H copy that was enqueued earlier on ANOTHER stream has finished (its
            # completion notification queues behind that copy) — enqueued before the synchronize of its own step, the 12 MB copy
            # added its full 0.22 ms to every step although nothing on the device waited for it.
            if i >= 1 and not a.e2e_blocking and not a.e2e_no_bulk:
                for sh in shards:
                    bulk_out(sh, b ^ 1)
            for sh in shards:
                sh["stream"].synchronize()  # the caller has the gating result of step i before it issues step i+1
        if steps >= 1 and not a.e2e_blocking and not a.e2e_no_bulk:
            for sh in shards:
                bulk_out(sh, (steps - 1) & 1)
        for sh in shards:
            main.wait_stream(sh["stream"])
            main.wait_stream(sh["copy_stream"])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    nsh = max(1, min(a.e2e_shards, n_env // 16))
    shards = build_shards(nsh)
    launches_e2e0 = eng.launch_count()
    run_e2e(shards, min(3, K))  # untimed: first replays, pinned pages touched
    h2d = h_act[0].numel() * 4
    d2h_gate = sum(sh["h_gate"].numel() for sh in shards)
    d2h = d2h_gate + sum(sh["h_bulk"][0].numel() for sh in shards)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_ms = run_e2e(shards, K)
    e2e_value, _ = aggregate_throughput(n_env * S * K, e2e_ms, dev)
    clk = clocks.stop() if (rank == 0 and primary) else None
    res = dict(spec=spec, value=value, ms_per_step=total_ms_max / K, step_ms=step_ms / K, share=step_ms / total_ms, launches=int(launches),
               e2e={"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "d2h_gating_bytes_per_step": d2h if a.e2e_blocking else d2h_gate, "ms_per_step": e2e_ms / K,
                    "readback": "blocking" if a.e2e_blocking else "gate (actor input + done mask, one packed copy) blocking; critic input, reward inputs and trajectory arrays double-buffered on a copy stream",
                    "host_calls_per_step": "1 cudaMemcpyAsync (actions) + 1 cudaGraphLaunch {step kernel, state bookkeeping, gating D2H} + 1 stream synchronize (+ the streamed bulk copy)"
                                           if not a.e2e_no_graph else "plain enqueues", "streams": nsh},
               clocks=clk, niter_gpu=niter_gpu, ncon=ncon_mean, resets=resets_per_step, snap=snap, snap_action=snap_action,
               params_np=params_np, eng=eng, dev=dev)
    del shards
    return res


def fp32_peak_probe(eng, dev):
    """FP32 FFMA peak (roofline denominator; MEASURED_PEAKS.json has HBM GB/s and bf16 tensor TF/s only)."""
    import ctypes

    import torch

    scratch = torch.empty(148 * 16 * 256, dtype=torch.float32, device=dev)
    best = 0.0
    for _ in range(4):
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        eng.lib.zb_bench_ffma(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.c_void_p(scratch.data_ptr()), 148 * 16, 20000)
        p1.record()
        torch.cuda.synchronize()
        best = max(best, 148 * 16 * 256 * 20000 * 32 / (p0.elapsed_time(p1) * 1e-3) / 1e12)
    return best


def roofline_of(res, model, fp32_peak, hbm_peak, have_peaks):
    spec = res["spec"]
    sp = {k: v[:64] for k, v in res["params_np"].items()} if res["params_np"] else None
    stage, niter_cpu = oracle_flops_per_env_step(model, spec.opts(), res["snap"], res["snap_action"], spec.S, sp)
    # the sample's solver ran niter_cpu iterations on average, the GPU niter_gpu over the timed region: the solver's share is scaled
    # to the GPU's count so that `achieved` counts only iterations the kernel executed
    scale = (res["niter_gpu"] / niter_cpu) if niter_cpu > 0 else 1.0
    stage_adj = dict(stage)
    stage_adj["solver"] = stage["solver"] * scale
    flops_per = float(sum(stage_adj.values()))
    step_s = res["step_ms"] * 1e-3
    n_env, S = spec.n_env, spec.S
    bytes_ctrl = algorithmic_bytes_per_ctrl_step(spec.randomized, spec.randomized, spec.obs)
    traffic, traffic_note = ncu_traffic(spec)
    ach = flops_per * n_env * S / step_s / 1e12
    return {
        "bound": "fp32", "unit": "TFLOP/s", "kernel": "zb_kernel<ZB_MODE_STEP>", "achieved": ach, "peak": fp32_peak,
        "peak_source": "measured in this run (FFMA probe zb_bench_ffma; MEASURED_PEAKS.json has no FP32 figure)", "frac": ach / fp32_peak,
        "traffic": traffic, "traffic_note": traffic_note + "; algorithmic bytes per launch = %d" % (bytes_ctrl * n_env),
        "flops_per_env_step": flops_per,
        "flops_source": "oracle op counters (fp32, 64-env sample of the state at the end of the timed region), solver share scaled by "
                        "GPU / sample iteration count = %.3f / %.3f" % (res["niter_gpu"], niter_cpu),
        "kernel_ms_per_launch": res["step_ms"], "kernel_share_of_step": res["share"],
        "hbm": {"bound": "hbm", "unit": "GB/s", "achieved": bytes_ctrl * n_env / step_s / 1e9, "peak": hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if have_peaks else "fallback", "frac": bytes_ctrl * n_env / step_s / 1e9 / hbm_peak,
                "bytes_per_env_ctrl_step": bytes_ctrl},
    }, stage_adj, niter_cpu


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import ksim_gym_zbot_b200  # noqa: F401
    from ksim_gym_zbot_b200 import model as M

    model = M.build_zbot_syn()
    spec = Spec(a)

    if a.impl == "reference":
        if rank == 0:
            run_reference(a, spec, model, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the physics step has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = a.steps, a.warmup
    res = measure(a, spec, model, K, W, rank, world, local, primary=True)
    fp32_peak = fp32_peak_probe(res["eng"], dev)

    # short runs of the other single-GPU configurations (same kernel, same timing rules), so that the driver's record carries them too
    extras = []
    if not a.no_extras and a.workload == "c2" and not a.envs_per_gpu and a.substeps == 4:
        Ke, We = min(K, 40), max(3, min(W, 10))
        for sp_ in (Spec(a, "c2", 8192, 4), Spec(a, "c3", 16384, 4), Spec(a, "c2", 4096, 20)):
            if sp_.S == 20:
                Ke = min(Ke, 20)
            extras.append(measure(a, sp_, model, Ke, We, rank, world, local, primary=False) | {"K": Ke, "W": We})
    rollout_actor = None
    if not a.no_extras and a.workload == "c2" and a.substeps == 4:
        rollout_actor = measure_rollout_actor(a, spec, model, rank, world, local)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roof, stage, niter_cpu = roofline_of(res, model, fp32_peak, hbm_peak, bool(peaks))

    # ---------------- CPU baseline (rank 0, N=1 only): the SAME closed loop as `--impl reference`, bounded in time ----------------
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        arm = CpuArm(spec, model, cores)
        arm.run_steps(2)
        k, secs = arm.run_seconds(a.cpu_baseline_seconds)
        cpu = {"value": spec.n_env * spec.S * k / secs, "unit": UNIT, **arm.describe(k, secs)}

    others = []
    for r in extras:
        rf, _, _ = roofline_of(r, model, fp32_peak, hbm_peak, bool(peaks))
        s_ = r["spec"]
        others.append({"workload": s_.short(), "config": s_.config(world), "value": r["value"], "unit": UNIT, "steps": r["K"], "warmup": r["W"],
                       "ms_per_step": r["ms_per_step"], "e2e": {k: r["e2e"][k] for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "ms_per_step")},
                       "roofline_frac": rf["frac"], "flops_per_env_step": rf["flops_per_env_step"], "solver_niter_mean": r["niter_gpu"],
                       "resets_per_step": r["resets"], "gpu_launches": r["launches"]})
    line = {
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": spec.config(world),
        "e2e": res["e2e"], "gpu_launches": res["launches"], "clocks": res["clocks"], "roofline": roof, "cpu_baseline": cpu,
        "stats": {"solver_niter_mean": res["niter_gpu"], "active_contacts_mean_last_step": res["ncon"], "resets_per_step": res["resets"],
                  "ctrl_steps_per_s": res["value"] / spec.S, "solver_niter_mean_flop_sample": niter_cpu, "flops_per_stage": stage},
        "other_workloads": others,
    }
    if rollout_actor is not None:
        a_ = rollout_actor["actor"]
        a_["achieved_tflops_note"] = "useful fp32-equivalent FLOPs (2 x MACs of the three matrix stages) per second; the 3xTF32 path executes 3x the tensor-core MACs"
        a_["fp32_cuda_core_kernel_frac_of_fp32_peak"] = (spec.n_env * a_["flops_per_env"] / (a_["kernel_ms_fp32_cuda_cores"] * 1e-3) / 1e12) / fp32_peak if fp32_peak else None
        line["rollout_actor"] = rollout_actor
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



def measure_rollout_actor(a, spec: "Spec", model, rank: int, world: int, local: int, T: int = 200, chunks: int = 3):
    """BASELINE.json configs[1] with the reference's actor in the loop: per control step zb_actor_f32 (random-init
    weights of the reference architecture, train.py:885-968, sampled actions) + the physics launch writing the Trajectory; the 2T launches of
    a T = 200 step chunk (rollout_length 4.0 s / ctrl_dt, train.py:1775) replay as ONE CUDA graph.  Timed: `chunks` replays with CUDA
    events, max over ranks.  The actor kernel is also timed alone (L2 flushed) for its FP32 roofline fraction."""
    import torch
    import torch.distributed as dist

    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import ZbotCudaEngine
    from ksim_gym_zbot_b200.policy import ActorParams, ActorPolicy
    from ksim_gym_zbot_b200.rollout import Rollout
    from ksim_gym_zbot_b200.sharding import aggregate_throughput, rank_seed

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import zb_policy_oracle as PO

    dev = torch.device("cuda", local)
    n = spec.n_env
    eng = ZbotCudaEngine(model, n, device=local, opts=spec.opts(), ctrl_dt=spec.ctrl_dt, seed=rank_seed(spec.seed, rank) + 5)
    ro = Rollout(eng, T)
    from ksim_gym_zbot_b200.policy import PRECISION_FP32, PRECISION_TF32X3

    prm = ActorParams.random(seed=spec.seed)
    pol = ActorPolicy(prm, n, dev, precision=PRECISION_TF32X3)  # GRU stack on the tcgen05 tensor cores (3xTF32: fp32-level accuracy)
    launches0 = eng.launch_count()
    ro.reset()
    ro.unroll_actor(pol, argmax=False)  # capture + first replay (warm-up chunk)
    ro.next_chunk()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(chunks):
        ro.sample_noise()
        ro.unroll_actor(pol, argmax=False)
        ro.next_chunk()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    value, ms_max = aggregate_throughput(n * spec.S * T * chunks, ms, dev)
    resets = float((ro.traj.done != 0).sum().item()) / T
    # the actor kernel alone, both precisions (L2 flushed between launches)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    actor_ms = {}
    for name, pr in (("tf32x3_tcgen05", PRECISION_TF32X3), ("fp32_cuda_cores", PRECISION_FP32)):
        p2 = pol if pr == PRECISION_TF32X3 else ActorPolicy(prm, n, dev, precision=pr)
        for _ in range(3):
            p2.enqueue(ro.traj.actor_obs[0], p2.action)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(20)]
        for x, y in ev:
            flush.zero_()
            x.record(); p2.enqueue(ro.traj.actor_obs[0], p2.action); y.record()
        torch.cuda.synchronize()
        actor_ms[name] = float(np.median([x.elapsed_time(y) for x, y in ev]))
    eng.close()
    return {"workload": f"c2 {n} envs/GPU S={spec.S} + actor in the loop", "value": value, "unit": UNIT, "chunks": chunks, "chunk_steps": T,
            "ms_per_ctrl_step": ms_max / (T * chunks), "launches_per_ctrl_step": 2, "graph": "one CUDA graph per chunk (2T launches + 1 copy)",
            "actions": "sampled from the mixture (pre-sampled U / N numbers per chunk), random-init weights", "resets_per_step": resets,
            "trajectory_bytes_per_env_ctrl_step": ro.traj.bytes_per_env_step(),
            "actor": {"kernel_ms": actor_ms["tf32x3_tcgen05"], "kernel_ms_fp32_cuda_cores": actor_ms["fp32_cuda_cores"], "flops_per_env": PO.flops_per_env(),
                      "achieved_tflops": n * PO.flops_per_env() / (actor_ms["tf32x3_tcgen05"] * 1e-3) / 1e12,
                      "dtype": "GRU stack: 3xTF32 tcgen05.mma (fp32-level accuracy, accumulators in TMEM); input / output projections: f32 CUDA cores",
                      "ctas": (n + 31) // 32}}

def run_reference(a, spec: Spec, model, world: int):
    """The reference arm: the reference's algorithm for this path on the host cores (CpuArm), full per-GPU batch, K timed control steps."""
    cores = os.cpu_count() or 1
    arm = CpuArm(spec, model, cores)
    K, W = a.steps, a.warmup
    arm.run_steps(W)
    dt = arm.run_steps(K)
    value = spec.n_env * spec.S * K / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dt * 1e3 / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": spec.config(world),
        "cpu_baseline": {"value": value, "unit": UNIT, **arm.describe(K, dt)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
