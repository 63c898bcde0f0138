#!/usr/bin/env python
"""bench_c4.py — BASELINE.json configs[3]: rollout + PPO update with the NCCL gradient all-reduce.

    python bench_c4.py --gpus 1 --steps 2 --warmup 1
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 bench_c4.py --gpus 8      # 8 x 8192 = 65 536 envs

What is the product here and what is a stand-in:
  * the physics (S fused substeps + termination + partial reset + BOTH network input vectors, one launch per control step) is
    the CUDA path of this repository (zb_step_obs_f32) — environments are sharded, no collective inside the step;
  * the actor / critic (Linear -> 5 x GRU-128 -> Linear; 539 308 / 556 289 parameters with equinox's bias layout,
    /root/reference/train.py:885-1057) and the PPO update (clipped surrogate + value loss, num_passes = 4, 4 minibatches,
    train.py:1770-1773) are a plain-PyTorch STAND-IN for the reference's JAX trainer, which SURVEY.md §8 puts out of scope: JAX /
    ksim are not installable here.  Simplifications: cuDNN GRU over the whole rollout without mid-sequence carry resets, a toy
    reward (upright + height) — the FLOPs, the launch pattern and the 4.4 MB gradient all-reduce per minibatch are the same;
  * the only collective of the whole job is that all-reduce (`torch.distributed` NCCL, flat fp32 gradient bucket).

One "step" = one PPO iteration: a rollout of T control steps of the whole batch + the update.  Prints one JSON line with
env-steps/s (physics substeps of all ranks / max-over-ranks device time) and the split physics / policy / update / all-reduce.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_JOINTS, NUM_MIX, HID, DEPTH = 20, 5, 128, 5  # train.py:27, 1065-1075


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2, help="timed PPO iterations")
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--envs-per-gpu", type=int, default=8192)
    ap.add_argument("--rollout-steps", type=int, default=200, help="control steps per rollout (4.0 s / 0.02 s, train.py:1775)")
    ap.add_argument("--substeps", type=int, default=20, help="dt = 0.001 (train.py:1777)")
    ap.add_argument("--num-passes", type=int, default=4)
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist
    import torch.nn as nn

    import ksim_gym_zbot_b200  # noqa: F401
    from ksim_gym_zbot_b200 import model as M
    from ksim_gym_zbot_b200.engine import EnvParams, PushEventConfig, ZbotCudaEngine
    from ksim_gym_zbot_b200.sharding import rank_seed

    if not torch.cuda.is_available():
        raise SystemExit("bench_c4.py needs a CUDA device")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True

    n, S, T = a.envs_per_gpu, a.substeps, a.rollout_steps
    model = M.build_zbot_syn()
    opts = M.make_opts(dt=0.02 / S)
    eng = ZbotCudaEngine(model, n, device=local, opts=opts, ctrl_dt=0.02, push=PushEventConfig(), seed=rank_seed(a.seed, rank),
                         outputs=("xpos", "sensordata", "contact_mask", "solver_niter", "flags"))
    ep = EnvParams(n, dev, **M.randomize_env_params(model, n, a.seed + 100 + rank))

    class Net(nn.Module):
        def __init__(self, nin, nout):
            super().__init__()
            self.inp, self.rnn, self.out = nn.Linear(nin, HID), nn.GRU(HID, HID, num_layers=DEPTH), nn.Linear(HID, nout)

        def forward(self, x, h):  # x [T, N, nin], h [DEPTH, N, HID]: the whole rollout through cuDNN (update)
            y, h = self.rnn(self.inp(x), h)
            return self.out(y), h

        def step(self, x, h):  # one control step, x [N, nin]: plain GEMMs + gate math with the same weights (rollout)
            x = self.inp(x)
            hs = []
            for l in range(DEPTH):
                gi = torch.addmm(getattr(self.rnn, f"bias_ih_l{l}"), x, getattr(self.rnn, f"weight_ih_l{l}").t())
                gh = torch.addmm(getattr(self.rnn, f"bias_hh_l{l}"), h[l], getattr(self.rnn, f"weight_hh_l{l}").t())
                r, z = torch.sigmoid(gi[:, :HID] + gh[:, :HID]), torch.sigmoid(gi[:, HID:2 * HID] + gh[:, HID:2 * HID])
                c = torch.tanh(gi[:, 2 * HID:] + r * gh[:, 2 * HID:])
                x = (1 - z) * c + z * h[l]
                hs.append(x)
            return self.out(x), torch.stack(hs)

    torch.manual_seed(1234)  # identical initial weights on every rank
    actor, critic = Net(50, NUM_JOINTS * 3 * NUM_MIX).to(dev), Net(484, 1).to(dev)
    params = list(actor.parameters()) + list(critic.parameters())
    n_params = sum(p.numel() for p in params)
    optim = torch.optim.AdamW(params, lr=1e-3, weight_decay=1e-5)  # train.py:1769, 1084-1087
    flat_grad = torch.zeros(n_params, device=dev)
    bias = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=dev)

    def mixture(out):  # train.py:955-968: means, stds, logits per joint and mixture
        m_, s_, l_ = out.split(NUM_JOINTS * NUM_MIX, dim=-1)
        shp = out.shape[:-1] + (NUM_JOINTS, NUM_MIX)
        std = 0.01 + (1.0 - 0.01) * torch.sigmoid(s_.reshape(shp))
        return m_.reshape(shp) + bias[:, None], std, torch.log_softmax(l_.reshape(shp), dim=-1)

    def log_prob(out, act):
        mean, std, logw = mixture(out)
        z = (act[..., None] - mean) / std
        return torch.logsumexp(logw - 0.5 * z * z - torch.log(std) - 0.5 * math.log(2 * math.pi), dim=-1).sum(-1)

    def sample(out, gen):
        mean, std, logw = mixture(out)
        u = torch.rand(logw.shape, device=dev, generator=gen).clamp_(1e-9, 1.0 - 1e-7)
        k = torch.argmax(logw - torch.log(-torch.log(u)), dim=-1, keepdim=True)  # Gumbel-max categorical sample (no host sync)
        mu, sd = mean.gather(-1, k)[..., 0], std.gather(-1, k)[..., 0]
        return mu + sd * torch.randn(mu.shape, device=dev, generator=gen)  # gen=None: default generator (CUDA-graph safe)

    gen = torch.Generator(device=dev)
    gen.manual_seed(rank_seed(a.seed, rank) + 5)
    rq, rv, rt = eng.sample_reset(n)
    obs = eng.new_observations(n)
    ps = eng.reset(ep, reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt, obs=obs)
    done = torch.zeros((n,), dtype=torch.uint8, device=dev)
    # trajectory storage (device resident, like ksim's Trajectory pytree)
    tr = dict(actor=torch.empty((T, n, 50), device=dev), critic=torch.empty((T, n, 484), device=dev), act=torch.empty((T, n, NUM_JOINTS), device=dev),
              logp=torch.empty((T, n), device=dev), rew=torch.empty((T, n), device=dev), done=torch.empty((T, n), device=dev))
    pushes = [eng.sample_push(n) for _ in range(8)]
    lag = eng.sample_imu_lag(n, (0.0, 0.1))
    ha = torch.zeros((DEPTH, n, HID), device=dev)

    # policy forward + sampling + log-prob of one control step as ONE CUDA graph (static buffers): the stand-in's ~60 small torch
    # kernels per step would otherwise be launch-bound (8 ms per step eager vs the 6.5 ms physics launch)
    torch.manual_seed(rank_seed(a.seed, rank) + 5)  # default generator: graph-safe Philox state
    g_act, g_logp = torch.zeros((n, NUM_JOINTS), device=dev), torch.zeros((n,), device=dev)

    def policy_step():
        nonlocal ha
        out, h = actor.step(obs.actor_obs, ha)
        act = sample(out, None)
        g_act.copy_(act); g_logp.copy_(log_prob(out, act)); ha.copy_(h)

    with torch.no_grad():
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                policy_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            policy_step()
        ha.zero_()

    E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    t_phys = t_pol = t_upd = t_ar = 0.0
    launches0 = 0

    def iteration(timed: bool):
        nonlocal ha, t_phys, t_pol, t_upd, t_ar
        ev = []
        h0 = ha.clone()
        with torch.no_grad():
            for t in range(T):
                e0, e1, e2 = E(), E(), E()
                e0.record()
                tr["actor"][t].copy_(obs.actor_obs); tr["critic"][t].copy_(obs.critic_obs)
                graph.replay()
                tr["act"][t].copy_(g_act); tr["logp"][t].copy_(g_logp)
                e1.record()
                eng.step(g_act, ep, ps, push_rand=pushes[t % 8], auto_reset=(rq, rv, rt), done_out=done, obs=obs, reset_imu_lag=lag)
                e2.record()
                d = done.float()
                # toy reward from the fields the reference's rewards read (upright + height, train.py:469-483, 389-413)
                up = 1.0 - 2.0 * (ps.data.qpos[:, 4] ** 2 + ps.data.qpos[:, 5] ** 2)
                tr["rew"][t].copy_(up + torch.exp(-torch.abs(ps.data.xpos[:, 1, 2] - 0.27) / 0.25) - d)
                tr["done"][t].copy_(d)
                ha.mul_((1.0 - d)[None, :, None])
                ev.append((e0, e1, e2))
        u0, u1 = E(), E()
        u0.record()
        ar_ev = []
        # GAE with the pre-update critic
        with torch.no_grad():
            v, _ = critic(tr["critic"], torch.zeros_like(h0))
            v = v[..., 0]
            adv = torch.zeros_like(v)
            last = torch.zeros((n,), device=dev)
            for t in range(T - 1, -1, -1):
                nv = v[t + 1] if t + 1 < T else v[t]
                delta = tr["rew"][t] + 0.97 * nv * (1 - tr["done"][t]) - v[t]
                last = delta + 0.97 * 0.95 * (1 - tr["done"][t]) * last
                adv[t] = last
            ret = adv + v
            adv = (adv - adv.mean()) / (adv.std() + 1e-6)
        mb = n // a.minibatches
        for _ in range(a.num_passes):
            for k in range(a.minibatches):
                sl = slice(k * mb, (k + 1) * mb)
                out, _ = actor(tr["actor"][:, sl], h0[:, sl].contiguous())
                val, _ = critic(tr["critic"][:, sl], torch.zeros((DEPTH, mb, HID), device=dev))
                ratio = torch.exp(log_prob(out, tr["act"][:, sl]) - tr["logp"][:, sl])
                pg = -torch.minimum(ratio * adv[:, sl], torch.clamp(ratio, 0.8, 1.2) * adv[:, sl]).mean()
                loss = pg + 0.5 * ((val[..., 0] - ret[:, sl]) ** 2).mean()
                optim.zero_grad(set_to_none=False)
                loss.backward()
                # the job's ONLY collective: mean of the flat fp32 gradient over the ranks (NCCL over NVLink / NVSwitch)
                a0, a1 = E(), E()
                off = 0
                for p in params:
                    flat_grad[off:off + p.numel()].copy_(p.grad.reshape(-1)); off += p.numel()
                a0.record()
                if world > 1:
                    dist.all_reduce(flat_grad)
                    flat_grad.div_(world)
                a1.record()
                off = 0
                for p in params:
                    p.grad.copy_(flat_grad[off:off + p.numel()].view_as(p.grad)); off += p.numel()
                ar_ev.append((a0, a1))
                torch.nn.utils.clip_grad_norm_(params, 2.0)  # train.py:1080
                optim.step()
        u1.record()
        if timed:
            torch.cuda.synchronize()
            t_pol += sum(e[0].elapsed_time(e[1]) for e in ev)
            t_phys += sum(e[1].elapsed_time(e[2]) for e in ev)
            t_upd += u0.elapsed_time(u1)
            t_ar += sum(x[0].elapsed_time(x[1]) for x in ar_ev)

    for _ in range(a.warmup):
        iteration(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = eng.launch_count()
    s0, s1 = E(), E()
    s0.record()
    for _ in range(a.steps):
        iteration(True)
    s1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([s0.elapsed_time(s1)], device=dev)
    units = torch.tensor([float(n * S * T * a.steps)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(units)
    if rank == 0:
        K = a.steps
        line = {
            "metric": "zbot_ppo_iteration_env_steps_per_sec", "value": float(units.item()) / (float(ms.item()) * 1e-3), "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": a.warmup, "ms_per_step": float(ms.item()) / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 physics, tf32 networks", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[3]: rollout + PPO update, envs sharded, NCCL gradient all-reduce (trainer = PyTorch stand-in)",
                       "envs_per_gpu": n, "rollout_steps": T, "substeps": S, "num_passes": a.num_passes, "minibatches": a.minibatches,
                       "parameters": n_params, "parallelism": f"dp{world}"},
            "split_ms_per_iteration": {"physics_incl_obs_epilogue": t_phys / K, "policy_forward_and_sampling": t_pol / K, "ppo_update": t_upd / K,
                                       "allreduce_inside_update": t_ar / K, "allreduce_calls": a.num_passes * a.minibatches,
                                       "allreduce_bytes": 4 * n_params},
            "gpu_launches": int(eng.launch_count() - launches0),
            "physics_env_steps_per_sec_inside_rollout": n * S * T * world / (t_phys / K * 1e-3) if t_phys > 0 else None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
