"""Where does the streamed bulk read-back cost time in bench.py's end-to-end loop?  Per-step device events around the pieces.
usage: python scripts/e2e_probe2.py"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import numpy as np
import torch
import ksim_gym_zbot_b200  # noqa
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.engine import ZbotCudaEngine

n, S = 4096, 4
dev = torch.device("cuda", 0)
eng = ZbotCudaEngine(M.build_zbot_syn(), n, device=0, opts=M.make_opts(dt=0.005), ctrl_dt=0.02,
                     outputs=("xpos", "xquat", "qacc", "actuator_force", "sensordata", "site_xpos", "contact_mask", "solver_niter", "flags"))
rq, rv, rt = eng.sample_reset(n)
obs = eng.new_observations(n, rewards=True)
lag = eng.sample_imu_lag(n, (0.0, 0.1))
ps = eng.reset(reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt, obs=obs, reset_imu_lag=lag)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
acts = (torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=dev) + 0.3 * (torch.rand((64, n, 20), generator=g, device=dev) * 2 - 1))
h_acts = acts.cpu().pin_memory()
act = torch.empty((n, 20), device=dev)
plan = eng.plan_step(ps, None, auto_reset=(rq, rv, rt), done_out=done, obs=obs, reset_imu_lag=lag, track_action=False)
gate = torch.empty(823296 // 4, device=dev); h_gate = torch.empty_like(gate, device="cpu").pin_memory()
s_main, s_copy = torch.cuda.Stream(), torch.cuda.Stream()
s_hi = torch.cuda.Stream(priority=-1)
E = lambda: torch.cuda.Event(enable_timing=True)


def loop(bulk_mb, chunks=1, steps=100, copy_stream=None, tag="", zero_copy=False):
    big = torch.empty(int(bulk_mb * 1e6) // 4, device=dev) if bulk_mb else None
    h_big = torch.empty_like(big, device="cpu").pin_memory() if bulk_mb else None
    cs = copy_stream or s_copy
    torch.cuda.synchronize()
    ev = [[E() for _ in range(4)] for _ in range(steps)]
    evk = [torch.cuda.Event() for _ in range(2)]
    t0 = time.perf_counter()
    t_enq, t_bulk, t_sync = 0.0, 0.0, 0.0
    for i in range(steps):
        h0 = time.perf_counter()
        with torch.cuda.stream(s_main):
            ev[i][0].record(s_main)
            if not zero_copy:
                act.copy_(h_acts[i % 64], non_blocking=True)
            ev[i][1].record(s_main)
            plan(h_acts[i % 64] if zero_copy else act, None)  # zero_copy: the kernel reads the pinned host buffer itself
            ev[i][2].record(s_main)
            h_gate.copy_(gate, non_blocking=True)
            ev[i][3].record(s_main)
            evk[i & 1].record(s_main)
        h1 = time.perf_counter()
        if big is not None:
            with torch.cuda.stream(cs):
                cs.wait_event(evk[i & 1])
                if chunks == 1:
                    h_big.copy_(big, non_blocking=True)
                else:
                    m = big.numel() // chunks
                    for c in range(chunks):
                        h_big[c * m:(c + 1) * m].copy_(big[c * m:(c + 1) * m], non_blocking=True)
        h2 = time.perf_counter()
        s_main.synchronize()
        h3 = time.perf_counter()
        t_enq += h1 - h0; t_bulk += h2 - h1; t_sync += h3 - h2
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps * 1e3
    sk = slice(10, None)
    h2d = np.mean([e[0].elapsed_time(e[1]) for e in ev[sk]])
    ker = np.mean([e[1].elapsed_time(e[2]) for e in ev[sk]])
    gt = np.mean([e[2].elapsed_time(e[3]) for e in ev[sk]])
    gap = np.mean([ev[i][3].elapsed_time(ev[i + 1][0]) for i in range(10, steps - 1)])
    print(f"{tag:28s} bulk {bulk_mb:5.1f} MB x{chunks}: {dt:.4f} ms/step wall | h2d {h2d:.4f} kernel {ker:.4f} gate {gt:.4f} gap(host) {gap:.4f} | host: enqueue {t_enq / steps * 1e3:.4f} bulk-enqueue {t_bulk / steps * 1e3:.4f} sync {t_sync / steps * 1e3:.4f}")


for _ in range(2):
    loop(0, tag="no bulk")
    loop(12.3, tag="bulk 12.3")
    loop(12.3, tag="bulk 12.3 hi-prio copy stream", copy_stream=s_hi)
