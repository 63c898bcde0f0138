"""Where does an end-to-end control step spend its time?  (diagnostic for bench.py's timed region 2)"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch
import ksim_gym_zbot_b200  # noqa
from ksim_gym_zbot_b200 import model as M
from ksim_gym_zbot_b200.engine import ZbotCudaEngine

n, S = 4096, 4
dev = torch.device("cuda", 0)
eng = ZbotCudaEngine(M.build_zbot_syn(), n, device=0, opts=M.make_opts(dt=0.005), ctrl_dt=0.02)
rq, rv, rt = eng.sample_reset(n)
ps = eng.reset(reset_qpos=rq, reset_qvel=rv, reset_t_rem=rt)
done = torch.zeros(n, dtype=torch.uint8, device=dev)
act = torch.tensor(M.BIAS_POSE, dtype=torch.float32, device=dev).expand(n, 20).contiguous()
h_act = act.cpu().pin_memory()
big = torch.empty(10_500_000 // 4, device=dev); h_big = torch.empty_like(big, device="cpu").pin_memory()
small = torch.empty(1_500_000 // 4, device=dev); h_small = torch.empty_like(small, device="cpu").pin_memory()
s_main, s_copy = torch.cuda.Stream(), torch.cuda.Stream()
E = lambda: torch.cuda.Event(enable_timing=True)

def loop(mode, steps=100):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evk = [torch.cuda.Event() for _ in range(2)]
    host_issue = 0.0
    for i in range(steps):
        h0 = time.perf_counter()
        with torch.cuda.stream(s_main):
            act.copy_(h_act, non_blocking=True)
            eng.step(act, None, ps, auto_reset=(rq, rv, rt), done_out=done)
            h_small.copy_(small, non_blocking=True)
            evk[i & 1].record(s_main)  # AFTER the gating copy: both D2H copies share one DMA queue, the bulk copy must not get in front
            if mode == "blocking":
                h_big.copy_(big, non_blocking=True)
        if mode == "pipelined":
            with torch.cuda.stream(s_copy):
                s_copy.wait_event(evk[i & 1])
                h_big.copy_(big, non_blocking=True)
        host_issue += time.perf_counter() - h0
        s_main.synchronize()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps * 1e3
    print(f"{mode:10s} {dt:.3f} ms/step wall, host issue {host_issue / steps * 1e3:.3f} ms/step")

for mode in ("kernel", "blocking", "pipelined", "kernel", "pipelined", "blocking"):
    loop(mode)
# copy speeds
for t, h, name in ((big, h_big, "10.5 MB"), (small, h_small, "1.5 MB")):
    e0, e1 = E(), E()
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): h.copy_(t, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(name, "D2H", e0.elapsed_time(e1) / 20, "ms")
