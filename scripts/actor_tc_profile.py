import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ksim_gym_zbot_b200  # noqa
from ksim_gym_zbot_b200.policy import ActorParams, ActorPolicy
n = 4096
lib = os.path.join(ROOT, "ksim-gym-zbot_b200/csrc/variants/libzbphys_tcprof.so")
pol = ActorPolicy(ActorParams.random(0), n, "cuda:0", lib_path=lib, precision=1)
obs = torch.randn((n, 50), device="cuda")
dist = torch.zeros((n, 3, 20, 5), device="cuda")
names = ["-", "layer prologue", "wait tile/conv", "wait mma(g-2)", "convert", "arrive", "issue/refill", "wait layer", "epilogue", "out epilogue+head", "setup", "obs load", "wait in_w", "in_proj", "TOTAL"]
for rep in range(3):
    pol.enqueue(obs, pol.action, dist_out=dist)
    torch.cuda.synchronize()
    d = dist.view(-1)[:32].cpu().numpy()
    for who, off in (("thread 0", 0), ("thread 200", 16)):
        print(rep, who, {k: int(v) for k, v in zip(names, d[off:off + 15])}, flush=True)
