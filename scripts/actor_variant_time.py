import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import ksim_gym_zbot_b200  # noqa
from ksim_gym_zbot_b200.policy import ActorParams, ActorPolicy
n = 4096
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
obs = torch.randn((n, 50), device="cuda")
for name in sys.argv[1:]:
    lib = os.path.join(ROOT, "ksim-gym-zbot_b200/csrc/variants", f"libzbphys_{name}.so")
    for prec in (0, 1):
        pol = ActorPolicy(ActorParams.random(0), n, "cuda:0", lib_path=lib, precision=prec)
        for _ in range(3): pol(obs)
        res = {}
        for fl in (True, False):
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(20)]
            for a, b in ev:
                if fl: flush.zero_()
                a.record(); pol(obs); b.record()
            torch.cuda.synchronize()
            res["flushed" if fl else "warm L2"] = round(float(np.median([a.elapsed_time(b) for a, b in ev])), 4)
        print(name, "precision", prec, res, flush=True)
