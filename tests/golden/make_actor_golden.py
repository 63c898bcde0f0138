#!/usr/bin/env python
"""Golden vectors for the actor (tests/golden/actor_golden.npz): the reference's Actor.forward (/root/reference/train.py:946-968) evaluated by an
implementation that shares no code with this repository's restatement or kernels — torch.nn.Linear / torch.nn.GRUCell on the CPU in float64
(equinox's GRUCell == torch's with bias_ih = bias and bias_hh = [0, 0, bias_n]; equinox itself is not installable here), three recurrent
steps with a carry restart in between.  Weights come from ActorParams.random(seed=21) so that the test can rebuild them.

    python tests/golden/make_actor_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ksim_gym_zbot_b200  # noqa: E402,F401
from ksim_gym_zbot_b200 import model as M  # noqa: E402
from ksim_gym_zbot_b200.policy import HIDDEN, ActorParams  # noqa: E402

SEED, N, T = 21, 16, 3


def main():
    prm = ActorParams.random(seed=SEED)
    a = prm.arrays
    g = np.random.default_rng(SEED + 1)
    obs = g.normal(0, 1, (T, N, 50)).astype(np.float32)
    carry0 = g.uniform(-1, 1, (N, prm.depth, HIDDEN)).astype(np.float32)
    reset = (g.uniform(size=(T, N)) < 0.2).astype(np.uint8)  # reset[t]: environment starts step t from the initial carry
    reset[0] = 0
    with torch.no_grad():
        lin_in, lin_out = torch.nn.Linear(50, HIDDEN).double(), torch.nn.Linear(HIDDEN, 300).double()
        lin_in.weight.copy_(torch.tensor(a["input_proj.weight"])); lin_in.bias.copy_(torch.tensor(a["input_proj.bias"]))
        lin_out.weight.copy_(torch.tensor(a["output_proj.weight"])); lin_out.bias.copy_(torch.tensor(a["output_proj.bias"]))
        cells = []
        for i in range(prm.depth):
            c = torch.nn.GRUCell(HIDDEN, HIDDEN).double()
            c.weight_ih.copy_(torch.tensor(a[f"rnns.{i}.weight_ih"])); c.weight_hh.copy_(torch.tensor(a[f"rnns.{i}.weight_hh"]))
            c.bias_ih.copy_(torch.tensor(a[f"rnns.{i}.bias"]))
            c.bias_hh.zero_(); c.bias_hh[2 * HIDDEN:].copy_(torch.tensor(a[f"rnns.{i}.bias_n"]))
            cells.append(c)
        carry = torch.tensor(carry0).double()
        means, stds, logits, carries = [], [], [], []
        for t in range(T):
            carry = torch.where(torch.tensor(reset[t] != 0)[:, None, None], torch.zeros_like(carry), carry)
            x = lin_in(torch.tensor(obs[t]).double())
            nxt = []
            for i, c in enumerate(cells):
                x = c(x, carry[:, i])
                nxt.append(x)
            carry = torch.stack(nxt, 1)
            out = lin_out(x)
            means.append((out[:, :100].reshape(N, 20, 5) + torch.tensor(M.BIAS_POSE.astype(np.float32)).double()[None, :, None]).numpy())
            stds.append(torch.clamp((torch.nn.functional.softplus(out[:, 100:200]) + 0.01) * 1.0, max=1.0).reshape(N, 20, 5).numpy())
            logits.append(out[:, 200:].reshape(N, 20, 5).numpy())
            carries.append(carry.numpy())
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "actor_golden.npz"), seed=SEED, obs=obs, carry0=carry0, reset=reset,
                        means=np.stack(means).astype(np.float32), stds=np.stack(stds).astype(np.float32), logits=np.stack(logits).astype(np.float32),
                        carries=np.stack(carries).astype(np.float32))
    print("wrote actor_golden.npz")


if __name__ == "__main__":
    main()
