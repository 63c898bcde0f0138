"""CPU restatement of the reference's Actor forward pass + action selection — TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's
cpu_baseline leg may import this; the product path is csrc/zb_policy.cuh through zb_actor_f32).

Follows /root/reference/train.py:
  Actor.forward                      train.py:946-968   (input_proj -> GRUCell stack -> output_proj -> mixture-of-Gaussians parameters)
  Actor.__init__ shapes              train.py:898-944   (Linear(50,128), depth x GRUCell(128,128), Linear(128, 20*3*M))
  Model / get_model hyper-parameters train.py:1026-1057, 1604-1614 (min_std 0.01, max_std 1.0, var_scale 1.0, hidden 128, depth 5, M 5)
  sample_action                      train.py:1737-1763 (mode() when argmax else sample())
  initial carry                      train.py:1731-1735 (zeros [depth, hidden])
eqx.nn.Linear / eqx.nn.GRUCell are equinox (un-vendored dependency, /root/reference/uv.lock) — restated from their published
definitions [R]:  Linear: y = W x + b.  GRUCell: ig = W_ih x + bias; hg = W_hh h; r = sigmoid(ig_r + hg_r); z = sigmoid(ig_z + hg_z);
n = tanh(ig_n + r * (hg_n + bias_n)); h' = n + z * (h - n)  (PyTorch's gate order r, z, n).  tests/test_policy.py pins this file
against torch.nn.Linear / torch.nn.GRUCell (an independent implementation of the same published cell).
ksim.MixtureOfGaussians (un-vendored) [R]: Categorical(logits) over M Normal(mean, std) components per joint; `mode` = the mean of the
most probable component, `sample` = component drawn from softmax(logits), then mean_k + std_k * eps.  The kernel takes the random
numbers from the caller (u in [0,1) by inverse CDF, eps ~ N(0,1)); parity unpinned against ksim for these two choices
(PARITY_RISKS.md).
"""
from __future__ import annotations

import numpy as np

NUM_JOINTS = 20


def softplus(x):
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_cell(w_ih, w_hh, bias, bias_n, x, h):
    """eqx.nn.GRUCell.__call__ for a batch: x [N, in], h [N, H]."""
    H = h.shape[-1]
    ig = x @ w_ih.T + bias
    hg = h @ w_hh.T
    r = sigmoid(ig[:, :H] + hg[:, :H])
    z = sigmoid(ig[:, H:2 * H] + hg[:, H:2 * H])
    n = np.tanh(ig[:, 2 * H:] + r * (hg[:, 2 * H:] + bias_n))
    return n + z * (h - n)


def actor_forward(p: dict, obs: np.ndarray, carry: np.ndarray, dtype=np.float64):
    """Actor.forward (train.py:946-968) for a batch.  p: equinox-layout arrays (input_proj.weight [128,50], .bias; rnns[i].weight_ih
    [384,128], .weight_hh, .bias [384], .bias_n [128]; output_proj.weight [20*3*M,128], .bias) + min_std / max_std / var_scale /
    mean_bias [20].  Returns (means [N,20,M], stds, logits, next_carry [N,depth,128])."""
    c = lambda a: np.asarray(a, dtype=dtype)  # noqa: E731
    x = c(obs) @ c(p["input_proj.weight"]).T + c(p["input_proj.bias"])
    carry = c(carry)
    out_c = []
    for i in range(carry.shape[1]):
        x = gru_cell(c(p[f"rnns.{i}.weight_ih"]), c(p[f"rnns.{i}.weight_hh"]), c(p[f"rnns.{i}.bias"]), c(p[f"rnns.{i}.bias_n"]), x, carry[:, i])
        out_c.append(x)
    out = x @ c(p["output_proj.weight"]).T + c(p["output_proj.bias"])
    M = out.shape[-1] // (3 * NUM_JOINTS)
    sl = NUM_JOINTS * M
    mean = out[:, :sl].reshape(-1, NUM_JOINTS, M) + c(p["mean_bias"])[None, :, None]
    std = np.minimum((softplus(out[:, sl:2 * sl].reshape(-1, NUM_JOINTS, M)) + dtype(p["min_std"])) * dtype(p["var_scale"]), dtype(p["max_std"]))
    logits = out[:, 2 * sl:].reshape(-1, NUM_JOINTS, M)
    return mean, std, logits, np.stack(out_c, axis=1)


def select_action(mean, std, logits, noise_u=None, noise_eps=None):
    """mode (noise None) or sample with caller-supplied noise: component = inverse CDF of softmax(logits) at u, mean_k + std_k * eps."""
    if noise_u is None:
        k = np.argmax(logits, axis=-1)
        return np.take_along_axis(mean, k[..., None], -1)[..., 0]
    w = np.exp(logits - logits.max(-1, keepdims=True))
    cum = np.cumsum(w, -1)
    u = np.asarray(noise_u, dtype=w.dtype)[..., None] * cum[..., -1:]
    k = np.minimum((u >= cum).sum(-1), logits.shape[-1] - 1)
    return np.take_along_axis(mean, k[..., None], -1)[..., 0] + np.take_along_axis(std, k[..., None], -1)[..., 0] * np.asarray(noise_eps, dtype=w.dtype)


def flops_per_env(n_in=50, hidden=128, depth=5, num_mixtures=5) -> int:
    """Multiply-adds x 2 of the three matrix stages of one actor evaluation (activations not counted)."""
    return 2 * (n_in * hidden + depth * 2 * 3 * hidden * hidden + hidden * NUM_JOINTS * 3 * num_mixtures)
