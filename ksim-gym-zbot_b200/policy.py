"""Policy inference in the rollout loop: the reference's ``Actor`` (/root/reference/train.py:885-968) on the GPU through
``zb_actor_f32`` (csrc/zb_policy.cuh) — SURVEY.md §8f rank 4.

``ActorParams`` holds the weights in the equinox layout the reference's checkpoint has (``input_proj.weight [128, 50]``,
``rnns.i.weight_ih [384, 128]`` …) and packs them k-major for the kernel; ``ActorPolicy`` mirrors ``sample_action``
(train.py:1737-1763): it owns the actor half of the model carry ``[N, depth, 128]`` (train.py:1731-1735) and turns the ``[N, 50]``
vector the observation epilogue wrote into the next action, with ``argmax`` (mode) or pre-sampled noise.  There is no CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import model as M
from .engine import NUM_ACTOR_INPUTS, ZbError, load_library

HIDDEN = 128  # ZB_ACTOR_HIDDEN (train.py:1065-1068)


class ZbActor(ctypes.Structure):
    _fields_ = [("n_in", ctypes.c_int32), ("depth", ctypes.c_int32), ("num_mixtures", ctypes.c_int32), ("min_std", ctypes.c_float),
                ("max_std", ctypes.c_float), ("var_scale", ctypes.c_float), ("precision", ctypes.c_int32), ("reserved", ctypes.c_int32)] + \
               [(k, ctypes.c_void_p) for k in ("in_wt", "in_b", "gru_wih_t", "gru_whh_t", "gru_b", "gru_bn", "out_wt", "out_b", "mean_bias", "gru_w_tc", "out_w_tc")]


PRECISION_FP32, PRECISION_TF32X3 = 0, 1  # ZB_ACTOR_FP32 / ZB_ACTOR_TF32X3


class ActorParams:
    """Weights of ``Actor`` in the equinox layout (numpy), plus the static fields of train.py:891-896."""

    def __init__(self, arrays: dict[str, np.ndarray], depth: int = 5, num_mixtures: int = 5, min_std: float = 0.01, max_std: float = 1.0,
                 var_scale: float = 1.0, mean_bias: np.ndarray | None = None):
        self.depth, self.num_mixtures = int(depth), int(num_mixtures)
        self.min_std, self.max_std, self.var_scale = float(min_std), float(max_std), float(var_scale)
        self.arrays = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in arrays.items()}
        self.n_in = self.arrays["input_proj.weight"].shape[1]
        self.mean_bias = np.asarray(M.BIAS_POSE if mean_bias is None else mean_bias, dtype=np.float32)  # JOINT_BIASES, train.py:59-82, 963
        n_out = M.NU * 3 * self.num_mixtures
        want = {"input_proj.weight": (HIDDEN, self.n_in), "input_proj.bias": (HIDDEN,), "output_proj.weight": (n_out, HIDDEN), "output_proj.bias": (n_out,)}
        for i in range(self.depth):
            want.update({f"rnns.{i}.weight_ih": (3 * HIDDEN, HIDDEN), f"rnns.{i}.weight_hh": (3 * HIDDEN, HIDDEN), f"rnns.{i}.bias": (3 * HIDDEN,),
                         f"rnns.{i}.bias_n": (HIDDEN,)})
        for k, shp in want.items():
            if k not in self.arrays or self.arrays[k].shape != shp:
                raise ValueError(f"actor weight '{k}': expected shape {shp}, got {self.arrays[k].shape if k in self.arrays else None}")

    @staticmethod
    def random(seed: int = 0, n_in: int = NUM_ACTOR_INPUTS, depth: int = 5, num_mixtures: int = 5, **kw) -> "ActorParams":
        """Random-init weights of the architecture (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) as eqx.nn.Linear / GRUCell draw them [R]); the
        reference re-uses one key for all GRU cells (train.py:920-931) — here every layer gets its own numbers so that a layer mix-up shows."""
        g = np.random.default_rng(seed)
        u = lambda shape, fan: g.uniform(-1, 1, shape).astype(np.float32) / np.float32(np.sqrt(fan))  # noqa: E731
        a = {"input_proj.weight": u((HIDDEN, n_in), n_in), "input_proj.bias": u((HIDDEN,), n_in),
             "output_proj.weight": u((M.NU * 3 * num_mixtures, HIDDEN), HIDDEN), "output_proj.bias": u((M.NU * 3 * num_mixtures,), HIDDEN)}
        for i in range(depth):
            a.update({f"rnns.{i}.weight_ih": u((3 * HIDDEN, HIDDEN), HIDDEN), f"rnns.{i}.weight_hh": u((3 * HIDDEN, HIDDEN), HIDDEN),
                      f"rnns.{i}.bias": u((3 * HIDDEN,), HIDDEN), f"rnns.{i}.bias_n": u((HIDDEN,), HIDDEN)})
        return ActorParams(a, depth=depth, num_mixtures=num_mixtures, **kw)

    def oracle_dict(self) -> dict:
        return dict(self.arrays, min_std=self.min_std, max_std=self.max_std, var_scale=self.var_scale, mean_bias=self.mean_bias)

    def pack(self) -> dict[str, np.ndarray]:
        """Kernel layout (include/zbphys.h ZbActor): every matrix k-major, the three gates of a GRU matrix interleaved per k."""
        a, H = self.arrays, HIDDEN
        gate_t = lambda w: np.ascontiguousarray(w.reshape(3, H, H).transpose(2, 0, 1))  # noqa: E731  [g*H+f][k] -> [k][g][f]

        def tc(i):  # tensor-core operand tiles: [k-tile t][matrix m][k-chunk c][f][j] = W_m[f][8 t + 4 c + j]  (ZbActor.gru_w_tc)
            w6 = np.concatenate([a[f"rnns.{i}.weight_ih"].reshape(3, H, H), a[f"rnns.{i}.weight_hh"].reshape(3, H, H)])  # [m][f][k]
            return np.ascontiguousarray(w6.reshape(6, H, H // 8, 2, 4).transpose(2, 0, 3, 1, 4))                            # [t][m][c][f][j]

        wo = np.zeros((3 * H, H), dtype=np.float32)  # output_proj rows padded to 3 M-tiles of 128
        wo[:a["output_proj.weight"].shape[0]] = a["output_proj.weight"]
        out_tc = np.ascontiguousarray(wo.reshape(3, H, H // 8, 2, 4).transpose(2, 0, 3, 1, 4))  # [t][mt][c][row][j]
        return {"gru_w_tc": np.stack([tc(i) for i in range(self.depth)]), "out_w_tc": out_tc,"in_wt": np.ascontiguousarray(a["input_proj.weight"].T), "in_b": a["input_proj.bias"],
                "gru_wih_t": np.stack([gate_t(a[f"rnns.{i}.weight_ih"]) for i in range(self.depth)]),
                "gru_whh_t": np.stack([gate_t(a[f"rnns.{i}.weight_hh"]) for i in range(self.depth)]),
                "gru_b": np.stack([a[f"rnns.{i}.bias"].reshape(3, H) for i in range(self.depth)]),
                "gru_bn": np.stack([a[f"rnns.{i}.bias_n"] for i in range(self.depth)]),
                "out_wt": np.ascontiguousarray(a["output_proj.weight"].T), "out_b": a["output_proj.bias"], "mean_bias": self.mean_bias}


class ActorPolicy:
    """``sample_action`` (train.py:1737-1763) for a batch of environments: ``policy(actor_obs) -> action``."""

    def __init__(self, params: ActorParams, n_envs: int, device, lib_path: str | None = None, precision: int = PRECISION_FP32):
        if not torch.cuda.is_available():
            raise RuntimeError("ActorPolicy needs a CUDA device: there is no CPU path for the policy kernel")
        self.lib = load_library(lib_path)
        self.lib.zb_actor_f32.argtypes = [ctypes.c_void_p] * 10 + [ctypes.c_int]
        assert self.lib.zb_sizeof_actor() == ctypes.sizeof(ZbActor), "ZbActor layout mismatch between policy.py and zbphys.h"
        self.params, self.n, self.device = params, int(n_envs), torch.device(device)
        self.weights = {k: torch.tensor(v, device=self.device) for k, v in params.pack().items()}
        self.desc = ZbActor(params.n_in, params.depth, params.num_mixtures, params.min_std, params.max_std, params.var_scale, int(precision), 0,
                            *[self.weights[k].data_ptr() for k in ("in_wt", "in_b", "gru_wih_t", "gru_whh_t", "gru_b", "gru_bn", "out_wt", "out_b", "mean_bias",
                                                                   "gru_w_tc", "out_w_tc")])
        self.carry = torch.zeros((self.n, params.depth, HIDDEN), dtype=torch.float32, device=self.device)  # get_initial_model_carry, train.py:1731
        self.action = torch.zeros((self.n, M.NU), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):  # n_envs = 0: sets the kernel's shared-memory attribute, so that later launches can be graph-captured
            z = ctypes.c_void_p(self.carry.data_ptr())
            rc = self.lib.zb_actor_f32(None, ctypes.byref(self.desc), z, z, z, None, None, None, z, None, 0)
        if rc != 0:
            raise ZbError(rc, self.lib.zb_last_error().decode())

    def reset_carry(self) -> None:
        self.carry.zero_()

    def enqueue(self, actor_obs: torch.Tensor, action_out: torch.Tensor, carry_reset: torch.Tensor | None = None, noise_u: torch.Tensor | None = None,
                noise_eps: torch.Tensor | None = None, dist_out: torch.Tensor | None = None) -> None:
        """One launch on the current stream (graph-capturable: no allocation, no synchronisation).  ``carry_reset`` [N] uint8: the
        previous step's ``done``; ``noise_*`` [N, 20] pre-sampled (None = mode)."""
        for t in (actor_obs, action_out, carry_reset, noise_u, noise_eps, dist_out):
            if t is not None:
                assert t.is_cuda and t.is_contiguous() and t.device == self.device
        assert actor_obs.shape == (self.n, self.params.n_in) and action_out.shape == (self.n, M.NU)
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(self.device):  # the launch goes to the policy's device whatever the caller's current device is
            rc = self.lib.zb_actor_f32(ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream), ctypes.byref(self.desc), p(actor_obs),
                                       p(self.carry), p(self.carry), p(carry_reset), p(noise_u), p(noise_eps), p(action_out), p(dist_out), self.n)
        if rc != 0:
            raise ZbError(rc, self.lib.zb_last_error().decode())

    def __call__(self, actor_obs: torch.Tensor, carry_reset: torch.Tensor | None = None, noise: tuple[torch.Tensor, torch.Tensor] | None = None) -> torch.Tensor:
        self.enqueue(actor_obs, self.action, carry_reset, *(noise if noise is not None else (None, None)))
        return self.action
