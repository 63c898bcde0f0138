"""Shared input generators for the parity tests (seeded, identical for the oracle and the CUDA path)."""

from __future__ import annotations

import math

import numpy as np

import ksim_gym_zbot_b200  # noqa: F401
from ksim_gym_zbot_b200 import model as M


def random_states(m, n: int, seed: int, joint_noise: float = 0.25, vel: float = 0.5, z_jitter: float = 0.004,
                  tilt: float = 0.05):
    """qpos/qvel around the JOINT_BIASES standing pose: some feet in contact, some in the air."""
    r = np.random.default_rng(seed)
    q = np.repeat(M.default_qpos(m)[None], n, 0)
    q[:, 7:] += r.uniform(-joint_noise, joint_noise, (n, M.NU))
    q[:, 2] += r.uniform(-z_jitter, z_jitter, n)
    q[:, 0:2] += r.uniform(-0.5, 0.5, (n, 2))
    ax = r.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    ang = r.uniform(-tilt, tilt, n)
    q[:, 3] = np.cos(ang / 2)
    q[:, 4:7] = ax * np.sin(ang / 2)[:, None]
    q[:, 3:7] *= r.uniform(0.9, 1.1, (n, 1))  # un-normalised on purpose: kinematics must normalise
    v = r.uniform(-vel, vel, (n, M.NV))
    return q.astype(np.float32), v.astype(np.float32)


def random_actions(n: int, seed: int, amp: float = 0.3):
    r = np.random.default_rng(seed)
    return (M.BIAS_POSE[None] + amp * r.uniform(-1, 1, (n, M.NU))).astype(np.float32)


def quat_from_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    return np.concatenate([[math.cos(angle / 2)], axis * math.sin(angle / 2)])


def rel_err(a, b, floor=1.0):
    """max |a-b| / max(floor, |b|) — 'relative with an absolute floor'."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(floor, np.abs(b)))) if a.size else 0.0
