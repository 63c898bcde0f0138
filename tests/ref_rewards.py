"""numpy transcription of the reward list the reference task uses (/root/reference/train.py:1546-1586, `get_rewards`) and of its
terminations (train.py:1588-1593), evaluated on the per-control-step fields a rollout records.  TEST INFRASTRUCTURE (checker only).

Provenance of each term:
  [V] code is in the reference (file:line given)           [R] ksim 0.1.99 built-in, formula recalled (un-vendored: uv.lock:1872-1873)

The [R] formulas cannot be checked against ksim here; for the parity question asked of them ("do the CUDA arm and the oracle arm
produce the same reward stream?") that does not matter: both arms go through this same function, so any fixed, smooth function of
(qpos, qvel) tests the same thing.  Scales and arguments are the reference's (train.py:1548-1585).
"""
from __future__ import annotations

import numpy as np

# (joint_name, reference_angle_rad, weight) — train.py:59-82
JOINT_BIASES = [
    ("right_hip_yaw", 0.0, 1.0), ("right_hip_roll", -0.1, 1.0), ("right_hip_pitch", -0.4, 0.01), ("right_knee_pitch", -0.8, 0.01),
    ("right_ankle_pitch", -0.4, 0.01), ("right_ankle_roll", -0.1, 0.01), ("left_hip_yaw", 0.0, 1.0), ("left_hip_roll", 0.1, 1.0),
    ("left_hip_pitch", -0.4, 0.01), ("left_knee_pitch", -0.8, 0.01), ("left_ankle_pitch", -0.4, 0.01), ("left_ankle_roll", 0.1, 0.01),
    ("left_shoulder_pitch", 0.0, 1.0), ("left_shoulder_roll", 0.2, 1.0), ("left_elbow_roll", -0.2, 1.0), ("left_gripper_roll", 0.0, 1.0),
    ("right_shoulder_pitch", 0.0, 1.0), ("right_shoulder_roll", -0.2, 1.0), ("right_elbow_roll", 0.2, 1.0), ("right_gripper_roll", 0.0, 1.0),
]
_NAMES = [n for n, _, _ in JOINT_BIASES]
STRAIGHT_LEG = ["left_hip_roll", "left_hip_yaw", "right_hip_roll", "right_hip_yaw"]                      # train.py:585-590
ANKLE_KNEE = ["left_knee_pitch", "left_ankle_pitch", "left_ankle_roll", "right_knee_pitch", "right_ankle_pitch", "right_ankle_roll"]  # :606-613
ARM_POSE = ["left_shoulder_pitch", "left_shoulder_roll", "left_elbow_roll", "left_gripper_roll", "right_shoulder_pitch",
            "right_shoulder_roll", "right_elbow_roll", "right_gripper_roll"]                             # :632-641

# name -> (scale, kind); kind "flag": the per-step value is a threshold decision (0/1), "cont": a smooth function of the state
SCALES = {  # train.py:1548-1585
    "stay_alive": (1.0, "flag"), "upright": (1.0, "cont"), "naive_forward": (5.0, "cont"), "naive_forward_orientation": (0.3, "cont"),
    "linvel_y_penalty": (-2.0, "cont"), "single_foot_contact": (0.3, "flag"), "feet_airtime": (2.5, "air"), "feet_orientation": (0.3, "cont"),
    "feet_too_close": (-0.5, "flag"), "straight_leg": (-0.5, "cont"), "ankle_knee": (-0.05, "cont"), "arm_pose": (-2.0, "cont"),
}


def _unit(q):  # xax normalises as q / (|q| + 1e-6) [R]
    return q / (np.linalg.norm(q, axis=-1, keepdims=True) + q.dtype.type(1e-6))


def rotate_inverse(v, q):  # xax.rotate_vector_by_quat(v, q, inverse=True) [R]
    q = _unit(q)
    w, u = q[..., 0:1], -q[..., 1:]
    t = 2 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


def quat_to_euler(q):  # xax.quat_to_euler [R]
    q = _unit(q)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    roll = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    sp = 2 * (w * y - z * x)
    pitch = np.where(np.abs(sp) >= 1, np.sign(sp) * np.pi / 2, np.arcsin(np.clip(sp, -1, 1)))
    return np.stack([roll, pitch], axis=-1)


def _joint_deviation(qpos, names):  # ksim.JointDeviationPenalty [R]: sum_j w_j (q_j - target_j)^2, targets / weights train.py:559-562
    idx = [7 + _NAMES.index(n) for n in names]
    tgt = np.array([JOINT_BIASES[_NAMES.index(n)][1] for n in names], qpos.dtype)
    w = np.array([JOINT_BIASES[_NAMES.index(n)][2] for n in names], qpos.dtype)
    return (w * (qpos[..., idx] - tgt) ** 2).sum(-1)


def terminations(qpos, time=None, flags=None):
    """train.py:1588-1593: BadZTermination(0.05, 0.5) | NotUprightTermination(60 deg) | EpisodeLengthTermination(80 s)."""
    up = qpos[..., 3] ** 2 - qpos[..., 4] ** 2 - qpos[..., 5] ** 2 + qpos[..., 6] ** 2
    d = (qpos[..., 2] < 0.05) | (qpos[..., 2] > 0.5) | (up < 0.5) | ~np.isfinite(qpos[..., 2]) | ~np.isfinite(up)
    if time is not None:
        d = d | (time > 80.0)
    if flags is not None:
        d = d | (flags != 0)
    return d


def step_terms(f, left_body, right_body):
    """Per-control-step reward terms (everything but the stateful FeetAirtimeReward) from one step's fields:
    f = dict(qpos [n,27], qvel [n,26], xquat [n,26,4], site_xpos [n,3,3], sensordata [n,18], done [n] bool)."""
    qpos, qvel = f["qpos"], f["qvel"]
    quat = qpos[:, 3:7]
    one = qpos.dtype.type(1)
    t = {}
    t["stay_alive"] = np.where(f["done"], -one, one / 10)                                        # ksim.StayAliveReward [R] (balance = 10)
    t["upright"] = rotate_inverse(np.array([0, 0, 1], qpos.dtype), quat)[:, 2]                   # ksim.UprightReward [R]
    t["naive_forward"] = np.minimum(qvel[:, 0], qpos.dtype.type(0.2))                            # ksim.NaiveForwardReward(clip_max=0.2) [R]
    fwd = rotate_inverse(np.array([1, 0, 0], qpos.dtype), quat)
    t["naive_forward_orientation"] = fwd[:, 0] - np.linalg.norm(fwd[:, 1:], axis=-1)             # ksim.NaiveForwardOrientationReward [R]
    t["linvel_y_penalty"] = np.abs(rotate_inverse(qvel[:, :3], quat)[:, 1])                      # ksim.LinearVelocityPenalty(y, robot frame, l1) [R]
    lc, rc = f["sensordata"][:, 10] > 0.1, f["sensordata"][:, 11] > 0.1                          # [V] train.py:699-703
    t["left_contact"], t["right_contact"] = lc, rc
    t["single_foot_contact"] = np.logical_xor(lc, rc).astype(qpos.dtype)                         # [V] train.py:704-711 (stand_still_threshold None)
    err = np.abs(quat_to_euler(f["xquat"][:, left_body])).sum(-1) + np.abs(quat_to_euler(f["xquat"][:, right_body])).sum(-1)
    t["feet_orientation"] = np.exp(-err / 0.25)                                                  # [V] train.py:680-689
    fl, fr = rotate_inverse(f["site_xpos"][:, 1], quat), rotate_inverse(f["site_xpos"][:, 2], quat)  # [V] train.py:452-465
    t["feet_dist"] = np.linalg.norm(fl - fr, axis=-1)
    t["feet_too_close"] = (t["feet_dist"] < 0.12).astype(qpos.dtype)                             # [V] train.py:740-748
    t["straight_leg"] = _joint_deviation(qpos, STRAIGHT_LEG)                                     # [V] names train.py:585-590, base class [R]
    t["ankle_knee"] = _joint_deviation(qpos, ANKLE_KNEE)                                         # [V] names train.py:606-613
    t["arm_pose"] = _joint_deviation(qpos, ARM_POSE)                                             # [V] names train.py:632-641
    return t


def feet_airtime(left_contact, right_contact, done, carry, ctrl_dt=0.02, touchdown_penalty=0.3):
    """FeetAirtimeReward.get_reward_stateful, train.py:503-546, for [T, n] boolean arrays (the oracle module holds the same
    restatement: zb_oracle.feet_airtime; this copy keeps the checker free of oracle/ imports)."""
    T, n = left_contact.shape
    time = np.asarray(carry, np.float32).copy()
    air = np.zeros((T, n, 2), np.float32)
    for t in range(T):
        cod = np.stack([left_contact[t] | done[t], right_contact[t] | done[t]], -1)
        time = np.where(cod, np.float32(0), time + np.float32(ctrl_dt)).astype(np.float32)
        air[t] = time
    c = np.stack([left_contact, right_contact], -1)
    prev = np.concatenate([np.zeros((1, n, 2), bool), c[:-1]], 0)
    td = (c & ~prev).astype(np.float32)
    r = (np.roll(air, 1, axis=0) - np.float32(touchdown_penalty)) * td
    return r.sum(-1), time


def total_reward(terms):
    """sum_i scale_i * term_i over the reference's reward list; terms: {name: [T, n]}."""
    return sum(s * np.asarray(terms[k], np.float64) for k, (s, _) in SCALES.items())
