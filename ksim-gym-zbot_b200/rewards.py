"""The reference task's reward list over a rollout chunk, as batched torch tensor ops on the device.

Reference: ``ZbotWalkingTask.get_rewards`` (/root/reference/train.py:1546-1586) evaluated on a ``ksim.Trajectory``.  Here the
trajectory is the time-major ``rollout.Trajectory`` the CUDA engine wrote.  The observation-based terms (contact flags, feet
distance) come from the step launch's observation epilogue (``ZbObs.reward_terms``), the scan over time from
``zb_feet_airtime_f32``; the rest are a handful of elementwise ops on ``qpos`` / ``qvel`` / ``xquat``.

Provenance: [V] = the term's code is in the reference (file:line), [R] = ksim 0.1.99 built-in whose formula is recalled (ksim is
un-vendored: /root/reference/uv.lock:1872-1873) — see PARITY_RISKS.md "ksim built-in rewards".  All terms are RAW per-step values;
``total`` applies the reference's ``scale`` factors (curriculum scaling of the three joint penalties is the caller's).
"""
from __future__ import annotations

import torch

from . import model as M

# (joint weights of JOINT_BIASES, train.py:59-82: 0.01 for hip_pitch / knee / ankle joints, 1.0 for the rest)
_JOINT_WEIGHT = {n: (0.01 if any(k in n for k in ("hip_pitch", "knee_pitch", "ankle_pitch", "ankle_roll")) else 1.0) for n in M.JOINT_NAMES}
STRAIGHT_LEG = ("left_hip_roll", "left_hip_yaw", "right_hip_roll", "right_hip_yaw")                                   # train.py:585-590
ANKLE_KNEE = ("left_knee_pitch", "left_ankle_pitch", "left_ankle_roll", "right_knee_pitch", "right_ankle_pitch", "right_ankle_roll")  # :606-613
ARM_POSE = ("left_shoulder_pitch", "left_shoulder_roll", "left_elbow_roll", "left_gripper_roll", "right_shoulder_pitch",
            "right_shoulder_roll", "right_elbow_roll", "right_gripper_roll")                                          # :632-641
SCALES = {  # train.py:1548-1585
    "stay_alive": 1.0, "upright": 1.0, "naive_forward": 5.0, "naive_forward_orientation": 0.3, "linvel_y_penalty": -2.0,
    "single_foot_contact": 0.3, "feet_airtime": 2.5, "feet_orientation": 0.3, "feet_too_close": -0.5, "straight_leg": -0.5,
    "ankle_knee": -0.05, "arm_pose": -2.0,
}


def _rotate_inverse(v: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """xax.rotate_vector_by_quat(v, q, inverse=True) [R]: q normalised as q / (|q| + 1e-6)."""
    q = q / (q.norm(dim=-1, keepdim=True) + 1e-6)
    w, u = q[..., 0:1], -q[..., 1:]
    t = 2 * torch.cross(u, v.expand_as(u), dim=-1)
    return v + w * t + torch.cross(u, t, dim=-1)


def _joint_deviation(qpos: torch.Tensor, names) -> torch.Tensor:
    """ksim.JointDeviationPenalty [R] with the targets / weights of train.py:559-562: sum_j w_j (q_j - target_j)^2."""
    idx = [7 + M.JOINT_NAMES.index(n) for n in names]
    tgt = torch.tensor([float(M.BIAS_POSE[M.JOINT_NAMES.index(n)]) for n in names], dtype=qpos.dtype, device=qpos.device)
    w = torch.tensor([_JOINT_WEIGHT[n] for n in names], dtype=qpos.dtype, device=qpos.device)
    return (w * (qpos[..., idx] - tgt) ** 2).sum(-1)


def _roll_pitch_abs(q: torch.Tensor) -> torch.Tensor:
    """|roll| + |pitch| of xax.quat_to_euler(q) [R]."""
    q = q / (q.norm(dim=-1, keepdim=True) + 1e-6)
    w, x, y, z = q.unbind(-1)
    roll = torch.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    sp = 2 * (w * y - z * x)
    pitch = torch.where(sp.abs() >= 1, torch.sign(sp) * (torch.pi / 2), torch.asin(sp.clamp(-1, 1)))
    return roll.abs() + pitch.abs()


def reward_terms(traj, feet_airtime: torch.Tensor, left_body: int, right_body: int) -> dict[str, torch.Tensor]:
    """{name: [T, N]} raw reward terms of the reference's list.  ``traj``: rollout.Trajectory; ``feet_airtime``: FeetAirtimeReward
    over the chunk ([T, N], from ZbotCudaEngine.feet_airtime_reward); left / right_body: bodies of the foot sites.

    Index alignment as in ksim's rollout [R]: terms on ``trajectory.qpos / qvel / xquat`` see the state AFTER step t (the terminal
    state where the episode ends), terms on ``trajectory.obs[...]`` see the observation the action of step t was computed from, i.e.
    slot t of the obs-like arrays (written by the launch of step t-1 / the reset launch)."""
    T = traj.qpos.shape[0]
    qpos, qvel, rt = traj.qpos, traj.qvel, traj.reward_terms[:T]
    quat = qpos[..., 3:7]
    f = qpos.dtype
    ez = torch.tensor([0.0, 0.0, 1.0], dtype=f, device=qpos.device)
    ex = torch.tensor([1.0, 0.0, 0.0], dtype=f, device=qpos.device)
    fwd = _rotate_inverse(ex, quat)
    done = traj.done != 0
    return {
        "stay_alive": torch.where(done, torch.full_like(qpos[..., 0], -1.0), torch.full_like(qpos[..., 0], 0.1)),  # ksim.StayAliveReward [R] (balance 10)
        "upright": _rotate_inverse(ez, quat)[..., 2],                                    # ksim.UprightReward [R]
        "naive_forward": qvel[..., 0].clamp(max=0.2),                                    # ksim.NaiveForwardReward(clip_max=0.2) [R]
        "naive_forward_orientation": fwd[..., 0] - fwd[..., 1:].norm(dim=-1),            # ksim.NaiveForwardOrientationReward [R]
        "linvel_y_penalty": _rotate_inverse(qvel[..., :3], quat)[..., 1].abs(),          # ksim.LinearVelocityPenalty(y, robot frame, l1) [R]
        "single_foot_contact": rt[..., 0],                                               # [V] train.py:691-713 (kernel epilogue)
        "feet_airtime": feet_airtime,                                                    # [V] train.py:485-546 (zb_feet_airtime_f32)
        "feet_orientation": torch.exp(-(_roll_pitch_abs(traj.xquat[:, :, left_body]) + _roll_pitch_abs(traj.xquat[:, :, right_body])) / 0.25),  # [V] train.py:647-689 on trajectory.xquat
        "feet_too_close": rt[..., 2],                                                    # [V] train.py:729-748 (kernel epilogue)
        "straight_leg": _joint_deviation(qpos, STRAIGHT_LEG),                            # [V] train.py:575-593 on [R] JointDeviationPenalty
        "ankle_knee": _joint_deviation(qpos, ANKLE_KNEE),                                # [V] train.py:596-617
        "arm_pose": _joint_deviation(qpos, ARM_POSE),                                    # [V] train.py:620-644
    }


def total(terms: dict[str, torch.Tensor]) -> torch.Tensor:
    """sum_i scale_i * term_i [T, N] (train.py:1548-1585)."""
    return sum(SCALES[k] * v for k, v in terms.items())
