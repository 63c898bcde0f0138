"""`zbot_syn` v1 — frozen synthetic Z-Bot-topology model + the host-side "model compile".

The real Z-Bot MJCF is not in the reference (it is downloaded at run time:
/root/reference/train.py:1327 ``ksim.get_mujoco_model_path("zbot", name="robot")``), so masses,
inertias, link offsets and gains below are a documented stand-in.  What IS taken from the
reference: the joint/action order (train.py:59-82), nq=27/nv=26/nu=20 (train.py:27, 461, 1252-1253),
25 non-world bodies (train.py:42-43), the sites and sensors (train.py:1453, 1495-1520), motor
actuators fed by the Feetech servo model (train.py:1121-1134, 1421-1431), the floor plane with
geom_priority 2 (train.py:1329-1330), the nominal pose JOINT_BIASES (train.py:61-82), vmax/amax
defaults (train.py:1347-1348) and the servo deadband (train.py:1111-1118).

This module also does what MuJoCo's ``mj_setConst`` does at compile time (dof_invweight0,
body_invweight0, meaninertia) with a small fp64 numpy kinematics — host logic, not the hot path.
"""

from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import numpy as np

NQ, NV, NU, NBODY, NSITE, NBOX, MAXCON = 27, 26, 20, 26, 3, 2, 8
NEFC_CON = 4 * MAXCON
NEFC = 2 * NV + NEFC_CON
NSENSORDATA = 18

# (joint_name, reference_angle_rad) — order of train.py:59-82 (== qpos[7:] == action order).
JOINT_BIASES: list[tuple[str, float]] = [
    ("right_hip_yaw", 0.0),
    ("right_hip_roll", -0.1),
    ("right_hip_pitch", -0.4),
    ("right_knee_pitch", -0.8),
    ("right_ankle_pitch", -0.4),
    ("right_ankle_roll", -0.1),
    ("left_hip_yaw", 0.0),
    ("left_hip_roll", 0.1),
    ("left_hip_pitch", -0.4),
    ("left_knee_pitch", -0.8),
    ("left_ankle_pitch", -0.4),
    ("left_ankle_roll", 0.1),
    ("left_shoulder_pitch", 0.0),
    ("left_shoulder_roll", 0.2),
    ("left_elbow_roll", -0.2),
    ("left_gripper_roll", 0.0),
    ("right_shoulder_pitch", 0.0),
    ("right_shoulder_roll", -0.2),
    ("right_elbow_roll", 0.2),
    ("right_gripper_roll", 0.0),
]
JOINT_NAMES = [n for n, _ in JOINT_BIASES]
BIAS_POSE = np.array([v for _, v in JOINT_BIASES], dtype=np.float64)

_f, _i = ctypes.c_float, ctypes.c_int32


class ZbOpts(ctypes.Structure):
    """Mirror of ``ZbOpts`` in include/zbphys.h."""

    _fields_ = [
        ("dt", _f),
        ("tolerance", _f),
        ("ls_tolerance", _f),
        ("impratio", _f),
        ("gravity", _f * 3),
        ("iterations", _i),
        ("ls_iterations", _i),
        ("solver", _i),
        ("integrator", _i),
        ("eulerdamp", _i),
        ("warmstart", _i),
        ("substep_order", _i),
        ("reserved", _i),
    ]


SOLVER_CG, SOLVER_NEWTON = 0, 1
GEOM_BOX, GEOM_SPHERE, GEOM_CAPSULE = 0, 1, 2  # ZbModel.box_type / pair_type
INT_EULER, INT_IMPLICITFAST = 0, 1


def make_opts(
    dt: float = 0.001,
    iterations: int = 8,
    ls_iterations: int = 8,
    solver: int = SOLVER_CG,
    integrator: int = INT_IMPLICITFAST,
    tolerance: float = 1e-8,
    ls_tolerance: float = 0.01,
    impratio: float = 1.0,
    eulerdamp: int = 1,
    warmstart: int = 1,
    substep_order: int = 0,
    gravity: tuple[float, float, float] = (0.0, 0.0, -9.81),
) -> ZbOpts:
    """Engine options; defaults = train.py:1777-1780 (dt, iterations, ls_iterations) + MuJoCo defaults."""
    o = ZbOpts()
    o.dt, o.tolerance, o.ls_tolerance, o.impratio = dt, tolerance, ls_tolerance, impratio
    o.gravity[:] = gravity
    o.iterations, o.ls_iterations, o.solver, o.integrator = iterations, ls_iterations, solver, integrator
    o.eulerdamp, o.warmstart, o.substep_order = eulerdamp, warmstart, substep_order
    return o


class ZbModel(ctypes.Structure):
    """Mirror of ``ZbModel`` in include/zbphys.h (field order must match exactly)."""

    _fields_ = [
        ("body_parent", _i * NBODY),
        ("body_dofadr", _i * NBODY),
        ("body_jnttype", _i * NBODY),
        ("dof_body", _i * NV),
        ("dof_parent", _i * NV),
        ("body_pos", _f * 3 * NBODY),
        ("body_quat", _f * 4 * NBODY),
        ("body_ipos", _f * 3 * NBODY),
        ("body_iquat", _f * 4 * NBODY),
        ("body_mass", _f * NBODY),
        ("body_inertia", _f * 3 * NBODY),
        ("body_invweight0", _f * 2 * NBODY),
        ("jnt_pos", _f * 3 * NBODY),
        ("jnt_axis", _f * 3 * NBODY),
        ("jnt_range", _f * 2 * NV),
        ("jnt_margin", _f * NV),
        ("jnt_solref", _f * 2 * NV),
        ("jnt_solimp", _f * 5 * NV),
        ("jnt_limited", _i * NV),
        ("dof_armature", _f * NV),
        ("dof_damping", _f * NV),
        ("dof_frictionloss", _f * NV),
        ("dof_invweight0", _f * NV),
        ("dof_solref", _f * 2 * NV),
        ("dof_solimp", _f * 5 * NV),
        ("qpos0", _f * NQ),
        ("act_dof", _i * NU),
        ("act_gear", _f * NU),
        ("act_ctrlrange", _f * 2 * NU),
        ("act_forcerange", _f * 2 * NU),
        ("act_ctrllimited", _i * NU),
        ("act_forcelimited", _i * NU),
        ("servo_kp", _f * NU),
        ("servo_kd", _f * NU),
        ("servo_max_pwm", _f * NU),
        ("servo_vin", _f * NU),
        ("servo_kt", _f * NU),
        ("servo_R", _f * NU),
        ("servo_vmax", _f * NU),
        ("servo_amax", _f * NU),
        ("servo_error_gain", _f * NU),
        ("servo_deadband_pos", _f),
        ("servo_deadband_neg", _f),
        ("floor_pos", _f * 3),
        ("floor_normal", _f * 3),
        ("floor_friction", _f * 3),
        ("floor_solref", _f * 2),
        ("floor_solimp", _f * 5),
        ("floor_priority", _i),
        ("floor_margin", _f),
        ("floor_gap", _f),
        ("box_body", _i * NBOX),
        ("box_pos", _f * 3 * NBOX),
        ("box_quat", _f * 4 * NBOX),
        ("box_size", _f * 3 * NBOX),
        ("box_friction", _f * 3 * NBOX),
        ("box_solref", _f * 2 * NBOX),
        ("box_solimp", _f * 5 * NBOX),
        ("box_priority", _i * NBOX),
        ("box_margin", _f * NBOX),
        ("box_gap", _f * NBOX),
        ("box_type", _i * NBOX),
        ("pair_body", _i * NBOX),
        ("pair_type", _i * NBOX),
        ("pair_pos", _f * 3 * NBOX),
        ("pair_quat", _f * 4 * NBOX),
        ("pair_size", _f * 3 * NBOX),
        ("pair_friction", _f * 3 * NBOX),
        ("pair_solref", _f * 2 * NBOX),
        ("pair_solimp", _f * 5 * NBOX),
        ("pair_priority", _i * NBOX),
        ("pair_margin", _f * NBOX),
        ("pair_gap", _f * NBOX),
        ("site_body", _i * NSITE),
        ("site_pos", _f * 3 * NSITE),
        ("site_quat", _f * 4 * NSITE),
        ("site_size", _f * 3 * NSITE),
        ("meaninertia", _f),
        ("reserved", _i),
    ]

    def field(self, name: str) -> np.ndarray:
        """numpy view (shares memory) of one array field."""
        return np.ctypeslib.as_array(getattr(self, name))

    def copy(self) -> "ZbModel":
        out = ZbModel()
        ctypes.memmove(ctypes.byref(out), ctypes.byref(self), ctypes.sizeof(ZbModel))
        return out


# ---------------------------------------------------------------------------------------------
# small fp64 helpers (host logic only)
# ---------------------------------------------------------------------------------------------


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    w1, x1, y1, z1 = a
    w2, x2, y2, z2 = b
    return np.array(
        [
            w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
            w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
            w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
            w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
        ]
    )


def quat_to_mat(q: np.ndarray) -> np.ndarray:
    w, x, y, z = q
    return np.array(
        [
            [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
        ]
    )


def axis_angle_quat(axis: np.ndarray, angle: float) -> np.ndarray:
    s, c = math.sin(angle * 0.5), math.cos(angle * 0.5)
    return np.array([c, axis[0] * s, axis[1] * s, axis[2] * s])


def _box_inertia(mass: float, dims: tuple[float, float, float]) -> np.ndarray:
    x, y, z = dims
    return mass / 12.0 * np.array([y * y + z * z, x * x + z * z, x * x + y * y])


@dataclass
class _Body:
    name: str
    parent: int
    pos: tuple[float, float, float]
    quat: np.ndarray
    mass: float
    ipos: tuple[float, float, float]
    dims: tuple[float, float, float]  # box used for the inertia
    iquat: np.ndarray
    joint: str | None = None  # "free" | hinge joint name | None
    axis: tuple[float, float, float] = (0.0, 0.0, 1.0)
    jpos: tuple[float, float, float] = (0.0, 0.0, 0.0)


_QI = np.array([1.0, 0.0, 0.0, 0.0])


def _bodies() -> list[_Body]:
    X, Y, Z = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)
    NX, NY = (-1.0, 0.0, 0.0), (0.0, -1.0, 0.0)
    b: list[_Body] = [
        _Body("world", 0, (0, 0, 0), _QI, 0.0, (0, 0, 0), (0.0, 0.0, 0.0), _QI),
        _Body("base", 0, (0, 0, 0.30), _QI, 1.2, (0.004, 0.0, 0.03), (0.09, 0.14, 0.16),
              axis_angle_quat(np.array(Z), 0.05), joint="free"),
    ]

    def leg(side: str, sy: float) -> None:
        root = 1
        # axes chosen so that JOINT_BIASES gives flat feet: pitch sum -0.4 + 0.8 - 0.4, roll -0.1 + 0.1
        chain = [
            (f"{side}_hip_yaw", (0.0, sy * 0.05, -0.04), 0.08, (0.0, 0.0, -0.01), (0.03, 0.03, 0.03), Z, (0, 0, 0)),
            (f"{side}_hip_roll", (0.0, 0.0, -0.03), 0.25, (0.0, 0.0, -0.005), (0.045, 0.035, 0.04), X, (0, 0, 0)),
            (f"{side}_hip_pitch", (0.0, 0.0, -0.02), 0.30, (0.002, 0.0, -0.04), (0.04, 0.04, 0.09), Y, (0, 0, 0)),
            (f"{side}_knee_pitch", (0.0, 0.0, -0.08), 0.30, (-0.002, 0.0, -0.04), (0.035, 0.04, 0.09), NY, (0, 0, 0)),
            (f"{side}_ankle_pitch", (0.0, 0.0, -0.08), 0.08, (0.0, 0.0, -0.008), (0.03, 0.03, 0.03), Y, (0, 0, 0.002)),
            (f"{side}_ankle_roll", (0.0, 0.0, -0.02), 0.25, (0.005, 0.0, -0.01), (0.045, 0.035, 0.035), NX, (0, 0, 0)),
        ]
        parent = root
        for k, (name, pos, mass, ipos, dims, axis, jpos) in enumerate(chain):
            iq = axis_angle_quat(np.array(Y), 0.2) if k in (2, 3) else _QI
            b.append(_Body(name, parent, pos, _QI, mass, ipos, dims, iq, joint=name, axis=axis, jpos=jpos))
            parent = len(b) - 1
        foot = "Right_Foot" if side == "right" else "Left_Foot"
        b.append(_Body(foot, parent, (0.01, 0.0, -0.025), _QI, 0.05, (0.0, 0.0, 0.0), (0.10, 0.05, 0.01), _QI))

    def arm(side: str, sy: float) -> None:
        chain = [
            (f"{side}_shoulder_pitch", (0.0, sy * 0.08, 0.08), axis_angle_quat(np.array(X), -sy * 0.15),
             0.15, (0.0, sy * 0.01, 0.0), (0.035, 0.03, 0.035), Y, (0, 0, 0)),
            (f"{side}_shoulder_roll", (0.0, sy * 0.02, 0.0), _QI, 0.15, (0.0, 0.0, -0.03), (0.03, 0.03, 0.07), X, (0, 0, 0)),
            (f"{side}_elbow_roll", (0.0, 0.0, -0.07), _QI, 0.12, (0.0, 0.0, -0.03), (0.028, 0.028, 0.07), X, (0, 0, 0.003)),
            (f"{side}_gripper_roll", (0.0, 0.0, -0.07), _QI, 0.08, (0.0, 0.0, -0.02), (0.03, 0.04, 0.05), Z, (0, 0, 0)),
        ]
        parent = 1
        for k, (name, pos, quat, mass, ipos, dims, axis, jpos) in enumerate(chain):
            iq = axis_angle_quat(np.array(X), 0.1 * sy) if k == 1 else _QI
            b.append(_Body(name, parent, pos, quat, mass, ipos, dims, iq, joint=name, axis=axis, jpos=jpos))
            parent = len(b) - 1

    leg("right", -1.0)
    leg("left", +1.0)
    arm("left", +1.0)
    arm("right", -1.0)
    b.append(_Body("imu_link", 1, (0.0, 0.0, 0.05), _QI, 0.01, (0, 0, 0), (0.02, 0.02, 0.005), _QI))
    b.append(_Body("head", 1, (0.01, 0.0, 0.13), _QI, 0.15, (0.0, 0.0, 0.01), (0.06, 0.07, 0.06), _QI))
    assert len(b) == NBODY
    return b


BODY_NAMES = [x.name for x in _bodies()]


def _fk_np(m: ZbModel, qpos: np.ndarray):
    """fp64 forward kinematics (host-side, for the compile-time constants only)."""
    parent, jtype = m.field("body_parent"), m.field("body_jnttype")
    dofadr = m.field("body_dofadr")
    xpos, xquat = np.zeros((NBODY, 3)), np.zeros((NBODY, 4))
    xanchor, xaxis = np.zeros((NBODY, 3)), np.zeros((NBODY, 3))
    xquat[0] = _QI
    qpos0 = m.field("qpos0").astype(np.float64)
    for b in range(1, NBODY):
        p = parent[b]
        if jtype[b] == 1:
            xpos[b] = qpos[0:3]
            xquat[b] = qpos[3:7] / np.linalg.norm(qpos[3:7])
            continue
        pos = xpos[p] + quat_to_mat(xquat[p]) @ m.field("body_pos")[b].astype(np.float64)
        quat = quat_mul(xquat[p], m.field("body_quat")[b].astype(np.float64))
        if jtype[b] == 2:
            qa = dofadr[b] + 1
            jp_, ax = m.field("jnt_pos")[b].astype(np.float64), m.field("jnt_axis")[b].astype(np.float64)
            R = quat_to_mat(quat)
            xanchor[b], xaxis[b] = pos + R @ jp_, R @ ax
            quat = quat_mul(quat, axis_angle_quat(ax, qpos[qa] - qpos0[qa]))
            pos = xanchor[b] - quat_to_mat(quat) @ jp_
        xpos[b], xquat[b] = pos, quat
    return xpos, xquat, xanchor, xaxis


def _mass_matrix_np(m: ZbModel, qpos: np.ndarray):
    """fp64 joint-space inertia from body-COM Jacobians (independent of the CRB path used at run time)."""
    xpos, xquat, xanchor, xaxis = _fk_np(m, qpos)
    parent, jtype, dofadr = m.field("body_parent"), m.field("body_jnttype"), m.field("body_dofadr")
    M = np.zeros((NV, NV))
    jacs = []
    for b in range(NBODY):
        R = quat_to_mat(xquat[b])
        com = xpos[b] + R @ m.field("body_ipos")[b].astype(np.float64)
        Ri = quat_to_mat(quat_mul(xquat[b], m.field("body_iquat")[b].astype(np.float64)))
        Iw = Ri @ np.diag(m.field("body_inertia")[b].astype(np.float64)) @ Ri.T
        jp_, jr = np.zeros((3, NV)), np.zeros((3, NV))
        a = b
        while a > 0:
            if jtype[a] == 1:
                Rb = quat_to_mat(xquat[a])
                jp_[:, 0:3] = np.eye(3)
                for k in range(3):
                    jr[:, 3 + k] = Rb[:, k]
                    jp_[:, 3 + k] = np.cross(Rb[:, k], com - xpos[a])
            elif jtype[a] == 2:
                d = dofadr[a]
                jr[:, d] = xaxis[a]
                jp_[:, d] = np.cross(xaxis[a], com - xanchor[a])
            a = parent[a]
        mass = float(m.field("body_mass")[b])
        M += mass * jp_.T @ jp_ + jr.T @ Iw @ jr
        jacs.append((jp_, jr))
    M += np.diag(m.field("dof_armature").astype(np.float64))
    return M, jacs


def build_zbot_syn() -> ZbModel:
    """Build the frozen synthetic model `zbot_syn` v1 (deterministic)."""
    m = ZbModel()
    bodies = _bodies()
    d = 0
    for i, b in enumerate(bodies):
        m.body_parent[i] = b.parent
        m.field("body_pos")[i] = b.pos
        m.field("body_quat")[i] = b.quat
        m.field("body_ipos")[i] = b.ipos
        m.field("body_iquat")[i] = b.iquat
        m.body_mass[i] = b.mass
        m.field("body_inertia")[i] = _box_inertia(b.mass, b.dims) if b.mass > 0 else 0.0
        m.field("jnt_pos")[i] = b.jpos
        m.field("jnt_axis")[i] = b.axis
        if b.joint == "free":
            m.body_jnttype[i], m.body_dofadr[i] = 1, d
            for k in range(6):
                m.dof_body[d + k] = i
                m.dof_parent[d + k] = d + k - 1
            d += 6
        elif b.joint is not None:
            assert JOINT_NAMES[d - 6] == b.joint, (JOINT_NAMES[d - 6], b.joint)
            m.body_jnttype[i], m.body_dofadr[i] = 2, d
            m.dof_body[d] = i
            # parent dof: last dof of the nearest ancestor with a joint
            a = b.parent
            while a > 0 and m.body_dofadr[a] < 0:
                a = m.body_parent[a]
            m.dof_parent[d] = (5 if m.body_jnttype[a] == 1 else m.body_dofadr[a]) if a > 0 else -1
            d += 1
        else:
            m.body_jnttype[i], m.body_dofadr[i] = 0, -1
    assert d == NV

    # joint limits (all 20 hinges limited, rad)
    rng_tbl = {
        "hip_yaw": (-0.8, 0.8), "hip_roll": (-0.6, 0.6), "hip_pitch": (-1.6, 1.0), "knee_pitch": (-2.0, 0.1),
        "ankle_pitch": (-1.0, 1.0), "ankle_roll": (-0.6, 0.6), "shoulder_pitch": (-2.0, 2.0),
        "shoulder_roll": (-0.4, 1.6), "elbow_roll": (-1.6, 1.6), "gripper_roll": (-1.0, 1.0),
    }
    solref, solimp = (0.02, 1.0), (0.9, 0.95, 0.001, 0.5, 2.0)
    for k in range(NV):
        m.field("jnt_solref")[k] = solref
        m.field("jnt_solimp")[k] = solimp
        m.field("dof_solref")[k] = solref
        m.field("dof_solimp")[k] = solimp
    for j, name in enumerate(JOINT_NAMES):
        dof = 6 + j
        side, kind = name.split("_", 1)
        lo, hi = rng_tbl[kind]
        if kind == "shoulder_roll" and side == "right":
            lo, hi = -hi, -lo
        m.field("jnt_range")[dof] = (lo, hi)
        m.jnt_limited[dof] = 1
        is_leg = j < 12
        m.dof_armature[dof] = 0.02 if is_leg else 0.008
        m.dof_damping[dof] = 0.15 if is_leg else 0.06
        m.dof_frictionloss[dof] = 0.06 if is_leg else 0.03
        # motors named <joint>_ctrl, ctrl == torque from FeetechActuators (train.py:1264-1280)
        m.act_dof[j] = dof
        m.act_gear[j] = 1.0
        tmax = 6.0 if is_leg else 2.2
        m.field("act_ctrlrange")[j] = (-tmax, tmax)
        m.field("act_forcerange")[j] = (-tmax, tmax)
        m.act_ctrllimited[j] = 1
        m.act_forcelimited[j] = 1
        # Feetech-like constants (field names train.py:1121-1134)
        m.servo_kp[j] = 20.0 if is_leg else 12.0
        m.servo_kd[j] = 0.08 if is_leg else 0.05
        m.servo_max_pwm[j] = 1.0
        m.servo_vin[j] = 12.0
        m.servo_kt[j] = 0.8 if is_leg else 0.4
        m.servo_R[j] = 1.5 if is_leg else 2.0
        m.servo_vmax[j] = 5.0  # train.py:1347
        m.servo_amax[j] = 17.45  # train.py:1348
        m.servo_error_gain[j] = 0.16
    enc = 0.087 * math.pi / 180.0  # train.py:1113
    m.servo_deadband_pos = 2 * enc
    m.servo_deadband_neg = 2 * enc

    m.field("qpos0")[:] = 0.0
    m.field("qpos0")[0:3] = bodies[1].pos
    m.field("qpos0")[3:7] = (1, 0, 0, 0)

    # floor: plane geom "floor" of the mujoco_scenes "smooth" scene, priority 2 (train.py:1328-1330)
    m.field("floor_pos")[:] = 0.0
    m.field("floor_normal")[:] = (0, 0, 1)
    m.field("floor_friction")[:] = (1.0, 0.005, 0.0001)
    m.field("floor_solref")[:] = solref
    m.field("floor_solimp")[:] = solimp
    m.floor_priority, m.floor_margin, m.floor_gap = 2, 0.0, 0.0
    for k, foot in enumerate(("Right_Foot", "Left_Foot")):  # geom-id order == body order
        bi = BODY_NAMES.index(foot)
        m.box_body[k] = bi
        m.field("box_pos")[k] = (0.0, 0.0, 0.0)
        m.field("box_quat")[k] = _QI
        m.field("box_size")[k] = (0.05, 0.025, 0.005)
        m.field("box_friction")[k] = (0.8, 0.005, 0.0001)
        m.field("box_solref")[k] = solref
        m.field("box_solimp")[k] = solimp
        m.box_priority[k], m.box_margin[k], m.box_gap[k] = 0, 0.0, 0.0
        m.box_type[k], m.pair_body[k] = GEOM_BOX, 0  # geom 1 = the floor plane (the pair_* fields are unused)
        m.field("pair_quat")[k] = _QI
    # sites: imu_site on imu_link; left_foot/right_foot touch boxes around the foot plates
    m.site_body[0] = BODY_NAMES.index("imu_link")
    m.field("site_pos")[0] = (0.0, 0.0, 0.002)
    m.field("site_quat")[0] = _QI
    m.field("site_size")[0] = (0.005, 0.005, 0.005)
    for k, foot in ((1, "Left_Foot"), (2, "Right_Foot")):
        m.site_body[k] = BODY_NAMES.index(foot)
        m.field("site_pos")[k] = (0.0, 0.0, 0.0)
        m.field("site_quat")[k] = _QI
        m.field("site_size")[k] = (0.06, 0.035, 0.02)

    set_const(m)
    return m


def set_collision_group(m: ZbModel, g: int, *, body: int | str, geom: str, size, pos=(0.0, 0.0, 0.0), quat=None, friction=None,
                        other_body: int | str = 0, other_geom: str = "sphere", other_size=(0.0, 0.0, 0.0), other_pos=(0.0, 0.0, 0.0),
                        other_quat=None, other_friction=None, priority: int = 0, other_priority: int = 0, margin: float = 0.0) -> None:
    """Configure collision group ``g`` (ZbModel.box_* / pair_*): geom 2 (``geom`` on ``body``) against the floor plane
    (``other_body=0``) or against geom 1 (``other_geom`` on ``other_body``).  MJX pair order: sphere before capsule before box."""
    types = {"box": GEOM_BOX, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE}
    bi = BODY_NAMES.index(body) if isinstance(body, str) else int(body)
    oi = BODY_NAMES.index(other_body) if isinstance(other_body, str) else int(other_body)
    m.box_body[g], m.box_type[g] = bi, types[geom]
    sz = np.zeros(3); sz[:len(np.atleast_1d(size))] = np.atleast_1d(size)
    m.field("box_size")[g] = sz
    m.field("box_pos")[g] = pos
    m.field("box_quat")[g] = _QI if quat is None else quat
    if friction is not None:
        m.field("box_friction")[g] = friction
    m.box_priority[g], m.box_margin[g] = priority, margin
    m.pair_body[g], m.pair_type[g] = oi, types[other_geom]
    osz = np.zeros(3); osz[:len(np.atleast_1d(other_size))] = np.atleast_1d(other_size)
    m.field("pair_size")[g] = osz
    m.field("pair_pos")[g] = other_pos
    m.field("pair_quat")[g] = _QI if other_quat is None else other_quat
    if oi != 0:
        m.field("pair_friction")[g] = (1.0, 0.005, 0.0001) if other_friction is None else other_friction
        m.field("pair_solref")[g] = m.field("box_solref")[g]
        m.field("pair_solimp")[g] = m.field("box_solimp")[g]
    m.pair_priority[g], m.pair_margin[g] = other_priority, margin
    if oi != 0:
        order = {GEOM_SPHERE: 0, GEOM_CAPSULE: 1, GEOM_BOX: 2}
        if types[geom] == GEOM_BOX or order[types[other_geom]] > order[types[geom]]:
            raise ValueError("body-body pairs: sphere-sphere, sphere-capsule (sphere = geom 1) and capsule-capsule only")


def set_const(m: ZbModel) -> None:
    """MuJoCo's ``mj_setConst`` for the fields this path reads ([R]): ``dof_invweight0`` (diagonal of ``M^-1`` at ``qpos0``,
    averaged over the translational / rotational triples of the free joint), ``body_invweight0`` (mean translational /
    rotational diagonal of ``J M^-1 J^T`` at the body COM) and ``meaninertia`` (mean diagonal of ``M``)."""
    qpos0 = m.field("qpos0").astype(np.float64)
    M, jacs = _mass_matrix_np(m, qpos0)
    Minv = np.linalg.inv(M)
    diag = np.diag(Minv)
    inv0 = m.field("dof_invweight0")
    inv0[0:3] = diag[0:3].mean()
    inv0[3:6] = diag[3:6].mean()
    inv0[6:] = diag[6:]
    m.field("body_invweight0")[0] = 0.0
    for b in range(1, NBODY):
        jp_, jr = jacs[b]
        m.field("body_invweight0")[b] = (np.trace(jp_ @ Minv @ jp_.T) / 3.0, np.trace(jr @ Minv @ jr.T) / 3.0)
    m.meaninertia = float(np.diag(M).mean())


def total_mass(m: ZbModel) -> float:
    return float(m.field("body_mass").sum())


def foot_bottom_z(m: ZbModel, qpos: np.ndarray) -> float:
    """Lowest box-vertex height (fp64 host FK) — used to place the base so the feet touch the floor."""
    xpos, xquat, _, _ = _fk_np(m, qpos)
    lo = np.inf
    for k in range(NBOX):
        b = m.box_body[k]
        R = quat_to_mat(quat_mul(xquat[b], m.field("box_quat")[k].astype(np.float64)))
        c = xpos[b] + quat_to_mat(xquat[b]) @ m.field("box_pos")[k].astype(np.float64)
        s = m.field("box_size")[k].astype(np.float64)
        for sx in (-1, 1):
            for sy in (-1, 1):
                for sz in (-1, 1):
                    lo = min(lo, (c + R @ (s * np.array([sx, sy, sz])))[2])
    return float(lo)


def default_qpos(m: ZbModel, penetration: float = 5e-4) -> np.ndarray:
    """JOINT_BIASES pose (train.py:61-82, reset train.py:1473) with the base lowered onto the floor."""
    q = m.field("qpos0").astype(np.float64).copy()
    q[7:] = BIAS_POSE
    q[2] -= foot_bottom_z(m, q) + penetration
    return q


def randomize_env_params(m: ZbModel, n: int, seed: int) -> dict[str, np.ndarray]:
    """Host-side stand-in for ksim's randomizers with the ranges of train.py:1439-1455.

    ksim defaults for the un-parameterised ones are recalled (PARITY_RISKS.md: ZB_RANDOMIZER_RANGES).
    """
    r = np.random.default_rng(seed)
    f32 = np.float32
    out: dict[str, np.ndarray] = {}
    out["dof_frictionloss"] = (m.field("dof_frictionloss")[None, :] * r.uniform(0.5, 2.0, (n, NV))).astype(f32)
    out["dof_armature"] = (m.field("dof_armature")[None, :] * r.uniform(0.95, 1.05, (n, NV))).astype(f32)
    scale = r.uniform(0.95, 1.15, (n, 1))  # train.py:1443
    out["body_mass"] = (m.field("body_mass")[None, :] * scale).astype(f32)
    out["body_inertia"] = np.repeat(m.field("body_inertia")[None], n, 0).astype(f32)
    out["dof_damping"] = (m.field("dof_damping")[None, :] * r.uniform(0.95, 1.05, (n, NV))).astype(f32)
    q0 = np.repeat(m.field("qpos0")[None], n, 0).astype(np.float64)
    q0[:, 7:] += r.uniform(math.radians(-2), math.radians(2), (n, NU))  # train.py:1445
    out["qpos0"] = q0.astype(f32)
    out["floor_friction"] = r.uniform(0.3, 1.5, (n,)).astype(f32)  # train.py:1446-1448
    # IMU alignment: tilt sigma 5 deg, yaw sigma 1 deg, 5 mm (train.py:1452-1454)
    pos = m.field("site_pos")[0][None, :] + r.normal(0, 0.005, (n, 3))
    quat = np.zeros((n, 4))
    for i in range(n):
        tilt, yaw = r.normal(0, math.radians(5)), r.normal(0, math.radians(1.0))
        phi = r.uniform(0, 2 * math.pi)
        qt = axis_angle_quat(np.array([math.cos(phi), math.sin(phi), 0.0]), tilt)
        qy = axis_angle_quat(np.array([0.0, 0.0, 1.0]), yaw)
        quat[i] = quat_mul(quat_mul(qy, qt), m.field("site_quat")[0].astype(np.float64))
    out["imu_site_pos"] = pos.astype(f32)
    out["imu_site_quat"] = quat.astype(f32)
    return out
