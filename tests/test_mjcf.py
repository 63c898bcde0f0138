"""MJCF -> ZbModel compiler: round trip of zbot_syn and the MuJoCo XML semantics it restates."""

import ctypes
import math

import numpy as np
import pytest

import ksim_gym_zbot_b200  # noqa: F401
from ksim_gym_zbot_b200 import mjcf
from ksim_gym_zbot_b200 import model as M


def _fields_equal(a, b, skip=()):
    bad = []
    for name, _ in M.ZbModel._fields_:
        if name in skip:
            continue
        x, y = getattr(a, name), getattr(b, name)
        x = np.asarray(x) if isinstance(x, (int, float)) else np.ctypeslib.as_array(x)
        y = np.asarray(y) if isinstance(y, (int, float)) else np.ctypeslib.as_array(y)
        if not np.array_equal(x, y):
            bad.append((name, float(np.max(np.abs(np.asarray(x, dtype=np.float64) - np.asarray(y, dtype=np.float64))))))
    return bad


def test_round_trip_is_exact(zmodel):
    """to_mjcf -> compile_mjcf reproduces every field of zbot_syn bit for bit (values are printed at fp32 round-trip precision;
    the derived constants come from the same mj_setConst restatement)."""
    xml = mjcf.to_mjcf(zmodel, opts=M.make_opts(dt=0.001))
    c = mjcf.compile_mjcf(xml, servo_params=mjcf.servo_params_of(zmodel))
    assert _fields_equal(c.model, zmodel) == []
    assert bytes(c.model) == bytes(zmodel) or ctypes.sizeof(c.model) == ctypes.sizeof(zmodel)
    assert c.body_names == M.BODY_NAMES and c.joint_names == M.JOINT_NAMES
    assert c.opts["dt"] == pytest.approx(0.001) and c.opts["iterations"] == 8 and c.opts["solver"] == M.SOLVER_CG
    assert c.sensor_names[:3] == ["imu_site_quat", "imu_acc", "imu_gyro"] and c.ignored_geoms == []


def _mini(body_extra="", joint_attr="", compiler='<compiler angle="radian"/>', default="", foot_geom_attr=""):
    """A valid 26-body file assembled from zbot_syn with hooks to vary single elements."""
    xml = mjcf.to_mjcf(M.build_zbot_syn())
    xml = xml.replace('<compiler angle="radian" autolimits="true"/>', compiler + default)
    if joint_attr:
        xml = xml.replace('<joint name="right_hip_yaw" ', f'<joint name="right_hip_yaw" {joint_attr} ')
    if body_extra:
        xml = xml.replace('<freejoint name="floating_base"/>', '<freejoint name="floating_base"/>' + body_extra)
    if foot_geom_attr:
        xml = xml.replace('<geom name="Right_Foot_box" ', f'<geom name="Right_Foot_box" {foot_geom_attr} ')
    return xml


def test_degrees_euler_and_defaults():
    base = M.build_zbot_syn()
    # degree files: ranges and refs are converted, euler orientation = intrinsic xyz
    xml = mjcf.to_mjcf(base).replace('angle="radian"', 'angle="degree"')
    c = mjcf.compile_mjcf(xml, servo_params=mjcf.servo_params_of(base))
    np.testing.assert_allclose(c.model.field("jnt_range")[6:], base.field("jnt_range")[6:] * math.pi / 180, rtol=1e-6)
    xml = mjcf.to_mjcf(base).replace('<body name="head" pos="0.009999999776482582 0.0 0.12999999523162842" quat="1.0 0.0 0.0 0.0">',
                                     '<body name="head" pos="0.01 0 0.13" euler="0.3 -0.2 0.1">')
    assert "euler=" in xml
    c = mjcf.compile_mjcf(xml, servo_params=mjcf.servo_params_of(base))
    q = M.quat_mul(M.quat_mul(M.axis_angle_quat(np.array([1.0, 0, 0]), 0.3), M.axis_angle_quat(np.array([0, 1.0, 0]), -0.2)),
                   M.axis_angle_quat(np.array([0, 0, 1.0]), 0.1))
    np.testing.assert_allclose(c.model.field("body_quat")[M.BODY_NAMES.index("head")], q, atol=1e-7)
    # default classes: nested class overrides the parent class, the element overrides both
    dflt = ('<default><joint damping="0.5" armature="0.01"/><default class="leg"><joint damping="0.7"/></default></default>')
    xml = mjcf.to_mjcf(base).replace('<compiler angle="radian" autolimits="true"/>', '<compiler angle="radian" autolimits="true"/>' + dflt)
    import re

    xml = re.sub(r' armature="[^"]*" damping="[^"]*"', "", xml)  # let the defaults decide
    xml = xml.replace('<joint name="left_knee_pitch" ', '<joint name="left_knee_pitch" class="leg" ')
    xml = xml.replace('<joint name="left_hip_yaw" ', '<joint name="left_hip_yaw" class="leg" damping="0.9" ')
    c = mjcf.compile_mjcf(xml, servo_params=mjcf.servo_params_of(base))
    d = {n: 6 + j for j, n in enumerate(M.JOINT_NAMES)}
    assert c.model.dof_damping[d["right_hip_yaw"]] == pytest.approx(0.5) and c.model.dof_armature[d["right_hip_yaw"]] == pytest.approx(0.01)
    assert c.model.dof_damping[d["left_knee_pitch"]] == pytest.approx(0.7) and c.model.dof_armature[d["left_knee_pitch"]] == pytest.approx(0.01)
    assert c.model.dof_damping[d["left_hip_yaw"]] == pytest.approx(0.9)
    # the compiled constants follow the changed model (armature enters M)
    assert c.model.dof_invweight0[d["right_hip_yaw"]] != base.dof_invweight0[d["right_hip_yaw"]]


def test_fullinertia_is_diagonalised():
    base = M.build_zbot_syn()
    b = M.BODY_NAMES.index("head")
    R = M.quat_to_mat(M.axis_angle_quat(np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8]), 0.7))
    I = R @ np.diag([3e-4, 2e-4, 1e-4]) @ R.T
    full = 'fullinertia="%r %r %r %r %r %r"' % tuple(float(x) for x in (I[0, 0], I[1, 1], I[2, 2], I[0, 1], I[0, 2], I[1, 2]))
    xml = mjcf.to_mjcf(base)
    import re

    xml = re.sub(r'(<body name="head"[^>]*>\s*<inertial pos="[^"]*") quat="[^"]*" (mass="[^"]*") diaginertia="[^"]*"', r"\1 \2 " + full, xml)
    assert "fullinertia" in xml
    c = mjcf.compile_mjcf(xml, servo_params=mjcf.servo_params_of(base))
    np.testing.assert_allclose(c.model.field("body_inertia")[b], [3e-4, 2e-4, 1e-4], rtol=1e-5)
    Rq = M.quat_to_mat(c.model.field("body_iquat")[b].astype(np.float64))
    np.testing.assert_allclose(Rq @ np.diag(c.model.field("body_inertia")[b].astype(np.float64)) @ Rq.T, I, atol=1e-9)
    assert np.linalg.det(Rq) == pytest.approx(1.0, abs=1e-6)


def test_inertia_from_box_geom_and_autolimits():
    base = M.build_zbot_syn()
    xml = mjcf.to_mjcf(base)
    import re

    # head: drop the <inertial>, give it a box geom -> mass = density * volume, box inertia about its centre
    xml = re.sub(r'(<body name="head"[^>]*>)\s*<inertial [^>]*/>', r'\1<geom type="box" size="0.03 0.035 0.03" pos="0 0 0.01" density="800" contype="0" conaffinity="0"/>', xml)
    # autolimits: a joint with a range and no `limited` attribute is limited; without a range it is not
    xml = xml.replace(' limited="true"', "")
    xml = re.sub(r'(<joint name="left_gripper_roll"[^>]*?) range="[^"]*"', r"\1", xml)
    c = mjcf.compile_mjcf(xml, servo_params=mjcf.servo_params_of(base))
    b = M.BODY_NAMES.index("head")
    mass = 800 * 0.06 * 0.07 * 0.06
    assert c.model.body_mass[b] == pytest.approx(mass, rel=1e-6)
    np.testing.assert_allclose(c.model.field("body_inertia")[b], mass / 12 * np.array([0.07**2 + 0.06**2, 0.06**2 + 0.06**2, 0.06**2 + 0.07**2]), rtol=1e-5)
    np.testing.assert_allclose(c.model.field("body_ipos")[b], [0, 0, 0.01], atol=1e-9)
    lim = np.ctypeslib.as_array(c.model.jnt_limited)
    d = {n: 6 + j for j, n in enumerate(M.JOINT_NAMES)}
    assert lim[d["left_gripper_roll"]] == 0 and lim[6:].sum() == M.NU - 1


@pytest.mark.parametrize("mutate,needle", [
    (lambda x: x.replace('<freejoint name="floating_base"/>', '<joint type="slide"/>'), "not supported"),
    (lambda x: x.replace('name="left_hip_yaw"', 'name="left_hip_yawX"'), "action order"),
    (lambda x: x.replace('<geom name="floor" type="plane"', '<geom name="floor" type="sphere"'), "no floor"),
    (lambda x: x.replace('<site name="imu_site"', '<site name="imu"'), "imu_site"),
    (lambda x: x.replace('<motor name="left_elbow_roll_ctrl" joint="left_elbow_roll"', '<motor name="m" joint="nope"'), "no motor"),
    (lambda x: x.replace("</worldbody>", '<body name="extra"><inertial mass="1" diaginertia="1 1 1"/></body></worldbody>'), "26 bodies"),
])
def test_errors_name_the_problem(mutate, needle):
    base = M.build_zbot_syn()
    xml = mutate(mjcf.to_mjcf(base))
    with pytest.raises(mjcf.MjcfError) as e:
        mjcf.compile_mjcf(xml, servo_params=mjcf.servo_params_of(base))
    assert needle in str(e.value)


def test_compiled_model_steps_like_the_original(zmodel):
    """The oracle cannot tell the compiled round-trip model from the original (same bytes -> same trajectory)."""
    import zb_oracle
    from helpers import random_actions, random_states

    c = mjcf.compile_mjcf(mjcf.to_mjcf(zmodel), servo_params=mjcf.servo_params_of(zmodel))
    opts = M.make_opts(dt=0.005)
    q, v = random_states(zmodel, 8, 3)
    res = []
    for mdl in (zmodel, c.model):
        o = zb_oracle.Oracle(mdl, opts, np.float32)
        st = o.new_state(8)
        o.reset(st, q, v, np.full(8, 1e9, np.float32))
        o.step(st, random_actions(8, 4), 4)
        res.append(st["qpos"].copy())
    np.testing.assert_array_equal(res[0], res[1])
