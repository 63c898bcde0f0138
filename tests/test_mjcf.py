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


# ---------------------------------------------------------------------------------------------------------------------------
# a file the compiler's own writer did NOT produce: an onshape-to-robot / K-Scale style export (degrees, euler, default classes with
# childclass, visual meshes, <asset>, fullinertia, motors with ctrlrange, a <sensor> block), generated here from a small table with its own
# formatting, and checked field by field against numbers computed from that table
# ---------------------------------------------------------------------------------------------------------------------------
def _export_style_xml():
    rng = np.random.default_rng(12)
    legs = ["hip_yaw", "hip_roll", "hip_pitch", "knee_pitch", "ankle_pitch", "ankle_roll"]
    arms = ["shoulder_pitch", "shoulder_roll", "elbow_roll", "gripper_roll"]
    chains = [("right", legs, "Right_Foot"), ("left", legs, "Left_Foot"), ("left", arms, None), ("right", arms, None)]
    table = {"bodies": {}, "joints": {}}
    L = ['<mujoco model="zbot_export">',
         '  <compiler angle="degree" eulerseq="xyz" autolimits="true" meshdir="meshes"/>',
         '  <option timestep="0.002" iterations="6" ls_iterations="7" solver="Newton" integrator="Euler"/>',
         '  <default>',
         '    <joint armature="0.005" damping="0.1"/>',
         '    <default class="visual"><geom type="mesh" contype="0" conaffinity="0" group="2"/></default>',
         '    <default class="collision"><geom type="box" friction="0.8 0.02 0.001" group="3"/></default>',
         '    <default class="leg"><joint damping="0.3" frictionloss="0.04"/>',
         '      <default class="knee"><joint armature="0.011"/></default>',
         '    </default>',
         '    <default class="arm"><joint frictionloss="0.01"/></default>',
         '    <motor ctrlrange="-3 3"/>',
         '  </default>',
         '  <asset><mesh name="body_mesh" file="body.stl"/><mesh name="link_mesh" file="link.stl"/></asset>',
         '  <worldbody>',
         '    <geom name="ground" type="plane" size="0 0 0.1" friction="0.9 0.005 0.0001" priority="2"/>',
         '    <body name="Z-BOT2_MASTER-BODY-SKELETON" pos="0 0 0.41" euler="0 0 10">',
         '      <freejoint name="floating_base"/>',
         '      <inertial pos="0.001 0 0.02" mass="1.5" fullinertia="0.004 0.005 0.003 0.0002 -0.0001 0.0003"/>',
         '      <geom class="visual" mesh="body_mesh"/>',
         '      <site name="imu_site" pos="0 0 0.03" euler="0 5 0" size="0.01"/>']
    table["bodies"]["Z-BOT2_MASTER-BODY-SKELETON"] = dict(mass=1.5, pos=(0, 0, 0.41))
    for side, names, foot in chains:
        sy = -1.0 if side == "right" else 1.0
        cls = "leg" if foot else "arm"
        indent = "      "
        L.append(f'{indent}<body name="{side}_{names[0]}_link" pos="0 {0.05 * sy} {-0.04 if foot else 0.08}" childclass="{cls}">')
        depth = 1
        for k, jn in enumerate(names):
            name = f"{side}_{jn}"
            if k > 0:
                L.append(f'{indent}{"  " * depth}<body name="{name}_link" pos="0 0 {-0.03 - 0.01 * k}" euler="{2 * k} 0 0">')
                depth += 1
            ax = [(0, 0, 1), (1, 0, 0), (0, 1, 0)][k % 3]
            lo, hi = -30.0 - 5 * k, 40.0 + 3 * k
            mass = float(np.round(rng.uniform(0.05, 0.3), 3))
            jcls = ' class="knee"' if jn == "knee_pitch" else ""
            extra = ' damping="0.45"' if jn == "ankle_roll" else ""
            ip = f'{indent}{"  " * depth}'
            L.append(f'{ip}<joint name="{name}" type="hinge" axis="{ax[0]} {ax[1]} {ax[2]}" range="{lo} {hi}" ref="{1.5 * k}"{jcls}{extra}/>')
            L.append(f'{ip}<inertial pos="0 0 -0.01" mass="{mass}" diaginertia="{2e-5 * (k + 1)} {3e-5 * (k + 1)} {1e-5 * (k + 1)}"/>')
            L.append(f'{ip}<geom class="visual" mesh="link_mesh"/>')
            table["joints"][name] = dict(axis=ax, range=(lo, hi), ref=1.5 * k, mass=mass, cls=cls, knee=jn == "knee_pitch", ankle_roll=jn == "ankle_roll")
        if foot:
            ip = f'{indent}{"  " * (depth + 1)}'
            L.append(f'{indent}{"  " * depth}<body name="{foot}" pos="0.01 0 -0.025">')
            L.append(f'{ip}<inertial pos="0 0 0" mass="0.05" diaginertia="1e-5 4e-5 5e-5"/>')
            L.append(f'{ip}<geom name="{foot}_collision" class="collision" size="0.05 0.025 0.005" pos="0.005 0 -0.002"/>')
            L.append(f'{ip}<geom class="visual" mesh="link_mesh"/>')
            L.append(f'{ip}<site name="{"right_foot" if side == "right" else "left_foot"}" pos="0 0 -0.005" size="0.05 0.025 0.004"/>')
            L.append(f'{indent}{"  " * depth}</body>')
        for _ in range(depth):
            depth -= 1
            L.append(f'{indent}{"  " * depth}</body>')
    for nm, pos in (("imu_link", "0 0 0.05"), ("head", "0.01 0 0.13")):
        L.append(f'      <body name="{nm}" pos="{pos}"><inertial pos="0 0 0" mass="0.02" diaginertia="1e-6 1e-6 1e-6"/><geom class="visual" mesh="link_mesh"/></body>')
    L += ['    </body>', '  </worldbody>', '  <actuator>']
    for side, names, _ in chains:
        for jn in names:
            L.append(f'    <motor name="{side}_{jn}_ctrl" joint="{side}_{jn}" gear="1"/>')
    L += ['  </actuator>', '  <sensor>', '    <framequat name="imu_site_quat" objtype="site" objname="imu_site"/>', '    <gyro name="imu_gyro" site="imu_site"/>',
          '    <accelerometer name="imu_acc" site="imu_site"/>', '  </sensor>', '</mujoco>']
    return "\n".join(L), table


def test_export_style_file_compiles_to_the_expected_model():
    xml, tab = _export_style_xml()
    servo = {n: dict(kp=20.0, kd=0.5, max_pwm=1.0, vin=12.0, kt=0.8, R=1.2, vmax=6.0, amax=40.0, error_gain=0.164) for n in M.JOINT_NAMES}
    c = mjcf.compile_mjcf(xml, servo_params=servo)
    m = c.model
    assert c.joint_names == M.JOINT_NAMES and not c.ignored_geoms  # visual meshes (contype = conaffinity = 0) are not collision geometry
    assert c.opts == {"dt": 0.002, "iterations": 6, "ls_iterations": 7, "solver": M.SOLVER_NEWTON, "integrator": M.INT_EULER}
    assert c.sensor_names == ["imu_site_quat", "imu_gyro", "imu_acc"]
    deg = math.pi / 180
    for j, name in enumerate(M.JOINT_NAMES):
        t, d = tab["joints"][name], 6 + j
        b = int(m.dof_body[d])
        assert c.body_names[b] == (name + "_link")
        np.testing.assert_allclose(m.field("jnt_axis")[b], t["axis"], atol=0)
        np.testing.assert_allclose(m.field("jnt_range")[d], np.array(t["range"]) * deg, rtol=1e-6)   # degrees -> radians
        assert m.jnt_limited[d] == 1                                                                      # autolimits: range given
        assert m.field("qpos0")[d + 1] == pytest.approx(t["ref"] * deg, rel=1e-6)
        assert m.body_mass[b] == pytest.approx(t["mass"], rel=1e-6)
        # default classes: childclass of the chain root, nested `knee` class, per-element override
        assert m.dof_damping[d] == pytest.approx(0.45 if t["ankle_roll"] else (0.3 if t["cls"] == "leg" else 0.1), rel=1e-6)
        assert m.dof_frictionloss[d] == pytest.approx(0.04 if t["cls"] == "leg" else 0.01, rel=1e-6)
        assert m.dof_armature[d] == pytest.approx(0.011 if t["knee"] else 0.005, rel=1e-6)
        assert m.act_ctrllimited[j] == 1 and tuple(m.field("act_ctrlrange")[j]) == (-3.0, 3.0) and m.act_gear[j] == 1.0
        assert m.servo_kp[j] == 20.0
    # tree: parents follow the nesting, dof_parent follows the kinematic chain (first hinge of a chain -> last free-joint dof)
    assert m.body_parent[1] == 0 and int(m.body_jnttype[1]) == 1
    for j, name in enumerate(M.JOINT_NAMES):
        first = name.split("_", 1)[1] in ("hip_yaw", "shoulder_pitch")
        assert m.dof_parent[6 + j] == (5 if first else 6 + j - 1)
    # base: euler z 10 deg -> quaternion; fullinertia diagonalised: same tensor, principal moments sorted as MuJoCo does (descending)
    np.testing.assert_allclose(m.field("qpos0")[:7], [0, 0, 0.41, math.cos(5 * deg), 0, 0, math.sin(5 * deg)], atol=1e-7)
    I = np.array([[0.004, 0.0002, -0.0001], [0.0002, 0.005, 0.0003], [-0.0001, 0.0003, 0.003]])
    R = M.quat_to_mat(np.asarray(m.field("body_iquat")[1], dtype=np.float64))
    np.testing.assert_allclose(R @ np.diag(np.asarray(m.field("body_inertia")[1], dtype=np.float64)) @ R.T, I, atol=2e-9)
    assert np.all(np.diff(m.field("body_inertia")[1]) <= 0)
    # feet: the two collision boxes (class defaults give type + friction), floor priority, sites
    for g, foot in enumerate(("Right_Foot", "Left_Foot")):
        assert c.body_names[int(m.box_body[g])] == foot and m.box_type[g] == M.GEOM_BOX and m.pair_body[g] == 0
        np.testing.assert_allclose(m.field("box_size")[g], (0.05, 0.025, 0.005), rtol=1e-6)
        np.testing.assert_allclose(m.field("box_friction")[g], (0.8, 0.02, 0.001), rtol=1e-6)
        np.testing.assert_allclose(m.field("box_pos")[g], (0.005, 0, -0.002), rtol=1e-6)
    assert m.floor_priority == 2 and m.field("floor_friction")[0] == pytest.approx(0.9)
    assert c.body_names[int(m.site_body[0])] == "Z-BOT2_MASTER-BODY-SKELETON"
    np.testing.assert_allclose(m.field("site_quat")[0], [math.cos(2.5 * deg), 0, math.sin(2.5 * deg), 0], atol=1e-7)
    assert c.body_names[int(m.site_body[1])] == "Left_Foot" and c.body_names[int(m.site_body[2])] == "Right_Foot"
    # mj_setConst ran: total mass, positive inverse weights, and the engine accepts the model's topology (host-side derive only)
    assert M.total_mass(m) == pytest.approx(1.5 + sum(t["mass"] for t in tab["joints"].values()) + 2 * 0.05 + 2 * 0.02, rel=1e-6)
    assert np.all(np.asarray(m.dof_invweight0)[:26] > 0)
    # ... and the oracle steps it: a drop from the compiled default pose stays finite and the feet touch down
    import zb_oracle

    o = zb_oracle.Oracle(m, M.make_opts(dt=0.002), np.float64)
    st, out = o.new_state(1), o.new_out(1)
    q0 = np.asarray(m.field("qpos0"), dtype=np.float64)[None].copy()
    q0[0, 2] = 0.42  # (the feet hang ~0.37 m below the base origin in this pose)
    o.reset(st, q0, np.zeros((1, 26)), np.full(1, 1e9), out=out)
    act = q0[:, 7:].copy()
    touched = 0
    for _ in range(250):
        o.step(st, act, 1, out=out)
        touched |= int(out["contact_mask"][0])
    assert np.isfinite(st["qpos"]).all() and touched != 0 and 0.0 < st["qpos"][0, 2] < 0.45
