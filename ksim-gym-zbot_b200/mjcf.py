"""MJCF -> ``ZbModel`` compiler — host logic, no MuJoCo needed.

The reference loads its robot with ``mujoco.MjModel.from_xml_path(ksim.get_mujoco_model_path("zbot", name="robot"))`` wrapped
in the ``mujoco_scenes`` "smooth" floor (/root/reference/train.py:1326-1331) and reads the servo constants from the K-Scale
metadata (train.py:1333-1437).  Neither the XML nor MuJoCo is available here, so this module restates the part of MuJoCo's
model compiler the hot path needs ([R] = recalled MuJoCo XML semantics):

* ``<compiler angle= eulerseq= autolimits=>``, ``<option timestep= iterations= ls_iterations= solver= integrator= impratio= gravity=>``
* ``<default>`` classes (nested, ``childclass``) for joint / geom / site / motor / general
* ``<body pos= quat=|euler=|axisangle=>`` with ``<inertial pos= quat= mass= diaginertia=|fullinertia=>`` (or inertia from box / sphere
  geoms), ``<freejoint>`` / ``<joint type="free|hinge" axis= pos= range= limited= armature= damping= frictionloss= margin=
  solreflimit= solimplimit= solreffriction= solimpfriction= ref=>``
* ``<geom type="plane|box" size= pos= quat= friction= priority= solref= solimp= margin= gap= contype= conaffinity=>``: the floor plane
  and the two foot boxes (every other geom only contributes inertia when a body has no ``<inertial>``)
* ``<site name= pos= quat= size=>`` (imu_site, left_foot, right_foot — train.py:1453, 1519-1520), ``<actuator><motor joint= gear=
  ctrlrange= forcerange= ctrllimited= forcelimited=>`` (train.py:1358-1362), ``<sensor>`` names are checked only
* mj_setConst: ``dof_invweight0``, ``body_invweight0``, ``meaninertia`` at ``qpos0`` (``model.set_const``)

Bodies are numbered depth-first in document order like MuJoCo does, which makes every serial chain a contiguous block — the
lane layout the kernel relies on; ``zb_engine_create`` re-validates the topology.  ``to_mjcf`` writes a ``ZbModel`` back out (used
to round-trip ``zbot_syn`` in the tests and as a template for the real robot file).
"""

from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from . import model as M

SERVO_FIELDS = ("kp", "kd", "max_pwm", "vin", "kt", "R", "vmax", "amax", "error_gain")  # train.py:1121-1134


class MjcfError(ValueError):
    pass


@dataclass
class CompiledModel:
    model: M.ZbModel
    opts: dict = field(default_factory=dict)        # keyword arguments for model.make_opts found in <option>
    body_names: list = field(default_factory=list)
    joint_names: list = field(default_factory=list)  # hinge joints in qpos order
    sensor_names: list = field(default_factory=list)
    ignored_geoms: list = field(default_factory=list)


# -------------------------------------------------------------------------------------------------
# attribute helpers
# -------------------------------------------------------------------------------------------------
def _floats(s: str | None, n: int | None = None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.size != n:
        if v.size < n and default is not None:  # MuJoCo pads short vectors (size, friction) with the defaults
            v = np.concatenate([v, np.array(default, dtype=np.float64)[v.size:]])
        else:
            raise MjcfError(f"expected {n} numbers, got '{s}'")
    return v


def _bool(s: str | None, default=None):
    if s is None:
        return default
    if s in ("true", "1"):
        return True
    if s in ("false", "0"):
        return False
    if s == "auto":
        return None
    raise MjcfError(f"bad boolean '{s}'")


class _Defaults:
    """<default> tree: class name -> {element tag -> attributes}, children inherit from parents."""

    def __init__(self, root: ET.Element | None):
        self.classes: dict[str, dict[str, dict[str, str]]] = {"main": {}}
        if root is not None:
            self._walk(root, None)

    def _walk(self, node: ET.Element, parent: str | None):
        name = node.get("class", "main")
        base = {k: dict(v) for k, v in self.classes.get(parent, {}).items()} if parent else {}
        for child in node:
            if child.tag == "default":
                continue
            base.setdefault(child.tag, {}).update(child.attrib)
        self.classes[name] = base
        for child in node:
            if child.tag == "default":
                self._walk(child, name)

    def resolve(self, elem: ET.Element, childclass: str | None) -> dict[str, str]:
        cls = elem.get("class") or childclass or "main"
        if cls not in self.classes:
            raise MjcfError(f"unknown default class '{cls}'")
        tag = "joint" if elem.tag == "freejoint" else elem.tag
        out = dict(self.classes[cls].get(tag, {}))
        out.update(elem.attrib)
        return out


def _orientation(a: dict, degree: bool, eulerseq: str) -> np.ndarray:
    if "quat" in a:
        q = _floats(a["quat"], 4)
        return q / np.linalg.norm(q)
    if "euler" in a:
        e = _floats(a["euler"], 3) * (math.pi / 180.0 if degree else 1.0)
        q = np.array([1.0, 0.0, 0.0, 0.0])
        for ch, ang in zip(eulerseq, e):  # lower case = intrinsic (rotating frame), upper case = extrinsic
            ax = np.eye(3)["xyz".index(ch.lower())]
            r = M.axis_angle_quat(ax, ang)
            q = M.quat_mul(q, r) if ch.islower() else M.quat_mul(r, q)
        return q
    if "axisangle" in a:
        v = _floats(a["axisangle"], 4)
        ang = v[3] * (math.pi / 180.0 if degree else 1.0)
        return M.axis_angle_quat(v[:3] / np.linalg.norm(v[:3]), ang)
    for k in ("xyaxes", "zaxis"):
        if k in a:
            raise MjcfError(f"orientation attribute '{k}' is not supported")
    return np.array([1.0, 0.0, 0.0, 0.0])


def _principal(full: np.ndarray):
    """fullinertia (xx yy zz xy xz yz) -> (diaginertia, quat) with a right-handed eigenbasis, largest moment first."""
    xx, yy, zz, xy, xz, yz = full
    w, v = np.linalg.eigh(np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]]))
    order = np.argsort(-w)
    w, v = w[order], v[:, order]
    if np.linalg.det(v) < 0:
        v[:, 2] = -v[:, 2]
    return w, _mat_to_quat(v)


def _mat_to_quat(R: np.ndarray) -> np.ndarray:
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def _geom_inertia(a: dict):
    """(mass, com, inertia about com in the body frame) of a box / sphere geom ([R] density default 1000)."""
    typ = a.get("type", "sphere")
    pos = _floats(a.get("pos"), 3, (0, 0, 0))
    size = _floats(a.get("size"))
    if typ == "box":
        vol = 8 * size[0] * size[1] * size[2]
        mass = float(a["mass"]) if "mass" in a else float(a.get("density", 1000.0)) * vol
        d = 2 * size
        I = mass / 12.0 * np.array([d[1] ** 2 + d[2] ** 2, d[0] ** 2 + d[2] ** 2, d[0] ** 2 + d[1] ** 2])
    elif typ == "sphere":
        vol = 4.0 / 3.0 * math.pi * size[0] ** 3
        mass = float(a["mass"]) if "mass" in a else float(a.get("density", 1000.0)) * vol
        I = np.full(3, 0.4 * mass * size[0] ** 2)
    else:
        raise MjcfError(f"inertia from geom type '{typ}' is not supported: give the body an <inertial>")
    return mass, pos, I


# -------------------------------------------------------------------------------------------------
# the compiler
# -------------------------------------------------------------------------------------------------
def compile_mjcf(xml: str, servo_params: dict | None = None, floor: dict | None = None,
                 site_names=("imu_site", "left_foot", "right_foot"), joint_order: list | None = None) -> CompiledModel:
    """Compile an MJCF string into a ``ZbModel``.

    ``servo_params``: joint name -> dict with the Feetech fields of train.py:1121-1134 (``kp kd max_pwm vin kt R vmax amax
    error_gain``) — the K-Scale metadata the reference reads at train.py:1340-1437.  ``floor``: overrides for the floor plane
    when the XML has none (the reference adds it through ``mujoco_scenes``, train.py:1328-1330: priority 2).
    ``joint_order``: expected hinge order (defaults to train.py:59-82); a mismatch is an error because actions are positional.
    """
    root = ET.fromstring(xml)
    if root.tag != "mujoco":
        raise MjcfError("root element must be <mujoco>")
    comp = root.find("compiler")
    degree = (comp.get("angle", "degree") if comp is not None else "degree") == "degree"
    eulerseq = comp.get("eulerseq", "xyz") if comp is not None else "xyz"
    autolimits = _bool(comp.get("autolimits"), True) if comp is not None else True
    defaults = _Defaults(root.find("default"))
    out = CompiledModel(model=M.ZbModel())
    m = out.model
    ang = math.pi / 180.0 if degree else 1.0

    opt = root.find("option")
    if opt is not None:
        if "timestep" in opt.attrib:
            out.opts["dt"] = float(opt.get("timestep"))
        for k in ("iterations", "ls_iterations"):
            if k in opt.attrib:
                out.opts[k] = int(opt.get(k))
        for k in ("tolerance", "ls_tolerance", "impratio"):
            if k in opt.attrib:
                out.opts[k] = float(opt.get(k))
        if "gravity" in opt.attrib:
            out.opts["gravity"] = tuple(_floats(opt.get("gravity"), 3))
        if "solver" in opt.attrib:
            out.opts["solver"] = {"CG": M.SOLVER_CG, "Newton": M.SOLVER_NEWTON}[opt.get("solver")]
        if "integrator" in opt.attrib:
            out.opts["integrator"] = {"Euler": M.INT_EULER, "implicitfast": M.INT_IMPLICITFAST}[opt.get("integrator")]
        if opt.get("cone", "pyramidal") != "pyramidal":
            raise MjcfError("only the pyramidal friction cone is built into the kernels")

    world = root.find("worldbody")
    if world is None:
        raise MjcfError("no <worldbody>")

    bodies: list[dict] = [dict(name="world", parent=0, pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]), joint=None, geoms=[], sites=[],
                               inertial=None)]
    floor_geom = None

    def visit(elem: ET.Element, parent: int, childclass: str | None):
        nonlocal floor_geom
        for child in elem:
            if child.tag == "geom":
                a = defaults.resolve(child, childclass)
                if parent == 0 and a.get("type", "sphere") == "plane":
                    floor_geom = a
                else:
                    bodies[parent]["geoms"].append(a)
            elif child.tag == "site":
                bodies[parent]["sites"].append(defaults.resolve(child, childclass))
            elif child.tag == "inertial":
                bodies[parent]["inertial"] = dict(child.attrib)
            elif child.tag in ("joint", "freejoint"):
                a = defaults.resolve(child, childclass)
                if child.tag == "freejoint":
                    a["type"] = "free"
                if bodies[parent]["joint"] is not None:
                    raise MjcfError(f"body '{bodies[parent]['name']}': more than one joint per body is not supported")
                bodies[parent]["joint"] = a
            elif child.tag == "body":
                a = child.attrib
                b = dict(name=a.get("name", f"body{len(bodies)}"), parent=parent, pos=_floats(a.get("pos"), 3, (0, 0, 0)),
                         quat=_orientation(a, degree, eulerseq), joint=None, geoms=[], sites=[], inertial=None)
                bodies.append(b)
                visit(child, len(bodies) - 1, a.get("childclass", childclass))

    visit(world, 0, None)
    if len(bodies) != M.NBODY:
        raise MjcfError(f"the kernels are built for {M.NBODY} bodies incl. the world (train.py:42-43), the file has {len(bodies)}")
    out.body_names = [b["name"] for b in bodies]

    # ---- tree, joints, inertias ----
    hinge_names, d = [], 0
    solref0, solimp0 = (0.02, 1.0), (0.9, 0.95, 0.001, 0.5, 2.0)
    for k in range(M.NV):
        m.field("jnt_solref")[k] = solref0
        m.field("jnt_solimp")[k] = solimp0
        m.field("dof_solref")[k] = solref0
        m.field("dof_solimp")[k] = solimp0
    m.field("qpos0")[:] = 0.0
    for i, b in enumerate(bodies):
        m.body_parent[i] = b["parent"]
        m.field("body_pos")[i] = b["pos"]
        m.field("body_quat")[i] = b["quat"]
        m.field("body_iquat")[i] = (1, 0, 0, 0)
        m.field("jnt_axis")[i] = (0, 0, 1)
        if i > 0:
            if b["inertial"] is not None:
                a = b["inertial"]
                m.body_mass[i] = float(a["mass"])
                m.field("body_ipos")[i] = _floats(a.get("pos"), 3, (0, 0, 0))
                iq = _orientation(a, degree, eulerseq)
                if "fullinertia" in a:
                    diag, q2 = _principal(_floats(a["fullinertia"], 6))
                    iq = M.quat_mul(iq, q2)
                else:
                    diag = _floats(a.get("diaginertia"), 3)
                    if diag is None:
                        raise MjcfError(f"body '{b['name']}': <inertial> needs diaginertia or fullinertia")
                m.field("body_inertia")[i] = diag
                m.field("body_iquat")[i] = iq
            elif b["geoms"]:
                parts = [_geom_inertia(g) for g in b["geoms"] if _orientation(g, degree, eulerseq)[0] > 1 - 1e-12]
                if len(parts) != len(b["geoms"]):
                    raise MjcfError(f"body '{b['name']}': inertia from rotated geoms is not supported: add an <inertial>")
                mass = sum(p[0] for p in parts)
                com = sum(p[0] * p[1] for p in parts) / mass
                I = np.zeros(3)
                for pm, pc, pI in parts:
                    r = pc - com
                    I += pI + pm * (r @ r - r * r)
                    if len(parts) > 1 and np.count_nonzero(r) > 1:
                        raise MjcfError(f"body '{b['name']}': off-axis geoms give a non-diagonal inertia: add an <inertial>")
                m.body_mass[i], m.field("body_ipos")[i], m.field("body_inertia")[i] = mass, com, I
            else:
                raise MjcfError(f"body '{b['name']}' has neither <inertial> nor geoms")
        j = b["joint"]
        if j is None:
            m.body_jnttype[i], m.body_dofadr[i] = 0, -1
            continue
        typ = j.get("type", "hinge")
        if typ == "free":
            if d != 0 or b["parent"] != 0:
                raise MjcfError("the free joint must belong to the first body under the world")
            m.body_jnttype[i], m.body_dofadr[i] = 1, 0
            for k in range(6):
                m.dof_body[k], m.dof_parent[k] = i, k - 1
            m.field("qpos0")[0:3] = b["pos"]
            m.field("qpos0")[3:7] = b["quat"]
            d = 6
        elif typ == "hinge":
            if d + 1 > M.NV:
                raise MjcfError(f"more than {M.NU} hinge joints")
            m.body_jnttype[i], m.body_dofadr[i] = 2, d
            m.dof_body[d] = i
            a_ = b["parent"]
            while a_ > 0 and m.body_dofadr[a_] < 0:
                a_ = m.body_parent[a_]
            m.dof_parent[d] = (5 if m.body_jnttype[a_] == 1 else m.body_dofadr[a_]) if a_ > 0 else -1
            ax = _floats(j.get("axis"), 3, (0, 0, 1))
            m.field("jnt_axis")[i] = ax / np.linalg.norm(ax)
            m.field("jnt_pos")[i] = _floats(j.get("pos"), 3, (0, 0, 0))
            rng = _floats(j.get("range"), 2)
            limited = _bool(j.get("limited"))
            if limited is None:
                limited = autolimits and rng is not None
            if limited and rng is None:
                raise MjcfError(f"joint '{j.get('name')}': limited without range")
            if rng is not None:
                m.field("jnt_range")[d] = rng * ang
            m.jnt_limited[d] = int(bool(limited))
            m.jnt_margin[d] = float(j.get("margin", 0.0))
            m.dof_armature[d] = float(j.get("armature", 0.0))
            m.dof_damping[d] = float(j.get("damping", 0.0))
            m.dof_frictionloss[d] = float(j.get("frictionloss", 0.0))
            m.field("qpos0")[d + 1] = float(j.get("ref", 0.0)) * ang
            for key, dst, n in (("solreflimit", "jnt_solref", 2), ("solimplimit", "jnt_solimp", 5), ("solreffriction", "dof_solref", 2),
                                ("solimpfriction", "dof_solimp", 5)):
                if key in j:
                    m.field(dst)[d] = _floats(j[key], n, solref0 if n == 2 else solimp0)
            hinge_names.append(j.get("name", f"joint{d}"))
            d += 1
        else:
            raise MjcfError(f"joint type '{typ}' is not supported (free + hinge only)")
    if d != M.NV:
        raise MjcfError(f"the kernels are built for nv = {M.NV} (free joint + {M.NU} hinges, train.py:27), the file has {d}")
    want = joint_order if joint_order is not None else M.JOINT_NAMES
    names = [n.removeprefix("dof_") for n in hinge_names]
    if names != list(want):
        raise MjcfError(f"hinge joint order {names} differs from the action order of train.py:59-82 {list(want)}")
    out.joint_names = hinge_names

    # ---- floor + foot boxes (collision pairs of the hot path: box vs plane) ----
    fl = dict(floor_geom or {})
    fl.update(floor or {})
    if not fl:
        raise MjcfError("no floor plane: add <geom type='plane'> to the worldbody or pass floor={...} (mujoco_scenes adds it, train.py:1328)")
    m.field("floor_pos")[:] = _floats(fl.get("pos"), 3, (0, 0, 0))
    fq = _orientation(fl, degree, eulerseq)
    m.field("floor_normal")[:] = M.quat_to_mat(fq)[:, 2]
    fr0 = (1.0, 0.005, 0.0001)
    m.field("floor_friction")[:] = _floats(fl.get("friction"), 3, fr0)
    m.field("floor_solref")[:] = _floats(fl.get("solref"), 2, solref0)
    m.field("floor_solimp")[:] = _floats(fl.get("solimp"), 5, solimp0)
    m.floor_priority = int(fl.get("priority", 2))  # train.py:1330 sets geom_priority[floor] = 2
    m.floor_margin, m.floor_gap = float(fl.get("margin", 0.0)), float(fl.get("gap", 0.0))
    # Collision groups (ZbModel.box_* / pair_*).  Explicit <contact><pair geom1 geom2> elements become body-body groups (geom 1 =
    # pair_*, geom 2 = box_*; MJX orders a pair by geom type: sphere before capsule); every other colliding box / sphere / capsule geom on a
    # chain body collides with the floor plane.  Anything else (meshes, cylinders, geoms beyond the ZB_NBOX groups) is reported in
    # `ignored_geoms` — the kernel has no primitive for it.
    types = {"box": M.GEOM_BOX, "sphere": M.GEOM_SPHERE, "capsule": M.GEOM_CAPSULE}
    by_name = {}
    for i, b in enumerate(bodies):
        for g in b["geoms"]:
            if g.get("name"):
                by_name[g["name"]] = (i, g)

    def geom_size(g):
        t = g.get("type", "sphere")
        n_ = {"box": 3, "sphere": 1, "capsule": 2}[t]
        sz = np.zeros(3)
        sz[:n_] = _floats(g.get("size"), n_)
        return sz

    def fill_geom2(k, i, g):
        m.box_body[k], m.box_type[k] = i, types[g.get("type", "sphere")]
        m.field("box_pos")[k] = _floats(g.get("pos"), 3, (0, 0, 0))
        m.field("box_quat")[k] = _orientation(g, degree, eulerseq)
        m.field("box_size")[k] = geom_size(g)
        m.field("box_friction")[k] = _floats(g.get("friction"), 3, fr0)
        m.field("box_solref")[k] = _floats(g.get("solref"), 2, solref0)
        m.field("box_solimp")[k] = _floats(g.get("solimp"), 5, solimp0)
        m.box_priority[k] = int(g.get("priority", 0))
        m.box_margin[k], m.box_gap[k] = float(g.get("margin", 0.0)), float(g.get("gap", 0.0))
        m.pair_body[k] = 0
        m.field("pair_quat")[k] = (1, 0, 0, 0)

    pairs = []  # (geom-2 slot order is fixed later: floor groups first in body order, like geom ids)
    contact = root.find("contact")
    paired = set()
    if contact is not None:
        for pe in contact.findall("pair"):
            n1, n2 = pe.get("geom1"), pe.get("geom2")
            if n1 not in by_name or n2 not in by_name:
                raise MjcfError(f"<pair>: unknown geom {n1!r} / {n2!r}")
            (i1, g1), (i2, g2) = by_name[n1], by_name[n2]
            order = {"sphere": 0, "capsule": 1}
            t1, t2 = g1.get("type", "sphere"), g2.get("type", "sphere")
            if t1 not in order or t2 not in order:
                raise MjcfError("<pair>: body-body groups support sphere-sphere, sphere-capsule and capsule-capsule")
            if order[t1] > order[t2]:
                (i1, g1, n1), (i2, g2, n2) = (i2, g2, n2), (i1, g1, n1)
            pairs.append((i1, g1, i2, g2))
            paired |= {n1, n2}
    nbox = 0
    for i, b in enumerate(bodies):
        for g in b["geoms"]:
            collides = int(g.get("contype", 1)) != 0 or int(g.get("conaffinity", 1)) != 0
            if g.get("name") in paired:
                continue
            if g.get("type", "sphere") in types and collides and i >= 2 and nbox + len(pairs) < M.NBOX:
                fill_geom2(nbox, i, g)
                nbox += 1
            elif collides:
                out.ignored_geoms.append((b["name"], g.get("name"), g.get("type", "sphere")))
    for (i1, g1, i2, g2) in pairs:
        if nbox >= M.NBOX:
            raise MjcfError(f"more than {M.NBOX} collision groups")
        fill_geom2(nbox, i2, g2)
        m.pair_body[nbox], m.pair_type[nbox] = i1, types[g1.get("type", "sphere")]
        m.field("pair_pos")[nbox] = _floats(g1.get("pos"), 3, (0, 0, 0))
        m.field("pair_quat")[nbox] = _orientation(g1, degree, eulerseq)
        m.field("pair_size")[nbox] = geom_size(g1)
        m.field("pair_friction")[nbox] = _floats(g1.get("friction"), 3, fr0)
        m.field("pair_solref")[nbox] = _floats(g1.get("solref"), 2, solref0)
        m.field("pair_solimp")[nbox] = _floats(g1.get("solimp"), 5, solimp0)
        m.pair_priority[nbox] = int(g1.get("priority", 0))
        m.pair_margin[nbox], m.pair_gap[nbox] = float(g1.get("margin", 0.0)), float(g1.get("gap", 0.0))
        nbox += 1
    if nbox != M.NBOX:
        raise MjcfError(f"expected {M.NBOX} collision groups (colliding box / sphere / capsule geoms on chain bodies, <pair> elements), found {nbox}")

    # ---- sites ----
    found = {}
    for i, b in enumerate(bodies):
        for s in b["sites"]:
            if s.get("name") in site_names:
                found[s["name"]] = (i, s)
    for k, name in enumerate(site_names):
        if name not in found:
            raise MjcfError(f"site '{name}' not found (train.py:1453, 1519-1520)")
        i, s = found[name]
        m.site_body[k] = i
        m.field("site_pos")[k] = _floats(s.get("pos"), 3, (0, 0, 0))
        m.field("site_quat")[k] = _orientation(s, degree, eulerseq)
        m.field("site_size")[k] = _floats(s.get("size"), 3, (0.005, 0.005, 0.005))

    # ---- motors (named <joint>_ctrl, train.py:1358-1362) in action order ----
    act = root.find("actuator")
    motors = {}
    if act is not None:
        for a in act:
            if a.tag not in ("motor", "general"):
                raise MjcfError(f"actuator type '{a.tag}' is not supported (the reference drives motors with servo torques)")
            r = defaults.resolve(a, None)
            motors[r.get("joint")] = r
    for jn, (dofname) in enumerate(hinge_names):
        if dofname not in motors:
            raise MjcfError(f"no motor for joint '{dofname}'")
        r = motors[dofname]
        m.act_dof[jn] = 6 + jn
        m.act_gear[jn] = _floats(r.get("gear"), None, (1.0,))[0]
        for key, dst, flag in (("ctrlrange", "act_ctrlrange", "act_ctrllimited"), ("forcerange", "act_forcerange", "act_forcelimited")):
            rng = _floats(r.get(key), 2)
            lim = _bool(r.get(key.replace("range", "limited")))
            if lim is None:
                lim = autolimits and rng is not None
            if rng is not None:
                m.field(dst)[jn] = rng
            getattr(m, flag)[jn] = int(bool(lim))
    sens = root.find("sensor")
    out.sensor_names = [s.get("name") for s in sens] if sens is not None else []

    # ---- Feetech constants (K-Scale metadata, train.py:1340-1437) ----
    if servo_params is not None:
        for jn, name in enumerate(names):
            p = servo_params.get(name) or servo_params.get(hinge_names[jn])
            if p is None:
                raise MjcfError(f"servo_params has no entry for joint '{name}'")
            for f in SERVO_FIELDS:
                getattr(m, "servo_" + f)[jn] = float(p[f])
    enc = 0.087 * math.pi / 180.0  # train.py:1111-1118
    m.servo_deadband_pos = m.servo_deadband_neg = 2 * enc
    M.set_const(m)
    return out


# -------------------------------------------------------------------------------------------------
# the inverse: ZbModel -> MJCF text
# -------------------------------------------------------------------------------------------------
def _fmt(v) -> str:
    return " ".join(repr(float(np.float32(x))) if abs(float(x)) < 1e30 else repr(float(x)) for x in np.atleast_1d(np.asarray(v, dtype=np.float64)))


def to_mjcf(m: M.ZbModel, body_names=None, joint_names=None, opts: M.ZbOpts | None = None) -> str:
    """Write a ``ZbModel`` as MJCF (radians, explicit inertials; fp32 values printed with round-trip precision)."""
    body_names = body_names or M.BODY_NAMES
    joint_names = joint_names or M.JOINT_NAMES
    site_names = ("imu_site", "left_foot", "right_foot")
    L = ["<mujoco model=\"zbot\">", "  <compiler angle=\"radian\" autolimits=\"true\"/>"]
    if opts is not None:
        L.append(f"  <option timestep=\"{_fmt(opts.dt)}\" iterations=\"{opts.iterations}\" ls_iterations=\"{opts.ls_iterations}\" "
                 f"solver=\"{'Newton' if opts.solver == M.SOLVER_NEWTON else 'CG'}\" integrator=\"{'implicitfast' if opts.integrator else 'Euler'}\" "
                 f"impratio=\"{_fmt(opts.impratio)}\" gravity=\"{_fmt(list(opts.gravity))}\"/>")
    L.append("  <worldbody>")
    L.append(f"    <geom name=\"floor\" type=\"plane\" size=\"0 0 0.05\" pos=\"{_fmt(m.field('floor_pos'))}\" friction=\"{_fmt(m.field('floor_friction'))}\" "
             f"solref=\"{_fmt(m.field('floor_solref'))}\" solimp=\"{_fmt(m.field('floor_solimp'))}\" priority=\"{m.floor_priority}\" "
             f"margin=\"{_fmt(m.floor_margin)}\" gap=\"{_fmt(m.floor_gap)}\"/>")
    children: dict[int, list[int]] = {}
    for b in range(1, M.NBODY):
        children.setdefault(int(m.body_parent[b]), []).append(b)

    def emit(b: int, ind: str):
        L.append(f"{ind}<body name=\"{body_names[b]}\" pos=\"{_fmt(m.field('body_pos')[b])}\" quat=\"{_fmt(m.field('body_quat')[b])}\">")
        L.append(f"{ind}  <inertial pos=\"{_fmt(m.field('body_ipos')[b])}\" quat=\"{_fmt(m.field('body_iquat')[b])}\" mass=\"{_fmt(m.body_mass[b])}\" "
                 f"diaginertia=\"{_fmt(m.field('body_inertia')[b])}\"/>")
        if m.body_jnttype[b] == 1:
            L.append(f"{ind}  <freejoint name=\"floating_base\"/>")
        elif m.body_jnttype[b] == 2:
            d = int(m.body_dofadr[b])
            L.append(f"{ind}  <joint name=\"{joint_names[d - 6]}\" type=\"hinge\" axis=\"{_fmt(m.field('jnt_axis')[b])}\" pos=\"{_fmt(m.field('jnt_pos')[b])}\" "
                     f"range=\"{_fmt(m.field('jnt_range')[d])}\" limited=\"{'true' if m.jnt_limited[d] else 'false'}\" margin=\"{_fmt(m.jnt_margin[d])}\" "
                     f"armature=\"{_fmt(m.dof_armature[d])}\" damping=\"{_fmt(m.dof_damping[d])}\" frictionloss=\"{_fmt(m.dof_frictionloss[d])}\" "
                     f"ref=\"{_fmt(m.field('qpos0')[d + 1])}\" solreflimit=\"{_fmt(m.field('jnt_solref')[d])}\" solimplimit=\"{_fmt(m.field('jnt_solimp')[d])}\" "
                     f"solreffriction=\"{_fmt(m.field('dof_solref')[d])}\" solimpfriction=\"{_fmt(m.field('dof_solimp')[d])}\"/>")
        tname, nsz = {M.GEOM_BOX: "box", M.GEOM_SPHERE: "sphere", M.GEOM_CAPSULE: "capsule"}, {M.GEOM_BOX: 3, M.GEOM_SPHERE: 1, M.GEOM_CAPSULE: 2}
        for g in range(M.NBOX):
            if m.box_body[g] == b:
                t = int(m.box_type[g])
                L.append(f"{ind}  <geom name=\"group{g}_geom2\" type=\"{tname[t]}\" size=\"{_fmt(m.field('box_size')[g][:nsz[t]])}\" pos=\"{_fmt(m.field('box_pos')[g])}\" "
                         f"quat=\"{_fmt(m.field('box_quat')[g])}\" friction=\"{_fmt(m.field('box_friction')[g])}\" solref=\"{_fmt(m.field('box_solref')[g])}\" "
                         f"solimp=\"{_fmt(m.field('box_solimp')[g])}\" priority=\"{m.box_priority[g]}\" margin=\"{_fmt(m.box_margin[g])}\" gap=\"{_fmt(m.box_gap[g])}\"/>")
            if m.pair_body[g] == b and b != 0:
                t = int(m.pair_type[g])
                L.append(f"{ind}  <geom name=\"group{g}_geom1\" type=\"{tname[t]}\" size=\"{_fmt(m.field('pair_size')[g][:nsz[t]])}\" pos=\"{_fmt(m.field('pair_pos')[g])}\" "
                         f"quat=\"{_fmt(m.field('pair_quat')[g])}\" friction=\"{_fmt(m.field('pair_friction')[g])}\" solref=\"{_fmt(m.field('pair_solref')[g])}\" "
                         f"solimp=\"{_fmt(m.field('pair_solimp')[g])}\" priority=\"{m.pair_priority[g]}\" margin=\"{_fmt(m.pair_margin[g])}\" gap=\"{_fmt(m.pair_gap[g])}\"/>")
        for s in range(M.NSITE):
            if m.site_body[s] == b:
                L.append(f"{ind}  <site name=\"{site_names[s]}\" type=\"box\" pos=\"{_fmt(m.field('site_pos')[s])}\" quat=\"{_fmt(m.field('site_quat')[s])}\" "
                         f"size=\"{_fmt(m.field('site_size')[s])}\"/>")
        for c in children.get(b, []):
            emit(c, ind + "  ")
        L.append(f"{ind}</body>")

    for b in children.get(0, []):
        emit(b, "    ")
    L.append("  </worldbody>")
    if any(m.pair_body[g] != 0 for g in range(M.NBOX)):
        L.append("  <contact>")
        L += [f"    <pair geom1=\"group{g}_geom1\" geom2=\"group{g}_geom2\"/>" for g in range(M.NBOX) if m.pair_body[g] != 0]
        L.append("  </contact>")
    L.append("  <actuator>")
    for j in range(M.NU):
        L.append(f"    <motor name=\"{joint_names[j]}_ctrl\" joint=\"{joint_names[j]}\" gear=\"{_fmt(m.act_gear[j])}\" ctrlrange=\"{_fmt(m.field('act_ctrlrange')[j])}\" "
                 f"ctrllimited=\"{'true' if m.act_ctrllimited[j] else 'false'}\" forcerange=\"{_fmt(m.field('act_forcerange')[j])}\" "
                 f"forcelimited=\"{'true' if m.act_forcelimited[j] else 'false'}\"/>")
    L.append("  </actuator>")
    L.append("  <sensor>")  # train.py:44-46, 1495-1516
    L += ["    <framequat name=\"imu_site_quat\" objtype=\"site\" objname=\"imu_site\"/>", "    <accelerometer name=\"imu_acc\" site=\"imu_site\"/>",
          "    <gyro name=\"imu_gyro\" site=\"imu_site\"/>", "    <touch name=\"left_foot_touch\" site=\"left_foot\"/>",
          "    <touch name=\"right_foot_touch\" site=\"right_foot\"/>", "    <force name=\"left_foot_force\" site=\"left_foot\"/>",
          "    <force name=\"right_foot_force\" site=\"right_foot\"/>"]
    L += ["  </sensor>", "</mujoco>"]
    return "\n".join(L) + "\n"


def servo_params_of(m: M.ZbModel, joint_names=None) -> dict:
    """The Feetech constants of a ``ZbModel`` as the per-joint metadata dict ``compile_mjcf`` accepts."""
    joint_names = joint_names or M.JOINT_NAMES
    return {n: {f: float(getattr(m, "servo_" + f)[j]) for f in SERVO_FIELDS} for j, n in enumerate(joint_names)}
