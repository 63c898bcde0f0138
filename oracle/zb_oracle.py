"""ctypes front-end of the CPU oracle (oracle/zb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Allowed importers: tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of bench.py.  The product package never imports this module.

PARITY UNPINNED — see the header of zb_oracle.c: the restated algorithm lives in mujoco-mjx 3.3.4 and
ksim 0.1.99 (/root/reference/uv.lock:2226-2227, 1872-1873), neither vendored nor installable here, and
the reference has no tests or golden vectors for this path.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

NQ, NV, NU, NBODY, NSITE, NBOX, MAXCON = 27, 26, 20, 26, 3, 2, 8
NEFC = 2 * NV + 4 * MAXCON
NSENSORDATA = 18
NTRACE, TRACE_POINTS, TRACE_SEARCH, TRACE_START = 160, 32, 128, 154  # include/zbphys.h ZB_NTRACE / ZB_TRACE_*
STAGES = [
    "kinematics", "com_pos", "crb", "factor_m", "collision", "make_constraint", "com_vel", "rne",
    "actuation", "fwd_acceleration", "solver", "sensor", "integrate",
]

_REAL_FIELDS = [
    ("qpos", (NQ,)), ("qvel", (NV,)), ("ctrl", (NU,)), ("qacc_warmstart", (NV,)),
    ("xpos", (NBODY, 3)), ("xquat", (NBODY, 4)), ("xmat", (NBODY, 9)), ("xipos", (NBODY, 3)), ("ximat", (NBODY, 9)),
    ("xanchor", (NBODY, 3)), ("xaxis", (NBODY, 3)), ("subtree_com", (NBODY, 3)),
    ("cinert", (NBODY, 10)), ("crb", (NBODY, 10)), ("cdof", (NV, 6)), ("cdof_dot", (NV, 6)), ("cvel", (NBODY, 6)),
    ("qM", (NV, NV)), ("qL", (NV, NV)),
    ("qfrc_passive", (NV,)), ("qfrc_bias", (NV,)), ("qfrc_actuator", (NV,)), ("qfrc_smooth", (NV,)), ("qacc_smooth", (NV,)),
    ("actuator_force", (NU,)),
    ("site_xpos", (NSITE, 3)), ("site_xmat", (NSITE, 9)), ("box_xpos", (NBOX, 3)), ("box_xmat", (NBOX, 9)), ("pair_xpos", (NBOX, 3)), ("pair_xmat", (NBOX, 9)),
    ("con_dist", (MAXCON,)), ("con_pos", (MAXCON, 3)), ("con_frame", (MAXCON, 9)), ("con_mu", (MAXCON,)),
    ("con_includemargin", (MAXCON,)), ("con_solref", (MAXCON, 2)), ("con_solimp", (MAXCON, 5)),
    ("efc_J", (NEFC, NV)), ("efc_pos", (NEFC,)), ("efc_D", (NEFC,)), ("efc_R", (NEFC,)), ("efc_aref", (NEFC,)),
    ("efc_floss", (NEFC,)), ("efc_force", (NEFC,)),
    ("qfrc_constraint", (NV,)), ("qacc", (NV,)),
    ("cacc", (NBODY, 6)), ("cfrc_int", (NBODY, 6)), ("cfrc_ext", (NBODY, 6)),
    ("sensordata", (NSENSORDATA,)),
    ("solver_cost", ()), ("trace", (NTRACE,)),
]
_INT_FIELDS = [
    ("con_body", (MAXCON,)), ("con_body1", (MAXCON,)), ("con_vert", (MAXCON,)), ("con_active", (MAXCON,)), ("efc_type", (NEFC,)),
    ("solver_niter", ()), ("ls_evals", ()), ("ls_iters", ()), ("flags", ()),
]


def _ctype(base, shape):
    t = base
    for n in reversed(shape):
        t = t * n
    return t


def _make_data(real):
    class ZboData(ctypes.Structure):
        _fields_ = [(n, _ctype(real, s)) for n, s in _REAL_FIELDS] + [(n, _ctype(ctypes.c_int32, s)) for n, s in _INT_FIELDS]

        def get(self, name):
            v = getattr(self, name)
            if isinstance(v, (int, float)):
                return v
            return np.ctypeslib.as_array(v).copy()

        def set(self, name, value):
            np.ctypeslib.as_array(getattr(self, name))[...] = value

    return ZboData


ZboDataF32 = _make_data(ctypes.c_float)
ZboDataF64 = _make_data(ctypes.c_double)

_PARAM_FIELDS = ["body_mass", "body_inertia", "dof_armature", "dof_damping", "dof_frictionloss", "qpos0",
                 "floor_friction", "imu_site_pos", "imu_site_quat", "prev_action", "latency_steps"]
_STATE_FIELDS = ["qpos", "qvel", "qacc_warmstart", "act_pos", "act_vel", "act_torque", "push_t_rem", "time"]
_OUT_REAL = ["xpos", "xquat", "cinert", "cvel", "qacc", "qfrc_constraint", "actuator_force", "ctrl", "sensordata",
             "site_xpos", "site_xmat", "contact_dist", "contact_pos", "efc_force"]
_OUT_INT = ["contact_mask", "solver_niter", "flags"]
OUT_SHAPES = {
    "xpos": (NBODY, 3), "xquat": (NBODY, 4), "cinert": (NBODY, 10), "cvel": (NBODY, 6), "qacc": (NV,),
    "qfrc_constraint": (NV,), "actuator_force": (NU,), "ctrl": (NU,), "sensordata": (NSENSORDATA,),
    "site_xpos": (NSITE, 3), "site_xmat": (NSITE, 9), "contact_dist": (MAXCON,), "contact_pos": (MAXCON, 3),
    "efc_force": (NEFC,), "contact_mask": (), "solver_niter": (), "flags": (),
}
STATE_SHAPES = {"qpos": (NQ,), "qvel": (NV,), "qacc_warmstart": (NV,), "act_pos": (NU,), "act_vel": (NU,),
                "act_torque": (NU,), "push_t_rem": (), "time": ()}


class _Ptrs(ctypes.Structure):
    pass


def _ptr_struct(names):
    return type("S", (ctypes.Structure,), {"_fields_": [(n, ctypes.c_void_p) for n in names]})


_ZboParams = _ptr_struct(_PARAM_FIELDS)
_ZboState = _ptr_struct(_STATE_FIELDS)
_ZboOut = _ptr_struct(_OUT_REAL + _OUT_INT + ["solver_trace"])


def build(force: bool = False) -> None:
    """Compile oracle/_build/libzb_oracle*.so with gcc (oracle/Makefile)."""
    if force or not os.path.exists(os.path.join(_BUILD, "libzb_oracle.so")) or (
        os.path.getmtime(os.path.join(_BUILD, "libzb_oracle.so")) < os.path.getmtime(os.path.join(_HERE, "zb_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def _host_key() -> str:
    """Identifies the CPU this process runs on (model name + ISA flags): -march=native objects are only valid there."""
    import hashlib

    txt = ""
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith(("model name", "flags")):
                txt += ln
                if ln.startswith("flags"):
                    break
    except OSError:
        pass
    return hashlib.sha1(txt.encode()).hexdigest()[:10]


def build_fast() -> str:
    """Timing build of the same source (``-O3 -march=native -ffp-contract=fast``) for bench.py's CPU arm, compiled
    on the host that runs it into ``_build/fast_<cpu key>/``.  Parity tests never load it."""
    out = os.path.join(_BUILD, "fast_" + _host_key())
    path = os.path.join(out, "libzb_oracle_fast.so")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "zb_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-s", "fast", "FASTOUT=" + out], check=True)
    return path


_libs: dict[str, ctypes.CDLL] = {}


def lib(count: bool = False, fast: bool = False) -> ctypes.CDLL:
    key = "cnt" if count else ("fast" if fast else "std")
    if key not in _libs:
        if fast:
            path = build_fast()
        else:
            path = os.path.join(_BUILD, "libzb_oracle_cnt.so" if count else "libzb_oracle.so")
            if not os.path.exists(path):
                build()
        _libs[key] = ctypes.CDLL(path)
    return _libs[key]


class Oracle:
    """Batched CPU oracle in one precision (``np.float32`` — JAX's dtype — or ``np.float64``)."""

    def __init__(self, model, opts, dtype=np.float32, count: bool = False, threads: int = 1, fast: bool = False):
        self.model, self.opts = model, opts
        self.dtype = np.dtype(dtype)
        self.suf = "_f64" if self.dtype == np.float64 else "_f32"
        self.lib = lib(count, fast)  # fast: the -O3 -march=native timing build (bench.py only)
        self.count = count
        self.threads = threads
        self.Data = ZboDataF64 if self.dtype == np.float64 else ZboDataF32
        assert getattr(self.lib, "zbo_sizeof_data" + self.suf)() == ctypes.sizeof(self.Data)

    # ---- helpers -------------------------------------------------------------------------
    def _arr(self, a):
        return None if a is None else np.ascontiguousarray(a, dtype=self.dtype)

    def _params(self, params):
        keep, s = [], _ZboParams()
        if params:
            for k in _PARAM_FIELDS:
                if params.get(k) is not None:
                    a = np.ascontiguousarray(params[k], dtype=np.int32) if k == "latency_steps" else self._arr(params[k])
                    keep.append(a)
                    setattr(s, k, a.ctypes.data)
        return s, keep

    def new_state(self, n):
        st = {k: np.zeros((n,) + STATE_SHAPES[k], dtype=self.dtype) for k in _STATE_FIELDS}
        return st

    def new_out(self, n):
        out = {k: np.zeros((n,) + OUT_SHAPES[k], dtype=self.dtype) for k in _OUT_REAL}
        out.update({k: np.zeros((n,), dtype=np.int32) for k in _OUT_INT})
        out["solver_trace"] = np.zeros((n, NTRACE), dtype=self.dtype)
        return out

    def _state(self, st):
        s = _ZboState()
        for k in _STATE_FIELDS:
            a = st[k]
            assert a.dtype == self.dtype and a.flags.c_contiguous, k
            setattr(s, k, a.ctypes.data)
        return s

    def _out(self, out):
        s = _ZboOut()
        if out is not None:
            for k in _OUT_REAL + _OUT_INT + ["solver_trace"]:
                if out.get(k) is not None:
                    setattr(s, k, out[k].ctypes.data)
        return s

    # ---- single-environment, full intermediates ------------------------------------------
    def forward_one(self, qpos, qvel, ctrl=None, warmstart=None, params=None, env=0):
        d = self.Data()
        ps, keep = self._params(params)
        q, v = self._arr(qpos), self._arr(qvel)
        c, w = self._arr(ctrl), self._arr(warmstart)
        fn = getattr(self.lib, "zbo_forward_one" + self.suf)
        rc = fn(ctypes.byref(self.model), ctypes.byref(self.opts), ctypes.byref(ps), env, q.ctypes.data_as(ctypes.c_void_p),
                v.ctypes.data_as(ctypes.c_void_p), None if c is None else c.ctypes.data_as(ctypes.c_void_p),
                None if w is None else w.ctypes.data_as(ctypes.c_void_p), ctypes.byref(d))
        assert rc == 0
        return d

    def step_one(self, d, params=None, env=0):
        """mjx.step on a ZboData holding qpos/qvel/ctrl/qacc_warmstart (in place)."""
        ps, keep = self._params(params)
        fn = getattr(self.lib, "zbo_step_one" + self.suf)
        rc = fn(ctypes.byref(self.model), ctypes.byref(self.opts), ctypes.byref(ps), env, ctypes.byref(d))
        assert rc == 0
        return d

    # ---- batched engine entry points (same semantics as include/zbphys.h) -----------------
    def step(self, state, action, n_substeps, params=None, push_rand=None, push_per_substep=False, out=None):
        n = state["qpos"].shape[0]
        ps, keep = self._params(params)
        a, pr = self._arr(action), self._arr(push_rand)
        fn = getattr(self.lib, "zbo_step" + self.suf)
        rc = fn(ctypes.byref(self.model), ctypes.byref(self.opts), ctypes.byref(ps), ctypes.byref(self._state(state)),
                a.ctypes.data_as(ctypes.c_void_p), None if pr is None else pr.ctypes.data_as(ctypes.c_void_p),
                int(bool(push_per_substep)), ctypes.byref(self._out(out)), n, int(n_substeps), self.threads)
        assert rc == 0, rc
        return out

    def step_follow(self, state, action, trace, niter, params=None, out=None, force_start=None):
        """One physics step whose solver takes the step sizes of ``trace`` [n, 128] (another implementation's ZbOut.solver_trace)
        and runs ``niter`` [n] iterations (zbo_step_follow).  Returns (out, start_costs [n, 2])."""
        n = state["qpos"].shape[0]
        ps, keep = self._params(params)
        a, tr = self._arr(action), self._arr(trace)
        ni = np.ascontiguousarray(niter, dtype=np.int32)
        fs = None if force_start is None else np.ascontiguousarray(force_start, dtype=np.int32)
        costs = np.zeros((n, 2), dtype=self.dtype)
        fn = getattr(self.lib, "zbo_step_follow" + self.suf)
        P = lambda x: None if x is None else x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        rc = fn(ctypes.byref(self.model), ctypes.byref(self.opts), ctypes.byref(ps), ctypes.byref(self._state(state)), P(a),
                ctypes.byref(self._out(out)), n, self.threads, P(tr), P(ni), P(fs), P(costs))
        assert rc == 0, rc
        return out, costs

    def ls_points(self, state, action, alphas, params=None, force_start=None, search=None):
        """First-line-search derivatives at the step sizes ``alphas`` [n, K] (zbo_ls_points): dict(d0, d1, scale, margin [n, K],
        start_costs [n, 2]).  The state is not modified."""
        n = state["qpos"].shape[0]
        ps, keep = self._params(params)
        a, al = self._arr(action), self._arr(alphas)
        K = al.shape[1]
        fs = None if force_start is None else np.ascontiguousarray(force_start, dtype=np.int32)
        r = {k: np.zeros((n, K), dtype=self.dtype) for k in ("d0", "d1", "scale", "margin", "cost")}
        r["start_costs"] = np.zeros((n, 2), dtype=self.dtype)
        r["search"] = np.zeros((n, NV), dtype=self.dtype)
        r["info"] = np.zeros((n, 4), dtype=self.dtype)
        sin = self._arr(search)
        fn = getattr(self.lib, "zbo_ls_points" + self.suf)
        P = lambda x: None if x is None else x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        rc = fn(ctypes.byref(self.model), ctypes.byref(self.opts), ctypes.byref(ps), ctypes.byref(self._state(state)), P(a), n, self.threads,
                P(fs), P(al), K, P(r["d0"]), P(r["d1"]), P(r["scale"]), P(r["margin"]), P(r["cost"]), P(r["start_costs"]), P(sin), P(r["search"]), P(r["info"]))
        assert rc == 0, rc
        return r

    def forward(self, state, params=None, out=None):
        n = state["qpos"].shape[0]
        ps, keep = self._params(params)
        fn = getattr(self.lib, "zbo_forward" + self.suf)
        rc = fn(ctypes.byref(self.model), ctypes.byref(self.opts), ctypes.byref(ps), ctypes.byref(self._state(state)),
                ctypes.byref(self._out(out)), n, self.threads)
        assert rc == 0, rc
        return out

    def reset(self, state, reset_qpos, reset_qvel, reset_t_rem=None, done=None, params=None, out=None):
        n = state["qpos"].shape[0]
        ps, keep = self._params(params)
        rq, rv, rt = self._arr(reset_qpos), self._arr(reset_qvel), self._arr(reset_t_rem)
        dn = None if done is None else np.ascontiguousarray(done, dtype=np.uint8)
        fn = getattr(self.lib, "zbo_reset" + self.suf)
        rc = fn(ctypes.byref(self.model), ctypes.byref(self.opts), ctypes.byref(ps), ctypes.byref(self._state(state)),
                None if dn is None else dn.ctypes.data_as(ctypes.c_void_p), rq.ctypes.data_as(ctypes.c_void_p),
                rv.ctypes.data_as(ctypes.c_void_p), None if rt is None else rt.ctypes.data_as(ctypes.c_void_p),
                ctypes.byref(self._out(out)), n, self.threads)
        assert rc == 0, rc
        return out

    def observe(self, state, out, imu_x, imu_lag, command=None, reinit=None, reset_lag=None, noise_quat=None, noise_acc=None,
                critic=True):
        """ZbObs epilogue on the arrays a step / reset produced: returns (actor_obs [n,50], critic_obs [n,484] or None);
        imu_x / imu_lag (arrays of this oracle's dtype) are updated in place."""
        n = state["qpos"].shape[0]
        assert imu_x.dtype == self.dtype and imu_lag.dtype == self.dtype and imu_x.flags.c_contiguous
        actor = np.zeros((n, 50), dtype=self.dtype)
        crit = np.zeros((n, 484), dtype=self.dtype) if critic else None
        keep = [self._arr(a) for a in (command, reset_lag, noise_quat, noise_acc)]
        rn = None if reinit is None else np.ascontiguousarray(reinit, dtype=np.uint8)
        P = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        fn = getattr(self.lib, "zbo_observe" + self.suf)
        rc = fn(n, P(state["qpos"]), P(state["qvel"]), P(out["cinert"]), P(out["cvel"]), P(out["sensordata"]), P(out["actuator_force"]),
                P(keep[0]), P(imu_x), P(imu_lag), P(rn), P(keep[1]), P(keep[2]), P(keep[3]), P(actor), P(crit))
        assert rc == 0, rc
        return actor, crit

    def reward_inputs(self, state, out, command=None):
        """ZbObs.feet_pos_obs [n,6] / reward_terms [n,6] from the arrays a step / reset produced."""
        n = state["qpos"].shape[0]
        feet, terms = np.zeros((n, 6), dtype=self.dtype), np.zeros((n, 6), dtype=self.dtype)
        cmd = self._arr(command)
        P = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        rc = getattr(self.lib, "zbo_reward_inputs" + self.suf)(
            n, P(state["qpos"]), P(out["xpos"]), P(out["xquat"]), P(out["site_xpos"]), P(out["sensordata"]), P(cmd),
            int(self.model.site_body[1]), int(self.model.site_body[2]), P(feet), P(terms))
        assert rc == 0, rc
        return feet, terms

    def set_variant(self, v: int) -> None:
        """bit 0: contact rows applied matrix-free like the CUDA kernel (diagnostics)."""
        getattr(self.lib, "zbo_set_variant" + self.suf)(int(v))

    def lsstats(self, reset: bool = True):
        """Line-search statistics collected with set_variant(4) (single-threaded runs), see zb_oracle.c."""
        buf = (ctypes.c_double * 8)()
        getattr(self.lib, "zbo_lsstats_get" + self.suf)(buf, int(reset))
        return list(buf)

    # ---- flop counters (count build only) ------------------------------------------------
    def flops_reset(self):
        self.lib.zbo_flops_reset()

    def flops(self) -> dict[str, float]:
        buf = (ctypes.c_double * len(STAGES))()
        self.lib.zbo_flops_get(buf)
        return dict(zip(STAGES, list(buf)))


def feet_airtime(left_contact, right_contact, done, carry, ctrl_dt=0.02, touchdown_penalty=0.3):
    """Restatement of FeetAirtimeReward.get_reward_stateful (/root/reference/train.py:503-546) for a batch of environments:
    contacts / done [T, N] booleans, carry [N, 2] float32 -> (reward [T, N], new carry [N, 2]), raw (before ``scale``),
    stand_still_threshold = None (train.py:1560-1565).  fp32 accumulation like jax's default dtype."""
    lc, rc, dn = (np.asarray(a, dtype=bool) for a in (left_contact, right_contact, done))
    T, n = lc.shape
    time = np.asarray(carry, dtype=np.float32).copy()
    air = np.zeros((T, n, 2), np.float32)
    dt, pen = np.float32(ctrl_dt), np.float32(touchdown_penalty)
    for t in range(T):
        cod = np.stack([lc[t] | dn[t], rc[t] | dn[t]], axis=-1)
        time = np.where(cod, np.float32(0), time + dt).astype(np.float32)
        air[t] = time
    c = np.stack([lc, rc], axis=-1)
    prev = np.concatenate([np.zeros((1, n, 2), bool), c[:-1]], axis=0)
    td = (c & ~prev).astype(np.float32)
    shifted = np.roll(air, 1, axis=0)
    reward = ((shifted - pen) * td).astype(np.float32)
    return (reward[..., 0] + reward[..., 1]).astype(np.float32), time
