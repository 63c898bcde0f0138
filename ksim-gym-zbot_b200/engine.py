"""Host-side mirror of ksim's ``PhysicsEngine`` for the CUDA Z-Bot physics step.

Reference interface (ksim 0.1.99 ``ksim/engine.py`` — un-vendored, /root/reference/uv.lock:1872-1873;
call sites /root/reference/train.py:1340-1344 actuators, :1457-1476 events/resets, :1439-1455 randomizers):

    PhysicsEngine.reset(physics_model, curriculum_level, rng) -> PhysicsState
    PhysicsEngine.step(action, physics_model, physics_state, curriculum_level, rng) -> PhysicsState

Here ``physics_model`` is the batch of per-environment model overrides the ksim randomizers produce
(``EnvParams``), the state containers are torch CUDA tensors and every call goes through the C ABI of
``include/zbphys.h`` (``libzbphys.so``).  There is no CPU path: importing works anywhere, but constructing an
engine without a CUDA device or without the compiled library raises.
"""

from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import model as M

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "libzbphys.so")

_PARAM_FIELDS = ["body_mass", "body_inertia", "dof_armature", "dof_damping", "dof_frictionloss", "qpos0",
                 "floor_friction", "imu_site_pos", "imu_site_quat", "prev_action", "latency_steps"]
_STATE_FIELDS = ["qpos", "qvel", "qacc_warmstart", "act_pos", "act_vel", "act_torque", "push_t_rem", "time"]
_OUT_F32 = ["xpos", "xquat", "cinert", "cvel", "qacc", "qfrc_constraint", "actuator_force", "ctrl", "sensordata",
            "site_xpos", "site_xmat", "contact_dist", "contact_pos", "efc_force"]
_OUT_I32 = ["contact_mask", "solver_niter", "flags"]

STATE_SHAPES = {"qpos": (M.NQ,), "qvel": (M.NV,), "qacc_warmstart": (M.NV,), "act_pos": (M.NU,), "act_vel": (M.NU,),
                "act_torque": (M.NU,), "push_t_rem": (), "time": ()}
OUT_SHAPES = {
    "xpos": (M.NBODY, 3), "xquat": (M.NBODY, 4), "cinert": (M.NBODY, 10), "cvel": (M.NBODY, 6), "qacc": (M.NV,),
    "qfrc_constraint": (M.NV,), "actuator_force": (M.NU,), "ctrl": (M.NU,), "sensordata": (M.NSENSORDATA,),
    "site_xpos": (M.NSITE, 3), "site_xmat": (M.NSITE, 9), "contact_dist": (M.MAXCON,), "contact_pos": (M.MAXCON, 3),
    "efc_force": (M.NEFC,), "contact_mask": (), "solver_niter": (), "flags": (),
}
PARAM_SHAPES = {"body_mass": (M.NBODY,), "body_inertia": (M.NBODY, 3), "dof_armature": (M.NV,), "dof_damping": (M.NV,),
                "dof_frictionloss": (M.NV,), "qpos0": (M.NQ,), "floor_friction": (), "imu_site_pos": (3,),
                "imu_site_quat": (4,), "prev_action": (M.NU,), "latency_steps": ()}


def _ptr_struct(name, names):
    return type(name, (ctypes.Structure,), {"_fields_": [(n, ctypes.c_void_p) for n in names]})


ZbEnvParams = _ptr_struct("ZbEnvParams", _PARAM_FIELDS)
ZbState = _ptr_struct("ZbState", _STATE_FIELDS)
ZbOut = _ptr_struct("ZbOut", _OUT_F32 + _OUT_I32 + ["solver_trace"])
_OBS_FIELDS = ["command", "imu_x", "imu_lag", "reset_imu_lag", "noise_imu_quat", "noise_imu_acc", "actor_obs", "critic_obs", "feet_pos_obs",
               "reward_terms"]
NUM_REWARD_TERMS = 6  # ZB_NREWARD_TERMS
ZbObs = _ptr_struct("ZbObs", _OBS_FIELDS)
NUM_ACTOR_INPUTS, NUM_CRITIC_INPUTS, NUM_COMMANDS = 50, 484, 7  # train.py:53-54, 416-427


class ZbAutoResetArgs(ctypes.Structure):
    _fields_ = [("reset_qpos", ctypes.c_void_p), ("reset_qvel", ctypes.c_void_p), ("reset_t_rem", ctypes.c_void_p),
                ("done_out", ctypes.c_void_p), ("z_lo", ctypes.c_float), ("z_hi", ctypes.c_float), ("max_tilt_rad", ctypes.c_float),
                ("max_episode_sec", ctypes.c_float), ("term_qpos", ctypes.c_void_p), ("term_qvel", ctypes.c_void_p),
                ("term_xpos", ctypes.c_void_p), ("term_xquat", ctypes.c_void_p)]


MAX_EPISODE_SEC = 80.0  # ksim.EpisodeLengthTermination(max_length_sec=80), train.py:1592


ZB_OK, ZB_EINVAL, ZB_ENODEV, ZB_ECUDA, ZB_EUNSUP = 0, -1, -2, -3, -4
EXPORTS = ["zb_engine_create", "zb_engine_destroy", "zb_engine_set_opts", "zb_step_f32", "zb_forward_f32", "zb_reset_f32",
           "zb_abi_version", "zb_sizeof_model", "zb_sizeof_opts", "zb_last_error", "zb_launch_count", "zb_bench_ffma", "zb_terminations_f32", "zb_step_autoreset_f32",
           "zb_step_obs_f32", "zb_reset_obs_f32", "zb_feet_airtime_f32", "zb_cta_shape", "zb_actor_f32", "zb_sizeof_actor"]


def _term_ptrs(terminal):
    """(qpos, qvel, xpos, xquat) tensors or None -> the four ZbAutoResetArgs.term_* pointers."""
    if terminal is None:
        return (None, None, None, None)
    out = []
    for t in terminal:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        out.append(None if t is None else t.data_ptr())
    return tuple(out)


class ZbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"zbphys error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: str | None = None) -> ctypes.CDLL:
    """dlopen ``libzbphys.so`` (built in-tree by ``__graft_entry__.build()``).  Fails loudly if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("ZBPHYS_LIB") or LIB_PATH  # ZBPHYS_LIB: tuning variants of the same library
    if not os.path.exists(p):
        raise RuntimeError(f"{p} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                           "There is no CPU fallback for the physics step.")
    lib = ctypes.CDLL(p)
    lib.zb_last_error.restype = ctypes.c_char_p
    lib.zb_launch_count.restype = ctypes.c_longlong
    lib.zb_engine_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.zb_engine_destroy.argtypes = [ctypes.c_void_p]
    lib.zb_engine_set_opts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.zb_step_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.zb_step_autoreset_f32.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + \
        [ctypes.c_void_p] * 4 + [ctypes.c_float] * 4
    lib.zb_step_obs_f32.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.zb_reset_obs_f32.argtypes = [ctypes.c_void_p] * 10 + [ctypes.c_int]
    LL = ctypes.c_longlong
    lib.zb_feet_airtime_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, LL, LL, ctypes.c_void_p, LL, LL, ctypes.c_void_p,
                                        ctypes.c_void_p, LL, LL, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float]
    lib.zb_forward_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.zb_reset_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.zb_terminations_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float]
    lib.zb_bench_ffma.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    assert lib.zb_sizeof_model() == ctypes.sizeof(M.ZbModel), "ZbModel layout mismatch between model.py and zbphys.h"
    assert lib.zb_sizeof_opts() == ctypes.sizeof(M.ZbOpts), "ZbOpts layout mismatch between model.py and zbphys.h"
    if path is None:
        _lib = lib
    return lib


# ---------------------------------------------------------------------------------------------
# state containers (duck-type ksim.types.PhysicsState / mjx.Data for the fields train.py reads)
# ---------------------------------------------------------------------------------------------
@dataclass
class PlannerState:
    """train.py:1103-1108."""

    position: torch.Tensor
    velocity: torch.Tensor
    last_computed_torque: torch.Tensor


@dataclass
class ZbotData:
    """The mjx.Data fields the task reads (train.py:452-461, 476, 681, 854, 882, 1252-1253)."""

    qpos: torch.Tensor
    qvel: torch.Tensor
    qacc_warmstart: torch.Tensor
    time: torch.Tensor
    xpos: torch.Tensor
    xquat: torch.Tensor
    cinert: torch.Tensor
    cvel: torch.Tensor
    qacc: torch.Tensor
    qfrc_constraint: torch.Tensor
    actuator_force: torch.Tensor
    ctrl: torch.Tensor
    sensordata: torch.Tensor
    site_xpos: torch.Tensor
    site_xmat: torch.Tensor
    contact_dist: torch.Tensor
    contact_pos: torch.Tensor
    efc_force: torch.Tensor
    contact_mask: torch.Tensor
    solver_niter: torch.Tensor
    flags: torch.Tensor


@dataclass
class PhysicsState:
    """ksim.types.PhysicsState, flattened to batched torch tensors."""

    data: ZbotData
    most_recent_action: torch.Tensor
    event_states: dict = field(default_factory=dict)  # {"push_t_rem": [N]}
    actuator_state: PlannerState | None = None
    action_latency: torch.Tensor | None = None


@dataclass
class Observations:
    """Network inputs of the reference produced by the kernel epilogue (``ZbObs`` in include/zbphys.h):
    ``actor_obs`` [N, 50] = run_actor's concatenation (train.py:1624-1648), ``critic_obs`` [N, 484] = run_critic's
    (train.py:1650-1681); ``imu_x`` / ``imu_lag`` are the carry of ImuOrientationObservation (train.py:846-848)."""

    imu_x: torch.Tensor
    imu_lag: torch.Tensor
    actor_obs: torch.Tensor
    critic_obs: torch.Tensor | None
    lag_range: tuple[float, float] = (0.0, 0.1)  # train.py:1497
    # reward inputs (ZbObs.feet_pos_obs / reward_terms): FeetPositionObservation train.py:430-465 and the per-step values of
    # SimpleSingleFootContactReward, FeetOrientationReward, FeetTooClosePenalty, BaseHeightReward + the two contact flags
    feet_pos_obs: torch.Tensor | None = None
    reward_terms: torch.Tensor | None = None

    def c_struct(self, command=None, reset_imu_lag=None, noise_imu_quat=None, noise_imu_acc=None) -> ZbObs:
        s = ZbObs()
        for k, t in (("command", command), ("imu_x", self.imu_x), ("imu_lag", self.imu_lag), ("reset_imu_lag", reset_imu_lag),
                     ("noise_imu_quat", noise_imu_quat), ("noise_imu_acc", noise_imu_acc), ("actor_obs", self.actor_obs),
                     ("critic_obs", self.critic_obs), ("feet_pos_obs", self.feet_pos_obs), ("reward_terms", self.reward_terms)):
            if t is not None:
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), k
                setattr(s, k, t.data_ptr())
        return s


@dataclass
class PushEventConfig:
    """ksim.PushEvent(...) as configured at train.py:1459-1468."""

    x_linvel: float = 0.1
    y_linvel: float = 0.1
    z_linvel: float = 0.05
    x_angvel: float = 0.0
    y_angvel: float = 0.0
    z_angvel: float = 0.0
    vel_range: tuple[float, float] = (0.05, 0.15)
    interval_range: tuple[float, float] = (2.0, 4.0)


@dataclass
class ResetConfig:
    """train.py:1471-1476: RandomJointPositionReset(JOINT_BIASES, scale=0) + RandomJointVelocityReset()."""

    joint_pos_scale: float = 0.0
    joint_vel_scale: float = 0.01  # ksim default for RandomJointVelocityReset is recalled, see PARITY_RISKS.md


class EnvParams:
    """Per-environment model overrides == what ksim's randomizers put into the batched mjx.Model."""

    def __init__(self, n_envs: int, device, **arrays):
        self.n_envs = n_envs
        self.tensors: dict[str, torch.Tensor] = {}
        for k, v in arrays.items():
            if k not in PARAM_SHAPES:
                raise KeyError(k)
            dt_ = torch.int32 if k == "latency_steps" else torch.float32
            t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v, dtype=dt_).to(device).contiguous()
            if tuple(t.shape) != (n_envs,) + PARAM_SHAPES[k]:
                raise ValueError(f"{k}: expected shape {(n_envs,) + PARAM_SHAPES[k]}, got {tuple(t.shape)}")
            self.tensors[k] = t

    def c_struct(self) -> ZbEnvParams:
        s = ZbEnvParams()
        for k, t in self.tensors.items():
            setattr(s, k, t.data_ptr())
        return s

    def slice(self, lo: int, hi: int) -> "EnvParams":
        out = EnvParams.__new__(EnvParams)
        out.n_envs = hi - lo
        out.tensors = {k: t[lo:hi].contiguous() for k, t in self.tensors.items()}
        return out


class ZbotCudaEngine:
    """Drop-in for ksim ``MjxEngine`` on the Z-Bot: same ``reset`` / ``step`` contract, CUDA kernels underneath.

    ``phys_steps_per_ctrl_steps = round(ctrl_dt / dt)`` physics substeps are fused into ONE kernel launch per
    control step (PushEvent -> FeetechActuators -> mjx.step per substep).
    """

    def __init__(self, model: M.ZbModel, n_envs: int, device: int | str | torch.device = 0, dt: float = 0.001,
                 ctrl_dt: float = 0.02, iterations: int = 8, ls_iterations: int = 8, solver: int = M.SOLVER_CG,
                 integrator: int = M.INT_IMPLICITFAST, push: PushEventConfig | None = None,
                 resets: ResetConfig | None = None, opts: M.ZbOpts | None = None, seed: int = 0,
                 outputs: tuple[str, ...] | None = None, action_latency_range: tuple[float, float] = (0.0, 0.0),
                 drop_action_prob: float = 0.0, lib_path: str | None = None):
        if not torch.cuda.is_available():
            raise RuntimeError("ZbotCudaEngine needs a CUDA device: the physics step has no CPU fallback")
        self.lib = load_library(lib_path)  # lib_path: a tuning / mutation variant of libzbphys.so (build.build_variants)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.model = model
        self.n_envs = int(n_envs)
        self.opts = opts if opts is not None else M.make_opts(dt=dt, iterations=iterations, ls_iterations=ls_iterations,
                                                              solver=solver, integrator=integrator)
        self.dt = float(self.opts.dt)
        self.ctrl_dt = float(ctrl_dt)
        self.phys_steps_per_ctrl_steps = max(1, int(round(self.ctrl_dt / self.dt)))
        self.push = push
        self.resets = resets or ResetConfig()
        self.action_latency_range = tuple(action_latency_range)  # seconds, sampled per reset (ksim get_physics_engine)
        self.drop_action_prob = float(drop_action_prob)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self.outputs = tuple(outputs) if outputs is not None else tuple(_OUT_F32 + _OUT_I32)
        handle = ctypes.c_void_p()
        rc = self.lib.zb_engine_create(ctypes.byref(model), ctypes.byref(self.opts), self.device.index or 0, ctypes.byref(handle))
        self._check(rc)
        self._h = handle
        self.trace: torch.Tensor | None = None  # set to a [N, ZB_NTRACE = 160] float32 CUDA tensor to record the solver trace
        self._default_qpos = torch.tensor(M.default_qpos(model), dtype=torch.float32, device=self.device)

    # ---- plumbing ------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != ZB_OK:
            raise ZbError(rc, self.lib.zb_last_error().decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.zb_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def new_state(self, n: int | None = None) -> PhysicsState:
        n = self.n_envs if n is None else n
        z = lambda shape, dt=torch.float32: torch.zeros((n,) + shape, dtype=dt, device=self.device)  # noqa: E731
        data = ZbotData(
            qpos=z((M.NQ,)), qvel=z((M.NV,)), qacc_warmstart=z((M.NV,)), time=z(()),
            **{k: z(OUT_SHAPES[k]) for k in _OUT_F32}, **{k: z((), torch.int32) for k in _OUT_I32},
        )
        data.qpos[:, 3] = 1.0
        return PhysicsState(
            data=data, most_recent_action=z((M.NU,)), event_states={"push_t_rem": z(())},
            actuator_state=PlannerState(z((M.NU,)), z((M.NU,)), z((M.NU,))), action_latency=z(()),
        )

    def new_observations(self, n: int | None = None, critic: bool = True, lag_range: tuple[float, float] = (0.0, 0.1),
                         rewards: bool = False) -> Observations:
        """Buffers + carry of the observation epilogue (ImuOrientationObservation.initial_carry, train.py:846-848)."""
        n = self.n_envs if n is None else n
        z = lambda *shape: torch.zeros((n,) + shape, dtype=torch.float32, device=self.device)  # noqa: E731
        return Observations(imu_x=z(4), imu_lag=self.sample_imu_lag(n, lag_range), actor_obs=z(NUM_ACTOR_INPUTS),
                            critic_obs=z(NUM_CRITIC_INPUTS) if critic else None, lag_range=tuple(lag_range),
                            feet_pos_obs=z(6) if rewards else None, reward_terms=z(NUM_REWARD_TERMS) if rewards else None)

    def sample_imu_lag(self, n: int, lag_range: tuple[float, float]) -> torch.Tensor:
        lo, hi = lag_range
        return (lo + (hi - lo) * torch.rand((n,), generator=self.gen, device=self.device)).contiguous()

    def _c_state(self, ps: PhysicsState) -> ZbState:
        s = ZbState()
        d, a = ps.data, ps.actuator_state
        for name, t in (("qpos", d.qpos), ("qvel", d.qvel), ("qacc_warmstart", d.qacc_warmstart), ("act_pos", a.position),
                        ("act_vel", a.velocity), ("act_torque", a.last_computed_torque),
                        ("push_t_rem", ps.event_states["push_t_rem"]), ("time", d.time)):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), name
            setattr(s, name, t.data_ptr())
        return s

    def _c_out(self, ps: PhysicsState) -> ZbOut:
        o = ZbOut()
        if self.trace is not None:
            o.solver_trace = self.trace.data_ptr()
        for k in self.outputs:
            t = getattr(ps.data, k)
            assert t.is_cuda and t.is_contiguous(), k
            setattr(o, k, t.data_ptr())
        return o

    @staticmethod
    def _c_params(params: EnvParams | None):
        return ctypes.byref(params.c_struct()) if params is not None else None

    # ---- host-side sampling (ksim resets / events; stays outside the kernel on purpose: the RNG stream is
    #      the caller's, the kernel only consumes the pre-sampled numbers) ----------------------------------
    def sample_reset(self, n: int, physics_model: EnvParams | None = None):
        """ksim resets of train.py:1471-1476 -> (qpos [n,27], qvel [n,26], push_t_rem [n])."""
        qpos = self._default_qpos.expand(n, M.NQ).clone()
        if self.resets.joint_pos_scale > 0:
            qpos[:, 7:] += (torch.rand((n, M.NU), generator=self.gen, device=self.device) * 2 - 1) * self.resets.joint_pos_scale
        qvel = torch.zeros((n, M.NV), device=self.device)
        if self.resets.joint_vel_scale > 0:
            qvel[:, 6:] += (torch.rand((n, M.NU), generator=self.gen, device=self.device) * 2 - 1) * self.resets.joint_vel_scale
        if self.push is not None:
            lo, hi = self.push.interval_range
            t_rem = lo + (hi - lo) * torch.rand((n,), generator=self.gen, device=self.device)
        else:
            t_rem = torch.full((n,), 1e30, device=self.device)
        return qpos, qvel, t_rem

    def sample_push(self, n: int, curriculum_level: float = 1.0) -> torch.Tensor:
        """Pre-sampled PushEvent numbers for one control step: [n, 7] = (delta qvel[0:6], next interval)."""
        p = self.push
        scale = torch.tensor([p.x_linvel, p.y_linvel, p.z_linvel, p.x_angvel, p.y_angvel, p.z_angvel], device=self.device)
        u = torch.rand((n, 8), generator=self.gen, device=self.device)
        direction = (u[:, :6] * 2 - 1) * scale
        mag = p.vel_range[0] + (p.vel_range[1] - p.vel_range[0]) * u[:, 6:7]
        dq = direction * (mag / max(p.vel_range[1], 1e-12)) * curriculum_level
        nxt = p.interval_range[0] + (p.interval_range[1] - p.interval_range[0]) * u[:, 7:8]
        return torch.cat([dq, nxt], dim=1).contiguous()

    # ---- PhysicsEngine interface -----------------------------------------------------------
    def reset(self, physics_model: EnvParams | None = None, curriculum_level: float = 1.0, rng=None,
              physics_state: PhysicsState | None = None, done: torch.Tensor | None = None,
              reset_qpos: torch.Tensor | None = None, reset_qvel: torch.Tensor | None = None,
              reset_t_rem: torch.Tensor | None = None, obs: Observations | None = None,
              command: torch.Tensor | None = None, reset_imu_lag: torch.Tensor | None = None,
              obs_noise: tuple | None = None) -> PhysicsState:
        """MjxEngine.reset.  With ``physics_state`` and ``done`` given, only the finished environments are reset
        (what ksim's rollout does with ``jnp.where(done, reset_state, state)``)."""
        ps = physics_state if physics_state is not None else self.new_state()
        n = ps.data.qpos.shape[0]
        if reset_qpos is None:
            reset_qpos, rv_, t = self.sample_reset(n, physics_model)
            reset_qvel = rv_ if reset_qvel is None else reset_qvel
            reset_t_rem = t if reset_t_rem is None else reset_t_rem
        if reset_qvel is None:
            reset_qvel = torch.zeros((n, M.NV), device=self.device)
        if reset_t_rem is None:  # a PushEvent timer of 0 would fire in the very first substep: sample it (ksim: U(interval_range)) or park it
            if self.push is not None:
                lo, hi = self.push.interval_range
                reset_t_rem = lo + (hi - lo) * torch.rand((n,), generator=self.gen, device=self.device)
            else:
                reset_t_rem = torch.full((n,), 1e30, device=self.device)
        reset_qpos = reset_qpos.to(self.device, torch.float32).contiguous()
        reset_qvel = reset_qvel.to(self.device, torch.float32).contiguous()
        reset_t_rem = reset_t_rem.to(self.device, torch.float32).contiguous()
        dn = None
        if done is not None:
            dn = done.to(self.device, torch.uint8).contiguous()
        st, out = self._c_state(ps), self._c_out(ps)
        args = (self._h, self._stream(), self._c_params(physics_model), ctypes.byref(st),
                None if dn is None else ctypes.c_void_p(dn.data_ptr()), ctypes.c_void_p(reset_qpos.data_ptr()),
                ctypes.c_void_p(reset_qvel.data_ptr()),
                None if reset_t_rem is None else ctypes.c_void_p(reset_t_rem.data_ptr()), ctypes.byref(out))
        if obs is None:
            rc = self.lib.zb_reset_f32(*args, n)
        else:  # first observation of the episode; the carry of the reset environments restarts (train.py:846-848)
            if reset_imu_lag is None:
                reset_imu_lag = self.sample_imu_lag(n, obs.lag_range)
            nq, na = obs_noise if obs_noise is not None else (None, None)
            ob = obs.c_struct(command, reset_imu_lag, nq, na)
            rc = self.lib.zb_reset_obs_f32(*args, ctypes.byref(ob), n)
        self._check(rc)
        lo_, hi_ = self.action_latency_range
        new_lat = lo_ + (hi_ - lo_) * torch.rand((n,), generator=self.gen, device=self.device) if hi_ > 0 else None
        if done is None:
            ps.most_recent_action.copy_(ps.data.qpos[:, 7:])  # get_default_action train.py:1282-1283
            if new_lat is not None:
                ps.action_latency = new_lat
        else:
            torch.where(dn.bool()[:, None], ps.data.qpos[:, 7:], ps.most_recent_action, out=ps.most_recent_action)
            if new_lat is not None:
                ps.action_latency = torch.where(dn.bool(), new_lat, ps.action_latency)
        return ps

    def step(self, action: torch.Tensor, physics_model: EnvParams | None, physics_state: PhysicsState,
             curriculum_level: float = 1.0, rng=None, push_rand: torch.Tensor | None = None,
             n_substeps: int | None = None, auto_reset: tuple | None = None, done_out: torch.Tensor | None = None,
             z_lo: float = 0.05, z_hi: float = 0.5, max_tilt: float = math.radians(60), obs: Observations | None = None,
             command: torch.Tensor | None = None, reset_imu_lag: torch.Tensor | None = None,
             obs_noise: tuple | None = None, max_episode_sec: float = MAX_EPISODE_SEC, terminal: tuple | None = None) -> PhysicsState:
        """MjxEngine.step: one control step = ``phys_steps_per_ctrl_steps`` fused physics substeps (in place).

        ``terminal=(qpos, qvel, xpos, xquat)`` (with ``auto_reset`` and ``obs``; each may be None): receives the Trajectory record of
        the step — the state BEFORE the done -> reset select (ZbAutoResetArgs.term_*).

        ``auto_reset=(reset_qpos, reset_qvel, reset_t_rem)`` folds the rollout's termination check + partial reset
        (train.py:1588-1593, 1471-1476) into the same launch; ``done_out`` (uint8 [N]) receives the mask.
        ``obs`` adds the observation epilogue: the next step's actor / critic input vectors (train.py:1624-1681) are
        written by the same launch (``command`` [N, 7], ``obs_noise`` = (imu_quat [N, 4], imu_acc [N, 3]) pre-scaled)."""
        ps = physics_state
        n = ps.data.qpos.shape[0]
        if not (action.is_cuda and action.dtype == torch.float32 and action.is_contiguous()):
            action = action.to(self.device, torch.float32).contiguous()
        if tuple(action.shape) != (n, M.NU):
            raise ValueError(f"action must be [{n}, {M.NU}]")
        # MjxEngine.step: drop-action bernoulli and action latency (both off in the reference's config)
        if self.drop_action_prob > 0:
            drop = torch.rand((n, 1), generator=self.gen, device=self.device) < self.drop_action_prob
            action = torch.where(drop, ps.most_recent_action, action).contiguous()
        if ps.action_latency is not None and self.action_latency_range[1] > 0:
            lat = EnvParams.__new__(EnvParams)
            lat.n_envs = n
            lat.tensors = dict(physics_model.tensors) if physics_model is not None else {}
            lat.tensors["prev_action"] = ps.most_recent_action  # (the launch reads it before the stream-ordered update below)
            lat.tensors["latency_steps"] = torch.round(ps.action_latency / self.dt).to(torch.int32).contiguous()
            physics_model = lat
        per_substep = 0
        if push_rand is None and self.push is not None:
            push_rand = self.sample_push(n, curriculum_level)
        if push_rand is not None:
            push_rand = push_rand.to(self.device, torch.float32).contiguous()
            per_substep = 1 if push_rand.dim() == 3 else 0
        S = self.phys_steps_per_ctrl_steps if n_substeps is None else int(n_substeps)
        st, out = self._c_state(ps), self._c_out(ps)
        pr = None if push_rand is None else ctypes.c_void_p(push_rand.data_ptr())
        if obs is not None:
            ar = None
            if auto_reset is not None:
                rq, rv, rt = auto_reset
                for t in (rq, rv) + ((rt,) if rt is not None else ()):
                    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                ar = ZbAutoResetArgs(rq.data_ptr(), rv.data_ptr(), None if rt is None else rt.data_ptr(),
                                     None if done_out is None else done_out.data_ptr(), z_lo, z_hi, max_tilt, max_episode_sec,
                                     *_term_ptrs(terminal))
                if reset_imu_lag is None:
                    reset_imu_lag = self.sample_imu_lag(n, obs.lag_range)
            nq, na = obs_noise if obs_noise is not None else (None, None)
            ob = obs.c_struct(command, reset_imu_lag, nq, na)
            rc = self.lib.zb_step_obs_f32(self._h, self._stream(), self._c_params(physics_model), ctypes.byref(st),
                                          ctypes.c_void_p(action.data_ptr()), pr, per_substep, ctypes.byref(out), ctypes.byref(ob), n, S,
                                          None if ar is None else ctypes.byref(ar))
        elif terminal is not None:
            raise ValueError("terminal= (Trajectory record) needs the zb_step_obs_f32 path: pass obs=")
        elif auto_reset is None:
            rc = self.lib.zb_step_f32(self._h, self._stream(), self._c_params(physics_model), ctypes.byref(st),
                                      ctypes.c_void_p(action.data_ptr()), pr, per_substep, ctypes.byref(out), n, S)
        else:
            rq, rv, rt = auto_reset
            for t in (rq, rv) + ((rt,) if rt is not None else ()):
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
            rc = self.lib.zb_step_autoreset_f32(
                self._h, self._stream(), self._c_params(physics_model), ctypes.byref(st), ctypes.c_void_p(action.data_ptr()), pr,
                per_substep, ctypes.byref(out), n, S, ctypes.c_void_p(rq.data_ptr()), ctypes.c_void_p(rv.data_ptr()),
                None if rt is None else ctypes.c_void_p(rt.data_ptr()),
                None if done_out is None else ctypes.c_void_p(done_out.data_ptr()), z_lo, z_hi, max_tilt, max_episode_sec)
        self._check(rc)
        ps.most_recent_action.copy_(action)  # engine-owned buffer: never alias the caller's action tensor
        if auto_reset is not None and done_out is not None and self.action_latency_range[1] > 0:
            # reset environments restart from the default action (get_default_action, train.py:1282-1283) with a fresh latency
            dn = done_out.bool()
            torch.where(dn[:, None], ps.data.qpos[:, 7:], ps.most_recent_action, out=ps.most_recent_action)
            lo_, hi_ = self.action_latency_range
            ps.action_latency = torch.where(dn, lo_ + (hi_ - lo_) * torch.rand((n,), generator=self.gen, device=self.device), ps.action_latency)
        return ps

    def plan_step(self, physics_state: PhysicsState, physics_model: EnvParams | None = None, n_substeps: int | None = None,
                  auto_reset: tuple | None = None, done_out: torch.Tensor | None = None, z_lo: float = 0.05, z_hi: float = 0.5,
                  max_tilt: float = math.radians(60), obs: Observations | None = None, command: torch.Tensor | None = None,
                  reset_imu_lag: torch.Tensor | None = None, obs_noise: tuple | None = None, push_per_substep: bool = False,
                  max_episode_sec: float = MAX_EPISODE_SEC, track_action: bool = True, terminal: tuple | None = None):
        """Pre-bind one control step: all argument structs of the C call are built once, the returned function
        ``run(action, push_rand=None)`` only passes the two pointers that change per step and enqueues the launch on the
        current stream (what an XLA custom call does with its operand buffers; ~10x less host time per step than ``step``).
        The tensors bound here must stay alive and in place (refill ``reset_imu_lag`` / the reset arrays / ``command`` in place to
        change them between steps; a ``reset_imu_lag`` of None is sampled ONCE here); action latency / drop-action are not handled
        by a plan."""
        if self.drop_action_prob > 0 or self.action_latency_range[1] > 0:
            raise ValueError("plan_step does not cover action latency / drop-action: use step()")
        ps = physics_state
        n = ps.data.qpos.shape[0]
        S = self.phys_steps_per_ctrl_steps if n_substeps is None else int(n_substeps)
        st, out = self._c_state(ps), self._c_out(ps)
        params = physics_model.c_struct() if physics_model is not None else None
        keep = [ps, physics_model, auto_reset, done_out, obs, command, reset_imu_lag, obs_noise, st, out, params]
        p_params = ctypes.byref(params) if params is not None else None
        p_st, p_out, h, lib, pps = ctypes.byref(st), ctypes.byref(out), self._h, self.lib, int(bool(push_per_substep))
        ar = None
        if auto_reset is not None:
            rq, rv, rt = auto_reset
            ar = ZbAutoResetArgs(rq.data_ptr(), rv.data_ptr(), None if rt is None else rt.data_ptr(),
                                 None if done_out is None else done_out.data_ptr(), z_lo, z_hi, max_tilt, max_episode_sec,
                                 *_term_ptrs(terminal))
            keep += [ar, terminal]
            if terminal is not None and obs is None:
                raise ValueError("terminal= (Trajectory record) needs the zb_step_obs_f32 path: pass obs=")
        if obs is not None:
            if auto_reset is not None and reset_imu_lag is None:
                reset_imu_lag = self.sample_imu_lag(n, obs.lag_range)
                keep.append(reset_imu_lag)
            nq, na = obs_noise if obs_noise is not None else (None, None)
            ob = obs.c_struct(command, reset_imu_lag, nq, na)
            keep.append(ob)
            p_ob, p_ar = ctypes.byref(ob), (ctypes.byref(ar) if ar is not None else None)
            fn = lib.zb_step_obs_f32

            def call(stream, a_ptr, pr_ptr):
                return fn(h, stream, p_params, p_st, a_ptr, pr_ptr, pps, p_out, p_ob, n, S, p_ar)
        elif ar is not None:
            fn = lib.zb_step_autoreset_f32

            def call(stream, a_ptr, pr_ptr):
                return fn(h, stream, p_params, p_st, a_ptr, pr_ptr, pps, p_out, n, S, ar.reset_qpos, ar.reset_qvel, ar.reset_t_rem, ar.done_out,
                          z_lo, z_hi, max_tilt, max_episode_sec)
        else:
            fn = lib.zb_step_f32

            def call(stream, a_ptr, pr_ptr):
                return fn(h, stream, p_params, p_st, a_ptr, pr_ptr, pps, p_out, n, S)
        dev, current_stream, check = self.device, torch.cuda.current_stream, self._check

        def run(action: torch.Tensor, push_rand: torch.Tensor | None = None, _keep=keep):
            rc = call(current_stream(dev).cuda_stream, action.data_ptr(), None if push_rand is None else push_rand.data_ptr())
            if rc != ZB_OK:
                check(rc)
            if track_action:  # (a device-to-device copy of 80 B per environment; track_action=False when the caller keeps the action)
                ps.most_recent_action.copy_(action)

        return run

    def forward(self, physics_model: EnvParams | None, physics_state: PhysicsState) -> PhysicsState:
        """mjx.forward on the current state (ctrl = last computed actuator torque)."""
        ps = physics_state
        st, out = self._c_state(ps), self._c_out(ps)
        rc = self.lib.zb_forward_f32(self._h, self._stream(), self._c_params(physics_model), ctypes.byref(st), ctypes.byref(out),
                                     ps.data.qpos.shape[0])
        self._check(rc)
        return ps

    def terminations(self, physics_state: PhysicsState, done: torch.Tensor | None = None, z_lo: float = 0.05,
                     z_hi: float = 0.5, max_tilt: float = math.radians(60), max_episode_sec: float = MAX_EPISODE_SEC) -> torch.Tensor:
        """BadZTermination | NotUprightTermination | EpisodeLengthTermination (train.py:1588-1593) as one kernel; returns a uint8 [N]
        mask: 1 = failure, 2 = episode length reached (ksim: -1 / +1), 0 = running."""
        n = physics_state.data.qpos.shape[0]
        if done is None:
            done = torch.empty((n,), dtype=torch.uint8, device=self.device)
        rc = self.lib.zb_terminations_f32(self._stream(), ctypes.c_void_p(physics_state.data.qpos.data_ptr()),
                                          ctypes.c_void_p(physics_state.data.flags.data_ptr()), ctypes.c_void_p(physics_state.data.time.data_ptr()),
                                          ctypes.c_void_p(done.data_ptr()), n, z_lo, z_hi, max_tilt, max_episode_sec)
        self._check(rc)
        return done

    def feet_airtime_reward(self, left_contact: torch.Tensor, right_contact: torch.Tensor, done: torch.Tensor, carry: torch.Tensor,
                            reward: torch.Tensor | None = None, ctrl_dt: float | None = None, touchdown_penalty: float = 0.3) -> torch.Tensor:
        """FeetAirtimeReward.get_reward_stateful over one rollout chunk (train.py:503-546, raw: before ``scale``).
        ``left_contact`` / ``right_contact`` [T, N] float (any strides, e.g. ``terms[:, :, 4]`` of stacked ``reward_terms``),
        ``done`` [T, N] uint8, ``carry`` [N, 2] float32 contiguous (updated in place); returns ``reward`` [T, N]."""
        T, n = left_contact.shape
        assert right_contact.shape == (T, n) and done.shape == (T, n) and done.dtype == torch.uint8 and carry.shape == (n, 2) and carry.is_contiguous()
        assert left_contact.stride() == right_contact.stride() and left_contact.dtype == torch.float32 == right_contact.dtype
        if reward is None:
            reward = torch.empty((T, n), dtype=torch.float32, device=self.device)
        rc = self.lib.zb_feet_airtime_f32(self._stream(), ctypes.c_void_p(left_contact.data_ptr()), ctypes.c_void_p(right_contact.data_ptr()),
                                          left_contact.stride(0), left_contact.stride(1), ctypes.c_void_p(done.data_ptr()), done.stride(0),
                                          done.stride(1), ctypes.c_void_p(carry.data_ptr()), ctypes.c_void_p(reward.data_ptr()), reward.stride(0),
                                          reward.stride(1), n, T, self.ctrl_dt if ctrl_dt is None else ctrl_dt, touchdown_penalty)
        self._check(rc)
        return reward

    def launch_count(self) -> int:
        return int(self.lib.zb_launch_count())


def get_physics_engine(model: M.ZbModel, n_envs: int, **kw) -> ZbotCudaEngine:
    """Counterpart of ``ksim.get_physics_engine`` for this path."""
    return ZbotCudaEngine(model, n_envs, **kw)


def terminations(ps: PhysicsState, z_lo: float = 0.05, z_hi: float = 0.5, max_tilt: float = math.radians(60),
                 max_episode_sec: float = MAX_EPISODE_SEC) -> torch.Tensor:
    """BadZTermination + NotUprightTermination + EpisodeLengthTermination of train.py:1588-1593 (used by the rollout harness)."""
    q = ps.data.qpos
    z = q[:, 2]
    w, x, y = q[:, 3], q[:, 4], q[:, 5]
    # z-axis of the base in the world frame: R[2,2] = w^2 - x^2 - y^2 + z^2 = 1 - 2(x^2 + y^2) for unit quats
    up = 1.0 - 2.0 * (x * x + y * y)
    bad = (z < z_lo) | (z > z_hi) | (up < math.cos(max_tilt)) | (ps.data.flags != 0) | ~torch.isfinite(z) | ~torch.isfinite(w)
    if max_episode_sec > 0:
        bad = bad | (ps.data.time > max_episode_sec)
    return bad
