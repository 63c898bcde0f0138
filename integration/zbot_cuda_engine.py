"""ksim-side binding of the CUDA Z-Bot physics step: ``ZbotCudaEngine(ksim.PhysicsEngine)``.

UNEXERCISED: JAX, ksim and the XLA FFI headers are not installable in the image this repository is built in, so this module
cannot be imported there.  What IS checked there (tests/test_integration_cpu.py): it parses, its operand / result tables equal the
``Bind()`` chains of integration/zb_xla_ffi.cc in count, order and dtype, and every ``mjx.Data`` field the reference task reads
exists on ``ZbotData``.  ksim internals used here are recalled from ksim 0.1.99 (un-vendored: /root/reference/uv.lock:1872-1873)
and marked [R].

The reference builds its engine in ``ksim.RLTask.get_engine`` -> ``ksim.get_physics_engine(...)``.  There is no plugin registry, so
the drop-in route is this ``PhysicsEngine`` subclass returned from an overridden ``get_engine`` — the ONLY edit in
/root/reference/train.py:

    class ZbotWalkingTask(ksim.PPOTask[ZbotWalkingTaskConfig]):
        def get_engine(self, physics_model, metadata=None):
            return ZbotCudaEngine.create(self, physics_model, metadata)

``reset`` / ``step`` keep ksim's signatures and return a ``ksim.PhysicsState`` whose ``data`` is a ``ZbotData`` pytree that duck-types
``mjx.Data`` for every field the task reads (train.py:452-461, 476, 681, 854, 882, 1252-1253, 1307).  Each is ONE XLA custom call
under ``jax.vmap`` (registered with ``vmap_method="broadcast_all"``: every operand arrives with the leading environment dimension
the kernel expects) inside ksim's ``jax.lax.scan``.  Random numbers stay JAX's: PushEvent draws, reset states and the IMU lag of
re-initialised environments are sampled here and handed to the kernel as plain arrays.  The termination check and the partial
reset are folded into the step call (``data.done``); ksim's own ``jnp.where(done, reset_state, state)`` then selects between two
identical states for those environments, i.e. stays correct and becomes a no-op.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass

import jax
import jax.numpy as jnp
import ksim
import numpy as np
import xax

NQ, NV, NU, NB, NSITE = 27, 26, 20, 26, 3

# ---- operand / result tables: the single source of truth on the Python side (compared with zb_xla_ffi.cc by the CPU test) --------
# (name, per-environment shape, dtype)
_STATE = (("qpos", (NQ,), "f32"), ("qvel", (NV,), "f32"), ("qacc_warmstart", (NV,), "f32"), ("act_pos", (NU,), "f32"), ("act_vel", (NU,), "f32"),
          ("act_torque", (NU,), "f32"), ("push_t_rem", (), "f32"), ("time", (), "f32"), ("imu_x", (4,), "f32"), ("imu_lag", (), "f32"))
_STEP_INPUTS = (("action", (NU,), "f32"), ("push_rand", (7,), "f32"), ("command", (7,), "f32"), ("reset_qpos", (NQ,), "f32"),
                ("reset_qvel", (NV,), "f32"), ("reset_t_rem", (), "f32"), ("reset_imu_lag", (), "f32"))
_RESET_INPUTS = (("reset_qpos", (NQ,), "f32"), ("reset_qvel", (NV,), "f32"), ("reset_t_rem", (), "f32"), ("reset_imu_lag", (), "f32"),
                 ("command", (7,), "f32"))
_MODEL_FIELDS = (("body_mass", (NB,), "f32"), ("body_inertia", (NB, 3), "f32"), ("dof_armature", (NV,), "f32"), ("dof_damping", (NV,), "f32"),
                 ("dof_frictionloss", (NV,), "f32"), ("qpos0", (NQ,), "f32"), ("floor_friction", (), "f32"), ("imu_site_pos", (3,), "f32"),
                 ("imu_site_quat", (4,), "f32"))
_DATA_OUT = (("xpos", (NB, 3), "f32"), ("xquat", (NB, 4), "f32"), ("cinert", (NB, 10), "f32"), ("cvel", (NB, 6), "f32"), ("qacc", (NV,), "f32"),
             ("actuator_force", (NU,), "f32"), ("ctrl", (NU,), "f32"), ("sensordata", (18,), "f32"), ("site_xpos", (NSITE, 3), "f32"),
             ("site_xmat", (NSITE, 9), "f32"))
_OBS_OUT = (("actor_obs", (50,), "f32"), ("critic_obs", (484,), "f32"), ("feet_pos_obs", (6,), "f32"), ("reward_terms", (6,), "f32"))
# zb_actor (policy inference in the rollout, train.py:946-968, 1737-1763): weights in kernel layout (policy.ActorParams.pack), then the
# per-step arrays; `carry` is aliased with its result
_ACTOR_ATTRS = (("depth", "i32"), ("num_mixtures", "i32"), ("precision", "i32"), ("argmax", "i32"), ("min_std", "f32"), ("max_std", "f32"), ("var_scale", "f32"))
_ACTOR_WEIGHTS = (("in_wt", (50, 128), "f32"), ("in_b", (128,), "f32"), ("gru_wih_t", (5, 128, 3, 128), "f32"), ("gru_whh_t", (5, 128, 3, 128), "f32"),
                  ("gru_b", (5, 3, 128), "f32"), ("gru_bn", (5, 128), "f32"), ("out_wt", (128, 300), "f32"), ("out_b", (300,), "f32"),
                  ("mean_bias", (NU,), "f32"), ("gru_w_tc", (5, 16, 6, 2, 128, 4), "f32"), ("out_w_tc", (16, 3, 2, 128, 4), "f32"))
_ACTOR_INPUTS = (("actor_obs", (50,), "f32"), ("carry", (5, 128), "f32"), ("carry_reset", (), "u8"), ("noise_u", (NU,), "f32"), ("noise_eps", (NU,), "f32"))
_ACTOR_RESULTS = (("carry", (5, 128), "f32"), ("action", (NU,), "f32"))
_STEP_RESULTS = _STATE + _DATA_OUT + (("done", (), "u8"),) + _OBS_OUT
_RESET_RESULTS = _STATE + _DATA_OUT + _OBS_OUT
_STEP_ATTRS = (("engine", "i64"), ("n_substeps", "i32"), ("z_lo", "f32"), ("z_hi", "f32"), ("max_tilt_rad", "f32"), ("max_episode_sec", "f32"))
_RESET_ATTRS = (("engine", "i64"),)
_DT = {"f32": jnp.float32, "u8": jnp.uint8, "i32": jnp.int32}

_lib = ctypes.CDLL("libzb_ffi.so")
for _name in ("zb_step_obs", "zb_reset_obs", "zb_feet_airtime", "zb_actor"):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_lib, _name + "_ffi")), platform="CUDA")


@jax.tree_util.register_dataclass
@dataclass(frozen=True)
class ZbotData:
    """Duck-types ``mjx.Data`` for the fields the task reads (train.py:452-461 site poses, :476 / :681 xpos / xquat, :854 sensordata,
    :882 qpos, :1252-1253 qpos / qvel in the actuators; ksim's built-in observations: qpos, qvel, cinert, cvel, actuator_force,
    sensordata [R]) at the point of the pipeline where ``mjx.step`` leaves them (forward() of the last substep),
    plus what the fused launch produces on top: the termination mask and the network / reward inputs."""

    qpos: jax.Array
    qvel: jax.Array
    qacc_warmstart: jax.Array
    time: jax.Array
    xpos: jax.Array
    xquat: jax.Array
    cinert: jax.Array
    cvel: jax.Array
    qacc: jax.Array
    actuator_force: jax.Array
    ctrl: jax.Array
    sensordata: jax.Array
    site_xpos: jax.Array
    site_xmat: jax.Array  # [3, 9] row-major; ksim.get_site_pose reshapes to 3x3 [R]
    done: jax.Array       # uint8: 0 running, 1 failure, 2 episode length (the state is ALREADY the reset one where non-zero)
    imu_x: jax.Array      # carry of ImuOrientationObservation (train.py:846-848), maintained by the kernel epilogue
    imu_lag: jax.Array
    actor_obs: jax.Array
    critic_obs: jax.Array
    feet_pos_obs: jax.Array
    reward_terms: jax.Array


def _flatten(ps: ksim.PhysicsState) -> list:
    """PhysicsState -> the ten state operands of ``_STATE``, in order."""
    d, a = ps.data, ps.actuator_state  # a: PlannerState, train.py:1103-1108
    return [d.qpos, d.qvel, d.qacc_warmstart, a.position, a.velocity, a.last_computed_torque, ps.event_states["push"], d.time, d.imu_x, d.imu_lag]


def _unflatten(outs, names, most_recent_action, action_latency, planner_cls) -> ksim.PhysicsState:
    """Custom-call results (in table order) -> PhysicsState(data=ZbotData, ...)."""
    r = {n: v for (n, _, _), v in zip(names, outs)}
    data = ZbotData(**{k: r[k] for k in ("qpos", "qvel", "qacc_warmstart", "time", "xpos", "xquat", "cinert", "cvel", "qacc", "actuator_force",
                                         "ctrl", "sensordata", "site_xpos", "site_xmat", "imu_x", "imu_lag", "actor_obs", "critic_obs",
                                         "feet_pos_obs", "reward_terms")},
                    done=r.get("done", jnp.zeros((), jnp.uint8)))
    return ksim.PhysicsState(
        data=data, most_recent_action=most_recent_action, event_states=xax.FrozenDict({"push": r["push_t_rem"]}),
        actuator_state=planner_cls(position=r["act_pos"], velocity=r["act_vel"], last_computed_torque=r["act_torque"]),
        action_latency=action_latency)


def _results(table):
    return [jax.ShapeDtypeStruct(shape, _DT[dt]) for _, shape, dt in table]


class ZbotCudaEngine(ksim.PhysicsEngine):
    """Same ``reset`` / ``step`` signatures as ksim's ``MjxEngine`` [R]; one custom call each."""

    handle: int                    # ZbEngine* from zb_engine_create (held by the task for the life of the process)
    base_fields: dict              # ZbModel values of the per-env fields, used where a randomizer does not override one
    default_qpos: jax.Array
    planner_cls: type              # train.PlannerState
    push: dict                     # PushEvent parameters, train.py:1459-1468
    reset_joint_vel_scale: float
    imu_lag_range: tuple[float, float]
    z_lo: float = 0.05             # train.py:1590
    z_hi: float = 0.5
    max_tilt_rad: float = math.radians(60)  # train.py:1591
    max_episode_sec: float = 80.0  # train.py:1592

    @classmethod
    def create(cls, task, physics_model, metadata=None):
        """What ``ksim.get_physics_engine`` does for MJX [R] + ``zb_engine_create`` on the compiled model.  The MJCF -> ZbModel
        compiler is ksim-gym-zbot_b200/mjcf.py (pure Python, no MuJoCo); `task.get_mujoco_model` has already produced the XML."""
        import ksim_gym_zbot_b200  # noqa: F401
        from ksim_gym_zbot_b200 import mjcf, model as M
        from ksim_gym_zbot_b200.engine import load_library

        import train  # the reference module: PlannerState lives there (train.py:1103-1108)

        c = mjcf.compile_mjcf(open(ksim.get_mujoco_model_path("zbot", name="robot")).read(), servo_params=metadata, floor={"priority": 2})
        opts = M.make_opts(dt=task.config.dt, iterations=task.config.iterations, ls_iterations=task.config.ls_iterations)
        lib, h = load_library(), ctypes.c_void_p()
        assert lib.zb_engine_create(ctypes.byref(c.model), ctypes.byref(opts), jax.devices()[0].id, ctypes.byref(h)) == 0, lib.zb_last_error()
        m = c.model
        base = {k: jnp.asarray(np.ctypeslib.as_array(getattr(m, k)).copy()) for k in ("body_mass", "body_inertia", "dof_armature", "dof_damping",
                                                                                      "dof_frictionloss", "qpos0")}
        base["floor_friction"] = jnp.float32(m.floor_friction[0])
        base["imu_site_pos"] = jnp.asarray(np.ctypeslib.as_array(m.site_pos)[0].copy())
        base["imu_site_quat"] = jnp.asarray(np.ctypeslib.as_array(m.site_quat)[0].copy())
        return cls(
            actuators=task.get_actuators(physics_model, metadata), resets=task.get_resets(physics_model), events=task.get_events(physics_model),
            phys_steps_per_ctrl_steps=round(task.config.ctrl_dt / task.config.dt), drop_action_prob=0.0, action_latency_range=(0.0, 0.0),
            handle=int(h.value), base_fields=base, default_qpos=jnp.asarray(M.default_qpos(m), jnp.float32), planner_cls=train.PlannerState,
            push=dict(scale=jnp.array([0.1, 0.1, 0.05, 0.0, 0.0, 0.0], jnp.float32), vel_range=(0.05, 0.15), interval_range=(2.0, 4.0)),
            reset_joint_vel_scale=0.01, imu_lag_range=(0.0, 0.1))

    # ---- the pieces ksim's resets / events / randomizers compute, as plain arrays -------------------------------------------
    def _model_fields(self, physics_model) -> list:
        """Per-environment randomizer outputs (train.py:1439-1455) in ZbEnvParams order; ``physics_model`` is the mjx.Model-like
        pytree ksim merged them into [R] (field names of mjx.Model), falling back to the compiled model's values."""
        g = lambda name, default: getattr(physics_model, name, default)  # noqa: E731
        b = self.base_fields
        geom_friction = g("geom_friction", None)
        site_pos, site_quat = g("site_pos", None), g("site_quat", None)
        return [g("body_mass", b["body_mass"]), g("body_inertia", b["body_inertia"]), g("dof_armature", b["dof_armature"]),
                g("dof_damping", b["dof_damping"]), g("dof_frictionloss", b["dof_frictionloss"]), g("qpos0", b["qpos0"]),
                b["floor_friction"] if geom_friction is None else geom_friction[0, 0],
                b["imu_site_pos"] if site_pos is None else site_pos[0], b["imu_site_quat"] if site_quat is None else site_quat[0]]

    def _sample_reset(self, rng):
        """train.py:1471-1476: RandomJointPositionReset(JOINT_BIASES, scale 0) + RandomJointVelocityReset [R: default scale]."""
        k1, k2, k3 = jax.random.split(rng, 3)
        qvel = jnp.zeros((NV,), jnp.float32).at[6:].set(jax.random.uniform(k1, (NU,), minval=-1.0, maxval=1.0) * self.reset_joint_vel_scale)
        t_rem = jax.random.uniform(k2, (), minval=self.push["interval_range"][0], maxval=self.push["interval_range"][1])
        lag = jax.random.uniform(k3, (), minval=self.imu_lag_range[0], maxval=self.imu_lag_range[1])  # train.py:1497
        return self.default_qpos, qvel, t_rem, lag

    def _sample_push(self, rng, curriculum_level):
        """PushEvent numbers consumed when the timer expires (train.py:1459-1468; combination rule [R], PARITY_RISKS.md ZB_PUSH_SAMPLING)."""
        k1, k2, k3 = jax.random.split(rng, 3)
        lo, hi = self.push["vel_range"]
        direction = jax.random.uniform(k1, (6,), minval=-1.0, maxval=1.0) * self.push["scale"]
        mag = jax.random.uniform(k2, (), minval=lo, maxval=hi) / hi
        nxt = jax.random.uniform(k3, (1,), minval=self.push["interval_range"][0], maxval=self.push["interval_range"][1])
        return jnp.concatenate([direction * mag * curriculum_level, nxt])

    # ---- PhysicsEngine interface ------------------------------------------------------------------------------------------
    def reset(self, physics_model, curriculum_level, rng):
        rq, rv, rt, lag = self._sample_reset(rng)
        command = jnp.zeros((7,), jnp.float32)  # ConstantZeroCommand, train.py:416-427
        call = jax.ffi.ffi_call("zb_reset_obs", _results(_RESET_RESULTS), vmap_method="broadcast_all")
        outs = call(rq, rv, rt, lag, command, *self._model_fields(physics_model), engine=np.int64(self.handle))
        return _unflatten(outs, _RESET_RESULTS, most_recent_action=rq[7:], action_latency=jnp.zeros(()), planner_cls=self.planner_cls)

    def step(self, action, physics_model, physics_state, curriculum_level, rng):
        k_push, k_reset = jax.random.split(rng)
        rq, rv, rt, lag = self._sample_reset(k_reset)
        command = jnp.zeros((7,), jnp.float32)
        call = jax.ffi.ffi_call("zb_step_obs", _results(_STEP_RESULTS), vmap_method="broadcast_all",
                                input_output_aliases={i: i for i in range(len(_STATE))})  # the state is updated in place
        outs = call(*_flatten(physics_state), action, self._sample_push(k_push, curriculum_level), command, rq, rv, rt, lag,
                    *self._model_fields(physics_model),
                    engine=np.int64(self.handle), n_substeps=np.int32(self.phys_steps_per_ctrl_steps), z_lo=np.float32(self.z_lo),
                    z_hi=np.float32(self.z_hi), max_tilt_rad=np.float32(self.max_tilt_rad),
                    max_episode_sec=np.float32(self.max_episode_sec))  # attributes are static (numpy scalars), operands are traced
        return _unflatten(outs, _STEP_RESULTS, most_recent_action=action, action_latency=physics_state.action_latency, planner_cls=self.planner_cls)


def feet_airtime_reward(left_contact, right_contact, done, carry, ctrl_dt=0.02, touchdown_penalty=0.3):
    """FeetAirtimeReward.get_reward_stateful over one rollout chunk (train.py:503-546) as one custom call: contacts / done [T, N],
    carry [N, 2] -> (new carry, reward [T, N])."""
    call = jax.ffi.ffi_call("zb_feet_airtime", [jax.ShapeDtypeStruct(carry.shape, jnp.float32), jax.ShapeDtypeStruct(left_contact.shape, jnp.float32)],
                            input_output_aliases={3: 0})
    return call(left_contact, right_contact, done, carry, ctrl_dt=np.float32(ctrl_dt), touchdown_penalty=np.float32(touchdown_penalty))


def zb_sample_action(packed_actor: dict, observations_actor: "jax.Array", actor_carry: "jax.Array", done_prev: "jax.Array", rng: "jax.Array", argmax: bool,
                     precision: int = 1, min_std: float = 0.01, max_std: float = 1.0, var_scale: float = 1.0):
    """Drop-in body for ``ZbotWalkingTask.sample_action`` (train.py:1737-1763): one ``zb_actor`` custom call per control step instead of
    ``Actor.forward`` + ``MixtureOfGaussians.mode() / .sample()``.  ``packed_actor`` = the eleven weight arrays of ``_ACTOR_WEIGHTS`` in kernel
    layout (``ksim-gym-zbot_b200/policy.py: ActorParams.pack`` — re-pack after every policy update); the random numbers of the sample path
    come from ``rng`` here (JAX's PRNG), the kernel only consumes them.  Returns (action, next_actor_carry)."""
    k1, k2 = jax.random.split(rng)
    u, eps = jax.random.uniform(k1, (NU,)), jax.random.normal(k2, (NU,))
    depth, num_mixtures = packed_actor["gru_b"].shape[0], packed_actor["out_b"].shape[0] // (3 * NU)
    call = jax.ffi.ffi_call("zb_actor", _results(_ACTOR_RESULTS), vmap_method="broadcast_all", input_output_aliases={12: 0})
    carry, action = call(*[packed_actor[n] for n, _, _ in _ACTOR_WEIGHTS], observations_actor, actor_carry, done_prev.astype(jnp.uint8), u, eps,
                         depth=np.int32(depth), num_mixtures=np.int32(num_mixtures), precision=np.int32(precision), argmax=np.int32(bool(argmax)),
                         min_std=np.float32(min_std), max_std=np.float32(max_std), var_scale=np.float32(var_scale))
    return action, carry
