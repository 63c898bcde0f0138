"""ksim-side binding of the CUDA Z-Bot physics step (to be used where JAX + ksim exist; NOT importable in this image).

The reference builds its engine in ``ksim.RLTask.get_engine`` -> ``ksim.get_physics_engine(...)`` (ksim 0.1.99, un-vendored:
/root/reference/uv.lock:1872-1873).  There is no plugin registry, so the drop-in route is this ``PhysicsEngine`` subclass returned
from an overridden ``get_engine`` — the ONLY edit in /root/reference/train.py:

    class ZbotWalkingTask(ksim.PPOTask[ZbotWalkingTaskConfig]):
        def get_engine(self, physics_model, metadata=None):
            return ZbotCudaEngine.create(self, physics_model, metadata)

``step`` issues ONE XLA custom call (integration/zb_xla_ffi.cc -> zb_step_obs_f32) per control step under ``jax.vmap`` (the
call is registered with ``vmap_method="broadcast_all"``: every operand arrives with the leading environment dimension the kernel
expects) inside ksim's ``jax.lax.scan``.  Random numbers stay JAX's: PushEvent draws, reset states and the IMU lag of
re-initialised environments are sampled here and handed to the kernel as plain arrays.
"""
from __future__ import annotations

import ctypes

import jax
import jax.numpy as jnp
import numpy as np
import ksim  # noqa: F401  (PhysicsEngine, PhysicsState)

_lib = ctypes.CDLL("libzb_ffi.so")
jax.ffi.register_ffi_target("zb_step_obs", jax.ffi.pycapsule(_lib.zb_step_obs_ffi), platform="CUDA")
jax.ffi.register_ffi_target("zb_feet_airtime", jax.ffi.pycapsule(_lib.zb_feet_airtime_ffi), platform="CUDA")

NQ, NV, NU, NB, NSITE = 27, 26, 20, 26, 3
_STATE = ("qpos", "qvel", "qacc_warmstart", "act_pos", "act_vel", "act_torque", "push_t_rem", "time", "imu_x", "imu_lag")
_STATE_SHAPES = ((NQ,), (NV,), (NV,), (NU,), (NU,), (NU,), (), (), (4,), ())
_OUT = (("xpos", (NB, 3)), ("xquat", (NB, 4)), ("qacc", (NV,)), ("actuator_force", (NU,)), ("sensordata", (18,)), ("site_xpos", (NSITE, 3)),
        ("site_xmat", (NSITE, 9)), ("done", ()), ("actor_obs", (50,)), ("critic_obs", (484,)), ("feet_pos_obs", (6,)), ("reward_terms", (6,)))


def zb_step_obs(engine_handle: int, n_substeps: int, state: dict, action, push_rand, command, reset: dict, model_fields: dict,
                z_lo=0.05, z_hi=0.5, max_tilt_rad=1.0471975512):
    """One control step for ONE environment (call it under jax.vmap).  ``state``: the ten arrays of ``_STATE``; ``reset``:
    qpos / qvel / t_rem / imu_lag of a fresh episode (ksim resets, train.py:1471-1476); ``model_fields``: the per-environment
    randomizer outputs (train.py:1439-1455) in ZbEnvParams order.  Returns (new_state dict, outputs dict)."""
    f32, u8 = jnp.float32, jnp.uint8
    results = [jax.ShapeDtypeStruct(s, f32) for s in _STATE_SHAPES] + [jax.ShapeDtypeStruct(s, u8 if k == "done" else f32) for k, s in _OUT]
    call = jax.ffi.ffi_call("zb_step_obs", results, vmap_method="broadcast_all",
                            input_output_aliases={i: i for i in range(len(_STATE))})  # state is updated in place
    outs = call(*[state[k] for k in _STATE], action, push_rand, command, reset["qpos"], reset["qvel"], reset["t_rem"], reset["imu_lag"],
                *[model_fields[k] for k in ("body_mass", "body_inertia", "dof_armature", "dof_damping", "dof_frictionloss", "qpos0",
                                            "floor_friction", "imu_site_pos", "imu_site_quat")],
                engine=np.int64(engine_handle), n_substeps=np.int32(n_substeps), z_lo=np.float32(z_lo), z_hi=np.float32(z_hi),
                max_tilt_rad=np.float32(max_tilt_rad))  # attributes are static (numpy scalars), operands are traced
    new_state = dict(zip(_STATE, outs[:len(_STATE)]))
    return new_state, {k: v for (k, _), v in zip(_OUT, outs[len(_STATE):])}
