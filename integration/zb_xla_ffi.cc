// zb_xla_ffi.cc — XLA FFI custom-call handlers around the C ABI of include/zbphys.h (libzbphys.so).
//
// UNEXERCISED IN THIS IMAGE: jaxlib (and with it xla/ffi/api/ffi.h) is not installed here, nor on the GPU boxes
// (SURVEY.md §8b "What can be exercised in this image").  What IS checked here (tests/test_integration_cpu.py): the file
// type-checks under `g++ -fsyntax-only` against integration/stub/xla/ffi/api/ffi.h — a minimal stand-in of the FFI API that
// static_asserts every handler against its Bind() chain — and the operand / result order, count and dtypes of each handler
// equal the tables of integration/zbot_cuda_engine.py.  This file is the binding a ksim / JAX maintainer compiles where JAX
// exists:
//
//   g++ -O2 -std=c++17 -fPIC -shared zb_xla_ffi.cc -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I../include -I/usr/local/cuda/include -L../ksim-gym-zbot_b200/csrc -lzbphys -lcudart -o libzb_ffi.so
//
// and registers from Python (integration/zbot_cuda_engine.py):
//   jax.ffi.register_ffi_target("zb_step_obs", jax.ffi.pycapsule(lib.zb_step_obs_ffi), platform="CUDA")
//
// What it replaces on the reference side: the body of ksim's rollout scan around the engine — `MjxEngine.step`
// (ksim 0.1.99 ksim/engine.py, /root/reference/uv.lock:1872-1873), the termination check
// (/root/reference/train.py:1588-1593), the `done -> engine.reset` select (train.py:1471-1476) and the gathering of the
// network inputs (train.py:1624-1681) — as ONE custom call per control step; `MjxEngine.reset` as one call per rollout start.
//
// Contract with XLA: all buffers are XLA-owned device memory, the handler only enqueues on the stream XLA passes, holds no
// host state (re-entrant, legal inside scan / while, command-buffer compatible), needs no VJP (PPO never differentiates the
// physics, train.py:1683-1729).  State operands are declared as input_output_aliases of the corresponding results on the
// Python side, so the kernel updates them in place; if XLA could not alias (it then hands us distinct buffers) the state is
// copied operand -> result first.
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include "xla/ffi/api/ffi.h"
#include "zbphys.h"

namespace ffi = xla::ffi;
using F32 = ffi::Buffer<ffi::F32>;
using U8 = ffi::Buffer<ffi::U8>;
using S32 = ffi::Buffer<ffi::S32>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;
using RS32 = ffi::ResultBuffer<ffi::S32>;

static void alias_or_copy(cudaStream_t s, const F32& in, RF32& out) {
  if (out->typed_data() != in.typed_data())
    cudaMemcpyAsync(out->typed_data(), in.typed_data(), in.size_bytes(), cudaMemcpyDeviceToDevice, s);
}

static ffi::Error status(int rc) { return rc == ZB_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, zb_last_error()); }

// One control step: n_substeps x { PushEvent -> Feetech servo -> mjx.step } + terminations + partial reset + observations.
static ffi::Error ZbStepObsImpl(
    cudaStream_t stream, int64_t engine, int32_t n_substeps, float z_lo, float z_hi, float max_tilt_rad, float max_episode_sec,
    // state (aliased with the first ten results)
    F32 qpos, F32 qvel, F32 qacc_warmstart, F32 act_pos, F32 act_vel, F32 act_torque, F32 push_t_rem, F32 time, F32 imu_x, F32 imu_lag,
    // per-step inputs
    F32 action, F32 push_rand, F32 command, F32 reset_qpos, F32 reset_qvel, F32 reset_t_rem, F32 reset_imu_lag,
    // per-environment model (ksim randomizers, train.py:1439-1455)
    F32 body_mass, F32 body_inertia, F32 dof_armature, F32 dof_damping, F32 dof_frictionloss, F32 qpos0, F32 floor_friction,
    F32 imu_site_pos, F32 imu_site_quat,
    // results: state
    RF32 qpos_o, RF32 qvel_o, RF32 qacc_warmstart_o, RF32 act_pos_o, RF32 act_vel_o, RF32 act_torque_o, RF32 push_t_rem_o, RF32 time_o,
    RF32 imu_x_o, RF32 imu_lag_o,
    // results: mjx.Data fields the task reads (train.py:452-461, 476, 681, 854, 882) and the rollout outputs
    RF32 xpos, RF32 xquat, RF32 cinert, RF32 cvel, RF32 qacc, RF32 actuator_force, RF32 ctrl, RF32 sensordata, RF32 site_xpos, RF32 site_xmat,
    RU8 done, RF32 actor_obs, RF32 critic_obs, RF32 feet_pos_obs, RF32 reward_terms) {
  const int n = static_cast<int>(qpos.dimensions()[0]);
  alias_or_copy(stream, qpos, qpos_o); alias_or_copy(stream, qvel, qvel_o); alias_or_copy(stream, qacc_warmstart, qacc_warmstart_o);
  alias_or_copy(stream, act_pos, act_pos_o); alias_or_copy(stream, act_vel, act_vel_o); alias_or_copy(stream, act_torque, act_torque_o);
  alias_or_copy(stream, push_t_rem, push_t_rem_o); alias_or_copy(stream, time, time_o);
  alias_or_copy(stream, imu_x, imu_x_o); alias_or_copy(stream, imu_lag, imu_lag_o);

  ZbState st = {qpos_o->typed_data(), qvel_o->typed_data(), qacc_warmstart_o->typed_data(), act_pos_o->typed_data(), act_vel_o->typed_data(),
                act_torque_o->typed_data(), push_t_rem_o->typed_data(), time_o->typed_data()};
  ZbEnvParams p = {};
  p.body_mass = body_mass.typed_data(); p.body_inertia = body_inertia.typed_data(); p.dof_armature = dof_armature.typed_data();
  p.dof_damping = dof_damping.typed_data(); p.dof_frictionloss = dof_frictionloss.typed_data(); p.qpos0 = qpos0.typed_data();
  p.floor_friction = floor_friction.typed_data(); p.imu_site_pos = imu_site_pos.typed_data(); p.imu_site_quat = imu_site_quat.typed_data();
  ZbOut out = {};
  out.xpos = xpos->typed_data(); out.xquat = xquat->typed_data(); out.cinert = cinert->typed_data(); out.cvel = cvel->typed_data();
  out.qacc = qacc->typed_data(); out.actuator_force = actuator_force->typed_data(); out.ctrl = ctrl->typed_data();
  out.sensordata = sensordata->typed_data(); out.site_xpos = site_xpos->typed_data(); out.site_xmat = site_xmat->typed_data();
  ZbObs ob = {};
  ob.command = command.typed_data(); ob.imu_x = imu_x_o->typed_data(); ob.imu_lag = imu_lag_o->typed_data();
  ob.reset_imu_lag = reset_imu_lag.typed_data(); ob.actor_obs = actor_obs->typed_data(); ob.critic_obs = critic_obs->typed_data();
  ob.feet_pos_obs = feet_pos_obs->typed_data(); ob.reward_terms = reward_terms->typed_data();
  ZbAutoResetArgs ar = {reset_qpos.typed_data(), reset_qvel.typed_data(), reset_t_rem.typed_data(), done->typed_data(), z_lo, z_hi, max_tilt_rad,
                        max_episode_sec};
  return status(zb_step_obs_f32(reinterpret_cast<ZbEngine*>(engine), stream, &p, &st, action.typed_data(), push_rand.typed_data(),
                                /*push_per_substep=*/0, &out, &ob, n, n_substeps, &ar));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    zb_step_obs_ffi, ZbStepObsImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("engine").Attr<int32_t>("n_substeps").Attr<float>("z_lo").Attr<float>("z_hi").Attr<float>("max_tilt_rad").Attr<float>("max_episode_sec")
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()   // state + imu carry
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()                                   // action ... reset_imu_lag
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()             // per-env model
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()   // state + imu carry
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()   // xpos ... site_xmat
        .Ret<U8>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());                                                        // done, actor, critic, feet, terms

// MjxEngine.reset (+ get_initial_state, train.py:1293-1298) for a whole batch: the state results are written from the
// host-sampled reset arrays (ksim resets, train.py:1471-1476), forward() runs on them and the first observation is produced.
static ffi::Error ZbResetObsImpl(
    cudaStream_t stream, int64_t engine,
    // inputs
    F32 reset_qpos, F32 reset_qvel, F32 reset_t_rem, F32 reset_imu_lag, F32 command,
    // per-environment model
    F32 body_mass, F32 body_inertia, F32 dof_armature, F32 dof_damping, F32 dof_frictionloss, F32 qpos0, F32 floor_friction,
    F32 imu_site_pos, F32 imu_site_quat,
    // results: state
    RF32 qpos_o, RF32 qvel_o, RF32 qacc_warmstart_o, RF32 act_pos_o, RF32 act_vel_o, RF32 act_torque_o, RF32 push_t_rem_o, RF32 time_o,
    RF32 imu_x_o, RF32 imu_lag_o,
    // results: mjx.Data fields + observations of the fresh episode
    RF32 xpos, RF32 xquat, RF32 cinert, RF32 cvel, RF32 qacc, RF32 actuator_force, RF32 ctrl, RF32 sensordata, RF32 site_xpos, RF32 site_xmat,
    RF32 actor_obs, RF32 critic_obs, RF32 feet_pos_obs, RF32 reward_terms) {
  const int n = static_cast<int>(reset_qpos.dimensions()[0]);
  // the filter carry of a fresh episode starts from x = 0 (train.py:846-848); the kernel writes x after the first observation
  cudaMemsetAsync(imu_x_o->typed_data(), 0, imu_x_o->size_bytes(), stream);
  cudaMemcpyAsync(imu_lag_o->typed_data(), reset_imu_lag.typed_data(), reset_imu_lag.size_bytes(), cudaMemcpyDeviceToDevice, stream);
  ZbState st = {qpos_o->typed_data(), qvel_o->typed_data(), qacc_warmstart_o->typed_data(), act_pos_o->typed_data(), act_vel_o->typed_data(),
                act_torque_o->typed_data(), push_t_rem_o->typed_data(), time_o->typed_data()};
  ZbEnvParams p = {};
  p.body_mass = body_mass.typed_data(); p.body_inertia = body_inertia.typed_data(); p.dof_armature = dof_armature.typed_data();
  p.dof_damping = dof_damping.typed_data(); p.dof_frictionloss = dof_frictionloss.typed_data(); p.qpos0 = qpos0.typed_data();
  p.floor_friction = floor_friction.typed_data(); p.imu_site_pos = imu_site_pos.typed_data(); p.imu_site_quat = imu_site_quat.typed_data();
  ZbOut out = {};
  out.xpos = xpos->typed_data(); out.xquat = xquat->typed_data(); out.cinert = cinert->typed_data(); out.cvel = cvel->typed_data();
  out.qacc = qacc->typed_data(); out.actuator_force = actuator_force->typed_data(); out.ctrl = ctrl->typed_data();
  out.sensordata = sensordata->typed_data(); out.site_xpos = site_xpos->typed_data(); out.site_xmat = site_xmat->typed_data();
  ZbObs ob = {};
  ob.command = command.typed_data(); ob.imu_x = imu_x_o->typed_data(); ob.imu_lag = imu_lag_o->typed_data();
  ob.reset_imu_lag = reset_imu_lag.typed_data(); ob.actor_obs = actor_obs->typed_data(); ob.critic_obs = critic_obs->typed_data();
  ob.feet_pos_obs = feet_pos_obs->typed_data(); ob.reward_terms = reward_terms->typed_data();
  return status(zb_reset_obs_f32(reinterpret_cast<ZbEngine*>(engine), stream, &p, &st, /*done=*/nullptr, reset_qpos.typed_data(),
                                 reset_qvel.typed_data(), reset_t_rem.typed_data(), &out, &ob, n));
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    zb_reset_obs_ffi, ZbResetObsImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int64_t>("engine")
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()                                                         // reset arrays, command
        .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()             // per-env model
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()   // state + imu carry
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>()   // xpos ... site_xmat
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>());                                                                  // actor, critic, feet, terms

// FeetAirtimeReward over a rollout chunk (train.py:503-546): contacts / done / reward as [T, N] arrays, carry [N, 2].
static ffi::Error ZbFeetAirtimeImpl(cudaStream_t stream, float ctrl_dt, float touchdown_penalty, F32 left_contact, F32 right_contact,
                                    U8 done, F32 carry, RF32 carry_o, RF32 reward) {
  const int T = static_cast<int>(left_contact.dimensions()[0]), n = static_cast<int>(left_contact.dimensions()[1]);
  alias_or_copy(stream, carry, carry_o);
  return status(zb_feet_airtime_f32(stream, left_contact.typed_data(), right_contact.typed_data(), n, 1, done.typed_data(), n, 1,
                                    carry_o->typed_data(), reward->typed_data(), n, 1, n, T, ctrl_dt, touchdown_penalty));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(zb_feet_airtime_ffi, ZbFeetAirtimeImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<float>("ctrl_dt").Attr<float>("touchdown_penalty")
                                  .Arg<F32>().Arg<F32>().Arg<U8>().Arg<F32>().Ret<F32>().Ret<F32>());
// Actor.forward + action selection per control step (train.py:946-968, 1737-1763; zb_actor_f32): weights in the kernel layout
// (ksim-gym-zbot_b200/policy.py ActorParams.pack; re-packed once per policy update), carry aliased with its result.
static ffi::Error ZbActorImpl(cudaStream_t stream, int32_t depth, int32_t num_mixtures, int32_t precision, int32_t argmax, float min_std, float max_std,
                              float var_scale, F32 in_wt, F32 in_b, F32 gru_wih_t, F32 gru_whh_t, F32 gru_b, F32 gru_bn, F32 out_wt, F32 out_b,
                              F32 mean_bias, F32 gru_w_tc, F32 out_w_tc, F32 actor_obs, F32 carry, U8 carry_reset, F32 noise_u, F32 noise_eps,
                              RF32 carry_o, RF32 action) {
  const int n = static_cast<int>(actor_obs.dimensions()[0]);
  ZbActor a;
  a.n_in = static_cast<int32_t>(actor_obs.dimensions()[1]);
  a.depth = depth; a.num_mixtures = num_mixtures; a.min_std = min_std; a.max_std = max_std; a.var_scale = var_scale;
  a.precision = precision; a.reserved = 0;
  a.in_wt = in_wt.typed_data(); a.in_b = in_b.typed_data(); a.gru_wih_t = gru_wih_t.typed_data(); a.gru_whh_t = gru_whh_t.typed_data();
  a.gru_b = gru_b.typed_data(); a.gru_bn = gru_bn.typed_data(); a.out_wt = out_wt.typed_data(); a.out_b = out_b.typed_data();
  a.mean_bias = mean_bias.typed_data(); a.gru_w_tc = gru_w_tc.typed_data(); a.out_w_tc = out_w_tc.typed_data();
  return status(zb_actor_f32(stream, &a, actor_obs.typed_data(), carry.typed_data(), carry_o->typed_data(), carry_reset.typed_data(),
                             argmax ? nullptr : noise_u.typed_data(), argmax ? nullptr : noise_eps.typed_data(), action->typed_data(), nullptr, n));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(zb_actor_ffi, ZbActorImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int32_t>("depth").Attr<int32_t>("num_mixtures").Attr<int32_t>("precision").Attr<int32_t>("argmax")
                                  .Attr<float>("min_std").Attr<float>("max_std").Attr<float>("var_scale")
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()  // weights
                                  .Arg<F32>().Arg<F32>().Arg<U8>().Arg<F32>().Arg<F32>()                                                                      // obs, carry, reset, noise
                                  .Ret<F32>().Ret<F32>());
#else
#error "xla/ffi/api/ffi.h not found: build this file where jaxlib is installed (-I$(python -c 'import jax; print(jax.ffi.include_dir())'))"
#endif
