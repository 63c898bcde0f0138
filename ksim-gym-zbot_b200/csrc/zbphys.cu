// zbphys.cu — kernels + the C ABI of include/zbphys.h (libzbphys.so).
//
// Drop-in boundary for ksim's MjxEngine.reset/step (ksim 0.1.99 ksim/engine.py, un-vendored:
// /root/reference/uv.lock:1872-1873).  Plain C entry points, caller-owned device memory, no torch.
// There is NO CPU fallback: without a CUDA device zb_engine_create returns ZB_ENODEV.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <new>

#include "../../include/zbphys.h"
#include "zb_kernel.cuh"
#if ZB_TU == 0
#include "zb_policy.cuh"
#include "zb_policy_tc.cuh"
#endif

#define ZB_ABI_VERSION 5
#define ZB_FSM_BYTES_PER_WARP ((31 + (ZB_CJ_SMEM ? 12 : 0)) * 32 * 4) /* WarpEnv::F_N floats per lane */

// auto-reset pass of the step kernel (zb_step_autoreset_f32)
struct ZbAutoReset {
  int enabled;
  float z_lo, z_hi, cos_tilt, max_time;
  const float *reset_qpos, *reset_qvel, *reset_t_rem;
  uint8_t* done_out;
  float *term_qpos, *term_qvel, *term_xpos, *term_xquat;  // Trajectory record (ZbAutoResetArgs)
};

struct ZbEngine {
  ZbDev host;
  ZbDev* dev;
  int device;
  int num_sms;
  int force_wpb;  // ZBPHYS_WPB environment override (experiments), 0 = automatic
};

// CTA shape of a launch: (profile, warps per CTA).  One CTA is resident per SM and a warp's speed hardly depends on its
// co-residents (the kernel is latency-bound per warp: 6.4 cycles per instruction even at 7 warps per SM), so the time of a batch
// is (number of waves) x (time of one CTA) and the shape is chosen to minimise the waves:
//   * n <= 16 SMs  : latency profile (128 registers), ceil(n / SMs) warps per CTA: every SM gets a small CTA
//                    (1024 envs: 147 CTAs of 7 warps, 0.35 ms instead of 64 CTAs of 16, 0.43 ms);
//   * larger       : throughput profile (72 registers, up to 28 warps): spills make one CTA ~1.6x slower (0.60 vs 0.37 ms per
//                    4 substeps) but it carries 1.75x the environments — 4096 envs run in ONE wave instead of 1.73
//                    (0.605 vs 0.737 ms), 8192 in two instead of 3.5 (1.18 vs 1.39 ms), 16 384: 2.29 vs 2.48 ms
//                    (profiles/r1_occupancy_sweep2.txt).  The CTAs are balanced over the waves: ceil(n / (waves SMs)) warps.
struct CtaShape { bool wide; int wpb; };
static CtaShape shape_policy(int n_envs, int sms, int force_wpb, bool wide_available) {
  if (sms <= 0) sms = 148;
  if (force_wpb > 0) return {wide_available && force_wpb > ZB_WPB, force_wpb > ZB_WPB && !wide_available ? ZB_WPB : force_wpb};
  if (n_envs <= ZB_WPB * sms || !wide_available) {
    const int w = (n_envs + sms - 1) / sms;
    return {false, w < 1 ? 1 : (w > ZB_WPB ? ZB_WPB : w)};
  }
  const long long per_wave = (long long)ZB_WPB_WIDE * sms;
  const long long waves = (n_envs + per_wave - 1) / per_wave;
  int w = (int)((n_envs + waves * sms - 1) / (waves * sms));
  if (w > ZB_WPB_WIDE) w = ZB_WPB_WIDE;
  return {true, w};
}
static bool uses_gen(const ZbEngine* e) { return e->host.any_prim || e->host.any_pair; }  // general collision instantiations (TU 1)
static int cflags_of(const ZbEngine* e) { return (e->host.any_box ? 1 : 0) | (e->host.any_prim ? 2 : 0) | (e->host.any_pair ? 4 : 0); }
static CtaShape pick_shape(const ZbEngine* e, int n_envs, bool wide_available) { return shape_policy(n_envs, e->num_sms, e->force_wpb, wide_available); }

// Launches go to the engine's device whatever the caller's current device is (an engine on device k used from a thread whose
// current device is another one launched on the wrong GPU).  The stream handed in must belong to the engine's device.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, const char* a = "") {
  snprintf(g_err, sizeof(g_err), fmt, a);
  return code;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// OBS: the observation epilogue is a separate instantiation so that the plain step keeps its register allocation
template <int MODE, bool TRACE, bool OBS, int WPB, bool GEN>
__global__ void __launch_bounds__(WPB * 32, 1) zb_kernel(const ZbDev* __restrict__ D, ZbOpts o, float meaninertia, int cflags, ZbEnvParams P, ZbState S,
                                                       const float* __restrict__ action, const float* __restrict__ push_rand,
                                                       int push_per_substep, ZbOut O, int n_envs, int n_substeps,
                                                       const uint8_t* __restrict__ done, const float* __restrict__ rq,
                                                       const float* __restrict__ rv, const float* __restrict__ rt, ZbAutoReset AR, ZbObs OB) {
  __shared__ float scratch[WPB][32 * ZB_SCR];
  __shared__ float imu_keep[OBS ? WPB : 1][16];  // sensors() -> observe() hand-over of the IMU sensor values
  const int L = threadIdx.x & 31, wib = threadIdx.x >> 5;
#ifndef ZB_BAR_GROUPS
#define ZB_BAR_GROUPS 2
#endif
  // lockstep groups: the warps of a CTA meet at substep boundaries in ZB_BAR_GROUPS groups (named barriers 1..G).  The
  // lockstep exists for instruction-fetch sharing, but every barrier also waits for the slowest environment of its group; two
  // groups measured +1 % over one (4096 and 16 384 envs), four no better (scripts/sweep_variants.sh).
  const int nwarp = blockDim.x >> 5;
#ifdef ZB_BAR_INTERLEAVE
  const int bar_g = wib % ZB_BAR_GROUPS;
  const int bar_cnt = (bar_g < nwarp ? (nwarp - 1 - bar_g) / ZB_BAR_GROUPS + 1 : 0) * 32;
#else
  const int bar_g = (ZB_BAR_GROUPS > 1) ? (wib * ZB_BAR_GROUPS) / nwarp : 0;
  const int bar_lo = (bar_g * nwarp + ZB_BAR_GROUPS - 1) / ZB_BAR_GROUPS, bar_hi = ((bar_g + 1) * nwarp + ZB_BAR_GROUPS - 1) / ZB_BAR_GROUPS;
  const int bar_cnt = (bar_hi - bar_lo) * 32;
#endif
#define ZB_LOCKSTEP_BARRIER() do { if (ZB_BAR_GROUPS > 1) asm volatile("bar.sync %0, %1;" ::"r"(bar_g + 1), "r"(bar_cnt) : "memory"); else __syncthreads(); } while (0)
  const int env = blockIdx.x * (blockDim.x >> 5) + wib;  // warps per CTA is a launch-time choice (<= ZB_WPB, see pick_wpb)
  // (vote: makes the flag provably warp-uniform for ptxas; a divergent-looking exit around barriers would wrap every
  //  later shuffle in WARPSYNC)
  const bool live = __all_sync(ZB_FULL, env < n_envs);
  if (!live) {
#if ZB_CTA_LOCKSTEP
    // a spare warp of the last CTA only keeps the barrier count: one per substep + one for the auto-reset pass
    if (MODE == ZB_MODE_STEP)
      for (int s = 0; s <= n_substeps; s++) ZB_LOCKSTEP_BARRIER();
#endif
    return;
  }
  if (MODE == ZB_MODE_RESET && done && !done[env]) return;
  const ZbModel& m = D->m;
  // throughput profile: factor + mass-matrix rows of every lane in (dynamic) shared memory, 31 x 128 B per warp
  constexpr bool FSM = WPB > ZB_WPB;
  extern __shared__ float dyn_smem[];
  WarpEnv<TRACE, FSM, GEN> E;
  E.init_topology(D, o, meaninertia, scratch[wib], L, cflags);
  E.fsm = FSM ? dyn_smem + (size_t)wib * (ZB_FSM_BYTES_PER_WARP / 4) + L : nullptr;
  const int b = E.isbody ? L : 0;
  const int d = E.hasdof ? E.dof : 0;
  // ---- per-environment model (randomizers, train.py:1439-1455) ----
  E.mass = E.isbody ? (P.body_mass ? P.body_mass[(size_t)env * NB + b] : m.body_mass[b]) : 0.f;
#pragma unroll
  for (int k = 0; k < 3; k++)
    E.inertia[k] = E.isbody ? (P.body_inertia ? P.body_inertia[((size_t)env * NB + b) * 3 + k] : m.body_inertia[b][k]) : 0.f;
  E.arm = E.hasdof ? (P.dof_armature ? P.dof_armature[(size_t)env * NV + d] : m.dof_armature[d]) : 0.f;
  E.damp = E.hasdof ? (P.dof_damping ? P.dof_damping[(size_t)env * NV + d] : m.dof_damping[d]) : 0.f;
  E.floss = E.hasdof ? (P.dof_frictionloss ? P.dof_frictionloss[(size_t)env * NV + d] : m.dof_frictionloss[d]) : 0.f;
  E.q0 = E.hinge ? (P.qpos0 ? P.qpos0[(size_t)env * NQ + d + 1] : m.qpos0[d + 1]) : 0.f;
  E.mu = P.floor_friction ? P.floor_friction[env] : m.floor_friction[0];
  E.mass_tot = wsum(E.mass);
  E.inv_mass_tot = 1.f / E.mass_tot;
  {
    const int s = E.isbody ? D->body_site[L] : -1;
    const int ss = s >= 0 ? s : 0;
#pragma unroll
    for (int k = 0; k < 3; k++) E.spos[k] = (s == 0 && P.imu_site_pos) ? P.imu_site_pos[(size_t)env * 3 + k] : m.site_pos[ss][k];
#pragma unroll
    for (int k = 0; k < 4; k++) E.squat[k] = (s == 0 && P.imu_site_quat) ? P.imu_site_quat[(size_t)env * 4 + k] : m.site_quat[ss][k];
  }
  // ---- state ----
  const float* qsrc = MODE == ZB_MODE_RESET ? rq : S.qpos;
  const float* vsrc = MODE == ZB_MODE_RESET ? rv : S.qvel;
#pragma unroll
  for (int k = 0; k < 7; k++) E.bq[k] = qsrc[(size_t)env * NQ + k];
  E.q = E.hinge ? qsrc[(size_t)env * NQ + d + 1] : 0.f;
  E.qv = E.hasdof ? vsrc[(size_t)env * NV + d] : 0.f;
  E.warm = (MODE != ZB_MODE_RESET && E.hasdof && S.qacc_warmstart) ? S.qacc_warmstart[(size_t)env * NV + d] : 0.f;
  const int a = E.act >= 0 ? E.act : 0;
  E.apos = E.avel = E.atq = E.tgt = E.ctrl = 0.f;
  float prev_tgt = 0.f;
  const int latency = (MODE == ZB_MODE_STEP && P.prev_action && P.latency_steps) ? P.latency_steps[env] : 0;
  if (E.act >= 0) {
    if (MODE == ZB_MODE_RESET) {  // get_initial_state train.py:1293-1298
      E.apos = E.q; E.avel = E.qv; E.atq = 0.f;
    } else {
      E.apos = S.act_pos[(size_t)env * NU + a]; E.avel = S.act_vel[(size_t)env * NU + a]; E.atq = S.act_torque[(size_t)env * NU + a];
    }
    if (MODE == ZB_MODE_STEP) E.tgt = action[(size_t)env * NU + a];
    if (MODE == ZB_MODE_STEP && P.prev_action) prev_tgt = P.prev_action[(size_t)env * NU + a];
    if (MODE == ZB_MODE_FORWARD) E.ctrl = E.atq;
  }
  float trem = 0.f, tm = 0.f;
  if (MODE == ZB_MODE_STEP) {
    trem = S.push_t_rem ? S.push_t_rem[env] : 0.f;
    tm = S.time ? S.time[env] : 0.f;
  }
  const float dt = o.dt;
  const bool implicit_damp = o.integrator == ZB_INT_IMPLICITFAST || (o.integrator == ZB_INT_EULER && o.eulerdamp && D->any_damping);
  const int nsub = MODE == ZB_MODE_STEP ? n_substeps : 1;
  float* sd = O.sensordata ? O.sensordata + (size_t)env * ZB_NSENSORDATA : nullptr;
  E.trace = (TRACE && O.solver_trace) ? O.solver_trace + (size_t)env * ZB_NTRACE : nullptr;
  const bool want_obs = OBS && MODE != ZB_MODE_FORWARD && (OB.actor_obs || OB.critic_obs || OB.imu_x || OB.feet_pos_obs || OB.reward_terms);
  bool reinit = MODE == ZB_MODE_RESET;  // observation carry restarts for environments this launch (re)initialises

  // s == nsub is the optional auto-reset pass (zb_step_autoreset_f32): terminate -> reset state -> forward(), i.e.
  // zb_terminations_f32 + zb_reset_f32(done) folded into the step launch (first half).
  for (int s = 0; s <= nsub; s++) {
    const bool last = s == nsub - 1;
    const bool extra = s == nsub;
#if ZB_CTA_LOCKSTEP
    if (MODE == ZB_MODE_STEP) ZB_LOCKSTEP_BARRIER();
#endif
    if (extra) {
      if (!(MODE == ZB_MODE_STEP && AR.enabled)) break;
      // Trajectory record of this control step: the integrated state + the poses forward() of the last substep left, before the
      // termination test can replace them (what ksim's rollout stores from next_physics_state)
      if (AR.term_qpos) {
        float v = E.bq[0];
#pragma unroll
        for (int k = 1; k < 7; k++) v = (L == k) ? E.bq[k] : v;
        if (L < 7) AR.term_qpos[(size_t)env * NQ + L] = v;
        if (E.hinge) AR.term_qpos[(size_t)env * NQ + d + 1] = E.q;
      }
      if (AR.term_qvel && E.hasdof) AR.term_qvel[(size_t)env * NV + d] = E.qv;
      if (E.isbody) {
        if (AR.term_xpos) { float* p = AR.term_xpos + ((size_t)env * NB + L) * 3; p[0] = E.xpos[0]; p[1] = E.xpos[1]; p[2] = E.xpos[2]; }
        if (AR.term_xquat) { float* p = AR.term_xquat + ((size_t)env * NB + L) * 4; p[0] = E.xquat[0]; p[1] = E.xquat[1]; p[2] = E.xquat[2]; p[3] = E.xquat[3]; }
      }
      const float up = E.bq[3] * E.bq[3] - E.bq[4] * E.bq[4] - E.bq[5] * E.bq[5] + E.bq[6] * E.bq[6];
      const bool nonfinite = !(isfinite(E.q) && isfinite(E.qv) && isfinite(E.bq[0]) && isfinite(E.bq[1]) && isfinite(E.bq[2]) && isfinite(E.bq[3]));
      bool bad = (E.bq[2] < AR.z_lo) || (E.bq[2] > AR.z_hi) || (up < AR.cos_tilt) || !isfinite(E.bq[2]) || !isfinite(up);
      bad = __any_sync(ZB_FULL, bad || nonfinite);
      const bool timeout = __all_sync(ZB_FULL, AR.max_time > 0.f && tm > AR.max_time);  // EpisodeLengthTermination, train.py:1592
      if (L == 0 && AR.done_out) AR.done_out[env] = bad ? 1 : (timeout ? 2 : 0);
      if (!(bad || timeout)) break;
      reinit = true;
#pragma unroll
      for (int k = 0; k < 7; k++) E.bq[k] = AR.reset_qpos[(size_t)env * NQ + k];
      E.q = E.hinge ? AR.reset_qpos[(size_t)env * NQ + d + 1] : 0.f;
      E.qv = E.hasdof ? AR.reset_qvel[(size_t)env * NV + d] : 0.f;
      E.warm = 0.f;
      E.apos = E.act >= 0 ? E.q : 0.f; E.avel = E.act >= 0 ? E.qv : 0.f; E.atq = 0.f; E.ctrl = 0.f;
      trem = AR.reset_t_rem ? AR.reset_t_rem[env] : 0.f;
      tm = 0.f;
    }
    const bool stepping = MODE == ZB_MODE_STEP && !extra;
    if (stepping) {
#pragma unroll
      for (int phase = 0; phase < 2; phase++) {
        const bool do_event = (o.substep_order == ZB_ORDER_EVENT_ACT_STEP) ? phase == 0 : phase == 1;
        if (do_event) {
          if (push_rand && S.push_t_rem) {  // PushEvent (train.py:1459-1468)
            const float* pr = push_rand + ((size_t)env * (push_per_substep ? n_substeps : 1) + (push_per_substep ? s : 0)) * 7;
            float t = trem - dt;
            if (t <= 0.f) {
              if (E.bdof) E.qv += pr[L - BL0];
              t = pr[6];
            }
            trem = t;
          }
        } else {
          const float new_tgt = E.tgt;
          if (s < latency) E.tgt = prev_tgt;  // action latency: the servo still tracks the previous action
          E.actuator(dt);
          E.tgt = new_tgt;
        }
      }
    }
    // ---- mjx.forward ----
    E.kinematics();
    E.mass_matrix();
    E.factor_m(0.f);  // (M + dt diag(damping) of the implicit integrators is factored inside solve_all, when needed)
    E.collision();
    float qfrc_passive;
    E.velocity(qfrc_passive);
    E.make_constraint();
#if ZB_EXPLICIT_J
    E.contact_rows_jacobian();
#endif
    E.bias_forces();
    const float qfrc_act = E.actuation();
    E.qfrc_smooth = E.hasdof ? qfrc_passive - E.qfrc_bias + qfrc_act : 0.f;
    const float qa = E.solve_all(implicit_damp && stepping);
    if (last || extra) E.final_rows();
    if (last || extra) E.sensors(sd, imu_keep[OBS ? wib : 0], O.sensordata != nullptr, OBS && want_obs);
    if (extra) E.warm = 0.f;
    if (stepping) {
      // ---- euler / implicitfast + _advance ----
      if (E.hasdof) E.qv += dt * qa;
      if (E.hinge) E.q += dt * E.qv;
      float vb[6];
#pragma unroll
      for (int k = 0; k < 6; k++) vb[k] = shf(E.qv, BL0 + k);
      E.bq[0] += dt * vb[0]; E.bq[1] += dt * vb[1]; E.bq[2] += dt * vb[2];
      {  // math.quat_integrate
        const float n = sqrtf(vb[3] * vb[3] + vb[4] * vb[4] + vb[5] * vb[5]);
        const float sc = 1.f / (n + 1e-6f * (n == 0.f));
        float sn, cs;
        sincosf(dt * n * 0.5f, &sn, &cs);
        const float qr[4] = {cs, vb[3] * sc * sn, vb[4] * sc * sn, vb[5] * sc * sn};
        float q2[4];
        quat_mul(q2, E.bq + 3, qr);
        normalize4(q2);
        E.bq[3] = q2[0]; E.bq[4] = q2[1]; E.bq[5] = q2[2]; E.bq[6] = q2[3];
      }
      E.warm = E.qacc;
      tm += dt;
    }
  }

  // ---- state write-back ----
  {  // base pose (MJX writes the normalised free-joint quaternion back into qpos, also in forward())
    float v = E.bq[0];
#pragma unroll
    for (int k = 1; k < 7; k++) v = (L == k) ? E.bq[k] : v;
    if (L < 7) S.qpos[(size_t)env * NQ + L] = v;
  }
  if (MODE != ZB_MODE_FORWARD) {
    if (E.hinge) S.qpos[(size_t)env * NQ + d + 1] = E.q;
    if (E.hasdof) {
      S.qvel[(size_t)env * NV + d] = E.qv;
      if (S.qacc_warmstart) S.qacc_warmstart[(size_t)env * NV + d] = MODE == ZB_MODE_RESET ? 0.f : E.warm;
    }
    if (E.act >= 0) {
      S.act_pos[(size_t)env * NU + a] = E.apos; S.act_vel[(size_t)env * NU + a] = E.avel; S.act_torque[(size_t)env * NU + a] = E.atq;
    }
    if (L == 0) {
      if (S.push_t_rem) S.push_t_rem[env] = MODE == ZB_MODE_RESET ? (rt ? rt[env] : 0.f) : trem;
      if (S.time) S.time[env] = MODE == ZB_MODE_RESET ? 0.f : tm;
    }
  }
  if (OBS && want_obs) E.observe(OB, env, reinit, imu_keep[OBS ? wib : 0]);
  // ---- outputs: mjx.Data fields as forward() of the last substep left them ----
  if (E.isbody) {
    if (O.xpos) { float* p = O.xpos + ((size_t)env * NB + L) * 3; p[0] = E.xpos[0]; p[1] = E.xpos[1]; p[2] = E.xpos[2]; }
    if (O.xquat) { float* p = O.xquat + ((size_t)env * NB + L) * 4; p[0] = E.xquat[0]; p[1] = E.xquat[1]; p[2] = E.xquat[2]; p[3] = E.xquat[3]; }
    if (O.cinert) {
      float* p = O.cinert + ((size_t)env * NB + L) * 10;
#pragma unroll
      for (int k = 0; k < 10; k++) p[k] = E.cinert[k];
    }
    if (O.cvel) {
      float* p = O.cvel + ((size_t)env * NB + L) * 6;
#pragma unroll
      for (int k = 0; k < 6; k++) p[k] = E.cvel[k];
    }
    const int st = D->body_site[L];
    if (st >= 0) {
      if (O.site_xpos) { float* p = O.site_xpos + ((size_t)env * ZB_NSITE + st) * 3; p[0] = E.sxpos[0]; p[1] = E.sxpos[1]; p[2] = E.sxpos[2]; }
      if (O.site_xmat) {
        float* p = O.site_xmat + ((size_t)env * ZB_NSITE + st) * 9;
#pragma unroll
        for (int k = 0; k < 9; k++) p[k] = E.sxmat[k];
      }
    }
  }
  if (E.hasdof) {
    if (O.qacc) O.qacc[(size_t)env * NV + d] = E.qacc;
    if (O.qfrc_constraint) O.qfrc_constraint[(size_t)env * NV + d] = E.qfrc_c;
    if (O.efc_force) {
      O.efc_force[(size_t)env * (2 * NV + ZB_NEFC_CON) + d] = E.fr_force;
      O.efc_force[(size_t)env * (2 * NV + ZB_NEFC_CON) + NV + d] = E.lm_force;
    }
  }
  if (E.act >= 0) {
    if (O.actuator_force) O.actuator_force[(size_t)env * NU + a] = E.act_force;
    if (O.ctrl) O.ctrl[(size_t)env * NU + a] = E.ctrl;
  }
  if (O.efc_force) O.efc_force[(size_t)env * (2 * NV + ZB_NEFC_CON) + 2 * NV + L] = E.c_force;
  if ((L & 3) == 0) {
    if (O.contact_dist) O.contact_dist[(size_t)env * ZB_MAXCON + (L >> 2)] = E.cdist;
    if (O.contact_pos) { float* p = O.contact_pos + ((size_t)env * ZB_MAXCON + (L >> 2)) * 3; p[0] = E.cpos[0]; p[1] = E.cpos[1]; p[2] = E.cpos[2]; }
  }
  {
    const unsigned bal = __ballot_sync(ZB_FULL, E.c_on && (L & 3) == 0);
    const bool bad = !(isfinite(E.q) && isfinite(E.qv) && isfinite(E.bq[0]) && isfinite(E.bq[1]) && isfinite(E.bq[2]) && isfinite(E.bq[3]));
    const unsigned nb = __ballot_sync(ZB_FULL, bad);
    if (L == 0) {
      int mask = 0;
#pragma unroll
      for (int c = 0; c < ZB_MAXCON; c++) mask |= ((bal >> (4 * c)) & 1u) << c;
      if (O.contact_mask) O.contact_mask[env] = mask;
      if (O.solver_niter) O.solver_niter[env] = E.niter;
      if (O.flags) O.flags[env] = nb ? 1 : 0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// launch dispatch.  The library is built from TWO translation units of this one source (build.py compiles them in parallel):
//   ZB_TU == 0: host code + the kernels for models whose collision groups are all plane-box (GEN = false: zbot_syn);
//   ZB_TU == 1: the kernels of the general collision groups (GEN = true), reached through zb_gen_launch / zb_gen_init.
// ZB_MINIMAL builds (variants, mutants) are TU 0 alone: plain step (both profiles, + trace) and plain reset, GEN = false only.
// ---------------------------------------------------------------------------------------------
struct KArgs {
  const ZbDev* D; ZbOpts o; float meaninertia; int cflags; ZbEnvParams P; ZbState S; const float* action; const float* push_rand;
  int push_per_substep; ZbOut O; int n_envs, n_substeps; const uint8_t* done; const float *rq, *rv, *rt; ZbAutoReset AR; ZbObs OB;
};
#define ZB_KLAUNCH(MODE, TR, OBSV, W, GENV)                                                                                        \
  zb_kernel<MODE, TR, OBSV, W, GENV><<<grid, threads, smem, st>>>(a.D, a.o, a.meaninertia, a.cflags, a.P, a.S, a.action, a.push_rand, \
                                                                 a.push_per_substep, a.O, a.n_envs, a.n_substeps, a.done, a.rq, a.rv, a.rt, a.AR, a.OB)
// returns false when the requested instantiation is not part of this build
template <bool GENV>
static bool launch_kernel(int mode, bool trace, bool obs, bool wide, int grid, int threads, size_t smem, cudaStream_t st, const KArgs& a) {
  if (mode == ZB_MODE_STEP) {
    if (!trace && !obs) { if (wide) ZB_KLAUNCH(ZB_MODE_STEP, false, false, ZB_WPB_WIDE, GENV); else ZB_KLAUNCH(ZB_MODE_STEP, false, false, ZB_WPB, GENV); return true; }
    if (trace && !obs && !wide) { ZB_KLAUNCH(ZB_MODE_STEP, true, false, ZB_WPB, GENV); return true; }
#if !defined(ZB_MINIMAL) || defined(ZB_MINIMAL_OBS)
    if (!trace && obs && wide) { ZB_KLAUNCH(ZB_MODE_STEP, false, true, ZB_WPB_WIDE, GENV); return true; }
#endif
#ifndef ZB_MINIMAL
    if (!trace && obs && !wide) { ZB_KLAUNCH(ZB_MODE_STEP, false, true, ZB_WPB, GENV); return true; }
    if constexpr (!GENV) { if (trace && obs && !wide) { ZB_KLAUNCH(ZB_MODE_STEP, true, true, ZB_WPB, false); return true; } }
#endif
    return false;
  }
  if (wide) return false;
  if (mode == ZB_MODE_RESET) {
    if (!trace && !obs) { ZB_KLAUNCH(ZB_MODE_RESET, false, false, ZB_WPB, GENV); return true; }
#ifndef ZB_MINIMAL
    if (!trace && obs) { ZB_KLAUNCH(ZB_MODE_RESET, false, true, ZB_WPB, GENV); return true; }
    if constexpr (!GENV) { if (trace) { if (obs) ZB_KLAUNCH(ZB_MODE_RESET, true, true, ZB_WPB, false); else ZB_KLAUNCH(ZB_MODE_RESET, true, false, ZB_WPB, false); return true; } }
#endif
    return false;
  }
#ifndef ZB_MINIMAL
  if (mode == ZB_MODE_FORWARD && !obs) {
    if (!trace) { ZB_KLAUNCH(ZB_MODE_FORWARD, false, false, ZB_WPB, GENV); return true; }
    if constexpr (!GENV) { ZB_KLAUNCH(ZB_MODE_FORWARD, true, false, ZB_WPB, false); return true; }
  }
#endif
  return false;
}
template <bool GENV>
static cudaError_t init_kernels() {  // the throughput-profile kernels keep 31 floats per lane in dynamic shared memory: 28 warps x 3968 B = 108.5 KB > the 48 KB default
  const int bytes = ZB_WPB_WIDE * ZB_FSM_BYTES_PER_WARP;
  cudaError_t a1 = cudaFuncSetAttribute(zb_kernel<ZB_MODE_STEP, false, false, ZB_WPB_WIDE, GENV>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
#if !defined(ZB_MINIMAL) || defined(ZB_MINIMAL_OBS)
  cudaError_t a2 = cudaFuncSetAttribute(zb_kernel<ZB_MODE_STEP, false, true, ZB_WPB_WIDE, GENV>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
#else
  cudaError_t a2 = cudaSuccess;
#endif
  return a1 != cudaSuccess ? a1 : a2;
}
#ifndef ZB_TU
#define ZB_TU 0
#endif
#if ZB_TU == 1
bool zb_gen_launch(int mode, bool trace, bool obs, bool wide, int grid, int threads, size_t smem, cudaStream_t st, const KArgs& a) {
  return launch_kernel<true>(mode, trace, obs, wide, grid, threads, smem, st, a);
}
cudaError_t zb_gen_init() { return init_kernels<true>(); }
#else
#ifdef ZB_MINIMAL
static bool zb_gen_launch(int, bool, bool, bool, int, int, size_t, cudaStream_t, const KArgs&) { return false; }
static cudaError_t zb_gen_init() { return cudaSuccess; }
#else
bool zb_gen_launch(int mode, bool trace, bool obs, bool wide, int grid, int threads, size_t smem, cudaStream_t st, const KArgs& a);  // TU 1
cudaError_t zb_gen_init();
#endif

// ksim BadZTermination + NotUprightTermination + EpisodeLengthTermination as configured at train.py:1588-1593 (+ non-finite guard),
// one thread per env.
__global__ void zb_term_kernel(const float* __restrict__ qpos, const int32_t* __restrict__ flags, const float* __restrict__ time,
                               uint8_t* __restrict__ done, int n_envs, float z_lo, float z_hi, float cos_tilt, float max_time) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_envs) return;
  const float* q = qpos + (size_t)i * NQ;
  const float z = q[2], w = q[3], x = q[4], y = q[5], zz = q[6];
  const float up = w * w - x * x - y * y + zz * zz;  // R[2][2]: world z-axis component of the base z-axis
  bool bad = (z < z_lo) || (z > z_hi) || (up < cos_tilt) || !isfinite(z) || !isfinite(up);
  if (flags) bad = bad || flags[i] != 0;
  const bool timeout = time && max_time > 0.f && time[i] > max_time;
  done[i] = bad ? 1 : (timeout ? 2 : 0);
}

// FeetAirtimeReward.get_reward_stateful of one rollout chunk (train.py:503-546), one thread per environment.
__global__ void zb_airtime_kernel(const float* __restrict__ lc, const float* __restrict__ rc, long long cts, long long ces,
                                  const uint8_t* __restrict__ done, long long dts, long long des, float* __restrict__ carry,
                                  float* __restrict__ reward, long long rts, long long res, int n_envs, int T, float ctrl_dt, float penalty) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_envs || T <= 0) return;
  float tl = carry[2 * i], tr = carry[2 * i + 1];  // time since lift-off
  float pl = 0.f, pr = 0.f;                        // airtime[t-1]
  bool cl_prev = false, cr_prev = false, cl0 = false, cr0 = false;
  for (int t = 0; t < T; t++) {
    const bool cl = lc[t * cts + i * ces] > 0.5f, cr = rc[t * cts + i * ces] > 0.5f, dn = done[t * dts + i * des] != 0;
    tl = (cl || dn) ? 0.f : tl + ctrl_dt;
    tr = (cr || dn) ? 0.f : tr + ctrl_dt;
    if (t == 0) { cl0 = cl; cr0 = cr; }
    else reward[t * rts + i * res] = (pl - penalty) * ((cl && !cl_prev) ? 1.f : 0.f) + (pr - penalty) * ((cr && !cr_prev) ? 1.f : 0.f);
    pl = tl; pr = tr; cl_prev = cl; cr_prev = cr;
  }
  // t = 0: touchdown against contact[-1] = false, airtime shifted in from the END of the chunk (jnp.roll)
  reward[i * res] = (tl - penalty) * (cl0 ? 1.f : 0.f) + (tr - penalty) * (cr0 ? 1.f : 0.f);
  carry[2 * i] = tl; carry[2 * i + 1] = tr;
}

// FP32 FFMA throughput probe: the roofline denominator for this path is the CUDA-core FP32 peak, which
// MEASURED_PEAKS.json does not carry (it has HBM GB/s and bf16 tensor TF/s only).  16 independent FMA chains per thread.
__global__ void __launch_bounds__(256) zb_ffma_kernel(float* __restrict__ out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = (float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = fmaf(x[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; k++) s += x[k];
  if (s == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------------------------------------
// host side: model validation + derived tables (what mjx.put_model + ksim.get_physics_engine do)
// ---------------------------------------------------------------------------------------------
static void h_quat_to_mat(float* m, const float* q) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}

// constraint._kbi constants (refsafe uses opt.timestep -> recomputed by zb_engine_set_opts)
static ZbSolPrm h_solprm(const ZbOpts& o, const float* solref, const float* solimp) {
  ZbSolPrm s;
  const float timeconst = fmaxf(solref[0], 2.f * o.dt), dampratio = solref[1];
  s.dmin = fminf(fmaxf(solimp[0], MJ_MINIMP), MJ_MAXIMP);
  s.dmax = fminf(fmaxf(solimp[1], MJ_MINIMP), MJ_MAXIMP);
  s.width = fmaxf(MJ_MINVAL, solimp[2]);
  s.mid = fminf(fmaxf(solimp[3], MJ_MINIMP), MJ_MAXIMP);
  s.power = fmaxf(1.f, solimp[4]);
  s.k = 1.f / (s.dmax * s.dmax * timeconst * timeconst * dampratio * dampratio);
  s.b = 2.f / (s.dmax * timeconst);
  if (solref[0] <= 0.f) { s.k = -solref[0] / (s.dmax * s.dmax); s.b = -solref[1] / s.dmax; }
  s.c1 = 1.f / powf(s.mid, s.power - 1.f);
  s.c2 = 1.f / powf(1.f - s.mid, s.power - 1.f);
  return s;
}
static void derive_solprm(ZbDev* D) {
  const ZbModel& m = D->m;
  for (int L = 0; L < 32; L++) {
    D->fr_R[L] = 0.f; D->fr_b[L] = 0.f;
    memset(&D->lim[L], 0, sizeof(ZbSolPrm));
    const int d = D->lane_dof[L];
    if (d < 0) continue;
    if (m.dof_frictionloss[d] > 0.f) {  // pos = 0 -> imp = dmin
      const ZbSolPrm s = h_solprm(D->o, m.dof_solref[d], m.dof_solimp[d]);
      const float imp = s.dmin;
      D->fr_R[L] = fmaxf(m.dof_invweight0[d] * (1.f - imp) / imp, MJ_MINVAL);
      D->fr_b[L] = s.b;
    }
    D->lim[L] = h_solprm(D->o, m.jnt_solref[d], m.jnt_solimp[d]);
  }
  for (int g = 0; g < ZB_NBOX; g++) {
    // contact parameter mixing (collision_driver): the higher geom_priority wins, a tie takes the larger friction and geom 1's solref / solimp
    const bool world = m.pair_body[g] == 0;
    const int p1 = world ? m.floor_priority : m.pair_priority[g], p2 = m.box_priority[g];
    const float* sr1 = world ? m.floor_solref : m.pair_solref[g];
    const float* si1 = world ? m.floor_solimp : m.pair_solimp[g];
    D->con[g] = h_solprm(D->o, p1 >= p2 ? sr1 : m.box_solref[g], p1 >= p2 ? si1 : m.box_solimp[g]);
    if (world) {  // the floor friction is a per-environment field (train.py:1446-1448)
      D->con_mu_mode[g] = p1 > p2 ? 0 : (p1 < p2 ? 1 : 2);
      D->con_mu_const[g] = m.box_friction[g][0];
    } else {
      D->con_mu_mode[g] = 1;
      D->con_mu_const[g] = p1 > p2 ? m.pair_friction[g][0] : (p1 < p2 ? m.box_friction[g][0] : fmaxf(m.pair_friction[g][0], m.box_friction[g][0]));
    }
    D->con_includemargin[g] = fmaxf(world ? m.floor_margin : m.pair_margin[g], m.box_margin[g]) - fmaxf(world ? m.floor_gap : m.pair_gap[g], m.box_gap[g]);
    D->con_invweight[g] = m.body_invweight0[m.pair_body[g]][0] + m.body_invweight0[m.box_body[g]][0];
  }
}

static int derive(ZbDev* D) {
  const ZbModel& m = D->m;
  if (m.body_jnttype[0] != 0 || m.body_jnttype[1] != 1 || m.body_parent[1] != 0 || m.body_dofadr[1] != 0)
    return fail(ZB_EUNSUP, "model: body 1 must be the free-joint base attached to the world%s");
  for (int L = 0; L < 32; L++) {
    D->cs[L] = D->ce[L] = D->gs[L] = L;
    D->gp[L] = 0; D->gn[L] = 0;
    D->lane_dof[L] = D->lane_act[L] = D->lane_box[L] = D->body_site[L] = D->body_box[L] = -1;
    D->bpj[L][0] = D->bpj[L][1] = D->bpj[L][2] = 0.f;
  }
  for (int a = 0; a < 6; a++) { D->lane_dof[BL0 + a] = a; D->gs[BL0 + a] = BL0; D->gp[BL0 + a] = a; D->gn[BL0 + a] = 6; }
  // chains: every body b >= 2 hangs off the base or continues the chain of body b-1
  for (int b = 2; b < NB; b++) {
    if (m.body_jnttype[b] == 1) return fail(ZB_EUNSUP, "model: only one free joint supported%s");
    if (m.body_parent[b] == 1) D->cs[b] = b;
    else if (m.body_parent[b] == b - 1) D->cs[b] = D->cs[b - 1];
    else return fail(ZB_EUNSUP, "model: tree must be a floating base with serial chains%s");
  }
  for (int b = NB - 1; b >= 2; b--) D->ce[b] = (b + 1 < NB && D->cs[b + 1] == D->cs[b]) ? D->ce[b + 1] : b;
  for (int b = 2; b < NB; b++) {
    if (D->ce[b] - D->cs[b] + 1 > 8) return fail(ZB_EUNSUP, "model: chains longer than 8 bodies unsupported%s");
    float R[9];
    h_quat_to_mat(R, m.body_quat[b]);
    for (int r = 0; r < 3; r++)
      D->bpj[b][r] = m.body_pos[b][r] + R[r * 3] * m.jnt_pos[b][0] + R[r * 3 + 1] * m.jnt_pos[b][1] + R[r * 3 + 2] * m.jnt_pos[b][2];
    if (m.body_jnttype[b] == 2) {
      const int pos = b - D->cs[b];
      if (pos >= CHD) return fail(ZB_EUNSUP, "model: more than 6 hinge dofs in one chain%s");
      // every body between the chain start and a hinge body must carry a hinge (contiguous dof positions)
      for (int c = D->cs[b]; c < b; c++)
        if (m.body_jnttype[c] != 2) return fail(ZB_EUNSUP, "model: jointless body inside a chain before a hinge%s");
      D->lane_dof[b] = m.body_dofadr[b];
      D->gs[b] = D->cs[b]; D->gp[b] = pos;
      if (m.dof_body[m.body_dofadr[b]] != b) return fail(ZB_EINVAL, "model: dof_body inconsistent%s");
    }
  }
  for (int b = 2; b < NB; b++)
    if (m.body_jnttype[b] == 2) {
      int n = 0;
      for (int c = D->cs[b]; c <= D->ce[b]; c++) n += m.body_jnttype[c] == 2;
      D->gn[b] = n;
    }
  for (int j = 0; j < NU; j++) {
    const int dof = m.act_dof[j];
    if (dof < 6 || dof >= NV) return fail(ZB_EINVAL, "model: act_dof out of range%s");
    D->lane_act[m.dof_body[dof]] = j;
  }
  D->any_box = D->any_prim = D->any_pair = 0;
  for (int L = 0; L < 32; L++) {
    D->body_pair[L] = -1;
    for (int g = 0; g < ZB_NBOX; g++) D->lane_jt[L][g] = -1;
  }
  for (int g = 0; g < ZB_NBOX; g++) {
    const int bb = m.box_body[g], b1 = m.pair_body[g], t2 = m.box_type[g], t1 = m.pair_type[g];
    if (bb < 2 || bb >= NB) return fail(ZB_EUNSUP, "model: the geom of a collision group must sit on a chain body%s");
    if (D->body_box[bb] >= 0) return fail(ZB_EUNSUP, "model: one box per body%s");
    if (t2 != ZB_GEOM_BOX && t2 != ZB_GEOM_SPHERE && t2 != ZB_GEOM_CAPSULE) return fail(ZB_EINVAL, "model: unknown box_type%s");
    if (b1 == 0) {
      D->grp_kind[g] = t2 == ZB_GEOM_BOX ? ZB_PAIR_PLANE_BOX : (t2 == ZB_GEOM_SPHERE ? ZB_PAIR_PLANE_SPHERE : ZB_PAIR_PLANE_CAPSULE);
      if (t2 == ZB_GEOM_BOX) D->any_box = 1; else D->any_prim = 1;
    } else {
      // mjx orders a pair by geom type (sphere < capsule): geom 1 = pair_*, geom 2 = box_*
      if (b1 < 2 || b1 >= NB || b1 == bb) return fail(ZB_EUNSUP, "model: pair_body must be another chain body (or 0 = the floor)%s");
      if (t1 == ZB_GEOM_SPHERE && t2 == ZB_GEOM_SPHERE) D->grp_kind[g] = ZB_PAIR_SPHERE_SPHERE;
      else if (t1 == ZB_GEOM_SPHERE && t2 == ZB_GEOM_CAPSULE) D->grp_kind[g] = ZB_PAIR_SPHERE_CAPSULE;
      else if (t1 == ZB_GEOM_CAPSULE && t2 == ZB_GEOM_CAPSULE) D->grp_kind[g] = ZB_PAIR_CAPSULE_CAPSULE;
      else return fail(ZB_EUNSUP, "model: body-body groups: sphere-sphere, sphere-capsule (sphere = geom 1), capsule-capsule%s");
      if (D->body_pair[b1] >= 0) return fail(ZB_EUNSUP, "model: one geom 1 per body%s");
      D->body_pair[b1] = g;
      D->any_pair = 1;
    }
    D->body_box[bb] = g;
    for (int c = D->cs[bb]; c <= bb; c++)
      if (m.body_jnttype[c] == 2) {
        if (D->lane_box[c] >= 0) return fail(ZB_EUNSUP, "model: one box per chain%s");
        D->lane_box[c] = g;
      }
    D->box_site[g] = -1;
  }
  for (int s = 0; s < ZB_NSITE; s++) {
    const int sb = m.site_body[s];
    if (sb < 1 || sb >= NB || D->body_site[sb] >= 0) return fail(ZB_EUNSUP, "model: one site per (non-world) body%s");
    D->body_site[sb] = s;
    if (s > 0 && D->body_box[sb] >= 0) D->box_site[D->body_box[sb]] = s;
  }
  // lane_jt: which entry of group g's transposed J^T f totals belongs to the dof of lane L (contact_jtf, contact_rows_jacobian)
  for (int g = 0; g < ZB_NBOX; g++) {
    const int bb = m.box_body[g], b1 = m.pair_body[g];
    for (int c = D->cs[bb]; c <= bb; c++)
      if (m.body_jnttype[c] == 2) D->lane_jt[c][g] = D->gp[c];
    if (b1 == 0) {
      for (int a = 0; a < 6; a++) D->lane_jt[BL0 + a][g] = 6 + a;
    } else {
      const bool same_chain = D->cs[b1] == D->cs[bb];
      for (int c = D->cs[b1]; c <= b1; c++)
        if (m.body_jnttype[c] == 2) D->lane_jt[c][g] = same_chain ? D->gp[c] : 6 + D->gp[c];  // (same chain: the two blocks are merged into the first)
    }
  }
  {  // math.make_frame(floor_normal)
    float a[3] = {m.floor_normal[0], m.floor_normal[1], m.floor_normal[2]};
    float n = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (!(n > 0)) return fail(ZB_EINVAL, "model: floor_normal is zero%s");
    for (int k = 0; k < 3; k++) a[k] /= n;
    float bb[3] = {0, 0, 0};
    if (-0.5f < a[1] && a[1] < 0.5f) bb[1] = 1; else bb[2] = 1;
    const float ab = a[0] * bb[0] + a[1] * bb[1] + a[2] * bb[2];
    for (int k = 0; k < 3; k++) bb[k] -= a[k] * ab;
    n = sqrtf(bb[0] * bb[0] + bb[1] * bb[1] + bb[2] * bb[2]);
    for (int k = 0; k < 3; k++) bb[k] /= n;
    const float c[3] = {a[1] * bb[2] - a[2] * bb[1], a[2] * bb[0] - a[0] * bb[2], a[0] * bb[1] - a[1] * bb[0]};
    for (int k = 0; k < 3; k++) { D->frame[k] = a[k]; D->frame[3 + k] = bb[k]; D->frame[6 + k] = c[k]; }
  }
  D->any_damping = 0;
  for (int i = 0; i < NV; i++) D->any_damping |= m.dof_damping[i] > 0;
  derive_solprm(D);
  return ZB_OK;
}

static int check_model_opts(const ZbDev* D, const ZbOpts* o) {
  if (D->any_pair && o->solver == ZB_SOLVER_NEWTON)
    return fail(ZB_EUNSUP, "Newton solver with body-body collision groups: the Hessian loses its block-arrow structure (use CG)%s");
  return ZB_OK;
}
static int check_opts(const ZbOpts* o) {
  if (!(o->dt > 0) || o->iterations < 0 || o->ls_iterations < 0 || !(o->impratio > 0)) return fail(ZB_EINVAL, "opts: dt/iterations/impratio invalid%s");
  if (o->solver != ZB_SOLVER_CG && o->solver != ZB_SOLVER_NEWTON) return fail(ZB_EINVAL, "opts: unknown solver%s");
  if (o->integrator != ZB_INT_EULER && o->integrator != ZB_INT_IMPLICITFAST) return fail(ZB_EINVAL, "opts: unknown integrator%s");
  if (o->substep_order != ZB_ORDER_EVENT_ACT_STEP && o->substep_order != ZB_ORDER_ACT_EVENT_STEP) return fail(ZB_EINVAL, "opts: unknown substep_order%s");
  return ZB_OK;
}

extern "C" {

int zb_abi_version(void) { return ZB_ABI_VERSION; }
int zb_cta_shape(int n_envs, int num_sms, int* wide, int* warps_per_cta) {
  if (n_envs < 1 || !wide || !warps_per_cta) return fail(ZB_EINVAL, "zb_cta_shape: bad argument%s");
  const CtaShape cs = shape_policy(n_envs, num_sms, 0, true);
  *wide = cs.wide ? 1 : 0;
  *warps_per_cta = cs.wpb;
  return ZB_OK;
}
int zb_sizeof_model(void) { return (int)sizeof(ZbModel); }
int zb_sizeof_opts(void) { return (int)sizeof(ZbOpts); }
const char* zb_last_error(void) { return g_err; }
long long zb_launch_count(void) { return g_launches.load(); }

int zb_engine_create(const ZbModel* model, const ZbOpts* opts, int device, ZbEngine** out) {
  if (!model || !opts || !out) return fail(ZB_EINVAL, "zb_engine_create: null argument%s");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(ZB_ENODEV, "zb_engine_create: no CUDA device (there is no CPU fallback)%s");
  }
  if (device < 0 || device >= ndev) return fail(ZB_EINVAL, "zb_engine_create: bad device index%s");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(ZB_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(cudaGetLastError()));
  if (prop.major != 10) return fail(ZB_ENODEV, "zb_engine_create: kernels are built for sm_100a only%s");
  int rc = check_opts(opts);
  if (rc) return rc;
  ZbEngine* e = new (std::nothrow) ZbEngine;
  if (!e) return fail(ZB_EINVAL, "out of host memory%s");
  memset(&e->host, 0, sizeof(e->host));
  e->host.m = *model;
  e->host.o = *opts;
  e->device = device;
  e->dev = nullptr;
  e->num_sms = prop.multiProcessorCount;
  e->force_wpb = 0;
  if (const char* w = getenv("ZBPHYS_WPB")) {
    const int v = atoi(w);
    if (v >= 1 && v <= ZB_WPB_WIDE) e->force_wpb = v;  // > ZB_WPB selects the throughput profile
  }
  rc = derive(&e->host);
  if (!rc) rc = check_model_opts(&e->host, opts);
  if (rc) { delete e; return rc; }
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  {
    cudaError_t a1 = init_kernels<false>();
    if (a1 == cudaSuccess && uses_gen(e)) a1 = zb_gen_init();
    if (a1 != cudaSuccess) {
      cudaSetDevice(prev);
      delete e;
      return fail(ZB_ECUDA, "zb_engine_create: cudaFuncSetAttribute(MaxDynamicSharedMemorySize): %s", cudaGetErrorString(a1));
    }
  }
  cudaError_t ce = cudaMalloc(&e->dev, sizeof(ZbDev));
  if (ce == cudaSuccess) ce = cudaMemcpy(e->dev, &e->host, sizeof(ZbDev), cudaMemcpyHostToDevice);
  cudaSetDevice(prev);
  if (ce != cudaSuccess) {
    if (e->dev) cudaFree(e->dev);
    delete e;
    return fail(ZB_ECUDA, "zb_engine_create: %s", cudaGetErrorString(ce));
  }
  *out = e;
  return ZB_OK;
}

int zb_engine_destroy(ZbEngine* e) {
  if (!e) return ZB_OK;
  if (e->dev) cudaFree(e->dev);
  delete e;
  return ZB_OK;
}

int zb_engine_set_opts(ZbEngine* e, const ZbOpts* opts) {
  if (!e || !opts) return fail(ZB_EINVAL, "zb_engine_set_opts: null argument%s");
  int rc = check_opts(opts);
  if (!rc) rc = check_model_opts(&e->host, opts);
  if (rc) return rc;
  e->host.o = *opts;
  derive_solprm(&e->host);
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(e->device);
  cudaError_t ce = cudaMemcpy(e->dev, &e->host, sizeof(ZbDev), cudaMemcpyHostToDevice);
  cudaSetDevice(prev);
  if (ce != cudaSuccess) return fail(ZB_ECUDA, "zb_engine_set_opts: %s", cudaGetErrorString(ce));
  return ZB_OK;
}

static int check_state(const ZbState* s) {
  if (!s || !s->qpos || !s->qvel || !s->act_pos || !s->act_vel || !s->act_torque) return fail(ZB_EINVAL, "state: qpos/qvel/act_* must be non-null%s");
  return ZB_OK;
}
static int after_launch(const char* what) {
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(ZB_ECUDA, what, cudaGetErrorString(ce));
  g_launches.fetch_add(1);
  return ZB_OK;
}

static int dispatch(const ZbEngine* e, int mode, bool trace, bool obs, bool wide, int grid, int threads, size_t smem, void* stream, const KArgs& a,
                    const char* what) {
  const bool ok = uses_gen(e) ? zb_gen_launch(mode, trace, obs, wide, grid, threads, smem, (cudaStream_t)stream, a)
                              : launch_kernel<false>(mode, trace, obs, wide, grid, threads, smem, (cudaStream_t)stream, a);
  if (!ok) return fail(ZB_EUNSUP, "this combination (mode / trace / observation epilogue / profile / collision groups) is not instantiated in this build%s");
  return after_launch(what);
}

static int step_impl(ZbEngine* e, void* stream, const ZbEnvParams* params, const ZbState* state, const float* action,
                     const float* push_rand, int push_per_substep, const ZbOut* out, int n_envs, int n_substeps, ZbAutoReset AR,
                     const ZbObs* obs = nullptr) {
  if (!e || !action) return fail(ZB_EINVAL, "zb_step_f32: null engine/action%s");
  if (n_envs < 0 || n_substeps < 1) return fail(ZB_EINVAL, "zb_step_f32: n_envs >= 0 and n_substeps >= 1 required%s");
  int rc = check_state(state);
  if (rc) return rc;
  if (n_envs == 0) return ZB_OK;
  DeviceGuard guard(e->device);
  ZbEnvParams P; memset(&P, 0, sizeof(P)); if (params) P = *params;
  ZbOut O; memset(&O, 0, sizeof(O)); if (out) O = *out;
  ZbObs OB; memset(&OB, 0, sizeof(OB)); if (obs) OB = *obs;
  const bool with_obs = OB.actor_obs || OB.critic_obs || OB.imu_x || OB.feet_pos_obs || OB.reward_terms;
  const CtaShape cs = pick_shape(e, n_envs, !O.solver_trace);  // the diagnostics instantiation exists in the latency profile only
  const int wpb = cs.wpb, grid = (n_envs + wpb - 1) / wpb;
  KArgs a = {e->dev, e->host.o, e->host.m.meaninertia, cflags_of(e), P, *state, action, push_rand, push_per_substep, O, n_envs, n_substeps,
             nullptr, nullptr, nullptr, nullptr, AR, OB};
  // (the diagnostics instantiation exists in the latency profile only)
  return dispatch(e, ZB_MODE_STEP, O.solver_trace != nullptr, with_obs, cs.wide && !O.solver_trace, grid, wpb * 32,
                  (cs.wide && !O.solver_trace) ? (size_t)wpb * ZB_FSM_BYTES_PER_WARP : 0, stream, a, "zb_step_f32 launch: %s");
}

int zb_step_f32(ZbEngine* e, void* stream, const ZbEnvParams* params, const ZbState* state, const float* action,
                const float* push_rand, int push_per_substep, const ZbOut* out, int n_envs, int n_substeps) {
  ZbAutoReset AR;
  memset(&AR, 0, sizeof(AR));
  return step_impl(e, stream, params, state, action, push_rand, push_per_substep, out, n_envs, n_substeps, AR);
}

int zb_step_autoreset_f32(ZbEngine* e, void* stream, const ZbEnvParams* params, const ZbState* state, const float* action,
                          const float* push_rand, int push_per_substep, const ZbOut* out, int n_envs, int n_substeps,
                          const float* reset_qpos, const float* reset_qvel, const float* reset_t_rem, uint8_t* done_out,
                          float z_lo, float z_hi, float max_tilt_rad, float max_episode_sec) {
  if (!reset_qpos || !reset_qvel) return fail(ZB_EINVAL, "zb_step_autoreset_f32: null reset_qpos/reset_qvel%s");
  if (max_episode_sec > 0 && state && !state->time) return fail(ZB_EINVAL, "zb_step_autoreset_f32: max_episode_sec needs ZbState.time%s");
  ZbAutoReset AR;
  memset(&AR, 0, sizeof(AR));
  AR.enabled = 1; AR.z_lo = z_lo; AR.z_hi = z_hi; AR.cos_tilt = cosf(max_tilt_rad); AR.max_time = max_episode_sec;
  AR.reset_qpos = reset_qpos; AR.reset_qvel = reset_qvel; AR.reset_t_rem = reset_t_rem; AR.done_out = done_out;
  return step_impl(e, stream, params, state, action, push_rand, push_per_substep, out, n_envs, n_substeps, AR);
}

int zb_forward_f32(ZbEngine* e, void* stream, const ZbEnvParams* params, const ZbState* state, const ZbOut* out, int n_envs) {
  if (!e) return fail(ZB_EINVAL, "zb_forward_f32: null engine%s");
  if (n_envs < 0) return fail(ZB_EINVAL, "zb_forward_f32: n_envs < 0%s");
  int rc = check_state(state);
  if (rc) return rc;
  if (n_envs == 0) return ZB_OK;
  DeviceGuard guard(e->device);
  ZbEnvParams P; memset(&P, 0, sizeof(P)); if (params) P = *params;
  ZbOut O; memset(&O, 0, sizeof(O)); if (out) O = *out;
  const int wpb = pick_shape(e, n_envs, false).wpb, grid = (n_envs + wpb - 1) / wpb;
  KArgs a = {e->dev, e->host.o, e->host.m.meaninertia, cflags_of(e), P, *state, nullptr, nullptr, 0, O, n_envs, 1, nullptr, nullptr, nullptr, nullptr,
             ZbAutoReset{}, ZbObs{}};
  return dispatch(e, ZB_MODE_FORWARD, O.solver_trace != nullptr, false, false, grid, wpb * 32, 0, stream, a, "zb_forward_f32 launch: %s");
}

static int check_obs(const ZbObs* obs) {
  if (obs && (obs->actor_obs || obs->critic_obs) && !obs->imu_x) return fail(ZB_EINVAL, "obs: imu_x (filter carry) is required when actor_obs / critic_obs are requested%s");
  if (obs && obs->imu_x && !obs->imu_lag) return fail(ZB_EINVAL, "obs: imu_lag is required with imu_x%s");
  return ZB_OK;
}

static int reset_impl(ZbEngine* e, void* stream, const ZbEnvParams* params, const ZbState* state, const uint8_t* done,
                      const float* reset_qpos, const float* reset_qvel, const float* reset_t_rem, const ZbOut* out, const ZbObs* obs,
                      int n_envs) {
  if (!e || !reset_qpos || !reset_qvel) return fail(ZB_EINVAL, "zb_reset_f32: null engine/reset_qpos/reset_qvel%s");
  if (n_envs < 0) return fail(ZB_EINVAL, "zb_reset_f32: n_envs < 0%s");
  int rc = check_state(state);
  if (rc) return rc;
  if (n_envs == 0) return ZB_OK;
  DeviceGuard guard(e->device);
  ZbEnvParams P; memset(&P, 0, sizeof(P)); if (params) P = *params;
  ZbOut O; memset(&O, 0, sizeof(O)); if (out) O = *out;
  ZbObs OB; memset(&OB, 0, sizeof(OB)); if (obs) OB = *obs;
  const int wpb = pick_shape(e, n_envs, false).wpb, grid = (n_envs + wpb - 1) / wpb;
  const bool with_obs = OB.actor_obs || OB.critic_obs || OB.imu_x || OB.feet_pos_obs || OB.reward_terms;
  KArgs a = {e->dev, e->host.o, e->host.m.meaninertia, cflags_of(e), P, *state, nullptr, nullptr, 0, O, n_envs, 1, done, reset_qpos, reset_qvel, reset_t_rem,
             ZbAutoReset{}, OB};
  return dispatch(e, ZB_MODE_RESET, O.solver_trace != nullptr, with_obs, false, grid, wpb * 32, 0, stream, a, "zb_reset_f32 launch: %s");
}

int zb_reset_f32(ZbEngine* e, void* stream, const ZbEnvParams* params, const ZbState* state, const uint8_t* done,
                 const float* reset_qpos, const float* reset_qvel, const float* reset_t_rem, const ZbOut* out, int n_envs) {
  return reset_impl(e, stream, params, state, done, reset_qpos, reset_qvel, reset_t_rem, out, nullptr, n_envs);
}

int zb_reset_obs_f32(ZbEngine* e, void* stream, const ZbEnvParams* params, const ZbState* state, const uint8_t* done,
                     const float* reset_qpos, const float* reset_qvel, const float* reset_t_rem, const ZbOut* out, const ZbObs* obs,
                     int n_envs) {
  int rc = check_obs(obs);
  if (rc) return rc;
  return reset_impl(e, stream, params, state, done, reset_qpos, reset_qvel, reset_t_rem, out, obs, n_envs);
}

int zb_step_obs_f32(ZbEngine* e, void* stream, const ZbEnvParams* params, const ZbState* state, const float* action,
                    const float* push_rand, int push_per_substep, const ZbOut* out, const ZbObs* obs, int n_envs, int n_substeps,
                    const ZbAutoResetArgs* ar) {
  int rc = check_obs(obs);
  if (rc) return rc;
  ZbAutoReset AR;
  memset(&AR, 0, sizeof(AR));
  if (ar) {
    if (!ar->reset_qpos || !ar->reset_qvel) return fail(ZB_EINVAL, "zb_step_obs_f32: null reset_qpos/reset_qvel%s");
    if (ar->max_episode_sec > 0 && state && !state->time) return fail(ZB_EINVAL, "zb_step_obs_f32: max_episode_sec needs ZbState.time%s");
    AR.enabled = 1; AR.z_lo = ar->z_lo; AR.z_hi = ar->z_hi; AR.cos_tilt = cosf(ar->max_tilt_rad); AR.max_time = ar->max_episode_sec;
    AR.reset_qpos = ar->reset_qpos; AR.reset_qvel = ar->reset_qvel; AR.reset_t_rem = ar->reset_t_rem; AR.done_out = ar->done_out;
    AR.term_qpos = ar->term_qpos; AR.term_qvel = ar->term_qvel; AR.term_xpos = ar->term_xpos; AR.term_xquat = ar->term_xquat;
  }
  return step_impl(e, stream, params, state, action, push_rand, push_per_substep, out, n_envs, n_substeps, AR, obs);
}

int zb_terminations_f32(void* stream, const float* qpos, const int32_t* flags, const float* time, uint8_t* done, int n_envs, float z_lo,
                        float z_hi, float max_tilt_rad, float max_episode_sec) {
  if (!qpos || !done || n_envs < 0) return fail(ZB_EINVAL, "zb_terminations_f32: bad argument%s");
  if (n_envs == 0) return ZB_OK;
  zb_term_kernel<<<(n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(qpos, flags, time, done, n_envs, z_lo, z_hi, cosf(max_tilt_rad), max_episode_sec);
  return after_launch("zb_terminations_f32 launch: %s");
}

int zb_feet_airtime_f32(void* stream, const float* left_contact, const float* right_contact, long long cts, long long ces,
                        const uint8_t* done, long long dts, long long des, float* carry, float* reward, long long rts, long long res,
                        int n_envs, int n_steps, float ctrl_dt, float touchdown_penalty) {
  if (!left_contact || !right_contact || !done || !carry || !reward || n_envs < 0 || n_steps < 0) return fail(ZB_EINVAL, "zb_feet_airtime_f32: bad argument%s");
  if (n_envs == 0 || n_steps == 0) return ZB_OK;
  zb_airtime_kernel<<<(n_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(left_contact, right_contact, cts, ces, done, dts, des, carry, reward, rts, res,
                                                                          n_envs, n_steps, ctrl_dt, touchdown_penalty);
  return after_launch("zb_feet_airtime_f32 launch: %s");
}

/* Actor network of the reference for one control step (zb_policy.cuh). */
int zb_actor_f32(void* stream, const ZbActor* A, const float* actor_obs, const float* carry_in, float* carry_out, const uint8_t* carry_reset,
                 const float* noise_u, const float* noise_eps, float* action, float* dist, int n_envs) {
  if (!A || !actor_obs || !carry_in || !carry_out || !action || n_envs < 0) return fail(ZB_EINVAL, "zb_actor_f32: bad argument%s");
  if (!A->in_wt || !A->in_b || !A->gru_wih_t || !A->gru_whh_t || !A->gru_b || !A->gru_bn || !A->out_wt || !A->out_b)
    return fail(ZB_EINVAL, "zb_actor_f32: NULL weight pointer%s");
  if ((((uintptr_t)A->in_wt | (uintptr_t)A->gru_wih_t | (uintptr_t)A->gru_whh_t | (uintptr_t)A->out_wt) & 15) != 0)
    return fail(ZB_EINVAL, "zb_actor_f32: weight matrices must be 16-byte aligned (they are fetched by TMA bulk copies)%s");
  if ((noise_u == nullptr) != (noise_eps == nullptr)) return fail(ZB_EINVAL, "zb_actor_f32: noise_u and noise_eps come together (both NULL = mode)%s");
  if (A->n_in < 1 || A->n_in > ZB_ACTOR_MAX_IN || A->depth < 1 || A->num_mixtures < 1 || A->num_mixtures > ZB_ACTOR_MAX_MIXTURES)
    return fail(ZB_EUNSUP, "zb_actor_f32: n_in <= 128, depth >= 1, 1 <= num_mixtures <= 5 (hidden size 128 is built in)%s");
  const bool tc = A->precision == ZB_ACTOR_TF32X3;
  if (A->precision != ZB_ACTOR_FP32 && !tc) return fail(ZB_EINVAL, "zb_actor_f32: unknown ZbActor.precision%s");
  if (tc && (!A->gru_w_tc || !A->out_w_tc || (((uintptr_t)A->gru_w_tc | (uintptr_t)A->out_w_tc) & 15) != 0))
    return fail(ZB_EINVAL, "zb_actor_f32: ZB_ACTOR_TF32X3 needs gru_w_tc and out_w_tc (16-byte aligned)%s");
  if (tc && A->n_in > ZBT_MAX_IN) return fail(ZB_EUNSUP, "zb_actor_f32: ZB_ACTOR_TF32X3 is built for n_in <= 96%s");
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return fail(ZB_ENODEV, "zb_actor_f32: no CUDA device%s");
  static std::atomic<unsigned long long> attr_set{0};
  if (dev < 64 && !((attr_set.load() >> dev) & 1ull)) {
    cudaError_t ce = cudaFuncSetAttribute(zb_actor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZbpSmem));
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(zb_actor_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ZbtSmem));
    if (ce != cudaSuccess) return fail(ZB_ECUDA, "zb_actor_f32: cudaFuncSetAttribute(MaxDynamicSharedMemorySize): %s", cudaGetErrorString(ce));
    attr_set.fetch_or(1ull << dev);
  }
  if (n_envs == 0) return ZB_OK;  // (a call with n_envs = 0 only prepares the kernels: do one before capturing launches into a CUDA graph)
  const int grid = (n_envs + ZBP_TE - 1) / ZBP_TE;
  if (tc)
    zb_actor_tc_kernel<<<grid, ZBP_THREADS, sizeof(ZbtSmem), (cudaStream_t)stream>>>(*A, actor_obs, carry_in, carry_out, carry_reset, noise_u, noise_eps, action, dist, n_envs);
  else
    zb_actor_kernel<<<grid, ZBP_THREADS, sizeof(ZbpSmem), (cudaStream_t)stream>>>(*A, actor_obs, carry_in, carry_out, carry_reset, noise_u, noise_eps, action, dist, n_envs);
  return after_launch("zb_actor_f32 launch: %s");
}
int zb_sizeof_actor(void) { return (int)sizeof(ZbActor); }

/* Launches the FFMA probe: grid x 256 threads, `iters` x 16 FMAs per thread.  The caller times it with CUDA events;
 * flops = grid * 256 * iters * 16 * 2. */
int zb_bench_ffma(void* stream, float* scratch, int grid, int iters) {
  if (!scratch || grid <= 0 || iters <= 0) return fail(ZB_EINVAL, "zb_bench_ffma: bad argument%s");
  zb_ffma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scratch, iters, 0.999f, 1e-3f);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) return fail(ZB_ECUDA, "zb_bench_ffma launch: %s", cudaGetErrorString(ce));
  return ZB_OK;
}

}  // extern "C"
#endif  // ZB_TU == 0
