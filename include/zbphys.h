/*
 * zbphys.h — C ABI of the B200-native batched Z-Bot rigid-body physics step.
 *
 * This is the drop-in boundary for the ONE hot path of schnetzlerjoe/ksim-gym-zbot:
 * what ksim's MjxEngine.reset / MjxEngine.step do for every environment
 * (upstream ksim 0.1.99 ksim/engine.py — un-vendored dependency pinned at
 * /root/reference/uv.lock:1872-1873; physics = mujoco-mjx 3.3.4, uv.lock:2226-2227),
 * with the reference's own per-substep servo model fused in
 * (/root/reference/train.py:1137-1196 trapezoidal_step, :1242-1280 get_stateful_ctrl).
 *
 * Rules of the boundary (what an XLA-FFI / ctypes / cgo binding may rely on):
 *   - plain C types only; every array pointer is CALLER-OWNED DEVICE memory, fp32/int32,
 *     row-major with the environment index leading ([N, ...] — the layout jax.vmap gives);
 *   - calls only enqueue work on the given cudaStream_t; they never synchronise, never
 *     allocate and never throw.  Return value: 0 = ok, negative = ZB_E* error code;
 *   - the state arrays are updated in place; outputs never alias state;
 *   - there is NO CPU fallback: without a CUDA device zb_engine_create fails with ZB_ENODEV.
 */
#ifndef ZBPHYS_H_
#define ZBPHYS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- compile-time sizes of the Z-Bot topology (train.py:27 NUM_JOINTS, :59-82 joint order,
 *      :42-43 CRITIC_DIM com_inertia=250=10*25 bodies, com_velocity=150=6*25) ---------------- */
#define ZB_NQ 27          /* 7 (free joint) + 20 hinges                       */
#define ZB_NV 26          /* 6 + 20                                           */
#define ZB_NU 20          /* actuators == hinge joints, train.py:27           */
#define ZB_NBODY 26       /* world + 25 bodies                                */
#define ZB_NSITE 3        /* imu_site, left_foot, right_foot  (train.py:1453, 1519-1520) */
#define ZB_NBOX 2         /* foot collision boxes vs. the floor plane         */
#define ZB_MAXCON 8       /* 4 manifold slots per foot box (MJX plane_convex) */
#define ZB_NEFC_CON (4 * ZB_MAXCON) /* pyramidal condim-3 rows                */
#define ZB_NTRACE 160     /* solver_trace floats per environment              */
#define ZB_TRACE_POINTS 32  /* first index of the line-search points in solver_trace */
#define ZB_TRACE_SEARCH 128 /* first index of the first search direction        */
#define ZB_TRACE_START 154  /* start point of the solve: 1 qacc_warmstart, 2 qacc_smooth */
#define ZB_NSENSORDATA 18 /* quat4 acc3 gyro3 touchL touchR forceL3 forceR3   */
#define ZB_NCOMMAND 7     /* commands["zero_command"], train.py:416-427       */
#define ZB_NACTOR_OBS 50  /* NUM_ACTOR_INPUTS, train.py:31-38, 53             */
#define ZB_NCRITIC_OBS 484 /* NUM_CRITIC_INPUTS, train.py:40-52, 54           */
#define ZB_NREWARD_TERMS 6 /* see ZbObs.reward_terms                          */

enum { ZB_SOLVER_CG = 0, ZB_SOLVER_NEWTON = 1 };
enum { ZB_GEOM_BOX = 0, ZB_GEOM_SPHERE = 1, ZB_GEOM_CAPSULE = 2 }; /* ZbModel.box_type / pair_type (0 = box keeps zero-filled models valid) */
enum { ZB_INT_EULER = 0, ZB_INT_IMPLICITFAST = 1 };
enum { ZB_ORDER_EVENT_ACT_STEP = 0, ZB_ORDER_ACT_EVENT_STEP = 1 };

enum {
  ZB_OK = 0,
  ZB_EINVAL = -1,  /* null/ill-formed argument                      */
  ZB_ENODEV = -2,  /* no CUDA device / wrong architecture           */
  ZB_ECUDA = -3,   /* a CUDA runtime call failed (see zb_last_error)*/
  ZB_EUNSUP = -4   /* option combination not built into the kernels */
};

/* Sensor layout inside sensordata[ZB_NSENSORDATA] (train.py:44-46,1495-1516,516-517,724-725) */
enum {
  ZB_SENS_IMU_QUAT = 0,   /* framequat imu_site (4) */
  ZB_SENS_IMU_ACC = 4,    /* accelerometer (3)      */
  ZB_SENS_IMU_GYRO = 7,   /* gyro (3)               */
  ZB_SENS_TOUCH_L = 10,   /* left_foot_touch (1)    */
  ZB_SENS_TOUCH_R = 11,   /* right_foot_touch (1)   */
  ZB_SENS_FORCE_L = 12,   /* left_foot_force (3)    */
  ZB_SENS_FORCE_R = 15    /* right_foot_force (3)   */
};

/* Physics options == the mjOption fields ksim sets from its config
 * (train.py:1777-1780 dt/ctrl_dt/iterations/ls_iterations; the rest are ksim/MuJoCo defaults,
 * each one a named parity switch, see PARITY_RISKS.md). */
typedef struct ZbOpts {
  float dt;            /* physics timestep                                   */
  float tolerance;     /* solver tolerance (MuJoCo default 1e-8)             */
  float ls_tolerance;  /* line-search tolerance (MuJoCo default 0.01)        */
  float impratio;      /* MuJoCo default 1                                    */
  float gravity[3];    /* (0,0,-9.81)                                         */
  int32_t iterations;    /* train.py:1779 -> 8                                */
  int32_t ls_iterations; /* train.py:1780 -> 8                                */
  int32_t solver;        /* ZB_SOLVER_*                                       */
  int32_t integrator;    /* ZB_INT_*                                          */
  int32_t eulerdamp;     /* Euler only: 1 = implicit-in-damping               */
  int32_t warmstart;     /* 1 = pick cheaper of qacc_warmstart / qacc_smooth  */
  int32_t substep_order; /* ZB_ORDER_*                                        */
  int32_t reserved;
} ZbOpts;

/* The compiled model (what mjx.put_model holds for the fields this path reads), fp32. */
typedef struct ZbModel {
  /* tree */
  int32_t body_parent[ZB_NBODY];
  int32_t body_dofadr[ZB_NBODY];  /* first dof of the body's joint, -1 = jointless  */
  int32_t body_jnttype[ZB_NBODY]; /* 0 = none, 1 = free, 2 = hinge                  */
  int32_t dof_body[ZB_NV];
  int32_t dof_parent[ZB_NV];      /* parent dof in the kinematic chain, -1 = none   */
  float body_pos[ZB_NBODY][3];
  float body_quat[ZB_NBODY][4];
  float body_ipos[ZB_NBODY][3];
  float body_iquat[ZB_NBODY][4];
  float body_mass[ZB_NBODY];
  float body_inertia[ZB_NBODY][3];
  float body_invweight0[ZB_NBODY][2];
  float jnt_pos[ZB_NBODY][3];     /* hinge anchor in the body frame (per body)      */
  float jnt_axis[ZB_NBODY][3];
  /* per dof (hinge dofs 6..25 are meaningful for range/solref/solimp) */
  float jnt_range[ZB_NV][2];
  float jnt_margin[ZB_NV];
  float jnt_solref[ZB_NV][2];
  float jnt_solimp[ZB_NV][5];
  int32_t jnt_limited[ZB_NV];
  float dof_armature[ZB_NV];
  float dof_damping[ZB_NV];
  float dof_frictionloss[ZB_NV];
  float dof_invweight0[ZB_NV];
  float dof_solref[ZB_NV][2];
  float dof_solimp[ZB_NV][5];
  float qpos0[ZB_NQ];
  /* motors: <joint>_ctrl, train.py:1358-1362 */
  int32_t act_dof[ZB_NU];
  float act_gear[ZB_NU];
  float act_ctrlrange[ZB_NU][2];
  float act_forcerange[ZB_NU][2];
  int32_t act_ctrllimited[ZB_NU];
  int32_t act_forcelimited[ZB_NU];
  /* Feetech servo constants, train.py:1421-1431 (per joint, action order) */
  float servo_kp[ZB_NU];
  float servo_kd[ZB_NU];
  float servo_max_pwm[ZB_NU];
  float servo_vin[ZB_NU];
  float servo_kt[ZB_NU];
  float servo_R[ZB_NU];
  float servo_vmax[ZB_NU];
  float servo_amax[ZB_NU];
  float servo_error_gain[ZB_NU];
  float servo_deadband_pos; /* train.py:1111-1118 */
  float servo_deadband_neg;
  /* floor plane: geom "floor", priority 2 (train.py:1329-1330) */
  float floor_pos[3];
  float floor_normal[3];
  float floor_friction[3];
  float floor_solref[2];
  float floor_solimp[5];
  int32_t floor_priority;
  float floor_margin;
  float floor_gap;
  /* Collision groups (mjx.collision_driver's static candidate list).  Group g collides geom 2 = the box_*
   * fields (on body box_body[g], shape box_type[g]) with geom 1 = the floor plane (pair_body[g] == 0) or a second geom on body
   * pair_body[g] (self-collision; pair_* fields).  MJX orders a pair by geom type (plane < sphere < capsule < box) and the
   * contact normal points from geom 1 to geom 2; primitives and contacts per group [R: mjx collision_primitive / collision_convex]:
   *   plane-box 4 (plane_convex + _manifold_points) | plane-capsule 2 | plane-sphere 1 | sphere-sphere, sphere-capsule,
   *   capsule-capsule 1.   zbot_syn: two foot boxes against the floor. */
  int32_t box_body[ZB_NBOX];
  float box_pos[ZB_NBOX][3];
  float box_quat[ZB_NBOX][4];
  float box_size[ZB_NBOX][3];
  float box_friction[ZB_NBOX][3];
  float box_solref[ZB_NBOX][2];
  float box_solimp[ZB_NBOX][5];
  int32_t box_priority[ZB_NBOX];
  float box_margin[ZB_NBOX];
  float box_gap[ZB_NBOX];
  int32_t box_type[ZB_NBOX];      /* ZB_GEOM_*: box (size = half extents) | sphere (size[0] = radius) | capsule (radius, half length) */
  int32_t pair_body[ZB_NBOX];     /* 0: geom 1 is the floor plane; else the body of geom 1                                        */
  int32_t pair_type[ZB_NBOX];     /* ZB_GEOM_SPHERE / ZB_GEOM_CAPSULE (unused for the plane)                                      */
  float pair_pos[ZB_NBOX][3];
  float pair_quat[ZB_NBOX][4];
  float pair_size[ZB_NBOX][3];
  float pair_friction[ZB_NBOX][3];
  float pair_solref[ZB_NBOX][2];
  float pair_solimp[ZB_NBOX][5];
  int32_t pair_priority[ZB_NBOX];
  float pair_margin[ZB_NBOX];
  float pair_gap[ZB_NBOX];
  /* sites: 0 = imu_site, 1 = left_foot, 2 = right_foot (box-shaped touch volumes) */
  int32_t site_body[ZB_NSITE];
  float site_pos[ZB_NSITE][3];
  float site_quat[ZB_NSITE][4];
  float site_size[ZB_NSITE][3];
  /* statistics */
  float meaninertia;
  int32_t reserved;
} ZbModel;

/* Per-environment model overrides == the fields ksim's randomizers batch
 * (train.py:1439-1455).  Any pointer may be NULL -> the ZbModel value is used. */
typedef struct ZbEnvParams {
  const float* body_mass;        /* [N, ZB_NBODY]    AllBodiesMassMultiplicationRandomizer */
  const float* body_inertia;     /* [N, ZB_NBODY, 3]                                        */
  const float* dof_armature;     /* [N, ZB_NV]       ArmatureRandomizer                     */
  const float* dof_damping;      /* [N, ZB_NV]       JointDampingRandomizer                 */
  const float* dof_frictionloss; /* [N, ZB_NV]       StaticFrictionRandomizer               */
  const float* qpos0;            /* [N, ZB_NQ]       JointZeroPositionRandomizer            */
  const float* floor_friction;   /* [N]              FloorFrictionRandomizer (sliding mu)   */
  const float* imu_site_pos;     /* [N, 3]           IMUAlignmentRandomizer                 */
  const float* imu_site_quat;    /* [N, 4]                                                  */
  /* MjxEngine.step action latency (ksim `action_latency_range`, commented out at train.py:1786):
   * substeps s < latency_steps[i] still track prev_action, later ones the new action.  NULL = no latency. */
  const float* prev_action;      /* [N, ZB_NU]       PhysicsState.most_recent_action        */
  const int32_t* latency_steps;  /* [N]              round(action_latency / dt)             */
} ZbEnvParams;

/* ksim PhysicsState, flattened (data.qpos/qvel/qacc_warmstart/time, actuator_state ==
 * PlannerState train.py:1103-1108, event_states == PushEvent time remaining). In/out. */
typedef struct ZbState {
  float* qpos;            /* [N, ZB_NQ] */
  float* qvel;            /* [N, ZB_NV] */
  float* qacc_warmstart;  /* [N, ZB_NV] */
  float* act_pos;         /* [N, ZB_NU] PlannerState.position             */
  float* act_vel;         /* [N, ZB_NU] PlannerState.velocity             */
  float* act_torque;      /* [N, ZB_NU] PlannerState.last_computed_torque */
  float* push_t_rem;      /* [N]        PushEvent state                   */
  float* time;            /* [N]                                          */
} ZbState;

/* mjx.Data fields the task reads (train.py:452-461, 476, 681, 854, 882, 1252-1253, 1307),
 * produced by forward() of the LAST substep, i.e. at the pre-integration state exactly as
 * mjx.step leaves them.  Any pointer may be NULL (that output is skipped). */
typedef struct ZbOut {
  float* xpos;            /* [N, ZB_NBODY, 3]  */
  float* xquat;           /* [N, ZB_NBODY, 4]  */
  float* cinert;          /* [N, ZB_NBODY, 10] */
  float* cvel;            /* [N, ZB_NBODY, 6]  */
  float* qacc;            /* [N, ZB_NV]        */
  float* qfrc_constraint; /* [N, ZB_NV]        */
  float* actuator_force;  /* [N, ZB_NU]        */
  float* ctrl;            /* [N, ZB_NU]        */
  float* sensordata;      /* [N, ZB_NSENSORDATA] */
  float* site_xpos;       /* [N, ZB_NSITE, 3]  */
  float* site_xmat;       /* [N, ZB_NSITE, 9]  */
  float* contact_dist;    /* [N, ZB_MAXCON]    */
  float* contact_pos;     /* [N, ZB_MAXCON, 3] */
  float* efc_force;       /* [N, ZB_NV*2 + ZB_NEFC_CON]  friction | limit | contact rows */
  int32_t* contact_mask;  /* [N] bit c set <=> contact slot c active (dist < margin)    */
  int32_t* solver_niter;  /* [N]                                                         */
  int32_t* flags;         /* [N] bit0: non-finite state detected                         */
  float* solver_trace;    /* [N, ZB_NTRACE] diagnostics of the LAST substep: per solver iteration i (alpha, cost, ls_iters, |search|) at 4i;
                           * from ZB_TRACE_POINTS on the (alpha, d0, d1) of every point the FIRST line search evaluates (p0, newton(p0), then
                           * lo_next / hi_next / mid per bracket iteration); from ZB_TRACE_SEARCH on the first search direction by dof;
                           * [ZB_TRACE_START] the start point the solve picked */
} ZbOut;

/* Observation epilogue: the two network input vectors of the reference, assembled by the step /
 * reset kernel from the values it already holds in registers, so that the rollout does not re-read cinert / cvel / ... from
 * HBM and runs no per-observation fusions.  All pointers may be NULL (that part is skipped).
 *
 *   actor_obs  [N, 50]  = run_actor's concatenation, /root/reference/train.py:1624-1648:
 *       joint_pos(20) | joint_vel(20) | imu_orientation(4) | cmd[0:2] | cmd[3] | cmd[4:7]
 *   critic_obs [N, 484] = run_critic's concatenation, train.py:1650-1681:
 *       joint_pos(20) | joint_vel/10 (20) | cinert[1:] (250) | cvel[1:] (150) | imu_acc(3) | imu_gyro(3) |
 *       imu_orientation(4) | cmd(7) | actuator_force/100 (20) | base_pos(3) | base_quat(4)
 *   with joint_pos = qpos[7:], joint_vel = qvel[6:], base_pos/quat = qpos[0:3]/[3:7] of the NEW state and the mjx.Data
 *   fields (cinert, cvel, sensordata, actuator_force) exactly as ZbOut reports them (ksim's *Observation classes, [R]).
 *
 *   imu_orientation = ImuOrientationObservation.observe_stateful, train.py:837-866: the framequat sensor, spun back by the
 *   commanded heading cmd[3] (rotate_quat_by_quat(..., inverse=True), train.py:751-787), flipped to the w >= 0 hemisphere,
 *   then x <- x*lag + q*(1-lag) with the carry (x, lag) of train.py:846-848.  The carry lives in imu_x / imu_lag (in/out);
 *   environments re-initialised by the launch (zb_reset_*, the auto-reset pass) restart from x = 0 and
 *   lag = reset_imu_lag[i] (host-sampled U(lag_range), train.py:1497) before their first observation.
 *   noise_*: pre-scaled additive observation noise (ksim adds it after observe(); imu_orientation 1 deg and imu_acc 0.5 at
 *   train.py:1498, 1503-1505 are the only non-zero ones among the network inputs); the carry stays noise-free. */
typedef struct ZbObs {
  const float* command;        /* [N, ZB_NCOMMAND]; NULL = zeros (ConstantZeroCommand, train.py:416-427) */
  float* imu_x;                /* [N, 4]  filter state, in/out (required for imu_orientation)             */
  float* imu_lag;              /* [N]     EMA factor, in/out                                              */
  const float* reset_imu_lag;  /* [N]     lag of re-initialised environments; NULL = keep imu_lag         */
  const float* noise_imu_quat; /* [N, 4]  NULL = none                                                     */
  const float* noise_imu_acc;  /* [N, 3]  NULL = none                                                     */
  float* actor_obs;            /* [N, ZB_NACTOR_OBS]                                                      */
  float* critic_obs;           /* [N, ZB_NCRITIC_OBS]                                                     */
  /* Reward inputs (first part): per-step quantities of the reference's OWN reward / observation
   * classes, raw (before ksim's scale / curriculum factors), as functions of this launch's mjx.Data:
   *   feet_pos_obs [N, 6]  FeetPositionObservation, /root/reference/train.py:430-465 (in_robot_frame, floor_threshold 0):
   *                        left_foot | right_foot site positions rotated by the inverse base quaternion qpos[3:7]
   *   reward_terms [N, ZB_NREWARD_TERMS]:
   *     [0] SimpleSingleFootContactReward (stand_still_threshold=None), train.py:691-713: xor(touch_l > 0.1, touch_r > 0.1)
   *     [1] FeetOrientationReward(target_rp=(0,0), error_scale=0.25), train.py:647-689, on the bodies of the two foot sites
   *     [2] FeetTooClosePenalty(threshold_m=0.12), train.py:729-748, on feet_pos_obs
   *     [3] BaseHeightReward(error_scale=0.25, standard_height=0.27), train.py:468-483: exp(-|xpos[1].z - (cmd[4] + 0.27)| / 0.25)
   *     [4], [5] left / right contact flags (touch > 0.1): what FeetAirtimeReward's scan consumes, train.py:524-525 */
  float* feet_pos_obs;         /* [N, 6]                                                                  */
  float* reward_terms;         /* [N, ZB_NREWARD_TERMS]                                                   */
} ZbObs;

/* Arguments of the auto-reset pass (see zb_step_autoreset_f32). */
typedef struct ZbAutoResetArgs {
  const float* reset_qpos;  /* [N, ZB_NQ] */
  const float* reset_qvel;  /* [N, ZB_NV] */
  const float* reset_t_rem; /* [N] or NULL */
  uint8_t* done_out;        /* [N] or NULL: 0 = running, 1 = terminated (bad z / not upright / non-finite), 2 = episode length reached */
  float z_lo, z_hi, max_tilt_rad;
  float max_episode_sec;    /* ksim.EpisodeLengthTermination(max_length_sec), train.py:1592: time > max ends the episode; <= 0 disables */
  /* Trajectory record: what ksim's rollout stores in its Trajectory for this control step —
   * qpos / qvel / xpos / xquat of `next_physics_state` BEFORE the done -> reset select, i.e. the terminal state of an environment
   * that ends here (train.py:297, 476, 681 read exactly these four).  Written for EVERY environment just before the termination
   * test; pointing them at slice t of time-major [T, N, ...] arrays makes the step launch the trajectory writer.  Each may be NULL. */
  float* term_qpos;         /* [N, ZB_NQ]        */
  float* term_qvel;         /* [N, ZB_NV]        */
  float* term_xpos;         /* [N, ZB_NBODY, 3]  */
  float* term_xquat;        /* [N, ZB_NBODY, 4]  */
} ZbAutoResetArgs;

typedef struct ZbEngine ZbEngine; /* opaque: device copy of ZbModel + ZbOpts */

/* Engine construction == mjx.put_model + ksim.get_physics_engine(...) for this path. */
int zb_engine_create(const ZbModel* model, const ZbOpts* opts, int device, ZbEngine** out);
int zb_engine_destroy(ZbEngine* engine);
int zb_engine_set_opts(ZbEngine* engine, const ZbOpts* opts);

/* MjxEngine.step: n_substeps x { PushEvent -> FeetechActuators.get_stateful_ctrl -> mjx.step }.
 *   action    [N, ZB_NU]   position targets (train.py:1252-1260)
 *   push_rand [N, 7] (push_per_substep == 0) or [N, n_substeps, 7]: pre-sampled
 *             (dqvel[6], next interval) consumed when the push timer expires; NULL = no PushEvent. */
int zb_step_f32(ZbEngine* engine, void* cuda_stream, const ZbEnvParams* params, const ZbState* state,
                const float* action, const float* push_rand, int push_per_substep, const ZbOut* out,
                int n_envs, int n_substeps);

/* zb_step_f32 followed, inside the same launch, by the rollout's termination check and partial reset
 * (ksim RLTask._single_unroll: `done -> engine.reset`; BadZTermination / NotUprightTermination / EpisodeLengthTermination of
 * /root/reference/train.py:1588-1593): for every environment whose integrated state terminates
 * (z outside [z_lo, z_hi], base tilt above max_tilt_rad, non-finite, or time > max_episode_sec when that is > 0), the state is replaced by
 * reset_qpos/reset_qvel/reset_t_rem (host-sampled per ksim resets, train.py:1471-1476), PlannerState re-initialised
 * (train.py:1293-1298), time/warm-start cleared and forward() run on it — exactly zb_terminations_f32 +
 * zb_reset_f32(done).  done_out [N] (may be NULL) receives the termination mask (0 running, 1 failure, 2 episode length — ksim's
 * terminations return -1 / +1 for failure / success, `done` is "non-zero"). */
int zb_step_autoreset_f32(ZbEngine* engine, void* cuda_stream, const ZbEnvParams* params, const ZbState* state,
                          const float* action, const float* push_rand, int push_per_substep, const ZbOut* out, int n_envs,
                          int n_substeps, const float* reset_qpos, const float* reset_qvel, const float* reset_t_rem,
                          uint8_t* done_out, float z_lo, float z_hi, float max_tilt_rad, float max_episode_sec);

/* zb_step_f32 / zb_step_autoreset_f32 (autoreset != NULL) with the observation epilogue: ONE launch per control step produces
 * the next state, the mjx.Data outputs, the termination mask, the partial reset and both network input vectors. */
int zb_step_obs_f32(ZbEngine* engine, void* cuda_stream, const ZbEnvParams* params, const ZbState* state, const float* action,
                    const float* push_rand, int push_per_substep, const ZbOut* out, const ZbObs* obs, int n_envs,
                    int n_substeps, const ZbAutoResetArgs* autoreset);

/* zb_reset_f32 with the observation epilogue (first observation of an episode; carry re-initialised for the reset envs). */
int zb_reset_obs_f32(ZbEngine* engine, void* cuda_stream, const ZbEnvParams* params, const ZbState* state,
                     const uint8_t* done, const float* reset_qpos, const float* reset_qvel, const float* reset_t_rem,
                     const ZbOut* out, const ZbObs* obs, int n_envs);

/* mjx.forward on the current state (no integration, state untouched except the free-joint
 * quaternion normalisation MJX writes back) — what MjxEngine.reset runs after the resets. */
int zb_forward_f32(ZbEngine* engine, void* cuda_stream, const ZbEnvParams* params, const ZbState* state,
                   const ZbOut* out, int n_envs);

/* MjxEngine.reset for the environments with done[i] != 0 (done == NULL: all):
 * qpos/qvel <- reset_qpos/reset_qvel (host-sampled per ksim resets, train.py:1471-1476),
 * PlannerState <- (qpos[7:], qvel[6:], 0) (train.py:1293-1298), push_t_rem <- reset_t_rem,
 * time <- 0, qacc_warmstart <- 0, then forward() for those environments. */
int zb_reset_f32(ZbEngine* engine, void* cuda_stream, const ZbEnvParams* params, const ZbState* state,
                 const uint8_t* done, const float* reset_qpos, const float* reset_qvel,
                 const float* reset_t_rem, const ZbOut* out, int n_envs);

/* Introspection for bindings and tests. */
int zb_abi_version(void);
int zb_sizeof_model(void);
int zb_sizeof_opts(void);
const char* zb_last_error(void);
/* The CTA shape a step launch of n_envs environments gets on a GPU with num_sms SMs (host logic, no device needed):
 * *wide = 1 for the 28-warp / 72-register throughput profile, 0 for the 16-warp / 128-register latency profile. */
int zb_cta_shape(int n_envs, int num_sms, int* wide, int* warps_per_cta);
/* Number of kernel launches issued through this library since load (bench accounting). */
long long zb_launch_count(void);

/* ksim.BadZTermination(z_lo, z_hi) | ksim.NotUprightTermination(max_tilt) | ksim.EpisodeLengthTermination(max_episode_sec) as the
 * task configures them (/root/reference/train.py:1588-1593) plus the non-finite guard (flags from ZbOut, may be NULL):
 * done[i] = 1 when environment i fails, 2 when only its episode length ran out (time [N] may be NULL / max_episode_sec <= 0: no
 * episode-length check), 0 otherwise.  Feeds zb_reset_f32's done mask without a host round trip. */
int zb_terminations_f32(void* cuda_stream, const float* qpos, const int32_t* flags, const float* time, uint8_t* done, int n_envs,
                        float z_lo, float z_hi, float max_tilt_rad, float max_episode_sec);

/* FeetAirtimeReward.get_reward_stateful over one rollout chunk (/root/reference/train.py:503-546), raw
 * (before ksim's `scale`), with stand_still_threshold = None as the task configures it (train.py:1560-1565):
 *   airtime scan   t <- (contact | done) ? 0 : t + ctrl_dt, per foot, starting from carry[i] (train.py:503-516)
 *   touchdown      contact[t] & !contact[t-1], with contact[-1] := false at the start of the chunk (train.py:528-533)
 *   reward[t]      sum over the feet of (airtime[t-1] - touchdown_penalty) * touchdown[t], where airtime[-1] is airtime[T-1] of
 *                  the SAME chunk (the reference shifts with jnp.roll, train.py:535-536 — kept as it is)
 * left_contact / right_contact / done / reward are [T, N]-shaped views given by strides in elements (element (t, i) at
 * t * time_stride + i * env_stride), contacts as floats (> 0.5 = contact, e.g. ZbObs.reward_terms[:, 4] / [:, 5] stacked over
 * the rollout), done as uint8.  carry [N, 2] (left, right) is updated in place.  One thread per environment. */
int zb_feet_airtime_f32(void* cuda_stream, const float* left_contact, const float* right_contact, long long contact_time_stride,
                        long long contact_env_stride, const uint8_t* done, long long done_time_stride, long long done_env_stride,
                        float* carry, float* reward, long long reward_time_stride, long long reward_env_stride, int n_envs,
                        int n_steps, float ctrl_dt, float touchdown_penalty);

/* ------------------------------------------------------------------------------------------------------------------------------
 * Policy inference in the rollout loop: the reference's Actor (/root/reference/train.py:885-968) evaluated for
 * every environment of a control step by one kernel — input_proj Linear(n_in, 128) -> depth x GRUCell(128, 128) -> output_proj
 * Linear(128, 20 * 3 * num_mixtures) -> mixture-of-Gaussians head (mean + JOINT_BIASES, std = clip((softplus + min_std) * var_scale,
 * max_std), logits) -> action.  It is what run_actor / sample_action (train.py:1624-1648, 1737-1763) do per control step, reading the
 * [N, 50] vector the observation epilogue wrote (ZbObs.actor_obs).
 * Weights are device pointers in kernel layout (k-major; ksim-gym-zbot_b200/policy.py packs them from the equinox layout):
 *   in_wt     [n_in][128]            = input_proj.weight[f][k] transposed          in_b  [128]
 *   gru_wih_t [depth][128][3][128]   = rnns[l].weight_ih[g * 128 + f][k]   (gate order r, z, n as in equinox / PyTorch)
 *   gru_whh_t [depth][128][3][128]   = rnns[l].weight_hh[g * 128 + f][k]
 *   gru_b     [depth][3][128]        = rnns[l].bias        gru_bn [depth][128] = rnns[l].bias_n
 *   out_wt    [128][n_out]           = output_proj.weight[j][k] transposed         out_b [n_out],  n_out = 20 * 3 * num_mixtures
 *   mean_bias [20]                   = JOINT_BIASES (train.py:59-82, added at train.py:963), may be NULL */
#define ZB_ACTOR_HIDDEN 128
#define ZB_ACTOR_MAX_IN 128
#define ZB_ACTOR_MAX_MIXTURES 5
typedef struct ZbActor {
  int32_t n_in;          /* NUM_ACTOR_INPUTS = 50 (train.py:53) */
  int32_t depth;         /* GRU layers (train.py:1069-1072: 5) */
  int32_t num_mixtures;  /* train.py:1073-1076: 5 */
  float min_std, max_std, var_scale; /* train.py:1609-1610: 0.01, 1.0; Model passes var_scale = 1.0 (train.py:1046) */
  int32_t precision;     /* ZB_ACTOR_FP32: GRU stack on the CUDA cores in fp32; ZB_ACTOR_TF32X3: on the tensor cores (tcgen05, 3xTF32) */
  int32_t reserved;
  const float *in_wt, *in_b, *gru_wih_t, *gru_whh_t, *gru_b, *gru_bn, *out_wt, *out_b, *mean_bias;
  const float* gru_w_tc; /* ZB_ACTOR_TF32X3 only: [depth][16 k-tiles][6 matrices: W_ih r,z,n, W_hh r,z,n][2 k-chunks][128 f][4 k] =
                            W_m[f][8 t + 4 c + j] — the tcgen05 K-major no-swizzle operand layout, one 24 KB TMA bulk copy per k-tile */
  const float* out_w_tc; /* ZB_ACTOR_TF32X3 only: output_proj in the same layout, [16 k-tiles][3 m-tiles][2 k-chunks][128 rows][4 k] =
                            output_proj.weight[128 mt + row][8 t + 4 c + j], rows beyond n_out zero (12 KB per k-tile) */
} ZbActor;
#define ZB_ACTOR_FP32 0
#define ZB_ACTOR_TF32X3 1
/* One control step of the actor for n_envs environments.
 *   actor_obs   [N, n_in]          in
 *   carry_in    [N, depth, 128]    in   the actor half of the model carry (train.py:1731-1735); carry_out may alias it (in place)
 *   carry_out   [N, depth, 128]    out
 *   carry_reset [N] or NULL        in   non-zero: this environment starts from the initial carry (zeros) — the `done` of the previous
 *                                       control step, as the rollout re-initialises the carry of finished episodes [R]
 *   noise_u, noise_eps [N, 20] or NULL  pre-sampled U[0,1) / N(0,1): mixture component by inverse CDF of softmax(logits) at u, action =
 *                                       mean_k + std_k * eps  (`sample`); both NULL: the mean of the most probable component (`mode`,
 *                                       sample_action(argmax=True), train.py:1757).  The reference draws them from a JAX PRNG key.
 *   action      [N, 20]            out
 *   dist        [N, 3, 20, M] or NULL  out: means (incl. JOINT_BIASES), stds, logits as handed to ksim.MixtureOfGaussians (train.py:965)
 * One CTA per 32 environments.  precision = ZB_ACTOR_FP32: everything in fp32 on the CUDA cores (matches a CPU float32 evaluation to
 * rounding).  ZB_ACTOR_TF32X3: the GRU stack (92 % of the FLOPs) on the tcgen05 tensor cores with every product issued as
 * hi.hi + hi.lo + lo.hi of TF32 halves (fp32-level accuracy, ~1e-6 relative), accumulators in TMEM — the GRU stack and output_proj;
 * needs gru_w_tc, out_w_tc and n_in <= 96.
 * Errors: ZB_EINVAL (NULL / sizes / alignment), ZB_EUNSUP (limits above), ZB_ECUDA. */
int zb_actor_f32(void* cuda_stream, const ZbActor* actor, const float* actor_obs, const float* carry_in, float* carry_out,
                 const uint8_t* carry_reset, const float* noise_u, const float* noise_eps, float* action, float* dist, int n_envs);
int zb_sizeof_actor(void);

/* FP32 FFMA throughput probe (bench.py: the FP32 roofline denominator).  Enqueues grid x 256 threads doing
 * iters x 16 independent FMAs each; scratch needs grid * 256 floats.  flops = grid * 256 * iters * 32. */
int zb_bench_ffma(void* cuda_stream, float* scratch, int grid, int iters);

#ifdef __cplusplus
}
#endif
#endif /* ZBPHYS_H_ */
