/*
 * zb_oracle.h — data layout of the CPU oracle (TEST INFRASTRUCTURE ONLY, see zb_oracle.c).
 * `real` is float (-DZBO_F64=0) or double (-DZBO_F64=1); oracle/zb_oracle.py mirrors these
 * structs with ctypes for both precisions.
 */
#ifndef ZB_ORACLE_H_
#define ZB_ORACLE_H_
#include <stdint.h>

#include "../include/zbphys.h"

#ifndef ZBO_F64
#define ZBO_F64 0
#endif
#if ZBO_F64
typedef double real;
#else
typedef float real;
#endif

#define ZBO_NEFC (2 * ZB_NV + ZB_NEFC_CON) /* friction | limit | contact rows */

enum {
  ZBO_ST_KINEMATICS = 0, ZBO_ST_COM_POS, ZBO_ST_CRB, ZBO_ST_FACTOR, ZBO_ST_COLLISION, ZBO_ST_CONSTRAINT,
  ZBO_ST_COM_VEL, ZBO_ST_RNE, ZBO_ST_ACTUATION, ZBO_ST_ACCEL, ZBO_ST_SOLVER, ZBO_ST_SENSOR, ZBO_ST_INTEGRATE,
  ZBO_NSTAGE
};

/* host mirrors of ZbEnvParams / ZbState / ZbOut with `real` arrays */
typedef struct ZboParams {
  const real *body_mass, *body_inertia, *dof_armature, *dof_damping, *dof_frictionloss, *qpos0, *floor_friction,
      *imu_site_pos, *imu_site_quat, *prev_action;
  const int32_t* latency_steps;
} ZboParams;
typedef struct ZboState {
  real *qpos, *qvel, *qacc_warmstart, *act_pos, *act_vel, *act_torque, *push_t_rem, *time;
} ZboState;
typedef struct ZboOut {
  real *xpos, *xquat, *cinert, *cvel, *qacc, *qfrc_constraint, *actuator_force, *ctrl, *sensordata, *site_xpos,
      *site_xmat, *contact_dist, *contact_pos, *efc_force;
  int32_t *contact_mask, *solver_niter, *flags;
  real* solver_trace;
} ZboOut;

/* everything mjx.forward leaves in mjx.Data for one environment (plus a few debugging extras) */
typedef struct ZboData {
  real qpos[ZB_NQ], qvel[ZB_NV], ctrl[ZB_NU], qacc_warmstart[ZB_NV];
  real xpos[ZB_NBODY][3], xquat[ZB_NBODY][4], xmat[ZB_NBODY][9], xipos[ZB_NBODY][3], ximat[ZB_NBODY][9];
  real xanchor[ZB_NBODY][3], xaxis[ZB_NBODY][3], subtree_com[ZB_NBODY][3];
  real cinert[ZB_NBODY][10], crb[ZB_NBODY][10], cdof[ZB_NV][6], cdof_dot[ZB_NV][6], cvel[ZB_NBODY][6];
  real qM[ZB_NV][ZB_NV], qL[ZB_NV][ZB_NV];
  real qfrc_passive[ZB_NV], qfrc_bias[ZB_NV], qfrc_actuator[ZB_NV], qfrc_smooth[ZB_NV], qacc_smooth[ZB_NV];
  real actuator_force[ZB_NU];
  real site_xpos[ZB_NSITE][3], site_xmat[ZB_NSITE][9], box_xpos[ZB_NBOX][3], box_xmat[ZB_NBOX][9], pair_xpos[ZB_NBOX][3], pair_xmat[ZB_NBOX][9];
  real con_dist[ZB_MAXCON], con_pos[ZB_MAXCON][3], con_frame[ZB_MAXCON][9], con_mu[ZB_MAXCON];
  real con_includemargin[ZB_MAXCON], con_solref[ZB_MAXCON][2], con_solimp[ZB_MAXCON][5];
  real efc_J[ZBO_NEFC][ZB_NV], efc_pos[ZBO_NEFC], efc_D[ZBO_NEFC], efc_R[ZBO_NEFC], efc_aref[ZBO_NEFC];
  real efc_floss[ZBO_NEFC], efc_force[ZBO_NEFC];
  real qfrc_constraint[ZB_NV], qacc[ZB_NV];
  real cacc[ZB_NBODY][6], cfrc_int[ZB_NBODY][6], cfrc_ext[ZB_NBODY][6];
  real sensordata[ZB_NSENSORDATA];
  real solver_cost;
  real trace[ZB_NTRACE];
  int32_t con_body[ZB_MAXCON], con_body1[ZB_MAXCON], con_vert[ZB_MAXCON], con_active[ZB_MAXCON]; /* con_body = body of geom 2, con_body1 = body of geom 1 (0 = world) */
  int32_t efc_type[ZBO_NEFC]; /* 0 inert, 1 friction-loss, 2 limit, 3 contact */
  int32_t solver_niter, ls_evals, ls_iters, flags;
} ZboData;

#endif
