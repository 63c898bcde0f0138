/*
 * zb_oracle.c — CPU restatement of the Z-Bot physics step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product path (ksim-gym-zbot_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED: the algorithm this file restates lives in un-vendored third-party packages
 * that are absent from /root/reference and from this image:
 *   mujoco-mjx 3.3.4  (/root/reference/uv.lock:2226-2227)  — mjx.step / mjx.forward
 *   ksim 0.1.99       (/root/reference/uv.lock:1872-1873)  — MjxEngine.step / reset, PushEvent
 * and the reference ships no tests, golden vectors or fixtures for this path.
 * Every function below names the upstream function it restates ("[R]" = recalled published
 * algorithm, no source here) or the reference line it follows ("[V]" = /root/reference file:line).
 * What pins it instead: analytic known-answer tests and dynamics identities in tests/test_oracle.py.
 *
 * Compiled twice by oracle/Makefile: -DZBO_F64=0 (fp32, JAX's default dtype — the parity oracle
 * and CPU baseline) and -DZBO_F64=1 (fp64, the accuracy yard-stick).  Scalar, one environment
 * at a time; the batch entry points loop over environments with OpenMP.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/zbphys.h"
#include "zb_oracle.h"

#if ZBO_F64
#define R_SQRT sqrt
#define R_SIN sin
#define R_COS cos
#define R_POW pow
#define R_ABS fabs
#define R_FMAX fmax
#define R_FMIN fmin
#else
#define R_SQRT sqrtf
#define R_SIN sinf
#define R_COS cosf
#define R_POW powf
#define R_ABS fabsf
#define R_FMAX fmaxf
#define R_FMIN fminf
#endif

#define NQ ZB_NQ
#define NV ZB_NV
#define NU ZB_NU
#define NB ZB_NBODY
#define NS ZB_NSITE
#define NBOX ZB_NBOX
#define MAXCON ZB_MAXCON
#define NEFC ZBO_NEFC

#define MJ_MINVAL ((real)1e-15)
#define MJ_MINIMP ((real)0.0001)
#define MJ_MAXIMP ((real)0.9999)

/* optional flop accounting ("authoritative figure is the oracle's op counters").
 * FMA = 2, div/sqrt/sincos/pow = 1.  Only the build with -DZBO_COUNT pays for it. */
#ifdef ZBO_COUNT
static double g_flops[ZBO_NSTAGE];
static int g_stage;
#pragma omp threadprivate(g_flops, g_stage)
#define FL(n) (g_flops[g_stage] += (n))
#define STAGE(s) (g_stage = (s))
#else
#define FL(n) ((void)0)
#define STAGE(s) ((void)0)
#endif

static int g_variant = 0;                    /* see zbo_set_variant */
static _Thread_local real g_w[ZBO_NEFC][6];   /* contact-row wrench directions [off x dir; dir] */
static _Thread_local const ZbModel* g_model;

/* ------------------------------------------------------------------------------------------
 * small math — mujoco/mjx/_src/math.py [R]
 * ---------------------------------------------------------------------------------------- */
static inline real dot3(const real* a, const real* b) { FL(5); return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(real* o, const real* a, const real* b) {
  FL(9);
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline real dot6(const real* a, const real* b) {
  FL(11);
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
/* math.quat_mul */
static inline void quat_mul(real* o, const real* u, const real* v) {
  FL(28);
  real w = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  real x = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  real y = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  real z = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
/* math.quat_to_mat (row-major 3x3) */
static inline void quat_to_mat(real* m, const real* q) {
  FL(30);
  real w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void mat_vec(real* o, const real* m, const real* v) {
  FL(15);
  real x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  real y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  real z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static inline void matT_vec(real* o, const real* m, const real* v) {
  FL(15);
  real x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  real y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  real z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
/* math.rotate(vec, quat) == quat_to_mat(quat) @ vec */
static inline void rotate(real* o, const real* v, const real* q) {
  real m[9];
  quat_to_mat(m, q);
  mat_vec(o, m, v);
}
static inline void normalize4(real* q) {
  FL(12);
  real n = R_SQRT(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  real s = (real)1 / (n + (real)1e-6 * (n == 0));
  q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
/* math.axis_angle_to_quat */
static inline void axis_angle_to_quat(real* q, const real* axis, real angle) {
  FL(7);
  real s = R_SIN(angle * (real)0.5), c = R_COS(angle * (real)0.5);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* math.inert_mul: cinert (10) times motion vector (6) -> force vector (6) */
static inline void inert_mul(real* o, const real* I, const real* v) {
  FL(15 + 18 + 3 + 3 + 3);
  const real* pos = I + 6;
  real mass = I[9];
  real a0 = I[0] * v[0] + I[3] * v[1] + I[4] * v[2];
  real a1 = I[3] * v[0] + I[1] * v[1] + I[5] * v[2];
  real a2 = I[4] * v[0] + I[5] * v[1] + I[2] * v[2];
  real c[3], e[3];
  cross3(c, pos, v + 3); FL(-9);
  cross3(e, pos, v); FL(-9);
  o[0] = a0 + c[0]; o[1] = a1 + c[1]; o[2] = a2 + c[2];
  o[3] = mass * v[3] - e[0]; o[4] = mass * v[4] - e[1]; o[5] = mass * v[5] - e[2];
}
/* math.motion_cross(u, v) */
static inline void motion_cross(real* o, const real* u, const real* v) {
  real a[3], b[3], c[3];
  cross3(a, u, v);
  cross3(b, u, v + 3);
  cross3(c, u + 3, v);
  FL(3);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
/* math.motion_cross_force(v, f) */
static inline void motion_cross_force(real* o, const real* v, const real* f) {
  real a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
  FL(3);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}

/* ------------------------------------------------------------------------------------------
 * per-environment model: ZbModel + the randomizer overrides of train.py:1439-1455 [V]
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  real body_mass[NB], body_inertia[NB][3];
  real dof_armature[NV], dof_damping[NV], dof_frictionloss[NV];
  real qpos0[NQ];
  real floor_mu;
  real imu_pos[3], imu_quat[4];
} EnvModel;

static void env_model(EnvModel* e, const ZbModel* m, const ZboParams* p, int i) {
  for (int b = 0; b < NB; b++) {
    e->body_mass[b] = (p && p->body_mass) ? p->body_mass[i * NB + b] : (real)m->body_mass[b];
    for (int k = 0; k < 3; k++)
      e->body_inertia[b][k] = (p && p->body_inertia) ? p->body_inertia[(i * NB + b) * 3 + k] : (real)m->body_inertia[b][k];
  }
  for (int d = 0; d < NV; d++) {
    e->dof_armature[d] = (p && p->dof_armature) ? p->dof_armature[i * NV + d] : (real)m->dof_armature[d];
    e->dof_damping[d] = (p && p->dof_damping) ? p->dof_damping[i * NV + d] : (real)m->dof_damping[d];
    e->dof_frictionloss[d] = (p && p->dof_frictionloss) ? p->dof_frictionloss[i * NV + d] : (real)m->dof_frictionloss[d];
  }
  for (int k = 0; k < NQ; k++) e->qpos0[k] = (p && p->qpos0) ? p->qpos0[i * NQ + k] : (real)m->qpos0[k];
  e->floor_mu = (p && p->floor_friction) ? p->floor_friction[i] : (real)m->floor_friction[0];
  for (int k = 0; k < 3; k++) e->imu_pos[k] = (p && p->imu_site_pos) ? p->imu_site_pos[i * 3 + k] : (real)m->site_pos[0][k];
  for (int k = 0; k < 4; k++) e->imu_quat[k] = (p && p->imu_site_quat) ? p->imu_site_quat[i * 4 + k] : (real)m->site_quat[0][k];
}

/* ------------------------------------------------------------------------------------------
 * smooth.kinematics [R]
 * ---------------------------------------------------------------------------------------- */
static void kinematics(const ZbModel* m, const EnvModel* e, ZboData* d) {
  STAGE(ZBO_ST_KINEMATICS);
  memset(d->xpos[0], 0, sizeof(real) * 3);
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  for (int b = 1; b < NB; b++) {
    int p = m->body_parent[b];
    real pos[3], quat[4];
    if (m->body_jnttype[b] == 1) { /* free joint: pose straight from qpos, quaternion normalised in place */
      normalize4(d->qpos + 3);
      for (int k = 0; k < 3; k++) pos[k] = d->qpos[k];
      for (int k = 0; k < 4; k++) quat[k] = d->qpos[3 + k];
      for (int k = 0; k < 3; k++) d->xanchor[b][k] = pos[k];
      d->xaxis[b][0] = 0; d->xaxis[b][1] = 0; d->xaxis[b][2] = 1;
    } else {
      real bp[3] = {(real)m->body_pos[b][0], (real)m->body_pos[b][1], (real)m->body_pos[b][2]};
      real bq[4] = {(real)m->body_quat[b][0], (real)m->body_quat[b][1], (real)m->body_quat[b][2], (real)m->body_quat[b][3]};
      real r[3];
      rotate(r, bp, d->xquat[p]);
      for (int k = 0; k < 3; k++) pos[k] = d->xpos[p][k] + r[k];
      FL(3);
      quat_mul(quat, d->xquat[p], bq);
      if (m->body_jnttype[b] == 2) {
        int dof = m->body_dofadr[b], qa = dof + 1;
        real jp[3] = {(real)m->jnt_pos[b][0], (real)m->jnt_pos[b][1], (real)m->jnt_pos[b][2]};
        real ax[3] = {(real)m->jnt_axis[b][0], (real)m->jnt_axis[b][1], (real)m->jnt_axis[b][2]};
        rotate(r, jp, quat);
        for (int k = 0; k < 3; k++) d->xanchor[b][k] = pos[k] + r[k];
        rotate(d->xaxis[b], ax, quat);
        real qa_[4], q2[4];
        axis_angle_to_quat(qa_, ax, d->qpos[qa] - e->qpos0[qa]);
        quat_mul(q2, quat, qa_);
        memcpy(quat, q2, sizeof(q2));
        rotate(r, jp, quat);
        for (int k = 0; k < 3; k++) pos[k] = d->xanchor[b][k] - r[k];
        FL(7);
      } else {
        for (int k = 0; k < 3; k++) { d->xanchor[b][k] = 0; d->xaxis[b][k] = 0; }
      }
      normalize4(quat);
    }
    memcpy(d->xpos[b], pos, sizeof(pos));
    memcpy(d->xquat[b], quat, sizeof(quat));
  }
  for (int b = 0; b < NB; b++) {
    quat_to_mat(d->xmat[b], d->xquat[b]);
    real ip[3] = {(real)m->body_ipos[b][0], (real)m->body_ipos[b][1], (real)m->body_ipos[b][2]};
    real iq[4] = {(real)m->body_iquat[b][0], (real)m->body_iquat[b][1], (real)m->body_iquat[b][2], (real)m->body_iquat[b][3]};
    real r[3], q[4];
    mat_vec(r, d->xmat[b], ip);
    for (int k = 0; k < 3; k++) d->xipos[b][k] = d->xpos[b][k] + r[k];
    FL(3);
    if (b == 0) { iq[0] = 1; iq[1] = iq[2] = iq[3] = 0; }
    quat_mul(q, d->xquat[b], iq);
    quat_to_mat(d->ximat[b], q);
  }
  /* sites (imu_site pose is per-env: IMUAlignmentRandomizer train.py:1452-1454 [V]) and foot boxes */
  for (int s = 0; s < NS; s++) {
    int b = m->site_body[s];
    real sp[3], sq[4], r[3], q[4];
    for (int k = 0; k < 3; k++) sp[k] = s == 0 ? e->imu_pos[k] : (real)m->site_pos[s][k];
    for (int k = 0; k < 4; k++) sq[k] = s == 0 ? e->imu_quat[k] : (real)m->site_quat[s][k];
    mat_vec(r, d->xmat[b], sp);
    for (int k = 0; k < 3; k++) d->site_xpos[s][k] = d->xpos[b][k] + r[k];
    FL(3);
    quat_mul(q, d->xquat[b], sq);
    quat_to_mat(d->site_xmat[s], q);
  }
  for (int g = 0; g < NBOX; g++) {
    int b = m->box_body[g];
    real gp[3] = {(real)m->box_pos[g][0], (real)m->box_pos[g][1], (real)m->box_pos[g][2]};
    real gq[4] = {(real)m->box_quat[g][0], (real)m->box_quat[g][1], (real)m->box_quat[g][2], (real)m->box_quat[g][3]};
    real r[3], q[4];
    mat_vec(r, d->xmat[b], gp);
    for (int k = 0; k < 3; k++) d->box_xpos[g][k] = d->xpos[b][k] + r[k];
    FL(3);
    quat_mul(q, d->xquat[b], gq);
    quat_to_mat(d->box_xmat[g], q);
    if (m->pair_body[g] > 0) { /* geom 1 of a body-body pair */
      int b1 = m->pair_body[g];
      real pp[3] = {(real)m->pair_pos[g][0], (real)m->pair_pos[g][1], (real)m->pair_pos[g][2]};
      real pq[4] = {(real)m->pair_quat[g][0], (real)m->pair_quat[g][1], (real)m->pair_quat[g][2], (real)m->pair_quat[g][3]};
      mat_vec(r, d->xmat[b1], pp);
      for (int k = 0; k < 3; k++) d->pair_xpos[g][k] = d->xpos[b1][k] + r[k];
      quat_mul(q, d->xquat[b1], pq);
      quat_to_mat(d->pair_xmat[g], q);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * smooth.com_pos [R]
 * ---------------------------------------------------------------------------------------- */
static void com_pos(const ZbModel* m, const EnvModel* e, ZboData* d) {
  STAGE(ZBO_ST_COM_POS);
  real pos[NB][3], mass[NB];
  for (int b = 0; b < NB; b++) {
    mass[b] = e->body_mass[b];
    for (int k = 0; k < 3; k++) pos[b][k] = d->xipos[b][k] * mass[b];
    FL(3);
  }
  for (int b = NB - 1; b > 0; b--) { /* leaf -> root */
    int p = m->body_parent[b];
    mass[p] += mass[b];
    for (int k = 0; k < 3; k++) pos[p][k] += pos[b][k];
    FL(4);
  }
  for (int b = 0; b < NB; b++)
    for (int k = 0; k < 3; k++) {
      d->subtree_com[b][k] = mass[b] < MJ_MINVAL ? d->xipos[b][k] : pos[b][k] / mass[b];
      FL(1);
    }
  for (int b = 0; b < NB; b++) {
    int root = b == 0 ? 0 : 1; /* body_rootid: the single floating-base tree */
    real off[3];
    for (int k = 0; k < 3; k++) off[k] = d->xipos[b][k] - d->subtree_com[root][k];
    FL(3);
    /* inert = ximat diag(I) ximat^T + mass * (|off|^2 1 - off off^T) */
    const real* R = d->ximat[b];
    const real* I = e->body_inertia[b];
    real in[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = r; c < 3; c++) {
        in[r][c] = R[r * 3 + 0] * I[0] * R[c * 3 + 0] + R[r * 3 + 1] * I[1] * R[c * 3 + 1] + R[r * 3 + 2] * I[2] * R[c * 3 + 2];
        FL(8);
      }
    real ms = e->body_mass[b], o2 = dot3(off, off);
    in[0][0] += ms * (o2 - off[0] * off[0]); in[1][1] += ms * (o2 - off[1] * off[1]); in[2][2] += ms * (o2 - off[2] * off[2]);
    in[0][1] -= ms * off[0] * off[1]; in[0][2] -= ms * off[0] * off[2]; in[1][2] -= ms * off[1] * off[2];
    FL(21);
    real* ci = d->cinert[b];
    ci[0] = in[0][0]; ci[1] = in[1][1]; ci[2] = in[2][2]; ci[3] = in[0][1]; ci[4] = in[0][2]; ci[5] = in[1][2];
    ci[6] = off[0] * ms; ci[7] = off[1] * ms; ci[8] = off[2] * ms; ci[9] = ms;
    FL(3);
  }
  /* cdof: motion dofs in the frame centred at the tree COM */
  const real* c0 = d->subtree_com[1];
  for (int b = 1; b < NB; b++) {
    int dof = m->body_dofadr[b];
    if (m->body_jnttype[b] == 1) {
      real off[3];
      for (int k = 0; k < 3; k++) off[k] = c0[k] - d->xanchor[b][k];
      FL(3);
      for (int i = 0; i < 3; i++) {
        for (int k = 0; k < 6; k++) d->cdof[dof + i][k] = (k == 3 + i);
        real a[3] = {d->xmat[b][i], d->xmat[b][3 + i], d->xmat[b][6 + i]}; /* column i of xmat */
        real c[3];
        cross3(c, a, off);
        for (int k = 0; k < 3; k++) { d->cdof[dof + 3 + i][k] = a[k]; d->cdof[dof + 3 + i][3 + k] = c[k]; }
      }
    } else if (m->body_jnttype[b] == 2) {
      real off[3], c[3];
      for (int k = 0; k < 3; k++) off[k] = c0[k] - d->xanchor[b][k];
      FL(3);
      cross3(c, d->xaxis[b], off);
      for (int k = 0; k < 3; k++) { d->cdof[dof][k] = d->xaxis[b][k]; d->cdof[dof][3 + k] = c[k]; }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * smooth.crb + support.make_m (dense, nv < 60) [R]
 * ---------------------------------------------------------------------------------------- */
static void crb(const ZbModel* m, const EnvModel* e, ZboData* d) {
  STAGE(ZBO_ST_CRB);
  memcpy(d->crb, d->cinert, sizeof(d->crb));
  for (int b = NB - 1; b > 0; b--) {
    int p = m->body_parent[b];
    for (int k = 0; k < 10; k++) d->crb[p][k] += d->crb[b][k];
    FL(10);
  }
  for (int k = 0; k < 10; k++) d->crb[0][k] = 0;
  memset(d->qM, 0, sizeof(d->qM));
  for (int i = 0; i < NV; i++) {
    real f[6];
    inert_mul(f, d->crb[m->dof_body[i]], d->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parent[j]) {
      real v = dot6(d->cdof[j], f);
      d->qM[i][j] = v;
      d->qM[j][i] = v;
    }
    d->qM[i][i] += e->dof_armature[i];
    FL(1);
  }
}

/* smooth.factor_m: dense Cholesky (jax.scipy.linalg.cho_factor) [R].  Returns 0 if SPD. */
static int chol(real L[NV][NV], const real A[NV][NV]) {
  int bad = 0;
  for (int i = 0; i < NV; i++)
    for (int j = 0; j <= i; j++) {
      real s = A[i][j];
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      FL(2 * j);
      if (i == j) {
        if (!(s > 0)) bad = 1;
        L[i][i] = R_SQRT(s);
        FL(1);
      } else {
        L[i][j] = s / L[j][j];
        L[j][i] = 0;
        FL(1);
      }
    }
  return bad;
}
/* smooth.solve_m: cho_solve */
static void chol_solve(const real L[NV][NV], real* x) {
  for (int i = 0; i < NV; i++) {
    real s = x[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
    x[i] = s / L[i][i];
    FL(2 * i + 1);
  }
  for (int i = NV - 1; i >= 0; i--) {
    real s = x[i];
    for (int k = i + 1; k < NV; k++) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
    FL(2 * (NV - 1 - i) + 1);
  }
}
static void mul_m(const ZboData* d, real* out, const real* v) {
  for (int i = 0; i < NV; i++) {
    real s = 0;
    for (int j = 0; j < NV; j++) s += d->qM[i][j] * v[j];
    out[i] = s;
  }
  FL(2 * 203 * 2 - NV); /* tree-sparse count: structural zeros are not algorithmic flops */
}

/* ------------------------------------------------------------------------------------------
 * collision_driver.collision -> collision_convex.plane_convex + _manifold_points [R] (A.7)
 * ---------------------------------------------------------------------------------------- */
static void make_frame(real* f, const real* n_in) {
  real a[3] = {n_in[0], n_in[1], n_in[2]};
  real n = R_SQRT(dot3(a, a));
  real s = (real)1 / (n + (real)1e-6 * (n == 0));
  a[0] *= s; a[1] *= s; a[2] *= s;
  real b[3] = {0, 0, 0};
  if ((real)-0.5 < a[1] && a[1] < (real)0.5) b[1] = 1; else b[2] = 1;
  real ab = dot3(a, b);
  for (int k = 0; k < 3; k++) b[k] -= a[k] * ab;
  n = R_SQRT(dot3(b, b));
  s = (real)1 / (n + (real)1e-6 * (n == 0));
  b[0] *= s; b[1] *= s; b[2] *= s;
  real c[3];
  cross3(c, a, b);
  for (int k = 0; k < 3; k++) { f[k] = a[k]; f[3 + k] = b[k]; f[6 + k] = c[k]; }
  FL(20);
}
static int argmax8(const real* v) { /* jp.argmax: first maximum */
  int bi = 0;
  for (int i = 1; i < 8; i++)
    if (v[i] > v[bi]) bi = i;
  return bi;
}
/* contact parameter mixing of group g (collision_driver: the higher geom_priority wins — floor = 2, train.py:1330 [V]; equal
 * priority: max friction, margins / gaps max; solref / solimp of geom 1 on a tie: solmix default 1:1 would average, documented risk) */
static void mix_params(const ZbModel* m, const EnvModel* e, ZboData* d, int g, int ci) {
  int world = m->pair_body[g] == 0;
  real mu1 = world ? e->floor_mu : (real)m->pair_friction[g][0], mu2 = (real)m->box_friction[g][0];
  int p1 = world ? m->floor_priority : m->pair_priority[g], p2 = m->box_priority[g];
  const float* sr1 = world ? m->floor_solref : m->pair_solref[g];
  const float* si1 = world ? m->floor_solimp : m->pair_solimp[g];
  real margin = R_FMAX(world ? (real)m->floor_margin : (real)m->pair_margin[g], (real)m->box_margin[g]);
  real gap = R_FMAX(world ? (real)m->floor_gap : (real)m->pair_gap[g], (real)m->box_gap[g]);
  d->con_includemargin[ci] = margin - gap;
  const float *sr, *si;
  if (p1 > p2) { d->con_mu[ci] = mu1; sr = sr1; si = si1; }
  else if (p1 < p2) { d->con_mu[ci] = mu2; sr = m->box_solref[g]; si = m->box_solimp[g]; }
  else { d->con_mu[ci] = R_FMAX(mu1, mu2); sr = sr1; si = si1; }
  for (int k = 0; k < 2; k++) d->con_solref[ci][k] = (real)sr[k];
  for (int k = 0; k < 5; k++) d->con_solimp[ci][k] = (real)si[k];
}
static void set_contact(const ZbModel* m, const EnvModel* e, ZboData* d, int g, int s, real dist, const real* pos, const real* frame, int vert) {
  int ci = g * 4 + s;
  d->con_dist[ci] = dist;
  for (int k = 0; k < 3; k++) d->con_pos[ci][k] = pos[k];
  memcpy(d->con_frame[ci], frame, 9 * sizeof(real));
  d->con_vert[ci] = vert;
  d->con_body[ci] = m->box_body[g];
  d->con_body1[ci] = m->pair_body[g];
  mix_params(m, e, d, g, ci);
}
/* collision_primitive._plane_sphere [R] */
static void plane_sphere_(const real* n, const real* ppos, const real* spos, real radius, real* dist, real* pos) {
  real rel[3] = {spos[0] - ppos[0], spos[1] - ppos[1], spos[2] - ppos[2]};
  *dist = dot3(rel, n) - radius;
  for (int k = 0; k < 3; k++) pos[k] = spos[k] - n[k] * (radius + (real)0.5 * *dist);
  FL(12);
}
/* collision_primitive._sphere_sphere [R]: normal from sphere 1 to sphere 2 */
static void sphere_sphere_(const real* p1, real r1, const real* p2, real r2, real* dist, real* pos, real* n) {
  real dl[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  real len = R_SQRT(dot3(dl, dl));
  real sc = (real)1 / (len + (real)1e-6 * (len == 0));
  for (int k = 0; k < 3; k++) n[k] = dl[k] * sc;
  if (len == 0) { n[0] = 1; n[1] = 0; n[2] = 0; }
  *dist = len - (r1 + r2);
  for (int k = 0; k < 3; k++) pos[k] = p1[k] + n[k] * (r1 + (real)0.5 * *dist);
  FL(20);
}
/* math.closest_segment_point [R] */
static void closest_segment_point(const real* a, const real* b, const real* pt, real* out) {
  real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ap[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  real t = dot3(ap, ab) / (dot3(ab, ab) + (real)1e-6);
  t = R_FMIN(R_FMAX(t, (real)0), (real)1);
  for (int k = 0; k < 3; k++) out[k] = a[k] + t * ab[k];
  FL(15);
}
/* math.closest_segment_to_segment_points [R] */
static void closest_segment_to_segment(const real* a0, const real* a1, const real* b0, const real* b1, real* best_a, real* best_b) {
  real da[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, db[3] = {b1[0] - b0[0], b1[1] - b0[1], b1[2] - b0[2]};
  real la = R_SQRT(dot3(da, da)), lb = R_SQRT(dot3(db, db));
  real sa = (real)1 / (la + (real)1e-6 * (la == 0)), sb = (real)1 / (lb + (real)1e-6 * (lb == 0));
  for (int k = 0; k < 3; k++) { da[k] *= sa; db[k] *= sb; }
  real ha = la * (real)0.5, hb = lb * (real)0.5;
  real am[3], bm[3], tr[3];
  for (int k = 0; k < 3; k++) { am[k] = a0[k] + da[k] * ha; bm[k] = b0[k] + db[k] * hb; tr[k] = am[k] - bm[k]; }
  real dd = dot3(da, db), dat = dot3(da, tr), dbt = dot3(db, tr);
  real denom = (real)1 - dd * dd;
  real ta = (-dat + dd * dbt) / (denom + (real)1e-6);
  real tb = dbt + ta * dd;
  ta = R_FMIN(R_FMAX(ta, -ha), ha);
  tb = R_FMIN(R_FMAX(tb, -hb), hb);
  real ba[3], bb[3], na[3], nb[3];
  for (int k = 0; k < 3; k++) { ba[k] = am[k] + da[k] * ta; bb[k] = bm[k] + db[k] * tb; }
  closest_segment_point(a0, a1, bb, na);
  closest_segment_point(b0, b1, ba, nb);
  real e1[3] = {na[0] - bb[0], na[1] - bb[1], na[2] - bb[2]}, e2[3] = {ba[0] - nb[0], ba[1] - nb[1], ba[2] - nb[2]};
  real d1 = dot3(e1, e1), d2 = dot3(e2, e2);
  for (int k = 0; k < 3; k++) { best_a[k] = d1 < d2 ? na[k] : ba[k]; best_b[k] = d1 < d2 ? bb[k] : nb[k]; }
  FL(80);
}
/* collision_convex.plane_convex + _manifold_points for a box [R] (A.7): four contacts */
static void plane_box(const ZbModel* m, const EnvModel* e, ZboData* d, int g, const real* pn, const real* pp) {
    const real* mat = d->box_xmat[g];
    const real* pos = d->box_xpos[g];
    /* plane into the convex frame */
    real rel[3] = {pp[0] - pos[0], pp[1] - pos[1], pp[2] - pos[2]}, plane_pos[3], n[3];
    matT_vec(plane_pos, mat, rel);
    matT_vec(n, mat, pn);
    real vert[8][3], support[8], smax = 0;
    for (int v = 0; v < 8; v++) { /* itertools.product((-1, 1), repeat=3) order */
      vert[v][0] = ((v & 4) ? 1 : -1) * (real)m->box_size[g][0];
      vert[v][1] = ((v & 2) ? 1 : -1) * (real)m->box_size[g][1];
      vert[v][2] = ((v & 1) ? 1 : -1) * (real)m->box_size[g][2];
      real dv[3] = {plane_pos[0] - vert[v][0], plane_pos[1] - vert[v][1], plane_pos[2] - vert[v][2]};
      support[v] = dot3(dv, n);
      if (v == 0 || support[v] > smax) smax = support[v];
      FL(3);
    }
    real thr = R_FMAX((real)0, smax - (real)1e-3);
    real dist_mask[8];
    for (int v = 0; v < 8; v++) dist_mask[v] = support[v] > thr ? (real)0 : (real)-1e6;
    /* _manifold_points */
    int idx[4];
    real tmp[16];
    idx[0] = argmax8(dist_mask);
    const real* a = vert[idx[0]];
    for (int v = 0; v < 8; v++) {
      real dv[3] = {a[0] - vert[v][0], a[1] - vert[v][1], a[2] - vert[v][2]};
      tmp[v] = dot3(dv, dv) + dist_mask[v];
      FL(4);
    }
    idx[1] = argmax8(tmp);
    const real* b = vert[idx[1]];
    real amb[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]}, ab[3];
    cross3(ab, n, amb);
    real ap[8][3];
    for (int v = 0; v < 8; v++) {
      for (int k = 0; k < 3; k++) ap[v][k] = a[k] - vert[v][k];
      tmp[v] = R_ABS(dot3(ap[v], ab)) + dist_mask[v];
      FL(4);
    }
    idx[2] = argmax8(tmp);
    const real* c = vert[idx[2]];
    real amc[3] = {a[0] - c[0], a[1] - c[1], a[2] - c[2]}, bmc[3] = {b[0] - c[0], b[1] - c[1], b[2] - c[2]}, ac[3], bc[3];
    cross3(ac, n, amc);
    cross3(bc, n, bmc);
    for (int v = 0; v < 8; v++) {
      real bp[3] = {b[0] - vert[v][0], b[1] - vert[v][1], b[2] - vert[v][2]};
      tmp[v] = R_ABS(dot3(bp, bc)) + dist_mask[v];
      tmp[8 + v] = R_ABS(dot3(ap[v], ac)) + dist_mask[v];
      FL(8);
    }
    {
      int bi = 0;
      for (int i = 1; i < 16; i++)
        if (tmp[i] > tmp[bi]) bi = i;
      idx[3] = bi % 8;
    }
    real frame[9];
    make_frame(frame, pn);
    for (int s = 0; s < 4; s++) {
      int unique = 1;
      for (int t = 0; t < s; t++)
        if (idx[t] == idx[s]) unique = 0;
      real dist = unique ? -support[idx[s]] : (real)1;
      real w[3], cp[3];
      mat_vec(w, mat, vert[idx[s]]);
      for (int k = 0; k < 3; k++) cp[k] = pos[k] + w[k] - (real)0.5 * dist * pn[k];
      FL(9);
      set_contact(m, e, d, g, s, dist, cp, frame, idx[s]);
    }
}
/* collision_driver.collision over the static candidate list = the collision groups of the model (ZbModel.box_* / pair_*) */
static void collision(const ZbModel* m, const EnvModel* e, ZboData* d) {
  STAGE(ZBO_ST_COLLISION);
  real pn[3] = {(real)m->floor_normal[0], (real)m->floor_normal[1], (real)m->floor_normal[2]};
  real pp[3] = {(real)m->floor_pos[0], (real)m->floor_pos[1], (real)m->floor_pos[2]};
  const real zero3[3] = {0, 0, 0};
  real idf[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int g = 0; g < NBOX; g++) {
    const real* mat = d->box_xmat[g];
    const real* pos = d->box_xpos[g];
    const int t2 = m->box_type[g], world = m->pair_body[g] == 0;
    real r2 = (real)m->box_size[g][0], h2 = (real)m->box_size[g][1];
    real ax2[3] = {mat[2], mat[5], mat[8]};
    for (int s = 0; s < 4; s++) set_contact(m, e, d, g, s, (real)1, zero3, idf, 0); /* inert slots */
    if (world && t2 == ZB_GEOM_BOX) {
      plane_box(m, e, d, g, pn, pp);
    } else if (world && t2 == ZB_GEOM_SPHERE) { /* collision_primitive.plane_sphere [R] */
      real dist, cp[3], frame[9];
      plane_sphere_(pn, pp, pos, r2, &dist, cp);
      make_frame(frame, pn);
      set_contact(m, e, d, g, 0, dist, cp, frame, 0);
    } else if (world && t2 == ZB_GEOM_CAPSULE) { /* collision_primitive.plane_capsule [R]: two contacts, frame aligned with the axis */
      real na = dot3(pn, ax2), bv[3], frame[9];
      for (int k = 0; k < 3; k++) bv[k] = ax2[k] - pn[k] * na;
      real bn = R_SQRT(dot3(bv, bv));
      real sc = (real)1 / (bn + (real)1e-6 * (bn == 0));
      for (int k = 0; k < 3; k++) bv[k] *= sc;
      if (bn < (real)0.5) {
        bv[0] = 0; bv[1] = 0; bv[2] = 0;
        if ((real)-0.5 < pn[1] && pn[1] < (real)0.5) bv[1] = 1; else bv[2] = 1;
      }
      real cr[3];
      cross3(cr, pn, bv);
      for (int k = 0; k < 3; k++) { frame[k] = pn[k]; frame[3 + k] = bv[k]; frame[6 + k] = cr[k]; }
      FL(30);
      for (int s = 0; s < 2; s++) {
        real c[3], dist, cp[3], sg = s == 0 ? (real)1 : (real)-1;
        for (int k = 0; k < 3; k++) c[k] = pos[k] + sg * ax2[k] * h2;
        plane_sphere_(pn, pp, c, r2, &dist, cp);
        set_contact(m, e, d, g, s, dist, cp, frame, s);
      }
    } else if (!world) { /* sphere_sphere / sphere_capsule / capsule_capsule [R]: one contact, normal from geom 1 to geom 2 */
      const real* mat1 = d->pair_xmat[g];
      const real* pos1 = d->pair_xpos[g];
      const int t1 = m->pair_type[g];
      real r1 = (real)m->pair_size[g][0], h1 = (real)m->pair_size[g][1];
      real ax1[3] = {mat1[2], mat1[5], mat1[8]};
      real p1[3] = {pos1[0], pos1[1], pos1[2]}, p2[3] = {pos[0], pos[1], pos[2]};
      if (t1 == ZB_GEOM_SPHERE && t2 == ZB_GEOM_CAPSULE) {
        real a[3], b[3];
        for (int k = 0; k < 3; k++) { a[k] = pos[k] - ax2[k] * h2; b[k] = pos[k] + ax2[k] * h2; }
        closest_segment_point(a, b, pos1, p2);
      } else if (t1 == ZB_GEOM_CAPSULE && t2 == ZB_GEOM_CAPSULE) {
        real a0[3], a1[3], b0[3], b1[3];
        for (int k = 0; k < 3; k++) { a0[k] = pos1[k] - ax1[k] * h1; a1[k] = pos1[k] + ax1[k] * h1; b0[k] = pos[k] - ax2[k] * h2; b1[k] = pos[k] + ax2[k] * h2; }
        closest_segment_to_segment(a0, a1, b0, b1, p1, p2);
      }
      real dist, cp[3], n[3], frame[9];
      sphere_sphere_(p1, r1, p2, r2, &dist, cp, n);
      make_frame(frame, n);
      set_contact(m, e, d, g, 0, dist, cp, frame, 0);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * constraint.make_constraint [R]: rows = friction-loss | joint limits | contacts
 * ---------------------------------------------------------------------------------------- */
static void kbi(const ZbOpts* o, const real* solref, const real* solimp, real pos, real* k_, real* b_, real* imp_) {
  real timeconst = solref[0], dampratio = solref[1];
  timeconst = R_FMAX(timeconst, 2 * (real)o->dt); /* refsafe */
  real dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  dmin = R_FMIN(R_FMAX(dmin, MJ_MINIMP), MJ_MAXIMP);
  dmax = R_FMIN(R_FMAX(dmax, MJ_MINIMP), MJ_MAXIMP);
  width = R_FMAX(MJ_MINVAL, width);
  mid = R_FMIN(R_FMAX(mid, MJ_MINIMP), MJ_MAXIMP);
  power = R_FMAX((real)1, power);
  real k = (real)1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  real b = (real)2 / (dmax * timeconst);
  if (solref[0] <= 0) { k = -solref[0] / (dmax * dmax); b = -solref[1] / dmax; }
  real x = R_ABS(pos) / width;
  real ia = ((real)1 / R_POW(mid, power - 1)) * R_POW(x, power);
  real ib = (real)1 - ((real)1 / R_POW((real)1 - mid, power - 1)) * R_POW((real)1 - x, power);
  real y = x < mid ? ia : ib;
  real imp = dmin + y * (dmax - dmin);
  imp = R_FMIN(R_FMAX(imp, dmin), dmax);
  if (x > (real)1) imp = dmax;
  *k_ = k; *b_ = b; *imp_ = imp;
  FL(30);
}
static void finish_row(const ZbOpts* o, ZboData* d, int r, real invweight, const real* solref, const real* solimp) {
  real k, b, imp, vel = 0;
  kbi(o, solref, solimp, d->efc_pos[r], &k, &b, &imp);
  for (int j = 0; j < NV; j++) vel += d->efc_J[r][j] * d->qvel[j];
  real R = R_FMAX(invweight * ((real)1 - imp) / imp, MJ_MINVAL);
  d->efc_R[r] = R;
  d->efc_D[r] = (real)1 / R;
  d->efc_aref[r] = -b * vel - k * imp * d->efc_pos[r];
  FL(10);
}
static void make_constraint(const ZbModel* m, const ZbOpts* o, const EnvModel* e, ZboData* d) {
  STAGE(ZBO_ST_CONSTRAINT);
  memset(d->efc_J, 0, sizeof(d->efc_J));
  for (int r = 0; r < NEFC; r++) { d->efc_type[r] = 0; d->efc_pos[r] = 0; d->efc_D[r] = 0; d->efc_R[r] = 0; d->efc_aref[r] = 0; d->efc_floss[r] = 0; }
  for (int dof = 0; dof < NV; dof++) {
    /* friction loss: dofs whose (compile-time) frictionloss > 0 */
    if (m->dof_frictionloss[dof] > 0) {
      int r = dof;
      real sr[2] = {(real)m->dof_solref[dof][0], (real)m->dof_solref[dof][1]}, si[5];
      for (int k = 0; k < 5; k++) si[k] = (real)m->dof_solimp[dof][k];
      d->efc_type[r] = 1;
      d->efc_J[r][dof] = 1;
      d->efc_pos[r] = 0;
      d->efc_floss[r] = e->dof_frictionloss[dof];
      finish_row(o, d, r, (real)m->dof_invweight0[dof], sr, si);
      FL(2);
    }
    /* joint limit (hinge), always instantiated, active iff pos < 0 */
    if (m->jnt_limited[dof]) {
      int r = NV + dof;
      real q = d->qpos[dof + 1];
      real dmin = q - (real)m->jnt_range[dof][0], dmax = (real)m->jnt_range[dof][1] - q;
      real pos = R_FMIN(dmin, dmax) - (real)m->jnt_margin[dof];
      FL(4);
      if (pos < 0) {
        real sr[2] = {(real)m->jnt_solref[dof][0], (real)m->jnt_solref[dof][1]}, si[5];
        for (int k = 0; k < 5; k++) si[k] = (real)m->jnt_solimp[dof][k];
        d->efc_type[r] = 2;
        d->efc_J[r][dof] = dmin < dmax ? (real)1 : (real)-1;
        d->efc_pos[r] = pos;
        finish_row(o, d, r, (real)m->dof_invweight0[dof], sr, si);
        FL(2);
      }
    }
  }
  /* contacts: pyramidal, condim 3 -> 4 rows each */
  const real* c0 = d->subtree_com[1];
  for (int c = 0; c < MAXCON; c++) {
    real dist = d->con_dist[c] - d->con_includemargin[c];
    int active = dist < 0;
    d->con_active[c] = active;
    if (!active) continue;
    int body = d->con_body[c];
    /* support.jac_dif_pair: jacp(body2) - jacp(body1); body1 = world contributes nothing */
    real jp[3][NV];
    memset(jp, 0, sizeof(jp));
    real off[3] = {d->con_pos[c][0] - c0[0], d->con_pos[c][1] - c0[1], d->con_pos[c][2] - c0[2]};
    for (int side = 0; side < 2; side++) {
      int b = side == 0 ? body : d->con_body1[c];
      real sg = side == 0 ? (real)1 : (real)-1;
      while (b > 0 && m->body_dofadr[b] < 0) b = m->body_parent[b];
      int dof = b > 0 ? m->body_dofadr[b] + (m->body_jnttype[b] == 1 ? 5 : 0) : -1;
      for (int j = dof; j >= 0; j = m->dof_parent[j]) {
        real cr[3];
        cross3(cr, d->cdof[j], off);
        for (int k = 0; k < 3; k++) jp[k][j] += sg * (cr[k] + d->cdof[j][3 + k]);
        FL(3);
      }
    }
    real dc[3][NV]; /* frame @ diff */
    const real* fr = d->con_frame[c];
    for (int k = 0; k < 3; k++)
      for (int j = 0; j < NV; j++) dc[k][j] = fr[k * 3] * jp[0][j] + fr[k * 3 + 1] * jp[1][j] + fr[k * 3 + 2] * jp[2][j];
    FL(5 * 3 * 12);
    real mu = d->con_mu[c];
    real t = (real)m->body_invweight0[d->con_body1[c]][0] + (real)m->body_invweight0[body][0];
    real invweight = (t + mu * mu * t) * 2 * mu * mu / (real)o->impratio;
    FL(8);
    for (int k = 0; k < 4; k++) {
      int r = 2 * NV + 4 * c + k;
      real sgn = (k & 1) ? (real)-1 : (real)1;
      const real* dt_ = dc[1 + k / 2];
      for (int j = 0; j < NV; j++) d->efc_J[r][j] = dc[0][j] + sgn * dt_[j] * mu;
      {
        real dir[3] = {fr[0] + sgn * mu * fr[3 * (1 + k / 2)], fr[1] + sgn * mu * fr[3 * (1 + k / 2) + 1], fr[2] + sgn * mu * fr[3 * (1 + k / 2) + 2]};
        real tq[3];
        cross3(tq, off, dir); FL(-9);
        for (int q_ = 0; q_ < 3; q_++) { g_w[r][q_] = tq[q_]; g_w[r][3 + q_] = dir[q_]; }
      }
      FL(2 * 12);
      d->efc_type[r] = 3;
      d->efc_pos[r] = dist;
      finish_row(o, d, r, invweight, d->con_solref[c], d->con_solimp[c]);
      FL(2 * 12);
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * smooth.com_vel, passive.passive, smooth.rne [R] (A.4, A.5)
 * ---------------------------------------------------------------------------------------- */
static void com_vel(const ZbModel* m, ZboData* d) {
  STAGE(ZBO_ST_COM_VEL);
  memset(d->cvel[0], 0, sizeof(real) * 6);
  for (int b = 1; b < NB; b++) {
    real v[6];
    memcpy(v, d->cvel[m->body_parent[b]], sizeof(v));
    int dof = m->body_dofadr[b];
    if (m->body_jnttype[b] == 1) {
      for (int i = 0; i < 3; i++) {
        for (int k = 0; k < 6; k++) { d->cdof_dot[dof + i][k] = 0; v[k] += d->cdof[dof + i][k] * d->qvel[dof + i]; }
        FL(12);
      }
      for (int i = 3; i < 6; i++) motion_cross(d->cdof_dot[dof + i], v, d->cdof[dof + i]);
      for (int i = 3; i < 6; i++) {
        for (int k = 0; k < 6; k++) v[k] += d->cdof[dof + i][k] * d->qvel[dof + i];
        FL(12);
      }
    } else if (m->body_jnttype[b] == 2) {
      motion_cross(d->cdof_dot[dof], v, d->cdof[dof]);
      for (int k = 0; k < 6; k++) v[k] += d->cdof[dof][k] * d->qvel[dof];
      FL(12);
    }
    memcpy(d->cvel[b], v, sizeof(v));
  }
}
static void passive(const EnvModel* e, ZboData* d) {
  for (int i = 0; i < NV; i++) d->qfrc_passive[i] = -e->dof_damping[i] * d->qvel[i];
  FL(NV);
}
static void rne(const ZbModel* m, const ZbOpts* o, ZboData* d) {
  STAGE(ZBO_ST_RNE);
  real cacc[NB][6], cfrc[NB][6];
  for (int k = 0; k < 3; k++) { cacc[0][k] = 0; cacc[0][3 + k] = -(real)o->gravity[k]; }
  for (int b = 1; b < NB; b++) {
    memcpy(cacc[b], cacc[m->body_parent[b]], sizeof(cacc[b]));
    int dof = m->body_dofadr[b], nd = m->body_jnttype[b] == 1 ? 6 : (m->body_jnttype[b] == 2 ? 1 : 0);
    for (int i = 0; i < nd; i++) {
      for (int k = 0; k < 6; k++) cacc[b][k] += d->cdof_dot[dof + i][k] * d->qvel[dof + i];
      FL(12);
    }
  }
  for (int b = 0; b < NB; b++) {
    real t[6], u[6];
    inert_mul(cfrc[b], d->cinert[b], cacc[b]);
    inert_mul(t, d->cinert[b], d->cvel[b]);
    motion_cross_force(u, d->cvel[b], t);
    for (int k = 0; k < 6; k++) cfrc[b][k] += u[k];
    FL(6);
  }
  for (int b = NB - 1; b > 0; b--) {
    int p = m->body_parent[b];
    for (int k = 0; k < 6; k++) cfrc[p][k] += cfrc[b][k];
    FL(6);
  }
  for (int i = 0; i < NV; i++) d->qfrc_bias[i] = dot6(d->cdof[i], cfrc[m->dof_body[i]]);
}

/* forward.fwd_actuation (motors) + fwd_acceleration [R] (A.6) */
static void fwd_actuation(const ZbModel* m, ZboData* d) {
  STAGE(ZBO_ST_ACTUATION);
  memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
  for (int j = 0; j < NU; j++) {
    real c = d->ctrl[j];
    if (m->act_ctrllimited[j]) c = R_FMIN(R_FMAX(c, (real)m->act_ctrlrange[j][0]), (real)m->act_ctrlrange[j][1]);
    real f = c; /* motor: gain 1, no bias */
    if (m->act_forcelimited[j]) f = R_FMIN(R_FMAX(f, (real)m->act_forcerange[j][0]), (real)m->act_forcerange[j][1]);
    d->actuator_force[j] = f;
    d->qfrc_actuator[m->act_dof[j]] += (real)m->act_gear[j] * f;
    FL(6);
  }
}
static void fwd_acceleration(ZboData* d) {
  for (int i = 0; i < NV; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  FL(2 * NV);
  chol_solve(d->qL, d->qacc_smooth);
}

/* Experiment switch (tests/diagnostics only): bit 0 = apply the contact rows matrix-free, the way the CUDA kernel does
 * (J v = w_r . V_box(v), J^T f = cdof . sum_r f_r w_r), instead of through the explicit efc_J rows. */
static void box_motion(const ZbModel* m, const ZboData* d, int body, const real* v, real* V) {
  for (int k = 0; k < 6; k++) V[k] = 0;
  int b = body;
  while (b > 0 && m->body_dofadr[b] < 0) b = m->body_parent[b];
  int dof = b > 0 ? m->body_dofadr[b] + (m->body_jnttype[b] == 1 ? 5 : 0) : -1;
  for (int j = dof; j >= 0; j = m->dof_parent[j])
    for (int k = 0; k < 6; k++) V[k] += d->cdof[j][k] * v[j];
}
static void jv_contacts(const ZboData* d, const real* v, real* jv) {
  for (int c = 0; c < MAXCON; c++) {
    if (!d->con_active[c]) continue;
    real V[6];
    box_motion(g_model, d, d->con_body[c], v, V);
    for (int k = 0; k < 4; k++) jv[2 * NV + 4 * c + k] = dot6(g_w[2 * NV + 4 * c + k], V);
  }
}

/* ------------------------------------------------------------------------------------------
 * solver.solve — primal CG / Newton with MJX's bracketing line search [R]
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  real qacc[NV], Ma[NV], Jaref[NEFC], force[NEFC], qfrc_c[NV], grad[NV], Mgrad[NV], search[NV];
  real gauss, cost, prev_cost;
  int active[NEFC]; /* quadratic-zone membership (Newton Hessian) */
  int niter;
  real jascale[NEFC]; /* magnitude of the terms summed into Jaref (verification hooks: how well its sign is determined) */
} Ctx;

/* ---- verification hooks (zbo_step_follow / zbo_ls_points, see the entry points at the end of the file) ---- */
typedef struct { int K; const real* alphas; real *d0, *d1, *scale, *margin, *cost, *search_out, *info; } LSQuery;
static _Thread_local const real* g_follow_trace; /* solver trace of another implementation: take ITS step sizes */
static _Thread_local int g_follow_niter;
static _Thread_local int g_force_start;          /* 0: own choice, 1: qacc_warmstart, 2: qacc_smooth */
static _Thread_local real* g_start_costs;        /* [2] cost at qacc_warmstart / qacc_smooth */
static _Thread_local const LSQuery* g_ls_query;
static _Thread_local const real* g_search_override; /* [NV] first search direction of another implementation */

static void update_constraint(const ZboData* d, Ctx* c) {
  real cost = 0;
  for (int r = 0; r < NEFC; r++) {
    real x = c->Jaref[r];
    c->force[r] = 0;
    c->active[r] = 0;
    if (d->efc_type[r] == 0) continue;
    if (d->efc_type[r] == 1) {
      real f = d->efc_floss[r], rf = d->efc_R[r] * f;
      if (x <= -rf && f > 0) { c->force[r] = f; cost += f * ((real)-0.5 * rf - x); }
      else if (x >= rf && f > 0) { c->force[r] = -f; cost += f * ((real)-0.5 * rf + x); }
      else { c->force[r] = -d->efc_D[r] * x; cost += (real)0.5 * d->efc_D[r] * x * x; c->active[r] = 1; }
      FL(8);
    } else if (x < 0) {
      c->force[r] = -d->efc_D[r] * x;
      cost += (real)0.5 * d->efc_D[r] * x * x;
      c->active[r] = 1;
      FL(6);
    }
  }
  for (int j = 0; j < NV; j++) c->qfrc_c[j] = 0;
  const int matrix_free = (g_variant & 1) && g_model->pair_body[0] == 0 && g_model->pair_body[1] == 0;
  if (matrix_free) {
    for (int g = 0; g < NBOX; g++) {
      real W[6] = {0, 0, 0, 0, 0, 0};
      int body = g_model->box_body[g];
      for (int r = 2 * NV + 16 * g; r < 2 * NV + 16 * g + 16; r++)
        if (c->force[r] != 0)
          for (int k = 0; k < 6; k++) W[k] += c->force[r] * g_w[r][k];
      int b = body;
      while (b > 0 && g_model->body_dofadr[b] < 0) b = g_model->body_parent[b];
      int dof = b > 0 ? g_model->body_dofadr[b] + (g_model->body_jnttype[b] == 1 ? 5 : 0) : -1;
      for (int j = dof; j >= 0; j = g_model->dof_parent[j]) c->qfrc_c[j] += dot6(d->cdof[j], W);
    }
  }
  for (int r = 0; r < NEFC; r++)
    if (c->force[r] != 0 && !(matrix_free && d->efc_type[r] == 3)) {
      for (int j = 0; j < NV; j++) c->qfrc_c[j] += d->efc_J[r][j] * c->force[r];
      FL(d->efc_type[r] == 3 ? 24 : 2);
    }
  real g = 0;
  for (int j = 0; j < NV; j++) g += (c->Ma[j] - d->qfrc_smooth[j]) * (c->qacc[j] - d->qacc_smooth[j]);
  FL(4 * NV);
  c->gauss = (real)0.5 * g;
  c->prev_cost = c->cost;
  c->cost = c->gauss + cost;
}
static void ctx_create(const ZboData* d, Ctx* c, const real* qacc) {
  memcpy(c->qacc, qacc, sizeof(c->qacc));
  for (int r = 0; r < NEFC; r++) {
    real s = 0;
    if (d->efc_type[r])
      for (int j = 0; j < NV; j++) s += d->efc_J[r][j] * qacc[j];
    c->Jaref[r] = s - d->efc_aref[r];
    if (d->efc_type[r]) FL(d->efc_type[r] == 3 ? 25 : 3);
    c->jascale[r] = 0;
    if (g_ls_query && d->efc_type[r]) {
      real a = R_ABS(d->efc_aref[r]);
      for (int j = 0; j < NV; j++) a += R_ABS(d->efc_J[r][j] * qacc[j]);
      c->jascale[r] = a;
    }
  }
  if (g_variant & 1) {
    real jv[NEFC];
    jv_contacts(d, qacc, jv);
    for (int r = 2 * NV; r < NEFC; r++)
      if (d->efc_type[r] == 3) c->Jaref[r] = jv[r] - d->efc_aref[r];
  }
  mul_m(d, c->Ma, qacc);
  memset(c->grad, 0, sizeof(c->grad));
  memset(c->Mgrad, 0, sizeof(c->Mgrad));
  memset(c->search, 0, sizeof(c->search));
  c->gauss = 0;
  c->cost = (real)INFINITY;
  c->prev_cost = 0;
  c->niter = 0;
  update_constraint(d, c);
}
static void update_gradient(const ZbOpts* o, const ZboData* d, Ctx* c) {
  for (int j = 0; j < NV; j++) c->grad[j] = c->Ma[j] - d->qfrc_smooth[j] - c->qfrc_c[j];
  FL(2 * NV);
  memcpy(c->Mgrad, c->grad, sizeof(c->grad));
  if (o->solver == ZB_SOLVER_NEWTON) {
    static _Thread_local real H[NV][NV], L[NV][NV];
    memcpy(H, d->qM, sizeof(H));
    for (int r = 0; r < NEFC; r++)
      if (c->active[r]) {
        for (int i = 0; i < NV; i++) {
          real ji = d->efc_J[r][i] * d->efc_D[r];
          if (ji != 0)
            for (int j = 0; j < NV; j++) H[i][j] += ji * d->efc_J[r][j];
        }
        FL(d->efc_type[r] == 3 ? 12 + 2 * 78 : 3);
      }
    chol(L, H);
    chol_solve(L, c->Mgrad);
  } else {
    chol_solve(d->qL, c->Mgrad);
  }
}

typedef struct { real alpha, cost, deriv_0, deriv_1; } LSPoint;

static _Thread_local real* g_lstrace; /* where ls_point logs (alpha, d0, d1) during the first line search, else NULL */
static _Thread_local int g_lscount;
static LSPoint ls_point(const ZboData* d, const Ctx* c, real alpha, const real* jv, const real* quad_gauss) {
  real q0 = quad_gauss[0], q1 = quad_gauss[1], q2 = quad_gauss[2];
  for (int r = 0; r < NEFC; r++) {
    if (d->efc_type[r] == 0) continue;
    real x0 = c->Jaref[r], x = x0 + alpha * jv[r], D = d->efc_D[r];
    if (d->efc_type[r] == 1) {
      real f = d->efc_floss[r], rf = d->efc_R[r] * f;
      if (x <= -rf && f > 0) { q0 += f * ((real)-0.5 * rf - x0); q1 += -f * jv[r]; }
      else if (x >= rf && f > 0) { q0 += f * ((real)-0.5 * rf + x0); q1 += f * jv[r]; }
      else { q0 += (real)0.5 * D * x0 * x0; q1 += D * x0 * jv[r]; q2 += (real)0.5 * D * jv[r] * jv[r]; }
      FL(14);
    } else if (x < 0) {
      q0 += (real)0.5 * D * x0 * x0; q1 += D * x0 * jv[r]; q2 += (real)0.5 * D * jv[r] * jv[r];
      FL(12);
    }
  }
  LSPoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.deriv_0 = 2 * alpha * q2 + q1;
  p.deriv_1 = 2 * q2 + (q2 == 0 ? MJ_MINVAL : (real)0);
  FL(10);
  if (g_lstrace && 32 + 3 * g_lscount + 2 < ZB_TRACE_SEARCH) {
    g_lstrace[32 + 3 * g_lscount] = alpha; g_lstrace[32 + 3 * g_lscount + 1] = p.deriv_0; g_lstrace[32 + 3 * g_lscount + 2] = p.deriv_1;
    g_lscount++;
  }
  return p;
}
static inline int in_bracket(real x, real y) { return (x < y && y < 0) || (x > y && y > 0); }

/* Line-search statistics (diagnostics for the kernel design, single-threaded runs only; zbo_lsstats_get):
 * [0] line searches, [1] bracket iterations, [2] iterations entered with an end on an exactly zero derivative,
 * [3] iterations entered with both ends on the same pieces of every row ("same set"), [4] both, [5] same set and all three
 * candidate step sizes inside the bracket, [6] same set + zero end + candidates inside, [7] line searches that ran out of iterations. */
static double g_lsstat[8];
static void zone_sig(const ZboData* d, const Ctx* c, real alpha, const real* jv, signed char* sig) {
  for (int r = 0; r < NEFC; r++) {
    sig[r] = 0;
    if (d->efc_type[r] == 0) continue;
    real x = c->Jaref[r] + alpha * jv[r];
    if (d->efc_type[r] == 1) {
      real f = d->efc_floss[r], rf = d->efc_R[r] * f;
      sig[r] = (x <= -rf && f > 0) ? 1 : ((x >= rf && f > 0) ? 2 : 0);
    } else sig[r] = x < 0 ? 4 : 0;
  }
}

/* J search, M search and the Gauss part of the 1-D quadratic: the prelude of solver._linesearch */
static _Thread_local real g_jvscale[NEFC]; /* magnitude of the terms summed into J search */
static void ls_prelude(const ZboData* d, const Ctx* c, real* mv, real* jv, real* qg) {
  mul_m(d, mv, c->search);
  for (int r = 0; r < NEFC; r++) {
    real s = 0, a = 0;
    if (d->efc_type[r])
      for (int j = 0; j < NV; j++) { s += d->efc_J[r][j] * c->search[j]; a += R_ABS(d->efc_J[r][j] * c->search[j]); }
    jv[r] = s;
    g_jvscale[r] = a;
  }
  if ((g_variant & 1) && g_model->pair_body[0] == 0 && g_model->pair_body[1] == 0) jv_contacts(d, c->search, jv);
  qg[0] = c->gauss; qg[1] = 0; qg[2] = 0;
  for (int j = 0; j < NV; j++) {
    qg[1] += c->search[j] * c->Ma[j] - c->search[j] * d->qfrc_smooth[j];
    qg[2] += c->search[j] * mv[j];
  }
  qg[2] *= (real)0.5;
}
/* take a given step size along the current search direction (what _linesearch does with its own alpha) */
static void follow_step(const ZboData* d, Ctx* c, real alpha) {
  real mv[NV], jv[NEFC], qg[3];
  ls_prelude(d, c, mv, jv, qg);
  if (alpha != 0) {
    for (int j = 0; j < NV; j++) { c->qacc[j] += c->search[j] * alpha; c->Ma[j] += mv[j] * alpha; }
    for (int r = 0; r < NEFC; r++) c->Jaref[r] += jv[r] * alpha;
  }
}
/* first and second derivative of the 1-D cost along the current search direction at given step sizes, with the magnitude of
 * the summed terms (tolerance scale) and the relative distance of the nearest row to a change of its quadratic piece */
static void ls_query(const ZboData* d, const Ctx* c, const LSQuery* Q) {
  real mv[NV], jv[NEFC], qg[3];
  ls_prelude(d, c, mv, jv, qg);
  if (Q->info) {
    /* info[0] contact set; info[1] / info[3]: distance of the nearest joint limit / contact slot to its activation threshold
     * (rows that exist or not depending on rounding); info[2]: backward error of the search direction in use against
     * M search = -grad, relative to the magnitude of the terms that were summed */
    int mask = 0;
    real lim = (real)1e30, con = (real)1e30, be = 0;
    for (int k = 0; k < MAXCON; k++) {
      mask |= d->con_active[k] << k;
      con = R_FMIN(con, R_ABS(d->con_dist[k] - d->con_includemargin[k]));
    }
    for (int dof = 0; dof < NV; dof++)
      if (g_model->jnt_limited[dof]) {
        real q = d->qpos[dof + 1];
        real dmin = q - (real)g_model->jnt_range[dof][0], dmax = (real)g_model->jnt_range[dof][1] - q;
        lim = R_FMIN(lim, R_ABS(R_FMIN(dmin, dmax) - (real)g_model->jnt_margin[dof]));
      }
    for (int j = 0; j < NV; j++) {
      real ga = R_ABS(c->Ma[j]) + R_ABS(d->qfrc_smooth[j]), ms = 0;
      for (int r = 0; r < NEFC; r++) if (c->force[r] != 0) ga += R_ABS(d->efc_J[r][j] * c->force[r]);
      for (int k = 0; k < NV; k++) ms += R_ABS(d->qM[j][k] * c->search[k]);
      real den = ga + ms;
      if (den > 0) be = R_FMAX(be, R_ABS(mv[j] + c->grad[j]) / den);
    }
    Q->info[0] = (real)mask; Q->info[1] = lim; Q->info[2] = be; Q->info[3] = con;
  }
  real g1 = 0, g2 = 0;
  for (int j = 0; j < NV; j++) {
    g1 += R_ABS(c->search[j] * c->Ma[j]) + R_ABS(c->search[j] * d->qfrc_smooth[j]);
    g2 += (real)0.5 * R_ABS(c->search[j] * mv[j]);
  }
  for (int k = 0; k < Q->K; k++) {
    real alpha = Q->alphas[k], q0 = qg[0], q1 = qg[1], q2 = qg[2], a1 = g1, a2 = g2, margin = (real)1e30;
    for (int r = 0; r < NEFC; r++) {
      if (d->efc_type[r] == 0) continue;
      real x0 = c->Jaref[r], x = x0 + alpha * jv[r], D = d->efc_D[r];
      real sc = c->jascale[r] + R_ABS(alpha) * g_jvscale[r]; /* magnitude of everything summed into x */
      if (d->efc_type[r] == 1) {
        real f = d->efc_floss[r], rf = d->efc_R[r] * f;
        if (x <= -rf && f > 0) { q0 += f * ((real)-0.5 * rf - x0); q1 += -f * jv[r]; a1 += R_ABS(f * jv[r]); }
        else if (x >= rf && f > 0) { q0 += f * ((real)-0.5 * rf + x0); q1 += f * jv[r]; a1 += R_ABS(f * jv[r]); }
        else { q0 += (real)0.5 * D * x0 * x0; q1 += D * x0 * jv[r]; q2 += (real)0.5 * D * jv[r] * jv[r]; a1 += R_ABS(D * jv[r]) * c->jascale[r]; a2 += (real)0.5 * D * jv[r] * jv[r]; }
        if (f > 0 && sc + rf > 0) {
          real m1 = R_ABS(x + rf) / (sc + rf), m2 = R_ABS(x - rf) / (sc + rf);
          margin = R_FMIN(margin, R_FMIN(m1, m2));
        }
      } else {
        if (x < 0) { q0 += (real)0.5 * D * x0 * x0; q1 += D * x0 * jv[r]; q2 += (real)0.5 * D * jv[r] * jv[r]; a1 += R_ABS(D * jv[r]) * c->jascale[r]; a2 += (real)0.5 * D * jv[r] * jv[r]; }
        if (sc > 0) margin = R_FMIN(margin, R_ABS(x) / sc);
      }
    }
    Q->d0[k] = 2 * alpha * q2 + q1;
    Q->d1[k] = 2 * q2 + (q2 == 0 ? MJ_MINVAL : (real)0);
    Q->scale[k] = a1 + 2 * R_ABS(alpha) * a2;
    Q->margin[k] = margin;
    if (Q->cost) Q->cost[k] = alpha * alpha * q2 + alpha * q1 + q0;
  }
}

static void linesearch(const ZbModel* m, const ZbOpts* o, ZboData* d, Ctx* c) {
  real nrm = 0;
  for (int j = 0; j < NV; j++) nrm += c->search[j] * c->search[j];
  real smag = R_SQRT(nrm) * (real)m->meaninertia * (real)(NV > 1 ? NV : 1);
  real gtol = (real)o->tolerance * (real)o->ls_tolerance * smag;
  real mv[NV], jv[NEFC];
  mul_m(d, mv, c->search);
  for (int r = 0; r < NEFC; r++) {
    real s = 0;
    if (d->efc_type[r]) {
      for (int j = 0; j < NV; j++) s += d->efc_J[r][j] * c->search[j];
      FL(d->efc_type[r] == 3 ? 24 : 1);
    }
    jv[r] = s;
  }
  if ((g_variant & 1) && g_model->pair_body[0] == 0 && g_model->pair_body[1] == 0) jv_contacts(d, c->search, jv);
  real qg[3] = {c->gauss, 0, 0};
  for (int j = 0; j < NV; j++) {
    qg[1] += c->search[j] * c->Ma[j] - c->search[j] * d->qfrc_smooth[j];
    qg[2] += c->search[j] * mv[j];
  }
  qg[2] *= (real)0.5;
  FL(8 * NV);
  g_lstrace = c->niter == 0 ? d->trace : NULL;
  if (c->niter == 0) for (int j = 0; j < NV; j++) d->trace[ZB_TRACE_SEARCH + j] = c->search[j];
  g_lscount = 0;
  LSPoint p0 = ls_point(d, c, 0, jv, qg);
  LSPoint lo = ls_point(d, c, p0.alpha - p0.deriv_0 / p0.deriv_1, jv, qg), hi;
  d->ls_evals += 2;
  if (lo.deriv_0 < p0.deriv_0) { hi = p0; } else { hi = lo; lo = p0; }
  int swap = 1, it = 0;
  g_lsstat[0] += 1;
  while (1) {
    int done = it >= o->ls_iterations;
    done |= !swap;
    done |= (lo.deriv_0 < 0) && (lo.deriv_0 > -gtol);
    done |= (hi.deriv_0 > 0) && (hi.deriv_0 < gtol);
    if ((g_variant & 2) && (lo.deriv_0 == 0 || hi.deriv_0 == 0)) done = 1; /* experiment: stop once an end sits on the exact minimum */
    if (done) break;
    if (g_variant & 4) { /* statistics */
      signed char za[NEFC], zb[NEFC];
      zone_sig(d, c, lo.alpha, jv, za);
      zone_sig(d, c, hi.alpha, jv, zb);
      int same = memcmp(za, zb, sizeof(za)) == 0, zero = lo.deriv_0 == 0 || hi.deriv_0 == 0;
      real a0 = lo.alpha - lo.deriv_0 / lo.deriv_1, a1 = hi.alpha - hi.deriv_0 / hi.deriv_1;
      real mn = lo.alpha < hi.alpha ? lo.alpha : hi.alpha, mx = lo.alpha < hi.alpha ? hi.alpha : lo.alpha;
      int inside = a0 >= mn && a0 <= mx && a1 >= mn && a1 <= mx;
      g_lsstat[1] += 1; g_lsstat[2] += zero; g_lsstat[3] += same; g_lsstat[4] += same && zero; g_lsstat[5] += same && inside;
      g_lsstat[6] += same && zero && inside;
      if (it == o->ls_iterations - 1) g_lsstat[7] += 1;
    }
    LSPoint lo_next = ls_point(d, c, lo.alpha - lo.deriv_0 / lo.deriv_1, jv, qg);
    LSPoint hi_next = ls_point(d, c, hi.alpha - hi.deriv_0 / hi.deriv_1, jv, qg);
    LSPoint mid = ls_point(d, c, (real)0.5 * (lo.alpha + hi.alpha), jv, qg);
    d->ls_evals += 3;
    int s1 = in_bracket(lo.deriv_0, lo_next.deriv_0);
    if (s1) lo = lo_next;
    int s2 = in_bracket(lo.deriv_0, mid.deriv_0);
    if (s2) lo = mid;
    int s3 = in_bracket(lo.deriv_0, hi_next.deriv_0);
    if (s3) lo = hi_next;
    int s4 = in_bracket(hi.deriv_0, hi_next.deriv_0);
    if (s4) hi = hi_next;
    int s5 = in_bracket(hi.deriv_0, mid.deriv_0);
    if (s5) hi = mid;
    int s6 = in_bracket(hi.deriv_0, lo_next.deriv_0);
    if (s6) hi = lo_next;
    swap = s1 | s2 | s3 | s4 | s5 | s6;
    it++;
  }
  d->ls_iters += it;
  g_lstrace = NULL;
  int improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  real alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (4 * c->niter + 3 < ZB_NTRACE) {
    d->trace[4 * c->niter] = improved ? alpha : 0;
    d->trace[4 * c->niter + 2] = (real)it;
    d->trace[4 * c->niter + 3] = R_SQRT(nrm);
  }
  if (improved) {
    for (int j = 0; j < NV; j++) { c->qacc[j] += c->search[j] * alpha; c->Ma[j] += mv[j] * alpha; }
    for (int r = 0; r < NEFC; r++) c->Jaref[r] += jv[r] * alpha;
    FL(4 * NV + 2 * 72);
  }
}

static void solve(const ZbModel* m, const ZbOpts* o, ZboData* d) {
  STAGE(ZBO_ST_SOLVER);
  g_model = m;
  static _Thread_local Ctx c, c2;
  const real* start = d->qacc_smooth;
  if (o->warmstart) {
    ctx_create(d, &c, d->qacc_warmstart);
    ctx_create(d, &c2, d->qacc_smooth);
    int warm = c.cost < c2.cost;
    if (g_start_costs) { g_start_costs[0] = c.cost; g_start_costs[1] = c2.cost; }
    if (g_force_start == 1) warm = 1;
    if (g_force_start == 2) warm = 0;
    start = warm ? d->qacc_warmstart : d->qacc_smooth;
    d->trace[ZB_TRACE_START] = warm ? 1 : 2;
  }
  ctx_create(d, &c, start);
  update_gradient(o, d, &c);
  for (int j = 0; j < NV; j++) c.search[j] = -c.Mgrad[j];
  if (g_ls_query && g_ls_query->search_out) memcpy(g_ls_query->search_out, c.search, sizeof(c.search));
  if (g_search_override) memcpy(c.search, g_search_override, sizeof(c.search));
  real scale = (real)1 / ((real)m->meaninertia * (real)(NV > 1 ? NV : 1));
  d->ls_evals = 0;
  d->ls_iters = 0;
  while (1) {
    real improvement = scale * (c.prev_cost - c.cost);
    real gn = 0;
    for (int j = 0; j < NV; j++) gn += c.grad[j] * c.grad[j];
    real gradient = scale * R_SQRT(gn);
    FL(2 * NV + 4);
    int done = c.niter >= o->iterations;
    done |= improvement < (real)o->tolerance;
    done |= gradient < (real)o->tolerance;
    if (g_follow_trace) done = c.niter >= g_follow_niter; /* follow mode: as many iterations as the traced implementation ran */
    if (g_ls_query) { ls_query(d, &c, g_ls_query); break; }
    if (done) break;
    if (g_follow_trace) follow_step(d, &c, g_follow_trace[4 * c.niter]);
    else linesearch(m, o, d, &c);
    real pg[NV], pMg[NV];
    memcpy(pg, c.grad, sizeof(pg));
    memcpy(pMg, c.Mgrad, sizeof(pMg));
    update_constraint(d, &c);
    if (4 * c.niter + 3 < ZB_NTRACE) d->trace[4 * c.niter + 1] = c.cost;
    update_gradient(o, d, &c);
    if (o->solver == ZB_SOLVER_NEWTON) {
      for (int j = 0; j < NV; j++) c.search[j] = -c.Mgrad[j];
    } else { /* Polak-Ribiere+ */
      real num = 0, den = 0;
      for (int j = 0; j < NV; j++) { num += c.grad[j] * (c.Mgrad[j] - pMg[j]); den += pg[j] * pMg[j]; }
      real beta = num / R_FMAX(MJ_MINVAL, den);
      beta = R_FMAX((real)0, beta);
      for (int j = 0; j < NV; j++) c.search[j] = -c.Mgrad[j] + beta * c.search[j];
      FL(7 * NV + 2);
    }
    c.niter++;
  }
  memcpy(d->qacc, c.qacc, sizeof(c.qacc));
  memcpy(d->efc_force, c.force, sizeof(c.force));
  memcpy(d->qfrc_constraint, c.qfrc_c, sizeof(c.qfrc_c));
  d->solver_niter = c.niter;
  d->solver_cost = c.cost;
}

/* ------------------------------------------------------------------------------------------
 * sensor.sensor_pos / sensor_vel / sensor_acc + smooth.rne_postconstraint [R] (A.11)
 * sensordata layout: include/zbphys.h ZB_SENS_* (train.py:44-46, 1495-1516 [V])
 * ---------------------------------------------------------------------------------------- */
static void sensors(const ZbModel* m, const ZbOpts* o, const EnvModel* e, ZboData* d) {
  STAGE(ZBO_ST_SENSOR);
  const real* c0 = d->subtree_com[1];
  /* framequat(imu_site) */
  {
    real q[4];
    quat_mul(q, d->xquat[m->site_body[0]], e->imu_quat);
    normalize4(q);
    memcpy(d->sensordata + ZB_SENS_IMU_QUAT, q, sizeof(q));
  }
  /* gyro */
  matT_vec(d->sensordata + ZB_SENS_IMU_GYRO, d->site_xmat[0], d->cvel[m->site_body[0]]);
  /* rne_postconstraint: cacc with qacc, cfrc_ext from contacts, cfrc_int accumulated leaf -> root */
  for (int k = 0; k < 3; k++) { d->cacc[0][k] = 0; d->cacc[0][3 + k] = -(real)o->gravity[k]; }
  for (int b = 1; b < NB; b++) {
    memcpy(d->cacc[b], d->cacc[m->body_parent[b]], sizeof(d->cacc[b]));
    int dof = m->body_dofadr[b], nd = m->body_jnttype[b] == 1 ? 6 : (m->body_jnttype[b] == 2 ? 1 : 0);
    for (int i = 0; i < nd; i++) {
      for (int k = 0; k < 6; k++) d->cacc[b][k] += d->cdof_dot[dof + i][k] * d->qvel[dof + i] + d->cdof[dof + i][k] * d->qacc[dof + i];
      FL(24);
    }
  }
  memset(d->cfrc_ext, 0, sizeof(d->cfrc_ext));
  real con_normal[MAXCON];
  for (int c = 0; c < MAXCON; c++) {
    con_normal[c] = 0;
    if (!d->con_active[c]) continue;
    const real* f = d->efc_force + 2 * NV + 4 * c;
    real mu = d->con_mu[c];
    /* pyramid -> contact-frame force: normal = sum of edges, tangents = mu * (f+ - f-) */
    real fc[3] = {f[0] + f[1] + f[2] + f[3], mu * (f[0] - f[1]), mu * (f[2] - f[3])}, fw[3], tq[3];
    con_normal[c] = fc[0];
    matT_vec(fw, d->con_frame[c], fc); /* world force on body2 (the foot) */
    real off[3] = {d->con_pos[c][0] - c0[0], d->con_pos[c][1] - c0[1], d->con_pos[c][2] - c0[2]};
    cross3(tq, off, fw);
    int b = d->con_body[c], b1 = d->con_body1[c];
    for (int k = 0; k < 3; k++) { d->cfrc_ext[b][k] += tq[k]; d->cfrc_ext[b][3 + k] += fw[k]; }
    if (b1 > 0) /* the reaction on the body of geom 1 */
      for (int k = 0; k < 3; k++) { d->cfrc_ext[b1][k] -= tq[k]; d->cfrc_ext[b1][3 + k] -= fw[k]; }
    FL(17);
  }
  for (int b = 0; b < NB; b++) {
    real t[6], u[6];
    inert_mul(d->cfrc_int[b], d->cinert[b], d->cacc[b]);
    inert_mul(t, d->cinert[b], d->cvel[b]);
    motion_cross_force(u, d->cvel[b], t);
    for (int k = 0; k < 6; k++) d->cfrc_int[b][k] += u[k] - d->cfrc_ext[b][k];
    FL(12);
  }
  for (int b = NB - 1; b > 0; b--) {
    int p = m->body_parent[b];
    for (int k = 0; k < 6; k++) d->cfrc_int[p][k] += d->cfrc_int[b][k];
    FL(6);
  }
  /* accelerometer(imu_site) */
  {
    int b = m->site_body[0];
    const real* rot = d->site_xmat[0];
    real diff[3] = {d->site_xpos[0][0] - c0[0], d->site_xpos[0][1] - c0[1], d->site_xpos[0][2] - c0[2]};
    real ang[3], lin[3], acc[3], t[3], u[3], corr[3];
    matT_vec(ang, rot, d->cvel[b]);
    cross3(t, diff, d->cvel[b]);
    for (int k = 0; k < 3; k++) u[k] = d->cvel[b][3 + k] - t[k];
    matT_vec(lin, rot, u);
    cross3(t, diff, d->cacc[b]);
    for (int k = 0; k < 3; k++) u[k] = d->cacc[b][3 + k] - t[k];
    matT_vec(acc, rot, u);
    cross3(corr, ang, lin);
    for (int k = 0; k < 3; k++) d->sensordata[ZB_SENS_IMU_ACC + k] = acc[k] + corr[k];
    FL(12);
  }
  /* touch(left_foot, right_foot): normal force of the contacts on the site's body that lie inside the site box */
  for (int s = 1; s <= 2; s++) {
    real total = 0;
    for (int c = 0; c < MAXCON; c++) {
      if (!d->con_active[c] || d->con_body[c] != m->site_body[s]) continue;
      real rel[3] = {d->con_pos[c][0] - d->site_xpos[s][0], d->con_pos[c][1] - d->site_xpos[s][1], d->con_pos[c][2] - d->site_xpos[s][2]}, loc[3];
      matT_vec(loc, d->site_xmat[s], rel);
      if (R_ABS(loc[0]) <= (real)m->site_size[s][0] && R_ABS(loc[1]) <= (real)m->site_size[s][1] && R_ABS(loc[2]) <= (real)m->site_size[s][2])
        total += con_normal[c];
      FL(4);
    }
    d->sensordata[s == 1 ? ZB_SENS_TOUCH_L : ZB_SENS_TOUCH_R] = total;
  }
  /* force(left_foot, right_foot): cfrc_int force part in the site frame */
  for (int s = 1; s <= 2; s++)
    matT_vec(d->sensordata + (s == 1 ? ZB_SENS_FORCE_L : ZB_SENS_FORCE_R), d->site_xmat[s], d->cfrc_int[m->site_body[s]] + 3);
}

/* ------------------------------------------------------------------------------------------
 * forward.forward / euler / implicit(fast) / _advance [R]
 * ---------------------------------------------------------------------------------------- */
static void forward(const ZbModel* m, const ZbOpts* o, const EnvModel* e, ZboData* d) {
  kinematics(m, e, d);
  com_pos(m, e, d);
  crb(m, e, d);
  STAGE(ZBO_ST_FACTOR);
  d->flags |= chol(d->qL, d->qM);
  collision(m, e, d);
  make_constraint(m, o, e, d);
  com_vel(m, d);
  STAGE(ZBO_ST_RNE);
  passive(e, d);
  rne(m, o, d);
  fwd_actuation(m, d);
  STAGE(ZBO_ST_ACCEL);
  fwd_acceleration(d);
  solve(m, o, d);
  sensors(m, o, e, d);
}

static void integrate(const ZbModel* m, const ZbOpts* o, const EnvModel* e, ZboData* d) {
  STAGE(ZBO_ST_INTEGRATE);
  real qacc[NV];
  memcpy(qacc, d->qacc, sizeof(qacc));
  int damped = 0;
  for (int i = 0; i < NV; i++) damped |= m->dof_damping[i] > 0;
  if (o->integrator == ZB_INT_IMPLICITFAST || (o->integrator == ZB_INT_EULER && o->eulerdamp && damped)) {
    /* implicitfast with motors + joint damping only: qDeriv = -diag(damping)  (A.10) */
    static _Thread_local real H[NV][NV], L[NV][NV];
    memcpy(H, d->qM, sizeof(H));
    for (int i = 0; i < NV; i++) H[i][i] += (real)o->dt * e->dof_damping[i];
    FL(2 * NV);
    chol(L, H);
    for (int i = 0; i < NV; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    FL(NV);
    chol_solve(L, qacc);
  }
  real dt = (real)o->dt;
  for (int i = 0; i < NV; i++) d->qvel[i] += qacc[i] * dt;
  for (int k = 0; k < 3; k++) d->qpos[k] += dt * d->qvel[k];
  { /* math.quat_integrate */
    real* w = d->qvel + 3;
    real n = R_SQRT(dot3(w, w));
    real s = (real)1 / (n + (real)1e-6 * (n == 0));
    real ax[3] = {w[0] * s, w[1] * s, w[2] * s}, qr[4], q2[4];
    axis_angle_to_quat(qr, ax, dt * n);
    quat_mul(q2, d->qpos + 3, qr);
    normalize4(q2);
    memcpy(d->qpos + 3, q2, sizeof(q2));
  }
  for (int i = 6; i < NV; i++) d->qpos[i + 1] += dt * d->qvel[i];
  FL(4 * NV + 12);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(d->qacc)); /* solver.solve leaves qacc_warmstart = qacc */
}

/* ------------------------------------------------------------------------------------------
 * FeetechActuators.get_stateful_ctrl — /root/reference/train.py:1242-1280 [V]
 * trapezoidal_step                   — /root/reference/train.py:1137-1196 [V]
 * ---------------------------------------------------------------------------------------- */
static inline real sign_(real x) { return (real)((x > 0) - (x < 0)); }
static void feetech(const ZbModel* m, real dt, const real* action, const real* qpos, const real* qvel, real* pos, real* vel, real* torque, real* ctrl) {
  STAGE(ZBO_ST_ACTUATION);
  for (int j = 0; j < NU; j++) {
    real vmax = (real)m->servo_vmax[j], amax = (real)m->servo_amax[j];
    real err = action[j] - pos[j];
    real thr = err >= 0 ? (real)m->servo_deadband_pos : (real)m->servo_deadband_neg;
    int in_db = R_ABS(err) <= thr;
    real db_v = vel[j] * (real)0.8;
    real db_p = pos[j] + db_v * dt;
    real tdir = sign_(err);
    real stop = R_ABS(vel[j] * vel[j]) / (2 * amax);
    real vdir = sign_(vel[j]);
    int towards = vdir * tdir >= 0;
    int accel = towards && R_ABS(err) > stop;
    real acc = accel ? tdir * amax : -vdir * amax;
    if (R_ABS(vel[j]) < (real)1e-6) acc = tdir * amax;
    real pv = vel[j] + acc * dt;
    pv = R_FMIN(R_FMAX(pv, -vmax), vmax);
    real pp = pos[j] + pv * dt;
    real nv = in_db ? db_v : pv, np = in_db ? db_p : pp;
    pos[j] = np;
    vel[j] = nv;
    real perr = np - qpos[7 + j], verr = nv - qvel[6 + j];
    real duty = (real)m->servo_kp[j] * (real)m->servo_error_gain[j] * perr + (real)m->servo_kd[j] * verr;
    duty = R_FMIN(R_FMAX(duty, -(real)m->servo_max_pwm[j]), (real)m->servo_max_pwm[j]);
    real volts = duty * (real)m->servo_vin[j];
    real tq = volts * (real)m->servo_kt[j] / (real)m->servo_R[j];
    torque[j] = tq;
    ctrl[j] = tq; /* torque_noise_type "none": train.py:1433-1436 [V] */
    FL(30);
  }
}

static int finite_state(const ZboData* d) {
  for (int i = 0; i < NQ; i++) if (!isfinite((double)d->qpos[i])) return 0;
  for (int i = 0; i < NV; i++) if (!isfinite((double)d->qvel[i])) return 0;
  return 1;
}

static void write_out(const ZboOut* out, const ZboData* d, int i) {
  if (!out) return;
#define CP(field, src, n) if (out->field) memcpy(out->field + (size_t)i * (n), src, sizeof(real) * (n))
  CP(xpos, d->xpos, NB * 3);
  CP(xquat, d->xquat, NB * 4);
  CP(cinert, d->cinert, NB * 10);
  CP(cvel, d->cvel, NB * 6);
  CP(qacc, d->qacc, NV);
  CP(qfrc_constraint, d->qfrc_constraint, NV);
  CP(actuator_force, d->actuator_force, NU);
  CP(ctrl, d->ctrl, NU);
  CP(sensordata, d->sensordata, ZB_NSENSORDATA);
  CP(site_xpos, d->site_xpos, NS * 3);
  CP(site_xmat, d->site_xmat, NS * 9);
  CP(contact_dist, d->con_dist, MAXCON);
  CP(contact_pos, d->con_pos, MAXCON * 3);
  CP(efc_force, d->efc_force, NEFC);
  CP(solver_trace, d->trace, ZB_NTRACE);
#undef CP
  if (out->contact_mask) {
    int mask = 0;
    for (int c = 0; c < MAXCON; c++) mask |= d->con_active[c] << c;
    out->contact_mask[i] = mask;
  }
  if (out->solver_niter) out->solver_niter[i] = d->solver_niter;
  if (out->flags) out->flags[i] = d->flags;
}

static void load_state(ZboData* d, const ZboState* st, int i) {
  memset(d, 0, sizeof(*d));
  memcpy(d->qpos, st->qpos + (size_t)i * NQ, sizeof(d->qpos));
  memcpy(d->qvel, st->qvel + (size_t)i * NV, sizeof(d->qvel));
  if (st->qacc_warmstart) memcpy(d->qacc_warmstart, st->qacc_warmstart + (size_t)i * NV, sizeof(d->qacc_warmstart));
}
static void store_state(const ZboData* d, const ZboState* st, int i) {
  memcpy(st->qpos + (size_t)i * NQ, d->qpos, sizeof(d->qpos));
  memcpy(st->qvel + (size_t)i * NV, d->qvel, sizeof(d->qvel));
  if (st->qacc_warmstart) memcpy(st->qacc_warmstart + (size_t)i * NV, d->qacc_warmstart, sizeof(d->qacc_warmstart));
}

#define SUF2(a, b) a##b
#if ZBO_F64
#define SUF(a) SUF2(a, _f64)
#else
#define SUF(a) SUF2(a, _f32)
#endif

/* mjx.forward for one environment with explicit ctrl; fills the full ZboData (tests). */
int SUF(zbo_forward_one)(const ZbModel* m, const ZbOpts* o, const ZboParams* p, int env, const real* qpos, const real* qvel,
                         const real* ctrl, const real* warmstart, ZboData* d) {
  EnvModel e;
  env_model(&e, m, p, env);
  memset(d, 0, sizeof(*d));
  memcpy(d->qpos, qpos, sizeof(d->qpos));
  memcpy(d->qvel, qvel, sizeof(d->qvel));
  if (ctrl) memcpy(d->ctrl, ctrl, sizeof(d->ctrl));
  if (warmstart) memcpy(d->qacc_warmstart, warmstart, sizeof(d->qacc_warmstart));
  forward(m, o, &e, d);
  return 0;
}
/* mjx.step for one environment with explicit ctrl (no actuator model, no events). */
int SUF(zbo_step_one)(const ZbModel* m, const ZbOpts* o, const ZboParams* p, int env, ZboData* d) {
  EnvModel e;
  env_model(&e, m, p, env);
  forward(m, o, &e, d);
  integrate(m, o, &e, d);
  return 0;
}

/* MjxEngine.step [R]: S x { PushEvent -> FeetechActuators -> mjx.step } */
int SUF(zbo_step)(const ZbModel* m, const ZbOpts* o, const ZboParams* p, const ZboState* st, const real* action,
                  const real* push_rand, int push_per_substep, const ZboOut* out, int n_envs, int n_substeps, int n_threads) {
  if (!m || !o || !st || !action || n_envs < 0 || n_substeps < 1) return ZB_EINVAL;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
  for (int i = 0; i < n_envs; i++) {
    ZboData* d = (ZboData*)malloc(sizeof(ZboData));
    EnvModel e;
    env_model(&e, m, p, i);
    load_state(d, st, i);
    real dt = (real)o->dt;
    real* apos = st->act_pos + (size_t)i * NU;
    real* avel = st->act_vel + (size_t)i * NU;
    real* atq = st->act_torque + (size_t)i * NU;
    for (int s = 0; s < n_substeps; s++) {
      for (int phase = 0; phase < 2; phase++) {
        int do_event = (o->substep_order == ZB_ORDER_EVENT_ACT_STEP) ? phase == 0 : phase == 1;
        if (do_event) {
          if (push_rand && st->push_t_rem) { /* PushEvent train.py:1459-1468 [V] config; ksim/events.py [R] semantics */
            const real* pr = push_rand + ((size_t)i * (push_per_substep ? n_substeps : 1) + (push_per_substep ? s : 0)) * 7;
            real t = st->push_t_rem[i] - dt;
            if (t <= 0) {
              for (int k = 0; k < 6; k++) d->qvel[k] += pr[k];
              t = pr[6];
            }
            st->push_t_rem[i] = t;
          }
        } else {
          const real* tgt = action + (size_t)i * NU; /* action latency: early substeps keep the previous action */
          if (p && p->prev_action && p->latency_steps && s < p->latency_steps[i]) tgt = p->prev_action + (size_t)i * NU;
          feetech(m, dt, tgt, d->qpos, d->qvel, apos, avel, atq, d->ctrl);
        }
      }
      forward(m, o, &e, d);
      integrate(m, o, &e, d);
      if (st->time) st->time[i] += dt;
    }
    if (!finite_state(d)) d->flags |= 1;
    store_state(d, st, i);
    write_out(out, d, i);
    free(d);
  }
  return 0;
}

/* ONE physics step like zbo_step (n_substeps = 1, no PushEvent), but the solver takes the STEP SIZES of another implementation's
 * solver trace (ZbOut.solver_trace layout: alpha of iteration i at [4 i]) and runs exactly niter[i] iterations: everything except
 * the line search's bracket decisions — the one chaotic element of the capped solver, DESIGN.md §5 — is then a continuous function
 * of the inputs, and the traced implementation must agree with this fp64 evaluation at rounding level on EVERY environment.
 * force_start [n] (or NULL): 0 = own warm-start choice, 1 = qacc_warmstart, 2 = qacc_smooth; start_costs [n, 2] (or NULL). */
int SUF(zbo_step_follow)(const ZbModel* m, const ZbOpts* o, const ZboParams* p, const ZboState* st, const real* action, const ZboOut* out,
                         int n_envs, int n_threads, const real* trace, const int32_t* niter, const int32_t* force_start, real* start_costs) {
  if (!m || !o || !st || !action || !trace || !niter || n_envs < 0) return ZB_EINVAL;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
  for (int i = 0; i < n_envs; i++) {
    ZboData* d = (ZboData*)malloc(sizeof(ZboData));
    EnvModel e;
    env_model(&e, m, p, i);
    load_state(d, st, i);
    real dt = (real)o->dt;
    feetech(m, dt, action + (size_t)i * NU, d->qpos, d->qvel, st->act_pos + (size_t)i * NU, st->act_vel + (size_t)i * NU,
            st->act_torque + (size_t)i * NU, d->ctrl);
    g_follow_trace = trace + (size_t)i * ZB_NTRACE;
    g_follow_niter = niter[i];
    g_force_start = force_start ? force_start[i] : 0;
    g_start_costs = start_costs ? start_costs + 2 * (size_t)i : NULL;
    forward(m, o, &e, d);
    g_follow_trace = NULL; g_force_start = 0; g_start_costs = NULL;
    integrate(m, o, &e, d);
    if (st->time) st->time[i] += dt;
    if (!finite_state(d)) d->flags |= 1;
    store_state(d, st, i);
    write_out(out, d, i);
    free(d);
  }
  return 0;
}

/* Derivatives of the FIRST line search's 1-D cost at given step sizes (alphas [n, K]) for the state + action a step would start
 * from (the state is not modified): d0, d1 [n, K] as solver._linesearch's point function defines them, scale [n, K] = magnitude of
 * the summed terms of d0, margin [n, K] = relative distance of the nearest constraint row to a change of its quadratic piece,
 * cost [n, K] (or NULL) = the 1-D cost itself.  search_in [n, NV] (or NULL): take this first search direction (another implementation's,
 * ZbOut.solver_trace[ZB_TRACE_SEARCH..]) instead of the oracle's own -M^-1 grad; search_out [n, NV] (or NULL) receives the oracle's own;
 * info [n, 4] (or NULL): see ls_query. */
int SUF(zbo_ls_points)(const ZbModel* m, const ZbOpts* o, const ZboParams* p, const ZboState* st, const real* action, int n_envs,
                       int n_threads, const int32_t* force_start, const real* alphas, int K, real* d0, real* d1, real* scale,
                       real* margin, real* cost, real* start_costs, const real* search_in, real* search_out, real* info) {
  if (!m || !o || !st || !action || !alphas || K < 1 || n_envs < 0) return ZB_EINVAL;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
  for (int i = 0; i < n_envs; i++) {
    ZboData* d = (ZboData*)malloc(sizeof(ZboData));
    EnvModel e;
    env_model(&e, m, p, i);
    load_state(d, st, i);
    real pos[NU], vel[NU], tq[NU];
    memcpy(pos, st->act_pos + (size_t)i * NU, sizeof(pos));
    memcpy(vel, st->act_vel + (size_t)i * NU, sizeof(vel));
    memcpy(tq, st->act_torque + (size_t)i * NU, sizeof(tq));
    feetech(m, (real)o->dt, action + (size_t)i * NU, d->qpos, d->qvel, pos, vel, tq, d->ctrl);
    LSQuery Q = {K, alphas + (size_t)i * K, d0 + (size_t)i * K, d1 + (size_t)i * K, scale + (size_t)i * K, margin + (size_t)i * K, cost ? cost + (size_t)i * K : NULL,
                 search_out ? search_out + (size_t)i * NV : NULL, info ? info + 4 * (size_t)i : NULL};
    g_ls_query = &Q;
    g_search_override = search_in ? search_in + (size_t)i * NV : NULL;
    g_force_start = force_start ? force_start[i] : 0;
    g_start_costs = start_costs ? start_costs + 2 * (size_t)i : NULL;
    forward(m, o, &e, d);
    g_ls_query = NULL; g_search_override = NULL; g_force_start = 0; g_start_costs = NULL;
    free(d);
  }
  return 0;
}

int SUF(zbo_forward)(const ZbModel* m, const ZbOpts* o, const ZboParams* p, const ZboState* st, const ZboOut* out, int n_envs, int n_threads) {
  if (!m || !o || !st || n_envs < 0) return ZB_EINVAL;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
  for (int i = 0; i < n_envs; i++) {
    ZboData* d = (ZboData*)malloc(sizeof(ZboData));
    EnvModel e;
    env_model(&e, m, p, i);
    load_state(d, st, i);
    if (st->act_torque) memcpy(d->ctrl, st->act_torque + (size_t)i * NU, sizeof(d->ctrl));
    forward(m, o, &e, d);
    memcpy(st->qpos + (size_t)i * NQ, d->qpos, sizeof(d->qpos)); /* normalised free-joint quaternion */
    write_out(out, d, i);
    free(d);
  }
  return 0;
}

/* MjxEngine.reset [R] + get_initial_state train.py:1293-1298 [V] */
int SUF(zbo_reset)(const ZbModel* m, const ZbOpts* o, const ZboParams* p, const ZboState* st, const uint8_t* done,
                   const real* reset_qpos, const real* reset_qvel, const real* reset_t_rem, const ZboOut* out, int n_envs, int n_threads) {
  if (!m || !o || !st || !reset_qpos || !reset_qvel || n_envs < 0) return ZB_EINVAL;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads > 0 ? n_threads : 1)
  for (int i = 0; i < n_envs; i++) {
    if (done && !done[i]) continue;
    ZboData* d = (ZboData*)malloc(sizeof(ZboData));
    EnvModel e;
    env_model(&e, m, p, i);
    memset(d, 0, sizeof(*d));
    memcpy(d->qpos, reset_qpos + (size_t)i * NQ, sizeof(d->qpos));
    memcpy(d->qvel, reset_qvel + (size_t)i * NV, sizeof(d->qvel));
    for (int j = 0; j < NU; j++) {
      st->act_pos[(size_t)i * NU + j] = d->qpos[7 + j];
      st->act_vel[(size_t)i * NU + j] = d->qvel[6 + j];
      st->act_torque[(size_t)i * NU + j] = 0;
    }
    if (st->push_t_rem) st->push_t_rem[i] = reset_t_rem ? reset_t_rem[i] : 0;
    if (st->time) st->time[i] = 0;
    forward(m, o, &e, d);
    memset(d->qacc_warmstart, 0, sizeof(d->qacc_warmstart));
    store_state(d, st, i);
    write_out(out, d, i);
    free(d);
  }
  return 0;
}

/* Observation epilogue (include/zbphys.h ZbObs): the network input vectors of the reference, from the arrays zbo_step /
 * zbo_reset produced.  [V] run_actor train.py:1624-1648, run_critic train.py:1650-1681, ImuOrientationObservation
 * .observe_stateful train.py:837-866 (carry train.py:846-848), rotate_quat_by_quat train.py:751-787;
 * [R] ksim's JointPosition/JointVelocity/CenterOfMassInertia/CenterOfMassVelocity/ActuatorForce/BasePosition/
 * BaseOrientation/SensorObservation = qpos[7:], qvel[6:], cinert[1:], cvel[1:], actuator_force, qpos[0:3], qpos[3:7],
 * sensordata[range]; xax.euler_to_quat([0, 0, yaw]) = (cos(yaw/2), 0, 0, sin(yaw/2)); additive observation noise.
 * reinit[i] != 0: carry of environment i restarts from x = 0, lag = reset_lag[i] before this observation. */
int SUF(zbo_observe)(int n_envs, const real* qpos, const real* qvel, const real* cinert, const real* cvel, const real* sensordata,
                     const real* actuator_force, const real* command, real* imu_x, real* imu_lag, const uint8_t* reinit,
                     const real* reset_lag, const real* noise_quat, const real* noise_acc, real* actor_obs, real* critic_obs) {
  if (n_envs < 0 || !qpos || !qvel || !sensordata || !imu_x || !imu_lag) return ZB_EINVAL;
  const real eps = (real)1e-6;
  for (int i = 0; i < n_envs; i++) {
    const real* q = qpos + (size_t)i * NQ;
    const real* v = qvel + (size_t)i * NV;
    const real* sd = sensordata + (size_t)i * ZB_NSENSORDATA;
    real cmd[ZB_NCOMMAND];
    for (int k = 0; k < ZB_NCOMMAND; k++) cmd[k] = command ? command[(size_t)i * ZB_NCOMMAND + k] : 0;
    /* heading quaternion and its conjugate; both operands normalised as q / (|q| + eps) */
    real fq[4], hq[4] = {R_COS((real)0.5 * cmd[3]), 0, 0, R_SIN((real)0.5 * cmd[3])};
    const real* sq = sd + ZB_SENS_IMU_QUAT;
    real nf = R_SQRT(sq[0] * sq[0] + sq[1] * sq[1] + sq[2] * sq[2] + sq[3] * sq[3]) + eps;
    for (int k = 0; k < 4; k++) fq[k] = sq[k] / nf;
    real nh = R_SQRT(hq[0] * hq[0] + hq[1] * hq[1] + hq[2] * hq[2] + hq[3] * hq[3]) + eps;
    for (int k = 0; k < 4; k++) hq[k] = hq[k] / nh;
    for (int k = 1; k < 4; k++) hq[k] = -hq[k];
    real r[4];
    r[0] = hq[0] * fq[0] - hq[1] * fq[1] - hq[2] * fq[2] - hq[3] * fq[3];
    r[1] = hq[0] * fq[1] + hq[1] * fq[0] + hq[2] * fq[3] - hq[3] * fq[2];
    r[2] = hq[0] * fq[2] - hq[1] * fq[3] + hq[2] * fq[0] + hq[3] * fq[1];
    r[3] = hq[0] * fq[3] + hq[1] * fq[2] - hq[2] * fq[1] + hq[3] * fq[0];
    real nr = R_SQRT(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]) + eps;
    for (int k = 0; k < 4; k++) r[k] = r[k] / nr;
    if (r[0] < 0) for (int k = 0; k < 4; k++) r[k] = -r[k];
    real* x = imu_x + (size_t)i * 4;
    if (reinit && reinit[i]) {
      for (int k = 0; k < 4; k++) x[k] = 0;
      if (reset_lag) imu_lag[i] = reset_lag[i];
    }
    const real lag = imu_lag[i];
    real ob[4], acc[3];
    for (int k = 0; k < 4; k++) { x[k] = x[k] * lag + r[k] * ((real)1 - lag); ob[k] = x[k] + (noise_quat ? noise_quat[(size_t)i * 4 + k] : 0); }
    for (int k = 0; k < 3; k++) acc[k] = sd[ZB_SENS_IMU_ACC + k] + (noise_acc ? noise_acc[(size_t)i * 3 + k] : 0);
    if (actor_obs) {
      real* a = actor_obs + (size_t)i * ZB_NACTOR_OBS;
      int t = 0;
      for (int j = 0; j < NU; j++) a[t++] = q[7 + j];
      for (int j = 0; j < NU; j++) a[t++] = v[6 + j];
      for (int k = 0; k < 4; k++) a[t++] = ob[k];
      a[t++] = cmd[0]; a[t++] = cmd[1]; a[t++] = cmd[3]; a[t++] = cmd[4]; a[t++] = cmd[5]; a[t++] = cmd[6];
    }
    if (critic_obs) {
      if (!cinert || !cvel || !actuator_force) return ZB_EINVAL;
      real* c = critic_obs + (size_t)i * ZB_NCRITIC_OBS;
      int t = 0;
      for (int j = 0; j < NU; j++) c[t++] = q[7 + j];
      for (int j = 0; j < NU; j++) c[t++] = v[6 + j] / (real)10;
      for (int k = 0; k < (NB - 1) * 10; k++) c[t++] = cinert[(size_t)i * NB * 10 + 10 + k];
      for (int k = 0; k < (NB - 1) * 6; k++) c[t++] = cvel[(size_t)i * NB * 6 + 6 + k];
      for (int k = 0; k < 3; k++) c[t++] = acc[k];
      for (int k = 0; k < 3; k++) c[t++] = sd[ZB_SENS_IMU_GYRO + k];
      for (int k = 0; k < 4; k++) c[t++] = ob[k];
      for (int k = 0; k < ZB_NCOMMAND; k++) c[t++] = cmd[k];
      for (int j = 0; j < NU; j++) c[t++] = actuator_force[(size_t)i * NU + j] / (real)100;
      for (int k = 0; k < 7; k++) c[t++] = q[k];
    }
  }
  return 0;
}

/* Reward inputs (include/zbphys.h ZbObs.feet_pos_obs / reward_terms).  [V] FeetPositionObservation train.py:430-465,
 * SimpleSingleFootContactReward train.py:691-713 (stand_still_threshold=None, train.py:1559), FeetOrientationReward
 * train.py:647-689, FeetTooClosePenalty train.py:729-748, BaseHeightReward train.py:468-483, contact flags of
 * FeetAirtimeReward train.py:524-525; [R] xax.quat_to_euler (roll = atan2(2(wx+yz), 1-2(x^2+y^2)), pitch = asin(2(wy-zx))
 * clamped to +-pi/2) and xax.rotate_vector_by_quat(v, q, inverse=True) = R(q/|q|)^T v.
 * left_body / right_body: body ids of the left_foot / right_foot sites. */
int SUF(zbo_reward_inputs)(int n_envs, const real* qpos, const real* xpos, const real* xquat, const real* site_xpos, const real* sensordata,
                           const real* command, int left_body, int right_body, real* feet_pos_obs, real* reward_terms) {
  if (n_envs < 0 || !qpos || !xpos || !xquat || !site_xpos || !sensordata) return ZB_EINVAL;
  const real eps = (real)1e-6;
  for (int i = 0; i < n_envs; i++) {
    const real* q = qpos + (size_t)i * NQ;
    real bq[4];
    real nq = R_SQRT(q[3] * q[3] + q[4] * q[4] + q[5] * q[5] + q[6] * q[6]) + eps;
    for (int k = 0; k < 4; k++) bq[k] = q[3 + k] / nq;
    real R[9], feet[6];
    quat_to_mat(R, bq);
    for (int f = 0; f < 2; f++) matT_vec(feet + 3 * f, R, site_xpos + ((size_t)i * NS + 1 + f) * 3);
    if (feet_pos_obs) memcpy(feet_pos_obs + (size_t)i * 6, feet, sizeof(feet));
    if (reward_terms) {
      const real* sd = sensordata + (size_t)i * ZB_NSENSORDATA;
      const int cl = sd[ZB_SENS_TOUCH_L] > (real)0.1, cr = sd[ZB_SENS_TOUCH_R] > (real)0.1;
      real err = 0;
      const int bodies[2] = {left_body, right_body};
      for (int f = 0; f < 2; f++) {
        const real* qq = xquat + ((size_t)i * NB + bodies[f]) * 4;
        real n2 = R_SQRT(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]) + eps;
        real w = qq[0] / n2, x = qq[1] / n2, y = qq[2] / n2, z = qq[3] / n2;
#if ZBO_F64
        real roll = atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y));
        real sp = 2 * (w * y - z * x);
        real pitch = fabs(sp) >= 1 ? copysign(1.57079632679489662, sp) : asin(sp);
#else
        real roll = atan2f(2 * (w * x + y * z), 1 - 2 * (x * x + y * y));
        real sp = 2 * (w * y - z * x);
        real pitch = fabsf(sp) >= 1 ? copysignf(1.57079632679489662f, sp) : asinf(sp);
#endif
        err += R_ABS(roll) + R_ABS(pitch);
      }
      real d[3] = {feet[0] - feet[3], feet[1] - feet[4], feet[2] - feet[5]};
      real* t = reward_terms + (size_t)i * ZB_NREWARD_TERMS;
      real c4 = command ? command[(size_t)i * ZB_NCOMMAND + 4] : 0;
#if ZBO_F64
      t[1] = exp(-err / 0.25);
      t[3] = exp(-fabs(xpos[((size_t)i * NB + 1) * 3 + 2] - (c4 + 0.27)) / 0.25);
#else
      t[1] = expf(-err / 0.25f);
      t[3] = expf(-fabsf(xpos[((size_t)i * NB + 1) * 3 + 2] - (c4 + 0.27f)) / 0.25f);
#endif
      t[0] = (cl != cr) ? 1 : 0;
      t[2] = R_SQRT(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < (real)0.12 ? 1 : 0;
      t[4] = cl; t[5] = cr;
    }
  }
  return 0;
}

int SUF(zbo_sizeof_data)(void) { return (int)sizeof(ZboData); }

/* bit 0: contact rows applied matrix-free like the CUDA kernel; bit 1: the line search stops once a bracket end sits on an exactly
 * zero derivative (experiment); bit 2: collect line-search statistics (single-threaded) */
void SUF(zbo_set_variant)(int v) { g_variant = v; }
void SUF(zbo_lsstats_get)(double* out, int reset) {
  for (int k = 0; k < 8; k++) { out[k] = g_lsstat[k]; if (reset) g_lsstat[k] = 0; }
}

#if !ZBO_F64
/* flop counters (meaningful only in the -DZBO_COUNT build, single-threaded) */
void zbo_flops_reset(void) {
#ifdef ZBO_COUNT
  memset(g_flops, 0, sizeof(g_flops));
#endif
}
void zbo_flops_get(double* out) {
  for (int s = 0; s < ZBO_NSTAGE; s++)
#ifdef ZBO_COUNT
    out[s] = g_flops[s];
#else
    out[s] = -1;
#endif
}
#endif
