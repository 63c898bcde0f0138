// zb_kernel.cuh — the fused Z-Bot physics step: ONE WARP PER ENVIRONMENT.
//
// What it replaces: ksim MjxEngine.step / reset (ksim 0.1.99, un-vendored:
// /root/reference/uv.lock:1872-1873) around mjx.step (mujoco-mjx 3.3.4, uv.lock:2226-2227), with the
// reference's own servo model fused in (/root/reference/train.py:1137-1196 trapezoidal_step,
// :1242-1280 FeetechActuators.get_stateful_ctrl) and PushEvent (train.py:1459-1468).
//
// Lane mapping (ZB_NBODY + 6 == 32):
//   lane b  in [0, 26)  : body b; if the body has a hinge joint the lane also owns that dof
//   lane 26+a, a in [0,6): the six dofs of the free joint (base body = lane 1)
//   lane r  in [0, 32)  : contact row r = pyramid edge (r & 3) of contact slot (r >> 2); box = r >> 4
// Tree structure exploited: floating base + serial chains hanging off it (validated on the host).
//   chain scans  : Hillis-Steele over contiguous body lanes with shfl_up / shfl_down
//   mass matrix  : M = [[A, B^T], [B, C]], C block-diagonal per chain -> block LDL^T
//                  (chain factors in parallel, Schur complement S = A - Wt^T D^-1 Wt on the base lanes)
//   execution    : 16 (128 registers) or 28 (72 registers) environments per CTA, chosen per launch; the warps meet at a barrier
//                  at every substep so that they stream the same instructions (instruction fetch was the top stall of the
//                  first version, see profiles/README.md)
//   constraints  : friction-loss / limit rows are single-entry rows owned by the dof lane; contact rows
//                  are applied matrix-free: J[r, :] = w_r . cdof[:] (w_r = contact wrench direction), so
//                  J v = w_r . V_foot(v) and J^T f = cdof . sum_r f_r w_r.
#pragma once
#include <stdint.h>

#include "../../include/zbphys.h"
#include "zb_math.cuh"

#define NQ ZB_NQ
#define NV ZB_NV
#define NU ZB_NU
#define NB ZB_NBODY
#define BL0 (32 - 6) /* first base-dof lane */
#define CHD 6        /* dof positions per chain group (legs: 6, arms: 4) */
#ifndef ZB_WPB
#define ZB_WPB 16 /* latency profile: 16 warps (environments) per 512-thread CTA, one CTA per SM at 128 registers */
#endif
#ifndef ZB_WPB_WIDE
#define ZB_WPB_WIDE 28 /* throughput profile: 28 warps per 896-thread CTA at 72 registers (spills, but 4096 envs fit ONE wave of 148 SMs) */
#endif
#ifndef ZB_CTA_LOCKSTEP
#define ZB_CTA_LOCKSTEP 1 /* 1: the warps of a CTA meet at every substep (shared instruction fetch: +30 %).  (A barrier before every solver pass as well was measured slower — it waits for the slowest line search — and has been removed.) */
#endif
#define ZB_SCR 8     /* scratch floats per lane */
#ifndef ZB_MUTANT
#define ZB_MUTANT 0 /* mutation builds (build.py MUTANTS): one planted bug each, the parity tests must catch every one */
#endif
#ifndef ZB_EXPLICIT_J
#define ZB_EXPLICIT_J 1 /* contact rows as explicit 12-column Jacobian rows (chain dofs + base dofs) instead of matrix-free wrench directions */
#endif
#ifndef ZB_RECOLLIDE
#define ZB_RECOLLIDE 0 /* 1: sensors() re-run collision() instead of keeping contact point + frame live across the solve (more code, no gain measured) */
#endif
#ifndef ZB_CJ_SMEM
#define ZB_CJ_SMEM 0 /* throughput profile: the 12 Jacobian-row entries of a lane's contact row in shared memory too (43 instead of 31 floats per lane) */
#endif
#ifndef ZB_LS_FIRST
#define ZB_LS_FIRST 1 /* line search: the first Newton point reuses point(0)'s summed quadratic when every row sits on the same piece there (vote instead of a reduction; bit-exact) */
#endif
#ifndef ZB_PRED
#define ZB_PRED 1 /* lane-masked updates as predicated FFMA (inline PTX; ptxas otherwise emits FFMA + FSEL) and lane masks folded into one factor instead of one select per product */
#endif
#ifndef ZB_LS_SAME
#define ZB_LS_SAME 2 /* line search: 1 = no zone evaluation while both bracket ends share one active set, 2 = + the one-sided loop (see linesearch) */
#endif

static_assert(ZB_NBODY + 6 == 32, "lane mapping assumes 26 bodies + 6 free-joint dofs == one warp");

#define MJ_MINVAL 1e-15f
#define MJ_MINIMP 0.0001f
#define MJ_MAXIMP 0.9999f

enum { ZB_MODE_STEP = 0, ZB_MODE_FORWARD = 1, ZB_MODE_RESET = 2 };

// constraint._kbi constants of one (solref, solimp) pair, precomputed on the host (depend on opt.timestep via refsafe)
struct ZbSolPrm {
  float k, b, dmin, dmax, width, mid, power, c1, c2;  // c1 = 1/mid^(power-1), c2 = 1/(1-mid)^(power-1)
};

// Host-derived tables (zb_engine_create) living in device global memory next to the model.
struct ZbDev {
  ZbModel m;
  ZbOpts o;
  int cs[32], ce[32];   // body-chain start / end lane (self for world, base, base-dof lanes)
  int gs[32], gp[32];   // dof-group start lane and my position in it (chains: cs; base dofs: BL0)
  int gn[32];           // dof positions in my group (0: lane owns no dof)
  int lane_dof[32];     // dof index owned by the lane, -1 none
  int lane_act[32];     // actuator index driving that dof, -1 none
  int lane_box[32];     // hinge lanes: foot box whose body descends from my dof, -1 none
  int body_site[32];    // site carried by body lane, -1 none
  int body_box[32];     // box geom carried by body lane, -1 none
  float bpj[32][3];     // body_pos + R(body_quat) jnt_pos
  float frame[9];       // make_frame(floor_normal): rows n, t1, t2
  int box_site[ZB_NBOX];
  int any_damping;
  float fr_R[32], fr_b[32];        // friction-loss rows by lane: R (0 = no row) and b of the dof's solref/solimp
  ZbSolPrm lim[32];                // joint-limit rows by lane
  ZbSolPrm con[ZB_NBOX];           // contact rows by box (after geom-priority mixing)
  float con_includemargin[ZB_NBOX];
  float con_invweight[ZB_NBOX];    // body_invweight0[world] + body_invweight0[box body]
  int con_mu_mode[ZB_NBOX];        // 0: floor friction (per-env), 1: con_mu_const, 2: max of both
  float con_mu_const[ZB_NBOX];     // friction of the group after mixing, where it does not depend on the per-env floor friction
  // collision groups (ZbModel.box_* / pair_*): primitive of each group and what the model contains at all (uniform branches)
  int grp_kind[ZB_NBOX];           // ZB_PAIR_*
  int any_box, any_prim, any_pair; // a plane-box group / a sphere or capsule group / a body-body group exists
  int body_pair[32];               // group whose geom 1 this body lane carries, -1 none
  int lane_jt[32][ZB_NBOX];        // dof lanes: index of my entry in group g's transposed J^T f totals (0..5 chain of geom 2,
                                   // 6..11 chain of geom 1 / base dofs), -1: the group does not touch my dof
};
enum { ZB_PAIR_PLANE_BOX = 0, ZB_PAIR_PLANE_SPHERE, ZB_PAIR_PLANE_CAPSULE, ZB_PAIR_SPHERE_SPHERE, ZB_PAIR_SPHERE_CAPSULE, ZB_PAIR_CAPSULE_CAPSULE };

struct Factor {      // block LDL^T of M (or M + dt D); every entry is zero where it does not apply to the lane
  float lc[CHD];  // Lt[gp][k], k < gp: row of the unit lower factor of the lane's group (chain block or base block)
  float lt[CHD];  // Lt[k][gp], k > gp: column
  float inv;      // 1 / D[gp]
  float W[6];     // hinge lanes: row of G = D_c^-1 Lt_c^-1 B
};

struct LSPoint {
  float alpha, d0;  // step size and first derivative of the 1-D cost model (the second derivative belongs to the active set: LSSet.d1)
};
struct LSSet {          // an active set of the line search: its summed quadratic and this lane's zone bits
  float t0, t1, t2;
  float d1;            // second derivative 2 t2 (+ mjMINVAL if t2 == 0): constant on the set
  int z;
};

// TRACE: compile the solver-trace diagnostics in (scripts/gpu_probe*.py); the production instantiation carries none of it.
// FSM: the factor and the mass-matrix rows of a lane live in shared memory (throughput profile) instead of registers.
// GEN: general collision groups (sphere / capsule primitives, geom 1 on a body).  GEN = false is the instantiation for models whose
// groups are all plane-box (zbot_syn): the frame is the constant floor frame and none of the extra per-lane state exists — the
// general code costs 3.7 % there (measured), so it is compiled separately and chosen per engine.
template <bool TRACE, bool FSM, bool GEN>
struct WarpEnv {
  const ZbDev* D;
  ZbOpts o;           // by value from the kernel parameters: uniform-datapath registers, provably warp-uniform control flow
  bool any_box, any_prim, any_pair;  // ZbDev.any_*, handed in as a kernel parameter for the same reason (a flag LOADED from the model
                                     // table is not provably uniform: ptxas then wraps every later shuffle in WARPSYNC and clones the solver)
  float meaninertia;
  float* scr;  // per-warp scratch, 32 * ZB_SCR floats
  int L;
  // topology of this lane
  int cs, ce, gs, gp, gn, dof, act, lbox, boxlane;
  bool isbody, hinge, bdof, hasdof, chainbody;
  // per-environment model (randomizers, train.py:1439-1455)
  float mass, inertia[3], arm, damp, floss, q0, mu, mass_tot, inv_mass_tot;
  float spos[3], squat[4];
  // state
  float q, qv, warm, bq[7];
  float apos, avel, atq, tgt, ctrl;
  // kinematics
  float xpos[3], xquat[4], xmat[9], xipos[3], xb[3], Rb[9], c0[3];
  float cinert[10], cdof[6], cdofd[6], cvel[6];
  float sxpos[3], sxmat[9];
  // dynamics
  // Mass-matrix rows (Bm[6], C[6]) and the factor F0 (lc[6], lt[6], inv, W[6]) of this lane: 31 floats that are read once or
  // twice per solver pass.  In the 128-register profile they sit in registers (fst).  In the 72-register profile ptxas spilled
  // exactly these (17 local loads in front of every solve, profiles/r1_final_hotspots.txt) and the 666 KB of stack per SM
  // thrash the L1; there they live in shared memory instead (fsm: lane column, element i at fsm[32 i], conflict-free).
  // F0 = factor of M, replaced by the Newton Hessian's / by the factor of M + dt D when those are needed (solve_all).
  enum { F_LC = 0, F_LT = 6, F_INV = 12, F_W = 13, F_BM = 19, F_C = 25, F_CJ = 31, F_N = 31 + (ZB_CJ_SMEM ? 12 : 0) };
  float fst[FSM ? 1 : F_N];
  float* fsm;
  ZB_DI float fget(int i) const { return FSM ? fsm[32 * i] : fst[FSM ? 0 : i]; }
  ZB_DI void fput(int i, float v) { if (FSM) fsm[32 * i] = v; else fst[FSM ? 0 : i] = v; }
  ZB_DI void load_factor(Factor& F) const {
#pragma unroll
    for (int k = 0; k < CHD; k++) { F.lc[k] = fget(F_LC + k); F.lt[k] = fget(F_LT + k); F.W[k] = fget(F_W + k); }
    F.inv = fget(F_INV);
  }
  ZB_DI void load_m(float (&Bm)[6], float (&C)[CHD]) const {
#pragma unroll
    for (int k = 0; k < CHD; k++) { Bm[k] = fget(F_BM + k); C[k] = fget(F_C + k); }
  }
  float qfrc_smooth, qacc_smooth, act_force, qfrc_bias;
  // constraint rows
  bool fr_on, lm_on, c_on, any_limit;
  float fr_D, fr_rf, fr_f, fr_aref, lm_sign, lm_D, lm_aref, c_D, c_aref, w[6];
  // Explicit Jacobian row of this lane's contact row.  Tree sparsity: the row of a foot contact touches only the (<= 6) dofs of its
  // leg chain and the 6 base dofs, J[r, d] = w_r . cdof_d.  cj[0..5]: chain dofs of the box (lanes jsrc + k, zero for non-ancestors),
  // cj[6..11]: base dofs.  With it J v is 6 shuffles + 12 FMAs (instead of a 6-vector chain prefix scan + the base motion) and
  // neither cdof, w nor the base pose are read inside the solver loop.
  float cjst[(FSM && ZB_CJ_SMEM) ? 1 : 12];
  ZB_DI float cjget(int k) const { return (FSM && ZB_CJ_SMEM) ? fsm[32 * (F_CJ + k)] : cjst[(FSM && ZB_CJ_SMEM) ? 0 : k]; }
  ZB_DI void cjput(int k, float v) { if (FSM && ZB_CJ_SMEM) fsm[32 * (F_CJ + k)] = v; else cjst[(FSM && ZB_CJ_SMEM) ? 0 : k] = v; }
  int jsrc, jsrc2;      // first lane of the chain of geom 2 / of geom 1 (BL0: the base dofs, geom 1 = the floor)
  int pairlane;         // body lane of geom 1 of my group (0 = world)
  float cn[3], ct[3];   // contact normal and the tangent of this lane's pyramid edge (collision())
  float cdist, cpos[3];
  // solver
  float qacc, Ma, grad, Mgrad, search, mv, qfrc_c;
  float fr_ja, lm_ja, c_ja;
  float fr_force, lm_force, c_force;  // set by final_rows() only (after the solve of the last substep)
  // (row forces and the per-foot contact wrench are NOT members: they are consumed right where a solver pass computes them and
  //  recomputed from the final Jaref for the sensors / efc_force output — nine registers less through the line search)
  float gauss, cost, prev_cost;
  int niter;
  float* trace;  // optional global [ZB_NTRACE] solver trace (diagnostics), nullptr in production
  int lscount;

  // ------------------------------------------------------------------------------------------
  ZB_DI void init_topology(const ZbDev* D_, const ZbOpts& o_, float meaninertia_, float* scr_, int lane, int cflags) {
    D = D_; o = o_; meaninertia = meaninertia_; scr = scr_; L = lane;
    any_box = GEN ? (cflags & 1) != 0 : true; any_prim = GEN && (cflags & 2) != 0; any_pair = GEN && (cflags & 4) != 0;
    isbody = L < NB;
    const int jt = isbody ? D->m.body_jnttype[L] : 0;
    hinge = jt == 2;
    bdof = L >= BL0;
    cs = D->cs[L]; ce = D->ce[L]; gs = D->gs[L]; gp = D->gp[L]; gn = D->gn[L];
    dof = D->lane_dof[L]; act = D->lane_act[L]; lbox = D->lane_box[L];
    hasdof = dof >= 0;
    chainbody = isbody && L >= 2;
    boxlane = D->m.box_body[L >> 4];
    pairlane = D->m.pair_body[L >> 4];
  }

  // chain-scan helpers (bodies of one chain are contiguous lanes)
  template <int N>
  ZB_DI void chain_prefix(float (&v)[N]) {  // inclusive, root -> leaf
#pragma unroll
    for (int d = 1; d <= 4; d <<= 1) {
      const bool take = chainbody && (L - d >= cs);
#pragma unroll
      for (int k = 0; k < N; k++) {
        const float u = __shfl_up_sync(ZB_FULL, v[k], d);
        if (take) v[k] += u;
      }
    }
  }
  template <int N>
  ZB_DI void chain_suffix(float (&v)[N]) {  // inclusive, leaf -> root
#pragma unroll
    for (int d = 1; d <= 4; d <<= 1) {
      const bool take = chainbody && (L + d <= ce);
#pragma unroll
      for (int k = 0; k < N; k++) {
        const float u = __shfl_down_sync(ZB_FULL, v[k], d);
        if (take) v[k] += u;
      }
    }
  }

  // cdof of the three rotational free-joint dofs (columns of the base rotation about the tree COM), uniform in all lanes
  ZB_DI void base_rot_cdofs(float (&cb)[3][6]) const {
    const float r[3] = {c0[0] - xb[0], c0[1] - xb[1], c0[2] - xb[2]};
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float col[3] = {Rb[i], Rb[3 + i], Rb[6 + i]};
      float c[3];
      cross3(c, col, r);
      cb[i][0] = col[0]; cb[i][1] = col[1]; cb[i][2] = col[2];
      cb[i][3] = c[0]; cb[i][4] = c[1]; cb[i][5] = c[2];
    }
  }

  // spatial velocity of the base body produced by generalized base velocity vb[6] (uniform compute)
  ZB_DI void base_motion(float* V, const float* vb) const {
    float om[3], r[3], c[3];
    mat_vec(om, Rb, vb + 3);
    r[0] = c0[0] - xb[0]; r[1] = c0[1] - xb[1]; r[2] = c0[2] - xb[2];
    cross3(c, om, r);
    V[0] = om[0]; V[1] = om[1]; V[2] = om[2];
    V[3] = vb[0] + c[0]; V[4] = vb[1] + c[1]; V[5] = vb[2] + c[2];
  }

  // ------------------------------------------------------------------------------------------
  // smooth.kinematics + smooth.com_pos
  // ------------------------------------------------------------------------------------------
  ZB_DI void kinematics() {
    const ZbModel& m = D->m;
    normalize4(bq + 3);  // MJX normalises the free-joint quaternion in qpos
    xb[0] = bq[0]; xb[1] = bq[1]; xb[2] = bq[2];
    quat_to_mat(Rb, bq + 3);
    float lp[3] = {0.f, 0.f, 0.f}, lq[4] = {1.f, 0.f, 0.f, 0.f};
    const int b = isbody ? L : 0;
    if (chainbody) {
      const float bqt[4] = {m.body_quat[b][0], m.body_quat[b][1], m.body_quat[b][2], m.body_quat[b][3]};
      lp[0] = D->bpj[b][0]; lp[1] = D->bpj[b][1]; lp[2] = D->bpj[b][2];
      if (hinge) {
        float s, c;
        sincosf((q - q0) * 0.5f, &s, &c);
        const float aa[4] = {c, m.jnt_axis[b][0] * s, m.jnt_axis[b][1] * s, m.jnt_axis[b][2] * s};
        quat_mul(lq, bqt, aa);
        float R[9], r[3];
        const float jp[3] = {m.jnt_pos[b][0], m.jnt_pos[b][1], m.jnt_pos[b][2]};
        quat_to_mat(R, lq);
        mat_vec(r, R, jp);
        lp[0] -= r[0]; lp[1] -= r[1]; lp[2] -= r[2];
      } else {
        lq[0] = bqt[0]; lq[1] = bqt[1]; lq[2] = bqt[2]; lq[3] = bqt[3];
      }
    }
    // compose local transforms down each chain (pointer doubling)
#pragma unroll
    for (int d = 1; d <= 4; d <<= 1) {
      float up[3], uq[4];
#pragma unroll
      for (int k = 0; k < 3; k++) up[k] = __shfl_up_sync(ZB_FULL, lp[k], d);
#pragma unroll
      for (int k = 0; k < 4; k++) uq[k] = __shfl_up_sync(ZB_FULL, lq[k], d);
      if (chainbody && (L - d >= cs)) {
        float R[9], r[3], nq[4];
        quat_to_mat(R, uq);
        mat_vec(r, R, lp);
        lp[0] = up[0] + r[0]; lp[1] = up[1] + r[1]; lp[2] = up[2] + r[2];
        quat_mul(nq, uq, lq);
        lq[0] = nq[0]; lq[1] = nq[1]; lq[2] = nq[2]; lq[3] = nq[3];
      }
    }
    if (chainbody) {
      float r[3];
      mat_vec(r, Rb, lp);
      xpos[0] = xb[0] + r[0]; xpos[1] = xb[1] + r[1]; xpos[2] = xb[2] + r[2];
      quat_mul(xquat, bq + 3, lq);
      normalize4(xquat);
    } else if (L == 1) {
      xpos[0] = xb[0]; xpos[1] = xb[1]; xpos[2] = xb[2];
      xquat[0] = bq[3]; xquat[1] = bq[4]; xquat[2] = bq[5]; xquat[3] = bq[6];
    } else {
      xpos[0] = xpos[1] = xpos[2] = 0.f;
      xquat[0] = 1.f; xquat[1] = xquat[2] = xquat[3] = 0.f;
    }
    quat_to_mat(xmat, xquat);
    {
      const float ip[3] = {m.body_ipos[b][0], m.body_ipos[b][1], m.body_ipos[b][2]};
      float r[3];
      mat_vec(r, xmat, ip);
      xipos[0] = xpos[0] + r[0]; xipos[1] = xpos[1] + r[1]; xipos[2] = xpos[2] + r[2];
    }
    // tree centre of mass (the only subtree_com the pipeline needs: body_rootid is the base)
    {
      float s[3] = {isbody ? mass * xipos[0] : 0.f, isbody ? mass * xipos[1] : 0.f, isbody ? mass * xipos[2] : 0.f};
#pragma unroll
      for (int k = 0; k < 3; k++) c0[k] = wsum(s[k]) * inv_mass_tot;
    }
    // cinert about c0
    {
      float iq[4] = {1.f, 0.f, 0.f, 0.f}, qq[4], R[9];
      if (isbody && L > 0) { iq[0] = m.body_iquat[b][0]; iq[1] = m.body_iquat[b][1]; iq[2] = m.body_iquat[b][2]; iq[3] = m.body_iquat[b][3]; }
      quat_mul(qq, xquat, iq);
      quat_to_mat(R, qq);
      const float off[3] = {xipos[0] - c0[0], xipos[1] - c0[1], xipos[2] - c0[2]};
      const float o2 = dot3(off, off);
      const float I0 = inertia[0], I1 = inertia[1], I2 = inertia[2];
      cinert[0] = R[0] * I0 * R[0] + R[1] * I1 * R[1] + R[2] * I2 * R[2] + mass * (o2 - off[0] * off[0]);
      cinert[1] = R[3] * I0 * R[3] + R[4] * I1 * R[4] + R[5] * I2 * R[5] + mass * (o2 - off[1] * off[1]);
      cinert[2] = R[6] * I0 * R[6] + R[7] * I1 * R[7] + R[8] * I2 * R[8] + mass * (o2 - off[2] * off[2]);
      cinert[3] = R[0] * I0 * R[3] + R[1] * I1 * R[4] + R[2] * I2 * R[5] - mass * off[0] * off[1];
      cinert[4] = R[0] * I0 * R[6] + R[1] * I1 * R[7] + R[2] * I2 * R[8] - mass * off[0] * off[2];
      cinert[5] = R[3] * I0 * R[6] + R[4] * I1 * R[7] + R[5] * I2 * R[8] - mass * off[1] * off[2];
      cinert[6] = off[0] * mass; cinert[7] = off[1] * mass; cinert[8] = off[2] * mass; cinert[9] = mass;
      if (!isbody) {
#pragma unroll
        for (int k = 0; k < 10; k++) cinert[k] = 0.f;
      }
    }
    // cdof.  The three rotational free-joint dofs are computed by every lane (uniform)
    {
      float cb3[3][6];
      base_rot_cdofs(cb3);
#pragma unroll
      for (int k = 0; k < 6; k++) cdof[k] = 0.f;
      if (hinge) {
        const float jp[3] = {m.jnt_pos[b][0], m.jnt_pos[b][1], m.jnt_pos[b][2]};
        const float ax[3] = {m.jnt_axis[b][0], m.jnt_axis[b][1], m.jnt_axis[b][2]};
        float a[3], an[3], c[3];
        mat_vec(a, xmat, jp);
        mat_vec(an, xmat, ax);
        const float o[3] = {c0[0] - (xpos[0] + a[0]), c0[1] - (xpos[1] + a[1]), c0[2] - (xpos[2] + a[2])};
        cross3(c, an, o);
        cdof[0] = an[0]; cdof[1] = an[1]; cdof[2] = an[2]; cdof[3] = c[0]; cdof[4] = c[1]; cdof[5] = c[2];
      } else if (bdof) {
        const int a = L - BL0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          cdof[3 + i] = (a == i) ? 1.f : cdof[3 + i];
#pragma unroll
          for (int k = 0; k < 6; k++) cdof[k] = (a == 3 + i) ? cb3[i][k] : cdof[k];
        }
      }
    }
    // site carried by this body (imu_site pose is per-env)
    if (isbody && D->body_site[L] >= 0) {
      float r[3], qq[4];
      mat_vec(r, xmat, spos);
      sxpos[0] = xpos[0] + r[0]; sxpos[1] = xpos[1] + r[1]; sxpos[2] = xpos[2] + r[2];
      quat_mul(qq, xquat, squat);
      quat_to_mat(sxmat, qq);
    }
  }

  // Rows of a tree-structured symmetric matrix  X[i][j] = cdof_j . f_i  (j an ancestor-or-self dof of i), in block form:
  // hinge lanes get their chain-block row (full, by symmetry) and their coupling to the six base dofs; base lanes get their
  // 6x6 row.  f = (composite inertia) x cdof gives the mass matrix, f = (sum_r D_r w_r w_r^T) x cdof the Newton J^T D J.
  ZB_DI void assemble_rows(const float (&f)[6], float (&rb)[6], float (&rc)[CHD]) {
#pragma unroll
    for (int k = 0; k < CHD; k++) {
      float ck[6];
#pragma unroll
      for (int j = 0; j < 6; j++) ck[j] = shf(cdof[j], gs + k);
      rc[k] = (hinge && k <= gp) ? dot6(ck, f) : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CHD; k++) scr[L * ZB_SCR + k] = rc[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CHD; k++) {
      const float t = scr[((gs + k) & 31) * ZB_SCR + gp];
      if (hinge && k > gp && k < gn) rc[k] = t;
    }
    __syncwarp();
    float cb3[3][6];
    base_rot_cdofs(cb3);
    rb[0] = f[3]; rb[1] = f[4]; rb[2] = f[5];
    rb[3] = dot6(cb3[0], f); rb[4] = dot6(cb3[1], f); rb[5] = dot6(cb3[2], f);
    if (!hasdof) {
#pragma unroll
      for (int k = 0; k < 6; k++) rb[k] = 0.f;
    }
  }

  // ------------------------------------------------------------------------------------------
  // smooth.crb + support.make_m in block form
  // ------------------------------------------------------------------------------------------
  ZB_DI void mass_matrix() {
    float crb[10];
#pragma unroll
    for (int k = 0; k < 10; k++) crb[k] = cinert[k];
    chain_suffix<10>(crb);
    {
      // composite inertia of the whole tree for the base dofs (world's cinert is zero)
      float tot[16];
#pragma unroll
      for (int k = 0; k < 10; k++) tot[k] = cinert[k];
#pragma unroll
      for (int k = 10; k < 16; k++) tot[k] = 0.f;
      reduce_scatter<16>(tot, L);  // lane L holds total of index L >> 1
#pragma unroll
      for (int k = 0; k < 10; k++) {
        const float t = shf(tot[0], 2 * k);
        if (bdof) crb[k] = t;
      }
    }
    float f[6];
    inert_mul(f, crb, cdof);
    float Bm[6], C[CHD];
    assemble_rows(f, Bm, C);
    // armature on the diagonal
#pragma unroll
    for (int k = 0; k < CHD; k++) {
      // (select form on purpose: `if (k == gp) C[k] += ..` gets turned into a dynamic index -> local memory)
      fput(F_C + k, C[k] + ((hinge && k == gp) ? arm : 0.f));
      fput(F_BM + k, Bm[k] + ((bdof && k == gp) ? arm : 0.f));
    }
  }

  // In-group right-looking LDL^T on the rows held in r[] (lower part), lanes gs..gs+gn-1:  A = Lt D Lt^T, Lt unit lower.
  // Afterwards r[k] (k < gp) = Lt[gp][k], invd = 1 / D[gp].  No square roots; entries above the diagonal are scratch.
  ZB_DI void group_ldl(float (&r)[CHD], float& invd, bool valid) {
#pragma unroll
    for (int k = 0; k < CHD; k++) {
      float d = shf(r[k], gs + k);
      if (!(valid && k < gn)) d = 1.f;
      const float id = __frcp_rn(d);
      float ajk[CHD];
#pragma unroll
      for (int j = k + 1; j < CHD; j++) ajk[j] = shf(r[k], gs + j);  // a_jk BEFORE scaling = Lt_jk d_k
      r[k] = (valid && gp > k) ? r[k] * id : r[k];
      invd = (valid && gp == k) ? id : invd;
#pragma unroll
      for (int j = k + 1; j < CHD; j++) r[j] = fmaf(-r[k], ajk[j], r[j]);  // rows j > gp only collect scratch
    }
  }

  // smooth.factor_m as a block LDL^T of (M + diag(extra)), M = [[A, B^T], [B, C]]:
  //   chains  C = Lt_c D_c Lt_c^T,  Wt = Lt_c^-1 B,  G = D_c^-1 Wt,  S = A - Wt^T G = Lt_s D_s Lt_s^T.
  // All stored coefficients are ZERO where they do not apply (other lanes, k >= gp, short chains), so the substitutions in
  // solve() are bare SHFL + FFMA steps without per-step predicates.
  ZB_DI void factor(const float (&Bsrc)[6], const float (&Csrc)[CHD], float extra) {
    Factor F;
    float r[CHD];
#pragma unroll
    for (int k = 0; k < CHD; k++) r[k] = hinge ? Csrc[k] + (k == gp ? extra : 0.f) : 0.f;
    float invd = 0.f;
    group_ldl(r, invd, hinge);
#pragma unroll
    for (int k = 0; k < CHD; k++) F.lc[k] = (hinge && k < gp) ? r[k] : 0.f;
    // Wt = Lt_c^-1 B by unit forward substitution, six right-hand sides
    float W[6];
#pragma unroll
    for (int a = 0; a < 6; a++) W[a] = hinge ? Bsrc[a] : 0.f;
#pragma unroll
    for (int k = 0; k < CHD - 1; k++) {
#pragma unroll
      for (int a = 0; a < 6; a++) {
        const float wk = shf(W[a], gs + k);
        W[a] = fmaf(-F.lc[k], wk, W[a]);
      }
    }
#pragma unroll
    for (int a = 0; a < 6; a++) F.W[a] = W[a] * invd;  // G = D_c^-1 Wt (zero outside the hinge lanes)
    // Schur complement S = A - Wt^T G: 21 unique sums over the hinge lanes, one result per lane
    float rb[CHD];
    {
      float p[32];
      int t = 0;
#pragma unroll
      for (int a = 0; a < 6; a++)
#pragma unroll
        for (int b = a; b < 6; b++) p[t++] = W[a] * F.W[b];
#pragma unroll
      for (int k = 21; k < 32; k++) p[k] = 0.f;
      reduce_scatter<32>(p, L);  // lane t holds sum_t
      const int a = bdof ? L - BL0 : 0;
#pragma unroll
      for (int b = 0; b < 6; b++) {
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const int idx = lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);
        const float sv = shf(p[0], idx);
        rb[b] = bdof ? Bsrc[b] - sv + (b == gp ? extra : 0.f) : 0.f;
      }
    }
    float invb = 0.f;
    group_ldl(rb, invb, bdof);
#pragma unroll
    for (int k = 0; k < CHD; k++) F.lc[k] = (bdof && k < gp) ? rb[k] : F.lc[k];
    F.inv = hinge ? invd : (bdof ? invb : 0.f);
    // column copies Lt[k][gp] (k > gp) for the back substitutions
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CHD; k++) scr[L * ZB_SCR + k] = F.lc[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CHD; k++) {
      const float t = scr[((gs + k) & 31) * ZB_SCR + gp];
      F.lt[k] = (hasdof && k > gp && k < gn) ? t : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < CHD; k++) { fput(F_LC + k, F.lc[k]); fput(F_LT + k, F.lt[k]); fput(F_W + k, F.W[k]); }
    fput(F_INV, F.inv);
  }
  // factor of M + diag(extra) from the stored mass-matrix rows
  ZB_DI void factor_m(float extra) {
    float Bm[6], C[CHD];
    load_m(Bm, C);
    factor(Bm, C, extra);
  }

  // smooth.solve_m: x <- (M + diag)^-1 x, x distributed over the dof lanes
  ZB_DI float solve(float x) {
    Factor F;  // solve_all re-factors F0 (Newton Hessian, M + dt D) when another matrix is to be solved
    load_factor(F);
    if (!hasdof) x = 0.f;
    // chains: unit forward substitution  y_c = Lt_c^-1 r_c
#pragma unroll
    for (int k = 0; k < CHD - 1; k++) {
      const float yk = shf(x, gs + k);
      x = pfnma(hinge, F.lc[k], yk, x);
    }
    // base right-hand side: r_b - sum_i G_i y_i
    {
      float p[8];
#pragma unroll
      for (int a = 0; a < 6; a++) p[a] = F.W[a] * x;  // G is zero outside the hinge lanes
      p[6] = p[7] = 0.f;
      reduce_scatter<8>(p, L);  // lanes 4a..4a+3 hold total a
      const float t = shf(p[0], bdof ? 4 * (L - BL0) : 0);
      if (bdof) x -= t;
    }
    // base group: unit forward, scale by D^-1 (chain lanes too: D_c^-1 y_c), unit backward; the backward pulls also
    // broadcast x_base to every lane
#pragma unroll
    for (int k = 0; k < 5; k++) {
      const float yk = shf(x, BL0 + k);
      x = pfnma(bdof, F.lc[k], yk, x);
    }
    x *= F.inv;
    float xbv[6];
#pragma unroll
    for (int k = 5; k >= 0; k--) {
      xbv[k] = shf(x, BL0 + k);
      x = pfnma(bdof, F.lt[k], xbv[k], x);
    }
    // chains: Lt_c^T x_c = D_c^-1 y_c - G x_b, unit backward substitution
#pragma unroll
    for (int a = 0; a < 6; a++) x = fmaf(-F.W[a], xbv[a], x);
#pragma unroll
    for (int k = CHD - 1; k >= 1; k--) {
      const float xk = shf(x, gs + k);
      x = pfnma(hinge, F.lt[k], xk, x);
    }
    return x;
  }

  // support.mul_m: y = M v; also returns the broadcast base part of v
  ZB_DI float mul_m(float v, float (&vb)[6]) {
    if (!hasdof) v = 0.f;
    float Bm[6], C[CHD];
    load_m(Bm, C);
#pragma unroll
    for (int a = 0; a < 6; a++) vb[a] = shf(v, BL0 + a);
    float y = 0.f;
#pragma unroll
    for (int a = 0; a < 6; a++) y += Bm[a] * vb[a];
#pragma unroll
    for (int k = 0; k < CHD; k++) {
      const float vk = shf(v, gs + k);
      y = fmaf(C[k], vk, y);  // C is zero outside the lane's chain block
    }
    float p[8];
#if ZB_PRED
    const float vh = hinge ? v : 0.f;  // (base lanes hold A in Bm: must stay masked)
#pragma unroll
    for (int a = 0; a < 6; a++) p[a] = Bm[a] * vh;
#else
#pragma unroll
    for (int a = 0; a < 6; a++) p[a] = hinge ? Bm[a] * v : 0.f;  // (base lanes hold A in Bm: must stay masked)
#endif
    p[6] = p[7] = 0.f;
    reduce_scatter<8>(p, L);
    const float t = shf(p[0], bdof ? 4 * (L - BL0) : 0);
    if (bdof) y += t;
    return hasdof ? y : 0.f;
  }

  // contact rows: jv[r] = sum_d J[r, d] v_d over the chain dofs of the box and the base dofs (vb = broadcast base part of v)
  ZB_DI float contact_jv(float v, const float (&vb)[6]) {
#if ZB_EXPLICIT_J
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < CHD; k++) s = fmaf(cjget(k), shf(v, jsrc + k), s);
    if (any_pair) {  // (kernel-uniform) some group has geom 1 on a body: its second block are that chain's dofs
#pragma unroll
      for (int k = 0; k < CHD; k++) s = fmaf(cjget(6 + k), shf(v, jsrc2 + k), s);
    } else {
#pragma unroll
      for (int a = 0; a < 6; a++) s = fmaf(cjget(6 + a), vb[a], s);
    }
    return s;
#else
    float pre[6];
#pragma unroll
    for (int k = 0; k < 6; k++) pre[k] = hinge ? cdof[k] * v : 0.f;
    chain_prefix<6>(pre);
    float Vb[6], Vf[6];
    base_motion(Vb, vb);
#pragma unroll
    for (int k = 0; k < 6; k++) Vf[k] = shf(pre[k], boxlane) + Vb[k];
    return dot6(w, Vf);
#endif
  }

  // J[r, :] of this lane's contact row (once per substep, after make_constraint): for a hinge with axis a through the anchor p,
  // J = dir . (a x (pos - p)) — the lever arm taken from the joint anchor itself, not as the difference of two arms about the tree
  // COM (w . cdof = dir . (a x (pos - c0)) + dir . (a x (c0 - p)) loses three bits to cancellation at the ankle, 3 cm from the
  // contact and 25 cm from the COM).  Base dofs: translations J = dir, rotations (body-frame axes through the base origin) likewise.
  ZB_DI void contact_rows_jacobian() {
    jsrc = D->cs[boxlane];
    float anc[3] = {0.f, 0.f, 0.f};
    if (hinge) {
      const float jp[3] = {D->m.jnt_pos[L][0], D->m.jnt_pos[L][1], D->m.jnt_pos[L][2]};
      float a[3];
      mat_vec(a, xmat, jp);
      anc[0] = xpos[0] + a[0]; anc[1] = xpos[1] + a[1]; anc[2] = xpos[2] + a[2];
    }
    const float dir[3] = {w[3], w[4], w[5]};
    // chain of geom 2: the hinge of lane jsrc + k moves the geom iff it lies on the path base -> body of the geom
#pragma unroll
    for (int k = 0; k < CHD; k++) {
      float ax[3], rel[3], c[3];
#pragma unroll
      for (int j = 0; j < 3; j++) { ax[j] = shf(cdof[j], jsrc + k); rel[j] = cpos[j] - shf(anc[j], jsrc + k); }
      const int hk = __shfl_sync(ZB_FULL, (int)hinge, jsrc + k);
      cross3(c, ax, rel);
      cjput(k, (hk && jsrc + k <= boxlane) ? dot3(dir, c) : 0.f);
    }
    // second block: base dofs (geom 1 = the floor plane: translations J = dir, rotations about the body-frame axes through the base origin)
    const float rel[3] = {cpos[0] - xb[0], cpos[1] - xb[1], cpos[2] - xb[2]};
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float col[3] = {Rb[a], Rb[3 + a], Rb[6 + a]};
      float c[3];
      cross3(c, col, rel);
      cjput(6 + a, dir[a]);
      cjput(9 + a, dot3(dir, c));
    }
    jsrc2 = BL0;
    if (any_pair) {
      // ... or the chain of geom 1 with the opposite sign (support.jac_dif_pair; the base dofs move both geoms alike and cancel).
      // The shuffles are executed by every lane; lanes of floor groups keep their base block.
      const bool pair = pairlane != 0;
      const int j2 = pair ? D->cs[pairlane] : BL0;
#pragma unroll
      for (int k = 0; k < CHD; k++) {
        float ax[3], rl[3], c[3];
#pragma unroll
        for (int j = 0; j < 3; j++) { ax[j] = shf(cdof[j], (j2 + k) & 31); rl[j] = cpos[j] - shf(anc[j], (j2 + k) & 31); }
        const int hk = __shfl_sync(ZB_FULL, (int)hinge, (j2 + k) & 31);
        cross3(c, ax, rl);
        const float v = (hk && j2 + k <= pairlane) ? -dot3(dir, c) : 0.f;
        if (pair) {
          if (j2 == jsrc) { cjput(k, cjget(k) + v); cjput(6 + k, 0.f); }  // both geoms hang on the same chain
          else cjput(6 + k, v);
        }
      }
      jsrc2 = j2;
    }
  }

  // Total contact wrench of the foot box of my half-warp, sum_r f_r w_r over its 16 row lanes, as a transpose-reduce inside the
  // half: 8 shuffles instead of 24.  Afterwards component k of box g sits in lanes 16 g + 2k (+1).  Every component is still
  // summed over the lanes in the butterfly order 8, 4, 2, 1, i.e. bit for bit what six separate butterflies give.
  ZB_DI float contact_wrench(float f) const {
    float v[8];
#pragma unroll
    for (int k = 0; k < 6; k++) v[k] = c_on ? f * w[k] : 0.f;
    v[6] = v[7] = 0.f;
    int m = 8;
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1, m >>= 1) {
      const bool up = (L & m) != 0;
#pragma unroll
      for (int j = 0; j < half; j++) {
        const float send = up ? v[j] : v[j + half], keep = up ? v[j + half] : v[j];
        v[j] = keep + __shfl_xor_sync(ZB_FULL, send, m);
      }
    }
    return v[0] + shx(v[0], 1);
  }

  // J^T f for the contact rows
  ZB_DI float contact_jtf(float f) {
#if ZB_EXPLICIT_J
    // sum_r J[r, d] f_r over the 16 row lanes of each box as a transpose-reduce inside the half-warp: total t lands in lane 16 g + t
    float p[16];
#if ZB_PRED
    const float fm = c_on ? f : 0.f;
#pragma unroll
    for (int k = 0; k < 12; k++) p[k] = fm * cjget(k);
#else
#pragma unroll
    for (int k = 0; k < 12; k++) p[k] = c_on ? f * cjget(k) : 0.f;
#endif
    p[12] = p[13] = p[14] = p[15] = 0.f;
    int m = 8;
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1, m >>= 1) {
      const bool up = (L & m) != 0;
#pragma unroll
      for (int j = 0; j < half; j++) {
        const float send = up ? p[j] : p[j + half], keep = up ? p[j + half] : p[j];
        p[j] = keep + __shfl_xor_sync(ZB_FULL, send, m);
      }
    }
    if (any_pair) {  // (kernel-uniform) general candidate list: every dof lane picks its entry of each group's totals (ZbDev.lane_jt)
      float r = 0.f;
#pragma unroll
      for (int g = 0; g < ZB_NBOX; g++) {
        const int idx = D->lane_jt[L][g];
        const float t = shf(p[0], 16 * g + (idx < 0 ? 0 : idx));
        r += idx < 0 ? 0.f : t;
      }
      return r;
    }
    // floor groups only: chain dof at position gp of box lbox reads lane 16 lbox + gp; base dof a reads the sum over both boxes from lane 6 + a
    const float both = p[0] + shx(p[0], 16);  // (executed by all lanes)
    const float pub = ((L & 15) >= 6) ? both : p[0];
    const bool leg = hinge && lbox >= 0;
    const float r = shf(pub, leg ? 16 * lbox + gp : (bdof ? 6 + (L - BL0) : 0));
    return (leg || bdof) ? r : 0.f;
#else
    const float tot = contact_wrench(f);
    // even lanes keep the wrench of their own box, odd lanes the sum over both boxes: a leg dof pulls component k from lane
    // 16 box + 2k, a base dof from lane 2k + 1 — one shuffle per component for either kind of lane
    const float both_t = tot + shx(tot, 16);  // (the shuffle must be executed by ALL lanes: not inside the select)
    const float pub = (L & 1) ? both_t : tot;
    const bool leg = hinge && lbox >= 0;
    const int src = leg ? 16 * lbox : 1;
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const float wk = shf(pub, src + 2 * k);
      if (leg || bdof) r += cdof[k] * wk;
    }
    return r;
#endif
  }

  // ------------------------------------------------------------------------------------------
  // smooth.com_vel (A.4), passive, smooth.rne (A.5)
  // ------------------------------------------------------------------------------------------
  float cvel_base[6], cacc_base_lin[3];
  ZB_DI void velocity(float& qfrc_passive) {
    float vb[6];
#pragma unroll
    for (int a = 0; a < 6; a++) vb[a] = shf(qv, BL0 + a);
    base_motion(cvel_base, vb);
    float pre[6];
#pragma unroll
    for (int k = 0; k < 6; k++) pre[k] = hinge ? cdof[k] * qv : 0.f;
    chain_prefix<6>(pre);
#pragma unroll
    for (int k = 0; k < 6; k++) cvel[k] = (isbody && L >= 1) ? cvel_base[k] + (chainbody ? pre[k] : 0.f) : 0.f;
    // cdof_dot: hinge = cvel(parent) x cdof == cvel(body) x cdof; free rotations = [0; v_lin] x cdof
#pragma unroll
    for (int k = 0; k < 6; k++) cdofd[k] = 0.f;
    if (hinge) motion_cross(cdofd, cvel, cdof);
    if (bdof && L - BL0 >= 3) {
      const float u[6] = {0.f, 0.f, 0.f, vb[0], vb[1], vb[2]};
      motion_cross(cdofd, u, cdof);
    }
    // base part of cacc (gravity + free-joint cdof_dot * qvel): [0; -g + v_lin x omega_world]
    {
      const float vl[3] = {vb[0], vb[1], vb[2]};
      float c[3];
      cross3(c, vl, cvel_base);
      cacc_base_lin[0] = -o.gravity[0] + c[0];
      cacc_base_lin[1] = -o.gravity[1] + c[1];
      cacc_base_lin[2] = -o.gravity[2] + c[2];
    }
    qfrc_passive = hasdof ? -damp * qv : 0.f;
  }

  ZB_DI void bias_forces() {
    float pre[6], cacc[6];
#pragma unroll
    for (int k = 0; k < 6; k++) pre[k] = hinge ? cdofd[k] * qv : 0.f;
    chain_prefix<6>(pre);
#pragma unroll
    for (int k = 0; k < 6; k++) cacc[k] = chainbody ? pre[k] : 0.f;
    if (L == 0) { cacc[3] = -o.gravity[0]; cacc[4] = -o.gravity[1]; cacc[5] = -o.gravity[2]; }
    else if (isbody) { cacc[3] += cacc_base_lin[0]; cacc[4] += cacc_base_lin[1]; cacc[5] += cacc_base_lin[2]; }
    float cfrc[6], t[6], u[6];
    inert_mul(cfrc, cinert, cacc);
    inert_mul(t, cinert, cvel);
    motion_cross_force(u, cvel, t);
#pragma unroll
    for (int k = 0; k < 6; k++) cfrc[k] = isbody ? cfrc[k] + u[k] : 0.f;
    float tot[8];
#pragma unroll
    for (int k = 0; k < 6; k++) tot[k] = (isbody && L >= 1) ? cfrc[k] : 0.f;
    tot[6] = tot[7] = 0.f;
    chain_suffix<6>(cfrc);
    reduce_scatter<8>(tot, L);
    float r = hinge ? dot6(cdof, cfrc) : 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const float tk = shf(tot[0], 4 * k);
      if (bdof) r += cdof[k] * tk;
    }
    qfrc_bias = r;
  }

  // ------------------------------------------------------------------------------------------
  // FeetechActuators.get_stateful_ctrl + trapezoidal_step (train.py:1137-1196, 1242-1280)
  // ------------------------------------------------------------------------------------------
  ZB_DI void actuator(float dt) {
    if (act < 0) return;
    const ZbModel& m = D->m;
    const float vmax = m.servo_vmax[act], amax = m.servo_amax[act];
    const float err = tgt - apos;
    const float thr = err >= 0.f ? m.servo_deadband_pos : m.servo_deadband_neg;
    const bool in_db = fabsf(err) <= thr;
    const float db_v = avel * 0.8f;
    const float db_p = apos + db_v * dt;
    const float tdir = signf(err);
    const float stop = fabsf(avel * avel) / (2.f * amax);
    const float vdir = signf(avel);
    const bool towards = vdir * tdir >= 0.f;
    const bool accel = towards && fabsf(err) > stop;
    float acc = accel ? tdir * amax : -vdir * amax;
    if (fabsf(avel) < 1e-6f) acc = tdir * amax;
    float pv = avel + acc * dt;
    pv = clampf(pv, -vmax, vmax);
    const float pp = apos + pv * dt;
    avel = in_db ? db_v : pv;
    apos = in_db ? db_p : pp;
    const float perr = apos - q, verr = avel - qv;
    float duty = m.servo_kp[act] * m.servo_error_gain[act] * perr + m.servo_kd[act] * verr;
    duty = clampf(duty, -m.servo_max_pwm[act], m.servo_max_pwm[act]);
    const float volts = duty * m.servo_vin[act];
    atq = volts * m.servo_kt[act] / m.servo_R[act];
    ctrl = atq;
  }

  // forward.fwd_actuation for motors (A.6)
  ZB_DI float actuation() {
    act_force = 0.f;
    if (act < 0) return 0.f;
    const ZbModel& m = D->m;
    float c = ctrl;
    if (m.act_ctrllimited[act]) c = clampf(c, m.act_ctrlrange[act][0], m.act_ctrlrange[act][1]);
    float f = c;
    if (m.act_forcelimited[act]) f = clampf(f, m.act_forcerange[act][0], m.act_forcerange[act][1]);
    act_force = f;
    return m.act_gear[act] * f;
  }

  // ------------------------------------------------------------------------------------------
  // collision: plane_convex for the foot boxes + _manifold_points (A.7).  Every lane evaluates
  // vertex (L & 7) of box (L >> 4); 8-lane groups pick the manifold; lane r then becomes row r.
  // ------------------------------------------------------------------------------------------
  ZB_DI int gargmax(float val, int idx) {
#pragma unroll
    for (int d = 1; d <= 4; d <<= 1) {
      const float ov = shx(val, d);
      const int oi = __shfl_xor_sync(ZB_FULL, idx, d);
      if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
    }
    return idx;
  }
  ZB_DI void collision() {
    const ZbModel& m = D->m;
    const int g = L >> 4;
    // world pose of the geoms: the body lanes compute theirs, the row lanes fetch the two of their group
    float gpos[3] = {0.f, 0.f, 0.f}, gmat[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (isbody && D->body_box[L] >= 0) {
      const int gb = D->body_box[L];
      const float bp[3] = {m.box_pos[gb][0], m.box_pos[gb][1], m.box_pos[gb][2]};
      const float bqq[4] = {m.box_quat[gb][0], m.box_quat[gb][1], m.box_quat[gb][2], m.box_quat[gb][3]};
      float r[3], qq[4];
      mat_vec(r, xmat, bp);
      gpos[0] = xpos[0] + r[0]; gpos[1] = xpos[1] + r[1]; gpos[2] = xpos[2] + r[2];
      quat_mul(qq, xquat, bqq);
      quat_to_mat(gmat, qq);
    }
    float bp[3], bm[9];
#pragma unroll
    for (int k = 0; k < 3; k++) bp[k] = shf(gpos[k], boxlane);
#pragma unroll
    for (int k = 0; k < 9; k++) bm[k] = shf(gmat[k], boxlane);
    const float pn[3] = {m.floor_normal[0], m.floor_normal[1], m.floor_normal[2]};
    // defaults: inert slot, floor frame (D->frame = make_frame(floor normal): rows n, t1, t2)
    const int e = L & 3, s = (L >> 2) & 3;
    if (GEN) {
      const float* fr = D->frame;
      cdist = 1.f; cpos[0] = cpos[1] = cpos[2] = 0.f;
      cn[0] = fr[0]; cn[1] = fr[1]; cn[2] = fr[2];
      ct[0] = fr[3 * (1 + (e >> 1))]; ct[1] = fr[3 * (1 + (e >> 1)) + 1]; ct[2] = fr[3 * (1 + (e >> 1)) + 2];
    }
    if (any_box) collide_plane_box(bp, bm);
    if (GEN && (any_prim || any_pair)) {
      // sphere / capsule primitives (mjx collision_primitive [R]): plain per-lane arithmetic, every row lane evaluates its own slot
      const int kind = D->grp_kind[g];
      const float r2 = m.box_size[g][0], h2 = m.box_size[g][1];
      const float ax2[3] = {bm[2], bm[5], bm[8]};
      const float ppl[3] = {m.floor_pos[0], m.floor_pos[1], m.floor_pos[2]};
      if (kind == ZB_PAIR_PLANE_SPHERE || kind == ZB_PAIR_PLANE_CAPSULE) {
        // plane_sphere / plane_capsule: capsule = two spheres at the segment ends (slots 0, 1), frame aligned with the capsule axis
        const float sg = (kind == ZB_PAIR_PLANE_CAPSULE) ? (s == 0 ? 1.f : -1.f) * h2 : 0.f;
        const float c[3] = {bp[0] + sg * ax2[0], bp[1] + sg * ax2[1], bp[2] + sg * ax2[2]};
        const float rel[3] = {c[0] - ppl[0], c[1] - ppl[1], c[2] - ppl[2]};
        const float dist = dot3(rel, pn) - r2;
        const bool live = s < (kind == ZB_PAIR_PLANE_CAPSULE ? 2 : 1);
        cdist = live ? dist : 1.f;
        const float back = r2 + 0.5f * dist;
        cpos[0] = c[0] - pn[0] * back; cpos[1] = c[1] - pn[1] * back; cpos[2] = c[2] - pn[2] * back;
        if (kind == ZB_PAIR_PLANE_CAPSULE) {
          const float na = dot3(pn, ax2);
          float b[3] = {ax2[0] - pn[0] * na, ax2[1] - pn[1] * na, ax2[2] - pn[2] * na};
          const float bn = sqrtf(dot3(b, b));
          const float sc = 1.f / (bn + 1e-6f * (bn == 0.f));
          b[0] *= sc; b[1] *= sc; b[2] *= sc;
          if (bn < 0.5f) { const bool yy = -0.5f < pn[1] && pn[1] < 0.5f; b[0] = 0.f; b[1] = yy ? 1.f : 0.f; b[2] = yy ? 0.f : 1.f; }
          float t2[3];
          cross3(t2, pn, b);
          cn[0] = pn[0]; cn[1] = pn[1]; cn[2] = pn[2];
#pragma unroll
          for (int k = 0; k < 3; k++) ct[k] = (e >> 1) ? t2[k] : b[k];
        }
      }
      if (any_pair) {
        // geom 1 on a body (self-collision): sphere_sphere / sphere_capsule / capsule_capsule, one contact, normal from geom 1 to geom 2
        float qpos_[3] = {0.f, 0.f, 0.f}, qax[3] = {0.f, 0.f, 1.f};
        if (isbody && D->body_pair[L] >= 0) {
          const int gb = D->body_pair[L];
          const float pp_[3] = {m.pair_pos[gb][0], m.pair_pos[gb][1], m.pair_pos[gb][2]};
          const float pq_[4] = {m.pair_quat[gb][0], m.pair_quat[gb][1], m.pair_quat[gb][2], m.pair_quat[gb][3]};
          float r[3], qq[4], R[9];
          mat_vec(r, xmat, pp_);
          qpos_[0] = xpos[0] + r[0]; qpos_[1] = xpos[1] + r[1]; qpos_[2] = xpos[2] + r[2];
          quat_mul(qq, xquat, pq_);
          quat_to_mat(R, qq);
          qax[0] = R[2]; qax[1] = R[5]; qax[2] = R[8];
        }
        float p1[3], ax1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { p1[k] = shf(qpos_[k], pairlane); ax1[k] = shf(qax[k], pairlane); }
        if (kind >= ZB_PAIR_SPHERE_SPHERE) {
          const float r1 = m.pair_size[g][0], h1 = m.pair_size[g][1];
          float q1[3] = {p1[0], p1[1], p1[2]}, q2[3] = {bp[0], bp[1], bp[2]};
          if (kind == ZB_PAIR_SPHERE_CAPSULE) {
            const float a[3] = {bp[0] - ax2[0] * h2, bp[1] - ax2[1] * h2, bp[2] - ax2[2] * h2};
            const float b[3] = {bp[0] + ax2[0] * h2, bp[1] + ax2[1] * h2, bp[2] + ax2[2] * h2};
            closest_segment_point(q2, a, b, p1);
          } else if (kind == ZB_PAIR_CAPSULE_CAPSULE) {
            const float a0[3] = {p1[0] - ax1[0] * h1, p1[1] - ax1[1] * h1, p1[2] - ax1[2] * h1};
            const float a1[3] = {p1[0] + ax1[0] * h1, p1[1] + ax1[1] * h1, p1[2] + ax1[2] * h1};
            const float b0[3] = {bp[0] - ax2[0] * h2, bp[1] - ax2[1] * h2, bp[2] - ax2[2] * h2};
            const float b1[3] = {bp[0] + ax2[0] * h2, bp[1] + ax2[1] * h2, bp[2] + ax2[2] * h2};
            closest_segment_to_segment(q1, q2, a0, a1, b0, b1);
          }
          // _sphere_sphere
          const float dl[3] = {q2[0] - q1[0], q2[1] - q1[1], q2[2] - q1[2]};
          const float len = sqrtf(dot3(dl, dl));
          const float sc = 1.f / (len + 1e-6f * (len == 0.f));
          float n[3] = {dl[0] * sc, dl[1] * sc, dl[2] * sc};
          if (len == 0.f) { n[0] = 1.f; n[1] = 0.f; n[2] = 0.f; }
          const float dist = len - (r1 + r2);
          cdist = s == 0 ? dist : 1.f;
          const float fwd = r1 + 0.5f * dist;
          cpos[0] = q1[0] + n[0] * fwd; cpos[1] = q1[1] + n[1] * fwd; cpos[2] = q1[2] + n[2] * fwd;
          float t1[3], t2[3];
          make_frame_tangents(n, t1, t2);
          cn[0] = n[0]; cn[1] = n[1]; cn[2] = n[2];
#pragma unroll
          for (int k = 0; k < 3; k++) ct[k] = (e >> 1) ? t2[k] : t1[k];
        }
      }
    }
  }

  // math.make_frame(n) [R]: n normalised, b = y (or z when n is close to y), Gram-Schmidt, t2 = n x b
  static ZB_DI void make_frame_tangents(float (&a)[3], float (&b)[3], float (&c)[3]) {
    float nn = sqrtf(dot3(a, a));
    float sc = 1.f / (nn + 1e-6f * (nn == 0.f));
    a[0] *= sc; a[1] *= sc; a[2] *= sc;
    const bool yy = -0.5f < a[1] && a[1] < 0.5f;
    b[0] = 0.f; b[1] = yy ? 1.f : 0.f; b[2] = yy ? 0.f : 1.f;
    const float ab = dot3(a, b);
    b[0] -= a[0] * ab; b[1] -= a[1] * ab; b[2] -= a[2] * ab;
    nn = sqrtf(dot3(b, b));
    sc = 1.f / (nn + 1e-6f * (nn == 0.f));
    b[0] *= sc; b[1] *= sc; b[2] *= sc;
    cross3(c, a, b);
  }
  // math.closest_segment_point [R]
  static ZB_DI void closest_segment_point(float (&out)[3], const float (&a)[3], const float (&b)[3], const float (&pt)[3]) {
    const float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ap[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
    float t = dot3(ap, ab) / (dot3(ab, ab) + 1e-6f);
    t = fminf(fmaxf(t, 0.f), 1.f);
    out[0] = a[0] + t * ab[0]; out[1] = a[1] + t * ab[1]; out[2] = a[2] + t * ab[2];
  }
  // math.closest_segment_to_segment_points [R]
  static ZB_DI void closest_segment_to_segment(float (&best_a)[3], float (&best_b)[3], const float (&a0)[3], const float (&a1)[3],
                                               const float (&b0)[3], const float (&b1)[3]) {
    float da[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, db[3] = {b1[0] - b0[0], b1[1] - b0[1], b1[2] - b0[2]};
    const float la = sqrtf(dot3(da, da)), lb = sqrtf(dot3(db, db));
    const float sa = 1.f / (la + 1e-6f * (la == 0.f)), sb = 1.f / (lb + 1e-6f * (lb == 0.f));
    da[0] *= sa; da[1] *= sa; da[2] *= sa; db[0] *= sb; db[1] *= sb; db[2] *= sb;
    const float ha = la * 0.5f, hb = lb * 0.5f;
    float am[3], bmid[3], tr[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { am[k] = a0[k] + da[k] * ha; bmid[k] = b0[k] + db[k] * hb; tr[k] = am[k] - bmid[k]; }
    const float dd = dot3(da, db), dat = dot3(da, tr), dbt = dot3(db, tr);
    const float denom = 1.f - dd * dd;
    float ta = (-dat + dd * dbt) / (denom + 1e-6f);
    float tb = dbt + ta * dd;
    ta = fminf(fmaxf(ta, -ha), ha);
    tb = fminf(fmaxf(tb, -hb), hb);
    float ba[3], bb[3], na[3], nb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { ba[k] = am[k] + da[k] * ta; bb[k] = bmid[k] + db[k] * tb; }
    closest_segment_point(na, a0, a1, bb);
    closest_segment_point(nb, b0, b1, ba);
    const float e1[3] = {na[0] - bb[0], na[1] - bb[1], na[2] - bb[2]}, e2[3] = {ba[0] - nb[0], ba[1] - nb[1], ba[2] - nb[2]};
    const bool first = dot3(e1, e1) < dot3(e2, e2);
#pragma unroll
    for (int k = 0; k < 3; k++) { best_a[k] = first ? na[k] : ba[k]; best_b[k] = first ? bb[k] : nb[k]; }
  }

  // collision_convex.plane_convex + _manifold_points for the box groups: every lane evaluates vertex (L & 7) of the box of its
  // half-warp; 8-lane groups pick the manifold; lane r then becomes row r.  (Shuffles inside: executed by ALL lanes; lanes of
  // other primitives discard the result.)
  ZB_DI void collide_plane_box(const float (&bp)[3], const float (&bm)[9]) {
    const ZbModel& m = D->m;
    const int g = L >> 4;
    const float pn[3] = {m.floor_normal[0], m.floor_normal[1], m.floor_normal[2]};
    const float rel[3] = {m.floor_pos[0] - bp[0], m.floor_pos[1] - bp[1], m.floor_pos[2] - bp[2]};
    float ppl[3], nl[3];
    matT_vec(ppl, bm, rel);
    matT_vec(nl, bm, pn);
    const float sz[3] = {m.box_size[g][0], m.box_size[g][1], m.box_size[g][2]};
#define ZB_VERT(o, i) { o[0] = ((i) & 4) ? sz[0] : -sz[0]; o[1] = ((i) & 2) ? sz[1] : -sz[1]; o[2] = ((i) & 1) ? sz[2] : -sz[2]; }
#define ZB_SUPPORT(vv) ((ppl[0] - vv[0]) * nl[0] + (ppl[1] - vv[1]) * nl[1] + (ppl[2] - vv[2]) * nl[2])
    const int v = L & 7;
    float pv[3];
    ZB_VERT(pv, v);
    const float sup = ZB_SUPPORT(pv);
    float smax = sup;
    smax = fmaxf(smax, shx(smax, 1)); smax = fmaxf(smax, shx(smax, 2)); smax = fmaxf(smax, shx(smax, 4));
    const float thr = fmaxf(0.f, smax - 1e-3f);
    const float dm = sup > thr ? 0.f : -1e6f;
    const int ia = gargmax(dm, v);
    float a[3]; ZB_VERT(a, ia);
    const float ap[3] = {a[0] - pv[0], a[1] - pv[1], a[2] - pv[2]};
    const int ib = gargmax(dot3(ap, ap) + dm, v);
    float b[3]; ZB_VERT(b, ib);
    const float amb[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    float ab[3]; cross3(ab, nl, amb);
    const int ic = gargmax(fabsf(dot3(ap, ab)) + dm, v);
    float c[3]; ZB_VERT(c, ic);
    const float amc[3] = {a[0] - c[0], a[1] - c[1], a[2] - c[2]}, bmc[3] = {b[0] - c[0], b[1] - c[1], b[2] - c[2]};
    float ac[3], bc[3];
    cross3(ac, nl, amc); cross3(bc, nl, bmc);
    const float bpv[3] = {b[0] - pv[0], b[1] - pv[1], b[2] - pv[2]};
    const float v1 = fabsf(dot3(bpv, bc)) + dm, v2 = fabsf(dot3(ap, ac)) + dm;
    const int id = gargmax(v2 > v1 ? v2 : v1, v2 > v1 ? 8 + v : v) & 7;
    // this lane as contact row: slot s of box g
    const int s = (L >> 2) & 3;
    const int mi = s == 0 ? ia : (s == 1 ? ib : (s == 2 ? ic : id));
    const bool uniq = s == 0 ? true : (s == 1 ? ib != ia : (s == 2 ? (ic != ia && ic != ib) : (id != ia && id != ib && id != ic)));
    float mvv[3]; ZB_VERT(mvv, mi);
    const float dist = uniq ? -ZB_SUPPORT(mvv) : 1.f;
    float vw[3];
    mat_vec(vw, bm, mvv);
    if (!GEN || D->grp_kind[g] == ZB_PAIR_PLANE_BOX) {
      cdist = dist;
      cpos[0] = bp[0] + vw[0] - 0.5f * dist * pn[0];
      cpos[1] = bp[1] + vw[1] - 0.5f * dist * pn[1];
      cpos[2] = bp[2] + vw[2] - 0.5f * dist * pn[2];
    }
#undef ZB_VERT
#undef ZB_SUPPORT
  }

  // constraint._kbi impedance with host-precomputed constants (A.8)
  ZB_DI float impedance(const ZbSolPrm& s, float pos) const {
    const float x = fabsf(pos) / s.width;
    const float z = x < s.mid ? x : 1.f - x;
    const float pw = s.power == 2.f ? z * z : zb_pow(z, s.power);
    const float y = x < s.mid ? s.c1 * pw : 1.f - s.c2 * pw;
    float imp = s.dmin + y * (s.dmax - s.dmin);
    imp = clampf(imp, s.dmin, s.dmax);
    return x > 1.f ? s.dmax : imp;
  }

  // friction coefficient of this lane's contact (geom-priority mixing done on the host: con_mu_mode)
  ZB_DI float contact_mu() const {
    const int g = L >> 4;
    const float cf = D->con_mu_const[g];
    return D->con_mu_mode[g] == 0 ? mu : (D->con_mu_mode[g] == 1 ? cf : fmaxf(mu, cf));
  }
  // w = [(pos - c0) x dir ; dir], dir = n +- mu t: the wrench (about the tree COM) a unit force along this lane's pyramid edge exerts
  ZB_DI void row_wrench() {
    const float cmu = contact_mu();
    const int e = L & 3;
    float sgn = (e & 1) ? -1.f : 1.f;
    if (ZB_MUTANT == 1 && ((L >> 2) & 3) == 2 && e == 3) sgn = 1.f;  // (mutant 1: manifold slot 2 loses a facet of its friction pyramid)
    float dir[3];
    if (GEN) {
      dir[0] = cn[0] + sgn * cmu * ct[0]; dir[1] = cn[1] + sgn * cmu * ct[1]; dir[2] = cn[2] + sgn * cmu * ct[2];
    } else {  // the floor frame, a model constant
      const float* fr = D->frame;
      const float* tt = fr + 3 * (1 + (e >> 1));
      dir[0] = fr[0] + sgn * cmu * tt[0]; dir[1] = fr[1] + sgn * cmu * tt[1]; dir[2] = fr[2] + sgn * cmu * tt[2];
    }
    const float off[3] = {cpos[0] - c0[0], cpos[1] - c0[1], cpos[2] - c0[2]};
    float tq[3];
    cross3(tq, off, dir);
    w[0] = tq[0]; w[1] = tq[1]; w[2] = tq[2]; w[3] = dir[0]; w[4] = dir[1]; w[5] = dir[2];
  }

  ZB_DI void make_constraint() {
    const ZbModel& m = D->m;
    const int d = hasdof ? dof : 0;
    // friction-loss row of my dof: pos == 0 -> impedance, R, D and b are model constants (host-precomputed)
    fr_on = hasdof && D->fr_R[L] > 0.f;
    fr_D = 0.f; fr_rf = 0.f; fr_f = 0.f; fr_aref = 0.f;
    if (fr_on) {
      const float R = D->fr_R[L];
      fr_D = 1.f / R;
      fr_f = floss;
      fr_rf = R * floss;
      fr_aref = -D->fr_b[L] * qv;
    }
    // joint-limit row of my hinge
    lm_on = false; lm_D = 0.f; lm_aref = 0.f; lm_sign = 1.f;
    if (hinge && m.jnt_limited[d]) {
      const float dmin = q - m.jnt_range[d][0], dmax = m.jnt_range[d][1] - q;
      const float pos = fminf(dmin, dmax) - m.jnt_margin[d];
      if (pos < 0.f) {
        const ZbSolPrm& sp = D->lim[L];
        const float imp = impedance(sp, pos);
        lm_on = true;
        lm_sign = dmin < dmax ? 1.f : -1.f;
        const float R = fmaxf(m.dof_invweight0[d] * (1.f - imp) / imp, MJ_MINVAL);
        lm_D = 1.f / R;
        lm_aref = -sp.b * (lm_sign * qv) - sp.k * imp * pos;
      }
    }
    any_limit = __any_sync(ZB_FULL, lm_on);
    // contact row r = L: pyramid edge (L & 3) of slot (L >> 2)
    {
      const int g = L >> 4;
      const ZbSolPrm& sp = D->con[g];
      const float cmu = contact_mu();
      const float dist = cdist - D->con_includemargin[g];
      c_on = dist < 0.f;
      row_wrench();
      // velocity of the box body (needs velocity() first)
      float vf[6];
#pragma unroll
      for (int k = 0; k < 6; k++) vf[k] = shf(cvel[k], boxlane);
      if (any_pair) {  // (kernel-uniform) geom 1 on a body: relative velocity (cvel of the world lane is zero)
#pragma unroll
        for (int k = 0; k < 6; k++) vf[k] -= shf(cvel[k], pairlane);
      }
      c_D = 0.f; c_aref = 0.f;
      if (c_on) {
        const float imp = impedance(sp, dist);
        const float t = D->con_invweight[g];
        const float invweight = (t + cmu * cmu * t) * 2.f * cmu * cmu / o.impratio;
        const float R = fmaxf(invweight * (1.f - imp) / imp, MJ_MINVAL);
        c_D = 1.f / R;
        c_aref = -sp.b * dot6(w, vf) - sp.k * imp * dist;
      }
    }
  }

  // ------------------------------------------------------------------------------------------
  // solver.solve, CG (A.9) — ONE loop, every heavy primitive instantiated once (I-cache):
  //   phase -1 : qacc_smooth = M^-1 qfrc_smooth
  //   phase 0/1: cost of the two warm-start candidates (qacc_smooth, qacc_warmstart)
  //   phase 2  : Context.create at the chosen start + first gradient/search direction
  //   phase 3  : one solver iteration (line search, constraint update, gradient, Polak-Ribiere)
  //   after the loop: implicit integration solve (M + dt D)^-1 (qfrc_smooth + qfrc_constraint)
  // ------------------------------------------------------------------------------------------
  // forces and cost of the rows at the current Jaref; returns this lane's partial constraint cost
  ZB_DI float row_update(float& fr_force, float& lm_force, float& c_force) const {
    float cst = 0.f;
    fr_force = 0.f; lm_force = 0.f; c_force = 0.f;
    if (fr_on) {
      const float x = fr_ja;
      if (x <= -fr_rf && fr_f > 0.f) { fr_force = fr_f; cst += fr_f * (-0.5f * fr_rf - x); }
      else if (x >= fr_rf && fr_f > 0.f) { fr_force = -fr_f; cst += fr_f * (-0.5f * fr_rf + x); }
      else { fr_force = -fr_D * x; cst += 0.5f * fr_D * x * x; }
    }
    if (lm_on && lm_ja < 0.f) { lm_force = -lm_D * lm_ja; cst += 0.5f * lm_D * lm_ja * lm_ja; }
    if (c_on && c_ja < 0.f) { c_force = -c_D * c_ja; cst += 0.5f * c_D * c_ja * c_ja; }
    return cst;
  }
  // efc_force of this lane's rows at the final Jaref (what the last solver pass computed), for the sensors and the export
  ZB_DI void final_rows() {
    float a, b, c;
    row_update(a, b, c);
    fr_force = a; lm_force = b; c_force = c;
  }
  // is the friction row of this lane on its quadratic piece (Newton Hessian)
  ZB_DI bool fr_quadratic() const { return !(fr_f > 0.f && (fr_ja <= -fr_rf || fr_ja >= fr_rf)); }
  // MJX: (x < y && y < 0) || (x > y && y > 0)  <=>  y lies strictly between x and 0  <=>  (y - x) * y < 0.
  // (the sign of a rounded difference/product is exact; a product underflowing to 0 would need |d0| < 1e-19)
  static ZB_DI bool in_bracket(float x, float y) { return __fmul_rn(y - x, y) < 0.f; }

  // Per-row quadratic data of one line search (MJX: quad = [0.5 D x0^2, D x0 jv, 0.5 D jv^2]; friction rows add two
  // linear zones).
  struct LSRows {
    float fr_jv, lm_jv, c_jv;
    float f1, f2, fl, c1, c2, l1, l2;  // D*x0*jv, 0.5*D*jv^2 (and f*jv) of the friction / contact / limit row of this lane
    float qg0, qg1, qg2;
  };

  // Which piece of its piecewise-quadratic cost each row of this lane is on at step size a:
  // bit0/bit1 friction row in the negative/positive linear zone, bit2 limit row active, bit3 contact row active.
  ZB_DI int zone_at(const LSRows& r, float a) const {
    int z = 0;
    if (fr_on && fr_f > 0.f) {
      const float x = fr_ja + a * r.fr_jv;
      z |= (x <= -fr_rf) ? 1 : 0;
      z |= (x >= fr_rf) ? 2 : 0;
    }
    if (ZB_MUTANT != 3 && any_limit && lm_on && (lm_ja + a * r.lm_jv) < 0.f) z |= 4;  // (mutant 3: limit rows invisible to the line search)
    if (c_on && (c_ja + a * r.c_jv) < 0.f) z |= 8;
    return z;
  }

  // Point of the 1-D model at step size a on the active set S.
  static ZB_DI LSPoint ls_make_point(float a, const LSSet& S) {
    LSPoint p;
    p.alpha = a;
    // NOT fused on purpose (__fmul_rn/__fadd_rn are never contracted into FMA): when the active set did not change, the
    // Newton point alpha = -t1 / (2 t2) makes 2*alpha*t2 + t1 round to EXACTLY 0 with separately rounded products, and
    // the reference's bracket logic treats an exact zero as "not in bracket" (the point is rejected).  An FMA would
    // return the tiny residual instead and flip that decision in ~14 % of the states (scripts/gpu_probe3.py).
    p.d0 = __fadd_rn(__fmul_rn(__fmul_rn(2.f, a), S.t2), S.t1);
    return p;
  }
  static ZB_DI float ls_set_d1(float t2) { return __fadd_rn(__fmul_rn(2.f, t2), (t2 == 0.f ? MJ_MINVAL : 0.f)); }
  static ZB_DI float ls_cost(float a, const LSSet& S) {
    return __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(a, a), S.t2), __fmul_rn(a, S.t1)), S.t0);
  }

  // Full evaluation of NP step sizes: sum the quadratics of the rows that are active at each of them.
  template <int NP>
  ZB_DI void ls_eval(const LSRows& r, const float (&al)[3], const int (&zz)[3], LSPoint (&pt)[3], LSSet (&st)[3]) {
    constexpr int NS = NP == 1 ? 4 : 16;
    float s[NS];
#pragma unroll
    for (int j = 0; j < NP; j++) {
      const int z = zz[j];
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      if (fr_on) {
        const bool neg = (z & 1) != 0, pos = (z & 2) != 0, quad = !(neg || pos);
        s0 = quad ? 0.5f * fr_D * fr_ja * fr_ja : fr_f * (-0.5f * fr_rf + (neg ? -fr_ja : fr_ja));
        s1 = quad ? r.f1 : (neg ? -r.fl : r.fl);
        s2 = quad ? r.f2 : 0.f;
      }
      if (z & 4) { s0 += 0.5f * lm_D * lm_ja * lm_ja; s1 += r.l1; s2 += r.l2; }
      if (z & 8) { s0 += 0.5f * c_D * c_ja * c_ja; s1 += r.c1; s2 += r.c2; }
      s[3 * j] = s0; s[3 * j + 1] = s1; s[3 * j + 2] = s2;
    }
#pragma unroll
    for (int k = 3 * NP; k < NS; k++) s[k] = 0.f;
    reduce_scatter<NS>(s, L);
    constexpr int STR = 32 / NS;
#pragma unroll
    for (int j = 0; j < NP; j++) {
      st[j].t0 = shf(s[0], STR * (3 * j)) + r.qg0;
      st[j].t1 = shf(s[0], STR * (3 * j + 1)) + r.qg1;
      st[j].t2 = shf(s[0], STR * (3 * j + 2)) + r.qg2;
      st[j].z = zz[j];
      st[j].d1 = ls_set_d1(st[j].t2);
      pt[j] = ls_make_point(al[j], st[j]);
    }
  }

  // exact line search along `search` (MJX _linesearch): three candidate points per iteration.
  // 91 % of the iterations (measured on the bench workload) propose points whose rows all sit on the same quadratic pieces
  // as the bracket end they were derived from; their summed quadratic is then that end's, bit for bit what a fresh
  // summation over the same active set gives, so only a per-lane zone check and four votes are needed — no reduction.
  ZB_DI void linesearch(float cjv) {
    LSRows r;
    r.fr_jv = search; r.lm_jv = lm_sign * search; r.c_jv = cjv;
    r.f1 = fr_D * fr_ja * r.fr_jv; r.f2 = 0.5f * fr_D * r.fr_jv * r.fr_jv; r.fl = fr_f * r.fr_jv;
    r.l1 = lm_D * lm_ja * r.lm_jv; r.l2 = 0.5f * lm_D * r.lm_jv * r.lm_jv;
    r.c1 = c_D * c_ja * r.c_jv; r.c2 = 0.5f * c_D * r.c_jv * r.c_jv;
    LSPoint p0, lo, hi, pt[3];
    LSSet S0, Slo, Shi, st[3];
    float al[3] = {0.f, 0.f, 0.f};
    int zz[3] = {0, 0, 0};
    float snorm;
    {
      // |search|^2, the Gauss part of the quadratic and the row sums of p0 = point(0) in ONE transpose-reduce
      const int z = zone_at(r, 0.f);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      if (fr_on) {
        const bool neg = (z & 1) != 0, pos = (z & 2) != 0, quad = !(neg || pos);
        s0 = quad ? 0.5f * fr_D * fr_ja * fr_ja : fr_f * (-0.5f * fr_rf + (neg ? -fr_ja : fr_ja));
        s1 = quad ? r.f1 : (neg ? -r.fl : r.fl);
        s2 = quad ? r.f2 : 0.f;
      }
      if (z & 4) { s0 += 0.5f * lm_D * lm_ja * lm_ja; s1 += r.l1; s2 += r.l2; }
      if (z & 8) { s0 += 0.5f * c_D * c_ja * c_ja; s1 += r.c1; s2 += r.c2; }
      float p[8] = {search * search, search * Ma - search * qfrc_smooth, search * mv, s0, s1, s2, 0.f, 0.f};
      if (!hasdof) { p[0] = p[1] = p[2] = 0.f; }
      reduce_scatter<8>(p, L);  // value k in lanes 4k..4k+3
      snorm = shf(p[0], 0);
      r.qg0 = gauss; r.qg1 = shf(p[0], 4); r.qg2 = 0.5f * shf(p[0], 8);
      S0.t0 = shf(p[0], 12) + r.qg0; S0.t1 = shf(p[0], 16) + r.qg1; S0.t2 = shf(p[0], 20) + r.qg2; S0.z = z;
      S0.d1 = ls_set_d1(S0.t2);
      p0 = ls_make_point(0.f, S0);
    }
    const float gtol = o.tolerance * o.ls_tolerance * (sqrtf(snorm) * meaninertia * (float)NV);
    al[0] = p0.alpha - p0.d0 / S0.d1;
    zz[0] = zone_at(r, al[0]);
#if ZB_LS_FIRST
    // every row on the same piece as at step size 0: the summed quadratic of this point is S0's, bit for bit (same rows, same order)
    if (!(TRACE && trace && niter == 0) && !__any_sync(ZB_FULL, zz[0] != S0.z)) { st[0] = S0; pt[0] = ls_make_point(al[0], S0); }
    else
#endif
    ls_eval<1>(r, al, zz, pt, st);  // lo = point(newton(p0)); order the bracket
    if (pt[0].d0 < p0.d0) { lo = pt[0]; Slo = st[0]; hi = p0; Shi = S0; } else { hi = pt[0]; Shi = st[0]; lo = p0; Slo = S0; }
    if (TRACE && trace && niter == 0 && hasdof) trace[ZB_TRACE_SEARCH + dof] = search;
    if (TRACE && trace && niter == 0 && L == 0) {
      trace[32] = p0.alpha; trace[33] = p0.d0; trace[34] = S0.d1; trace[35] = pt[0].alpha; trace[36] = pt[0].d0; trace[37] = st[0].d1;
      lscount = 2;
    }
    bool swap = true;
    int it = 0;
#if ZB_LS_SAME
    // `same`: both bracket ends sit on the same piece of EVERY row of the warp (their active sets are identical).  Each row's
    // piece is a monotone step function of the step size (x0 + a jv is monotone in a, rounding included), so every step size
    // BETWEEN the two ends then lies on those pieces as well: no zone evaluation, no votes, and the iteration is pure
    // warp-uniform scalar work.  Measured with the oracle on the bench workload (oracle set_variant(4)): 92 % of the bracket
    // iterations start in this state, 83 % also propose all three candidates inside the bracket, and 65 % have one end resting
    // on an exactly zero derivative (then only the other end can still move: the one-sided loop below).  Bit for bit the same
    // decisions and numbers as the general iteration.
    bool same = !__any_sync(ZB_FULL, Slo.z != Shi.z);
#endif
#pragma unroll 1
    while (true) {
#if ZB_LS_SAME >= 2
      if (!(TRACE && trace && niter == 0)) {
        const bool zlo = lo.d0 == 0.f, zhi = hi.d0 == 0.f;
        if (__all_sync(ZB_FULL, same && (zlo != zhi))) {
          // ---- one-sided loop: the end Z rests on the exact minimum of its quadratic (d0 == 0), so its Newton candidate is Z
          // itself and no candidate can replace it (in_bracket(0, y) and in_bracket(x, 0) are false); only the other end M moves:
          // first to its Newton candidate, then to the midpoint, exactly the s1/s2 (M = lo) or s4/s5 (M = hi) tests of the
          // general iteration.  The midpoint is evaluated on lo's active set, and hi takes that set along with the midpoint.
          const bool mhi = zlo;
          float Ma_ = mhi ? hi.alpha : lo.alpha, Md = mhi ? hi.d0 : lo.d0;
          const float Za = mhi ? lo.alpha : hi.alpha, sg = mhi ? 1.f : -1.f;
          float m0 = mhi ? Shi.t0 : Slo.t0, m1 = mhi ? Shi.t1 : Slo.t1, m2 = mhi ? Shi.t2 : Slo.t2, md1 = mhi ? Shi.d1 : Slo.d1;
          const float z0 = mhi ? Slo.t0 : Shi.t0, z1 = mhi ? Slo.t1 : Shi.t1, z2 = mhi ? Slo.t2 : Shi.t2, zd1 = mhi ? Slo.d1 : Shi.d1;
#pragma unroll 1
          while (true) {
            bool dn = it >= o.ls_iterations;
            dn |= !swap;
            const float sd = sg * Md;
            dn |= (sd > 0.f) && (sd < gtol);
            if (dn) break;
            const bool nz = Md != 0.f;
            float nn = nz ? Md : 1.f;
            asm volatile("" : "+f"(nn));
            float q = nn / md1;
            q = nz ? q : 0.f;
            const float c1 = Ma_ - q, c2 = 0.5f * (Za + Ma_);
            const float amin = fminf(Za, Ma_), amax = fmaxf(Za, Ma_);
            if (!(c1 >= amin && c1 <= amax)) break;  // Newton candidate outside the bracket: its pieces must be looked at (general iteration)
            const float e1 = __fadd_rn(__fmul_rn(__fmul_rn(2.f, c1), m2), m1);
            const float e2 = mhi ? __fadd_rn(__fmul_rn(__fmul_rn(2.f, c2), z2), z1) : __fadd_rn(__fmul_rn(__fmul_rn(2.f, c2), m2), m1);
            const bool t1 = in_bracket(Md, e1);
            if (t1) { Ma_ = c1; Md = e1; }
            const bool t2 = in_bracket(Md, e2);
            if (t2) { Ma_ = c2; Md = e2; if (mhi) { m0 = z0; m1 = z1; m2 = z2; md1 = zd1; } }
            swap = t1 | t2;
            it++;
          }
          if (mhi) { hi.alpha = Ma_; hi.d0 = Md; Shi.t0 = m0; Shi.t1 = m1; Shi.t2 = m2; Shi.d1 = md1; }
          else if (ZB_MUTANT == 4) { hi.alpha = Ma_; hi.d0 = Md; }  // (mutant 4: the moved end is written back to the wrong side)
          else { lo.alpha = Ma_; lo.d0 = Md; }
        }
      }
#endif
      bool done = it >= o.ls_iterations;
      done |= !swap;
      done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
      done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
      // (every lane holds the same values; the vote only makes the exit PROVABLY warp-uniform, otherwise ptxas wraps each
      //  shuffle of the loop in WARPSYNC/ENDCOLLECTIVE)
      if (__all_sync(ZB_FULL, done)) break;
      // Newton steps from both ends.  A bracket end often sits on an EXACTLY zero derivative (see ls_make_point); 0 / d1
      // sends the IEEE division into its out-of-line special-case path (same result, ~40 instructions more).
      // (ptxas if-converts a guarding branch and the FCHK of the inline division still sees the zero: 68 % of the
      //  iterations called the slow path, profiles/r1_g.  So the division itself is fed a non-zero numerator.)
      //  The empty asm keeps the compiler from folding the two selects back into d0 / d1.)
      const bool nz_lo = lo.d0 != 0.f, nz_hi = hi.d0 != 0.f;
      float n_lo = nz_lo ? lo.d0 : 1.f, n_hi = nz_hi ? hi.d0 : 1.f;
      asm volatile("" : "+f"(n_lo), "+f"(n_hi));
      float q_lo = n_lo / Slo.d1, q_hi = n_hi / Shi.d1;
      q_lo = nz_lo ? q_lo : 0.f;
      q_hi = nz_hi ? q_hi : 0.f;
      al[0] = lo.alpha - q_lo;
      al[1] = hi.alpha - q_hi;
      al[2] = 0.5f * (lo.alpha + hi.alpha);
      bool f0 = false, f1 = false, f2l = false, f2h = false;
#if ZB_LS_SAME
      const float amin = fminf(lo.alpha, hi.alpha), amax = fmaxf(lo.alpha, hi.alpha);
      const bool inside = same && al[0] >= amin && al[0] <= amax && al[1] >= amin && al[1] <= amax && al[2] >= amin && al[2] <= amax;
      if (!__all_sync(ZB_FULL, inside))
#endif
      {
        zz[0] = zone_at(r, al[0]); zz[1] = zone_at(r, al[1]); zz[2] = zone_at(r, al[2]);
        f0 = __any_sync(ZB_FULL, zz[0] != Slo.z); f1 = __any_sync(ZB_FULL, zz[1] != Shi.z);
        f2l = __any_sync(ZB_FULL, zz[2] != Slo.z); f2h = __any_sync(ZB_FULL, zz[2] != Shi.z);
      }
      if (!f0 && !f1 && !(f2l && f2h)) {
        // ---- fast path: every candidate lives on the active set of a bracket end ----
        const LSPoint lo_next = ls_make_point(al[0], Slo), hi_next = ls_make_point(al[1], Shi);
        const LSPoint mid = f2l ? ls_make_point(al[2], Shi) : ls_make_point(al[2], Slo);
        pt[0] = lo_next; pt[1] = hi_next; pt[2] = mid;
        if (TRACE) { st[0].d1 = Slo.d1; st[1].d1 = Shi.d1; st[2].d1 = f2l ? Shi.d1 : Slo.d1; }
        bool lo_hi = false, hi_lo = false;  // does the new lo (hi) take the OTHER end's active set?
        const LSPoint hi_old = hi;
        const bool s1 = in_bracket(lo.d0, lo_next.d0); if (s1) lo = lo_next;
        const bool s2 = in_bracket(lo.d0, mid.d0); if (s2) { lo = mid; lo_hi = f2l; }
        const bool s3 = in_bracket(lo.d0, hi_next.d0); if (s3) { lo = hi_next; lo_hi = true; }
        const bool s4 = in_bracket(hi_old.d0, hi_next.d0); if (s4) hi = hi_next;
        const bool s5 = in_bracket(hi.d0, mid.d0); if (s5) { hi = mid; hi_lo = !f2l; }
        const bool s6 = in_bracket(hi.d0, lo_next.d0); if (s6) { hi = lo_next; hi_lo = true; }
        swap = s1 | s2 | s3 | s4 | s5 | s6;
        const LSSet a = Slo, b = Shi;
        if (lo_hi) Slo = b;
        if (hi_lo) Shi = a;
#if ZB_LS_SAME
        same = same || (lo_hi != hi_lo);  // one end took over the other's set: identical; both did: the sets only changed sides
#endif
      } else {
        ls_eval<3>(r, al, zz, pt, st);
        const LSPoint lo_next = pt[0], hi_next = pt[1], mid = pt[2];
        const bool s1 = in_bracket(lo.d0, lo_next.d0); if (s1) { lo = lo_next; Slo = st[0]; }
        const bool s2 = in_bracket(lo.d0, mid.d0); if (s2) { lo = mid; Slo = st[2]; }
        const bool s3 = in_bracket(lo.d0, hi_next.d0); if (s3) { lo = hi_next; Slo = st[1]; }
        const bool s4 = in_bracket(hi.d0, hi_next.d0); if (s4) { hi = hi_next; Shi = st[1]; }
        const bool s5 = in_bracket(hi.d0, mid.d0); if (s5) { hi = mid; Shi = st[2]; }
        const bool s6 = in_bracket(hi.d0, lo_next.d0); if (s6) { hi = lo_next; Shi = st[ZB_MUTANT == 2 ? 1 : 0]; }  // (mutant 2: wrong active set)
        swap = s1 | s2 | s3 | s4 | s5 | s6;
#if ZB_LS_SAME
        same = !__any_sync(ZB_FULL, Slo.z != Shi.z);
#endif
      }
      if (TRACE && trace && niter == 0 && L == 0) {
        for (int j = 0; j < 3; j++)
          if (32 + 3 * lscount + 2 < ZB_TRACE_SEARCH) {
            trace[32 + 3 * lscount] = pt[j].alpha; trace[32 + 3 * lscount + 1] = pt[j].d0; trace[32 + 3 * lscount + 2] = st[j].d1;
            lscount++;
          }
      }
      it++;
    }
    const float c0_ = ls_cost(p0.alpha, S0), cl = ls_cost(lo.alpha, Slo), ch = ls_cost(hi.alpha, Shi);
    const bool improved = (cl < c0_) || (ch < c0_);
    const float alpha = cl < ch ? lo.alpha : hi.alpha;
    if (TRACE && trace && L == 0 && 4 * niter + 3 < 32) {
      trace[4 * niter] = improved ? alpha : 0.f;
      trace[4 * niter + 2] = (float)it;
      trace[4 * niter + 3] = sqrtf(snorm);
    }
    if (improved) {
      qacc += search * alpha;
      Ma += mv * alpha;
      fr_ja += r.fr_jv * alpha;
      lm_ja += r.lm_jv * alpha;
      c_ja += r.c_jv * alpha;
    }
  }

  // Newton: H = M + J^T diag(D * active) J (solver._update_gradient), same tree block structure as M.
  // Contact rows enter through the 6x6 "constraint inertia" of each box, Wm_g = sum_r D_r w_r w_r^T, exactly like the
  // composite inertia enters the mass matrix; friction-loss / limit rows add D on the diagonal.
  ZB_DI void newton_hessian(float (&Hb)[6], float (&Hc)[CHD]) {
    const float a = (c_on && c_ja < 0.f) ? c_D : 0.f;
    float p16[16], p8[8];
    {
      int t = 0;
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) {
          const float v = a * w[i] * w[j];
          if (t < 16) p16[t] = v; else p8[t - 16] = v;
          t++;
        }
      p8[5] = p8[6] = p8[7] = 0.f;
    }
    // sums over the 16 row lanes of each box: masks 8,4,2,1 only (the two halves stay separate)
    {
      int m = 8;
#pragma unroll
      for (int half = 8; half >= 1; half >>= 1, m >>= 1) {
        const bool up = (L & m) != 0;
#pragma unroll
        for (int j = 0; j < half; j++) {
          const float send = up ? p16[j] : p16[j + half], keep = up ? p16[j + half] : p16[j];
          p16[j] = keep + __shfl_xor_sync(ZB_FULL, send, m);
        }
      }
      m = 8;
#pragma unroll
      for (int half = 4; half >= 1; half >>= 1, m >>= 1) {
        const bool up = (L & m) != 0;
#pragma unroll
        for (int j = 0; j < half; j++) {
          const float send = up ? p8[j] : p8[j + half], keep = up ? p8[j + half] : p8[j];
          p8[j] = keep + __shfl_xor_sync(ZB_FULL, send, m);
        }
      }
      p8[0] += __shfl_xor_sync(ZB_FULL, p8[0], 1);
    }
    // value t < 16 of box g sits in lane 16 g + t, value 16 + u in lanes 16 g + 2u (+1): publish through the scratch
    __syncwarp();
    {
      const int g = L >> 4;
      scr[g * 24 + (L & 15)] = p16[0];
      if ((L & 1) == 0 && ((L >> 1) & 7) < 5) scr[g * 24 + 16 + ((L >> 1) & 7)] = p8[0];
    }
    __syncwarp();
    float Wm[21];
#pragma unroll
    for (int t = 0; t < 21; t++) {
      const float w0 = scr[t], w1 = scr[24 + t];
      Wm[t] = bdof ? w0 + w1 : (hinge && lbox == 0 ? w0 : (hinge && lbox == 1 ? w1 : 0.f));
    }
    __syncwarp();
    float f[6];
    {
      float Mx[6][6];
      int t = 0;
#pragma unroll
      for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = i; j < 6; j++) { Mx[i][j] = Wm[t]; Mx[j][i] = Wm[t]; t++; }
#pragma unroll
      for (int i = 0; i < 6; i++) f[i] = Mx[i][0] * cdof[0] + Mx[i][1] * cdof[1] + Mx[i][2] * cdof[2] + Mx[i][3] * cdof[3] + Mx[i][4] * cdof[4] + Mx[i][5] * cdof[5];
    }
    assemble_rows(f, Hb, Hc);
    const float dg = (fr_on && fr_quadratic() ? fr_D : 0.f) + (lm_on && lm_ja < 0.f ? lm_D : 0.f);
#pragma unroll
    for (int k = 0; k < CHD; k++) {
      Hc[k] = hinge ? Hc[k] + fget(F_C + k) + (k == gp ? dg : 0.f) : 0.f;
      Hb[k] = Hb[k] + fget(F_BM + k) + ((bdof && k == gp) ? dg : 0.f);
    }
  }

  // The whole constrained solve + (optionally) the implicit-integration solve.  Returns the integration acceleration.
  // (The solver loop exists twice in the instruction stream, once per solver type, selected by a warp-uniform branch: the CG loop
  //  then carries no Hessian assembly / factorisation and, with the integration solve moved behind the loop, is 1.6 k instructions
  //  = 26 KB instead of 3.1 k = 50 KB.  It is executed ~12 times per substep by every warp of the CTA and now fits the 32 KB
  //  instruction cache level below L0: +2.7 % at S = 4, +4 % at S = 20, bit-identical results.)
  ZB_DI float solve_all(bool implicit_damp) {
    if (o.solver == ZB_SOLVER_NEWTON) solver_loop<true>();
    else solver_loop<false>();
    // implicit integration: (M + dt D)^-1 (qfrc_smooth + qfrc_constraint).  M + dt D is factored only here, when the solve has
    // finished: an up-front factor would sit in 19 registers per lane through the whole solve.  F0 (factor of M / of the last
    // Newton Hessian) is dead from here on.
    if (!implicit_damp) return qacc;
    float Hb[6], Hc[CHD];
    load_m(Hb, Hc);
    factor(Hb, Hc, o.dt * damp);
    return solve(qfrc_smooth + qfrc_c);
  }
  template <bool newton>
  ZB_DI void solver_loop() {
    const float scale = 1.f / (meaninertia * (float)NV);
    int phase = -1;
    float v = 0.f, cost_s = 0.f, pg = 0.f, pMg = 0.f;
    cost = 0.f; prev_cost = 0.f; niter = 0; lscount = 0;
    qacc = 0.f; Ma = 0.f; fr_ja = lm_ja = c_ja = 0.f; qfrc_c = 0.f; grad = 0.f; Mgrad = 0.f; search = 0.f; mv = 0.f; gauss = 0.f;
#pragma unroll 1
    while (true) {
      float fr_force = 0.f, lm_force = 0.f, c_force = 0.f;  // row forces of THIS pass
      if (phase >= 0) {
        // ---- operator application: mv = M v, contact rows J v ----
        float vb[6];
        mv = mul_m(v, vb);
        const float cjv = contact_jv(v, vb);
        if (phase >= 3) {
          linesearch(cjv);
        } else {  // Context.create at qacc = v
          qacc = hasdof ? v : 0.f;
          Ma = mv;
          fr_ja = v - fr_aref;
          lm_ja = lm_sign * v - lm_aref;
          c_ja = cjv - c_aref;
        }
        // ---- _update_constraint: forces, cost ----
        const float part = row_update(fr_force, lm_force, c_force);
        float p2[2] = {part, hasdof ? 0.5f * (Ma - qfrc_smooth) * (qacc - qacc_smooth) : 0.f};
        reduce_scatter<2>(p2, L);
        gauss = shf(p2[0], 16);
        prev_cost = phase == 2 ? INFINITY : cost;
        cost = gauss + shf(p2[0], 0);
        if (TRACE && trace && L == 0 && phase >= 3 && 4 * niter + 3 < 32) trace[4 * niter + 1] = cost;
      }
      float x = 0.f;
      if (phase == -1 || phase >= 2) {
        if (phase >= 2) {
          const float jt = contact_jtf(c_force);
          qfrc_c = hasdof ? jt + (fr_on ? fr_force : 0.f) + (lm_on ? lm_sign * lm_force : 0.f) : 0.f;
          pg = grad; pMg = Mgrad;
          grad = hasdof ? Ma - qfrc_smooth - qfrc_c : 0.f;
        }
        // Newton: F0 (factor of M) is dead once qacc_smooth exists and is overwritten by the factor of every iteration's Hessian
        if (newton && phase >= 2) {
          float Hb[6], Hc[CHD];
          newton_hessian(Hb, Hc);
          factor(Hb, Hc, 0.f);
        }
        const float rhs = phase == -1 ? qfrc_smooth : grad;
        x = solve(rhs);
      }
      // ---- transitions ----
      if (phase == -1) { qacc_smooth = x; v = x; phase = o.warmstart ? 0 : 2; continue; }
      if (phase == 0) { cost_s = cost; v = warm; phase = 1; continue; }
      if (phase == 1) {
        const bool pick_warm = __all_sync(ZB_FULL, cost < cost_s);
        if (TRACE && trace && L == 0) trace[ZB_TRACE_START] = pick_warm ? 1.f : 2.f;
        v = pick_warm ? warm : qacc_smooth; warm = 0.f; /* dead from here on: frees its register for the iterations */ phase = 2; continue; }
      Mgrad = x;
      float gnorm2;
      if (phase == 2 || newton) {
        gnorm2 = wsum(grad * grad);
        search = -Mgrad;
        if (phase != 2) niter++;
      } else {  // Polak-Ribiere+ and the gradient norm of the convergence test in one transpose-reduce
        float p4[4] = {grad * (Mgrad - pMg), pg * pMg, grad * grad, 0.f};
        reduce_scatter<4>(p4, L);
        const float num = shf(p4[0], 0), den = shf(p4[0], 8);
        gnorm2 = shf(p4[0], 16);
        const float beta = fmaxf(0.f, num / fmaxf(MJ_MINVAL, den));
        search = -Mgrad + beta * search;
        niter++;
      }
      const float improvement = scale * (prev_cost - cost);
      const float gradient = scale * sqrtf(gnorm2);
      bool done = niter >= o.iterations;
      done |= improvement < o.tolerance;
      done |= gradient < o.tolerance;
      if (__all_sync(ZB_FULL, done)) break;
      phase = 3;
      v = search;
    }
  }

  // ------------------------------------------------------------------------------------------
  // ------------------------------------------------------------------------------------------
  // sensors + rne_postconstraint (A.11); only evaluated when outputs are written
  // ------------------------------------------------------------------------------------------
  // keep: shared [16] (or nullptr) that receives framequat (4), accelerometer (3), gyro (3), touch l / r (2) for the observation epilogue —
  // shared memory, not registers: the values would otherwise stay live across the substep loop and the solver
  // (has_sd / has_keep: the same facts as `sd != nullptr` / `keep != nullptr`, but derived from kernel parameters only.  The pointers
  //  themselves depend on the warp index, and a branch on them is not PROVABLY warp-uniform: ptxas then treats everything behind it
  //  as potentially diverged, wraps every shuffle of the kernel in WARPSYNC and duplicates the solver — the observation
  //  instantiation was 18.6 k instructions instead of 9.6 k and 10 % slower because of exactly that.)
  ZB_DI void sensors(float* sd /* global [18] or nullptr */, float* keep, bool has_sd, bool has_keep) {
    if (!has_sd && !has_keep) return;
    if (has_keep) {  // touch values default to 0 (a foot box without a touch site)
      if (L == 0) { keep[10] = 0.f; keep[11] = 0.f; }
      __syncwarp();
    }
    const ZbModel& m = D->m;
    // cacc with qacc: base part uniform, chain part by prefix scan
    float ab[6];
#pragma unroll
    for (int a = 0; a < 6; a++) ab[a] = shf(qacc, BL0 + a);
    float Ab[6];
    base_motion(Ab, ab);  // sum_a cdof_a qacc_a
    float pre[6];
#pragma unroll
    for (int k = 0; k < 6; k++) pre[k] = hinge ? cdofd[k] * qv + cdof[k] * qacc : 0.f;
    chain_prefix<6>(pre);
    float cacc[6];
#pragma unroll
    for (int k = 0; k < 6; k++) cacc[k] = (chainbody ? pre[k] : 0.f) + Ab[k];
    cacc[3] += cacc_base_lin[0]; cacc[4] += cacc_base_lin[1]; cacc[5] += cacc_base_lin[2];
    if (L == 0) { cacc[0] = cacc[1] = cacc[2] = 0.f; cacc[3] = -o.gravity[0]; cacc[4] = -o.gravity[1]; cacc[5] = -o.gravity[2]; }
    // cfrc_int = local - cfrc_ext, accumulated leaf -> root along the chains
    float cf[6], t[6], u[6];
    inert_mul(cf, cinert, cacc);
    inert_mul(t, cinert, cvel);
    motion_cross_force(u, cvel, t);
    const int mybox = isbody ? D->body_box[L] : -1;
#if ZB_EXPLICIT_J
#if ZB_RECOLLIDE
    collision();   // contact point and frame recomputed from the (still live) body poses instead of living in registers across the solve
#endif
    row_wrench();  // recomputed here (20 instructions, last substep only) instead of living in 6 registers across the solve
#endif
    const float Wtot = contact_wrench(c_force);  // cfrc_ext of a foot = the solver's total contact wrench of its group
    const int mypair = (any_pair && isbody) ? D->body_pair[L] : -1;  // ... and the body of geom 1 feels the reaction
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const float ext = shf(Wtot, (mybox >= 0 ? 16 * mybox : 0) + 2 * k);
      cf[k] = isbody ? cf[k] + u[k] - (mybox >= 0 ? ext : 0.f) : 0.f;
      if (any_pair) {
        const float rea = shf(Wtot, (mypair >= 0 ? 16 * mypair : 0) + 2 * k);
        cf[k] += mypair >= 0 ? rea : 0.f;
      }
    }
    chain_suffix<6>(cf);
    const int s = isbody ? D->body_site[L] : -1;
    if (s == 0) {  // imu_site: framequat, accelerometer, gyro
      float qq[4];
      quat_mul(qq, xquat, squat);
      normalize4(qq);
      float simu[10];
      simu[0] = qq[0]; simu[1] = qq[1]; simu[2] = qq[2]; simu[3] = qq[3];
      float gy[3];
      matT_vec(gy, sxmat, cvel);
      simu[7] = gy[0]; simu[8] = gy[1]; simu[9] = gy[2];
      const float diff[3] = {sxpos[0] - c0[0], sxpos[1] - c0[1], sxpos[2] - c0[2]};
      float ang[3], lin[3], acc[3], tt[3], uu[3], corr[3];
      matT_vec(ang, sxmat, cvel);
      cross3(tt, diff, cvel);
      uu[0] = cvel[3] - tt[0]; uu[1] = cvel[4] - tt[1]; uu[2] = cvel[5] - tt[2];
      matT_vec(lin, sxmat, uu);
      cross3(tt, diff, cacc);
      uu[0] = cacc[3] - tt[0]; uu[1] = cacc[4] - tt[1]; uu[2] = cacc[5] - tt[2];
      matT_vec(acc, sxmat, uu);
      cross3(corr, ang, lin);
      simu[4] = acc[0] + corr[0]; simu[5] = acc[1] + corr[1]; simu[6] = acc[2] + corr[2];
      if (has_keep) {
#pragma unroll
        for (int k = 0; k < 10; k++) keep[k] = simu[k];
      }
      if (has_sd) {
#pragma unroll
        for (int k = 0; k < 4; k++) sd[ZB_SENS_IMU_QUAT + k] = simu[k];
#pragma unroll
        for (int k = 0; k < 3; k++) { sd[ZB_SENS_IMU_ACC + k] = simu[4 + k]; sd[ZB_SENS_IMU_GYRO + k] = simu[7 + k]; }
      }
    }
    if (has_sd && (s == 1 || s == 2)) {  // foot force sensors
      float fo[3];
      matT_vec(fo, sxmat, cf + 3);
      const int o = s == 1 ? ZB_SENS_FORCE_L : ZB_SENS_FORCE_R;
      sd[o] = fo[0]; sd[o + 1] = fo[1]; sd[o + 2] = fo[2];
    }
    // touch: normal force (= sum of the four edge forces) of my half's contacts that lie inside the site box
    {
      const int g = L >> 4;
      const int ts = D->box_site[g];
      const int sl = ts >= 0 ? m.site_body[ts] : 0;
      float sp[3], sm[9];
#pragma unroll
      for (int k = 0; k < 3; k++) sp[k] = shf(sxpos[k], sl);
#pragma unroll
      for (int k = 0; k < 9; k++) sm[k] = shf(sxmat[k], sl);
      float val = 0.f;
      if (ts >= 0 && c_on) {
        const float rel[3] = {cpos[0] - sp[0], cpos[1] - sp[1], cpos[2] - sp[2]};
        float loc[3];
        matT_vec(loc, sm, rel);
        if (fabsf(loc[0]) <= m.site_size[ts][0] && fabsf(loc[1]) <= m.site_size[ts][1] && fabsf(loc[2]) <= m.site_size[ts][2]) val = c_force;
      }
      val += shx(val, 8); val += shx(val, 4); val += shx(val, 2); val += shx(val, 1);
      if ((L & 15) == 0 && ts >= 0) {
        if (has_sd) sd[ts == 1 ? ZB_SENS_TOUCH_L : ZB_SENS_TOUCH_R] = val;
        if (has_keep) keep[ts == 1 ? 10 : 11] = val;
      }
    }
  }

  // ------------------------------------------------------------------------------------------
  // Observation epilogue (include/zbphys.h ZbObs): run_actor / run_critic input vectors of train.py:1624-1681,
  // ImuOrientationObservation.observe_stateful train.py:837-866 with rotate_quat_by_quat train.py:751-787.
  // reinit: this environment was (re)initialised by the launch -> carry restarts from (0, reset lag) (train.py:846-848).
  // ------------------------------------------------------------------------------------------
  ZB_DI void observe(const ZbObs& OB, int env, bool reinit, const float* keep /* shared [16] written by sensors() */) {
    __syncwarp();
    float sq[10];
#pragma unroll
    for (int k = 0; k < 10; k++) sq[k] = keep[k];
    float cmd[ZB_NCOMMAND];
#pragma unroll
    for (int k = 0; k < ZB_NCOMMAND; k++) cmd[k] = OB.command ? OB.command[(size_t)env * ZB_NCOMMAND + k] : 0.f;
    float ob[4] = {0.f, 0.f, 0.f, 0.f};
    if (OB.imu_x) {
      // heading quaternion euler_to_quat([0, 0, yaw]) = (cos(yaw/2), 0, 0, sin(yaw/2)); both operands and the product are
      // normalised with q / (|q| + 1e-6)
      float sy, cy;
      sincosf(0.5f * cmd[3], &sy, &cy);
      const float nf = sqrtf(sq[0] * sq[0] + sq[1] * sq[1] + sq[2] * sq[2] + sq[3] * sq[3]) + 1e-6f;
      const float q2[4] = {sq[0] / nf, sq[1] / nf, sq[2] / nf, sq[3] / nf};
      const float nh = sqrtf(cy * cy + sy * sy) + 1e-6f;
      const float w1 = cy / nh, z1 = -(sy / nh);  // conjugate: x1 = y1 = -0
      float r[4] = {w1 * q2[0] - z1 * q2[3], w1 * q2[1] - z1 * q2[2], w1 * q2[2] + z1 * q2[1], w1 * q2[3] + z1 * q2[0]};
      const float nr = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]) + 1e-6f;
      const bool flip = r[0] / nr < 0.f;
      float lag = reinit ? (OB.reset_imu_lag ? OB.reset_imu_lag[env] : (OB.imu_lag ? OB.imu_lag[env] : 0.f)) : (OB.imu_lag ? OB.imu_lag[env] : 0.f);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const float b = flip ? -(r[k] / nr) : r[k] / nr;
        const float x0 = reinit ? 0.f : OB.imu_x[(size_t)env * 4 + k];
        ob[k] = x0 * lag + b * (1.f - lag);
      }
      __syncwarp();  // every lane has read the old carry
      if (L == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) OB.imu_x[(size_t)env * 4 + k] = ob[k];
      }
      if (L == 0 && OB.imu_lag && reinit && OB.reset_imu_lag) OB.imu_lag[env] = lag;
      if (OB.noise_imu_quat) {
#pragma unroll
        for (int k = 0; k < 4; k++) ob[k] += OB.noise_imu_quat[(size_t)env * 4 + k];
      }
    }
    float acc[3] = {sq[4], sq[5], sq[6]};
    if (OB.noise_imu_acc) {
#pragma unroll
      for (int k = 0; k < 3; k++) acc[k] += OB.noise_imu_acc[(size_t)env * 3 + k];
    }
    if (OB.actor_obs) {
      float* a = OB.actor_obs + (size_t)env * ZB_NACTOR_OBS;
      if (hinge) { a[dof - 6] = q; a[NU + dof - 6] = qv; }
      if (L == 0) {  // imu_orientation(4) | cmd[0:2] | cmd[3] | cmd[4:7]   (plain stores of one lane: a lane-selected
                     //  register array would be turned into a local-memory array)
        float* t = a + 2 * NU;
        t[0] = ob[0]; t[1] = ob[1]; t[2] = ob[2]; t[3] = ob[3];
        t[4] = cmd[0]; t[5] = cmd[1]; t[6] = cmd[3]; t[7] = cmd[4]; t[8] = cmd[5]; t[9] = cmd[6];
      }
    }
    if (OB.critic_obs) {
      float* c = OB.critic_obs + (size_t)env * ZB_NCRITIC_OBS;
      if (hinge) { c[dof - 6] = q; c[NU + dof - 6] = qv / 10.f; }
      if (isbody && L >= 1) {
#pragma unroll
        for (int k = 0; k < 10; k++) c[2 * NU + (L - 1) * 10 + k] = cinert[k];
#pragma unroll
        for (int k = 0; k < 6; k++) c[2 * NU + 250 + (L - 1) * 6 + k] = cvel[k];
      }
      if (L == 0) {  // imu_acc(3) | imu_gyro(3) | imu_orientation(4) | cmd(7)
        float* t = c + 2 * NU + 400;
        t[0] = acc[0]; t[1] = acc[1]; t[2] = acc[2]; t[3] = sq[7]; t[4] = sq[8]; t[5] = sq[9];
        t[6] = ob[0]; t[7] = ob[1]; t[8] = ob[2]; t[9] = ob[3];
#pragma unroll
        for (int k = 0; k < ZB_NCOMMAND; k++) t[10 + k] = cmd[k];
      }
      if (act >= 0) c[2 * NU + 417 + act] = act_force / 100.f;
      if (L == 1) {  // base_pos(3) | base_quat(4) = qpos[0:7]
        float* t = c + 2 * NU + 417 + NU;
#pragma unroll
        for (int k = 0; k < 7; k++) t[k] = bq[k];
      }
    }
    // ---- reward inputs (ZbObs.feet_pos_obs / reward_terms): the reference's own per-step reward / observation code ----
    if (OB.feet_pos_obs || OB.reward_terms) {
      const ZbModel& m = D->m;
      const int ll = m.site_body[1], lr = m.site_body[2];  // bodies of the left_foot / right_foot sites (train.py:1519-1520)
      float fl[3], fr[3], ql[4], qr[4];
#pragma unroll
      for (int k = 0; k < 3; k++) { fl[k] = shf(sxpos[k], ll); fr[k] = shf(sxpos[k], lr); }
#pragma unroll
      for (int k = 0; k < 4; k++) { ql[k] = shf(xquat[k], ll); qr[k] = shf(xquat[k], lr); }
      const float zb1 = shf(xpos[2], 1);  // xpos[1, 2]: height of the base body as forward() of the last substep left it
      // FeetPositionObservation train.py:452-465: site positions rotated by the inverse of the base quaternion qpos[3:7]
      // (xax.rotate_vector_by_quat(v, q, inverse=True) [R]: v' = R(q)^T v with the quaternion normalised first)
      float qn[4] = {bq[3], bq[4], bq[5], bq[6]};
      {
        const float nq = sqrtf(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]) + 1e-6f;
#pragma unroll
        for (int k = 0; k < 4; k++) qn[k] = qn[k] / nq;
      }
      float Rq[9], pl[3], pr[3];
      quat_to_mat(Rq, qn);
      matT_vec(pl, Rq, fl);
      matT_vec(pr, Rq, fr);
      if (OB.feet_pos_obs && L == 0) {
        float* t = OB.feet_pos_obs + (size_t)env * 6;
        t[0] = pl[0]; t[1] = pl[1]; t[2] = pl[2]; t[3] = pr[0]; t[4] = pr[1]; t[5] = pr[2];
      }
      if (OB.reward_terms && L == 0) {
        const float tl = keep[10], tr = keep[11];
        const bool cl = tl > 0.1f, cr = tr > 0.1f;  // train.py:699-703
        // FeetOrientationReward train.py:680-689: |roll| + |pitch| of both foot bodies (xax.quat_to_euler [R])
        float err = 0.f;
#pragma unroll
        for (int f = 0; f < 2; f++) {
          const float* qq = f ? qr : ql;
          const float nq = sqrtf(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]) + 1e-6f;
          const float w = qq[0] / nq, x = qq[1] / nq, y = qq[2] / nq, z = qq[3] / nq;
          const float roll = atan2f(2.f * (w * x + y * z), 1.f - 2.f * (x * x + y * y));
          const float sp = 2.f * (w * y - z * x);
          const float pitch = fabsf(sp) >= 1.f ? copysignf(1.57079632679489662f, sp) : asinf(sp);
          err += fabsf(roll) + fabsf(pitch);
        }
        const float dx = pl[0] - pr[0], dy = pl[1] - pr[1], dz = pl[2] - pr[2];
        float* t = OB.reward_terms + (size_t)env * ZB_NREWARD_TERMS;
        t[0] = (cl != cr) ? 1.f : 0.f;
        t[1] = expf(-err / 0.25f);
        t[2] = sqrtf(dx * dx + dy * dy + dz * dz) < 0.12f ? 1.f : 0.f;                 // train.py:741-748
        t[3] = expf(-fabsf(zb1 - (cmd[4] + 0.27f)) / 0.25f);                              // train.py:476-483
        t[4] = cl ? 1.f : 0.f;
        t[5] = cr ? 1.f : 0.f;
      }
    }
  }
};
