// zb_policy.cuh — the reference's actor network as one kernel per control step ("policy inference
// (GRU-128x5) fused into the rollout loop").
//
// What it replaces: Actor.forward (/root/reference/train.py:946-968) as called by run_actor / sample_action
// (train.py:1624-1648, 1737-1763) for every environment on every control step of the rollout:
//     x      = input_proj(obs)                                   eqx.nn.Linear(50, 128)
//     x, h_i = GRUCell_i(x, carry[i])            i = 0..depth-1   eqx.nn.GRUCell(128, 128) [R: equinox, un-vendored]
//     out    = output_proj(x)                                    eqx.nn.Linear(128, 20 * 3 * num_mixtures)
//     mean   = out[:20M].reshape(20, M) + JOINT_BIASES[:, None]
//     std    = clip((softplus(out[20M:40M]) + min_std) * var_scale, max = max_std)
//     logits = out[40M:]
//     action = MixtureOfGaussians(mean, std, logits).mode() | .sample()          (ksim [R], see PARITY_RISKS.md)
// GRUCell (equinox [R]; the PyTorch gate convention):
//     ig = W_ih x + bias;  hg = W_hh h;  r = sigmoid(ig_r + hg_r);  z = sigmoid(ig_z + hg_z)
//     n  = tanh(ig_n + r * (hg_n + bias_n));  h' = n + z * (h - n)
//
// Mapping: one CTA = 32 environments, 512 threads; thread (f = tid & 127, eg = tid >> 7) owns hidden unit f of the 8 environments
// 8 eg .. 8 eg + 7, so all six gate pre-activations of a unit meet in one thread and the gate math needs no exchange.  The layer input
// and the carry of the tile sit transposed in shared memory (k-major: one broadcast LDS.128 pair gives a thread its 8 environments'
// values of input k); weights are pre-packed k-major on the host (ZbActor), so a warp reads 32 consecutive floats per (k, gate).  The
// weights stream through a 4-stage shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier complete_tx, one thread
// issues, tiles of 8 k-rows of both gate matrices = 24 KB) so that their DRAM / L2 latency never meets the FMA loop.
// Per k: 6 LDS (weights, conflict-free) + 4 LDS.128 (broadcast) feed 48 FFMAs.
// fp32 throughout (accumulation in k order) — a real dense contraction (1.07 MFLOP per environment per control step, 2x the
// physics FLOPs of a 4-substep control step), kept on the CUDA cores so that it matches the reference's CPU float32 numbers to
// rounding; a TF32 tensor-core path would sit 1e-3 away from them.
#pragma once
#include <stdint.h>

#include "../../include/zbphys.h"

#define ZBP_TE 32        /* environments per CTA */
#define ZBP_EPT 8        /* environments per thread */
#define ZBP_H ZB_ACTOR_HIDDEN
#define ZBP_THREADS (ZBP_H * (ZBP_TE / ZBP_EPT))
#define ZBP_LD (ZBP_TE + 4) /* padded row: a quarter-warp's STS.128 at consecutive f spread over all banks */
#define ZBP_OUT_MAX (ZB_NU * 3 * ZB_ACTOR_MAX_MIXTURES)
#define ZBP_NST 4                               /* weight ring: stages */
#define ZBP_KT 8                                /* GRU tile: k rows of both gate matrices per stage */
#define ZBP_STAGE_FLOATS (2 * ZBP_KT * 3 * ZBP_H) /* 24 KB */
#define ZBP_KT_IN 32                            /* input_proj tile: k rows of [128] */
#define ZBP_KT_OUT 16                           /* output_proj tile: k rows of [n_out <= 300] */

static __device__ __forceinline__ float zbp_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
static __device__ __forceinline__ float zbp_softplus(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

// ---- TMA bulk copy + mbarrier (the weight stream: global -> shared memory ring, completion counted in bytes on an mbarrier) ----
static __device__ __forceinline__ uint32_t zbp_saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void zbp_mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(zbp_saddr(b)), "r"(count) : "memory");
}
static __device__ __forceinline__ void zbp_mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(zbp_saddr(b)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void zbp_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(zbp_saddr(dst)), "l"(src), "r"(bytes),
               "r"(zbp_saddr(b))
               : "memory");
}
static __device__ __forceinline__ void zbp_mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "ZBP_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra ZBP_DONE_%=;\n"
      "bra ZBP_WAIT_%=;\n"
      "ZBP_DONE_%=:\n"
      "}" ::"r"(zbp_saddr(b)),
      "r"(parity)
      : "memory");
}

struct ZbpSmem {
  float w[ZBP_NST][ZBP_STAGE_FLOATS];   // weight ring (TMA destination, 16-byte aligned)
  float xT[ZBP_H][ZBP_LD];              // layer input, k-major
  float hT[ZBP_H][ZBP_LD];              // carry of the current layer, k-major
  float out[ZBP_TE][ZBP_OUT_MAX + 1];   // output_proj result per environment (row-major, odd stride)
  uint64_t full[ZBP_NST];               // "stage s holds its tile" (transaction barrier)
};

// Weight tiles in stream order: input_proj rows, then per GRU layer 16 tiles of 8 k-rows of (W_ih, W_hh), then output_proj rows.
struct ZbpStream {
  int t_in, t_gru, t_out, n_in, n_out;
  const float *in_wt, *wih, *whh, *out_wt;
  __device__ __forceinline__ int total() const { return t_in + t_gru + t_out; }
  __device__ void issue(int g, ZbpSmem& S) const {  // one thread
    const int st = g % ZBP_NST;
    float* dst = S.w[st];
    uint64_t* b = &S.full[st];
    if (g < t_in) {
      const int k0 = g * ZBP_KT_IN, rows = min(ZBP_KT_IN, n_in - k0);
      const uint32_t bytes = (uint32_t)rows * ZBP_H * 4;
      zbp_mbar_expect(b, bytes);
      zbp_bulk_load(dst, in_wt + (size_t)k0 * ZBP_H, bytes, b);
    } else if (g < t_in + t_gru) {
      const size_t off = (size_t)(g - t_in) * (ZBP_KT * 3 * ZBP_H);
      const uint32_t bytes = ZBP_KT * 3 * ZBP_H * 4;
      zbp_mbar_expect(b, 2 * bytes);
      zbp_bulk_load(dst, wih + off, bytes, b);
      zbp_bulk_load(dst + ZBP_KT * 3 * ZBP_H, whh + off, bytes, b);
    } else {
      const int k0 = (g - t_in - t_gru) * ZBP_KT_OUT;
      const uint32_t bytes = (uint32_t)ZBP_KT_OUT * n_out * 4;
      zbp_mbar_expect(b, bytes);
      zbp_bulk_load(dst, out_wt + (size_t)k0 * n_out, bytes, b);
    }
  }
};

// mixture-of-Gaussians head: one (environment, joint) pair per thread; `out` = output_proj result of the tile [env][n_out]
static __device__ __forceinline__ void zbp_head(const ZbActor& A, float (*outp)[ZBP_OUT_MAX + 1], int env0, int n_envs, const float* __restrict__ noise_u,
                                                const float* __restrict__ noise_eps, float* __restrict__ action, float* __restrict__ dist) {
  const int tid = threadIdx.x;
  const int M = A.num_mixtures, sl = ZB_NU * M;
  for (int i = tid; i < ZBP_TE * ZB_NU; i += ZBP_THREADS) {
    const int e = i / ZB_NU, j = i - e * ZB_NU, env = env0 + e;
    if (env >= n_envs) continue;
    const float* o = outp[e];
    const float jb = A.mean_bias ? A.mean_bias[j] : 0.f;
    float lmax = -INFINITY;
    int kmax = 0;
    for (int m = 0; m < M; m++) {
      const float lg = o[2 * sl + j * M + m];
      if (lg > lmax) { lmax = lg; kmax = m; }  // first maximum (jnp.argmax)
    }
    int kk = kmax;
    if (noise_eps && noise_u) {  // categorical by inverse CDF of softmax(logits) at u in [0, 1)
      float tot = 0.f;
      for (int m = 0; m < M; m++) tot += expf(o[2 * sl + j * M + m] - lmax);
      const float u = noise_u[(size_t)env * ZB_NU + j] * tot;
      float cum = 0.f;
      kk = M - 1;
      for (int m = 0; m < M; m++) {
        cum += expf(o[2 * sl + j * M + m] - lmax);
        if (u < cum) { kk = m; break; }
      }
    }
    const float mean = o[j * M + kk] + jb;
    const float sd = fminf((zbp_softplus(o[sl + j * M + kk]) + A.min_std) * A.var_scale, A.max_std);
    action[(size_t)env * ZB_NU + j] = noise_eps ? fmaf(sd, noise_eps[(size_t)env * ZB_NU + j], mean) : mean;
    if (dist) {  // the distribution parameters as the reference hands them to MixtureOfGaussians: [N, 3, 20, M]
      for (int m = 0; m < M; m++) {
        float* d = dist + (size_t)env * 3 * sl;
        d[j * M + m] = o[j * M + m] + jb;
        d[sl + j * M + m] = fminf((zbp_softplus(o[sl + j * M + m]) + A.min_std) * A.var_scale, A.max_std);
        d[2 * sl + j * M + m] = o[2 * sl + j * M + m];
      }
    }
  }
}

__global__ void __launch_bounds__(ZBP_THREADS, 1)
zb_actor_kernel(ZbActor A, const float* __restrict__ obs, const float* carry_in, float* carry_out, const uint8_t* __restrict__ carry_reset,
                const float* __restrict__ noise_u, const float* __restrict__ noise_eps, float* __restrict__ action, float* __restrict__ dist, int n_envs) {
  extern __shared__ __align__(128) unsigned char zbp_raw[];
  ZbpSmem& S = *reinterpret_cast<ZbpSmem*>(zbp_raw);
  const int tid = threadIdx.x, f = tid & (ZBP_H - 1), eg = tid >> 7, e0 = eg * ZBP_EPT;
  const int env0 = blockIdx.x * ZBP_TE;
  const int n_in = A.n_in, n_out = ZB_NU * 3 * A.num_mixtures;
  ZbpStream W;
  W.n_in = n_in; W.n_out = n_out;
  W.t_in = (n_in + ZBP_KT_IN - 1) / ZBP_KT_IN; W.t_gru = A.depth * (ZBP_H / ZBP_KT); W.t_out = ZBP_H / ZBP_KT_OUT;
  W.in_wt = A.in_wt; W.wih = A.gru_wih_t; W.whh = A.gru_whh_t; W.out_wt = A.out_wt;
  const int G = W.total();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < ZBP_NST; s++) zbp_mbar_init(&S.full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int g = 0; g < ZBP_NST && g < G; g++) W.issue(g, S);
  }
  int g = 0;  // tile the CTA consumes next (every thread counts along)
  // a tile is consumed between acquire() and release(); release() refills the stage with tile g + NST once every thread has read it
#define ZBP_ACQUIRE() zbp_mbar_wait(&S.full[g % ZBP_NST], (uint32_t)((g / ZBP_NST) & 1))
#define ZBP_RELEASE() do { __syncthreads(); if (tid == 0 && g + ZBP_NST < G) W.issue(g + ZBP_NST, S); g++; } while (0)

  // ---- observations of the tile, transposed into hT (used as scratch: [k][e]) ----
  for (int i = tid; i < ZBP_TE * n_in; i += ZBP_THREADS) {
    const int e = i / n_in, k = i - e * n_in;
    S.hT[k][e] = (env0 + e < n_envs) ? obs[(size_t)(env0 + e) * n_in + k] : 0.f;
  }
  __syncthreads();  // (also publishes the mbarrier initialisation to every thread)
  // ---- input_proj ----
  {
    float acc[ZBP_EPT];
    const float b = A.in_b[f];
#pragma unroll
    for (int e = 0; e < ZBP_EPT; e++) acc[e] = b;
    for (int t = 0; t < W.t_in; t++) {
      ZBP_ACQUIRE();
      const float* w = S.w[g % ZBP_NST] + f;
      const int k0 = t * ZBP_KT_IN, rows = min(ZBP_KT_IN, n_in - k0);
      for (int kk = 0; kk < rows; kk++) {
        const float wv = w[kk * ZBP_H];
        const float4 a = *reinterpret_cast<const float4*>(&S.hT[k0 + kk][e0]), c = *reinterpret_cast<const float4*>(&S.hT[k0 + kk][e0 + 4]);
        acc[0] = fmaf(wv, a.x, acc[0]); acc[1] = fmaf(wv, a.y, acc[1]); acc[2] = fmaf(wv, a.z, acc[2]); acc[3] = fmaf(wv, a.w, acc[3]);
        acc[4] = fmaf(wv, c.x, acc[4]); acc[5] = fmaf(wv, c.y, acc[5]); acc[6] = fmaf(wv, c.z, acc[6]); acc[7] = fmaf(wv, c.w, acc[7]);
      }
      ZBP_RELEASE();
    }
    *reinterpret_cast<float4*>(&S.xT[f][e0]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(&S.xT[f][e0 + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  // ---- GRU stack ----
  for (int l = 0; l < A.depth; l++) {
    // (the release() of the previous tile was a CTA barrier: hT is free; xT gets its barrier below)
    float hold[ZBP_EPT];
#pragma unroll
    for (int e = 0; e < ZBP_EPT; e++) {
      const int env = env0 + e0 + e;
      const bool live = env < n_envs && !(carry_reset && carry_reset[env]);
      hold[e] = live ? carry_in[((size_t)env * A.depth + l) * ZBP_H + f] : 0.f;  // (a reset carry is the initial carry: zeros, train.py:1731-1735)
    }
    *reinterpret_cast<float4*>(&S.hT[f][e0]) = make_float4(hold[0], hold[1], hold[2], hold[3]);
    *reinterpret_cast<float4*>(&S.hT[f][e0 + 4]) = make_float4(hold[4], hold[5], hold[6], hold[7]);
    __syncthreads();  // xT (previous layer / input_proj) and hT complete
    float ar[ZBP_EPT], az[ZBP_EPT], ain[ZBP_EPT], ahn[ZBP_EPT];
    const float* bl = A.gru_b + (size_t)l * 3 * ZBP_H;
    const float br = bl[f], bz = bl[ZBP_H + f], bn = bl[2 * ZBP_H + f], bhn = A.gru_bn[(size_t)l * ZBP_H + f];
#pragma unroll
    for (int e = 0; e < ZBP_EPT; e++) { ar[e] = br; az[e] = bz; ain[e] = bn; ahn[e] = bhn; }
    for (int t = 0; t < ZBP_H / ZBP_KT; t++) {
      ZBP_ACQUIRE();
      const float* wi = S.w[g % ZBP_NST] + f;
      const float* wh = wi + ZBP_KT * 3 * ZBP_H;
#pragma unroll
      for (int kk = 0; kk < ZBP_KT; kk++) {
        const int k = t * ZBP_KT + kk;
        const float wir = wi[(3 * kk) * ZBP_H], wiz = wi[(3 * kk + 1) * ZBP_H], win = wi[(3 * kk + 2) * ZBP_H];
        const float whr = wh[(3 * kk) * ZBP_H], whz = wh[(3 * kk + 1) * ZBP_H], whn = wh[(3 * kk + 2) * ZBP_H];
        const float4 x0 = *reinterpret_cast<const float4*>(&S.xT[k][e0]), x1 = *reinterpret_cast<const float4*>(&S.xT[k][e0 + 4]);
        const float4 h0 = *reinterpret_cast<const float4*>(&S.hT[k][e0]), h1 = *reinterpret_cast<const float4*>(&S.hT[k][e0 + 4]);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
        for (int e = 0; e < ZBP_EPT; e++) {
          ar[e] = fmaf(wir, xv[e], ar[e]); az[e] = fmaf(wiz, xv[e], az[e]); ain[e] = fmaf(win, xv[e], ain[e]);
          ar[e] = fmaf(whr, hv[e], ar[e]); az[e] = fmaf(whz, hv[e], az[e]); ahn[e] = fmaf(whn, hv[e], ahn[e]);
        }
      }
      ZBP_RELEASE();
    }
    // (the last release() was a CTA barrier: every thread is done reading xT / hT of this layer)
    float hnew[ZBP_EPT];
#pragma unroll
    for (int e = 0; e < ZBP_EPT; e++) {
      const float r = zbp_sigmoid(ar[e]), z = zbp_sigmoid(az[e]);
      const float nn = tanhf(fmaf(r, ahn[e], ain[e]));
      hnew[e] = fmaf(z, hold[e] - nn, nn);
      const int env = env0 + e0 + e;
      if (env < n_envs) carry_out[((size_t)env * A.depth + l) * ZBP_H + f] = hnew[e];
    }
    *reinterpret_cast<float4*>(&S.xT[f][e0]) = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
    *reinterpret_cast<float4*>(&S.xT[f][e0 + 4]) = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
  }
  __syncthreads();
  // ---- output_proj: thread (f, eg) computes outputs f, f + 128, f + 256 of its 8 environments ----
  {
    float acc[3][ZBP_EPT];
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int j = f + m * ZBP_H;
      const float b = j < n_out ? A.out_b[j] : 0.f;
#pragma unroll
      for (int e = 0; e < ZBP_EPT; e++) acc[m][e] = b;
    }
    for (int t = 0; t < W.t_out; t++) {
      ZBP_ACQUIRE();
      const float* w = S.w[g % ZBP_NST];
#pragma unroll 4
      for (int kk = 0; kk < ZBP_KT_OUT; kk++) {
        const int k = t * ZBP_KT_OUT + kk;
        float wv[3];
#pragma unroll
        for (int m = 0; m < 3; m++) { const int j = f + m * ZBP_H; wv[m] = j < n_out ? w[kk * n_out + j] : 0.f; }
        const float4 x0 = *reinterpret_cast<const float4*>(&S.xT[k][e0]), x1 = *reinterpret_cast<const float4*>(&S.xT[k][e0 + 4]);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int m = 0; m < 3; m++)
#pragma unroll
          for (int e = 0; e < ZBP_EPT; e++) acc[m][e] = fmaf(wv[m], xv[e], acc[m][e]);
      }
      ZBP_RELEASE();
    }
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int j = f + m * ZBP_H;
      if (j < n_out)
#pragma unroll
        for (int e = 0; e < ZBP_EPT; e++) S.out[e0 + e][j] = acc[m][e];
    }
  }
  __syncthreads();
#undef ZBP_ACQUIRE
#undef ZBP_RELEASE
  zbp_head(A, S.out, env0, n_envs, noise_u, noise_eps, action, dist);
}
