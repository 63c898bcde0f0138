// zb_policy_tc.cuh — the actor's GRU stack on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32 so that the result keeps fp32
// accuracy (the fp32 CUDA-core kernel of zb_policy.cuh is the reference implementation and the checker).
//
// Per CTA: 32 environments (the MMA N dimension), the 128 hidden units of a gate are the M dimension:
//     D_gate[128 f x 32 env] (+)= W_gate[128 f x 8 k] . X^T[8 k x 32 env]            tcgen05.mma.cta_group::1.kind::tf32, M128 N32 K8
// Four fp32 accumulators live in TMEM: r (W_ir x + W_hr h), z, i_n (W_in x), h_n (W_hn h).  Every product is formed as
// hi.hi + hi.lo + lo.hi with v_hi = v & 0xffffe000 (exactly representable in TF32), v_lo = v - v_hi — which recovers ~21 of the 24
// mantissa bits of each factor (the tensor core accumulates in fp32).  The activation operand holds [hi rows | lo rows], so ONE N = 64
// MMA gives A_hi.B_hi and A_hi.B_lo (in two column halves of the accumulator, summed in the epilogue: the 4 KB weight tile is read
// once for both) and one N = 32 MMA adds A_lo.B_hi.  Per layer: 6 gate matrices x 16 K-steps x 2 = 192 MMAs, 256 TMEM columns,
// issued by ONE thread; the other 511 threads feed it:
//   * weights: host-packed in the UMMA canonical K-major no-swizzle layout (8-row x 16-byte core matrices; ZbActor.gru_w_tc
//     [layer][k-tile 16][matrix 6][k-chunk 2][f 128][4]) so that a plain TMA bulk copy of 24 KB lands a ready operand tile in shared
//     memory; a 3-stage ring (full = transaction mbarrier, free again = tcgen05.commit on an mbarrier); the threads split each
//     landed tile into hi (in place) and lo (second half of the stage);
//   * activations: the layer input and the carry as hi / lo pairs in the same canonical layout ([k-chunk 32][env 32][4]), written by
//     the epilogue threads of the previous layer / loaded from the carry;
//   * epilogue: tcgen05.ld (32 lanes x 8 columns per warp) hands thread (f, eg) the four gate sums of hidden unit f for its 8
//     environments — the same thread mapping as the CUDA-core kernel — then sigmoid / tanh / blend in fp32, carry out, next operand.
// input_proj (50 -> 128) and output_proj (128 -> 300) are 8 % of the FLOPs and stay on the CUDA cores (fp32), fed by the same ring.
#pragma once
#include "zb_policy.cuh"

#define ZBT_NST 4
#define ZBT_TILE_FLOATS (6 * 2 * ZBP_H * 4) /* 24 KB: [matrix 6][k-chunk 2][f 128][4 k] */
#define ZBT_B_CHUNK (2 * ZBP_TE * 4 + 4)     /* floats between k-chunks of an activation operand: 64 rows x 4 k + 16 bytes of padding, so that the 32
                                               hidden units of a warp (8 chunks x 4 k) write to 32 different banks (the chunk stride is the descriptor's LBO) */
#define ZBT_B_FLOATS (32 * ZBT_B_CHUNK)      /* [k-chunk 32][row 64 = hi of env 0..31 | lo of env 0..31][4 k] */
#define ZBT_MAX_IN 96
#define ZBT_TMEM_COLS 512 /* 4 accumulators x 64 columns + 2 x 48 columns of weight lo parts (A operand of the lo.hi MMAs) */
#define ZBT_TMEM_LO 256  /* first column of the lo parts: [slot 2][matrix 6][8 k] */
#define ZBT_CONV_WARPS 12 /* warps 1..12 split the weight tiles (6144 floats = 12 x 32 threads x 4 float4); warp 0 issues the MMAs */
static_assert(ZBT_TILE_FLOATS % (4 * ZBT_CONV_WARPS * 32) == 0, "converter split");

struct ZbtSmem {
  float w[ZBT_NST][ZBT_TILE_FLOATS];     // weight ring: TMA destination, masked to its TF32 hi part in place; the end of the kernel re-uses it for `out`
  float wlo[2][ZBT_TILE_FLOATS];         // input_proj: its weights land here (the lo parts of the GRU weight tiles live in TMEM)
  __align__(16) float bx[ZBT_B_FLOATS];  // layer input: hi rows | lo rows (one N = 64 operand, or N = 32 = the hi rows alone)
  __align__(16) float bh[ZBT_B_FLOATS];  // carry of the layer, same layout (input_proj: observation scratch; output_proj: its input, k-major)
  uint64_t full[ZBT_NST], conv[ZBT_NST], mma_done[ZBT_NST], layer_done, in_full;
  uint32_t tmem_base;
};
static_assert(sizeof(ZbtSmem) <= 232448, "ZbtSmem exceeds the 227 KB of shared memory a CTA can have");
static_assert(ZBP_TE * (ZBP_OUT_MAX + 1) <= ZBT_NST * ZBT_TILE_FLOATS, "`out` must fit the weight ring it aliases");
static_assert(ZBP_H * ZBP_LD <= ZBT_B_FLOATS && ZBT_MAX_IN * ZBP_H <= 2 * ZBT_TILE_FLOATS && ZBT_MAX_IN * ZBP_LD <= ZBT_B_FLOATS, "aliased regions");

static __device__ __forceinline__ uint64_t zbt_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: start address, leading (K-chunk) and stride (8-row group) byte offsets in
  // 16-byte units, version 1 (Blackwell)
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
static __device__ __forceinline__ void zbt_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
static __device__ __forceinline__ void zbt_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t accumulate) {  // A from TMEM
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
static __device__ __forceinline__ void zbt_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
               "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
static __device__ __forceinline__ void zbt_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(zbp_saddr(bar)) : "memory");
}
static __device__ __forceinline__ void zbp_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(zbp_saddr(b)) : "memory");
}
static __device__ __forceinline__ bool zbt_elect() {  // one lane of a converged warp
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(p));
  return p != 0;
}
static __device__ __forceinline__ void zbt_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
static __device__ __forceinline__ int zbt_bidx(int k, int env, int lo) { return (k >> 2) * ZBT_B_CHUNK + (lo * ZBP_TE + env) * 4 + (k & 3); }
static __device__ __forceinline__ float zbt_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// tiles of the ring in stream order: T_gru GRU tiles ([layer][k-tile] of gru_w_tc, 24 KB each), then output_proj rows (16 k-rows each)
struct ZbtStream {
  int t_gru, t_out, n_out;
  const float *gru, *out_wt;
  __device__ void issue(int g, ZbtSmem& S) const {  // one thread
    const int st = g % ZBT_NST;
    uint64_t* b = &S.full[st];
    if (g < t_gru) {
      zbp_mbar_expect(b, ZBT_TILE_FLOATS * 4);
      zbp_bulk_load(S.w[st], gru + (size_t)g * ZBT_TILE_FLOATS, ZBT_TILE_FLOATS * 4, b);
    } else {  // output_proj as a sixth layer: 3 M-tiles of 128 output rows (ZbActor.out_w_tc [k-tile 16][m-tile 3][k-chunk 2][row 128][4], 12 KB per k-tile)
      zbp_mbar_expect(b, ZBT_TILE_FLOATS * 2);
      zbp_bulk_load(S.w[st], out_wt + (size_t)(g - t_gru) * (ZBT_TILE_FLOATS / 2), ZBT_TILE_FLOATS * 2, b);
    }
  }
};

__global__ void __launch_bounds__(ZBP_THREADS, 1)
zb_actor_tc_kernel(ZbActor A, const float* __restrict__ obs, const float* carry_in, float* carry_out, const uint8_t* __restrict__ carry_reset,
                   const float* __restrict__ noise_u, const float* __restrict__ noise_eps, float* __restrict__ action, float* __restrict__ dist, int n_envs) {
#ifdef ZBT_PROFILE
  long long pf[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long pt = clock64(), pstart = pt;
#define ZBT_TICK(i) do { const long long n_ = clock64(); pf[i] += n_ - pt; pt = n_; } while (0)
#else
#define ZBT_TICK(i)
#endif
  extern __shared__ __align__(128) unsigned char zbt_raw[];
  ZbtSmem& S = *reinterpret_cast<ZbtSmem*>(zbt_raw);
  const int tid = threadIdx.x, warp = tid >> 5, f = tid & (ZBP_H - 1), eg = tid >> 7, e0 = eg * ZBP_EPT;
  const int env0 = blockIdx.x * ZBP_TE;
  const int n_in = A.n_in, n_out = ZB_NU * 3 * A.num_mixtures;
  ZbtStream W;
  W.t_gru = A.depth * (ZBP_H / 8); W.t_out = ZBP_H / 8; W.n_out = n_out; W.gru = A.gru_w_tc; W.out_wt = A.out_w_tc;
  const int G = W.t_gru + W.t_out;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 32, M = 128
  const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ZBP_TE >> 3) << 17) | ((uint32_t)(ZBP_H >> 4) << 24);
  const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * ZBP_TE) >> 3) << 17) | ((uint32_t)(ZBP_H >> 4) << 24);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < ZBT_NST; s++) { zbp_mbar_init(&S.full[s], 1); zbp_mbar_init(&S.mma_done[s], 1); zbp_mbar_init(&S.conv[s], ZBT_CONV_WARPS); }
    zbp_mbar_init(&S.layer_done, 1);
    zbp_mbar_init(&S.in_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(zbp_saddr(&S.tmem_base)), "r"((uint32_t)ZBT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;
  ZBT_TICK(10);  // mbarrier init + TMEM alloc + barrier
  const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);      // (provably warp-uniform copies for the issuing warp)
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)n_in * ZBP_H * 4;  // input_proj weights, whole, into the two lo buffers (48 KB, contiguous)
    zbp_mbar_expect(&S.in_full, bytes);
    zbp_bulk_load(S.wlo[0], A.in_wt, bytes, &S.in_full);
    for (int g0 = 0; g0 < ZBT_NST - 1; g0++) W.issue(g0, S);
  }
  // ---- observations of the tile, transposed ([k][e], CUDA-core layout) into the bh region ----
  float (*obsT)[ZBP_LD] = reinterpret_cast<float (*)[ZBP_LD]>(&S.bh[0]);
  for (int i = tid; i < ZBP_TE * n_in; i += ZBP_THREADS) {
    const int e = i / n_in, k = i - e * n_in;
    obsT[k][e] = (env0 + e < n_envs) ? obs[(size_t)(env0 + e) * n_in + k] : 0.f;
  }
  __syncthreads();
  ZBT_TICK(11);  // observations -> shared memory
  // ---- input_proj (CUDA cores) -> layer 0's input operand (hi / lo) ----
  {
    float acc[ZBP_EPT];
    const float b = A.in_b[f];
#pragma unroll
    for (int e = 0; e < ZBP_EPT; e++) acc[e] = b;
    zbp_mbar_wait(&S.in_full, 0);
    ZBT_TICK(12);  // wait: input_proj weights landed
    const float* w = S.wlo[0] + f;
    for (int k = 0; k < n_in; k++) {
      const float wv = w[k * ZBP_H];
      const float4 a = *reinterpret_cast<const float4*>(&obsT[k][e0]), c = *reinterpret_cast<const float4*>(&obsT[k][e0 + 4]);
      acc[0] = fmaf(wv, a.x, acc[0]); acc[1] = fmaf(wv, a.y, acc[1]); acc[2] = fmaf(wv, a.z, acc[2]); acc[3] = fmaf(wv, a.w, acc[3]);
      acc[4] = fmaf(wv, c.x, acc[4]); acc[5] = fmaf(wv, c.y, acc[5]); acc[6] = fmaf(wv, c.z, acc[6]); acc[7] = fmaf(wv, c.w, acc[7]);
    }
#pragma unroll
    for (int e = 0; e < ZBP_EPT; e++) {
      const float h = zbt_hi(acc[e]);
      S.bx[zbt_bidx(f, e0 + e, 0)] = h;
      S.bx[zbt_bidx(f, e0 + e, 1)] = acc[e] - h;
    }
  }
  __syncthreads();  // the lo buffers and the observation scratch are free

  ZBT_TICK(13);  // input_proj
  // ---- GRU stack on the tensor cores ----
  float hold[ZBP_EPT], hnext[ZBP_EPT], br = 0.f, bz = 0.f, bn = 0.f, bhn = 0.f;
  bool live[ZBP_EPT];
#pragma unroll
  for (int e = 0; e < ZBP_EPT; e++) {  // layer 0's carry (a reset carry is the initial carry: zeros, train.py:1731-1735)
    const int env = env0 + e0 + e;
    live[e] = env < n_envs && !(carry_reset && carry_reset[env]);
    hnext[e] = live[e] ? carry_in[((size_t)env * A.depth) * ZBP_H + f] : 0.f;
  }
  for (int g = 0; g < G; g++) {
    const int l = g >> 4, t = g & 15, st = g % ZBT_NST;
    const bool out_layer = l == A.depth;  // (warp-uniform) output_proj: 3 M-tiles of W_out against the last layer's output, no carry operand
    if (t == 0) {
      if (l == 0) {  // layer 0: its carry operand is written here (later layers: by the previous layer's epilogue, one barrier less per layer)
#pragma unroll
        for (int e = 0; e < ZBP_EPT; e++) {
          const float h = zbt_hi(hnext[e]);
          S.bh[zbt_bidx(f, e0 + e, 0)] = h;
          S.bh[zbt_bidx(f, e0 + e, 1)] = hnext[e] - h;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // bx (input_proj), bh -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
      }
      if (!out_layer) {  // this layer's carry (gate blend), and the loads that are needed 16 tiles from now: its biases and the next layer's carry
#pragma unroll
        for (int e = 0; e < ZBP_EPT; e++) hold[e] = hnext[e];
        const float* bl = A.gru_b + (size_t)l * 3 * ZBP_H;
        br = bl[f]; bz = bl[ZBP_H + f]; bn = bl[2 * ZBP_H + f]; bhn = A.gru_bn[(size_t)l * ZBP_H + f];
        if (l + 1 < A.depth) {
#pragma unroll
          for (int e = 0; e < ZBP_EPT; e++) hnext[e] = live[e] ? carry_in[((size_t)(env0 + e0 + e) * A.depth + l + 1) * ZBP_H + f] : 0.f;
        }
      }
    }
    ZBT_TICK(1);  // layer prologue / loop overhead
    if (warp_u == 0) {
      // ===== issuing warp: waits for the tile's hi / lo operands, issues its 12 MMAs (one elected lane; descriptors are formed
      // warp-uniformly, in uniform registers), frees stages, keeps the TMA stream going =====
      zbp_mbar_wait(&S.conv[st], (uint32_t)((g / ZBT_NST) & 1));
      ZBT_TICK(2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = zbp_saddr(S.w[st]);
      const uint32_t koff = (uint32_t)(2 * t) * (ZBT_B_CHUNK * 4);  // two 16-byte k-chunks per K = 8 step
      const uint64_t dBx = zbt_desc(zbp_saddr(S.bx) + koff, ZBT_B_CHUNK * 4, 128), dBh = zbt_desc(zbp_saddr(S.bh) + koff, ZBT_B_CHUNK * 4, 128);
      const uint64_t dA_hi0 = zbt_desc(a_hi, ZBP_H * 16, 128);
      const uint32_t acc0 = t == 0 ? 0u : 1u;
      const bool lead = zbt_elect();
      if (lead) {
        if (!out_layer) {
#pragma unroll
          for (int m = 0; m < 6; m++) {
            const bool is_h = m >= 3;
            const int gate = m % 3;
            const uint32_t col = (gate == 2 ? (is_h ? 3u : 2u) : (uint32_t)gate) * 64u;  // accumulator: columns [col, col+32) += hi.hi and lo.hi, [col+32, col+64) += hi.lo
            const uint64_t dstep = (uint64_t)(m * ((2 * ZBP_H * 16) >> 4));         // next matrix of the tile: + 4 KB in the address field
            const uint32_t first = (!is_h || gate == 2) ? acc0 : 1u;                   // r, z: W_h h accumulates onto W_i x
#ifndef ZBT_SKIP_MMA
            zbt_mma(tmem_u + col, dA_hi0 + dstep, is_h ? dBh : dBx, idesc64, first);  // A_hi . [B_hi | B_lo]  (N = 64: the weight tile is read once for both)
            zbt_mma_ts(tmem_u + col, tmem_u + ZBT_TMEM_LO + (uint32_t)((g & 1) * 48 + m * 8), is_h ? dBh : dBx, idesc32, 1u);  // A_lo (TMEM) . B_hi  (N = 32: the hi rows)
#endif
          }
        } else {
#pragma unroll
          for (int m = 0; m < 3; m++) {  // output rows 128 m .. 128 m + 127
            const uint64_t dstep = (uint64_t)(m * ((2 * ZBP_H * 16) >> 4));
            zbt_mma(tmem_u + (uint32_t)m * 64u, dA_hi0 + dstep, dBx, idesc64, acc0);
            zbt_mma_ts(tmem_u + (uint32_t)m * 64u, tmem_u + ZBT_TMEM_LO + (uint32_t)((g & 1) * 48 + m * 8), dBx, idesc32, 1u);
          }
        }
        zbt_commit(&S.mma_done[st]);
        if (t == 15) zbt_commit(&S.layer_done);  // (the output_proj layer is completion number depth + 1)
        if (g + ZBT_NST - 1 < G) {  // tile g + NST - 1 goes into the stage tile g - 1 used: free once that tile's MMAs have completed
          if (g >= 1) zbp_mbar_wait(&S.mma_done[(g - 1) % ZBT_NST], (uint32_t)(((g - 1) / ZBT_NST) & 1));
          W.issue(g + ZBT_NST - 1, S);
        }
      }
      __syncwarp();
    } else if (warp_u <= ZBT_CONV_WARPS) {
      // ===== converter warps: split the landed tile into its TF32 hi part (in place) and the remainder (the tile's lo buffer, last read by
      // the MMAs of tile g - 2), one tile ahead of the MMAs =====
      zbp_mbar_wait(&S.full[st], (uint32_t)((g / ZBT_NST) & 1));
      ZBT_TICK(2);  // wait: tile landed
      if (g >= 2) zbp_mbar_wait(&S.mma_done[(g - 2) % ZBT_NST], (uint32_t)(((g - 2) / ZBT_NST) & 1));
      ZBT_TICK(3);  // wait: MMAs of tile g - 2
#ifndef ZBT_SKIP_CONVERT
      {  // thread = (row f of its warp's TMEM lane quadrant, two of the six matrices): 8 k values per matrix -> lo parts -> 8 TMEM columns
        // converter warps 1..12: quadrant = warp & 3 (a warp can only store to the TMEM lanes of its own quadrant); the three warps of a
        // quadrant (1-4, 5-8, 9-12) take matrices {0,1}, {2,3}, {4,5}
        const int q = warp & 3, fr = 32 * q + (tid & 31), grp = (warp - 1) >> 2;
#pragma unroll
        for (int mm = 0; mm < 2; mm++) {
          const int m = out_layer ? grp : 2 * grp + mm;
          if (out_layer && mm == 1) break;
          const float4 v0 = *reinterpret_cast<const float4*>(S.w[st] + (m * 2 + 0) * (ZBP_H * 4) + fr * 4);
          const float4 v1 = *reinterpret_cast<const float4*>(S.w[st] + (m * 2 + 1) * (ZBP_H * 4) + fr * 4);
          const float lo[8] = {v0.x - zbt_hi(v0.x), v0.y - zbt_hi(v0.y), v0.z - zbt_hi(v0.z), v0.w - zbt_hi(v0.w),
                               v1.x - zbt_hi(v1.x), v1.y - zbt_hi(v1.y), v1.z - zbt_hi(v1.z), v1.w - zbt_hi(v1.w)};
          zbt_st8(tmem + ((uint32_t)(32 * q) << 16) + ZBT_TMEM_LO + (uint32_t)((g & 1) * 48 + m * 8), lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
#endif
      ZBT_TICK(4);  // convert
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes (operands) -> visible to the tensor core's reads
      __syncwarp();
      if ((tid & 31) == 0) zbp_mbar_arrive(&S.conv[st]);
      ZBT_TICK(5);
    }
    ZBT_TICK(6);  // thread 0: MMA issue + commit + refill
    if (t == 15 && !out_layer) {  // layer epilogue
      zbp_mbar_wait(&S.layer_done, (uint32_t)(l & 1));
      ZBT_TICK(7);  // wait: layer's MMAs done
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)e0;
      float ar[8], az[8], ain[8], ahn[8];
      {
        float t2[8];
        zbt_ld8(taddr, ar); zbt_ld8(taddr + 32, t2);
#pragma unroll
        for (int e = 0; e < 8; e++) ar[e] += t2[e];
        zbt_ld8(taddr + 64, az); zbt_ld8(taddr + 96, t2);
#pragma unroll
        for (int e = 0; e < 8; e++) az[e] += t2[e];
        zbt_ld8(taddr + 128, ain); zbt_ld8(taddr + 160, t2);
#pragma unroll
        for (int e = 0; e < 8; e++) ain[e] += t2[e];
        zbt_ld8(taddr + 192, ahn); zbt_ld8(taddr + 224, t2);
#pragma unroll
        for (int e = 0; e < 8; e++) ahn[e] += t2[e];
      }
#pragma unroll
      for (int e = 0; e < ZBP_EPT; e++) {
        const float r = zbp_sigmoid(ar[e] + br), z = zbp_sigmoid(az[e] + bz);
        const float nn = tanhf(fmaf(r, ahn[e] + bhn, ain[e] + bn));
        const float hn = fmaf(z, hold[e] - nn, nn);
        const int env = env0 + e0 + e;
        if (env < n_envs) carry_out[((size_t)env * A.depth + l) * ZBP_H + f] = hn;
        const float h = zbt_hi(hn);  // the next layer's (or output_proj's) input operand
        S.bx[zbt_bidx(f, e0 + e, 0)] = h;
        S.bx[zbt_bidx(f, e0 + e, 1)] = hn - h;
        if (l + 1 < A.depth) {       // ... and the next layer's carry operand (bh: this layer's MMAs, its last readers, have completed)
          const float c = zbt_hi(hnext[e]);
          S.bh[zbt_bidx(f, e0 + e, 0)] = c;
          S.bh[zbt_bidx(f, e0 + e, 1)] = hnext[e] - c;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the operands just written -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();  // TMEM read out, the next layer's operands complete
      ZBT_TICK(8);  // epilogue
    }
  }
  // ---- output_proj epilogue: accumulators (3 M-tiles x [hi.hi + lo.hi | hi.lo]) + bias -> `out` of the tile (aliases the drained ring) ----
  float (*outS)[ZBP_OUT_MAX + 1] = reinterpret_cast<float (*)[ZBP_OUT_MAX + 1]>(&S.w[0][0]);
  {
    zbp_mbar_wait(&S.layer_done, (uint32_t)(A.depth & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)e0;
#pragma unroll
    for (int m = 0; m < 3; m++) {
      const int j = f + m * ZBP_H;
      float a[8], b2[8];
      zbt_ld8(taddr + m * 64, a); zbt_ld8(taddr + m * 64 + 32, b2);
      if (j < n_out) {
        const float bj = A.out_b[j];
#pragma unroll
        for (int e = 0; e < ZBP_EPT; e++) outS[e0 + e][j] = a[e] + b2[e] + bj;
      }
    }
  }
  __syncthreads();
  zbp_head(A, outS, env0, n_envs, noise_u, noise_eps, action, dist);
  ZBT_TICK(9);  // output_proj + head
#ifdef ZBT_PROFILE
  __syncthreads();
  if (blockIdx.x == 0 && (tid == 0 || tid == 200) && dist) {
    float* o = dist + (tid == 0 ? 0 : 16);
    for (int i = 0; i < 14; i++) o[i] = (float)pf[i];
    o[14] = (float)(clock64() - pstart);
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)ZBT_TMEM_COLS) : "memory");
}
