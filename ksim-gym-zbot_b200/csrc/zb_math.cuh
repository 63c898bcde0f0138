// zb_math.cuh — register-resident small math + warp primitives for the one-warp-per-environment
// physics kernels.  Everything is __forceinline__ and operates on fixed-size arrays so that the
// compiler keeps all of it in registers (no local memory).
#pragma once
#include <cuda_runtime.h>

#define ZB_FULL 0xffffffffu
#define ZB_DI __device__ __forceinline__

ZB_DI float shf(float v, int src) { return __shfl_sync(ZB_FULL, v, src); }
ZB_DI int shfi(int v, int src) { return __shfl_sync(ZB_FULL, v, src); }
ZB_DI float shx(float v, int m) { return __shfl_xor_sync(ZB_FULL, v, m); }
// x <- x - a*b on the lanes where p holds, untouched elsewhere: ONE predicated FFMA (ptxas turns `if (p) x = fmaf(-a, b, x)` into
// FFMA + FSEL, the select on the half-rate ALU pipe)
ZB_DI float pfnma(bool p, float a, float b, float x) {
#if !defined(ZB_PRED) || ZB_PRED
  asm("{\n .reg .pred q;\n setp.ne.u32 q, %1, 0;\n @q fma.rn.f32 %0, %2, %3, %0;\n}" : "+f"(x) : "r"((unsigned)p), "f"(-a), "f"(b));
  return x;
#else
  return p ? fmaf(-a, b, x) : x;
#endif
}
ZB_DI float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(ZB_FULL, v, o);
  return v;
}

// Transpose-reduce: N values per lane are summed over the 32 lanes in N-1 (+ log2(32/N)) shuffles
// instead of 5N.  Afterwards v[0] of lane L is the total of value index (L / (32/N)).
template <int N>
ZB_DI void reduce_scatter(float (&v)[N], int lane) {
  int m = 16;
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1, m >>= 1) {
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int j = 0; j < half; j++) {
      const float send = up ? v[j] : v[j + half];
      const float keep = up ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(ZB_FULL, send, m);
    }
  }
  for (; m >= 1; m >>= 1) v[0] += __shfl_xor_sync(ZB_FULL, v[0], m);
}

ZB_DI float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
ZB_DI float dot6(const float* a, const float* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
ZB_DI void cross3(float* o, const float* a, const float* b) {
  const float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
ZB_DI void quat_mul(float* o, const float* u, const float* v) {
  const float w = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  const float x = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  const float y = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  const float z = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
ZB_DI void quat_to_mat(float* m, const float* q) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2.f * (x * y - w * z); m[2] = 2.f * (x * z + w * y);
  m[3] = 2.f * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2.f * (y * z - w * x);
  m[6] = 2.f * (x * z - w * y); m[7] = 2.f * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
ZB_DI void mat_vec(float* o, const float* m, const float* v) {
  const float x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  const float y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  const float z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
ZB_DI void matT_vec(float* o, const float* m, const float* v) {
  const float x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  const float y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  const float z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
ZB_DI void normalize4(float* q) {
  const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const float s = 1.f / (n + 1e-6f * (n == 0.f));
  q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
}
// cinert (10) x motion vector (6) -> force vector (6)
ZB_DI void inert_mul(float* o, const float* I, const float* v) {
  const float* pos = I + 6;
  const float mass = I[9];
  float c[3], e[3];
  cross3(c, pos, v + 3);
  cross3(e, pos, v);
  const float a0 = I[0] * v[0] + I[3] * v[1] + I[4] * v[2];
  const float a1 = I[3] * v[0] + I[1] * v[1] + I[5] * v[2];
  const float a2 = I[4] * v[0] + I[5] * v[1] + I[2] * v[2];
  const float l0 = mass * v[3] - e[0], l1 = mass * v[4] - e[1], l2 = mass * v[5] - e[2];
  o[0] = a0 + c[0]; o[1] = a1 + c[1]; o[2] = a2 + c[2];
  o[3] = l0; o[4] = l1; o[5] = l2;
}
ZB_DI void motion_cross(float* o, const float* u, const float* v) {
  float a[3], b[3], c[3];
  cross3(a, u, v);
  cross3(b, u, v + 3);
  cross3(c, u + 3, v);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
ZB_DI void motion_cross_force(float* o, const float* v, const float* f) {
  float a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
ZB_DI float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
ZB_DI float signf(float x) { return (float)((x > 0.f) - (x < 0.f)); }

// general x^p for the solimp power (default power == 2 never gets here); out of line so that the libdevice powf body
// exists once and stays out of the hot instruction stream
static __device__ __noinline__ float zb_pow(float x, float p) { return powf(x, p); }
