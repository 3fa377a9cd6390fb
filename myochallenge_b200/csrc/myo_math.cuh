// Small fp32 vector / quaternion / spatial-algebra helpers for the step kernel.
// Spatial vectors are [rotation(3); translation(3)] about a per-tree reference point, the same
// convention MuJoCo's cdof/cvel/cacc use (engine_core_smooth.c), so stage dumps compare directly.
#pragma once
#include <cuda_runtime.h>

namespace myo {

#define MYO_DI __device__ __forceinline__
// phases are real functions: one copy of each in the instruction stream keeps the kernel inside the instruction cache
#define MYO_PHASE __device__ __noinline__
constexpr float kMinVal = 1e-15f;
constexpr float kPi = 3.14159265358979323846f;

// 1/sqrt(x) for x known to be a normal positive number (callers clamp to kMinVal): the bare MUFU.RSQ, without the
// denormal rescaling rsqrtf() wraps around it
MYO_DI float rsqrt_pos(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.f / sqrtf(x);
#endif
}
MYO_DI float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
MYO_DI void cross3(float* r, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
MYO_DI void sub3(float* r, const float* a, const float* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
MYO_DI void add3(float* r, const float* a, const float* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
MYO_DI void cpy3(float* r, const float* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
MYO_DI float norm3(const float* a) { return sqrtf(dot3(a, a)); }
MYO_DI float normalize3(float* a) {
  float n = norm3(a);
  if (n < kMinVal) { a[0] = 1.f; a[1] = 0.f; a[2] = 0.f; }
  else { float inv = 1.f / n; a[0] *= inv; a[1] *= inv; a[2] *= inv; }
  return n;
}
MYO_DI void normalize4(float* q) {
  float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < kMinVal) { q[0] = 1.f; q[1] = q[2] = q[3] = 0.f; }
  else { float inv = 1.f / n; q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv; }
}
MYO_DI void mulquat(float* r, const float* a, const float* b) {
  float t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  float t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  float t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  float t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
MYO_DI void quat2mat(float* R, const float* q) {
  float q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3], q11 = q[1] * q[1],
        q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[4] = q00 - q11 + q22 - q33; R[8] = q00 - q11 - q22 + q33;
  R[1] = 2.f * (q12 - q03); R[2] = 2.f * (q13 + q02); R[3] = 2.f * (q12 + q03);
  R[5] = 2.f * (q23 - q01); R[6] = 2.f * (q13 - q02); R[7] = 2.f * (q23 + q01);
}
MYO_DI void mulmatvec3(float* r, const float* R, const float* v) {
  float x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  float y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  float z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
MYO_DI void mulmatTvec3(float* r, const float* R, const float* v) {
  float x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  float y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  float z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
MYO_DI void rotvecquat(float* r, const float* v, const float* q) {
  // r = v + 2 w (u x v) + 2 u x (u x v)
  float u[3] = {q[1], q[2], q[3]}, t[3], t2[3];
  cross3(t, u, v);
  t[0] *= 2.f; t[1] *= 2.f; t[2] *= 2.f;
  cross3(t2, u, t);
  r[0] = v[0] + q[0] * t[0] + t2[0]; r[1] = v[1] + q[0] * t[1] + t2[1]; r[2] = v[2] + q[0] * t[2] + t2[2];
}
MYO_DI void mulmat3(float* C, const float* A, const float* B) {
  float t[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) t[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
#pragma unroll
  for (int k = 0; k < 9; k++) C[k] = t[k];
}
MYO_DI float clipf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

MYO_DI void cross_motion(float* r, const float* vel, const float* v) {
  float a[3], b[3], c[3];
  cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
MYO_DI void cross_force(float* r, const float* vel, const float* f) {
  float a[3], b[3], c[3];
  cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
// 10-number inertia about the reference point: Ixx Iyy Izz Ixy Ixz Iyz m*cx m*cy m*cz m
MYO_DI void mul_inert_vec(float* r, const float* i, const float* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
MYO_DI void inert_com(float* res, const float* inert, const float* mat, const float* dif, float mass) {
  float t[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      t[3 * r + c] = mat[3 * r] * inert[0] * mat[3 * c] + mat[3 * r + 1] * inert[1] * mat[3 * c + 1] +
                     mat[3 * r + 2] * inert[2] * mat[3 * c + 2];
  res[0] = t[0] + mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] = t[4] + mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] = t[8] + mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] = t[1] - mass * dif[0] * dif[1];
  res[4] = t[2] - mass * dif[0] * dif[2];
  res[5] = t[5] - mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2]; res[9] = mass;
}

// Philox4x32-10 counter-based RNG: stream keyed by (seed, world), counter = (episode, draw block).
struct Philox {
  uint32_t key[2], ctr[4], out[4];
  int have;
  MYO_DI void init(uint64_t seed, uint32_t world, uint32_t episode) {
    key[0] = (uint32_t)seed ^ (world * 0x9E3779B9u); key[1] = (uint32_t)(seed >> 32) ^ 0xBB67AE85u ^ world;
    ctr[0] = 0; ctr[1] = episode; ctr[2] = world; ctr[3] = 0x5851F42Du;
    have = 0;
  }
  MYO_DI void round(uint32_t* c, const uint32_t* k) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  MYO_DI void refill() {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]}, k[2] = {key[0], key[1]};
#pragma unroll
    for (int r = 0; r < 10; r++) { round(c, k); k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    ctr[0]++; have = 4;
  }
  MYO_DI float uniform() {   // [0, 1)
    if (!have) refill();
    uint32_t u = out[--have];
    return (float)(u >> 8) * (1.0f / 16777216.0f);
  }
  MYO_DI float uniform(float lo, float hi) { return lo + (hi - lo) * uniform(); }
  MYO_DI float normal() {      // Box-Muller, one value per call
    const float u1 = 1.f - uniform(), u2 = uniform();
    return sqrtf(-2.f * logf(u1)) * cosf(6.283185307179586f * u2);
  }
  MYO_DI float gamma(float a) {   // Marsaglia-Tsang; a < 1 through gamma(a + 1) u^(1/a)
    float boost = 1.f;
    if (a < 1.f) { boost = powf(1.f - uniform(), 1.f / a); a += 1.f; }
    const float d = a - 1.f / 3.f, cc = 1.f / sqrtf(9.f * d);
    for (int it = 0; it < 64; it++) {
      const float x = normal();
      float v = 1.f + cc * x;
      if (v <= 0.f) continue;
      v = v * v * v;
      const float u = 1.f - uniform();
      if (logf(u) < 0.5f * x * x + d - d * v + d * logf(v)) return boost * d * v;
    }
    return boost * d;
  }
  MYO_DI float beta(float a, float b) {   // numpy Generator/RandomState.beta: X / (X + Y), X ~ Gamma(a), Y ~ Gamma(b)
    const float x = gamma(a), y = gamma(b);
    return x / fmaxf(x + y, 1e-30f);
  }
};

// MyoSuite utils/quat_math.py conventions (from mujoco-worldgen's rotations.py): extrinsic xyz Euler angles
MYO_DI void euler2quat(float* q, const float* e) {
  const float ai = 0.5f * e[2], aj = -0.5f * e[1], ak = 0.5f * e[0];
  float si, ci, sj, cj, sk, ck;
  sincosf(ai, &si, &ci); sincosf(aj, &sj, &cj); sincosf(ak, &sk, &ck);
  const float cc = ci * ck, cs = ci * sk, sc = si * ck, ss = si * sk;
  q[0] = cj * cc + sj * ss;
  q[3] = cj * sc - sj * cs;
  q[2] = -(cj * ss + sj * cc);
  q[1] = cj * cs - sj * sc;
}
MYO_DI void mat2euler(float* e, const float* R) {      // R row major
  const float cy = sqrtf(R[8] * R[8] + R[5] * R[5]);
  if (cy > 4.f * 1.1920929e-7f) {
    e[2] = -atan2f(R[1], R[0]);
    e[1] = -atan2f(-R[2], cy);
    e[0] = -atan2f(R[5], R[8]);
  } else {
    e[2] = -atan2f(-R[3], R[4]);
    e[1] = -atan2f(-R[2], cy);
    e[0] = 0.f;
  }
}

}  // namespace myo
