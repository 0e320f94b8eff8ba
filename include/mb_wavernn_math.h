/*
 * Canonical scalar math of the WaveRNN sample loop, shared by the CUDA kernel
 * (mockingbird_b200/csrc/wavernn.cu) and the CPU twin (oracle/wavernn_twin.c).
 *
 * Everything here is built from IEEE-754 binary32 +, -, *, /, fma and integer bit operations only,
 * so that gcc (-O2 -ffp-contract=off, fmaf from libm / the FMA unit) and nvcc (no fast-math,
 * fmaf -> FFMA.RN, '/' -> IEEE division) produce bit-identical results.  That is what makes
 * "kernel == twin, bit-exact, free running" a testable property (hard part 1).
 *
 * Accuracy vs the reference's libm/Sleef calls: exp/log <= ~2 ulp, sigmoid/tanh absolute error
 * <= 2e-7 - far below the logit noise floor that matters for argmax(p/q) sampling.
 */
#ifndef MB_WAVERNN_MATH_H
#define MB_WAVERNN_MATH_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define MB_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define MB_HD static inline
#endif

MB_HD float mb_u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
MB_HD uint32_t mb_f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

/* exp(x), x <= ~88.  Cody-Waite reduction, degree-6 polynomial, exact 2^n scaling. */
MB_HD float mb_expf(float x) {
  if (x < -87.0f) return 0.0f;
  if (x > 88.0f) x = 88.0f;
  const float n = rintf(x * 1.4426950408889634f);
  float r = fmaf(n, -0.693145751953125f, x);        /* ln2 high part (exact product) */
  r = fmaf(n, -1.42860682030941723212e-6f, r);       /* ln2 low part */
  float p = 1.3888889225e-3f;                        /* 1/720 */
  p = fmaf(p, r, 8.3333337680e-3f);                  /* 1/120 */
  p = fmaf(p, r, 4.1666667908e-2f);                  /* 1/24  */
  p = fmaf(p, r, 1.6666667163e-1f);                  /* 1/6   */
  p = fmaf(p, r, 0.5f);
  p = fmaf(p, r, 1.0f);
  p = fmaf(p, r, 1.0f);
  const int32_t e = (int32_t)n;
  /* 2^e by exponent-field construction; split so that the intermediate stays normal */
  const int32_t e1 = e / 2, e2 = e - e1;
  const float s1 = mb_u2f((uint32_t)(e1 + 127) << 23);
  const float s2 = mb_u2f((uint32_t)(e2 + 127) << 23);
  return (p * s1) * s2;
}

/* log(x) for x > 0 (normal).  x = 2^k * m, m in [sqrt(.5), sqrt(2)), log(m) via atanh series. */
MB_HD float mb_logf(float x) {
  uint32_t ix = mb_f2u(x);
  int32_t k = (int32_t)(ix >> 23) - 127;
  ix = (ix & 0x007fffffu) | 0x3f800000u;
  float m = mb_u2f(ix);
  if (m > 1.41421356f) {
    m = m * 0.5f;
    k += 1;
  }
  const float f = m - 1.0f;
  const float s = f / (2.0f + f);
  const float z = s * s;
  float p = 0.2222222222f;           /* 2/9 */
  p = fmaf(p, z, 0.2857142857f);     /* 2/7 */
  p = fmaf(p, z, 0.4f);              /* 2/5 */
  p = fmaf(p, z, 0.6666666667f);     /* 2/3 */
  p = fmaf(p, z, 2.0f);
  const float lm = s * p;            /* log(m) = 2 atanh(s) */
  const float fk = (float)k;
  return fmaf(fk, 0.693145751953125f, fmaf(fk, 1.42860682030941723212e-6f, lm));
}

MB_HD float mb_sigmoidf(float x) {
  /* 1 / (1 + exp(-x)); symmetric form keeps the exp argument <= 0 */
  const float e = mb_expf(x < 0.0f ? x : -x);
  const float s = 1.0f / (1.0f + e);
  return x < 0.0f ? e * s : s;
}

MB_HD float mb_tanhf(float x) {
  const float ax = x < 0.0f ? -x : x;
  float t;
  if (ax < 0.0625f) {
    const float z = ax * ax;           /* odd Taylor series: relative accuracy near 0 */
    float p = 0.0539682540f;           /* 17/315 */
    p = fmaf(p, z, -0.1333333333f);    /* -2/15 */
    p = fmaf(p, z, 0.3333333333f);
    t = fmaf(-(ax * z), p, ax);        /* ax - ax*z*(1/3 - 2/15 z + 17/315 z^2) */
  } else {
    const float e = mb_expf(-2.0f * ax);
    t = (1.0f - e) / (1.0f + e);
  }
  return x < 0.0f ? -t : t;
}

/* ---- counter-based generator (Philox-4x32-10) for the built-in Exp(1) noise ------------------- */
MB_HD void mb_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t out[4]) {
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Exp(1) draw number `cls` (0..511) of (row, step) under `seed`: -log(u), u in (0,1) */
MB_HD float mb_exp1_noise(uint64_t seed, uint32_t step, uint32_t row, uint32_t cls) {
  uint32_t o[4];
  mb_philox4x32(cls >> 2, row, step, 0x4d425756u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  const uint32_t bits = o[cls & 3];
  const float u = ((float)(bits >> 9) + 0.5f) * 1.1920928955078125e-7f; /* 2^-23; exact, in (0,1) */
  return -mb_logf(u);
}

#endif /* MB_WAVERNN_MATH_H */
