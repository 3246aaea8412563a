/*
 * orc_math.h — transcendental functions of the CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * GLSL leaves the precision of sin/cos/tan/exp/atan/asin to the implementation (the
 * reference runs them through shaderc -> SPIR-V -> a Vulkan driver).  The
 * oracle pins them to the specification below so that a second implementation can be
 * compared with it bit for bit; tests/test_math_spec.py bounds the distance from libm.
 *
 * SPECIFICATION (all arithmetic fp32, each operation rounded separately, in this order):
 *   sincos(a): j = floor(a*(2/pi) + 0.5); r = ((a - j*P1) - j*P2) - j*P3 with
 *              P1 = 1.5703125, P2 = 4.837512969970703125e-4, P3 = 7.54978995489188216e-8;
 *              z = r*r; S = r + (r*z)*(S0 + z*(S1 + z*S2)); C = (1 - 0.5*z) + (z*z)*(C0 + z*(C1 + z*C2));
 *              quadrant j mod 4 selects (S,C), (C,-S), (-S,-C), (-C,S).  |a| > 1e6 or NaN -> NaN.
 *   tan(a)   : S / C.
 *   exp(a)   : NaN -> NaN; a > 88.7228317 -> +inf; a < -87.3365402 -> 0;
 *              n = floor(a*log2(e) + 0.5); r = (a - n*0.693359375) - n*(-2.12194440e-4);
 *              p = Horner(E0..E5 in r); y = ((p*(r*r)) + r) + 1; y * 2^(n/2) * 2^(n - n/2).
 *   exp_bilateral(e) = exp(-e) of the reconstruction weight: fused-multiply-add specification, see its comment.
 *   atan, atan2, asin: Cephes atanf/asinf range reductions and polynomials (see code).
 * Coefficients are the Cephes single-precision ones.
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline float orc_bits_to_float(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline float orc_qnan(void) { return orc_bits_to_float(0x7FC00000u); }

static inline void orc_sincosf(float a, float* sn, float* cs) {
  if (a != a || a > 1.0e6f || a < -1.0e6f) {
    *sn = orc_qnan();
    *cs = orc_qnan();
    return;
  }
  const float two_over_pi = 0.636619772367581343f;
  const float P1 = 1.5703125f, P2 = 4.837512969970703125e-4f, P3 = 7.54978995489188216e-8f;
  float j = floorf(a * two_over_pi + 0.5f);
  float r = a - j * P1;
  r = r - j * P2;
  r = r - j * P3;
  float z = r * r;
  float sp = -1.9515295891e-4f;
  sp = 8.3321608736e-3f + z * sp;
  sp = -1.6666654611e-1f + z * sp;
  float rz = r * z;
  float S = r + rz * sp;
  float cp = 2.443315711809948e-5f;
  cp = -1.388731625493765e-3f + z * cp;
  cp = 4.166664568298827e-2f + z * cp;
  float zz = z * z;
  float half_z = 0.5f * z;
  float C = (1.0f - half_z) + zz * cp;
  switch (((int)j) & 3) {
    case 0: *sn = S; *cs = C; break;
    case 1: *sn = C; *cs = -S; break;
    case 2: *sn = -S; *cs = -C; break;
    default: *sn = -C; *cs = S; break;
  }
}
static inline float orc_sinf(float a) {
  float s, c;
  orc_sincosf(a, &s, &c);
  return s;
}
static inline float orc_cosf(float a) {
  float s, c;
  orc_sincosf(a, &s, &c);
  return c;
}
static inline float orc_tanf(float a) {
  float s, c;
  orc_sincosf(a, &s, &c);
  return s / c;
}

static inline float orc_expf(float a) {
  if (a != a) return a;
  if (a > 88.7228317f) return orc_bits_to_float(0x7F800000u);
  if (a < -87.3365402f) return 0.0f;
  float n = floorf(a * 1.44269504088896341f + 0.5f);
  float r = a - n * 0.693359375f;
  r = r - n * -2.12194440e-4f;
  float p = 1.9875691500e-4f * r + 1.3981999507e-3f;
  p = p * r + 8.3334519073e-3f;
  p = p * r + 4.1665795894e-2f;
  p = p * r + 1.6666665459e-1f;
  p = p * r + 5.0000001201e-1f;
  float rr = r * r;
  float y = p * rr + r;
  y = y + 1.0f;
  int ni = (int)n;
  int h = ni / 2;
  y = y * orc_bits_to_float((uint32_t)(h + 127) << 23);
  y = y * orc_bits_to_float((uint32_t)(ni - h + 127) << 23);
  return y;
}

/* exp(-e), e >= 0: the bilateral weight of reconstruction.glsl:54.  Specified on fused multiply-adds
 * (fmaf, IEEE 754-2008 — GLSL permits contracting a*b+c and leaves exp()'s precision open):
 *   e != e -> NaN; e > 87.3365402 -> 0; a = -e;
 *   t = fma(a, log2(e), 12582912); k = t - 12582912           (k = nearest integer to a*log2(e), ties to even)
 *   r = fma(k, -1.42860682030941723212e-6, fma(k, -0.693145751953125, a))
 *   q = B4; q = fma(q, r, B3); ... fma(q, r, B0); q = fma(q, r, 1); y = fma(q, r, 1)
 *   result = y * 2^k, 2^k taken from the low bits of t (t's ulp is 1, so its bit pattern is 0x4B400000 + k).
 * B0..B4 = 4.999999404e-01, 1.666652113e-01, 4.166791961e-02, 8.368702605e-03, 1.384070492e-03 (minimax fit
 * of (exp(r) - 1 - r) / r^2 on |r| <= ln(2)/2).  Measured: < 0.91 ulp from exp(-e) on all of [0, 87.34]. */
static inline float orc_exp_bilateral(float e) {
  if (e != e) return -e;
  if (e > 87.3365402f) return 0.0f;
  const float magic = 12582912.0f; /* 1.5 * 2^23 */
  float a = -e;
  float t = fmaf(a, 1.44269504088896341f, magic);
  float k = t - magic;
  float r = fmaf(k, -0.693145751953125f, a);
  r = fmaf(k, -1.42860682030941723212e-6f, r);
  static const float B[5] = {4.999999404e-01f, 1.666652113e-01f, 4.166791961e-02f, 8.368702605e-03f,
                             1.384070492e-03f};
  float q = B[4];
  for (int i = 3; i >= 0; i--) q = fmaf(q, r, B[i]);
  q = fmaf(q, r, 1.0f);
  float y = fmaf(q, r, 1.0f);
  uint32_t tb;
  memcpy(&tb, &t, 4);
  int32_t ki = (int32_t)(tb - 0x4B400000u); /* in [-126, 0] */
  return y * orc_bits_to_float((uint32_t)(ki + 127) << 23);
}

static inline float orc_atanf(float a) {
  if (a != a) return a;
  int neg = a < 0.0f;
  float t = neg ? -a : a;
  float base = 0.0f;
  if (t > 2.414213562373095f) {          /* tan(3 pi / 8) */
    base = 1.5707963267948966192f;
    t = -(1.0f / t);
  } else if (t > 0.4142135623730950f) {  /* tan(pi / 8) */
    base = 0.7853981633974483096f;
    t = (t - 1.0f) / (t + 1.0f);
  }
  float z = t * t;
  float p = 8.05374449538e-2f * z - 1.38776856032e-1f;
  p = p * z + 1.99777106478e-1f;
  p = p * z - 3.33329491539e-1f;
  float pz = p * z;
  float tail = pz * t + t;
  float y = base + tail;
  return neg ? -y : y;
}
static inline float orc_atan2f(float y, float x) {
  const float PI_F = 3.14159274101257324f;
  if (y != y || x != x) return orc_qnan();
  if (x > 0.0f) return orc_atanf(y / x);
  if (x < 0.0f) {
    float a = orc_atanf(y / x);
    return y >= 0.0f ? a + PI_F : a - PI_F;
  }
  if (y > 0.0f) return 1.5707963267948966192f;
  if (y < 0.0f) return -1.5707963267948966192f;
  return 0.0f;
}
static inline float orc_asinf(float a) {
  if (a != a) return a;
  int neg = a < 0.0f;
  float t = neg ? -a : a;
  if (t > 1.0f) return orc_qnan();
  int big = t > 0.5f;
  float z, w;
  if (big) {
    z = 0.5f * (1.0f - t);
    w = sqrtf(z);
  } else {
    w = t;
    z = w * w;
  }
  float p = 4.2163199048e-2f * z + 2.4181311049e-2f;
  p = p * z + 4.5470025998e-2f;
  p = p * z + 7.4953002686e-2f;
  p = p * z + 1.6666752422e-1f;
  float pz = p * z;
  float r = pz * w + w;
  if (big) {
    r = r + r;
    r = 1.5707963267948966192f - r;
  }
  return neg ? -r : r;
}
#endif
