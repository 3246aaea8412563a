// Exact fp32 arithmetic + the transcendental functions of the Hijiki hot path.
//
// The reference shaders (reference shader/*.glsl) leave three things to the GLSL
// implementation: whether a*b+c is contracted, the precision of sin/cos/tan/exp/atan/asin,
// and how normalize()/length() are evaluated.  To make the CUDA path comparable with the
// CPU oracle BIT FOR BIT they are fixed here once:
//   * every fp32 operation of the shader arithmetic is a separately rounded IEEE-754
//     operation in source order (x::mul/add/sub map to __fmul_rn/__fadd_rn/__fsub_rn, which
//     nvcc never contracts; division and sqrt are the correctly rounded __fdiv_rn/__fsqrt_rn);
//   * the transcendentals are the fixed polynomial kernels below (Cephes single-precision
//     coefficients, <= 2 ulp on the ranges the path uses), built from those exact ops only;
//   * normalize(v) = v * (1/sqrt(dot(v,v))), length(v) = sqrt(dot(v,v)).
// The oracle carries its own, independently written copy of the same specification
// (oracle/orc_math.h); tests/ compare the two bit-for-bit and both against libm.
//
// Everything here is HJK_HD so that tests/native can compile the device headers for the
// host (g++ -ffp-contract=off) and unit-test the kernels' logic without a GPU.  The
// product library never runs this code on the CPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define HJK_HD __host__ __device__ __forceinline__
#define HJK_D __device__ __forceinline__
#else
#define HJK_HD inline
#define HJK_D inline
#endif

namespace hjk {
namespace x {  // exact (never contracted) scalar ops

HJK_HD float mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
HJK_HD float add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
HJK_HD float sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
HJK_HD float div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
HJK_HD float sqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return ::sqrtf(a);
#endif
}
HJK_HD float u2f(uint32_t u) {  // GLSL float(uint): round to nearest even
#if defined(__CUDA_ARCH__)
  return __uint2float_rn(u);
#else
  return (float)u;
#endif
}
HJK_HD float floor(float a) { return ::floorf(a); }
HJK_HD float as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union {
    uint32_t u;
    float f;
  } c;
  c.u = u;
  return c.f;
#endif
}
HJK_HD uint32_t as_uint(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union {
    uint32_t u;
    float f;
  } c;
  c.f = f;
  return c.u;
#endif
}
// GLSL min/max as the oracle fixes them (NaN behaviour included)
HJK_HD float gmin(float a, float b) { return b < a ? b : a; }
HJK_HD float gmax(float a, float b) { return a < b ? b : a; }
HJK_HD bool is_nan(float a) { return a != a; }

}  // namespace x

// ------------------------------------------------------------------ vec3 on exact ops
struct vec3 {
  float x, y, z;
};
HJK_HD vec3 V3(float a, float b, float c) {
  vec3 r;
  r.x = a;
  r.y = b;
  r.z = c;
  return r;
}
HJK_HD vec3 V3(float s) { return V3(s, s, s); }
HJK_HD vec3 operator+(vec3 a, vec3 b) { return V3(x::add(a.x, b.x), x::add(a.y, b.y), x::add(a.z, b.z)); }
HJK_HD vec3 operator-(vec3 a, vec3 b) { return V3(x::sub(a.x, b.x), x::sub(a.y, b.y), x::sub(a.z, b.z)); }
HJK_HD vec3 operator-(vec3 a) { return V3(-a.x, -a.y, -a.z); }
HJK_HD vec3 operator*(vec3 a, vec3 b) { return V3(x::mul(a.x, b.x), x::mul(a.y, b.y), x::mul(a.z, b.z)); }
HJK_HD vec3 operator*(vec3 a, float s) { return V3(x::mul(a.x, s), x::mul(a.y, s), x::mul(a.z, s)); }
HJK_HD vec3 operator*(float s, vec3 a) { return V3(x::mul(s, a.x), x::mul(s, a.y), x::mul(s, a.z)); }
HJK_HD vec3 operator/(vec3 a, float s) { return V3(x::div(a.x, s), x::div(a.y, s), x::div(a.z, s)); }
HJK_HD vec3 operator/(vec3 a, vec3 b) { return V3(x::div(a.x, b.x), x::div(a.y, b.y), x::div(a.z, b.z)); }
HJK_HD float dot(vec3 a, vec3 b) {
  return x::add(x::add(x::mul(a.x, b.x), x::mul(a.y, b.y)), x::mul(a.z, b.z));
}
HJK_HD vec3 cross(vec3 a, vec3 b) {
  return V3(x::sub(x::mul(a.y, b.z), x::mul(b.y, a.z)), x::sub(x::mul(a.z, b.x), x::mul(b.z, a.x)),
            x::sub(x::mul(a.x, b.y), x::mul(b.x, a.y)));
}
HJK_HD float length(vec3 a) { return x::sqrt(dot(a, a)); }
HJK_HD vec3 normalize(vec3 a) { return a * x::div(1.0f, x::sqrt(dot(a, a))); }
HJK_HD vec3 reflect(vec3 I, vec3 N) { return I - x::mul(2.0f, dot(N, I)) * N; }

// math.glsl:1
#define HJK_PI_F 3.14159274101257324f /* fp32(3.1415926535897932384626433832795) */

// ------------------------------------------------------------------ transcendentals
// sin and cos of one argument.  Cody-Waite reduction by pi/2 in three exact pieces, then the
// Cephes sinf/cosf minimax kernels on [-pi/4, pi/4].
HJK_HD void sincos_det(float a, float* s_out, float* c_out) {
  if (x::is_nan(a) || a > 1.0e6f || a < -1.0e6f) {  // outside the reduction's range: defined as NaN
    *s_out = *c_out = x::as_float(0x7FC00000u);
    return;
  }
  float j = x::floor(x::add(x::mul(a, 0.636619772367581343f), 0.5f));
  float r = x::sub(a, x::mul(j, 1.5703125f));
  r = x::sub(r, x::mul(j, 4.837512969970703125e-4f));
  r = x::sub(r, x::mul(j, 7.54978995489188216e-8f));
  int q = ((int)j) & 3;
  float z = x::mul(r, r);
  // sin kernel: r + r*z*(S0 + z*(S1 + z*S2))
  float ps = x::add(8.3321608736e-3f, x::mul(z, -1.9515295891e-4f));
  ps = x::add(-1.6666654611e-1f, x::mul(z, ps));
  float s = x::add(r, x::mul(x::mul(r, z), ps));
  // cos kernel: 1 - z/2 + z*z*(C0 + z*(C1 + z*C2))
  float pc = x::add(-1.388731625493765e-3f, x::mul(z, 2.443315711809948e-5f));
  pc = x::add(4.166664568298827e-2f, x::mul(z, pc));
  float c = x::add(x::sub(1.0f, x::mul(0.5f, z)), x::mul(x::mul(z, z), pc));
  float so = (q & 1) ? c : s;
  float co = (q & 1) ? s : c;
  if (q & 2) so = -so;
  if (q == 1 || q == 2) co = -co;
  *s_out = so;
  *c_out = co;
}
HJK_HD float sin_det(float a) {
  float s, c;
  sincos_det(a, &s, &c);
  return s;
}
HJK_HD float cos_det(float a) {
  float s, c;
  sincos_det(a, &s, &c);
  return c;
}
HJK_HD float tan_det(float a) {
  float s, c;
  sincos_det(a, &s, &c);
  return x::div(s, c);
}

// exp: n = round(x*log2 e), r = x - n*ln2 (two exact pieces), Cephes expf polynomial, 2^n by
// exponent construction.  exp_det(0) == 1 exactly.  Results below 2^-126 are flushed to 0.
HJK_HD float exp_det(float a) {
  if (x::is_nan(a)) return a;
  if (a > 88.7228317f) return x::as_float(0x7F800000u);
  if (a < -87.3365402f) return 0.0f;
  float n = x::floor(x::add(x::mul(a, 1.44269504088896341f), 0.5f));
  float r = x::sub(a, x::mul(n, 0.693359375f));
  r = x::sub(r, x::mul(n, -2.12194440e-4f));
  float z = x::mul(r, r);
  float p = x::add(x::mul(1.9875691500e-4f, r), 1.3981999507e-3f);
  p = x::add(x::mul(p, r), 8.3334519073e-3f);
  p = x::add(x::mul(p, r), 4.1665795894e-2f);
  p = x::add(x::mul(p, r), 1.6666665459e-1f);
  p = x::add(x::mul(p, r), 5.0000001201e-1f);
  float y = x::add(x::add(x::mul(p, z), r), 1.0f);
  int ni = (int)n;
  int n1 = ni / 2, n2 = ni - n1;  // two factors so that n = 128 and n = -126 stay representable
  y = x::mul(y, x::as_float((uint32_t)(n1 + 127) << 23));
  y = x::mul(y, x::as_float((uint32_t)(n2 + 127) << 23));
  return y;
}

// Bilateral weight of the reconstruction pass, exp(-e) for e >= 0 (reconstruction.glsl:54), the one
// transcendental evaluated per filter tap (~10^8 times per 4K pass).  GLSL leaves both the precision of
// exp() and the contraction of a*b+c to the implementation, so this one is specified on FUSED multiply-adds
// (exactly defined in IEEE 754-2008; __fmaf_rn on the device, fmaf() in the oracle): 13 operations
// instead of the ~30 separately rounded ones of exp_det, and < 1 ulp from the true value over the whole range
// (tests/test_math_spec.py: 0.91 ulp).
//   e NaN -> NaN (any payload); e > 87.3365402 -> 0; otherwise a = -e,
//   t = fma(a, log2 e, 1.5*2^23); n = t - 1.5*2^23   (n = a*log2(e) rounded to the nearest integer, ties to even)
//   r = fma(n, -LN2_LO, fma(n, -LN2_HI, a))          (LN2_HI = 0x1.62e4p-1 has 16 significant bits: n*LN2_HI exact)
//   y = fma(fma(q(r), r, 1), r, 1), q = Horner(B4..B0) in fma;   result = y * 2^n (2^n built from the bits of t)
// exp_fma_neg(0) == 1 exactly.
HJK_HD float fma_rn(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return ::fmaf(a, b, c);
#endif
}
HJK_HD float exp_fma_neg(float e) {
  // straight-line: a NaN e flows through every operation below, an e past the cutoff is replaced at the end
  const float a = -e;
  const float t = fma_rn(a, 1.44269504088896341f, 12582912.0f);
  const float n = x::sub(t, 12582912.0f);
  float r = fma_rn(n, -0.693145751953125f, a);
  r = fma_rn(n, -1.42860682030941723212e-6f, r);
  float p = 1.384070492e-03f;
  p = fma_rn(p, r, 8.368702605e-03f);
  p = fma_rn(p, r, 4.166791961e-02f);
  p = fma_rn(p, r, 1.666652113e-01f);
  p = fma_rn(p, r, 4.999999404e-01f);
  p = fma_rn(p, r, 1.0f);
  const float y = fma_rn(p, r, 1.0f);
  const float v = x::mul(y, x::as_float((x::as_uint(t) << 23) + 0x3F800000u));
  return e > 87.3365402f ? 0.0f : v;
}

// atan on the whole line (Cephes atanf), then atan2 by quadrant.
HJK_HD float atan_det(float a) {
  if (x::is_nan(a)) return a;
  bool neg = a < 0.0f;
  float t = neg ? -a : a;
  float y;
  if (t > 2.414213562373095f) {
    y = 1.5707963267948966192f;
    t = -x::div(1.0f, t);
  } else if (t > 0.4142135623730950f) {
    y = 0.7853981633974483096f;
    t = x::div(x::sub(t, 1.0f), x::add(t, 1.0f));
  } else {
    y = 0.0f;
  }
  float z = x::mul(t, t);
  float p = x::sub(x::mul(8.05374449538e-2f, z), 1.38776856032e-1f);
  p = x::add(x::mul(p, z), 1.99777106478e-1f);
  p = x::sub(x::mul(p, z), 3.33329491539e-1f);
  y = x::add(y, x::add(x::mul(x::mul(p, z), t), t));
  return neg ? -y : y;
}
HJK_HD float atan2_det(float yy, float xx) {
  if (x::is_nan(yy) || x::is_nan(xx)) return x::as_float(0x7FC00000u);
  if (xx > 0.0f) return atan_det(x::div(yy, xx));
  if (xx < 0.0f) {
    float a = atan_det(x::div(yy, xx));
    return yy >= 0.0f ? x::add(a, HJK_PI_F) : x::sub(a, HJK_PI_F);
  }
  if (yy > 0.0f) return 1.5707963267948966192f;
  if (yy < 0.0f) return -1.5707963267948966192f;
  return 0.0f;
}
// asin on [-1, 1] (Cephes asinf)
HJK_HD float asin_det(float a) {
  if (x::is_nan(a)) return a;
  bool neg = a < 0.0f;
  float t = neg ? -a : a;
  if (t > 1.0f) return x::as_float(0x7FC00000u);
  bool big = t > 0.5f;
  float z, w;
  if (big) {
    z = x::mul(0.5f, x::sub(1.0f, t));
    w = x::sqrt(z);
  } else {
    w = t;
    z = x::mul(w, w);
  }
  float p = x::add(x::mul(4.2163199048e-2f, z), 2.4181311049e-2f);
  p = x::add(x::mul(p, z), 4.5470025998e-2f);
  p = x::add(x::mul(p, z), 7.4953002686e-2f);
  p = x::add(x::mul(p, z), 1.6666752422e-1f);
  float r = x::add(x::mul(x::mul(p, z), w), w);
  if (big) {
    r = x::add(r, r);
    r = x::sub(1.5707963267948966192f, r);
  }
  return neg ? -r : r;
}

}  // namespace hjk
