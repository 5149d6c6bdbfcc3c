// tclip_math.cuh — special-function arithmetic for the Dirichlet MM M-step (sm_100a).
//
// Reference behaviour being reproduced (SegoleneMartin/transductive-CLIP):
//   curvature()     src/methods/zero_shot/em_dirichlet.py:153-155
//   update_alpha()  src/methods/zero_shot/em_dirichlet.py:157-177
// One MM iteration, per element a = alpha[t,k,d], with s = sum_d alpha[t,k,:]:
//   psi1 = psi(a+1);  c = a > 1e-11 ? |2 (lnG(1) - lnG(a+1) + psi1 a) / a^2| : psi'(1)
//   b = psi1 - psi(s) - c a - y;   a_new = (-b + sqrt(b^2 + 4c)) / (2c)
//
// How it is evaluated here (same function, different arithmetic — see DESIGN.md "MM kernel"):
//   * x = a+1 >= 1 is shifted by 3:  X = a+4,  P = (a+1)(a+2)(a+3),  P' = dP/da, then the Stirling series at X>=4.
//       psi(a+1)        = ln X - E,           E = 1/(2X) + S_psi(X) + P'/P
//       N := a psi1 - lnG(a+1) = -3.5 ln X + ln P + X - ln(2pi)/2 - a E - S_gam(X)        (c = 2N/a^2)
//     5 MUFU ops per element: rcp(X P) (shared by 1/X and 1/P), lg2 X, lg2 P, sqrt, rcp.
//   * the quadratic root is taken in the a-scaled, cancellation-free form
//       bt = a g - 2N  (= a b),  Dt = bt^2 + 8N (= a^2 (b^2+4c)),  q = |bt| + sqrt(Dt)
//       a_new = bt >= 0 ? 2a / q : a q / (4N)
//     which equals the reference's (-b + sqrt(b^2+4c))/(2c) exactly in real arithmetic.
//   * psi(s) is a per-row scalar: it is computed once per row and iteration in float64 and enters
//     g = psi1 - psi(s) - y as a (hi, lo) float pair, so the only error common to a whole row — the one the
//     ill-conditioned "scale" direction of the Dirichlet MLE amplifies by ~2a — is removed.
//   * a < TCLIP_SMALL_A uses the Taylor series of N around 0 (the Stirling form loses N ~ a^2 pi^2/12 to
//     cancellation there); it covers the reference's a <= 1e-11 guard, where c -> psi'(1) = pi^2/6.
//
// The file also compiles as plain C++ (TCLIP_HOST_MATH) so tests/test_math_host.py can check the series against
// SciPy on CPU; there the MUFU approximations are replaced by correctly rounded libm calls.
#pragma once

#if defined(__CUDACC__)
#define TCLIP_HD __host__ __device__ __forceinline__
#define TCLIP_D __device__ __forceinline__
#else
#define TCLIP_HD inline
#define TCLIP_D inline
#include <cmath>
#endif

namespace tclip {

constexpr float kLn2 = 0.693147180559945309f;
constexpr float kHalfLn2Pi = 0.918938533204672742f;
constexpr float kSmallA = 0.0625f;  // TCLIP_SMALL_A: below this N(a) comes from its Taylor series

// ---- MUFU wrappers -------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
TCLIP_D float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
TCLIP_D float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
TCLIP_D float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#else
inline float fast_rcp(float x) { return 1.0f / x; }
inline float fast_lg2(float x) { return (float)std::log2((double)x); }
inline float fast_sqrt(float x) { return std::sqrt(x); }
#endif

// ---- N(a) = a psi(a+1) - lnGamma(a+1) for small a: sum_{k>=2} (-1)^k zeta(k) (1 - 1/k) a^k ------------------------
// |a| < 1/16: 8 terms leave a relative truncation error < 1e-9.
TCLIP_HD float curvature_num_small(float a) {
  const float c2 = 0.822467033424113218f;    // zeta(2)/2
  const float c3 = -0.801371268773062857f;   // -2 zeta(3)/3
  const float c4 = 0.811742425283353644f;    // 3 zeta(4)/4
  const float c5 = -0.829542204114695941f;   // -4 zeta(5)/5
  const float c6 = 0.847785884987040950f;    // 5 zeta(6)/6
  const float c7 = -0.864299380613076709f;   // -6 zeta(7)/7
  const float c8 = 0.878567686673201297f;    // 7 zeta(8)/8
  const float c9 = -0.890674126956517524f;   // -8 zeta(9)/9
  float p = fmaf(a, c9, c8);
  p = fmaf(a, p, c7);
  p = fmaf(a, p, c6);
  p = fmaf(a, p, c5);
  p = fmaf(a, p, c4);
  p = fmaf(a, p, c3);
  p = fmaf(a, p, c2);
  return p * a * a;
}

// psi(a+1) and N(a) = a psi(a+1) - lnGamma(a+1) for a >= 0 (fp32, ~1-2 ulp of ln X).
struct PsiN {
  float psi1;
  float N;
};

TCLIP_HD PsiN psi1_and_curvature_num(float a) {
  const float x2 = a + 2.0f;
  const float X = a + 4.0f;
  const float t = x2 * x2;
  const float P = fmaf(x2, t, -x2);       // (a+1)(a+2)(a+3) = x2 (x2^2 - 1)
  const float dP = fmaf(3.0f, t, -1.0f);  // d/da of the above = 3 x2^2 - 1
  const float R = fast_rcp(X * P);
  const float rX = P * R;
  const float rP = X * R;
  const float L = fast_lg2(X);
  const float LP = fast_lg2(P);
  const float z = rX * rX;
  // S_psi = 1/(12X^2) - 1/(120X^4) + 1/(252X^6) - 1/(240X^8)
  float sp = fmaf(z, -1.0f / 240.0f, 1.0f / 252.0f);
  sp = fmaf(z, sp, -1.0f / 120.0f);
  sp = fmaf(z, sp, 1.0f / 12.0f);
  sp *= z;
  // S_gam = 1/(12X) - 1/(360X^3) + 1/(1260X^5) - 1/(1680X^7)
  float sg = fmaf(z, -1.0f / 1680.0f, 1.0f / 1260.0f);
  sg = fmaf(z, sg, -1.0f / 360.0f);
  sg = fmaf(z, sg, 1.0f / 12.0f);
  sg *= rX;
  const float E = fmaf(dP, rP, fmaf(0.5f, rX, sp));
  PsiN out;
  out.psi1 = fmaf(L, kLn2, -E);
  float N = fmaf(LP, kLn2, fmaf(L, -3.5f * kLn2, X - kHalfLn2Pi)) - fmaf(a, E, sg);
  if (a < kSmallA) N = curvature_num_small(a);
  out.N = fabsf(N);
  return out;
}

// One MM update of one element.  psis = psi(sum_d alpha) as hi + lo.
TCLIP_HD float mm_update_element(float a, float y, float psis_hi, float psis_lo) {
  const PsiN pn = psi1_and_curvature_num(a);
  const float g = ((pn.psi1 - psis_hi) - y) - psis_lo;
  const float bt = fmaf(a, g, -2.0f * pn.N);
  const float Dt = fmaf(bt, bt, 8.0f * pn.N);
  const float q = fabsf(bt) + fast_sqrt(Dt);
  const bool pos = bt >= 0.0f;
  const float num = pos ? 2.0f * a : a * q;
  const float den = pos ? q : 4.0f * pn.N;
  return num * fast_rcp(den);
}

// psi(s) in float64, s > 0.  Used once per row and MM iteration (and by the host tests).
TCLIP_HD double digamma_f64(double s) {
  double acc = 0.0;
  while (s < 10.0) {  // only rows with a tiny total mass take this path
    acc -= 1.0 / s;
    s += 1.0;
  }
  const double r = 1.0 / s;
  const double r2 = r * r;
  double ser = fma(r2, -1.0 / 12.0, 691.0 / 32760.0);
  ser = fma(r2, ser, -1.0 / 132.0);
  ser = fma(r2, ser, 1.0 / 240.0);
  ser = fma(r2, ser, -1.0 / 252.0);
  ser = fma(r2, ser, 1.0 / 120.0);
  ser = fma(r2, ser, -1.0 / 12.0);
  return log(s) + fma(r2, ser, -0.5 * r) + acc;
}

}  // namespace tclip
