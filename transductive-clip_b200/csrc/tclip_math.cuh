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

// ---- packed (2 x fp32) form of the same update ---------------------------------------------------------------------
// Blackwell (sm_100) issues FFMA2 / FMUL2 / FADD2 on register pairs: one issue slot for two fp32 FMAs.  The scalar
// kernel is issue-bound (ncu: 83 % issue-active, profiles/r1_mm_chunk_scalar.md), so the M-step evaluates two
// elements of a row per instruction wherever the operation is an add / mul / fma; only the MUFU ops and the selects
// stay scalar.  The small-a Taylor series is evaluated unconditionally here (no divergent branch).
#if defined(__CUDA_ARCH__)
TCLIP_D float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
TCLIP_D float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
TCLIP_D float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
#else
struct float2 {
  float x, y;
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float2 f2fma(float2 a, float2 b, float2 c) { return float2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline float2 f2mul(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 f2add(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
#endif
TCLIP_HD float2 f2(float c) { return make_float2(c, c); }

// ny = -y (negated once at load time); returns the updated pair.
TCLIP_HD float2 mm_update_pair(float2 a, float2 ny, float psis_hi, float psis_lo) {
  // Stirling part at X = a + 4 (same series as psi1_and_curvature_num), accumulated as -E so no negation is needed
  const float2 x2 = f2add(a, f2(2.0f));
  const float2 X = f2add(a, f2(4.0f));
  const float2 t = f2mul(x2, x2);
  const float2 P = f2mul(x2, f2add(t, f2(-1.0f)));     // (a+1)(a+2)(a+3)
  const float2 ndP = f2fma(t, f2(-3.0f), f2(1.0f));     // -(3 x2^2 - 1)
  const float2 XP = f2mul(X, P);
  const float2 R = make_float2(fast_rcp(XP.x), fast_rcp(XP.y));
  const float2 rX = f2mul(P, R);
  const float2 rP = f2mul(X, R);
  const float2 L = make_float2(fast_lg2(X.x), fast_lg2(X.y));
  const float2 LP = make_float2(fast_lg2(P.x), fast_lg2(P.y));
  const float2 z = f2mul(rX, rX);
  float2 nsp = f2fma(z, f2(1.0f / 240.0f), f2(-1.0f / 252.0f));   // -S_psi / z
  nsp = f2fma(z, nsp, f2(1.0f / 120.0f));
  nsp = f2fma(z, nsp, f2(-1.0f / 12.0f));
  float2 nsg = f2fma(z, f2(1.0f / 1680.0f), f2(-1.0f / 1260.0f));  // -S_gam / rX
  nsg = f2fma(z, nsg, f2(1.0f / 360.0f));
  nsg = f2fma(z, nsg, f2(-1.0f / 12.0f));
  const float2 nE = f2fma(ndP, rP, f2fma(f2(-0.5f), rX, f2mul(nsp, z)));   // -E
  const float2 psi1 = f2fma(L, f2(kLn2), nE);
  // N = LP ln2 - 3.5 ln2 L + (a + 4 - ln(2 pi)/2) - a E - S_gam
  float2 Ns = f2fma(nsg, rX, f2(4.0f - kHalfLn2Pi));
  Ns = f2add(f2fma(a, nE, Ns), a);
  Ns = f2fma(L, f2(-3.5f * kLn2), Ns);
  Ns = f2fma(LP, f2(kLn2), Ns);
  // Taylor form for a < 1/16: a^2 (c2 + c3 a + ... + c7 a^5), truncation < 1e-7 relative; evaluated unconditionally
  float2 ts = f2fma(a, f2(-0.864299380613076709f), f2(0.847785884987040950f));
  ts = f2fma(a, ts, f2(-0.829542204114695941f));
  ts = f2fma(a, ts, f2(0.811742425283353644f));
  ts = f2fma(a, ts, f2(-0.801371268773062857f));
  ts = f2fma(a, ts, f2(0.822467033424113218f));
  ts = f2mul(ts, f2mul(a, a));
  float2 N;
  N.x = a.x < kSmallA ? ts.x : fabsf(Ns.x);
  N.y = a.y < kSmallA ? ts.y : fabsf(Ns.y);
  // quadratic root, a-scaled and cancellation-free (see mm_update_element)
  float2 g = f2add(psi1, f2(-psis_hi));
  g = f2add(g, ny);
  g = f2add(g, f2(-psis_lo));
  const float2 bt = f2fma(a, g, f2mul(N, f2(-2.0f)));
  const float2 Dt = f2fma(bt, bt, f2mul(N, f2(8.0f)));
  const float2 r = make_float2(fast_sqrt(Dt.x), fast_sqrt(Dt.y));
  const float2 q = make_float2(fabsf(bt.x) + r.x, fabsf(bt.y) + r.y);
  const float2 aq = f2mul(a, q);
  const float2 a2 = f2add(a, a);
  const float2 N4 = f2mul(N, f2(4.0f));
  const bool px = bt.x >= 0.0f, py = bt.y >= 0.0f;
  const float2 num = make_float2(px ? a2.x : aq.x, py ? a2.y : aq.y);
  const float2 den = make_float2(px ? q.x : N4.x, py ? q.y : N4.y);
  return f2mul(num, make_float2(fast_rcp(den.x), fast_rcp(den.y)));
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
