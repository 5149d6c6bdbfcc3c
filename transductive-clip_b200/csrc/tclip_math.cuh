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
#include <cstring>

#if defined(__CUDACC__)
#define TCLIP_HD __host__ __device__ __forceinline__
#define TCLIP_D __device__ __forceinline__
#else
#define TCLIP_HD inline
#define TCLIP_D inline
#include <cmath>
#include <cstring>
#endif

namespace tclip {

constexpr float kLn2 = 0.693147180559945309f;
constexpr float kHalfLn2Pi = 0.918938533204672742f;
constexpr float kSmallA = 0.0625f;  // TCLIP_SMALL_A: below this N(a) comes from its Taylor series

// ---- MUFU wrappers -------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
// TCLIP_EXACT_{RCP,LG2,SQRT}: diagnostic builds that swap one MUFU approximation for the correctly rounded operation
// (scripts/build_variants.sh) to attribute the alpha error budget; never defined in the product build.
TCLIP_D float fast_rcp(float x) {
#ifdef TCLIP_EXACT_RCP
  return __frcp_rn(x);
#endif
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
TCLIP_D float fast_lg2(float x) {
#ifdef TCLIP_EXACT_LG2
  return log2f(x);
#endif
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
TCLIP_D float fast_sqrt(float x) {
#ifdef TCLIP_EXACT_SQRT
  return __fsqrt_rn(x);
#endif
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
#else
inline float fast_rcp(float x) { return 1.0f / x; }
// Host model of MUFU.LG2 as measured on B200 (profiles/r1_lg2_error.txt): results with |r| >= 1 are truncated toward
// zero to float32 (<= 1 ulp, biased), results in (-1, 1) carry ~2e-7 absolute error (modelled as rounding to 2^-22).
inline int& host_lg2_exact() {  // tests may switch the model off (1) to separate series error from MUFU error
  static int exact = 0;
  return exact;
}
inline float fast_lg2(float x) {
  const double r = std::log2((double)x);
  if (host_lg2_exact()) return (float)r;
  if (std::fabs(r) < 1.0) return (float)(std::nearbyint(r * 4194304.0) / 4194304.0);
  float f = (float)r;
  if (std::fabs((double)f) > std::fabs(r)) f = std::nextafterf(f, 0.0f);
  return f;
}
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

// x = 2^e m with m in [sqrt(1/2), sqrt(2)): returns m and e * 2^23 as a float (exact), integer pipe only.
TCLIP_HD void split_exponent(float x, float& m, float& e23) {
#if defined(__CUDA_ARCH__)
  const int b = __float_as_int(x);
  const int e = (b - 0x3f3504f3) & 0xff800000;
  m = __int_as_float(b - e);
  e23 = (float)e;
#else
  int b;
  std::memcpy(&b, &x, 4);
  const int e = (int)((unsigned)(b - 0x3f3504f3) & 0xff800000u);
  const int mb = b - e;
  std::memcpy(&m, &mb, 4);
  e23 = (float)e;
#endif
}

// ny = -y (negated once at load time); returns the updated pair.
//
// Shift by 4 here (X = a + 5 >= 5, P = (a+1)(a+2)(a+3)(a+4) = w^2 - 1 with w = a^2 + 5a + 5), so both Stirling series
// need three terms only (next terms < 1e-8), and everything is accumulated with the signs / factors the root needs:
//   nE = -E,   E = 1/(2X) + S_psi(X) + P'/P,   psi(a+1) = ln X + nE
//   M  = 2N >= 2 * 0.003 for a >= 1/16, far above its rounding error, so the reference's |.| is a no-op there
//   M  = 2N = 2 [ln P - 4.5 ln X + (a + 5 - ln(2 pi)/2) + a nE - S_gam(X)]
//   bt = a g - M,  Dt = bt^2 + 4M,  q = |bt| + sqrt(Dt),  a_new = bt >= 0 ? 2a / q : 2a q / (4M)
//
// ln X is NOT taken from MUFU.LG2: that unit truncates results with |r| >= 1 to float32 (mean error -0.45 ulp, e.g.
// -9e-7 at x ~ 1e5) and is biased by +4..8e-8 even for results in (-1, 1) (measured: profiles/r1_lg2_error.txt).  A bias
// common to the heavy elements of a row moves the ill-conditioned precision direction of the Dirichlet fit by
// ~ 2 s / (D - 1) times the bias (3e3 at s = 1.7e5, D = 100): with MUFU.LG2 the alpha error against the float64
// restatement was 5x the reference's own float32 error, with a correctly rounded log it is below it
// (profiles/r1_alpha_error_attribution.txt).  So ln X = e ln2 + ln m with the exponent split off on the integer pipe and
// ln m from a degree-6 minimax polynomial (7e-9 absolute) in round-to-nearest FFMA2s, and the row total s enters as
//   psi(a+1) - psi(s) = (e - k) ln2 + ln m + nE - (psi(s) - k ln2),      2^k ~ s,
// so the heavy elements (X ~ s) see small, accurately rounded terms only.  lg2(P), which only feeds the curvature,
// stays on the MUFU.
struct RowPsi {
  float dpsi;  // psi(s) - k ln2, |.| < ~0.4 + 1/(2s): float32 carries it to ~3e-8, no lo word needed
  float k23;   // k * 2^23 (exact)
};

// kSmall selects how the a < 1/16 Taylor form is handled: 2 = warp vote + uniform branch per pair (any caller), 1 = always
// evaluated and selected per element, 0 = skipped — the caller has established that no element of the warp is below 1/16
// (mm_chunk_kernel votes once per row and iteration on the running minimum instead of once per pair).
template <int kSmall = 2>
TCLIP_HD float2 mm_update_pair(float2 a, float2 ny, RowPsi rp) {
  const float2 X = f2add(a, f2(5.0f));
  const float2 w = f2fma(a, X, f2(5.0f));
  const float2 P = f2fma(w, w, f2(-1.0f));
  const float2 ndP = f2mul(w, f2fma(a, f2(-4.0f), f2(-10.0f)));   // -dP/da = -2w(2a + 5)
  // 1/X and 1/P from two MUFU.RCP: a MUFU costs one dispatch slot, the shared rcp(X P) cost one plus three FMULs
  const float2 rX = make_float2(fast_rcp(X.x), fast_rcp(X.y));
  const float2 rP = make_float2(fast_rcp(P.x), fast_rcp(P.y));
  float2 m, e23;
  split_exponent(X.x, m.x, e23.x);
  split_exponent(X.y, m.y, e23.y);
  const float2 f = f2add(m, f2(-1.0f));                            // [-0.293, 0.414]
  float2 lp = f2fma(f, f2(0.08671870082616806f), f2(-0.14378608763217926f));
  lp = f2fma(f, lp, f2(0.14977926015853882f));
  lp = f2fma(f, lp, f2(-0.16564473509788513f));
  lp = f2fma(f, lp, f2(0.1995488703250885f));
  lp = f2fma(f, lp, f2(-0.250016987323761f));
  lp = f2fma(f, lp, f2(0.33334165811538696f));
  const float2 lnm = f2fma(f2mul(f, f), f2fma(f, lp, f2(-0.5f)), f);   // ln m = f - f^2/2 + f^3 p(f)
  const float2 lnX = f2fma(e23, f2(kLn2 / 8388608.0f), lnm);                          // e ln2 + ln m
  const float2 lnXs = f2fma(f2add(e23, f2(-rp.k23)), f2(kLn2 / 8388608.0f), lnm);     // (e - k) ln2 + ln m
  const float2 LP = make_float2(fast_lg2(P.x), fast_lg2(P.y));
  const float2 z = f2mul(rX, rX);
  float2 nsp = f2fma(z, f2(-1.0f / 252.0f), f2(1.0f / 120.0f));    // -S_psi / z = -1/12 + z/120 - z^2/252
  nsp = f2fma(z, nsp, f2(-1.0f / 12.0f));
  float2 nsg2 = f2fma(z, f2(-2.0f / 1260.0f), f2(2.0f / 360.0f));  // -2 S_gam / rX = 2(-1/12 + z/360 - z^2/1260)
  nsg2 = f2fma(z, nsg2, f2(-2.0f / 12.0f));
  // nEd = -E - dpsi: the row constant rides along for free in the innermost fma
  const float2 nEd = f2fma(ndP, rP, f2fma(f2(-0.5f), rX, f2fma(nsp, z, f2(-rp.dpsi))));
  const float2 a2 = f2add(a, a);
  float2 M = f2fma(nsg2, rX, f2(2.0f * (5.0f - kHalfLn2Pi)));
  M = f2fma(a2, f2(1.0f + rp.dpsi), f2fma(a2, nEd, M));            // 2a (-E) + 2a
  M = f2fma(lnX, f2(-9.0f), M);
  M = f2fma(LP, f2(2.0f * kLn2), M);
  bool any_small = kSmall == 1;
  if (kSmall == 2) {
#if defined(__CUDA_ARCH__)
    any_small = __any_sync(0xffffffffu, (a.x < kSmallA) | (a.y < kSmallA));
#else
    any_small = (a.x < kSmallA) | (a.y < kSmallA);
#endif
  }
  if (any_small) {
    // Taylor form of 2N for a < 1/16: 2 a^2 (c2 + c3 a + ... + c7 a^5), truncation < 1e-7 relative (warp-uniform branch:
    // with y >= log(1e-15) the fixed points sit above 1/35, so whole warps rarely come here)
    float2 ts = f2fma(a, f2(2.0f * -0.864299380613076709f), f2(2.0f * 0.847785884987040950f));
    ts = f2fma(a, ts, f2(2.0f * -0.829542204114695941f));
    ts = f2fma(a, ts, f2(2.0f * 0.811742425283353644f));
    ts = f2fma(a, ts, f2(2.0f * -0.801371268773062857f));
    ts = f2fma(a, ts, f2(2.0f * 0.822467033424113218f));
    ts = f2mul(ts, f2mul(a, a));
    M.x = a.x < kSmallA ? ts.x : M.x;
    M.y = a.y < kSmallA ? ts.y : M.y;
  }
  const float2 g = f2add(f2add(lnXs, nEd), ny);                    // psi(a+1) - psi(s) - y
  const float2 nM = make_float2(-M.x, -M.y);
  const float2 bt = f2fma(a, g, nM);
  const float2 M4 = f2mul(M, f2(4.0f));
  const float2 Dt = f2fma(bt, bt, M4);
  const float2 r = make_float2(fast_sqrt(Dt.x), fast_sqrt(Dt.y));
  const float2 q = make_float2(fabsf(bt.x) + r.x, fabsf(bt.y) + r.y);
  const bool px = bt.x >= 0.0f, py = bt.y >= 0.0f;
  const float2 num = f2mul(a2, make_float2(px ? 1.0f : q.x, py ? 1.0f : q.y));
  const float2 den = make_float2(px ? q.x : M4.x, py ? q.y : M4.y);
  return f2mul(num, make_float2(fast_rcp(den.x), fast_rcp(den.y)));
}

// The same update cut in two at the point where the row constant enters, for the latency-bound few-rows kernel
// (mm_spec_kernel): mm_update_pre needs the pair only, so it is issued in the shadow of the row-total reduction and of
// psi(s); mm_update_post is the short remainder.  Operation for operation the arithmetic of mm_update_pair (checked bit
// for bit by tests/test_host_twin.py); the Taylor form for small a is evaluated unconditionally (no vote, no branch:
// issue slots are free here).
struct PairPre {
  float2 ndP, rP, rX, z, nsp, M0, lnX, lnm, e23, LP, ts;
};

TCLIP_HD PairPre mm_update_pre(float2 a) {
  PairPre p;
  const float2 X = f2add(a, f2(5.0f));
  const float2 w = f2fma(a, X, f2(5.0f));
  const float2 P = f2fma(w, w, f2(-1.0f));
  p.ndP = f2mul(w, f2fma(a, f2(-4.0f), f2(-10.0f)));
  p.rX = make_float2(fast_rcp(X.x), fast_rcp(X.y));
  p.rP = make_float2(fast_rcp(P.x), fast_rcp(P.y));
  float2 m;
  split_exponent(X.x, m.x, p.e23.x);
  split_exponent(X.y, m.y, p.e23.y);
  const float2 f = f2add(m, f2(-1.0f));
  float2 lp = f2fma(f, f2(0.08671870082616806f), f2(-0.14378608763217926f));
  lp = f2fma(f, lp, f2(0.14977926015853882f));
  lp = f2fma(f, lp, f2(-0.16564473509788513f));
  lp = f2fma(f, lp, f2(0.1995488703250885f));
  lp = f2fma(f, lp, f2(-0.250016987323761f));
  lp = f2fma(f, lp, f2(0.33334165811538696f));
  p.lnm = f2fma(f2mul(f, f), f2fma(f, lp, f2(-0.5f)), f);
  p.lnX = f2fma(p.e23, f2(kLn2 / 8388608.0f), p.lnm);
  p.LP = make_float2(fast_lg2(P.x), fast_lg2(P.y));
  p.z = f2mul(p.rX, p.rX);
  float2 nsp = f2fma(p.z, f2(-1.0f / 252.0f), f2(1.0f / 120.0f));
  p.nsp = f2fma(p.z, nsp, f2(-1.0f / 12.0f));
  float2 nsg2 = f2fma(p.z, f2(-2.0f / 1260.0f), f2(2.0f / 360.0f));
  nsg2 = f2fma(p.z, nsg2, f2(-2.0f / 12.0f));
  p.M0 = f2fma(nsg2, p.rX, f2(2.0f * (5.0f - kHalfLn2Pi)));
  float2 ts = f2fma(a, f2(2.0f * -0.864299380613076709f), f2(2.0f * 0.847785884987040950f));
  ts = f2fma(a, ts, f2(2.0f * -0.829542204114695941f));
  ts = f2fma(a, ts, f2(2.0f * 0.811742425283353644f));
  ts = f2fma(a, ts, f2(2.0f * -0.801371268773062857f));
  ts = f2fma(a, ts, f2(2.0f * 0.822467033424113218f));
  p.ts = f2mul(ts, f2mul(a, a));
  return p;
}

TCLIP_HD float2 mm_update_post(const PairPre& p, float2 a, float2 ny, RowPsi rp) {
  const float2 nEd = f2fma(p.ndP, p.rP, f2fma(f2(-0.5f), p.rX, f2fma(p.nsp, p.z, f2(-rp.dpsi))));
  const float2 a2 = f2add(a, a);
  float2 M = f2fma(a2, f2(1.0f + rp.dpsi), f2fma(a2, nEd, p.M0));
  M = f2fma(p.lnX, f2(-9.0f), M);
  M = f2fma(p.LP, f2(2.0f * kLn2), M);
  M.x = a.x < kSmallA ? p.ts.x : M.x;
  M.y = a.y < kSmallA ? p.ts.y : M.y;
  const float2 lnXs = f2fma(f2add(p.e23, f2(-rp.k23)), f2(kLn2 / 8388608.0f), p.lnm);
  const float2 g = f2add(f2add(lnXs, nEd), ny);
  const float2 nM = make_float2(-M.x, -M.y);
  const float2 bt = f2fma(a, g, nM);
  const float2 M4 = f2mul(M, f2(4.0f));
  const float2 Dt = f2fma(bt, bt, M4);
  const float2 r = make_float2(fast_sqrt(Dt.x), fast_sqrt(Dt.y));
  const float2 q = make_float2(fabsf(bt.x) + r.x, fabsf(bt.y) + r.y);
  const bool px = bt.x >= 0.0f, py = bt.y >= 0.0f;
  const float2 num = f2mul(a2, make_float2(px ? 1.0f : q.x, py ? 1.0f : q.y));
  const float2 den = make_float2(px ? q.x : M4.x, py ? q.y : M4.y);
  return f2mul(num, make_float2(fast_rcp(den.x), fast_rcp(den.y)));
}

// psi(s) for the row total, s > 0, to ~1e-10 absolute: ln s from the (m-1)/(m+1) series in float64, reciprocals from a
// MUFU seed + one Newton step (no division subroutine, no libm call).
TCLIP_HD double rcp_f64(double x) {
  double r = (double)fast_rcp((float)x);   // 1e-7 relative (needs |x| inside the float32 range: row totals are)
  r = r * fma(-x, r, 2.0);                 // one Newton step: ~1e-14
  return r;
}

// ln(s) = e ln2 + lnm with s = 2^e m, m in [sqrt(1/2), sqrt(2)); lnm = 2 atanh((m-1)/(m+1)) from its series up to t^13
// (t^2 <= 0.0295: truncation 5e-11), evaluated in Estrin form: this sits on the serial path of every MM iteration.
struct LogParts {
  int e;
  double lnm;
};

TCLIP_HD LogParts log_parts_f64(double s) {
#if defined(__CUDA_ARCH__)
  int hi = __double2hiint(s);
  const int lo = __double2loint(s);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = __hiloint2double(hi, lo);   // [1, 2)
#else
  int e;
  double m = std::frexp(s, &e) * 2.0;    // [1, 2)
  e -= 1;
#endif
  const bool up = m > 1.4142135623730951;
  m = up ? m * 0.5 : m;
  e = up ? e + 1 : e;
  const double t = (m - 1.0) * rcp_f64(m + 1.0);
  const double t2 = t * t;
  const double t4 = t2 * t2;
  const double t8 = t4 * t4;
  const double p01 = fma(t2, 1.0 / 3.0, 1.0);
  const double p23 = fma(t2, 1.0 / 7.0, 1.0 / 5.0);
  const double p45 = fma(t2, 1.0 / 11.0, 1.0 / 9.0);
  const double p = fma(t8, fma(t4, 1.0 / 13.0, p45), fma(t4, p23, p01));
  LogParts out;
  out.e = e;
  out.lnm = (t + t) * p;
  return out;
}

// psi(s) for the row total s > 0 (normal float64), split as k ln2 + dpsi with 2^k ~ s; dpsi still in float64 here.
struct RowPsiD {
  double dpsi;
  float k23;
};

TCLIP_HD RowPsiD row_psi_f64(double s) {
  double acc = 0.0;
  double x = s;
  while (x < 10.0) {  // only rows with a tiny total mass take this path
    acc -= rcp_f64(x);
    x += 1.0;
  }
  const LogParts lx = log_parts_f64(x);  // one logarithm only: it sits on the serial path of every MM iteration
  int k = lx.e;                          // 2^k ~ s
  if (x != s) {
#if defined(__CUDA_ARCH__)
    k = ((__double2hiint(s) >> 20) & 0x7ff) - 1023;
#else
    int e;
    std::frexp(s, &e);
    k = e - 1;
#endif
  }
  k = k < -120 ? -120 : (k > 120 ? 120 : k);           // s is a sum of float32 values; keeps k * 2^23 exact
  const double r = rcp_f64(x);
  const double r2 = r * r;
  double ser = fma(r2, 1.0 / 240.0, -1.0 / 252.0);   // next term 1/(132 x^10) < 1e-12
  ser = fma(r2, ser, 1.0 / 120.0);
  ser = fma(r2, ser, -1.0 / 12.0);
  const double tail = fma(r2, ser, -0.5 * r) + acc;   // psi(x) - ln x + acc
  RowPsiD out;
  out.dpsi = fma((double)(lx.e - k), 0.693147180559945309417, lx.lnm) + tail;
  out.k23 = (float)k * 8388608.0f;
  return out;
}

TCLIP_HD RowPsi row_psi(double s) {
  const RowPsiD d = row_psi_f64(s);
  RowPsi out;
  out.dpsi = (float)d.dpsi;
  out.k23 = d.k23;
  return out;
}

// The same through a fourth-order Taylor expansion around an anchor total, for the latency-bound few-rows kernel: between
// two MM iterations of a row the total moves by 1e-3 .. 1e-7 of itself, so with d = s - s_a
//   psi(s) = psi(s_a) + d psi^(1)(s_a) + d^2 psi^(2)(s_a)/2 + d^3 psi^(3)(s_a)/6 + d^4 psi^(4)(s_a)/24
// replaces the float64 logarithm (a ~20-operation dependent chain) by five fused multiply-adds while |d| <= s_a / 64
// (remainder (d/s)^5 / 5 < 2e-10, far below the float32 ulp of dpsi, 3e-8); outside that window, and for s_a < 16 (where
// the asymptotic series of the derivatives would need more terms), the full evaluation runs and becomes the new anchor.
// The expansion is always taken from the anchor, never chained, so nothing accumulates.  The result can differ from
// row_psi(s) in the last float32 bit of dpsi (tests/test_math_host.py bounds it); the kernels that prove periodicity from
// the row state alone (mm_chunk_kernel) keep the stateless row_psi.
struct PsiAnchor {      // plain data: it also lives in shared memory (one per warp in mm_chunk_kernel)
  double s;             // 0: no anchor yet
  double dpsi;          // psi(s) - k ln2
  double c1, c2, c3, c4;  // derivatives 1..4 of psi at s, divided by 1, 2, 6, 24
  float k23;
};

TCLIP_HD void psi_anchor_reset(PsiAnchor& an) {
  an.s = 0.0;
  an.dpsi = 0.0;
  an.c1 = an.c2 = an.c3 = an.c4 = 0.0;
  an.k23 = 0.0f;
}

TCLIP_HD RowPsi row_psi_anchored(double s, PsiAnchor& an) {
  const double d = s - an.s;
  RowPsi out;
  if (an.s >= 16.0 && fabs(d) <= an.s * (1.0 / 64.0)) {
    double p = fma(d, an.c4, an.c3);
    p = fma(d, p, an.c2);
    p = fma(d, p, an.c1);
    out.dpsi = (float)fma(d, p, an.dpsi);
    out.k23 = an.k23;
    return out;
  }
  const RowPsiD f = row_psi_f64(s);
  an.s = s;
  an.dpsi = f.dpsi;
  an.k23 = f.k23;
  if (s >= 16.0) {
    // derivatives of psi from its asymptotic series (next terms: 1/(42 s^7) in the first one, < 1e-10 at s = 16)
    const double r = rcp_f64(s);
    const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2;
    an.c1 = r + fma(r2, 0.5, fma(r3, 1.0 / 6.0, -(r4 * r) * (1.0 / 30.0)));
    an.c2 = 0.5 * (-(r2 + r3) + fma(r4, -0.5, (r3 * r3) * (1.0 / 6.0)));
    an.c3 = (fma(r3, 2.0, fma(r4, 3.0, (r4 * r) * 2.0)) - r4 * r3) * (1.0 / 6.0);
    an.c4 = (fma(r4, -6.0, fma(r4 * r, -12.0, (r3 * r3) * -10.0)) + 7.0 * (r4 * r4)) * (1.0 / 24.0);
  }
  out.dpsi = (float)f.dpsi;
  out.k23 = f.k23;
  return out;
}

// psi(s) itself in float64 (host tests)
TCLIP_HD double digamma_row(double s) {
  const RowPsi rp = row_psi(s);
  return (double)rp.dpsi + (double)rp.k23 / 8388608.0 * 0.693147180559945309417;
}

}  // namespace tclip
