// TEST SHIM — exposes tclip_math.cuh to the CPU test-suite (tests/test_math_host.py) as plain C symbols.
// Not part of libtclip_b200.so; compiled on demand with g++.  MUFU rcp / sqrt are correctly rounded libm calls here and MUFU lg2 is
// modelled after its measured error (tclip_math.cuh), so this checks the series / algebra and the sensitivity to the lg2
// truncation; the hardware itself is covered by the -m gpu parity tests.
#define TCLIP_HOST_MATH 1
#include <math.h>
#include "tclip_math.cuh"

extern "C" {
void tclip_host_set_lg2_exact(int exact) { tclip::host_lg2_exact() = exact; }
void tclip_host_psi1_N(const float* a, float* psi1, float* N, int n) {
  for (int i = 0; i < n; ++i) {
    tclip::PsiN r = tclip::psi1_and_curvature_num(a[i]);
    psi1[i] = r.psi1;
    N[i] = r.N;
  }
}
void tclip_host_mm_update(const float* a, const float* y, float* out, int n, double psis) {
  const float hi = (float)psis;
  const float lo = (float)(psis - (double)hi);
  for (int i = 0; i < n; ++i) out[i] = tclip::mm_update_element(a[i], y[i], hi, lo);
}
double tclip_host_digamma(double s) { return tclip::digamma_row(s); }
// `s` is the row total (the kernel derives psi(s) from it itself)
void tclip_host_mm_update_pair(const float* a, const float* y, float* out, int n, double s) {
  const tclip::RowPsi rp = tclip::row_psi(s);
  for (int i = 0; i + 1 < n; i += 2) {
    tclip::float2 r = tclip::mm_update_pair(tclip::make_float2(a[i], a[i + 1]), tclip::make_float2(-y[i], -y[i + 1]), rp);
    out[i] = r.x;
    out[i + 1] = r.y;
  }
}
// the two-phase form used by mm_spec_kernel (must equal mm_update_pair bit for bit)
void tclip_host_mm_update_pair_split(const float* a, const float* y, float* out, int n, double s) {
  const tclip::RowPsi rp = tclip::row_psi(s);
  for (int i = 0; i + 1 < n; i += 2) {
    const tclip::float2 av = tclip::make_float2(a[i], a[i + 1]);
    const tclip::PairPre pre = tclip::mm_update_pre(av);
    tclip::float2 r = tclip::mm_update_post(pre, av, tclip::make_float2(-y[i], -y[i + 1]), rp);
    out[i] = r.x;
    out[i + 1] = r.y;
  }
}
// row_psi_anchored along a walk of row totals: out_dpsi[i], out_full[i] = dpsi of the anchored / the stateless evaluation,
// out_dk23 = difference of the two exponent splits; returns how many steps took the full evaluation (re-anchored)
int tclip_host_row_psi_walk(const double* s, float* out_dpsi, float* out_full, float* out_dk23, int n) {
  tclip::PsiAnchor an;
  tclip::psi_anchor_reset(an);
  int full = 0;
  for (int i = 0; i < n; ++i) {
    const double before = an.s;
    const tclip::RowPsi a = tclip::row_psi_anchored(s[i], an);
    if (an.s != before) ++full;
    const tclip::RowPsi f = tclip::row_psi(s[i]);
    out_dpsi[i] = a.dpsi;
    out_full[i] = f.dpsi;
    out_dk23[i] = a.k23 - f.k23;
  }
  return full;
}
// rows x D MM iterations on the host: the CPU twin of the kernel's inner loop (row sum in double).
// D = padded (even) row length, n_valid = real row length: like the kernel, padding never enters the row total.
void tclip_host_mm_rows(float* alpha, const float* y, int rows, int D, int n_valid, int iters) {
  for (int r = 0; r < rows; ++r) {
    float* a = alpha + (long)r * D;
    const float* yy = y + (long)r * D;
    for (int it = 0; it < iters; ++it) {
      double s = 0.0;
      for (int d = 0; d < n_valid; ++d) s += (double)a[d];
      const tclip::RowPsi rp = tclip::row_psi(s);
      for (int d = 0; d + 1 < D; d += 2) {  // the kernel's packed form (D even here)
        tclip::float2 r = tclip::mm_update_pair(tclip::make_float2(a[d], a[d + 1]), tclip::make_float2(-yy[d], -yy[d + 1]), rp);
        a[d] = r.x;
        a[d + 1] = r.y;
      }
    }
  }
}
// the same with psi(s) from the anchored expansion (what mm_spec_kernel and the dense kernel's non-free-running rows do);
// returns the number of full evaluations (re-anchorings) over all rows
int tclip_host_mm_rows_anchored(float* alpha, const float* y, int rows, int D, int n_valid, int iters) {
  int full = 0;
  for (int r = 0; r < rows; ++r) {
    float* a = alpha + (long)r * D;
    const float* yy = y + (long)r * D;
    tclip::PsiAnchor an;
    tclip::psi_anchor_reset(an);
    for (int it = 0; it < iters; ++it) {
      double s = 0.0;
      for (int d = 0; d < n_valid; ++d) s += (double)a[d];
      const double before = an.s;
      const tclip::RowPsi rp = tclip::row_psi_anchored(s, an);
      if (an.s != before) ++full;
      for (int d = 0; d + 1 < D; d += 2) {
        tclip::float2 r2 = tclip::mm_update_pair(tclip::make_float2(a[d], a[d + 1]), tclip::make_float2(-yy[d], -yy[d + 1]), rp);
        a[d] = r2.x;
        a[d + 1] = r2.y;
      }
    }
  }
  return full;
}
}
