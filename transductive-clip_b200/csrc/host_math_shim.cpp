// TEST SHIM — exposes tclip_math.cuh to the CPU test-suite (tests/test_math_host.py) as plain C symbols.
// Not part of libtclip_b200.so; compiled on demand with g++.  MUFU approximations are libm calls here, so this
// checks the series / algebra, not the hardware approximations (those are covered by the -m gpu parity tests).
#define TCLIP_HOST_MATH 1
#include <math.h>
#include "tclip_math.cuh"

extern "C" {
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
double tclip_host_digamma(double s) { return tclip::digamma_f64(s); }
// rows x D MM iterations on the host: the CPU twin of the kernel's inner loop (row sum in double).
void tclip_host_mm_rows(float* alpha, const float* y, int rows, int D, int iters) {
  for (int r = 0; r < rows; ++r) {
    float* a = alpha + (long)r * D;
    const float* yy = y + (long)r * D;
    for (int it = 0; it < iters; ++it) {
      double s = 0.0;
      for (int d = 0; d < D; ++d) s += (double)a[d];
      const double ps = tclip::digamma_f64(s);
      const float hi = (float)ps;
      const float lo = (float)(ps - (double)hi);
      for (int d = 0; d < D; ++d) a[d] = tclip::mm_update_element(a[d], yy[d], hi, lo);
    }
  }
}
}
