// scratch: dump err(m) = lg2.approx(m) - log2(m) for all float32 m in [sqrt(1/2), sqrt(2)) on a 2^18 grid + block means
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const float* x, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); out[i] = r; }
}
int main() {
  const int n = 1 << 22;
  float *hx = new float[n], *ho = new float[n], *dx, *dout;
  cudaMalloc(&dx, n * 4); cudaMalloc(&dout, n * 4);
  const double lo = sqrt(0.5), hi = sqrt(2.0);
  for (int i = 0; i < n; ++i) hx[i] = (float)(lo + (hi - lo) * (i + 0.5) / n);
  cudaMemcpy(dx, hx, n * 4, cudaMemcpyHostToDevice);
  k<<<n / 256, 256>>>(dx, dout, n);
  cudaMemcpy(ho, dout, n * 4, cudaMemcpyDeviceToHost);
  const int nb = 64;
  printf("# block  m_lo  mean_err  max_err  min_err  rms\n");
  double tot = 0;
  for (int b = 0; b < nb; ++b) {
    double mean = 0, mx = -1, mn = 1, sq = 0;
    for (int i = b * (n / nb); i < (b + 1) * (n / nb); ++i) {
      double e = (double)ho[i] - log2((double)hx[i]);
      mean += e; mx = fmax(mx, e); mn = fmin(mn, e); sq += e * e;
    }
    mean /= (n / nb); tot += mean;
    printf("%2d %.5f %+.3e %+.3e %+.3e %.3e\n", b, hx[b * (n / nb)], mean, mx, mn, sqrt(sq / (n / nb)));
  }
  printf("# overall mean %+.4e\n", tot / nb);
  // fine structure: first 64 consecutive floats above 1.0
  float x = 1.0f;
  for (int i = 0; i < 24; ++i) { hx[i] = x; x = nextafterf(x, 2.0f); }
  for (int i = 0; i < 24; ++i) hx[24 + i] = 1.25f + i * 1e-6f;
  cudaMemcpy(dx, hx, 48 * 4, cudaMemcpyHostToDevice);
  k<<<1, 64>>>(dx, dout, 48);
  cudaMemcpy(ho, dout, 48 * 4, cudaMemcpyDeviceToHost);
  for (int i = 0; i < 48; ++i) printf("x=%.9g lg2approx=%.9g exact=%.12g err=%+.3e\n", hx[i], ho[i], log2((double)hx[i]), (double)ho[i] - log2((double)hx[i]));
  return 0;
}
