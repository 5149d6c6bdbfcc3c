// scratch: error of lg2.approx.ftz.f32 vs double log2 over several ranges
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const float* x, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[i])); out[i] = r; }
}
int main() {
  const int n = 1 << 20;
  float *hx = new float[n], *ho = new float[n], *dx, *dout;
  cudaMalloc(&dx, n * 4); cudaMalloc(&dout, n * 4);
  double ranges[][2] = {{0.5, 1.0}, {1.0, 2.0}, {0.9, 1.1}, {4.0, 8.0}, {5.0, 5.5}, {1e3, 2e3}, {1e5, 2e5}, {1e-4, 2e-4}, {1e19, 2e19}};
  for (auto& r : ranges) {
    for (int i = 0; i < n; ++i) hx[i] = (float)(r[0] + (r[1] - r[0]) * (i + 0.5) / n);
    cudaMemcpy(dx, hx, n * 4, cudaMemcpyHostToDevice);
    k<<<n / 256, 256>>>(dx, dout, n);
    cudaMemcpy(ho, dout, n * 4, cudaMemcpyDeviceToHost);
    double maxabs = 0, maxulp = 0, mean = 0;
    for (int i = 0; i < n; ++i) {
      double ref = log2((double)hx[i]);
      double e = (double)ho[i] - ref;
      double ulp = ldexp(1.0, ilogb(fabs(ref) > 0 ? fabs(ref) : 1e-30) - 23);
      maxabs = fmax(maxabs, fabs(e)); maxulp = fmax(maxulp, fabs(e) / ulp); mean += e;
    }
    printf("x in [%g, %g): max abs err %.3e  max err in ulp(result) %.2f  mean err %.3e\n", r[0], r[1], maxabs, maxulp, mean / n);
  }
  return 0;
}
