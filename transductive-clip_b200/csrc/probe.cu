// probe.cu — register-only issue-rate microbenchmarks.  They give the roofline denominators of the MM M-step kernel
// (FP32 FMA pipe and MUFU pipe) at the clock the GPU actually sustains, measured next to the kernel in bench.py;
// MEASURED_PEAKS.json only carries HBM and bf16-tensor peaks, and this path is bound by neither.
#include <cuda_runtime.h>

#include "tclip_kernels.cuh"

namespace tclip {

namespace {

constexpr int kProbeThreads = 256;
constexpr int kProbeChains = 8;    // independent dependency chains per thread
constexpr int kProbeUnroll = 64;   // operations per chain and loop trip

// flop per CTA = 2 * kProbeThreads * kProbeChains * kProbeUnroll * iters
__global__ void __launch_bounds__(kProbeThreads) probe_ffma_kernel(float* sink, int iters) {
  float x[kProbeChains];
#pragma unroll
  for (int c = 0; c < kProbeChains; ++c) x[c] = 1.0f + 1e-3f * (float)(threadIdx.x + c);
  const float m = 0.999f + 1e-9f * (float)blockIdx.x, a = 1e-3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < kProbeUnroll; ++k)
#pragma unroll
      for (int c = 0; c < kProbeChains; ++c) x[c] = fmaf(x[c], m, a);
  }
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < kProbeChains; ++c) s += x[c];
  if (s == 123.456f) sink[0] = s;  // never true: keeps the chains alive without a store
}

// MUFU ops per CTA = kProbeThreads * kProbeChains * kProbeUnroll * iters  (alternating lg2 / rcp / sqrt like the M-step)
__global__ void __launch_bounds__(kProbeThreads) probe_mufu_kernel(float* sink, int iters) {
  float x[kProbeChains];
#pragma unroll
  for (int c = 0; c < kProbeChains; ++c) x[c] = 1.5f + 1e-3f * (float)(threadIdx.x + c);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < kProbeUnroll; ++k)
#pragma unroll
      for (int c = 0; c < kProbeChains; ++c) {
        float r;
        if ((k % 3) == 0) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[c]));
        else if ((k % 3) == 1) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[c]));
        else asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x[c]));
        x[c] = r;
      }
  }
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < kProbeChains; ++c) s += x[c];
  if (s == 123.456f) sink[0] = s;
}

// packed FFMA2: flop per CTA = 2 * 2 * kProbeThreads * kProbeChains * kProbeUnroll * iters
__global__ void __launch_bounds__(kProbeThreads) probe_ffma2_kernel(float* sink, int iters) {
  float2 x[kProbeChains];
#pragma unroll
  for (int c = 0; c < kProbeChains; ++c) x[c] = make_float2(1.0f + 1e-3f * (float)(threadIdx.x + c), 1.0f);
  const float2 m = make_float2(0.999f + 1e-9f * (float)blockIdx.x, 0.9999f), a = make_float2(1e-3f, 1e-4f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < kProbeUnroll; ++k)
#pragma unroll
      for (int c = 0; c < kProbeChains; ++c) x[c] = __ffma2_rn(x[c], m, a);
  }
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < kProbeChains; ++c) s += x[c].x + x[c].y;
  if (s == 123.456f) sink[0] = s;
}

// the M-step's own mix: per trip 4 FFMA2 + 1 MUFU per chain, to see whether the two pipes overlap
__global__ void __launch_bounds__(kProbeThreads) probe_mix_kernel(float* sink, int iters) {
  float2 x[kProbeChains];
  float u[kProbeChains];
#pragma unroll
  for (int c = 0; c < kProbeChains; ++c) {
    x[c] = make_float2(1.0f + 1e-3f * (float)(threadIdx.x + c), 1.0f);
    u[c] = 1.5f + 1e-3f * (float)c;
  }
  const float2 m = make_float2(0.999f + 1e-9f * (float)blockIdx.x, 0.9999f), a = make_float2(1e-3f, 1e-4f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < kProbeUnroll / 4; ++k)
#pragma unroll
      for (int c = 0; c < kProbeChains; ++c) {
        x[c] = __ffma2_rn(x[c], m, a);
        x[c] = __ffma2_rn(x[c], m, a);
        x[c] = __ffma2_rn(x[c], m, a);
        x[c] = __ffma2_rn(x[c], m, a);
        float r;
        asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u[c]));
        u[c] = r;
      }
  }
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < kProbeChains; ++c) s += x[c].x + x[c].y + u[c];
  if (s == 123.456f) sink[0] = s;
}

}  // namespace

cudaError_t probe_ffma2(float* sink, int n_blocks, int iters, cudaStream_t st) {
  probe_ffma2_kernel<<<n_blocks, kProbeThreads, 0, st>>>(sink, iters);
  note_launch();
  return cudaGetLastError();
}

cudaError_t probe_mix(float* sink, int n_blocks, int iters, cudaStream_t st) {
  probe_mix_kernel<<<n_blocks, kProbeThreads, 0, st>>>(sink, iters);
  note_launch();
  return cudaGetLastError();
}

cudaError_t probe_ffma(float* sink, int n_blocks, int iters, cudaStream_t st) {
  probe_ffma_kernel<<<n_blocks, kProbeThreads, 0, st>>>(sink, iters);
  note_launch();
  return cudaGetLastError();
}

cudaError_t probe_mufu(float* sink, int n_blocks, int iters, cudaStream_t st) {
  probe_mufu_kernel<<<n_blocks, kProbeThreads, 0, st>>>(sink, iters);
  note_launch();
  return cudaGetLastError();
}

}  // namespace tclip
