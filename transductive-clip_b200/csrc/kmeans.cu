// kmeans.cu — soft k-means / hard k-means / EM-Gaussian (identity covariance) on sm_100a.
//
// Reference call sites (SegoleneMartin/transductive-CLIP, src/methods/zero_shot/):
//   initial assignment on visual features   soft_kmeans.py:185-197 (same in hard_kmeans.py:171-183, em_gaussian.py:188-203)
//                                           u[t] = softmax(T * normalize(x[t]) @ text^T)       -> normalize_rows + similarity
//   centroids w = u^T x / sum u             soft_kmeans.py:135-166; hard_kmeans.py:138-151 (empty clusters zeroed);
//                                           em_gaussian.py:138-169                              -> centroids_kernel
//   squared distances ||w_k - x_n||^2       soft_kmeans.py:105-114, hard_kmeans.py:26-35, em_gaussian.py:106-115
//                                                                                               -> sqdist_kernel
//   assignment                              soft_kmeans.py:116-125  u = softmax(T * (-1/2 d2))
//                                           em_gaussian.py:117-128  u = softmax(T * (-1/2 d2) + lambda v / n)
//                                           hard_kmeans.py:127-136,197-199  u = one-hot(argmin softmax(+d2))
//                                                                                               -> assign_kernel
//   logged criterion of hard k-means        hard_kmeans.py:201-203  mean_t ||u_old - u||_F      -> udiff_kernel
//   EM-Gaussian, diagonal covariance        em_gaussian_cov.py:106-130 (E-step), :172-193 (s_init / s_update)
//                                                                                               -> precisions_kernel, pair_kernel<2>
//   KL k-means                              kl_kmeans.py:123-127,166-177                        -> centroids (mode 2), pair_kernel<3>
//
// Layouts: x [T,n,D], u / d2 [T,n,K], w [T,K,D], text [K,D], all float32 row-major.  In w-space the loop is bound by the
// CUDA-core rate of the direct-difference distance (3 flop per (n,k,d), the reference's own formulation: no cancellation),
// then by HBM (w is written and read once per iteration, 4 MB per task at K=1000, D=1024).
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>
#include <cstdlib>
#include <string>

#include "tclip_kernels.cuh"

namespace tclip {

namespace {

constexpr float kEps = 1e-15f;

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// out[r,:] = x[r,:] / ||x[r,:]||  (one warp per row; a zero row gives NaN like the reference's division)
__global__ void __launch_bounds__(128) normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ out, long rows,
                                                             int D) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = x + row * D;
  float s = 0.0f;
  for (int d = lane; d < D; d += 32) s = fmaf(p[d], p[d], s);
  s = sqrtf(warp_sum_f32(s));
  for (int d = lane; d < D; d += 32) out[row * D + d] = p[d] / s;
}

// Tiled [M x D] . [N x D]^T with 64 x 64 tiles, BK = 16, 256 threads, 4 x 4 outputs per thread.
//   OP 0: sum_d a b                      (similarity; B shared by all batches when b_batch_stride == 0)
//   OP 1: sum_d (b - a)^2                (squared distance, the reference's direct form)
//   OP 2: sum_d s (b - a)^2              (diagonal-precision distance, S laid out like B)
//   OP 3: sum_d p log(p / q), p = a + eps, q = b + eps   (KL divergence of kl_kmeans.py:123-127, division first)
constexpr int kTile = 64;
constexpr int kBK = 16;

template <int OP>
__global__ void __launch_bounds__(256)
pair_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ S, float* __restrict__ C,
            int M, int N, int D, long a_batch_stride, long b_batch_stride, long c_batch_stride) {
  __shared__ float as[kBK][kTile + 4];
  __shared__ float bs[kBK][kTile + 4];
  __shared__ float ss[OP == 2 ? kBK : 1][kTile + 4];
  const int t = blockIdx.z;
  const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* ab = A + (long)t * a_batch_stride;
  const float* bb = B + (long)t * b_batch_stride;
  const float* sb = OP == 2 ? S + (long)t * b_batch_stride : nullptr;
  float acc[4][4] = {};
  for (int d0 = 0; d0 < D; d0 += kBK) {
    for (int i = threadIdx.x; i < kTile * kBK; i += 256) {
      const int r = i / kBK, c = i % kBK;
      const int d = d0 + c;
      const bool ina = m0 + r < M && d < D, inb = n0 + r < N && d < D;
      // OP 3 pads with p = q = 1 (p log(p/q) = 0), the others with zeros
      as[c][r] = ina ? ab[(long)(m0 + r) * D + d] : (OP == 3 ? 1.0f : 0.0f);
      bs[c][r] = inb ? bb[(long)(n0 + r) * D + d] : (OP == 3 ? 1.0f : 0.0f);
      if (OP == 2) ss[c][r] = inb ? sb[(long)(n0 + r) * D + d] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kBK; ++c) {
      float av[4], bv[4], sv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = as[c][ty * 4 + i];
        bv[i] = bs[c][tx * 4 + i];
        sv[i] = OP == 2 ? ss[c][tx * 4 + i] : 0.0f;
        if (OP == 3) {
          av[i] += kEps;
          bv[i] += kEps;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (OP == 0) {
            acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
          } else if (OP == 1) {
            const float df = bv[j] - av[i];
            acc[i][j] = fmaf(df, df, acc[i][j]);
          } else if (OP == 2) {
            const float df = bv[j] - av[i];
            acc[i][j] = fmaf(df * df, sv[j], acc[i][j]);
          } else {
            acc[i][j] = fmaf(av[i], logf(av[i] / bv[j]), acc[i][j]);
          }
        }
    }
    __syncthreads();
  }
  float* cb = C + (long)t * c_batch_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn < N) cb[(long)m * N + nn] = acc[i][j];
    }
  }
}

// w[t,k,:] = sum_n u[t,n,k] x[t,n,:] / max(sum_n u, eps) for non-empty clusters; empty ones are zeroed (mode 0) or keep
// the previous row (mode 1).  mode 2 = KL k-means (kl_kmeans.py:166-171): divide by max(size, 1), zero iff size == 0.
// 64(k) x 64(d) tile per CTA, n staged through shared memory; the column sums of u come along.
// precisions_kernel (below) has the same shape for s[t,k,:] = sum_n u / max(sum_n u (w - x)^2, eps).
constexpr int kStage = 16;

__global__ void __launch_bounds__(256)
centroids_kernel(const float* __restrict__ u, const float* __restrict__ x, float* __restrict__ w, int n, int K, int D,
                 int mode) {
  __shared__ float us[kStage][kTile + 4];
  __shared__ float xs[kStage][kTile + 4];
  __shared__ float csum[kTile];
  const int t = blockIdx.z;
  const int k0 = blockIdx.y * kTile, d0 = blockIdx.x * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* ub = u + (long)t * n * K;
  const float* xb = x + (long)t * n * D;
  if (threadIdx.x < kTile) {  // cluster sizes in query order, like u.sum(1)
    float s = 0.0f;
    const int k = k0 + threadIdx.x;
    if (k < K)
      for (int i = 0; i < n; ++i) s += ub[(long)i * K + k];
    csum[threadIdx.x] = s;
  }
  float acc[4][4] = {};
  for (int n0 = 0; n0 < n; n0 += kStage) {
    for (int i = threadIdx.x; i < kStage * kTile; i += 256) {
      const int r = i / kTile, c = i % kTile;
      const int nn = n0 + r;
      us[r][c] = (nn < n && k0 + c < K) ? ub[(long)nn * K + k0 + c] : 0.0f;
      xs[r][c] = (nn < n && d0 + c < D) ? xb[(long)nn * D + d0 + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kStage; ++r) {
      float uu[4], xx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uu[i] = us[r][ty * 4 + i];
        xx[i] = xs[r][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(uu[i], xx[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= K) continue;
    const float cs = csum[ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + tx * 4 + j;
      if (d >= D) continue;
      const long o = ((long)t * K + k) * D + d;
      if (mode == 2) w[o] = cs > 0.0f ? acc[i][j] / fmaxf(cs, 1.0f) : 0.0f;
      else if (cs > kEps) w[o] = acc[i][j] / fmaxf(cs, kEps);
      else if (mode == 0) w[o] = 0.0f;
    }
  }
}

// s[t,k,d] = sum_n u[t,n,k] / max(sum_n u[t,n,k] (w[t,k,d] - x[t,n,d])^2, eps); empty clusters keep their row when
// keep_old (s_update, em_gaussian_cov.py:182-193), s_init (:172-180) has no mask (0 / eps = 0 for an empty cluster).
__global__ void __launch_bounds__(256)
precisions_kernel(const float* __restrict__ u, const float* __restrict__ x, const float* __restrict__ w,
                  float* __restrict__ s, int n, int K, int D, int keep_old) {
  __shared__ float us[kStage][kTile + 4];
  __shared__ float xs[kStage][kTile + 4];
  __shared__ float csum[kTile];
  const int t = blockIdx.z;
  const int k0 = blockIdx.y * kTile, d0 = blockIdx.x * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* ub = u + (long)t * n * K;
  const float* xb = x + (long)t * n * D;
  if (threadIdx.x < kTile) {
    float sum = 0.0f;
    const int k = k0 + threadIdx.x;
    if (k < K)
      for (int i = 0; i < n; ++i) sum += ub[(long)i * K + k];
    csum[threadIdx.x] = sum;
  }
  float wv[4][4], acc[4][4] = {};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + ty * 4 + i, d = d0 + tx * 4 + j;
      wv[i][j] = (k < K && d < D) ? w[((long)t * K + k) * D + d] : 0.0f;
    }
  for (int n0 = 0; n0 < n; n0 += kStage) {
    for (int i = threadIdx.x; i < kStage * kTile; i += 256) {
      const int r = i / kTile, c = i % kTile;
      const int nn = n0 + r;
      us[r][c] = (nn < n && k0 + c < K) ? ub[(long)nn * K + k0 + c] : 0.0f;
      xs[r][c] = (nn < n && d0 + c < D) ? xb[(long)nn * D + d0 + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kStage; ++r) {
      float uu[4], xx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uu[i] = us[r][ty * 4 + i];
        xx[i] = xs[r][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float df = wv[i][j] - xx[j];
          acc[i][j] = fmaf(df * df, uu[i], acc[i][j]);   // padded n have u = 0
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= K) continue;
    const float cs = csum[ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + tx * 4 + j;
      if (d >= D) continue;
      const long o = ((long)t * K + k) * D + d;
      if (!keep_old || cs > kEps) s[o] = cs / fmaxf(acc[i][j], kEps);
    }
  }
}

// det[row] = 1/2 sum_d log(s[row,d] + eps)  (one warp per (task, class) row; em_gaussian_cov.py:127)
__global__ void __launch_bounds__(128) half_logdet_kernel(const float* __restrict__ s, float* __restrict__ det, long rows,
                                                          int D) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = s + row * D;
  float acc = 0.0f;
  for (int d = lane; d < D; d += 32) acc += logf(p[d] + kEps);
  acc = warp_sum_f32(acc);
  if (lane == 0) det[row] = 0.5f * acc;
}

// One warp per (task, query): logits from the squared distances, soft-max over the classes, optional one-hot.
//   mode 0 (soft k-means)  u = softmax(T * (-1/2 d2))
//   mode 1 (EM-Gaussian)   u = softmax(T * (-1/2 d2) + lambd * v / n)
//   mode 2 (hard k-means)  u = one-hot(argmin_k softmax(+d2)), lowest index on ties, as torch.argmin of the soft-maxed values
//   mode 3 (similarity)    u = softmax(T * s)   (initial assignment / prototype probabilities on visual features)
//   mode 4 (EM-Gaussian, diagonal covariance)  u = softmax(-1/2 d2s + det + lambd * v / n)   (em_gaussian_cov.py:117-130)
//   mode 5 (KL k-means)    u = one-hot(argmin_k div), NaN counts as the minimum like torch.argmin (kl_kmeans.py:174-177)
// `u` may alias `d2`.  labels (optional) = argmax_k of the final u.
// The logit of one (query, class) pair, every operation rounded on its own like the reference's tensor expressions (no fused
// multiply-add across them: the two assignment kernels must agree bit for bit, and contraction is the compiler's choice)
__device__ __forceinline__ float assign_logit(int mode, float d, float temperature, float lambd, float vk, float bk, float fn) {
  if (mode == 2) return d;
  if (mode == 3) return __fmul_rn(temperature, d);
  if (mode == 4) return __fadd_rn(__fadd_rn(__fmul_rn(-0.5f, d), bk), __fdiv_rn(__fmul_rn(lambd, vk), fn));
  float l = __fmul_rn(temperature, __fmul_rn(-0.5f, d));
  if (mode == 1) l = __fadd_rn(l, __fdiv_rn(__fmul_rn(lambd, vk), fn));
  return l;
}

__global__ void __launch_bounds__(128)
assign_kernel(const float* d2, const float* __restrict__ v, const float* __restrict__ bias, float temperature, float lambd,
              float* u, int* __restrict__ labels, int rows, int n, int K, int mode, double* __restrict__ row_sq) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int t = row / n;
  const float* x = d2 + (long)row * K;
  const float* vv = v ? v + (long)t * K : nullptr;
  const float* bb = bias ? bias + (long)t * K : nullptr;
  float* out = u + (long)row * K;
  if (mode == 5) {  // plain arg-min of the divergences, NaN first
    float best = CUDART_INF_F;
    int best_k = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
      float d = x[k];
      d = (d != d) ? -CUDART_INF_F : d;
      if (d < best) {
        best = d;
        best_k = k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
      if (ob < best || (ob == best && ok < best_k)) {
        best = ob;
        best_k = ok;
      }
    }
    if (best_k == 0x7fffffff) best_k = 0;  // every divergence +inf: torch.argmin returns the first index
    for (int k = lane; k < K; k += 32) out[k] = (k == best_k) ? 1.0f : 0.0f;
    if (lane == 0 && labels) labels[row] = best_k;
    return;
  }
  const float fn = (float)n;
  auto logit = [&](int k) -> float {
    return assign_logit(mode, x[k], temperature, lambd, vv ? vv[k] : 0.0f, bb ? bb[k] : 0.0f, fn);
  };
  float mx = -CUDART_INF_F;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, logit(k));
  mx = warp_max_f32(mx);
  float sum = 0.0f;
  for (int k = lane; k < K; k += 32) sum += expf(__fsub_rn(logit(k), mx));
  sum = warp_sum_f32(sum);
  // arg-extremum of the soft-maxed values: max for the soft variants (label output), min for hard k-means
  float best = mode == 2 ? CUDART_INF_F : -1.0f;
  int best_k = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float p = __fdiv_rn(expf(__fsub_rn(logit(k), mx)), sum);
    const bool better = mode == 2 ? p < best : p > best;  // strict: the lowest k of this lane's stripe wins ties
    if (better) {
      best = p;
      best_k = k;
    }
    if (mode != 2) out[k] = p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    const bool better = mode == 2 ? ob < best : ob > best;
    if (better || (ob == best && ok < best_k)) {
      best = ob;
      best_k = ok;
    }
  }
  if (mode == 2) {
    if (row_sq) {   // || u_old - u ||^2 of this row while the old row is replaced (hard_kmeans.py:201: the logged criterion)
      // float32 per row (<= 1000 terms of a row that sums to <= 2; exact when both rows are one-hot), float64 across the
      // rows and tasks: the float64 units are too slow for a per-element term (measured: 32 -> 102 us per launch)
      float s = 0.0f;
      for (int k = lane; k < K; k += 32) {
        const float nw = (k == best_k) ? 1.0f : 0.0f;
        const float d = out[k] - nw;
        s = fmaf(d, d, s);
        out[k] = nw;
      }
      s = warp_sum_f32(s);
      if (lane == 0) row_sq[row] = (double)s;
    } else {
      for (int k = lane; k < K; k += 32) out[k] = (k == best_k) ? 1.0f : 0.0f;
    }
  }
  if (lane == 0 && labels) labels[row] = best_k;
}

// The same rows with the NV = ceil(K / 32) logits of a lane kept in registers (K <= 1024; modes 0-4): d2 is read once instead
// of three times.  Every sum, maximum and arg-extremum is taken in the order of assign_kernel (a lane's classes ascending,
// then the shuffle tree), so the two kernels agree bit for bit.
template <int NV>
__global__ void __launch_bounds__(128)
assign_reg_kernel(const float* d2, const float* __restrict__ v, const float* __restrict__ bias, float temperature, float lambd,
                  float* u, int* __restrict__ labels, int rows, int n, int K, int mode, double* __restrict__ row_sq) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int t = row / n;
  const float* x = d2 + (long)row * K;
  const float* vv = v ? v + (long)t * K : nullptr;
  const float* bb = bias ? bias + (long)t * K : nullptr;
  float* out = u + (long)row * K;
  const float fn = (float)n;
  float l[NV];
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int k = lane + 32 * j;
    float lg = -CUDART_INF_F;
    if (k < K) {
      lg = assign_logit(mode, x[k], temperature, lambd, vv ? vv[k] : 0.0f, bb ? bb[k] : 0.0f, fn);
      mx = fmaxf(mx, lg);
    }
    l[j] = lg;
  }
  mx = warp_max_f32(mx);
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (lane + 32 * j < K) {
      l[j] = expf(__fsub_rn(l[j], mx));
      sum += l[j];
    }
  }
  sum = warp_sum_f32(sum);
  float best = mode == 2 ? CUDART_INF_F : -1.0f;
  int best_k = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int k = lane + 32 * j;
    if (k < K) {
      const float p = __fdiv_rn(l[j], sum);
      const bool better = mode == 2 ? p < best : p > best;  // strict: the lowest k of this lane's stripe wins ties
      if (better) {
        best = p;
        best_k = k;
      }
      if (mode != 2) out[k] = p;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    const bool better = mode == 2 ? ob < best : ob > best;
    if (better || (ob == best && ok < best_k)) {
      best = ob;
      best_k = ok;
    }
  }
  if (mode == 2) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int k = lane + 32 * j;
      if (k < K) {
        const float nw = (k == best_k) ? 1.0f : 0.0f;
        if (row_sq) {
          const float d = out[k] - nw;
          s = fmaf(d, d, s);
        }
        out[k] = nw;
      }
    }
    if (row_sq) {
      s = warp_sum_f32(s);
      if (lane == 0) row_sq[row] = (double)s;
    }
  }
  if (lane == 0 && labels) labels[row] = best_k;
}

void launch_assign(const float* d2, const float* v, const float* bias, float temperature, float lambd, float* u, int* labels,
                   int rows, int n, int K, int mode, double* row_sq, cudaStream_t st) {
  const unsigned grid = (unsigned)((rows + 3) / 4);
  static const bool generic = [] {   // TCLIP_ASSIGN=generic: the any-K kernel for every shape (cross-check of the two forms)
    const char* e = std::getenv("TCLIP_ASSIGN");
    return e && std::string(e) == "generic";
  }();
  if (generic || mode == 5 || K > 1024)
    assign_kernel<<<grid, 128, 0, st>>>(d2, v, bias, temperature, lambd, u, labels, rows, n, K, mode, row_sq);
  else if (K <= 128)
    assign_reg_kernel<4><<<grid, 128, 0, st>>>(d2, v, bias, temperature, lambd, u, labels, rows, n, K, mode, row_sq);
  else if (K <= 512)
    assign_reg_kernel<16><<<grid, 128, 0, st>>>(d2, v, bias, temperature, lambd, u, labels, rows, n, K, mode, row_sq);
  else
    assign_reg_kernel<32><<<grid, 128, 0, st>>>(d2, v, bias, temperature, lambd, u, labels, rows, n, K, mode, row_sq);
}

// task_norm[t] = ||a[t] - b[t]||_F (one CTA per task, fixed reduction tree), then the mean over tasks in task order
__global__ void __launch_bounds__(256)
udiff_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ task_norm, long per_task) {
  __shared__ double red[256];
  const int t = blockIdx.x;
  const float* pa = a + (long)t * per_task;
  const float* pb = b + (long)t * per_task;
  double s = 0.0;
  for (long i = threadIdx.x; i < per_task; i += 256) {
    const float d = pa[i] - pb[i];
    s += (double)d * (double)d;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) task_norm[t] = sqrtf((float)red[0]);
}

// crit[2 it], crit[2 it + 1] = mean_t sqrt(sum_n row_sq[it, t, n]) (logged twice per iteration upstream,
// hard_kmeans.py:203,208-209): one CTA per iteration, one warp per task (lane-strided, then a shuffle tree: a fixed order),
// then the mean over the tasks in task order.  task_norm: [iters, T] scratch.
__global__ void __launch_bounds__(256)
row_norm_mean_kernel(const double* __restrict__ row_sq, float* __restrict__ task_norm, float* __restrict__ crit, int T, int n) {
  const int it = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* rs = row_sq + (long)it * T * n;
  float* tn = task_norm + (long)it * T;
  for (int t = warp; t < T; t += 8) {
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += rs[(long)t * n + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) tn[t] = sqrtf((float)s);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int t = 0; t < T; ++t) total += (double)tn[t];
    crit[2 * it] = crit[2 * it + 1] = (float)(total / (double)T);
  }
}

__global__ void mean_kernel(const float* __restrict__ vals, float* __restrict__ out, int T) {
  double total = 0.0;
  for (int t = 0; t < T; ++t) total += (double)vals[t];
  *out = (float)(total / (double)T);
}

}  // namespace

cudaError_t normalize_rows(const float* x, float* out, long rows, int D, cudaStream_t st) {
  normalize_rows_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, st>>>(x, out, rows, D);
  note_launch();
  return cudaGetLastError();
}

// u[m,:] = softmax_k(scale * a[m,:] . text[k,:]) for M rows (all tasks flattened; text is shared).  The product runs on the
// tensor cores (3 x TF32 tcgen05 tiles with the fp32 round-to-nearest running sum of contraction_tc.cu) whenever TMA can
// address the operands, else on the CUDA cores.
cudaError_t kmeans_similarity(const float* a, const float* text, float scale, float* u, long M, int K, int D,
                              cudaStream_t st) {
  // batches of <= 64 * 65535 rows through blockIdx.y
  if (M > 64L * 65535) return cudaErrorInvalidValue;
  const bool tc = D >= 4 && (D % 4) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(text)) & 15) == 0;
  if (tc) {
    cudaError_t e = gemm_nt_tc(a, text, u, 1, (int)M, K, D, 1, 0.0f, nullptr, false, st);
    if (e != cudaSuccess) return e;
  } else {
    pair_kernel<0><<<dim3((K + kTile - 1) / kTile, (unsigned)((M + kTile - 1) / kTile), 1), 256, 0, st>>>(
        a, text, nullptr, u, (int)M, K, D, 0, 0, 0);
    note_launch();
  }
  launch_assign(u, nullptr, nullptr, scale, 0.0f, u, nullptr, (int)M, 1, K, 3, nullptr, st);
  note_launch();
  return cudaGetLastError();
}

cudaError_t kmeans_centroids(const float* u, const float* x, float* w, int T, int n, int K, int D, int mode,
                             cudaStream_t st) {
  centroids_kernel<<<dim3((D + kTile - 1) / kTile, (K + kTile - 1) / kTile, T), 256, 0, st>>>(u, x, w, n, K, D, mode);
  note_launch();
  return cudaGetLastError();
}

cudaError_t kmeans_sqdist(const float* x, const float* w, float* d2, int T, int n, int K, int D, cudaStream_t st) {
  pair_kernel<1><<<dim3((K + kTile - 1) / kTile, (n + kTile - 1) / kTile, T), 256, 0, st>>>(
      x, w, nullptr, d2, n, K, D, (long)n * D, (long)K * D, (long)n * K);
  note_launch();
  return cudaGetLastError();
}

cudaError_t kmeans_precisions(const float* u, const float* x, const float* w, float* s, int T, int n, int K, int D,
                              int keep_old, cudaStream_t st) {
  precisions_kernel<<<dim3((D + kTile - 1) / kTile, (K + kTile - 1) / kTile, T), 256, 0, st>>>(u, x, w, s, n, K, D,
                                                                                              keep_old);
  note_launch();
  return cudaGetLastError();
}

// d2s[t,n,k] = sum_d s (w - x)^2 and det[t,k] = 1/2 sum_d log(s + eps)
cudaError_t kmeans_sqdist_cov(const float* x, const float* w, const float* s, float* d2s, float* det, int T, int n, int K,
                              int D, cudaStream_t st) {
  pair_kernel<2><<<dim3((K + kTile - 1) / kTile, (n + kTile - 1) / kTile, T), 256, 0, st>>>(
      x, w, s, d2s, n, K, D, (long)n * D, (long)K * D, (long)n * K);
  half_logdet_kernel<<<(unsigned)(((long)T * K + 3) / 4), 128, 0, st>>>(s, det, (long)T * K, D);
  note_launch(2);
  return cudaGetLastError();
}

cudaError_t kmeans_kl_div(const float* x, const float* w, float* div, int T, int n, int K, int D, cudaStream_t st) {
  pair_kernel<3><<<dim3((K + kTile - 1) / kTile, (n + kTile - 1) / kTile, T), 256, 0, st>>>(
      x, w, nullptr, div, n, K, D, (long)n * D, (long)K * D, (long)n * K);
  note_launch();
  return cudaGetLastError();
}

cudaError_t kmeans_assign(const float* d2, const float* v, const float* bias, float temperature, float lambd, float* u,
                          int* labels, int T, int n, int K, int mode, cudaStream_t st) {
  const int rows = T * n;
  launch_assign(d2, v, bias, temperature, lambd, u, labels, rows, n, K, mode, nullptr, st);
  note_launch();
  return cudaGetLastError();
}

// hard k-means u_update in place: u <- one-hot(argmin softmax(+d2)); row_sq [T * n] receives || u_old - u ||^2 per query
cudaError_t kmeans_assign_hard_tracked(const float* d2, float* u, int* labels, double* row_sq, int T, int n, int K,
                                       cudaStream_t st) {
  const int rows = T * n;
  launch_assign(d2, nullptr, nullptr, 1.0f, 0.0f, u, labels, rows, n, K, 2, row_sq, st);
  note_launch();
  return cudaGetLastError();
}

// the logged criteria of `iters` iterations from their per-query terms row_sq [iters, T, n]; task_norm [iters, T] scratch
cudaError_t kmeans_hard_criterions(const double* row_sq, float* task_norm, float* crit, int iters, int T, int n,
                                   cudaStream_t st) {
  if (iters <= 0) return cudaSuccess;
  row_norm_mean_kernel<<<iters, 256, 0, st>>>(row_sq, task_norm, crit, T, n);
  note_launch();
  return cudaGetLastError();
}

cudaError_t kmeans_udiff(const float* a, const float* b, float* task_norm, float* mean_out, int T, long per_task,
                         cudaStream_t st) {
  udiff_kernel<<<T, 256, 0, st>>>(a, b, task_norm, per_task);
  mean_kernel<<<1, 1, 0, st>>>(task_norm, mean_out, T);
  note_launch(2);
  return cudaGetLastError();
}

}  // namespace tclip
