// dirichlet_estep.cu — everything of the Dirichlet EM loop that is not the MM M-step (sm_100a).
//
// Reference call sites (SegoleneMartin/transductive-CLIP, src/methods/zero_shot/em_dirichlet.py unless noted):
//   cluster sizes, live mask, v_update       :145-151, :217-218      -> colsum_v_kernel
//   moments y_cst (+ few-shot support terms) :219-222, few_shot/em_dirichlet.py:196-200 -> moments_kernel
//   empty-cluster restore + outer criterion  :224-226, :236-238      -> commit_kernel, criterion_kernel
//   Dirichlet log-normaliser                 :35-36                  -> lognorm_kernel
//   contraction log z . (alpha-1)^T          :37-38                  -> logits_kernel
//   softmax / argmax one-hot                 :142-143, hard_em_dirichlet.py:256-258 -> softmax_kernel
//   prototypes for label matching            :61-70, utils.py:380-399 -> cluster_order_kernel, prototypes_kernel
//
// Layouts (all row-major, innermost last): u [T,n,K], logz [T,n,D], alpha/y/work [T,K,D], colsum/v [T,K].
#include <cuda_runtime.h>

#include <algorithm>
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
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- log(z + eps) -----------------------------------------------------------------------------------------------
__global__ void log_features_kernel(const float* __restrict__ x, float* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = logf(x[i] + kEps);
}

// ---- cluster sizes, live mask and the dual variable v ------------------------------------------------------------
// colsum[t,k] = sum_n u[t,n,k];  live = colsum > 1e-15;  v = log(colsum / n + 1e-15) + 1
__global__ void colsum_v_kernel(const float* __restrict__ u, float* __restrict__ colsum, float* __restrict__ v,
                                int* __restrict__ live, int n, int K) {
  const int t = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float* up = u + (long)t * n * K + k;
  float s = 0.0f;
  for (int i = 0; i < n; ++i) s += up[(long)i * K];
  colsum[(long)t * K + k] = s;
  if (v) v[(long)t * K + k] = logf(s / (float)n + kEps) + 1.0f;
  if (live) live[(long)t * K + k] = s > kEps ? 1 : 0;
}

// ---- moments: y[t,k,d] = sum_n u[t,n,k] logz[t,n,d] / colsum[t,k]  (a [K x n] . [n x D] product per task) --------
// 64(k) x 64(d) tile per CTA, 256 threads, 4x4 outputs per thread, n staged through shared memory.
constexpr int kMomTile = 128;   // clusters x feature columns per CTA
constexpr int kMomStage = 16;   // queries staged per step

// `gate` (optional, device): {n_live, row cap}; the dense kernel runs iff n_live > cap, the row-wise one otherwise, so
// the host can enqueue both without knowing how many clusters are alive.
__device__ __forceinline__ bool dense_selected(const int* gate) { return gate == nullptr || gate[0] > gate[1]; }

// 128 x 128 outputs per CTA, 256 threads, 8 x 8 per thread as two 4-wide groups 64 apart in each direction (16-byte
// shared-memory reads without bank conflicts, 16-byte global accesses).  Every output is one fma chain over the queries in
// index order, the order moments_rows_kernel uses too.
__global__ void __launch_bounds__(256)
moments_kernel(const float* __restrict__ u, const float* __restrict__ logz, const float* __restrict__ colsum,
               const float* __restrict__ support_sum, const float* __restrict__ support_count,
               float* __restrict__ y, int n, int K, int D, int few_shot, const int* __restrict__ gate) {
  if (!dense_selected(gate)) return;
  __shared__ __align__(16) float us[kMomStage][kMomTile + 4];
  __shared__ __align__(16) float ls[kMomStage][kMomTile + 4];
  const int t = blockIdx.z;
  const int k0 = blockIdx.y * kMomTile;
  const int d0 = blockIdx.x * kMomTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* ub = u + (long)t * n * K;
  const float* lb = logz + (long)t * n * D;
  const bool vec = ((K | D) & 3) == 0 &&
                   ((reinterpret_cast<unsigned long long>(u) | reinterpret_cast<unsigned long long>(logz) |
                     reinterpret_cast<unsigned long long>(y)) & 15ull) == 0;

  float acc[8][8] = {};
  for (int n0 = 0; n0 < n; n0 += kMomStage) {
    if (vec) {
      for (int i = threadIdx.x; i < kMomStage * kMomTile / 4; i += 256) {
        const int r = i / (kMomTile / 4), c = (i % (kMomTile / 4)) * 4;
        const int nn = n0 + r;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (nn < n && k0 + c < K) a = *reinterpret_cast<const float4*>(ub + (long)nn * K + k0 + c);
        if (nn < n && d0 + c < D) b = *reinterpret_cast<const float4*>(lb + (long)nn * D + d0 + c);
        *reinterpret_cast<float4*>(&us[r][c]) = a;
        *reinterpret_cast<float4*>(&ls[r][c]) = b;
      }
    } else {
      for (int i = threadIdx.x; i < kMomStage * kMomTile; i += 256) {
        const int r = i / kMomTile, c = i % kMomTile;
        const int nn = n0 + r;
        us[r][c] = (nn < n && k0 + c < K) ? ub[(long)nn * K + k0 + c] : 0.0f;
        ls[r][c] = (nn < n && d0 + c < D) ? lb[(long)nn * D + d0 + c] : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kMomStage; ++r) {
      float uu[8], ll[8];
      const float4 u0 = *reinterpret_cast<const float4*>(&us[r][ty * 4]);
      const float4 u1 = *reinterpret_cast<const float4*>(&us[r][64 + ty * 4]);
      const float4 l0 = *reinterpret_cast<const float4*>(&ls[r][tx * 4]);
      const float4 l1 = *reinterpret_cast<const float4*>(&ls[r][64 + tx * 4]);
      uu[0] = u0.x; uu[1] = u0.y; uu[2] = u0.z; uu[3] = u0.w; uu[4] = u1.x; uu[5] = u1.y; uu[6] = u1.z; uu[7] = u1.w;
      ll[0] = l0.x; ll[1] = l0.y; ll[2] = l0.z; ll[3] = l0.w; ll[4] = l1.x; ll[5] = l1.y; ll[6] = l1.z; ll[7] = l1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(uu[i], ll[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (k >= K) continue;
    const float cs = colsum[(long)t * K + k];
    const float sc = few_shot ? support_count[(long)t * K + k] : 0.0f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int dbase = d0 + h * 64 + tx * 4;
      float val[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = dbase + j;
        const long o = ((long)t * K + k) * D + d;
        const float a = acc[i][h * 4 + j];
        if (few_shot) {
          // (1 / (count_s + sum u)) * (support_sum + sum u logz): few_shot/em_dirichlet.py:196-200
          val[j] = d < D ? (1.0f / (sc + cs)) * (support_sum[o] + a) : 0.0f;
        } else {
          val[j] = cs > kEps ? a / fmaxf(cs, kEps) : -10.0f;
        }
      }
      const long o0 = ((long)t * K + k) * D + dbase;
      if (vec && dbase + 3 < D) {
        *reinterpret_cast<float4*>(y + o0) = make_float4(val[0], val[1], val[2], val[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (dbase + j < D) y[o0 + j] = val[j];
      }
    }
  }
}

// Row-wise form for the skip-dead schedule: one CTA per row of `rows` (live rows: the moments; newly dead rows: the -10
// fill they start their trajectory from).  Each output is the same sequential fma chain over n as in moments_kernel, so
// both forms give bit-identical y.
__global__ void __launch_bounds__(256)
moments_rows_kernel(const float* __restrict__ u, const float* __restrict__ logz, const float* __restrict__ colsum,
                    float* __restrict__ y, const int* __restrict__ rows, const int* __restrict__ n_rows, int fill_only,
                    int n, int K, int D, const int* __restrict__ gate) {
  if (dense_selected(gate)) return;
  if (fill_only) {  // newly empty clusters (any number of them): y = -10
    const int nr = *n_rows;
    for (int b = blockIdx.x; b < nr; b += gridDim.x) {
      float* o = y + (long)rows[b] * D;
      for (int d = threadIdx.x; d < D; d += blockDim.x) o[d] = -10.0f;
    }
    return;
  }
  if ((int)blockIdx.x >= *n_rows) return;
  const int row = rows[blockIdx.x];
  float* out = y + (long)row * D;
  extern __shared__ float ucol[];  // [n] responsibilities of this cluster
  const int t = row / K, k = row % K;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ucol[i] = u[((long)t * n + i) * K + k];
  __syncthreads();
  const float cs = colsum[row];
  const float* lb = logz + (long)t * n * D;
  // four strided columns per thread at a time: the 75-step fma chain of a column is latency-bound, four of them overlap;
  // every column still sums i = 0..n-1 in order (bit-identical to moments_kernel)
  for (int d0 = threadIdx.x; d0 < D; d0 += 4 * blockDim.x) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int i = 0; i < n; ++i) {
      const float ui = ucol[i];
      const float* li = lb + (long)i * D + d0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (d0 + j * (int)blockDim.x < D) acc[j] = fmaf(ui, li[j * blockDim.x], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (d0 + j * (int)blockDim.x < D) out[d0 + j * blockDim.x] = cs > kEps ? acc[j] / fmaxf(cs, kEps) : -10.0f;
  }
}

// dst[t, c, r] = src[t, r, c], rows r >= R of the padded pitch Rp zero: 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256)
transpose_pad_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C, int Rp, const int* __restrict__ gate) {
  if (!dense_selected(gate)) return;
  __shared__ float tile[32][33];
  const int t = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* sb = src + (long)t * R * C;
  float* db = dst + (long)t * C * Rp;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? sb[(long)r * C + c] : 0.0f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < Rp) db[(long)c * Rp + r] = tile[tx][i];
  }
}

// ---- few-shot support statistics (iteration invariant): per-class count and per-class sum of log-features -------
__global__ void support_stats_kernel(const float* __restrict__ log_support, const long long* __restrict__ y_s,
                                     float* __restrict__ support_sum, float* __restrict__ support_count, int S,
                                     int K, int D) {
  // one CTA per (task, class): samples are scanned in index order so the sum order is deterministic
  const int t = blockIdx.y, k = blockIdx.x;
  const long long* ys = y_s + (long)t * S;
  const float* xs = log_support + (long)t * S * D;
  __shared__ int members[1024];
  __shared__ int count;
  int total = 0;
  for (int base = 0; base < S; base += 1024) {
    if (threadIdx.x == 0) {
      int c = 0;
      for (int i = base; i < min(S, base + 1024); ++i)
        if (ys[i] == k) members[c++] = i;
      count = c;
    }
    __syncthreads();
    const int c = count;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float acc = (base == 0) ? 0.0f : support_sum[((long)t * K + k) * D + d];
      for (int m = 0; m < c; ++m) acc += xs[(long)members[m] * D + d];
      support_sum[((long)t * K + k) * D + d] = acc;
    }
    total += c;
    __syncthreads();
  }
  if (threadIdx.x == 0) support_count[(long)t * K + k] = (float)total;
}

// ---- commit: empty clusters keep their previous row; per-row pieces of the outer criterion ------------------------
// rowstat[row] = (||old - new||^2, ||old||^2) over the row; alpha[row] <- work[row] when the cluster is live.
__global__ void __launch_bounds__(128)
commit_kernel(float* __restrict__ alpha, const float* __restrict__ work, const int* __restrict__ live,
              const int* __restrict__ dead_age, double2* __restrict__ rowstat, int rows, int D) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 4 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 4) {   // capped grid, warps stride
    // a cluster that was already empty in the previous outer iteration has the same alpha row as then: its
    // (0, ||alpha||^2) entry written by that commit is still right
    if (dead_age && dead_age[row] >= 2) continue;
    const bool lv = live ? live[row] != 0 : true;
    float* a = alpha + (long)row * D;
    const float* w = work + (long)row * D;
    float dsq = 0.0f, osq = 0.0f;
    for (int d = lane; d < D; d += 32) {
      const float o = a[d];
      const float nw = lv ? w[d] : o;
      const float df = o - nw;
      dsq = fmaf(df, df, dsq);
      osq = fmaf(o, o, osq);
      if (lv) a[d] = nw;
    }
    const double ds = warp_sum_f64((double)dsq), os = warp_sum_f64((double)osq);
    if (lane == 0) rowstat[row] = make_double2(ds, os);
  }
}

// criterion[t] = ||alpha_old - alpha||_F / ||alpha_old||_F: one CTA per task (fixed reduction tree), then the mean over
// tasks accumulated in task order by one thread
__global__ void __launch_bounds__(256)
criterion_task_kernel(const double2* __restrict__ rowstat, float* __restrict__ task_crit, int K) {
  __shared__ double red[256];
  const int t = blockIdx.x;
  double dx = 0.0, dy = 0.0;
  for (int k = threadIdx.x; k < K; k += 256) {
    const double2 r = rowstat[(long)t * K + k];
    dx += r.x;
    dy += r.y;
  }
  red[threadIdx.x] = dx;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  const double num = red[0];
  __syncthreads();
  red[threadIdx.x] = dy;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  const double den = red[0];
  if (threadIdx.x == 0) task_crit[t] = sqrtf((float)num) / sqrtf((float)den);
}

__global__ void criterion_mean_kernel(const float* __restrict__ task_crit, float* __restrict__ crit_out, int T) {
  double total = 0.0;
  for (int t = 0; t < T; ++t) total += (double)task_crit[t];
  *crit_out = (float)(total / (double)T);
}

// ---- Dirichlet log-normaliser: norm[t,k] = lnGamma(sum_d a) - sum_d lnGamma(a), float64, one warp per row --------
__global__ void __launch_bounds__(128)
lognorm_kernel(const float* __restrict__ alpha, double* __restrict__ norm, const int* __restrict__ live, int rows,
               int D, const int* __restrict__ gate) {
  if (!dense_selected(gate)) return;  // few live rows: estep_task_kernel
  const int lane = threadIdx.x & 31;
  // warps stride over the rows (a capped grid: when the gate is closed the launch costs a few hundred CTAs, not rows / 4)
  for (int row = blockIdx.x * 4 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 4) {
    if (live && !live[row]) continue;  // empty cluster: alpha row unchanged, norm[row] of the previous E-step still holds
    const float* a = alpha + (long)row * D;
    double s = 0.0, lg = 0.0;
    for (int d = lane; d < D; d += 32) {
      const double v = (double)a[d];
      s += v;
      lg += lgamma(v);
    }
    s = warp_sum_f64(s);
    lg = warp_sum_f64(lg);
    if (lane == 0) norm[row] = lgamma(s) - lg;
  }
}

// ---- contraction l3[t,n,k] = sum_d logz[t,n,d] (alpha[t,k,d] - 1): [n x D] . [D x K] per task ---------------------
// 64(n) x 64(k) tile, BK = 16, 256 threads, 4x4 per thread; partial sums per BK slab are folded into the running
// total so the summation is blocked rather than a single 1000-term chain.
constexpr int kLgTile = 64;
constexpr int kLgBK = 16;

__global__ void __launch_bounds__(256)
logits_kernel(const float* __restrict__ logz, const float* __restrict__ alpha, float* __restrict__ l3, int n, int K,
              int D, const int* __restrict__ gate) {
  if (!dense_selected(gate)) return;
  __shared__ float as[kLgBK][kLgTile + 4];  // logz tile, transposed: [d][n]
  __shared__ float bs[kLgBK][kLgTile + 4];  // (alpha-1) tile, transposed: [d][k]
  const int t = blockIdx.z;
  const int n0 = blockIdx.y * kLgTile;
  const int k0 = blockIdx.x * kLgTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* ab = logz + (long)t * n * D;
  const float* bb = alpha + (long)t * K * D;
  float acc[4][4] = {};
  for (int d0 = 0; d0 < D; d0 += kLgBK) {
    for (int i = threadIdx.x; i < kLgTile * kLgBK; i += 256) {
      const int r = i / kLgBK, c = i % kLgBK;
      const int d = d0 + c;
      as[c][r] = (n0 + r < n && d < D) ? ab[(long)(n0 + r) * D + d] : 0.0f;
      bs[c][r] = (k0 + r < K && d < D) ? bb[(long)(k0 + r) * D + d] - 1.0f : 0.0f;
    }
    __syncthreads();
    float part[4][4] = {};
#pragma unroll
    for (int c = 0; c < kLgBK; ++c) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = as[c][ty * 4 + i];
        bv[i] = bs[c][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nn = n0 + ty * 4 + i;
    if (nn >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) l3[((long)t * n + nn) * K + k] = acc[i][j];
    }
  }
}

// ---- responsibilities: u = softmax_k(norm + l3 + lambda v / n); optional argmax -> one-hot ------------------------
// One warp per (task, query).  labels = argmax of the *softmaxed* float32 values, first index wins
// (hard_em_dirichlet.py:256-258).  `u` may alias `l3` (each warp reads its row before overwriting it).
// K <= 1024: the 32 logits of a lane stay in registers (one pass over l3 / norm / v, one expf per element); same
// arithmetic and the same per-lane summation order as the general kernel below => identical results.
// (l3, norm: no __restrict__ — estep_task_kernel writes both earlier in the same kernel and must read them coherently)
// live_t / dead_max_row (optional): also leave max over the dead classes of float(norm + l3) for this query (the bound of the
// sparse soft-max below).
__device__ __forceinline__ void softmax_row_reg(const float* l3, const double* norm, const float* __restrict__ v,
                                                float lambd, float* u, int* __restrict__ labels, int row, int lane, int n,
                                                int K, int hard, const int* __restrict__ live_t = nullptr,
                                                float* dead_max_row = nullptr) {
  const int t = row / n;
  const float* x = l3 + (long)row * K;
  const double* nm = norm + (long)t * K;
  const float* vv = v + (long)t * K;
  float* out = u + (long)row * K;
  const float fn = (float)n;
  float lg[32];
  float mx = -CUDART_INF_F, dm = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int k = lane + 32 * j;
    if (k < K) {
      const float base = (float)(nm[k] + (double)x[k]);
      if (live_t && !live_t[k]) dm = fmaxf(dm, base);
      lg[j] = base + (lambd * vv[k]) / fn;
      mx = fmaxf(mx, lg[j]);
    }
  }
  mx = warp_max_f32(mx);
  if (dead_max_row) {
    dm = warp_max_f32(dm);
    if (lane == 0) *dead_max_row = dm;
  }
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (lane + 32 * j < K) {
      lg[j] = expf(lg[j] - mx);
      sum += lg[j];
    }
  }
  sum = warp_sum_f32(sum);
  float best = -1.0f;
  int best_k = 0x7fffffff;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const int k = lane + 32 * j;
    if (k < K) {
      const float p = lg[j] / sum;
      if (p > best) {  // strict: the lowest k of this lane's stripe wins ties
        best = p;
        best_k = k;
      }
      if (!hard) out[k] = p;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    if (ob > best || (ob == best && ok < best_k)) {
      best = ob;
      best_k = ok;
    }
  }
  if (hard) {
    for (int k = lane; k < K; k += 32) out[k] = (k == best_k) ? 1.0f : 0.0f;
  }
  if (lane == 0 && labels) labels[row] = best_k;
}

// `gate`: in the skip-dead schedule the dense E-step kernels run iff more live rows than the row-wise cap (else
// estep_task_kernel takes the whole E-step of a task)
__global__ void __launch_bounds__(128)
softmax_reg_kernel(const float* l3, const double* __restrict__ norm, const float* __restrict__ v, float lambd,
                   float* u, int* __restrict__ labels, int rows, int n, int K, int hard, const int* __restrict__ gate,
                   int it, int* __restrict__ last_dense) {
  if (!dense_selected(gate)) return;
  if (last_dense && blockIdx.x == 0 && threadIdx.x == 0) *last_dense = it;   // full rows were written: see estep_task_kernel
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  softmax_row_reg(l3, norm, v, lambd, u, labels, row, threadIdx.x & 31, n, K, hard);
}

__device__ __forceinline__ void softmax_row_gen(const float* l3, const double* norm, const float* __restrict__ v,
                                                float lambd, float* u, int* __restrict__ labels, int row, int lane, int n,
                                                int K, int hard) {
  const int t = row / n;
  const float* x = l3 + (long)row * K;
  const double* nm = norm + (long)t * K;
  const float* vv = v + (long)t * K;
  float* out = u + (long)row * K;
  const float fn = (float)n;

  float mx = -CUDART_INF_F;
  for (int k = lane; k < K; k += 32) {
    const float logit = (float)(nm[k] + (double)x[k]) + (lambd * vv[k]) / fn;
    mx = fmaxf(mx, logit);
  }
  mx = warp_max_f32(mx);
  float sum = 0.0f;
  for (int k = lane; k < K; k += 32) {
    const float logit = (float)(nm[k] + (double)x[k]) + (lambd * vv[k]) / fn;
    sum += expf(logit - mx);
  }
  sum = warp_sum_f32(sum);
  float best = -1.0f;
  int best_k = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float logit = (float)(nm[k] + (double)x[k]) + (lambd * vv[k]) / fn;
    const float p = expf(logit - mx) / sum;
    if (p > best) {  // strict: the lowest k of this lane's stripe wins ties
      best = p;
      best_k = k;
    }
    if (!hard) out[k] = p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    if (ob > best || (ob == best && ok < best_k)) {
      best = ob;
      best_k = ok;
    }
  }
  if (hard) {
    for (int k = lane; k < K; k += 32) out[k] = (k == best_k) ? 1.0f : 0.0f;
  }
  if (lane == 0 && labels) labels[row] = best_k;
}

__global__ void __launch_bounds__(128)
softmax_kernel(const float* l3, const double* __restrict__ norm, const float* __restrict__ v, float lambd,
               float* u, int* __restrict__ labels, int rows, int n, int K, int hard, const int* __restrict__ gate,
               int it, int* __restrict__ last_dense) {
  if (!dense_selected(gate)) return;
  if (last_dense && blockIdx.x == 0 && threadIdx.x == 0) *last_dense = it;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  softmax_row_gen(l3, norm, v, lambd, u, labels, row, threadIdx.x & 31, n, K, hard);
}

// ---- the whole E-step of one task in ONE kernel, for the skip-dead schedule with few live rows --------------------------
// get_logits + u_update (em_dirichlet.py:28-40,132-143; hard arg-max hard_em_dirichlet.py:256-258) for the clusters that are
// alive: an empty cluster keeps its alpha row, hence its log-normaliser and its column of l3, so per task only the live
// classes are recomputed — log-normaliser (float64 lgamma sums), contraction column l3[t, :, k] — and then the soft-max rows of
// the task's queries are formed from the persistent l3 / norm.  One launch (a CTA of 512 threads per task and slice of 16
// queries: every slice recomputes the task's few log-normalisers, writes identical values, and owns its queries' l3 entries and
// soft-max rows) replaces three (log-normaliser and contraction with one CTA per live row, then the soft-max of all rows);
// every value is produced by the same per-thread operation sequence as in the dense kernels (same lane partition, same
// shuffle trees), so the results are bit-identical to them.
constexpr int kTaskThreads = 512;
constexpr int kTaskWarps = kTaskThreads / 32;
constexpr int kTaskRowsPerPass = 8;   // live classes whose alpha - 1 rows are staged in shared memory at a time
constexpr int kTaskSlice = 16;        // queries per CTA (a multiple of the 4-query groups of the contraction)

// Exact sparse soft-max.  In the tail of the EM ~3 of the K classes of a task are alive; the logits of the dead ones are
// frozen (their alpha row, log-normaliser and l3 column do not change while they stay dead; their v is log(~1e-15) + 1), far
// below the live ones.  A soft-max over all K classes gives them exp(logit - max) = +0.0f exactly whenever logit - max <= -104,
// and adding +0.0f leaves every partial sum unchanged, so the row equals the soft-max over the live classes alone — bit for
// bit, as long as each lane still sums its own classes in increasing order.  Per query the bound is
// float(dead_max + max_dead float(float(lambda v) / n)) >= every dead logit (rounding is monotone), with dead_max = max over
// the dead classes of float(norm + l3), recomputed by a pass over all classes whenever the set of dead classes changes.  A row
// whose bound is not 106 below its live maximum takes the full pass.  Rows written by a full pass may hold non-zero
// responsibilities of dead classes: the next sparse pass of that task zeroes its rows first (soft variant; the hard variant
// only moves the 1 of the one-hot row).
constexpr int kSparseMaxLive = 64;       // more live classes per task: full rows
constexpr float kExpUnderflow = -106.0f;  // expf(x) == +0.0f for x <= -104 (exp(-104) = 6.8e-46 < half the smallest denormal)

// returns false (warp-uniform) if the bound does not hold: the caller then runs the full row
__device__ __forceinline__ bool softmax_row_sparse(const float* l3, const double* norm, const float* lterm, float* u,
                                                   int* __restrict__ labels, int row, int lane, int n, int K, int hard,
                                                   const int* cls, int nc, float bound, bool zero_row) {
  const int t = row / n;
  const float* x = l3 + (long)row * K;
  const double* nm = norm + (long)t * K;
  float* out = u + (long)row * K;
  float mx = -CUDART_INF_F;
  for (int i = 0; i < nc; ++i) {
    const int k = cls[i];
    if ((k & 31) == lane) mx = fmaxf(mx, (float)(nm[k] + (double)x[k]) + lterm[k]);
  }
  mx = warp_max_f32(mx);
  if (!(bound - mx <= kExpUnderflow)) return false;   // (also taken when anything is NaN)
  float sum = 0.0f;
  for (int i = 0; i < nc; ++i) {          // cls is ascending: every lane adds its classes in the order of the full pass
    const int k = cls[i];
    if ((k & 31) == lane) sum += expf(((float)(nm[k] + (double)x[k]) + lterm[k]) - mx);
  }
  sum = warp_sum_f32(sum);
  const int prev = (hard && labels) ? labels[row] : -1;
  if (zero_row && !hard) {
    for (int k = lane; k < K; k += 32) out[k] = 0.0f;   // (the live entries below are written by the same lane afterwards)
  }
  float best = -1.0f;
  int best_k = 0x7fffffff;
  for (int i = 0; i < nc; ++i) {
    const int k = cls[i];
    if ((k & 31) == lane) {
      const float p = expf(((float)(nm[k] + (double)x[k]) + lterm[k]) - mx) / sum;
      if (p > best) {
        best = p;
        best_k = k;
      }
      if (!hard) out[k] = p;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    if (ob > best || (ob == best && ok < best_k)) {
      best = ob;
      best_k = ok;
    }
  }
  if (lane == 0) {
    if (hard) {
      if (prev >= 0 && prev < K) out[prev] = 0.0f;
      out[best_k] = 1.0f;
    }
    if (labels) labels[row] = best_k;
  }
  return true;
}

__global__ void __launch_bounds__(kTaskThreads, 2)
estep_task_kernel(const float* __restrict__ alpha, const float* __restrict__ logz, const float* __restrict__ v, float lambd,
                  double* __restrict__ norm, float* l3, float* u, int* labels, const int* __restrict__ live,
                  int n, int K, int D, int hard, const SparseRows sp) {
  if (dense_selected(sp.gate)) return;
  extern __shared__ float am1[];               // [kTaskRowsPerPass][D] alpha - 1 of the classes of this pass
  __shared__ int cls[1024];                    // live classes of this task (K <= 1024: see the launcher)
  __shared__ int sorted[kSparseMaxLive];
  __shared__ float lterm[1024];                // float(float(lambda v) / n) per class
  __shared__ int n_cls;
  __shared__ double ps[kTaskWarps], pl[kTaskWarps];
  __shared__ float wmax[kTaskWarps];
  const int t = blockIdx.x;
  const int q_lo = blockIdx.y * kTaskSlice, q_hi = min(n, q_lo + kTaskSlice);   // this CTA's queries
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* live_t = live + (long)t * K;
  if (tid == 0) n_cls = 0;
  __syncthreads();
  for (int k = tid; k < K; k += kTaskThreads)
    if (live_t[k]) cls[atomicAdd(&n_cls, 1)] = k;   // any order: every class is handled independently
  __syncthreads();
  const int nc = n_cls;
  // log-normalisers: four warps per class (the partition of the dense kernel's warp, four times), four classes at a time
  {
    const int grp = warp >> 2, gtid = tid & 127, gwarp = warp & 3;
    for (int c0 = 0; c0 < nc; c0 += kTaskWarps / 4) {
      const int c = c0 + grp;
      double s = 0.0, lg = 0.0;
      if (c < nc) {
        const float* a = alpha + ((long)t * K + cls[c]) * D;
        for (int d = gtid; d < D; d += 128) {
          const double x = (double)a[d];
          s += x;
          lg += lgamma(x);
        }
      }
      s = warp_sum_f64(s);
      lg = warp_sum_f64(lg);
      if (lane == 0) {
        ps[warp] = s;
        pl[warp] = lg;
      }
      __syncthreads();
      if (c < nc && gwarp == 0 && lane == 0) {
        const int w0 = grp * 4;
        norm[(long)t * K + cls[c]] = lgamma((ps[w0] + ps[w0 + 1]) + (ps[w0 + 2] + ps[w0 + 3])) -
                                     ((pl[w0] + pl[w0 + 1]) + (pl[w0 + 2] + pl[w0 + 3]));
      }
      __syncthreads();
    }
  }
  // contraction columns: l3[t, q, k] = sum_d logz[t, q, d] (alpha[t, k, d] - 1) for the live classes, a warp per (class, four
  // queries): lanes stride over d, one shuffle tree per query
  for (int c0 = 0; c0 < nc; c0 += kTaskRowsPerPass) {
    const int np = min(kTaskRowsPerPass, nc - c0);
    for (int i = tid; i < np * D; i += kTaskThreads) {
      const int c = i / D, d = i - c * D;
      am1[i] = alpha[((long)t * K + cls[c0 + c]) * D + d] - 1.0f;
    }
    __syncthreads();
    const int n_groups = (q_hi - q_lo + 3) / 4;
    for (int item = warp; item < np * n_groups; item += kTaskWarps) {
      const int c = item / n_groups, n0 = q_lo + (item - c * n_groups) * 4;
      const int k = cls[c0 + c];
      const float* z0 = logz + ((long)t * n + n0) * D;
      const float* b_row = am1 + c * D;
      const int nq = min(4, q_hi - n0);
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
      for (int d = lane; d < D; d += 32) {
        const float b = b_row[d];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nq) acc[q] = fmaf(z0[(long)q * D + d], b, acc[q]);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float tot = warp_sum_f32(acc[q]);
        if (lane == 0 && q < nq) l3[((long)t * n + n0 + q) * K + k] = tot;
      }
    }
    __syncthreads();
  }
  // ---- which soft-max this launch takes (uniform over the whole grid: every CTA reads the same device words, and the only
  // word written during the launch, a_iter <- it, reads as "not valid yet" whether a CTA sees the old or the new value)
  bool sparse = sp.dead_max != nullptr && nc <= kSparseMaxLive && (labels != nullptr || !hard);
  bool need_zero = false;
  if (sparse) {
    const int a_it = *sp.a_iter, s_it = *sp.set_iter;
    sparse = (*sp.changed == 0) && (s_it <= a_it) && (a_it < sp.it);
    // (last_full is double-buffered by the parity of the iteration: slices of this task that take a full row during this
    // launch write the other half)
    need_zero = (sp.last_full[((sp.it - 1) & 1) * (int)gridDim.x + t] == sp.it - 1) || (*sp.last_dense == sp.it - 1);
  }
  // per class the v term of the logit, and its maximum over the dead classes
  float dv = -CUDART_INF_F;
  const float fn = (float)n;
  for (int k = tid; k < K; k += kTaskThreads) {
    const float term = (lambd * v[(long)t * K + k]) / fn;
    lterm[k] = term;
    if (!live_t[k]) dv = fmaxf(dv, term);
  }
  dv = warp_max_f32(dv);
  if (lane == 0) wmax[warp] = dv;
  if (sparse) {   // ascending order of the live classes (rank sort; nc <= kSparseMaxLive)
    for (int i = tid; i < nc; i += kTaskThreads) {
      int rank = 0;
      for (int j = 0; j < nc; ++j) rank += cls[j] < cls[i];
      sorted[rank] = cls[i];
    }
  }
  // the norm / l3 entries written above by other threads of this CTA are read below: make them visible
  __threadfence_block();
  __syncthreads();
  float dead_v = wmax[0];
#pragma unroll
  for (int w = 1; w < kTaskWarps; ++w) dead_v = fmaxf(dead_v, wmax[w]);
  // responsibilities of the task's queries, one warp per query
  bool any_full = false;
  for (int q = q_lo + warp; q < q_hi; q += kTaskWarps) {
    const int row = t * n + q;
    bool done = false;
    if (sparse) done = softmax_row_sparse(l3, norm, lterm, u, labels, row, lane, n, K, hard, sorted, nc,
                                          sp.dead_max[row] + dead_v, need_zero);
    if (!done) {
      softmax_row_reg(l3, norm, v, lambd, u, labels, row, lane, n, K, hard, live_t, sp.dead_max ? sp.dead_max + row : nullptr);
      any_full = true;
    }
  }
  if (sp.dead_max) {
    if (any_full && lane == 0) sp.last_full[(sp.it & 1) * (int)gridDim.x + t] = sp.it;   // (every writer writes the same value)
    if (!sparse && tid == 0) *sp.a_iter = sp.it;                  // dead_max of every row was recomputed in this launch
  }
}

// ---- label matching inputs -----------------------------------------------------------------------------------------
// Per task: clusters in order of first appearance among the predictions (utils.py:389-396), the cluster index of
// every query, and the cluster sizes.  n <= 1024 queries per task.
__global__ void cluster_order_kernel(const int* __restrict__ labels, int* __restrict__ cluster_label,
                                     int* __restrict__ cluster_size, int* __restrict__ sample_cluster,
                                     int* __restrict__ n_clusters, int n) {
  extern __shared__ int sh[];
  int* lab = sh;          // [n]
  int* first = sh + n;    // [n] 1 when this sample opens a new cluster
  int* rank = sh + 2 * n; // [n] cluster index of the sample's cluster
  const int t = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    lab[i] = labels[(long)t * n + i];
    cluster_label[(long)t * n + i] = -1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int f = 1;
    for (int j = 0; j < i; ++j) f &= (lab[j] != lab[i]);
    first[i] = f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    // index of the first sample carrying my label, then the number of cluster openings before it
    int j0 = i;
    for (int j = 0; j < i; ++j)
      if (lab[j] == lab[i]) {
        j0 = j;
        break;
      }
    int r = 0;
    for (int j = 0; j < j0; ++j) r += first[j];
    rank[i] = r;
    sample_cluster[(long)t * n + i] = r;
    if (first[i]) cluster_label[(long)t * n + r] = lab[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    int sz = 0;
    for (int i = 0; i < n; ++i) sz += (rank[i] == c);
    cluster_size[(long)t * n + c] = sz;
  }
  if (threadIdx.x == 0) {
    int nc = 0;
    for (int i = 0; i < n; ++i) nc += first[i];
    n_clusters[t] = nc;
  }
}

// proto[t,c,:] = mean of the raw features of the queries in cluster c (index order), c < n_clusters[t]; rest zero.
// One CTA per (cluster, task): the cluster indices of the task's queries are staged once, a thread owns four feature
// dimensions and walks the queries (a CTA-uniform branch skips the non-members).  Sums in query order per dimension.
__global__ void __launch_bounds__(256)
prototypes_kernel(const float* __restrict__ feats, const int* __restrict__ sample_cluster,
                  const int* __restrict__ cluster_size, const int* __restrict__ n_clusters, float* __restrict__ proto, int n,
                  int D) {
  extern __shared__ int sc_s[];   // [n]
  const int t = blockIdx.y, c = blockIdx.x;
  float* out = proto + ((long)t * n + c) * D;
  const float* x = feats + (long)t * n * D;
  const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(feats) & 15) == 0 && (reinterpret_cast<uintptr_t>(proto) & 15) == 0;
  if (c >= n_clusters[t]) {
    if (vec)
      for (int d = threadIdx.x; d < D / 4; d += blockDim.x) reinterpret_cast<float4*>(out)[d] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    else
      for (int d = threadIdx.x; d < D; d += blockDim.x) out[d] = 0.0f;
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) sc_s[i] = sample_cluster[(long)t * n + i];
  __syncthreads();
  const float inv = fmaxf((float)cluster_size[(long)t * n + c], kEps);
  if (vec) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int D4 = D / 4;
    for (int d = threadIdx.x; d < D4; d += blockDim.x) {
      float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      for (int i = 0; i < n; ++i) {
        if (sc_s[i] != c) continue;
        const float4 v = __ldg(x4 + (long)i * D4 + d);
        acc.x += v.x;
        acc.y += v.y;
        acc.z += v.z;
        acc.w += v.w;
      }
      reinterpret_cast<float4*>(out)[d] = make_float4(acc.x / inv, acc.y / inv, acc.z / inv, acc.w / inv);
    }
  } else {
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float acc = 0.0f;
      for (int i = 0; i < n; ++i)
        if (sc_s[i] == c) acc += x[(long)i * D + d];
      out[d] = acc / inv;
    }
  }
}

}  // namespace

// ---- host-side launchers ------------------------------------------------------------------------------------------
cudaError_t log_features(const float* x, float* out, long count, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  log_features_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(x, out, count);
  note_launch(1);
  return cudaGetLastError();
}

cudaError_t colsum_v(const float* u, float* colsum, float* v, int* live, int T, int n, int K, cudaStream_t st) {
  colsum_v_kernel<<<dim3((K + 127) / 128, T), 128, 0, st>>>(u, colsum, v, live, n, K);
  note_launch(1);
  return cudaGetLastError();
}

cudaError_t transpose_pad(const float* src, float* dst, int T, int R, int C, int Rp, const int* gate, cudaStream_t st) {
  transpose_pad_kernel<<<dim3((C + 31) / 32, (Rp + 31) / 32, T), 256, 0, st>>>(src, dst, R, C, Rp, gate);
  note_launch(1);
  return cudaGetLastError();
}

// The dense form is a [K x n] . [n x D] product per task: with `tc` it runs on the tensor cores (u^T and (log z)^T as
// K-major operands of the 3 x TF32 tcgen05 kernel of contraction_tc.cu, the division / support terms / -10 fill in its
// epilogue), else on the CUDA cores in the summation order the row-wise kernel reproduces bit for bit.
cudaError_t moments(const float* u, const float* logz, const float* colsum, const float* support_sum,
                    const float* support_count, float* y, int T, int n, int K, int D, const SparseRows* sp,
                    cudaStream_t st, const MomentsTc* tc) {
  const int few = support_sum != nullptr;
  const int* gate = sp ? sp->gate : nullptr;
  if (tc) {
    if (cudaError_t e = transpose_pad(u, tc->uT, T, n, K, tc->np, gate, st)) return e;
    TcEpilogue ep;
    ep.mode = few ? 2 : 1;
    ep.colsum = colsum;
    ep.support_sum = support_sum;
    ep.support_count = support_count;
    if (cudaError_t e = gemm_nt_tc(tc->uT, tc->logzT, y, T, K, D, tc->np, T, 0.0f, gate, false, st, &ep)) return e;
  } else {
    moments_kernel<<<dim3((D + kMomTile - 1) / kMomTile, (K + kMomTile - 1) / kMomTile, T), 256, 0, st>>>(
        u, logz, colsum, support_sum, support_count, y, n, K, D, few, gate);
    note_launch(1);
  }
  if (sp) {
    moments_rows_kernel<<<sp->cap, 256, n * sizeof(float), st>>>(u, logz, colsum, y, sp->rows_live, sp->n_live, 0, n, K, D,
                                                               gate);
    moments_rows_kernel<<<sp->cap, 256, 0, st>>>(u, logz, colsum, y, sp->rows_new, sp->n_new, 1, n, K, D, gate);
    note_launch(2);
  }
  return cudaGetLastError();
}

cudaError_t support_stats(const float* log_support, const long long* y_s, float* support_sum, float* support_count,
                          int T, int S, int K, int D, cudaStream_t st) {
  support_stats_kernel<<<dim3(K, T), 256, 0, st>>>(log_support, y_s, support_sum, support_count, S, K, D);
  note_launch(1);
  return cudaGetLastError();
}

cudaError_t commit(float* alpha, const float* work, const int* live, const int* dead_age, double2* rowstat,
                   float* task_crit, float* crit_out, int T, int K, int D, cudaStream_t st) {
  const int rows = T * K;
  commit_kernel<<<std::min((rows + 3) / 4, 148 * 32), 128, 0, st>>>(alpha, work, live, dead_age, rowstat, rows, D);
  criterion_task_kernel<<<T, 256, 0, st>>>(rowstat, task_crit, K);
  criterion_mean_kernel<<<1, 1, 0, st>>>(task_crit, crit_out, T);
  note_launch(3);
  return cudaGetLastError();
}

cudaError_t logits_simt(const float* logz, const float* alpha, float* l3, int T, int n, int K, int D, const int* gate,
                        cudaStream_t st) {
  logits_kernel<<<dim3((K + kLgTile - 1) / kLgTile, (n + kLgTile - 1) / kLgTile, T), 256, 0, st>>>(logz, alpha, l3, n, K,
                                                                                                  D, gate);
  note_launch(1);
  return cudaGetLastError();
}

// The dense contraction runs on the tensor cores (contraction_tc.cu) whenever TMA can address the operands (D % 4 == 0,
// n <= 128); other shapes take the CUDA-core kernel.  TCLIP_CONTRACTION=simt forces the latter (measurements).
static bool use_tensor_cores(const float* logz, const float* alpha, int n, int K, int D) {
  static const bool forced_simt = [] {
    const char* e = std::getenv("TCLIP_CONTRACTION");
    return e && std::string(e) == "simt";
  }();
  const bool aligned = ((reinterpret_cast<unsigned long long>(logz) | reinterpret_cast<unsigned long long>(alpha)) & 15ull) == 0;
  return !forced_simt && aligned && logits_tc_supported(n, K, D);
}

// l3 == nullptr: the contraction is written into u and soft-maxed in place (stage entry point).  With a persistent l3
// buffer and `sp`, only live clusters are recomputed (norm and l3 of empty clusters carry over from the last E-step).
cudaError_t estep(const float* alpha, const float* logz, const float* v, float lambd, double* norm, float* l3, float* u,
                  int* labels, int T, int n, int K, int D, int hard, const int* live, const SparseRows* sp,
                  cudaStream_t st) {
  const int rows = T * K;
  float* dst = l3 ? l3 : u;
  const int* gate = (sp && l3) ? sp->gate : nullptr;
  lognorm_kernel<<<std::min((rows + 3) / 4, 148 * 32), 128, 0, st>>>(alpha, norm, l3 ? live : nullptr, rows, D, gate);
  note_launch(1);
  if (use_tensor_cores(logz, alpha, n, K, D)) {
    if (cudaError_t e = logits_tc(logz, alpha, dst, T, n, K, D, gate, false, st)) return e;
  } else {
    if (cudaError_t e = logits_simt(logz, alpha, dst, T, n, K, D, gate, st)) return e;
  }
  const int qrows = T * n;
  int* last_dense = gate ? sp->last_dense : nullptr;
  const int it = gate ? sp->it : 0;
  if (K <= 1024)
    softmax_reg_kernel<<<(qrows + 3) / 4, 128, 0, st>>>(dst, norm, v, lambd, u, labels, qrows, n, K, hard, gate, it, last_dense);
  else
    softmax_kernel<<<(qrows + 3) / 4, 128, 0, st>>>(dst, norm, v, lambd, u, labels, qrows, n, K, hard, gate, it, last_dense);
  note_launch(1);
  if (gate && K <= 1024) {   // (K > 1024: the driver's gate never selects the row-wise form, the class list of a task lives
                             // in a 1024-entry shared array)
    // few live rows: the whole E-step of a task in one kernel (the dense kernels above returned at once)
    const size_t smem = (size_t)kTaskRowsPerPass * D * sizeof(float);
    estep_task_kernel<<<dim3(T, (n + kTaskSlice - 1) / kTaskSlice), kTaskThreads, smem, st>>>(alpha, logz, v, lambd, norm, dst, u,
                                                                                            labels, live, n, K, D, hard, *sp);
    note_launch(1);
  }
  return cudaGetLastError();
}

cudaError_t cluster_prototypes(const int* labels, const float* feats, int* cluster_label, int* cluster_size,
                               int* sample_cluster, int* n_clusters, float* proto, int T, int n, int D,
                               cudaStream_t st) {
  if (n > 1024) return cudaErrorInvalidValue;
  cluster_order_kernel<<<T, 128, 3 * n * sizeof(int), st>>>(labels, cluster_label, cluster_size, sample_cluster,
                                                            n_clusters, n);
  prototypes_kernel<<<dim3(n, T), 256, n * sizeof(int), st>>>(feats, sample_cluster, cluster_size, n_clusters, proto, n, D);
  note_launch(2);
  return cudaGetLastError();
}

}  // namespace tclip
