// matching.cu — cluster -> class label matching and the task accuracy on the device.
//
// Replaces `compute_graph_matching` / `compute_basic_matching` of the reference (src/utils.py:380-417) and the accuracy
// line of `compute_acc_clustering` (src/methods/zero_shot/em_dirichlet.py:86-92).  The reference builds, per task, the
// cost matrix A[i, :] = -probs[task, cluster_i, :] in float64 (clusters in order of first appearance, #clusters <= n_query)
// and calls SciPy's `linear_sum_assignment` (third-party; scipy/optimize/rectangular_lsap, the shortest augmenting path
// algorithm of D. F. Crouse, "On implementing 2D rectangular assignment algorithms", IEEE TAES 52(4), 2016).  This is the
// same algorithm, restated for one CTA of kMatchThreads threads per task: the column scan of every Dijkstra step (update of
// the shortest path costs + arg-min) is dealt to the threads (shuffle reduction inside a warp, then across the warps through
// shared memory; the order (value, unassigned first, lowest index) is total, so the result does not depend on how the
// columns are dealt), the dual variables and the path live in shared memory, all arithmetic is float64 in SciPy's
// operation order.  An optimal assignment is unique unless reduced costs tie exactly; on ties this
// kernel takes an unassigned column before an assigned one and then the lowest column index (SciPy: order of its
// `remaining` array) — with float64 costs from distinct float32 cluster means ties do not occur in practice, and the GPU
// tests compare against SciPy itself.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "tclip_kernels.cuh"

namespace tclip {

namespace {

struct Best {
  double val;
  int unassigned;  // 1 if the column has no row yet
  int idx;
};

__device__ __forceinline__ bool better(const Best& a, const Best& b) {  // is a preferred to b
  if (a.val != b.val) return a.val < b.val;
  if (a.unassigned != b.unassigned) return a.unassigned > b.unassigned;
  return a.idx < b.idx;
}

constexpr int kMatchThreads = 256;
constexpr int kMatchWarps = kMatchThreads / 32;

// CTA-wide arg-min of `b` under `better`; every thread returns the winner.  red: [kMatchWarps] in shared memory.
__device__ __forceinline__ Best block_best(Best b, Best* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best c;
    c.val = __shfl_xor_sync(0xffffffffu, b.val, o);
    c.unassigned = __shfl_xor_sync(0xffffffffu, b.unassigned, o);
    c.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
    if (better(c, b)) b = c;
  }
  __syncthreads();                       // the previous round's readers are done with red[]
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = b;
  __syncthreads();
  b = red[0];
#pragma unroll
  for (int w = 1; w < kMatchWarps; ++w)
    if (better(red[w], b)) b = red[w];
  return b;
}

// One CTA per task.  Shared memory: v, spc [nc] double; path, row4col [nc] int; sc [nc] unsigned char;
// u [nr] double; col4row [nr] int; sr [nr] unsigned char.
__global__ void __launch_bounds__(kMatchThreads)
match_clusters_kernel(const float* __restrict__ proto, const int* __restrict__ n_clusters,
                      const int* __restrict__ sample_cluster, const long long* __restrict__ y_q, int graph_matching,
                      int* __restrict__ cluster_class, long long* __restrict__ new_labels, float* __restrict__ acc, int n,
                      int K, int proto_rows) {
  extern __shared__ double smem_d[];
  __shared__ Best red[kMatchWarps];
  __shared__ float redf[kMatchWarps];
  __shared__ int redi[kMatchWarps];
  constexpr int NT = kMatchThreads;
  const int t = blockIdx.x, lane = threadIdx.x;   // `lane`: index of the thread in the CTA
  const int nc = K;
  const int nr = min(min(n_clusters[t], n), proto_rows);
  double* v = smem_d;
  double* spc = v + nc;
  double* u = spc + nc;                      // [n]
  int* path = reinterpret_cast<int*>(u + n);
  int* row4col = path + nc;
  int* col4row = row4col + nc;               // [n]
  unsigned char* sc = reinterpret_cast<unsigned char*>(col4row + n);
  unsigned char* sr = sc + nc;               // [n]
  const float* P = proto + (long)t * proto_rows * K;

  if (!graph_matching) {
    // compute_basic_matching: every cluster takes the arg-max class of its prototype (first maximum)
    for (int i = 0; i < nr; ++i) {
      float bv = -CUDART_INF_F;
      int bi = 0x7fffffff;
      for (int j = lane; j < nc; j += NT) {
        const float x = P[(long)i * K + j];
        if (x > bv) {
          bv = x;
          bi = j;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      __syncthreads();
      if ((lane & 31) == 0) {
        redf[lane >> 5] = bv;
        redi[lane >> 5] = bi;
      }
      __syncthreads();
      if (lane == 0) {
        for (int w = 1; w < kMatchWarps; ++w)
          if (redf[w] > bv || (redf[w] == bv && redi[w] < bi)) {
            bv = redf[w];
            bi = redi[w];
          }
        col4row[i] = bi;
      }
    }
    __syncthreads();
  } else {
    for (int j = lane; j < nc; j += NT) {
      v[j] = 0.0;
      row4col[j] = -1;
    }
    for (int i = lane; i < nr; i += NT) {
      u[i] = 0.0;
      col4row[i] = -1;
    }
    __syncthreads();
    for (int cur = 0; cur < nr; ++cur) {
      for (int j = lane; j < nc; j += NT) {
        spc[j] = CUDART_INF;
        path[j] = -1;
        sc[j] = 0;
      }
      for (int i = lane; i < nr; i += NT) sr[i] = 0;
      __syncthreads();
      double min_val = 0.0;
      int i = cur, sink = -1;
      while (sink < 0) {
        if (lane == 0) sr[i] = 1;
        const double ui = u[i];
        const float* row = P + (long)i * K;
        Best b{CUDART_INF, 0, 0x7fffffff};
        for (int j = lane; j < nc; j += NT) {
          if (sc[j]) continue;
          const double r = ((min_val + (-(double)row[j])) - ui) - v[j];
          double s = spc[j];
          if (r < s) {
            path[j] = i;
            spc[j] = r;
            s = r;
          }
          const Best c{s, row4col[j] < 0 ? 1 : 0, j};
          if (better(c, b)) b = c;
        }
        b = block_best(b, red);
        min_val = b.val;
        if (!(min_val < CUDART_INF)) {  // infeasible (cannot happen with finite costs): leave the row unmatched
          sink = -2;
          break;
        }
        const int j = b.idx;
        if (lane == 0) sc[j] = 1;
        __syncthreads();
        if (row4col[j] < 0) sink = j;
        else i = row4col[j];
      }
      if (sink >= 0) {
        // dual update
        if (lane == 0) u[cur] += min_val;
        __syncthreads();
        for (int r = lane; r < nr; r += NT)
          if (sr[r] && r != cur) u[r] += min_val - spc[col4row[r]];
        for (int j = lane; j < nc; j += NT)
          if (sc[j]) v[j] -= min_val - spc[j];
        __syncthreads();
        // augment along the path
        if (lane == 0) {
          int j = sink;
          while (true) {
            const int r = path[j];
            row4col[j] = r;
            const int prev = col4row[r];
            col4row[r] = j;
            j = prev;
            if (r == cur) break;
          }
        }
        __syncthreads();
      }
    }
  }

  // relabel the queries and score the task
  __syncthreads();
  int hits = 0;
  for (int q = lane; q < n; q += NT) {
    const int c = sample_cluster[(long)t * n + q];
    const int lab = (c >= 0 && c < nr) ? col4row[c] : -1;
    if (new_labels) new_labels[(long)t * n + q] = lab;
    if (y_q) hits += (y_q[(long)t * n + q] == (long long)lab) ? 1 : 0;
  }
  for (int c = lane; c < n; c += NT) cluster_class[(long)t * n + c] = c < nr ? col4row[c] : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
  __syncthreads();
  if ((lane & 31) == 0) redi[lane >> 5] = hits;
  __syncthreads();
  if (lane == 0 && acc) {
    int total = 0;
    for (int w = 0; w < kMatchWarps; ++w) total += redi[w];
    acc[t] = (float)total / (float)n;
  }
}

size_t match_smem_bytes(int n, int K) {
  size_t b = sizeof(double) * ((size_t)2 * K + n) + sizeof(int) * ((size_t)2 * K + n) + (size_t)K + n;
  return (b + 15) & ~(size_t)15;
}

}  // namespace

cudaError_t match_clusters(const float* proto, const int* n_clusters, const int* sample_cluster, const long long* y_q,
                           int graph_matching, int* cluster_class, long long* new_labels, float* acc, int T, int n, int K,
                           int proto_rows, cudaStream_t st) {
  const size_t smem = match_smem_bytes(n, K);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(match_clusters_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  match_clusters_kernel<<<T, kMatchThreads, smem, st>>>(proto, n_clusters, sample_cluster, y_q, graph_matching, cluster_class,
                                            new_labels, acc, n, K, proto_rows);
  note_launch();
  return cudaGetLastError();
}

}  // namespace tclip
