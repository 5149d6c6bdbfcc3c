// tclip_kernels.cuh — internal C++ interface between the kernel translation units and the C-ABI (capi.cu).
// Nothing here is exported; the public surface is include/tclip_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace tclip {

// kernels launched by this library in this process (tclip_launch_count); bumped by every launcher
extern std::atomic<long long> g_launches;
inline void note_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Per-device facts (function attributes, SM count) are cached per device index: one process may drive several GPUs, and
// several host threads (tclip_b200.pipeline) may reach a launcher at once.  A slot is 0 until its fact is established.
constexpr int kMaxDevices = 64;
struct PerDeviceFlags {
  std::atomic<int> v[kMaxDevices];
  PerDeviceFlags() { for (auto& x : v) x.store(0, std::memory_order_relaxed); }
};
// current device index, or -1 (out of the cached range or no device): callers then skip the cache
inline int current_device_slot() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

// ---- Dirichlet MM M-step (dirichlet_mm.cu) --------------------------------------------------------------------
constexpr int kMMThreads = 128;    // 4 warps = 4 rows per CTA
constexpr int kMMMaxSlots = 32;    // register slots per lane => D <= 1024
#ifndef TCLIP_MM_MIN_BLOCKS
#define TCLIP_MM_MIN_BLOCKS 6
#endif
constexpr int kMMMinBlocks = TCLIP_MM_MIN_BLOCKS;    // CTAs per SM the M-step kernel is compiled for (<= 80 registers per thread)

struct MMState {          // lives in device memory; written only by the reset / decide kernels
  int done;               // 1 once the batch-global criterion fell below tol
  int iters_done;         // MM iterations executed so far in this M-step
  unsigned int ticket;    // CTAs of the running chunk that have finished (the last one folds the criterion)
  int fired;              // mm_spec path: index of the check point that met the criterion, -1 if none
  double last_num;        // ||a_new - a||^2 at the last check
  double last_den;        // ||a||^2      at the last check
};

struct MMLaunch {
  const float* alpha_in;  // [rows_total, D] state the M-step starts from
  float* alpha_out;       // [rows_total, D] working/result state (may alias alpha_in)
  const float* y;         // [rows_total, D] moments y_cst
  const int* row_list;    // optional: indices of the rows to iterate (nullptr = rows 0..n_rows-1)
  const int* n_rows_dev;  // optional: device-side row count (overrides n_rows)
  int n_rows;             // rows to iterate (upper bound when n_rows_dev is given)
  int D;
  int n_blocks;           // grid size = mm_num_blocks(n_rows) (persistent: <= kMMMinBlocks CTAs per SM)
  double2* partials;      // [n_blocks] scratch for the criterion
  MMState* state;
  // "free-running" mode (row_cache != nullptr): ignore the batch-global exit, run all iter_mm iterations and store
  // each row's own criterion terms per check point: row_cache[check * rows_total + row] = (||da||^2, ||a||^2)
  double2* row_cache;
  int n_checks;
  int rows_total;         // free-running mode: leading dimension of row_cache ([n_checks][rows_total])
  int* frozen;            // free-running mode, optional: [rows_total] period (in chunks) of rows proven periodic, 0 = still iterating
  float* snap;            // free-running mode: [rows_total, D] chunk-end snapshots used for the periodicity proof
  unsigned long long* work_ctr;  // optional: += row-iterations executed (work accounting for the roofline)
  // optional (needs row_list + n_rows_dev): device-side {n_rows, cap}; when n_rows <= cap the whole M-step is run by
  // mm_spec_kernel (one row per CTA, one launch) instead of the chunked one-warp-per-row kernel.
  const int* split_gate;
  int split_cap;
  double2* spec_terms;    // [n_checks][split_cap]
  float* spec_snap;       // [n_checks][split_cap][D]
  int4* spec_probe;       // optional [split_cap]: per-row convergence statistics (selects the measurement build of the kernel)
  bool spec_lean;         // register-lean few-rows kernel (several batches in flight): same results, more CTAs per SM
};

int mm_max_dim();
int mm_num_blocks(int n_rows, bool persistent = false);
cudaError_t mm_run(MMLaunch p, int iter_mm, int check_every, float tol, const double2* extra_checks, cudaStream_t st);

// ---- the rest of the Dirichlet EM loop (dirichlet_estep.cu) ---------------------------------------------------------
cudaError_t log_features(const float* x, float* out, long count, cudaStream_t st);
cudaError_t colsum_v(const float* u, float* colsum, float* v, int* live, int T, int n, int K, cudaStream_t st);
// Skip-dead schedule: row lists built on the device by classify_rows_kernel (capi.cu).  gate = {n_live, cap}: kernels
// take the row-wise form iff n_live <= cap, decided on the device.
struct SparseRows {
  const int* rows_live;
  const int* n_live;
  const int* rows_new;
  const int* n_new;
  const int* gate;
  int cap;
  // state of the exact sparse soft-max of estep_task_kernel (all device pointers; nullptr = always the full rows)
  int it;                 // outer iteration of this call
  const int* changed;     // != 0: the set of dead rows changed at this outer iteration
  const int* set_iter;    // last outer iteration at which it changed
  int* a_iter;            // outer iteration at which dead_max was last computed
  float* dead_max;        // [T, n] per query: max over the dead classes of float(norm + l3)
  int* last_full;         // [2, T] (by iteration parity) last outer iteration in which a query of the task took a full row
  int* last_dense;        // last outer iteration whose E-step ran through the dense kernels
};
// Tensor-core form of the dense moments: u^T and (log z)^T staged as [T, K, np] / [T, D, np] (np = n rounded up to 4)
struct MomentsTc {
  float* uT;            // scratch [T, K, np], rewritten by every call
  const float* logzT;   // [T, D, np], made once per batch (transpose_pad of logz)
  int np;
};
cudaError_t moments(const float* u, const float* logz, const float* colsum, const float* support_sum,
                    const float* support_count, float* y, int T, int n, int K, int D, const SparseRows* sp,
                    cudaStream_t st, const MomentsTc* tc = nullptr);
cudaError_t support_stats(const float* log_support, const long long* y_s, float* support_sum, float* support_count,
                          int T, int S, int K, int D, cudaStream_t st);
cudaError_t commit(float* alpha, const float* work, const int* live, const int* dead_age, double2* rowstat,
                   float* task_crit, float* crit_out, int T, int K, int D, cudaStream_t st);
cudaError_t estep(const float* alpha, const float* logz, const float* v, float lambd, double* norm, float* l3, float* u,
                  int* labels, int T, int n, int K, int D, int hard, const int* live, const SparseRows* sp,
                  cudaStream_t st);
// contraction_tc.cu: l3 = logz . (alpha - 1)^T on tcgen05 (3 x TF32, fp32 round-to-nearest running sum outside the tensor
// core); accumulate_in_tmem = true is the measurement-only variant that leaves the whole sum to the tensor core.
bool logits_tc_supported(int n, int K, int D);
cudaError_t logits_simt(const float* logz, const float* alpha, float* l3, int T, int n, int K, int D, const int* gate,
                        cudaStream_t st);   // the CUDA-core form (dirichlet_estep.cu), any shape
cudaError_t logits_tc(const float* logz, const float* alpha, float* l3, int T, int n, int K, int D, const int* gate,
                      bool accumulate_in_tmem, cudaStream_t st);
// the same kernel as a general batched "NT" product: C[t,m,k] = sum_d A[t,m,d] (B[tb,k,d] - b_shift), B shared by all
// tasks when b_tasks == 1 (needs D % 4 == 0 and 16-byte aligned operands)
// optional epilogue on the rows of C: mode 1 / 2 = the zero-shot / few-shot moments (rows are classes; see moments_tc)
struct TcEpilogue {
  int mode = 0;
  const float* colsum = nullptr;         // [T, M]
  const float* support_sum = nullptr;    // [T, M, N]  (mode 2)
  const float* support_count = nullptr;  // [T, M]     (mode 2)
};
cudaError_t gemm_nt_tc(const float* a, const float* b, float* c, int T, int M, int N, int D, int b_tasks, float b_shift,
                       const int* gate, bool accumulate_in_tmem, cudaStream_t st, const TcEpilogue* epilogue = nullptr);
// dst[t, c, r] = src[t, r, c] for r < R, 0 for R <= r < Rp (Rp = R rounded up to a multiple of 4: a TMA-addressable row pitch);
// the K-major operand layout the tensor-core kernel needs for a contraction over the LEADING index of src
cudaError_t transpose_pad(const float* src, float* dst, int T, int R, int C, int Rp, const int* gate, cudaStream_t st);
cudaError_t cluster_prototypes(const int* labels, const float* feats, int* cluster_label, int* cluster_size,
                               int* sample_cluster, int* n_clusters, float* proto, int T, int n, int D,
                               cudaStream_t st);

// matching.cu: rectangular assignment (clusters -> classes) per task on the device + relabelled accuracy
cudaError_t match_clusters(const float* proto, const int* n_clusters, const int* sample_cluster, const long long* y_q,
                           int graph_matching, int* cluster_class, long long* new_labels, float* acc, int T, int n, int K,
                           int proto_rows, cudaStream_t st);

// task_gather.cu: x_q[m, :] = features[idx[m], :], y_q[m] = labels[idx[m]] (device-side task construction)
cudaError_t gather_tasks(const float* features, const long long* labels, const long long* idx, float* x_q,
                         long long* y_q, long long n_rows, long long count, int F, int* bad, cudaStream_t st);

cudaError_t gather_tasks_remap(const float* features, const long long* labels, const long long* idx,
                               const long long* col_perm, const long long* label_map, float* x_out, long long* y_out,
                               long long n_rows, long long count, int per_task, int F, int U, int n_labels, int* bad,
                               cudaStream_t st);

// ---- soft / hard k-means and EM-Gaussian (kmeans.cu) -----------------------------------------------------------------
cudaError_t normalize_rows(const float* x, float* out, long rows, int D, cudaStream_t st);
cudaError_t kmeans_similarity(const float* a, const float* text, float scale, float* u, long M, int K, int D,
                              cudaStream_t st);
cudaError_t kmeans_centroids(const float* u, const float* x, float* w, int T, int n, int K, int D, int mode,
                             cudaStream_t st);
cudaError_t kmeans_precisions(const float* u, const float* x, const float* w, float* s, int T, int n, int K, int D,
                              int keep_old, cudaStream_t st);
cudaError_t kmeans_sqdist_cov(const float* x, const float* w, const float* s, float* d2s, float* det, int T, int n, int K,
                              int D, cudaStream_t st);
cudaError_t kmeans_kl_div(const float* x, const float* w, float* div, int T, int n, int K, int D, cudaStream_t st);
cudaError_t kmeans_sqdist(const float* x, const float* w, float* d2, int T, int n, int K, int D, cudaStream_t st);
cudaError_t kmeans_assign(const float* d2, const float* v, const float* bias, float temperature, float lambd, float* u,
                          int* labels, int T, int n, int K, int mode, cudaStream_t st);
cudaError_t kmeans_assign_hard_tracked(const float* d2, float* u, int* labels, double* row_sq, int T, int n, int K,
                                       cudaStream_t st);
cudaError_t kmeans_hard_criterions(const double* row_sq, float* task_norm, float* crit, int iters, int T, int n,
                                   cudaStream_t st);
cudaError_t kmeans_udiff(const float* a, const float* b, float* task_norm, float* mean_out, int T, long per_task,
                         cudaStream_t st);

// kmeans_run.cu: the whole soft k-means / EM-Gaussian / hard k-means loop on one stream, in the coordinates of the task's
// own samples when min(n, D) is small enough (kmeans_sample_coordinates), else with the feature-space kernels above
struct KMeansRun {
  int T, n, K, D;
  int iters;
  int method;             // 0 soft k-means, 1 EM-Gaussian, 2 hard k-means (TCLIP_KMEANS_*)
  float temperature;
  float lambd;
  const float* x;         // [T,n,D]
  float* u;               // [T,n,K] in: initial assignment, out: final
  float* v;               // [T,K] (method 1)
  int* labels;            // [T,n]
  float* coef;            // [T,n,K] w[t,k,:] = sum_n coef[t,n,k] x[t,n,:] (sample-coordinate form only)
  float* w;               // optional [T,K,D]
  float* criterions;      // [iters] ([2 iters] for method 2)
  void* const* iter_events;  // optional [iters + 1]
};
bool kmeans_sample_coordinates(int n, int D);
size_t kmeans_run_workspace_bytes(const KMeansRun& p);
cudaError_t kmeans_run(const KMeansRun& p, void* workspace, cudaStream_t st);
cudaError_t kmeans_expand_centroids(const float* coef, const float* x, float* w, int T, int n, int K, int D,
                                    cudaStream_t st);

// ---- issue-rate microbenchmarks used as roofline denominators (probe.cu) ------------------------------------------
cudaError_t probe_ffma(float* sink, int n_blocks, int iters, cudaStream_t st);   // 2 * 8 * 256 * iters * 64 flop / CTA... see probe.cu
cudaError_t probe_mufu(float* sink, int n_blocks, int iters, cudaStream_t st);
cudaError_t probe_ffma2(float* sink, int n_blocks, int iters, cudaStream_t st);  // packed 2 x fp32 FMA
cudaError_t probe_mix(float* sink, int n_blocks, int iters, cudaStream_t st);    // 4 FFMA2 : 1 MUFU

}  // namespace tclip
