/* tclip_b200.h — C ABI of libtclip_b200.so: the B200 (sm_100a) implementation of transductive-CLIP's batched EM
 * inference loop (EM-Dirichlet / Hard EM-Dirichlet first; k-means family next).
 *
 * The reference (SegoleneMartin/transductive-CLIP) is pure Python/PyTorch and has no FFI; the entry points below are
 * what a binding for its `src/methods` hot path would call, one per stage of `run_method`, plus one fused driver.
 * Each declaration cites the reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, sizes, and a `void* stream` that is a `cudaStream_t` (NULL = default stream);
 *   - every function returns 0 (TCLIP_OK) or a negative TCLIP_ERR_* code; `tclip_last_error()` has the message;
 *   - nothing allocates user-visible memory: scratch comes from a caller-provided workspace whose size is returned
 *     by the matching `*_workspace_bytes` query; nothing synchronises the stream or the device;
 *   - all tensors are dense row-major float32 unless stated: T tasks, n queries, K classes, D feature dim, S support
 *     samples;  u [T,n,K], x/logz [T,n,D], alpha/y [T,K,D], v/colsum [T,K];  D <= 1024 for the M-step;
 *   - there is no CPU path: calling into a device that is not compute capability 10.x fails with TCLIP_ERR_DEVICE.
 */
#ifndef TCLIP_B200_H_
#define TCLIP_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCLIP_OK 0
#define TCLIP_ERR_INVALID (-1)   /* bad argument (null pointer, size out of range, D > 1024, ...) */
#define TCLIP_ERR_CUDA (-2)      /* a CUDA runtime call or kernel launch failed */
#define TCLIP_ERR_DEVICE (-3)    /* current device is not an sm_100 part */
#define TCLIP_ERR_WORKSPACE (-4) /* workspace missing or too small */

#define TCLIP_MM_DENSE 0      /* iterate every (task, class) row in every M-step, like the reference */
#define TCLIP_MM_SKIP_DEAD 1  /* iterate empty-cluster rows once, replay their cached criterion terms afterwards */

/* ---- library / device --------------------------------------------------------------------------------------- */
int tclip_version(void);                 /* 100 * major + minor; 102 = 1.2 (1.1 + spec_probe, spec_rows_cap, kmeans_run) */
const char* tclip_last_error(void);      /* message of the last failing call on this thread ("" if none) */
int tclip_device_check(int device);      /* TCLIP_OK iff `device` is compute capability 10.x (B200) */
int tclip_mm_max_dim(void);              /* largest D the M-step kernel supports (1024) */
int tclip_spec_rows_cap(void);           /* rows the few-rows M-step kernel of the skip-dead schedule handles (1480) */
long long tclip_launch_count(void);      /* kernels launched by this library in this process so far */

/* Roofline denominators for the M-step (it is FP32/MUFU issue bound, not HBM or tensor bound): launches a
 * register-only microbenchmark, which = 0: dependent FFMA chains (*ops_out = flop executed), which = 1: MUFU
 * rcp/sqrt/lg2 chains (*ops_out = MUFU operations), which = 2: packed FFMA2 chains (*ops_out = flop), which = 3: the
 * M-step's mix, 4 FFMA2 : 1 MUFU (*ops_out = FFMA2 thread-instructions).  Time it with events on `stream`.  sink: >= 4 bytes of device
 * memory (never written in practice).  No reference counterpart (measurement only). */
int tclip_probe_issue_rate(int which, float* sink, int n_blocks, int iters, double* ops_out, void* stream);

/* ---- stage entry points ------------------------------------------------------------------------------------- */

/* out = log(x + 1e-15), `count` elements.
 * Replaces torch.log(query + self.eps): src/methods/zero_shot/em_dirichlet.py:38,219;
 * few_shot/em_dirichlet.py:187-190. */
int tclip_log_features(const float* x, float* out, long long count, void* stream);

/* colsum[t,k] = sum_n u[t,n,k]; live[t,k] = colsum > 1e-15; v[t,k] = log(colsum/n + 1e-15) + 1.
 * `v` and `live` may be NULL.  Replaces cluster_sizes / nonzero_clusters (zero_shot/em_dirichlet.py:217-218) and
 * v_update (:145-151). */
int tclip_dirichlet_colsum_v(const float* u, float* colsum, float* v, int* live, int T, int n, int K, void* stream);

/* Moments y_cst.  Zero-shot (support_sum == NULL): y = sum_n u logz / max(colsum,1e-15), rows with
 * colsum <= 1e-15 filled with -10 (zero_shot/em_dirichlet.py:219-222).  Few-shot: y = (1/(support_count + colsum))
 * * (support_sum + sum_n u logz) (few_shot/em_dirichlet.py:196-200). */
int tclip_dirichlet_moments(const float* u, const float* logz, const float* colsum, const float* support_sum,
                            const float* support_count, float* y, int T, int n, int K, int D, void* stream);

/* The same moments on the tensor cores: u^T and (log z)^T are staged as K-major operands ([T,K,np], [T,D,np], np = n rounded
 * up to 4) of the 3 x TF32 tcgen05 kernel (fp32 round-to-nearest running sum outside the tensor core), the division / support
 * terms / -10 fill run in its epilogue.  Equally accurate at the level of alpha but measured slower than the CUDA-core kernel
 * (contraction length 75: DESIGN.md §3.6), so tclip_dirichlet_em_run uses it only when TCLIP_MOMENTS=tc is set (outer
 * iteration 0 and the few-shot setting).  `workspace`: tclip_dirichlet_moments_tc_workspace_bytes bytes, 256-byte aligned. */
size_t tclip_dirichlet_moments_tc_workspace_bytes(int T, int n, int K, int D);
int tclip_dirichlet_moments_tc(const float* u, const float* logz, const float* colsum, const float* support_sum,
                               const float* support_count, float* y, int T, int n, int K, int D, void* workspace,
                               size_t workspace_bytes, void* stream);

/* Few-shot, iteration-invariant: support_count[t,k] = #{s: y_s[t,s] == k}, support_sum[t,k,:] = sum of
 * log_support[t,s,:] over those s.  Replaces the [T,S,K,D] one-hot product of few_shot/em_dirichlet.py:182,199.
 * y_s is int64 [T,S]. */
int tclip_dirichlet_support_stats(const float* log_support, const long long* y_s, float* support_sum,
                                  float* support_count, int T, int S, int K, int D, void* stream);

/* The MM M-step on n_rows = T*K rows of length D: up to iter_mm iterations of
 *   c = |2(lnG(1) - lnG(a+1) + psi(a+1) a)/a^2| (pi^2/6 for a <= 1e-11);  b = psi(a+1) - psi(sum_d a) - c a - y;
 *   a <- (-b + sqrt(b^2 + 4c)) / (2c)
 * stopping after iteration l in {check_every, 2 check_every, ...} iff ||a_new-a||^2/||a||^2 < tol over ALL rows
 * (batch-global); the result is the last a_new.  alpha_out may alias alpha_in.  *iters_done_dev (device int) receives
 * the number of iterations executed.  Replaces curvature + update_alpha (zero_shot/em_dirichlet.py:153-177). */
size_t tclip_dirichlet_mm_workspace_bytes(int n_rows);
int tclip_dirichlet_mm(const float* alpha_in, float* alpha_out, const float* y, int n_rows, int D, int iter_mm,
                       int check_every, float tol, int* iters_done_dev, void* workspace, size_t workspace_bytes,
                       void* stream);

/* alpha[row] <- work[row] for live rows (live == NULL: all rows), and the logged outer criterion
 * *criterion = mean_t ||alpha_old - alpha||_F / ||alpha_old||_F; task_criterion [T] may be NULL.
 * `rowstat` is scratch of T*K*16 bytes.  Replaces zero_shot/em_dirichlet.py:224-226,236-238. */
int tclip_dirichlet_commit(float* alpha, const float* work, const int* live, void* rowstat, float* task_criterion,
                           float* criterion, int T, int K, int D, void* stream);

/* E-step: u = softmax_k(lnG(sum_d a) - sum_d lnG(a) + sum_d (a-1) logz + lambd v / n); hard != 0 replaces u by the
 * one-hot of argmax_k u.  labels [T,n] int32 (may be NULL) always receives argmax_k of the softmaxed values, lowest
 * index on ties.  `norm` is scratch of T*K*8 bytes.  Replaces get_logits + u_update
 * (zero_shot/em_dirichlet.py:28-40,132-143) and zero_shot/hard_em_dirichlet.py:256-258. */
int tclip_dirichlet_estep(const float* alpha, const float* logz, const float* v, float lambd, void* norm, float* u,
                          int* labels, int T, int n, int K, int D, int hard, void* stream);

/* The contraction inside the E-step on its own: l3[t,i,k] = sum_d logz[t,i,d] * (alpha[t,k,d] - 1), l3 [T,n,K].
 * mode 0: tcgen05 tensor cores, 3 x TF32 split precision, running sum in round-to-nearest fp32 outside the tensor core
 *         (what tclip_dirichlet_estep / tclip_dirichlet_em_run use; needs D % 4 == 0 and n <= 128, else TCLIP_ERR_INVALID);
 * mode 1: the same with the sum left to the tensor core's accumulator (measurement of its truncation only);
 * mode 2: the CUDA-core fp32 kernel (any shape).
 * Replaces `(torch.log(query + eps) * (alpha - 1)).sum(-1)` of get_logits (zero_shot/em_dirichlet.py:37-38). */
int tclip_dirichlet_contraction(const float* logz, const float* alpha, float* l3, int T, int n, int K, int D, int mode,
                                void* stream);

/* Inputs of the label matching: per task the clusters in order of first appearance among `labels`, their sizes, the
 * cluster index of every query, and proto[t,c,:] = mean raw feature of cluster c (rows c >= n_clusters[t] are 0).
 * All int outputs are int32; cluster_label/cluster_size/sample_cluster are [T,n], n_clusters [T], proto [T,n,D].
 * Replaces compute_acc_clustering's prototypes and the cost-matrix rows of compute_graph_matching
 * (zero_shot/em_dirichlet.py:61-70; src/utils.py:380-399). */
int tclip_cluster_prototypes(const int* labels, const float* feats, int* cluster_label, int* cluster_size,
                             int* sample_cluster, int* n_clusters, float* proto, int T, int n, int D, void* stream);

/* ---- soft k-means / hard k-means / EM-Gaussian (identity covariance): src/methods/zero_shot/{soft_kmeans,hard_kmeans,
 * em_gaussian}.py.  x [T,n,D] features, u [T,n,K], w [T,K,D] centroids, v [T,K], text [K,D] unit text embeddings. ---------- */

/* out[r,:] = x[r,:] / ||x[r,:]||.  Replaces `query[task] / query[task].norm(dim=-1, keepdim=True)`
 * (soft_kmeans.py:192-193) and the prototype normalisation of compute_acc_clustering (soft_kmeans.py:53-54). */
int tclip_normalize_rows(const float* x, float* out, long long rows, int D, void* stream);

/* u[m,:] = softmax_k(scale * a[m,:] . text[k,:]), m < M (all tasks flattened).  Replaces the initial assignment on visual
 * features `(T * image_features @ text_features.T).softmax(-1)` (soft_kmeans.py:194-197; hard_kmeans.py:180-183;
 * em_gaussian.py:199-203) and the prototype probabilities of compute_acc_clustering (soft_kmeans.py:55-56). */
int tclip_kmeans_similarity(const float* a, const float* text, float scale, float* u, long long M, int K, int D,
                            void* stream);

/* w[t,k,:] = sum_n u[t,n,k] x[t,n,:] / max(sum_n u, 1e-15) for clusters with sum_n u > 1e-15; empty clusters are zeroed
 * (mode 0: hard_kmeans.py:138-151, and w_init, soft_kmeans.py:135-148) or keep their row (mode 1: soft_kmeans.py:150-166,
 * em_gaussian.py:153-169, em_gaussian_cov.py:153-170).  mode 2 = KL k-means (kl_kmeans.py:166-171): divide by
 * max(size, 1), zero iff size == 0. */
int tclip_kmeans_centroids(const float* u, const float* x, float* w, int T, int n, int K, int D, int mode,
                           void* stream);

/* Diagonal precisions s[t,k,d] = sum_n u / max(sum_n u (w - x)^2, 1e-15); with keep_old != 0 empty clusters keep their row
 * (s_update, em_gaussian_cov.py:182-193), keep_old == 0 is s_init (:172-180). */
int tclip_kmeans_precisions(const float* u, const float* x, const float* w, float* s, int T, int n, int K, int D,
                            int keep_old, void* stream);

/* EM-Gaussian with diagonal covariance: u = softmax_k(-1/2 sum_d s (w - x)^2 + 1/2 sum_d log(s + 1e-15) + lambd v / n)
 * (em_gaussian_cov.py:106-130).  det is scratch of T*K floats. */
int tclip_kmeans_assign_cov(const float* x, const float* w, const float* s, const float* v, float lambd, float* det,
                            float* u, int* labels, int T, int n, int K, int D, void* stream);

/* KL k-means: u = one-hot(argmin_k sum_d p log(p / q)), p = x + 1e-15, q = w + 1e-15; NaN counts as the minimum, lowest
 * index on ties, like torch.argmin (kl_kmeans.py:123-127,173-177). */
int tclip_kmeans_assign_kl(const float* x, const float* w, float* u, int* labels, int T, int n, int K, int D,
                           void* stream);

#define TCLIP_KMEANS_SOFT 0   /* u = softmax(T * (-1/2 d2))                       soft_kmeans.py:105-125 */
#define TCLIP_KMEANS_GAUSS 1  /* u = softmax(T * (-1/2 d2) + lambd v / n)         em_gaussian.py:106-128 */
#define TCLIP_KMEANS_HARD 2   /* u = one-hot(argmin_k softmax(+d2))               hard_kmeans.py:26-35,127-136,197-199 */
/* d2[t,n,k] = sum_d (w[t,k,d] - x[t,n,d])^2 (direct difference, as the reference), then the assignment of `mode`.
 * labels [T,n] int32 (may be NULL) = arg-max of the final u (hard: the arg-min cluster). */
int tclip_kmeans_assign(const float* x, const float* w, const float* v, float temperature, float lambd, int mode, float* u,
                        int* labels, int T, int n, int K, int D, void* stream);

/* task_norm[t] = ||a[t] - b[t]||_F over `per_task` elements, *mean_out = mean_t.  Replaces the logged criterion
 * `(u_old - self.u).norm(dim=(1, 2)).mean(0)` (hard_kmeans.py:201). */
int tclip_kmeans_udiff(const float* a, const float* b, float* task_norm, float* mean_out, int T, long long per_task,
                       void* stream);

/* ---- fused driver of the k-means family: the whole run_method loop of soft k-means (soft_kmeans.py:168-220), EM-Gaussian
 * with identity covariance (em_gaussian.py:171-229) or hard k-means (hard_kmeans.py:153-211), up to but excluding the
 * accuracy, enqueued on one stream without host synchronisation.  With n_query <= 96 the loop runs in the
 * coordinates of the task's own samples (Cholesky factor of the n x n Gram matrix; every centroid is a combination of the
 * task's samples and only distances to those samples are ever needed): same direct-difference distances as the reference,
 * in <= n dimensions instead of D, and w itself is never formed inside the loop — `coef` receives its coefficients and
 * tclip_kmeans_expand_centroids (or a non-NULL `w`) turns them into w [T,K,D].  Otherwise the feature-space kernels run
 * (tclip_kmeans_centroids / tclip_kmeans_assign) and `w` holds the centroids (`coef` is not written). */
typedef struct tclip_kmeans_problem {
  int n_task, n_query, n_class, dim; /* T, n, K, D */
  int iters;                         /* args.iter */
  int method;                        /* TCLIP_KMEANS_SOFT / TCLIP_KMEANS_GAUSS / TCLIP_KMEANS_HARD */
  float temperature;                 /* args.T */
  float lambd;                       /* EM-Gaussian: int(K/5) * n_query (em_gaussian.py:20); ignored otherwise */
  const float* x;                    /* [T,n,D] query features */
  float* u;                          /* in: initial assignment [T,n,K] (soft_kmeans.py:185-197), out: final u */
  float* v;                          /* out [T,K] (EM-Gaussian), may be NULL otherwise */
  int* labels;                       /* out [T,n] arg-max_k u (hard k-means: the arg-min cluster) */
  float* coef;                       /* out [T,n,K]: w[t,k,:] = sum_i coef[t,i,k] x[t,i,:] (sample-coordinate form) */
  float* w;                          /* optional out [T,K,D] */
  float* criterions;                 /* out [iters] (identically 0 upstream), hard k-means [2*iters] */
  void* const* iter_events;          /* optional [iters+1] cudaEvent_t: before the first and after every iteration */
} tclip_kmeans_problem;
int tclip_kmeans_sample_coordinates(int n_query, int dim);   /* 1 iff tclip_kmeans_run takes the sample-coordinate form */
size_t tclip_kmeans_workspace_bytes(const tclip_kmeans_problem* p);
int tclip_kmeans_run(const tclip_kmeans_problem* p, void* workspace, size_t workspace_bytes, void* stream);
/* w[t,k,:] = sum_i coef[t,i,k] x[t,i,:].  Replaces the [T,n,K,D] broadcast of w_update (soft_kmeans.py:161-166). */
int tclip_kmeans_expand_centroids(const float* coef, const float* x, float* w, int T, int n, int K, int D, void* stream);

/* Cluster -> class matching and the task accuracy.  probs [T, proto_rows, K] float32: row c of task t = class
 * probabilities of cluster c (clusters in order of first appearance, as tclip_cluster_prototypes delivers them);
 * n_clusters [T], sample_cluster [T,n] int32.  graph_matching != 0: minimum-cost assignment of the clusters to distinct
 * classes with cost -probs in float64 (SciPy's linear_sum_assignment algorithm, one warp per task); == 0: arg-max class
 * per cluster.  Outputs: cluster_class [T,n] int32 (-1 beyond n_clusters), new_labels [T,n] int64 (may be NULL), and, when
 * y_q [T,n] int64 is given, acc [T] = mean(new_labels == y_q).  Replaces compute_graph_matching / compute_basic_matching
 * (src/utils.py:380-417) and zero_shot/em_dirichlet.py:86-92. */
int tclip_match_clusters(const float* probs, const int* n_clusters, const int* sample_cluster, const long long* y_q,
                         int graph_matching, int* cluster_class, long long* new_labels, float* acc, int T, int n, int K,
                         int proto_rows, void* stream);

/* Device-side construction of a task batch from the cached features: x_q[m, :] = features[idx[m], :] (features [n_rows, F]
 * float32, idx [count] int64, count = T * n_query), y_q[m] = labels[idx[m]] (int64; labels / y_q may be NULL).  `bad`
 * (device int, may be NULL, caller zeroes it) counts indices outside [0, n_rows); their rows are written as zeros / -1.
 * Replaces all_features_query[indices, :] / all_labels_query[indices] (src/eval_zero_shot.py:158-163) and the torch.cat
 * of Tasks_Generator_zero_shot.generate_tasks (src/task_generator_zero_shot.py:36-65). */
int tclip_gather_tasks(const float* features, const long long* labels, const long long* idx, float* x_q, long long* y_q,
                       long long n_rows, long long count, int F, int* bad, void* stream);

/* The few-shot form of the same: Tasks_Generator_few_shot.get_task (src/task_generator_few_shot.py:27-58) relabels every
 * task and, with softmax features, re-orders the columns.  col_perm [T, U] int64 = the task's `unique_labels`, label_map
 * [T, n_labels] int64 = position of every label in that list (both built on the host with the reference's torch calls);
 * idx [T * per_task].  x_out[t, m, j] = features[idx[t, m], col_perm[t, j]] ([T, per_task, U]), y_out[t, m] =
 * label_map[t, labels[idx[t, m]]].  `bad` counts rows with an index, column or label out of range. */
int tclip_gather_tasks_remap(const float* features, const long long* labels, const long long* idx, const long long* col_perm,
                             const long long* label_map, float* x_out, long long* y_out, long long n_rows, long long count,
                             int per_task, int F, int U, int n_labels, int* bad, void* stream);

/* ---- fused driver: the whole run_method loop, enqueued on one stream without host synchronisation -------------- */
/* The caller keeps several batches in flight on different streams (tclip_b200.pipeline): the few-rows M-step kernel of the
 * skip-dead schedule then runs in its register-lean form (80 instead of 168 registers: six instead of three CTAs per SM, so
 * that the tails of all batches in flight fit on the GPU together).  Same results bit for bit; 1.7 % slower for one batch
 * alone, 9 % more tasks/s with 8 batches in flight. */
#define TCLIP_FLAG_IN_FLIGHT 1
/* Skip-dead schedule with few live rows: by default the soft-max of a query only visits the live classes once a bound proves
 * that every dead class underflows to exactly 0 (bit-identical to the soft-max over all K classes).  This bit
 * keeps the pass over all classes (cross-checks, measurements). */
#define TCLIP_FLAG_FULL_SOFTMAX 2
typedef struct tclip_dirichlet_problem {
  int n_task, n_query, n_class, dim; /* T, n, K, D (D == K: softmax features) */
  int n_support;                     /* S; 0 = zero-shot */
  int iters;                         /* outer EM iterations (args.iter) */
  int iter_mm;                       /* MM iterations per M-step (args.iter_mm) */
  int check_every;                   /* 50 in the reference */
  float tol;                         /* 1e-11 in the reference */
  float lambd;                       /* int(K/5)*n_query (zero-shot) or int(K/k_eff)*n_query (few-shot) */
  int hard;                          /* 1 = Hard EM-Dirichlet */
  int mm_mode;                       /* TCLIP_MM_DENSE / TCLIP_MM_SKIP_DEAD (zero-shot only) */
  const float* x_q;                  /* [T,n,D] query softmax features */
  const float* x_s;                  /* [T,S,D] support features or NULL */
  const long long* y_s;              /* [T,S] int64 support labels or NULL */
  float* u;                          /* out [T,n,K] */
  float* alpha;                      /* out [T,K,D] */
  float* v;                          /* out [T,K] */
  int* labels;                       /* out [T,n] argmax_k u */
  float* criterions;                 /* out [iters] */
  int* mm_iters;                     /* out [iters] MM iterations executed per outer iteration */
  int* n_live;                       /* out [iters] non-empty clusters per outer iteration (all tasks) */
  long long* mm_rows;                /* out [iters] rows actually iterated x iterations (work done), may be NULL */
  void* const* iter_events;          /* optional [iters] cudaEvent_t recorded after each outer iteration */
  void* const* mm_events;            /* optional [2*iters] cudaEvent_t recorded before / after each M-step */
  double* mm_crit;                   /* optional out [iters][2]: (||a_new - a||^2, ||a||^2) over the whole batch at the last
                                        check point each M-step evaluated (the two norms of em_dirichlet.py:170-171) */
  int* spec_probe;                   /* optional out [iters][tclip_spec_rows_cap()][4] (measurement only; skip-dead schedule):
                                        per live row of the few-rows M-step kernel {MM iterations executed, iteration at
                                        which the row reached a bit-exact fixed point or -1, iteration at which a longer
                                        cycle was first seen or -1, its period}.  Non-NULL selects the statistics build of
                                        that kernel (same arithmetic and results, ~10 % slower). */
  int flags;                         /* TCLIP_FLAG_* bits, 0 = defaults */
} tclip_dirichlet_problem;

/* Runs zero_shot/em_dirichlet.py:195-244 (hard: zero_shot/hard_em_dirichlet.py:215-269) or, with n_support > 0,
 * few_shot/em_dirichlet.py:166-218 (hard: few_shot/hard_em_dirichlet.py:187-249), up to but excluding the accuracy. */
size_t tclip_dirichlet_em_workspace_bytes(const tclip_dirichlet_problem* p);
int tclip_dirichlet_em_run(const tclip_dirichlet_problem* p, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TCLIP_B200_H_ */
