// capi.cu — the C ABI of libtclip_b200.so (declared in include/tclip_b200.h) and the fused EM driver.
//
// The driver enqueues the whole `run_method` loop of the reference (src/methods/zero_shot/em_dirichlet.py:195-244 and
// its hard / few-shot siblings) on one stream with no host synchronisation: the MM early exit is a device-side flag.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>

#include "../../include/tclip_b200.h"
#include "tclip_kernels.cuh"

namespace tclip {
std::atomic<long long> g_launches{0};
}

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(TCLIP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define TCLIP_CUDA(call)                                   \
  do {                                                     \
    cudaError_t e__ = (call);                              \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);  \
  } while (0)

// one capability probe per device and process; the kernels are built for sm_100a only
int current_device_ok() {
  static int verdict[64] = {};
  int dev = 0;
  TCLIP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(TCLIP_ERR_DEVICE, "device index %d out of range", dev);
  if (verdict[dev] == 0) {
    int major = 0;
    TCLIP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    verdict[dev] = (major == 10) ? 1 : -1;
  }
  if (verdict[dev] < 0)
    return fail(TCLIP_ERR_DEVICE, "device %d is not compute capability 10.x; libtclip_b200 has no other code path", dev);
  return TCLIP_OK;
}

inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

// bump allocator over the caller's workspace
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += align_up(count * sizeof(T));
    return r;
  }
};

constexpr int kSplitCap = 1480;   // the few-rows M-step kernel (one row per CTA) handles up to this many live rows
constexpr int kSparseCap = 4096; // row-wise E-step kernels handle up to this many live rows
constexpr int kMaxChecks = 64;  // cached criterion terms per dead row (iter_mm / check_every must stay below)

struct EmWorkspace {
  float* logz;
  float* work;
  float* y;
  float* colsum;
  int* live;
  double* norm;
  double2* rowstat;
  double2* partials;
  tclip::MMState* state;
  tclip::MMState* state_free;  // never raised: dead-row trajectories ignore the batch-global exit
  float* task_crit;
  float* log_support;
  float* support_sum;
  float* support_count;
  float* logzT;         // [T, D, np] (log z)^T, K-major operand of the tensor-core moments
  float* uT;            // [T, K, np] u^T, rewritten before every tensor-core moments call
  int np;               // n rounded up to a multiple of 4 (TMA row pitch)
  // skip-dead mode
  int* cache_valid;
  double2* cache;       // [n_checks, rows]
  double2* extra;       // [n_checks] sum of the cached terms over all dead rows
  float* l3;            // [T,n,K] persistent contraction log z . (alpha-1)^T (only live columns are recomputed)
  int* dead_age;        // [rows] consecutive outer iterations the cluster has been empty
  int4* tile_counts;    // [ceil(rows / 1024)] per-tile {n_live, n_new, changed} of the row classification
  int* gate;            // {n_live, row cap}: device-side choice between dense and row-wise E-step kernels
  int* split_gate;      // {n_live, kSplitCap}: device-side choice of the few-rows M-step kernel
  double2* spec_terms;  // [n_checks][kSplitCap] criterion terms of the speculated rows
  float* spec_snap;     // [n_checks][kSplitCap][D] their states right after every check iteration
  int* frozen;          // [rows] dead rows proven periodic (mm_chunk_kernel)
  float* snap;          // [rows, D] periodicity snapshots
  int* list_live;
  int* list_new;
  int* counts;          // [0] = live rows, [1] = newly dead rows
  unsigned long long* work_ctr;  // row-iterations executed by the free-running pass of this outer iteration
  float* dead_max;      // [T, n] bound of the sparse soft-max (estep_task_kernel)
  int* last_full;       // [2, T]
  int* ss_state;        // {set_iter, a_iter, last_dense}
  size_t bytes;
};

int num_checks(int iter_mm, int check_every) {
  if (check_every <= 0) return 0;
  return (iter_mm - 1) / check_every;  // check points l = ce, 2ce, ... <= iter_mm - 1
}

EmWorkspace carve(const tclip_dirichlet_problem& p, void* ws) {
  const size_t T = p.n_task, n = p.n_query, K = p.n_class, D = p.dim, S = p.n_support;
  const size_t rows = T * K;
  Carver c(ws);
  EmWorkspace w{};
  w.logz = c.take<float>(T * n * D);
  w.work = c.take<float>(rows * D);
  w.y = c.take<float>(rows * D);
  w.colsum = c.take<float>(rows);
  w.live = c.take<int>(rows);
  w.norm = c.take<double>(rows);
  w.rowstat = c.take<double2>(rows);
  w.partials = c.take<double2>(std::max(tclip::mm_num_blocks((int)rows), kSplitCap));  // largest grid of any chunk kernel
  w.state = c.take<tclip::MMState>(1);
  w.state_free = c.take<tclip::MMState>(1);
  w.task_crit = c.take<float>(T);
  if (S > 0) {
    w.log_support = c.take<float>(T * S * D);
    w.support_sum = c.take<float>(rows * D);
    w.support_count = c.take<float>(rows);
  }
  w.np = (int)((n + 3) & ~size_t(3));
  w.logzT = c.take<float>(T * D * (size_t)w.np);
  w.uT = c.take<float>(T * K * (size_t)w.np);
  if (p.mm_mode == TCLIP_MM_SKIP_DEAD && S == 0) {
    const size_t nc = num_checks(p.iter_mm, p.check_every);
    w.cache_valid = c.take<int>(rows);
    w.cache = c.take<double2>(rows * (nc ? nc : 1));
    w.extra = c.take<double2>(nc ? nc : 1);
    w.l3 = c.take<float>(T * n * K);
    w.dead_age = c.take<int>(rows);
    w.gate = c.take<int>(2);
    w.split_gate = c.take<int>(2);
    w.spec_terms = c.take<double2>((nc ? nc : 1) * (size_t)kSplitCap);
    w.spec_snap = c.take<float>((nc ? nc : 1) * (size_t)kSplitCap * D);
    w.frozen = c.take<int>(rows);
    w.snap = c.take<float>(rows * D);
    w.list_live = c.take<int>(rows);
    w.list_new = c.take<int>(rows);
    w.counts = c.take<int>(4);
    w.tile_counts = c.take<int4>((rows + 1023) / 1024);
    w.work_ctr = c.take<unsigned long long>(1);
    w.dead_max = c.take<float>(T * n);
    w.last_full = c.take<int>(2 * T);
    w.ss_state = c.take<int>(4);
  }
  w.bytes = c.off;
  return w;
}

int validate(const tclip_dirichlet_problem* p) {
  if (!p) return fail(TCLIP_ERR_INVALID, "problem is NULL");
  if (p->n_task < 1 || p->n_query < 1 || p->n_class < 1 || p->dim < 1)
    return fail(TCLIP_ERR_INVALID, "n_task/n_query/n_class/dim must be >= 1");
  if (p->dim > tclip::mm_max_dim())
    return fail(TCLIP_ERR_INVALID, "dim %d exceeds the M-step kernel limit %d", p->dim, tclip::mm_max_dim());
  if (p->n_query > 1024) return fail(TCLIP_ERR_INVALID, "n_query %d > 1024", p->n_query);
  if (p->iters < 0 || p->iter_mm < 1) return fail(TCLIP_ERR_INVALID, "iters >= 0 and iter_mm >= 1 required");
  if (p->n_support < 0) return fail(TCLIP_ERR_INVALID, "n_support < 0");
  if ((long long)p->n_task * p->n_class > 0x7fffffffLL / 4) return fail(TCLIP_ERR_INVALID, "n_task * n_class too large");
  if (p->mm_mode != TCLIP_MM_DENSE && p->mm_mode != TCLIP_MM_SKIP_DEAD) return fail(TCLIP_ERR_INVALID, "bad mm_mode");
  if (num_checks(p->iter_mm, p->check_every) > kMaxChecks)
    return fail(TCLIP_ERR_INVALID, "iter_mm / check_every > %d", kMaxChecks);
  return TCLIP_OK;
}

// ---- small driver-side kernels --------------------------------------------------------------------------------
__global__ void fill_kernel(float* p, float v, long n) {
  const long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n && (reinterpret_cast<unsigned long long>(p) & 15ull) == 0) {
    *reinterpret_cast<float4*>(p + i) = make_float4(v, v, v, v);
  } else {
    for (long j = i; j < n && j < i + 4; ++j) p[j] = v;  // this thread's four elements only (tail / unaligned buffer)
  }
}
__global__ void init_sparse_softmax_kernel(int* s) {   // {set_iter, a_iter, last_dense}
  s[0] = 0;
  s[1] = -1;
  s[2] = -1;
}
__global__ void zero_int_kernel(int* p, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

// Stable compaction of the row classes (ascending row order, so every later summation order is fixed):
//   live rows                      -> list_live   (iterated in the main loop)
//   dead rows without a valid cache -> list_new    (full trajectory once, terms cached)
// Two passes over tiles of 1024 consecutive rows, one CTA per tile: classify_count_kernel leaves per-tile counts,
// classify_write_kernel adds up the counts of the tiles before its own and scatters its rows (ballot + popc inside a warp,
// a 32-entry shuffle scan across the warps).  counts = {n_live, n_new, changed}: `changed` says whether the set of dead
// rows or any cached term differs from the previous outer iteration (else the cached sums stay valid).
struct RowFlags {
  bool in, live, fresh;
  int age;
};

__device__ __forceinline__ RowFlags row_flags(const int* live, const int* cache_valid, const int* dead_age, int r, int rows) {
  RowFlags f;
  f.in = r < rows;
  f.live = f.in && live[r] != 0;
  f.age = f.in ? dead_age[r] : 0;
  f.fresh = f.in && !f.live && !cache_valid[r];
  return f;
}

__global__ void __launch_bounds__(1024)
classify_count_kernel(const int* __restrict__ live, const int* __restrict__ cache_valid, const int* __restrict__ dead_age,
                      int4* __restrict__ tile_counts, int rows) {
  __shared__ int wl[32], wn[32], wc[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RowFlags f = row_flags(live, cache_valid, dead_age, blockIdx.x * 1024 + threadIdx.x, rows);
  const unsigned bl = __ballot_sync(0xffffffffu, f.live), bn = __ballot_sync(0xffffffffu, f.fresh);
  const unsigned bc = __ballot_sync(0xffffffffu, f.in && (f.live != (f.age == 0)));  // was live and is not, or the reverse
  if (lane == 0) {
    wl[warp] = __popc(bl);
    wn[warp] = __popc(bn);
    wc[warp] = bc != 0u;
  }
  __syncthreads();
  if (warp == 0) {
    int a = wl[lane], b = wn[lane], c = wc[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c |= __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) tile_counts[blockIdx.x] = make_int4(a, b, c, 0);
  }
}

__global__ void __launch_bounds__(1024)
classify_write_kernel(const int* __restrict__ live, int* __restrict__ cache_valid, int* __restrict__ frozen,
                      int* __restrict__ dead_age, const int4* __restrict__ tile_counts, int* __restrict__ list_live,
                      int* __restrict__ list_new, int* __restrict__ counts, int* __restrict__ gate, int cap,
                      int* __restrict__ split_gate, int split_cap, unsigned long long* __restrict__ work_ctr, int rows,
                      int it, int* __restrict__ set_iter) {
  __shared__ int wl[32], wn[32];
  __shared__ int4 red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  // counts of the tiles before this one (x, y) and of all tiles (z, w; `changed` folded into the sign bit-free w2 below)
  int4 acc = make_int4(0, 0, 0, 0);
  int chg = 0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 1024) {
    const int4 c = tile_counts[i];
    if (i < (int)blockIdx.x) {
      acc.x += c.x;
      acc.y += c.y;
    }
    acc.z += c.x;
    acc.w += c.y;
    chg |= c.z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    chg |= __shfl_xor_sync(0xffffffffu, chg, o);
  }
  const int r = blockIdx.x * 1024 + threadIdx.x;
  const RowFlags f = row_flags(live, cache_valid, dead_age, r, rows);
  const unsigned bl = __ballot_sync(0xffffffffu, f.live), bn = __ballot_sync(0xffffffffu, f.fresh);
  if (lane == 0) {
    red[warp] = make_int4(acc.x, acc.y, acc.z, acc.w | (chg << 30));
    wl[warp] = __popc(bl);
    wn[warp] = __popc(bn);
  }
  __syncthreads();
  int base_l = 0, base_n = 0, tot_l = 0, tot_n = 0, changed = 0;
#pragma unroll 4
  for (int w = 0; w < 32; ++w) {  // every thread folds the 32 warp partials (broadcast reads)
    const int4 c = red[w];
    base_l += c.x;
    base_n += c.y;
    tot_l += c.z;
    tot_n += c.w & 0x3fffffff;
    changed |= c.w >> 30;
  }
  const int vl = wl[lane], vn = wn[lane];  // lane i holds the totals of warp i of this tile
  int sl = vl, sn = vn;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, sl, o), c = __shfl_up_sync(0xffffffffu, sn, o);
    if (lane >= o) {
      sl += a;
      sn += c;
    }
  }
  const int off_l = __shfl_sync(0xffffffffu, sl - vl, warp), off_n = __shfl_sync(0xffffffffu, sn - vn, warp);
  if (f.in) {
    dead_age[r] = f.live ? 0 : f.age + 1;
    if (f.live) {
      list_live[base_l + off_l + __popc(bl & lt)] = r;
      cache_valid[r] = 0;
    } else if (f.fresh) {
      list_new[base_n + off_n + __popc(bn & lt)] = r;
      cache_valid[r] = 1;
      frozen[r] = 0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *work_ctr = 0ull;
    counts[0] = tot_l;
    counts[1] = tot_n;
    counts[2] = (changed || tot_n > 0) ? 1 : 0;
    if (changed || tot_n > 0) *set_iter = it;   // the set of dead rows differs from the previous outer iteration's
    gate[0] = tot_l;            // rows the row-wise kernels have to recompute (newly dead rows only get their y filled)
    gate[1] = cap;
    split_gate[0] = tot_l;
    split_gate[1] = split_cap;
  }
}

// extra[c] = sum over dead rows of cache[row, c]  (one CTA per check, fixed order); kept from the previous outer iteration
// when neither the set of dead rows nor any cached term changed (classify_rows_kernel)
__global__ void __launch_bounds__(256)
sum_cache_kernel(const int* __restrict__ live, const double2* __restrict__ cache, double2* __restrict__ extra, int rows,
                 int n_checks, const int* __restrict__ changed) {
  if (!*changed) return;
  __shared__ double2 red[256];
  const int c = blockIdx.x;
  double2 acc = make_double2(0.0, 0.0);
  for (int r = threadIdx.x; r < rows; r += 256) {
    if (!live[r]) {
      const double2 v = cache[(long)c * rows + r];
      acc.x += v.x;
      acc.y += v.y;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) {
      red[threadIdx.x].x += red[threadIdx.x + w].x;
      red[threadIdx.x].y += red[threadIdx.x + w].y;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) extra[c] = red[0];
}

__global__ void __launch_bounds__(256) count_live_kernel(const int* __restrict__ live, int rows, int* out) {
  __shared__ int red[256];
  int c = 0;
  for (int r = threadIdx.x; r < rows; r += 256) c += live[r];
  red[threadIdx.x] = c;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

__global__ void record_work_kernel(const tclip::MMState* state, const int* counts, const unsigned long long* work_ctr,
                                   const int* split_gate, const int* n_live_dev, int rows, int iter_mm, int* mm_iters,
                                   int* n_live, long long* mm_rows, double* mm_crit) {
  const int done = state->iters_done;
  if (mm_crit) {
    mm_crit[0] = state->last_num;
    mm_crit[1] = state->last_den;
  }
  *mm_iters = done;
  *n_live = n_live_dev ? *n_live_dev : rows;
  if (mm_rows) {
    // row-iterations really executed: the free-running dead rows and the speculated live rows count their own (a
    // speculated row runs all iter_mm iterations whatever check fires), the chunked live rows run `done` iterations each
    const bool spec = split_gate && split_gate[0] <= split_gate[1];
    if (counts) *mm_rows = (spec ? 0ll : (long long)counts[0] * done) + (long long)*work_ctr;
    else *mm_rows = (long long)rows * done;
  }
}

}  // namespace

// ======================================================================================================================
extern "C" {

int tclip_version(void) { return 102; }  // 1.2: + problem.spec_probe, tclip_spec_rows_cap, tclip_kmeans_run

const char* tclip_last_error(void) { return g_last_error.c_str(); }

int tclip_mm_max_dim(void) { return tclip::mm_max_dim(); }

int tclip_spec_rows_cap(void) { return kSplitCap; }

long long tclip_launch_count(void) { return tclip::g_launches.load(std::memory_order_relaxed); }

int tclip_probe_issue_rate(int which, float* sink, int n_blocks, int iters, double* ops_out, void* stream) {
  if (!sink || n_blocks < 1 || iters < 1 || which < 0 || which > 3)
    return fail(TCLIP_ERR_INVALID, "tclip_probe_issue_rate: bad arguments");
  if (int rc = current_device_ok()) return rc;
  const double per_cta = 256.0 * 8.0 * 64.0 * (double)iters;  // threads x chains x unroll x iters (probe.cu)
  if (which == 0) {
    TCLIP_CUDA(tclip::probe_ffma(sink, n_blocks, iters, (cudaStream_t)stream));
    if (ops_out) *ops_out = 2.0 * per_cta * n_blocks;  // flop
  } else if (which == 1) {
    TCLIP_CUDA(tclip::probe_mufu(sink, n_blocks, iters, (cudaStream_t)stream));
    if (ops_out) *ops_out = per_cta * n_blocks;  // MUFU operations
  } else if (which == 2) {
    TCLIP_CUDA(tclip::probe_ffma2(sink, n_blocks, iters, (cudaStream_t)stream));
    if (ops_out) *ops_out = 4.0 * per_cta * n_blocks;  // flop (2 lanes x 2)
  } else {
    TCLIP_CUDA(tclip::probe_mix(sink, n_blocks, iters, (cudaStream_t)stream));
    if (ops_out) *ops_out = per_cta * n_blocks;  // FFMA2 instructions per thread-chain (+ 1/4 as many MUFU)
  }
  return TCLIP_OK;
}

int tclip_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return fail(TCLIP_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(TCLIP_ERR_DEVICE, "device %d not present (%d devices)", device, count);
  int major = 0, minor = 0;
  TCLIP_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  TCLIP_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10)
    return fail(TCLIP_ERR_DEVICE, "device %d is sm_%d%d; libtclip_b200 is built for sm_100a only", device, major, minor);
  return TCLIP_OK;
}

int tclip_log_features(const float* x, float* out, long long count, void* stream) {
  if (!x || !out || count < 0) return fail(TCLIP_ERR_INVALID, "tclip_log_features: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::log_features(x, out, (long)count, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_dirichlet_colsum_v(const float* u, float* colsum, float* v, int* live, int T, int n, int K, void* stream) {
  if (!u || !colsum || T < 1 || n < 1 || K < 1) return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_colsum_v: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::colsum_v(u, colsum, v, live, T, n, K, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_dirichlet_moments(const float* u, const float* logz, const float* colsum, const float* support_sum,
                            const float* support_count, float* y, int T, int n, int K, int D, void* stream) {
  if (!u || !logz || !colsum || !y || T < 1 || n < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_moments: bad arguments");
  if ((support_sum == nullptr) != (support_count == nullptr))
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_moments: support_sum and support_count go together");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::moments(u, logz, colsum, support_sum, support_count, y, T, n, K, D, nullptr, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_dirichlet_support_stats(const float* log_support, const long long* y_s, float* support_sum,
                                  float* support_count, int T, int S, int K, int D, void* stream) {
  if (!log_support || !y_s || !support_sum || !support_count || T < 1 || S < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_support_stats: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::support_stats(log_support, y_s, support_sum, support_count, T, S, K, D, (cudaStream_t)stream));
  return TCLIP_OK;
}

size_t tclip_dirichlet_moments_tc_workspace_bytes(int T, int n, int K, int D) {
  if (T < 1 || n < 1 || K < 1 || D < 1) return 0;
  const size_t np = ((size_t)n + 3) & ~size_t(3);
  return align_up(sizeof(float) * (size_t)T * D * np) + align_up(sizeof(float) * (size_t)T * K * np);
}

int tclip_dirichlet_moments_tc(const float* u, const float* logz, const float* colsum, const float* support_sum,
                               const float* support_count, float* y, int T, int n, int K, int D, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (!u || !logz || !colsum || !y || T < 1 || n < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_moments_tc: bad arguments");
  if ((support_sum == nullptr) != (support_count == nullptr))
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_moments_tc: support_sum and support_count go together");
  if (!workspace || workspace_bytes < tclip_dirichlet_moments_tc_workspace_bytes(T, n, K, D) ||
      (reinterpret_cast<unsigned long long>(workspace) & 255ull) != 0)
    return fail(TCLIP_ERR_WORKSPACE, "tclip_dirichlet_moments_tc: workspace too small or not 256-byte aligned");
  if (int rc = current_device_ok()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(workspace);
  const int np = (n + 3) & ~3;
  float* logzT = c.take<float>((size_t)T * D * np);
  float* uT = c.take<float>((size_t)T * K * np);
  TCLIP_CUDA(tclip::transpose_pad(logz, logzT, T, n, D, np, nullptr, st));
  const tclip::MomentsTc mtc{uT, logzT, np};
  TCLIP_CUDA(tclip::moments(u, logz, colsum, support_sum, support_count, y, T, n, K, D, nullptr, st, &mtc));
  return TCLIP_OK;
}

size_t tclip_dirichlet_mm_workspace_bytes(int n_rows) {
  if (n_rows < 1) return 0;
  return align_up(sizeof(double2) * (size_t)tclip::mm_num_blocks(n_rows)) + align_up(sizeof(tclip::MMState));
}

int tclip_dirichlet_mm(const float* alpha_in, float* alpha_out, const float* y, int n_rows, int D, int iter_mm,
                       int check_every, float tol, int* iters_done_dev, void* workspace, size_t workspace_bytes,
                       void* stream) {
  if (!alpha_in || !alpha_out || !y || n_rows < 1 || iter_mm < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_mm: bad arguments");
  if (D < 1 || D > tclip::mm_max_dim())
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_mm: D=%d outside [1, %d]", D, tclip::mm_max_dim());
  if (!workspace || workspace_bytes < tclip_dirichlet_mm_workspace_bytes(n_rows))
    return fail(TCLIP_ERR_WORKSPACE, "tclip_dirichlet_mm: workspace too small (%zu < %zu)", workspace_bytes,
                tclip_dirichlet_mm_workspace_bytes(n_rows));
  if (int rc = current_device_ok()) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  Carver c(workspace);
  tclip::MMLaunch l{};
  l.alpha_in = alpha_in;
  l.alpha_out = alpha_out;
  l.y = y;
  l.n_rows = n_rows;
  l.D = D;
  l.n_blocks = tclip::mm_num_blocks(n_rows);
  l.partials = c.take<double2>(l.n_blocks);
  l.state = c.take<tclip::MMState>(1);
  TCLIP_CUDA(tclip::mm_run(l, iter_mm, check_every, tol, nullptr, st));
  if (iters_done_dev)
    TCLIP_CUDA(cudaMemcpyAsync(iters_done_dev, &l.state->iters_done, sizeof(int), cudaMemcpyDeviceToDevice, st));
  return TCLIP_OK;
}

int tclip_dirichlet_commit(float* alpha, const float* work, const int* live, void* rowstat, float* task_criterion,
                           float* criterion, int T, int K, int D, void* stream) {
  if (!alpha || !work || !rowstat || !criterion || T < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_commit: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::commit(alpha, work, live, nullptr, (double2*)rowstat, task_criterion, criterion, T, K, D,
                           (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_dirichlet_estep(const float* alpha, const float* logz, const float* v, float lambd, void* norm, float* u,
                          int* labels, int T, int n, int K, int D, int hard, void* stream) {
  if (!alpha || !logz || !v || !norm || !u || T < 1 || n < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_estep: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::estep(alpha, logz, v, lambd, (double*)norm, nullptr, u, labels, T, n, K, D, hard, nullptr, nullptr,
                          (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_dirichlet_contraction(const float* logz, const float* alpha, float* l3, int T, int n, int K, int D, int mode,
                                void* stream) {
  if (!logz || !alpha || !l3 || T < 1 || n < 1 || K < 1 || D < 1 || mode < 0 || mode > 2)
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_contraction: bad arguments");
  if (mode != 2 && !tclip::logits_tc_supported(n, K, D))
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_contraction: the tensor-core form needs D %% 4 == 0 and n <= 128 (n=%d, D=%d)", n, D);
  if (int rc = current_device_ok()) return rc;
  if (mode == 2)
    TCLIP_CUDA(tclip::logits_simt(logz, alpha, l3, T, n, K, D, nullptr, (cudaStream_t)stream));
  else
    TCLIP_CUDA(tclip::logits_tc(logz, alpha, l3, T, n, K, D, nullptr, mode == 1, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_cluster_prototypes(const int* labels, const float* feats, int* cluster_label, int* cluster_size,
                             int* sample_cluster, int* n_clusters, float* proto, int T, int n, int D, void* stream) {
  if (!labels || !feats || !cluster_label || !cluster_size || !sample_cluster || !n_clusters || !proto || T < 1 ||
      n < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_cluster_prototypes: bad arguments");
  if (n > 1024) return fail(TCLIP_ERR_INVALID, "tclip_cluster_prototypes: n=%d > 1024", n);
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::cluster_prototypes(labels, feats, cluster_label, cluster_size, sample_cluster, n_clusters, proto,
                                       T, n, D, (cudaStream_t)stream));
  return TCLIP_OK;
}

// ---- k-means family -------------------------------------------------------------------------------------------------
int tclip_match_clusters(const float* probs, const int* n_clusters, const int* sample_cluster, const long long* y_q,
                         int graph_matching, int* cluster_class, long long* new_labels, float* acc, int T, int n, int K,
                         int proto_rows, void* stream) {
  if (!probs || !n_clusters || !sample_cluster || !cluster_class || T < 1 || n < 1 || K < 1 || proto_rows < 1 ||
      (y_q && !acc))
    return fail(TCLIP_ERR_INVALID, "tclip_match_clusters: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::match_clusters(probs, n_clusters, sample_cluster, y_q, graph_matching, cluster_class, new_labels, acc,
                                   T, n, K, proto_rows, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_gather_tasks(const float* features, const long long* labels, const long long* idx, float* x_q, long long* y_q,
                       long long n_rows, long long count, int F, int* bad, void* stream) {
  if (!features || !idx || !x_q || n_rows < 1 || count < 0 || F < 1 || (y_q && !labels))
    return fail(TCLIP_ERR_INVALID, "tclip_gather_tasks: bad arguments");
  if (count == 0) return TCLIP_OK;  // nothing to gather
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::gather_tasks(features, labels, idx, x_q, y_q, n_rows, count, F, bad, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_gather_tasks_remap(const float* features, const long long* labels, const long long* idx, const long long* col_perm,
                             const long long* label_map, float* x_out, long long* y_out, long long n_rows, long long count,
                             int per_task, int F, int U, int n_labels, int* bad, void* stream) {
  if (!features || !labels || !idx || !col_perm || !label_map || !x_out || n_rows < 1 || count < 0 || per_task < 1 ||
      F < 1 || U < 1 || n_labels < 1 || count % per_task != 0)
    return fail(TCLIP_ERR_INVALID, "tclip_gather_tasks_remap: bad arguments");
  if (count == 0) return TCLIP_OK;
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::gather_tasks_remap(features, labels, idx, col_perm, label_map, x_out, y_out, n_rows, count, per_task,
                                       F, U, n_labels, bad, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_normalize_rows(const float* x, float* out, long long rows, int D, void* stream) {
  if (!x || !out || rows < 1 || D < 1) return fail(TCLIP_ERR_INVALID, "tclip_normalize_rows: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::normalize_rows(x, out, (long)rows, D, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_kmeans_similarity(const float* a, const float* text, float scale, float* u, long long M, int K, int D,
                            void* stream) {
  if (!a || !text || !u || M < 0 || K < 1 || D < 1) return fail(TCLIP_ERR_INVALID, "tclip_kmeans_similarity: bad arguments");
  if (M == 0) return TCLIP_OK;
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_similarity(a, text, scale, u, (long)M, K, D, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_kmeans_centroids(const float* u, const float* x, float* w, int T, int n, int K, int D, int mode,
                           void* stream) {
  if (!u || !x || !w || T < 1 || n < 1 || K < 1 || D < 1 || mode < 0 || mode > 2)
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_centroids: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_centroids(u, x, w, T, n, K, D, mode, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_kmeans_precisions(const float* u, const float* x, const float* w, float* s, int T, int n, int K, int D,
                            int keep_old, void* stream) {
  if (!u || !x || !w || !s || T < 1 || n < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_precisions: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_precisions(u, x, w, s, T, n, K, D, keep_old, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_kmeans_assign_cov(const float* x, const float* w, const float* s, const float* v, float lambd, float* det,
                            float* u, int* labels, int T, int n, int K, int D, void* stream) {
  if (!x || !w || !s || !v || !det || !u || T < 1 || n < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_assign_cov: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_sqdist_cov(x, w, s, u, det, T, n, K, D, (cudaStream_t)stream));
  TCLIP_CUDA(tclip::kmeans_assign(u, v, det, 1.0f, lambd, u, labels, T, n, K, 4, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_kmeans_assign_kl(const float* x, const float* w, float* u, int* labels, int T, int n, int K, int D,
                           void* stream) {
  if (!x || !w || !u || T < 1 || n < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_assign_kl: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_kl_div(x, w, u, T, n, K, D, (cudaStream_t)stream));
  TCLIP_CUDA(tclip::kmeans_assign(u, nullptr, nullptr, 1.0f, 0.0f, u, labels, T, n, K, 5, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_kmeans_assign(const float* x, const float* w, const float* v, float temperature, float lambd, int mode, float* u,
                        int* labels, int T, int n, int K, int D, void* stream) {
  if (!x || !w || !u || T < 1 || n < 1 || K < 1 || D < 1 || mode < 0 || mode > 2 || (mode == 1 && !v))
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_assign: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_sqdist(x, w, u, T, n, K, D, (cudaStream_t)stream));
  TCLIP_CUDA(tclip::kmeans_assign(u, v, nullptr, temperature, lambd, u, labels, T, n, K, mode, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_kmeans_udiff(const float* a, const float* b, float* task_norm, float* mean_out, int T, long long per_task,
                       void* stream) {
  if (!a || !b || !task_norm || !mean_out || T < 1 || per_task < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_udiff: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_udiff(a, b, task_norm, mean_out, T, (long)per_task, (cudaStream_t)stream));
  return TCLIP_OK;
}

namespace {
int kmeans_validate(const tclip_kmeans_problem* p, tclip::KMeansRun* r) {
  if (!p) return fail(TCLIP_ERR_INVALID, "problem is NULL");
  if (p->n_task < 1 || p->n_query < 1 || p->n_class < 1 || p->dim < 1 || p->iters < 0)
    return fail(TCLIP_ERR_INVALID, "n_task/n_query/n_class/dim must be >= 1 and iters >= 0");
  if (p->method < 0 || p->method > 2) return fail(TCLIP_ERR_INVALID, "method must be TCLIP_KMEANS_SOFT / GAUSS / HARD");
  r->T = p->n_task; r->n = p->n_query; r->K = p->n_class; r->D = p->dim;
  r->iters = p->iters; r->method = p->method; r->temperature = p->temperature; r->lambd = p->lambd;
  r->x = p->x; r->u = p->u; r->v = p->v; r->labels = p->labels; r->coef = p->coef; r->w = p->w;
  r->criterions = p->criterions; r->iter_events = p->iter_events;
  return TCLIP_OK;
}
}  // namespace

int tclip_kmeans_sample_coordinates(int n_query, int dim) { return tclip::kmeans_sample_coordinates(n_query, dim) ? 1 : 0; }

size_t tclip_kmeans_workspace_bytes(const tclip_kmeans_problem* p) {
  tclip::KMeansRun r{};
  if (kmeans_validate(p, &r) != TCLIP_OK) return 0;
  return tclip::kmeans_run_workspace_bytes(r) + 256;
}

int tclip_kmeans_run(const tclip_kmeans_problem* p, void* workspace, size_t workspace_bytes, void* stream) {
  tclip::KMeansRun r{};
  if (int rc = kmeans_validate(p, &r)) return rc;
  const bool coords = tclip::kmeans_sample_coordinates(r.n, r.D);
  if (!p->x || !p->u || !p->labels || !p->criterions || (p->method == 1 && !p->v) || (coords && !p->coef))
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_run: x/u/labels/criterions (+ v for EM-Gaussian, coef in sample coordinates) must be set");
  const size_t need = tclip::kmeans_run_workspace_bytes(r);
  if (!workspace || workspace_bytes < need)
    return fail(TCLIP_ERR_WORKSPACE, "tclip_kmeans_run: workspace too small (%zu < %zu)", workspace_bytes, need);
  if ((reinterpret_cast<unsigned long long>(workspace) & 255ull) != 0)
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_run: workspace must be 256-byte aligned");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_run(r, workspace, (cudaStream_t)stream));
  return TCLIP_OK;
}

int tclip_kmeans_expand_centroids(const float* coef, const float* x, float* w, int T, int n, int K, int D, void* stream) {
  if (!coef || !x || !w || T < 1 || n < 1 || K < 1 || D < 1)
    return fail(TCLIP_ERR_INVALID, "tclip_kmeans_expand_centroids: bad arguments");
  if (int rc = current_device_ok()) return rc;
  TCLIP_CUDA(tclip::kmeans_expand_centroids(coef, x, w, T, n, K, D, (cudaStream_t)stream));
  return TCLIP_OK;
}

size_t tclip_dirichlet_em_workspace_bytes(const tclip_dirichlet_problem* p) {
  if (validate(p) != TCLIP_OK) return 0;
  return carve(*p, nullptr).bytes;
}

int tclip_dirichlet_em_run(const tclip_dirichlet_problem* p, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = validate(p)) return rc;
  if (!p->x_q || !p->u || !p->alpha || !p->v || !p->criterions || !p->mm_iters || !p->n_live)
    return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_em_run: x_q/u/alpha/v/criterions/mm_iters/n_live must be set");
  const bool few = p->n_support > 0;
  if (few && (!p->x_s || !p->y_s)) return fail(TCLIP_ERR_INVALID, "tclip_dirichlet_em_run: few-shot needs x_s and y_s");
  const size_t need = carve(*p, nullptr).bytes;
  if (!workspace || workspace_bytes < need)
    return fail(TCLIP_ERR_WORKSPACE, "tclip_dirichlet_em_run: workspace too small (%zu < %zu)", workspace_bytes, need);
  if (int rc = current_device_ok()) return rc;

  cudaStream_t st = (cudaStream_t)stream;
  const int T = p->n_task, n = p->n_query, K = p->n_class, D = p->dim, S = p->n_support;
  const int rows = T * K;
  const bool skip = (p->mm_mode == TCLIP_MM_SKIP_DEAD) && !few;
  const int nc = num_checks(p->iter_mm, p->check_every);
  EmWorkspace w = carve(*p, workspace);

  // initialisation: v = 0, u = query, alpha = 1 (em_dirichlet.py:202-210); features logged once
  TCLIP_CUDA(cudaMemsetAsync(p->v, 0, sizeof(float) * (size_t)rows, st));
  TCLIP_CUDA(cudaMemcpyAsync(p->u, p->x_q, sizeof(float) * (size_t)T * n * K, cudaMemcpyDeviceToDevice, st));
  fill_kernel<<<(unsigned)(((long)rows * D + 1023) / 1024), 256, 0, st>>>(p->alpha, 1.0f, (long)rows * D);
  tclip::note_launch();
  TCLIP_CUDA(tclip::log_features(p->x_q, w.logz, (long)T * n * D, st));
  if (few) {
    TCLIP_CUDA(tclip::log_features(p->x_s, w.log_support, (long)T * S * D, st));
    TCLIP_CUDA(tclip::support_stats(w.log_support, p->y_s, w.support_sum, w.support_count, T, S, K, D, st));
  }
  // The moments u^T log z have a tensor-core form (tclip_dirichlet_moments_tc).  Measured (profiles/r2_moments_tc.md): the
  // alpha it leads to is as close to float64 as with the CUDA-core kernel, but it is SLOWER — 0.5 against 0.4 ms per dense
  // call: the contraction is only n = 75 long, so a 128 x 128 output tile is three 32-deep blocks and the kernel is all
  // prologue, epilogue and the transposition of u — so the CUDA-core kernel stays the default.  TCLIP_MOMENTS=tc selects the
  // tensor-core form where no row-wise kernel has to reproduce the result bit for bit: outer iteration 0 (every schedule
  // takes the dense form there) and the few-shot setting (always dense).
  static const bool moments_simt = [] {
    const char* e = std::getenv("TCLIP_MOMENTS");
    return !(e && std::string(e) == "tc");
  }();
  // TCLIP_SPARSE_SOFTMAX=0: the sparse regime always soft-maxes over all classes (measurements, cross-checks)
  static const bool sparse_softmax_off = [] {
    const char* e = std::getenv("TCLIP_SPARSE_SOFTMAX");
    return e && std::string(e) == "0";
  }();
  // the per-task E-step kernel keeps a task's class list in a 1024-entry shared array: more classes than that always take the
  // dense E-step (a cap of 0 never selects the row-wise kernels)
  const int sparse_cap = (K <= 1024) ? kSparseCap : 0;
  const tclip::MomentsTc mtc{w.uT, w.logzT, w.np};
  if (!moments_simt) TCLIP_CUDA(tclip::transpose_pad(w.logz, w.logzT, T, n, D, w.np, nullptr, st));
  if (skip) {
    zero_int_kernel<<<(rows + 255) / 256, 256, 0, st>>>(w.cache_valid, rows);
    zero_int_kernel<<<(rows + 255) / 256, 256, 0, st>>>(w.dead_age, rows);
    tclip::note_launch(2);
    TCLIP_CUDA(cudaMemsetAsync(w.state_free, 0, sizeof(tclip::MMState), st));
    TCLIP_CUDA(cudaMemsetAsync(w.last_full, 0xff, sizeof(int) * 2 * (size_t)T, st));          // -1: never
    init_sparse_softmax_kernel<<<1, 1, 0, st>>>(w.ss_state);
    tclip::note_launch();
    TCLIP_CUDA(cudaMemsetAsync(w.extra, 0, sizeof(double2) * (size_t)(nc ? nc : 1), st));  // no dead rows yet
  }

  for (int it = 0; it < p->iters; ++it) {
    // cluster sizes of the current u, live mask, and v (v_update uses the same u, em_dirichlet.py:230)
    TCLIP_CUDA(tclip::colsum_v(p->u, w.colsum, p->v, few ? nullptr : w.live, T, n, K, st));
    // skip-dead: row lists first (moments, M-step and E-step all work from them); from the second outer iteration on the
    // E-step side touches live / newly dead rows only, everything of an empty cluster carries over unchanged
    tclip::SparseRows sp{};
    const bool sparse = skip && it > 0;
    if (skip) {
      const int n_tiles = (rows + 1023) / 1024;
      classify_count_kernel<<<n_tiles, 1024, 0, st>>>(w.live, w.cache_valid, w.dead_age, w.tile_counts, rows);
      classify_write_kernel<<<n_tiles, 1024, 0, st>>>(w.live, w.cache_valid, w.frozen, w.dead_age, w.tile_counts, w.list_live,
                                                     w.list_new, w.counts, w.gate, sparse_cap, w.split_gate, kSplitCap,
                                                     w.work_ctr, rows, it, w.ss_state);
      tclip::note_launch(2);
      sp.rows_live = w.list_live;
      sp.n_live = w.counts;
      sp.rows_new = w.list_new;
      sp.n_new = w.counts + 1;
      sp.gate = w.gate;
      sp.cap = kSparseCap;   // (grid size of the row-wise kernels; the gate itself compares with sparse_cap)
      sp.it = it;
      if (!sparse_softmax_off && !(p->flags & TCLIP_FLAG_FULL_SOFTMAX)) {
        sp.changed = w.counts + 2;
        sp.set_iter = w.ss_state;
        sp.a_iter = w.ss_state + 1;
        sp.last_dense = w.ss_state + 2;
        sp.dead_max = w.dead_max;
        sp.last_full = w.last_full;
      }
    }
    TCLIP_CUDA(tclip::moments(p->u, w.logz, w.colsum, w.support_sum, w.support_count, w.y, T, n, K, D,
                              sparse ? &sp : nullptr, st, (!moments_simt && (it == 0 || few)) ? &mtc : nullptr));

    if (p->mm_events && p->mm_events[2 * it]) TCLIP_CUDA(cudaEventRecord((cudaEvent_t)p->mm_events[2 * it], st));
    tclip::MMLaunch l{};
    l.alpha_in = p->alpha;
    l.alpha_out = w.work;
    l.y = w.y;
    l.D = D;
    l.partials = w.partials;
    l.state = w.state;
    if (!skip) {
      l.n_rows = rows;
      l.n_blocks = tclip::mm_num_blocks(rows);
      TCLIP_CUDA(tclip::mm_run(l, p->iter_mm, p->check_every, p->tol, nullptr, st));
    } else {
      // newly dead rows: full trajectory from their kept row with y = -10, criterion terms cached per check
      tclip::MMLaunch d = l;
      d.row_list = w.list_new;
      d.n_rows_dev = w.counts + 1;
      d.n_rows = rows;
      d.n_blocks = tclip::mm_num_blocks(rows);
      d.state = w.state_free;
      d.row_cache = w.cache;
      d.n_checks = nc;
      d.rows_total = rows;
      d.frozen = w.frozen;
      d.snap = w.snap;
      d.work_ctr = w.work_ctr;
      TCLIP_CUDA(tclip::mm_run(d, p->iter_mm, p->check_every, p->tol, nullptr, st));
      if (nc > 0) {
        sum_cache_kernel<<<nc, 256, 0, st>>>(w.live, w.cache, w.extra, rows, nc, w.counts + 2);
        tclip::note_launch();
      }
      l.row_list = w.list_live;
      l.n_rows_dev = w.counts;
      l.split_gate = w.split_gate;
      l.split_cap = kSplitCap;
      l.spec_terms = w.spec_terms;
      l.spec_snap = w.spec_snap;
      l.work_ctr = w.work_ctr;
      l.spec_probe = p->spec_probe ? reinterpret_cast<int4*>(p->spec_probe) + (size_t)it * kSplitCap : nullptr;
      l.spec_lean = (p->flags & TCLIP_FLAG_IN_FLIGHT) != 0;
      l.n_rows = rows;
      l.n_blocks = tclip::mm_num_blocks(rows, true);
      TCLIP_CUDA(tclip::mm_run(l, p->iter_mm, p->check_every, p->tol, nc > 0 ? w.extra : nullptr, st));
    }
    if (p->mm_events && p->mm_events[2 * it + 1])
      TCLIP_CUDA(cudaEventRecord((cudaEvent_t)p->mm_events[2 * it + 1], st));
    int* n_live_dev = nullptr;
    if (skip) {
      n_live_dev = w.counts;  // classify_rows_kernel counted them
    } else if (!few) {
      count_live_kernel<<<1, 256, 0, st>>>(w.live, rows, p->n_live + it);
      tclip::note_launch();
      n_live_dev = p->n_live + it;
    }
    record_work_kernel<<<1, 1, 0, st>>>(w.state, skip ? w.counts : nullptr, w.work_ctr, skip ? w.split_gate : nullptr,
                                         n_live_dev, rows, p->iter_mm,
                                         p->mm_iters + it, p->n_live + it, p->mm_rows ? p->mm_rows + it : nullptr,
                                         p->mm_crit ? p->mm_crit + 2 * it : nullptr);
    tclip::note_launch();
    // empty clusters keep their previous row; logged criterion (em_dirichlet.py:224-226,236-238)
    TCLIP_CUDA(tclip::commit(p->alpha, w.work, few ? nullptr : w.live, skip ? w.dead_age : nullptr, w.rowstat,
                             w.task_crit, p->criterions + it, T, K, D, st));
    // u <- softmax(logits + lambda v / n) [-> one-hot]
    TCLIP_CUDA(tclip::estep(p->alpha, w.logz, p->v, p->lambd, w.norm, skip ? w.l3 : nullptr, p->u, p->labels, T, n, K, D,
                            p->hard, sparse ? w.live : nullptr, sparse ? &sp : nullptr, st));
    if (p->iter_events && p->iter_events[it]) TCLIP_CUDA(cudaEventRecord((cudaEvent_t)p->iter_events[it], st));
  }
  TCLIP_CUDA(cudaGetLastError());
  return TCLIP_OK;
}

}  // extern "C"
