// kmeans_run.cu — soft k-means / hard k-means / EM-Gaussian (identity covariance): the whole run_method loop enqueued on
// one stream, in the coordinates of the task's own samples.
//
// Reference (SegoleneMartin/transductive-CLIP, src/methods/zero_shot/): soft_kmeans.py:105-125,135-220,
// hard_kmeans.py:26-35,127-211, em_gaussian.py:106-136,138-229.  Per iteration the reference forms two [T,n,K,D]
// broadcasts: the centroids w = u^T x / sum u (w_update) and the squared distances ||w_k - x_n||^2 (get_logits).
//
// B200-first formulation.  Every centroid the loop ever holds is a linear combination of the n query samples of its task
// (w_init and every w_update are; an empty cluster keeps an earlier combination or is zeroed), and every quantity the loop
// needs from w is a distance to one of those samples.  With G = X X^T = L L^T (n x n Gram matrix of the task, Cholesky
// factor L), sample n is row n of L in an orthonormal basis of span{x}, a centroid is wt_k = sum_n c_kn L_n, and
//     ||w_k - x_n||^2 = || wt_k - L_n ||^2
// exactly: the same direct-difference form as the reference (no ||w||^2 + ||x||^2 - 2 x.w cancellation), in r = rank(X) <= n
// dimensions instead of D.  At RN50 shape (n = 75, D = 1024, K = 1000) that is 13.6x fewer flops, and the state of a task
// is [K, 80] + [n, 80] floats (0.34 MB) instead of w [K, D] (4 MB): the loop leaves HBM altogether (SURVEY.md §8(d) bounds
// the w-space loop by 8 MB of HBM traffic per task and iteration).  The coefficients c ("coef", laid out like u) are kept
// so that w = coef^T x can be produced when somebody asks for it (tclip_kmeans_expand_centroids).
//
// When D <= n the features are used as they are (Z = x, r = D); with more than kMaxR = 96 samples per task the loop falls
// back to the feature-space kernels of kmeans.cu.
//
// One outer iteration of soft k-means / EM-Gaussian = ONE launch of kproj_iter_kernel<CHAIN = true>: u tile rebuilt from the
// previous launch's exponentials and per-tile row statistics (u_update), cluster sizes [and v_update], centroids in sample
// coordinates with the reference's empty-cluster rule, coefficients, squared distances, exponentials of the logits + row
// statistics for the next launch; u, labels and v are materialised once after the last iteration (finish_chain_kernel,
// colsum_v).  Hard k-means, whose
// logged criterion needs every u: kproj_iter_kernel<false> + assign_kernel (arg-min rows, kmeans.cu) + criterion.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "tclip_kernels.cuh"

namespace tclip {

namespace {

constexpr float kEps = 1e-15f;
constexpr int kMaxR = 96;     // largest coordinate count (and sample count) of the sample-coordinate form
// G[t] = X_t X_t^T in float64 (n <= kMaxR), lower triangle: kGramSplits CTAs per task, each over a contiguous range of the D
// columns walked in slabs of 32.  A slab is staged as float64, column-major (sample index contiguous), and a thread owns a
// 4 x 4 block of sample pairs on or below the diagonal: per column 4 128-bit shared loads feed 16 DFMA, and the threads of a
// warp (consecutive blocks of one block row) read consecutive 32-byte pieces.  Every entry is the same d-ascending fma chain
// per split as a one-entry-per-thread loop would give; the partial matrices G[t][split] are added in split order by
// chol_kernel (a fixed order: the factor is reproducible bit for bit).
constexpr int kGramSplits = 4;
constexpr int kGramPitch = kMaxR + 2;                                       // doubles per staged column (even: 16-byte loads)
constexpr int kGramBlocks = (kMaxR / 4) * (kMaxR / 4 + 1) / 2;              // 4 x 4 blocks of the lower triangle at n = kMaxR
constexpr int kGramPerThread = (kGramBlocks + 255) / 256;

__global__ void __launch_bounds__(256)
gram_kernel(const float* __restrict__ x, double* __restrict__ G, int n, int D) {
  __shared__ __align__(16) double xs[32][kGramPitch];
  const int t = blockIdx.x, split = blockIdx.y;
  const float* xb = x + (long)t * n * D;
  const int nb = (n + 3) / 4, n_blocks = nb * (nb + 1) / 2;
  int bi[kGramPerThread], bj[kGramPerThread];
  double acc[kGramPerThread][4][4];
#pragma unroll
  for (int q = 0; q < kGramPerThread; ++q) {
    const int blk = threadIdx.x + 256 * q;
    // block row of triangular index blk: the largest i with i (i + 1) / 2 <= blk
    int i = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= blk) ++i;
    while (i * (i + 1) / 2 > blk) --i;
    bi[q] = i;
    bj[q] = blk - i * (i + 1) / 2;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[q][a][b] = 0.0;
  }
  const int slabs = (D + 31) / 32, per = (slabs + kGramSplits - 1) / kGramSplits;
  const int d_lo = split * per * 32, d_hi = min(D, (split + 1) * per * 32);
  const int n4 = 4 * nb;   // rows n .. n4 - 1 are zero
  for (int d0 = d_lo; d0 < d_hi; d0 += 32) {
    for (int i = threadIdx.x; i < n4 * 32; i += 256) {
      const int r = i >> 5, c = i & 31;
      xs[c][r] = (r < n && d0 + c < D) ? (double)xb[(long)r * D + d0 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kGramPerThread; ++q) {
      if (threadIdx.x + 256 * q < n_blocks) {
#pragma unroll 4
        for (int c = 0; c < 32; ++c) {
          const double2 a01 = *reinterpret_cast<const double2*>(&xs[c][4 * bi[q]]);
          const double2 a23 = *reinterpret_cast<const double2*>(&xs[c][4 * bi[q] + 2]);
          const double2 b01 = *reinterpret_cast<const double2*>(&xs[c][4 * bj[q]]);
          const double2 b23 = *reinterpret_cast<const double2*>(&xs[c][4 * bj[q] + 2]);
          const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[q][a][b] = fma(av[a], bv[b], acc[q][a][b]);
        }
      }
    }
    __syncthreads();
  }
  double* gb = G + ((long)t * kGramSplits + split) * n * n;
#pragma unroll
  for (int q = 0; q < kGramPerThread; ++q) {
    if (threadIdx.x + 256 * q < n_blocks) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = 4 * bi[q] + a, j = 4 * bj[q] + b;
          if (i < n && j <= i) gb[i * n + j] = acc[q][a][b];
        }
    }
  }
}

// Z[t] = Cholesky factor of G[t] (lower triangular, float64 arithmetic, stored as float32 [n, zs], zero above the diagonal
// and in the padding columns).  A pivot that is not above 1e-9 of its original diagonal entry marks a sample that is a
// linear combination of earlier ones (G is positive SEMI-definite in general: duplicated samples): its column is zero,
// which keeps L L^T = G to rounding accuracy.
__global__ void __launch_bounds__(256)
chol_kernel(const double* __restrict__ G, float* __restrict__ Z, int n, int zs) {
  extern __shared__ double A[];   // [n][n]
  __shared__ double piv;
  __shared__ double diag[kMaxR];
  const int t = blockIdx.x;
  const double* gb = G + (long)t * kGramSplits * n * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < n; i += 8)   // the lower triangle is all the factorisation reads (and all gram_kernel writes)
    for (int k = lane; k <= i; k += 32) {
      double s = gb[i * n + k];
#pragma unroll
      for (int sp = 1; sp < kGramSplits; ++sp) s += gb[(long)sp * n * n + i * n + k];
      A[i * n + k] = s;
      if (k == i) diag[i] = s;
    }
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      const double p = A[j * n + j], orig = diag[j];
      piv = (p > 1e-9 * orig && p > 0.0) ? sqrt(p) : 0.0;
    }
    __syncthreads();
    const double d = piv;
    for (int i = j + threadIdx.x; i < n; i += 256) A[i * n + j] = (d > 0.0) ? (i == j ? d : A[i * n + j] / d) : 0.0;
    __syncthreads();
    if (d > 0.0) {   // trailing update, lower triangle: one warp per row
      for (int i = j + 1 + warp; i < n; i += 8) {
        const double lij = A[i * n + j];
        for (int k = j + 1 + lane; k <= i; k += 32) A[i * n + k] -= lij * A[k * n + j];
      }
    }
    __syncthreads();
  }
  float* zb = Z + (long)t * n * zs;
  for (int p = threadIdx.x; p < n * zs; p += 256) {
    const int i = p / zs, j = p - i * zs;
    zb[p] = (j <= i && j < n) ? (float)A[i * n + j] : 0.0f;
  }
}

// Asynchronous global -> shared copies (cp.async, SASS LDGSTS); both addresses aligned to the request size
__device__ __forceinline__ void cp_async_4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// One outer iteration's M-step and distances for a tile of KT classes of one task, in sample coordinates.
//   Z    [T, n, zs]   coordinates of the samples (Cholesky rows, or the features themselves), r used columns
//   u    [T, n, K]    responsibilities the centroids are formed from
//   coef [T, n, K]    out: u / max(colsum, eps) for the clusters whose centroid is (re)formed; kept / zeroed otherwise
//   wt   [T, K, rq]   in/out: centroids in sample coordinates
//   d2   [T, n, K]    out (when want_d2): || wt_k - Z_n ||^2
// mode 0 (w_init, soft_kmeans.py:135-148): w = sum u x / max(sum u, eps) for every cluster, no mask;
// mode 1 (w_update of soft k-means / EM-Gaussian, soft_kmeans.py:150-166): clusters with sum u <= eps keep their centroid;
// mode 2 (hard k-means, hard_kmeans.py:138-151): those clusters are zeroed.
// Thread tile: 4 class PAIRS x MJ coordinates (centroids), 4 class pairs x MN samples (distances); all multiply-adds are
// packed FFMA2 / FADD2 over the class pair (sm_100 issues two fp32 lanes per instruction).  MJ = ceil(r / 16),
// MN = ceil(n / 16); rq = 16 MJ.
// KT classes per CTA (128: 4 class pairs per thread, 111 KB of shared memory = 2 CTAs per SM at RN50 shape; 64: 2 pairs, 72 KB =
// 3 CTAs per SM).  kUS: row pitch of the u tile (keeps float4 / float2 alignment, spreads the staging stores over the banks);
// kWS: row pitch of the transposed centroid tile (64-bit stores of 16 consecutive rows hit 16 bank pairs).
// The chained form (soft k-means, EM-Gaussian): from the logits of its distances,
//   l = T (-d2 / 2) [+ lambda v / n]                      (soft_kmeans.py:105-125, em_gaussian.py:106-128)
// the launch of iteration i leaves, per query and class tile, the tile's maximum m, the exponentials e = exp(l - m) and their
// sum s; the launch of iteration i + 1 turns them into its u tile itself, u = e exp(m - M) / S with M = max_tiles m and
// S = sum_tiles s exp(m - M) (one scale factor per query and tile: every exponential is computed once) — the launch boundary
// is the grid-wide barrier the soft-max over all K classes needs.  u, v and the labels are materialised once, after the last
// iteration; no assignment / column-sum launch and no u round trip inside the loop.  EM-Gaussian's v_update
// (em_gaussian.py:130-136) needs the column sums of u only, i.e. the cluster sizes the M-step forms anyway, per class: tile-local.
struct ChainArgs {
  const float2* stats_in;   // [T, n, tiles] (m, s) of the exponentials in `lg`; nullptr: the input is u (first iteration)
  float2* stats_out;        // [T, n, tiles] (the other buffer: CTAs of one task read all tiles while others write theirs)
  float* lg;                // [T, n, K] exponentials e, read (when stats_in) and rewritten in place: a CTA owns its class tile
  const float* v0;          // [T, K] EM-Gaussian, u input only: the v the first logits use (zeros); nullptr otherwise
  float temperature;
  float lambd;
  int method;               // 0 soft k-means, 1 EM-Gaussian
};

// TRI: Z is a Cholesky factor (lower triangular: sample n has no coordinate beyond n; MJ == MN).  In blocks of 16 the
// centroid sums then skip the samples below a coordinate block (they would add u * 0: bit-identical), and the distance of a
// sample group takes the coordinate blocks beyond it as the running sum of squares of the centroid alone, (0 - w)^2 = w^2:
// the same terms in another order.  57 % / 63 % of the multiply-adds of the two phases remain at n = 75.
template <int MJ, int MN, int KT, bool CHAIN, int NT, bool TRI>
__global__ void __launch_bounds__(NT)
kproj_iter_kernel(const float* __restrict__ Z, int zs, const float* __restrict__ u, float* __restrict__ coef,
                  float* __restrict__ wt, float* __restrict__ d2, int n, int K, int r, int mode, int want_d2,
                  const ChainArgs ch) {
  constexpr int RQ = 16 * MJ, NQ = 16 * MN, ZP = RQ + 1;
  constexpr int kUS = KT + 4, kWS = KT + 2, CPT = KT / (NT / 16), PQ = CPT / 2;   // classes / class pairs per thread
  extern __shared__ __align__(16) float sm[];
  float* Zs = sm;                    // [NQ][ZP]   samples (rows >= n and columns >= r are zero)
  float* us = Zs + ((NQ * ZP + 3) & ~3);   // [NQ][kUS]  u tile, later the d2 tile (16-byte aligned for the float4 reads)
  float* nwT = us + NQ * kUS;        // [RQ][kWS]  MINUS the centroids, coordinate-major (class pairs are contiguous)
  float* cs = nwT + RQ * kWS;        // [KT]      cluster sizes
  const int t = blockIdx.y, k0 = blockIdx.x * KT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* zb = Z + (long)t * n * zs;
  const bool from_logits = CHAIN && ch.stats_in != nullptr;
  const float* ub = (from_logits ? ch.lg : u) + (long)t * n * K;
  float* row_m = cs + KT;            // [NQ] chained form: exp(m_tile - M) / S of every query
  float* vt = row_m + 2 * NQ;        // [KT] EM-Gaussian: lambda v / n of this tile's classes
  // Every global request of the CTA is in flight before anything waits: the u / logits tile comes in by asynchronous copies
  // (LDGSTS, nothing staged in registers), the padding is zeroed by plain stores to the other addresses.  With a loop of
  // load -> store pairs this phase was a chain of ~10 dependent DRAM round trips and 32 % of the kernel's warp time
  // (profiles/r2_kmeans.md).
  // Samples: the Cholesky buffer has whole zero-padded rows of RQ floats — 128-bit loads into registers, issued first and
  // stored after the tile's copies are on their way (the odd row pitch of the shared tile, which keeps the distance loop
  // free of bank conflicts, rules out 16-byte async copies); any other Z: 4-byte async copies.
  constexpr int ZV = RQ / 4, ZPER = (NQ * ZV + NT - 1) / NT;
  const bool z_rows = zs == RQ && (reinterpret_cast<uintptr_t>(zb) & 15) == 0;
  float4 zr[ZPER];
  if (z_rows) {
#pragma unroll
    for (int q = 0; q < ZPER; ++q) {
      const int i = tid + NT * q;
      zr[q] = (i < n * ZV) ? __ldg(reinterpret_cast<const float4*>(zb) + i) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  } else {
    for (int i = tid; i < NQ * RQ; i += NT) {
      const int row = i / RQ, c = i - row * RQ;
      if (row < n && c < r) cp_async_4(Zs + row * ZP + c, zb + (long)row * zs + c);
      else Zs[row * ZP + c] = 0.0f;
    }
  }
  if ((K & 3) == 0 && (reinterpret_cast<uintptr_t>(ub) & 15) == 0) {   // 16-byte requests: rows and tiles start on 16 bytes
    for (int i = tid; i < NQ * (KT / 4); i += NT) {
      const int row = i / (KT / 4), c = 4 * (i - row * (KT / 4));
      if (row < n && k0 + c < K) cp_async_16(us + row * kUS + c, ub + (long)row * K + k0 + c);
      else *reinterpret_cast<float4*>(us + row * kUS + c) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  } else {
    for (int i = tid; i < NQ * KT; i += NT) {
      const int row = i / KT, c = i - row * KT;
      if (row < n && k0 + c < K) cp_async_4(us + row * kUS + c, ub + (long)row * K + k0 + c);
      else us[row * kUS + c] = 0.0f;
    }
  }
  if (z_rows) {
#pragma unroll
    for (int q = 0; q < ZPER; ++q) {
      const int i = tid + NT * q;
      if (i < NQ * ZV) {
        const int row = i / ZV, c = 4 * (i - row * ZV);
        float* dst = Zs + row * ZP + c;
        dst[0] = zr[q].x;
        dst[1] = zr[q].y;
        dst[2] = zr[q].z;
        dst[3] = zr[q].w;
      }
    }
  }
  if constexpr (CHAIN) {
    if (from_logits && tid < n) {   // in the shadow of the copies: the soft-max statistics of the query over all tiles
      const int tiles = gridDim.x;
      const float2* sp = ch.stats_in + ((long)t * n + tid) * tiles;
      float M = -CUDART_INF_F;
      for (int i = 0; i < tiles; ++i) M = fmaxf(M, sp[i].x);
      float S = 0.0f;
      for (int i = 0; i < tiles; ++i) S += sp[i].y * expf(sp[i].x - M);   // tile order: reproducible
      row_m[tid] = expf(sp[blockIdx.x].x - M) / S;   // the factor that turns this tile's exponentials into u
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if constexpr (CHAIN) {
    if (from_logits) {   // u tile = softmax row restricted to this tile (u_update); the padding stays zero
      for (int row = tid >> 5; row < n; row += NT / 32) {   // one warp per query
        const float f = row_m[row];
#pragma unroll
        for (int j = 0; j < KT / 32; ++j) {
          const int c = (tid & 31) + 32 * j;
          us[row * kUS + c] *= f;   // (the padding is zero and stays zero)
        }
      }
      __syncthreads();
    }
  }
  if (tid < KT) {   // cluster sizes in sample order, like u.sum(1)
    float s = 0.0f;
    for (int i = 0; i < n; ++i) s += us[i * kUS + tid];
    cs[tid] = s;
    if constexpr (CHAIN) {
      if (ch.method == 1) {   // v_update of the u this launch started from (log(colsum / n + eps) + 1), or the given first v
        const float v = from_logits ? logf(s / (float)n + kEps) + 1.0f : (k0 + tid < K ? ch.v0[(long)t * K + k0 + tid] : 0.0f);
        vt[tid] = (ch.lambd * v) / (float)n;
      }
    }
  }
  __syncthreads();
  // centroids of this tile: wt[k, j] = sum_n u[n, k] Z[n, j] / max(cs, eps)
  {
    float2 acc[PQ][MJ];
#pragma unroll
    for (int q = 0; q < PQ; ++q)
#pragma unroll
      for (int m = 0; m < MJ; ++m) acc[q][m] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int nb = 0; nb < (TRI ? MN : 1); ++nb) {   // TRI: sample block nb reaches the coordinate blocks m <= nb only
      const int n_lo = TRI ? 16 * nb : 0, n_hi = TRI ? min(n, 16 * nb + 16) : n;
#pragma unroll 4
      for (int nn = n_lo; nn < n_hi; ++nn) {
        float2 uv[PQ];
#pragma unroll
        for (int q = 0; q < PQ; q += 2) {
          const float4 u4 = *reinterpret_cast<const float4*>(us + nn * kUS + ty * CPT + 2 * q);
          uv[q] = make_float2(u4.x, u4.y);
          uv[q + 1] = make_float2(u4.z, u4.w);
        }
#pragma unroll
        for (int m = 0; m < MJ; ++m) {
          if (TRI && m > nb) continue;
          const float z = Zs[nn * ZP + tx + 16 * m];
          const float2 zz = make_float2(z, z);
#pragma unroll
          for (int q = 0; q < PQ; ++q) acc[q][m] = __ffma2_rn(uv[q], zz, acc[q][m]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < PQ; ++q) {
      const int kk = ty * CPT + 2 * q;
      // One reciprocal per class, a multiply per coordinate (these are sample coordinates: there is no reference value a
      // true division would reproduce bit for bit; the unrolled divisions and their branches were 26 % of the kernel's time).
      const float c0 = cs[kk], c1 = cs[kk + 1];
      const bool in0 = k0 + kk < K, in1 = k0 + kk + 1 < K;
      const bool f0 = in0 && ((mode == 0) || (c0 > kEps)), f1 = in1 && ((mode == 0) || (c1 > kEps));
      const bool keep0 = in0 && !f0 && mode == 1, keep1 = in1 && !f1 && mode == 1;   // empty cluster keeps its centroid
      const bool store0 = f0 || (in0 && mode == 2), store1 = f1 || (in1 && mode == 2);   // (mode 2: zeroed)
      const float2 inv = make_float2(f0 ? 1.0f / fmaxf(c0, kEps) : 0.0f, f1 ? 1.0f / fmaxf(c1, kEps) : 0.0f);
      float* w0 = wt + ((long)t * K + k0 + kk) * RQ;
      float* w1 = w0 + RQ;
#pragma unroll
      for (int m = 0; m < MJ; ++m) {
        const int j = tx + 16 * m;
        float2 v = __fmul2_rn(acc[q][m], inv);
        if (keep0) v.x = w0[j];
        if (keep1) v.y = w1[j];
        if (store0) w0[j] = v.x;
        if (store1) w1[j] = v.y;
        *reinterpret_cast<float2*>(nwT + j * kWS + kk) = make_float2(-v.x, -v.y);
      }
    }
  }
  // coefficients of the centroids in terms of the samples (what tclip_kmeans_expand_centroids turns into w)
  {
    // a thread keeps its class: one reciprocal of the cluster size, then a multiply per sample
    const int kk = tid & (KT - 1), k = k0 + kk;
    const float c = cs[kk];
    const bool formed = mode == 0 || c > kEps;
    if (k < K && (formed || mode == 2)) {
      const float inv = formed ? 1.0f / fmaxf(c, kEps) : 0.0f;
      float* cb = coef + (long)t * n * K + k;
      for (int row = tid / KT; row < n; row += NT / KT) cb[(long)row * K] = us[row * kUS + kk] * inv;
    }
  }
  if (!want_d2) return;
  __syncthreads();
  // distances: d2[n, k] = sum_j (Z[n, j] - wt[k, j])^2, the reference's direct-difference form
  float2 acc[PQ][MN];
#pragma unroll
  for (int q = 0; q < PQ; ++q)
#pragma unroll
    for (int m = 0; m < MN; ++m) acc[q][m] = make_float2(0.0f, 0.0f);
  if constexpr (!TRI) {
#pragma unroll 4
    for (int j = 0; j < r; ++j) {
      float2 nw[PQ];
#pragma unroll
      for (int q = 0; q < PQ; ++q) nw[q] = *reinterpret_cast<const float2*>(nwT + j * kWS + ty * CPT + 2 * q);
#pragma unroll
      for (int m = 0; m < MN; ++m) {
        const float z = Zs[(tx + 16 * m) * ZP + j];
        const float2 zz = make_float2(z, z);
#pragma unroll
        for (int q = 0; q < PQ; ++q) {
          const float2 df = __fadd2_rn(zz, nw[q]);
          acc[q][m] = __ffma2_rn(df, df, acc[q][m]);
        }
      }
    }
  } else {
    // coordinate blocks from the last to the first; sq = sum of w^2 over the blocks already walked, which is all that the
    // sample group entering at this block (m == jb: its samples are zero beyond coordinate 16 jb + 15) has of them
    float2 sq[PQ];
#pragma unroll
    for (int q = 0; q < PQ; ++q) sq[q] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int jb = MJ - 1; jb >= 0; --jb) {
#pragma unroll
      for (int q = 0; q < PQ; ++q) acc[q][jb] = sq[q];
      const int j_hi = min(r, 16 * jb + 16);
#pragma unroll 4
      for (int j = 16 * jb; j < j_hi; ++j) {
        float2 nw[PQ];
#pragma unroll
        for (int q = 0; q < PQ; ++q) nw[q] = *reinterpret_cast<const float2*>(nwT + j * kWS + ty * CPT + 2 * q);
#pragma unroll
        for (int m = 0; m < MN; ++m) {
          if (m < jb) continue;
          const float z = Zs[(tx + 16 * m) * ZP + j];
          const float2 zz = make_float2(z, z);
#pragma unroll
          for (int q = 0; q < PQ; ++q) {
            const float2 df = __fadd2_rn(zz, nw[q]);
            acc[q][m] = __ffma2_rn(df, df, acc[q][m]);
          }
        }
#pragma unroll
        for (int q = 0; q < PQ; ++q) sq[q] = __ffma2_rn(nw[q], nw[q], sq[q]);
      }
    }
  }
  // through shared memory (the u tile is dead) so that the rows go out coalesced
#pragma unroll
  for (int m = 0; m < MN; ++m)
#pragma unroll
    for (int q = 0; q < PQ; ++q) *reinterpret_cast<float2*>(us + (tx + 16 * m) * kUS + ty * CPT + 2 * q) = acc[q][m];
  __syncthreads();
  if constexpr (!CHAIN) {
    float* db = d2 + (long)t * n * K;
    for (int i = tid; i < n * KT; i += NT) {
      const int row = i / KT, kk = i - row * KT;
      if (k0 + kk < K) db[(long)row * K + k0 + kk] = us[row * kUS + kk];
    }
  } else {
    // logits of this tile, their per-query maximum, exponentials and sum (one warp per query)
    constexpr int LPT = KT / 32;
    const int lane = tid & 31, warp = tid >> 5;
    const bool gauss = ch.method == 1;
    for (int row = warp; row < n; row += NT / 32) {
      float l[LPT];
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < LPT; ++j) {
        const int kk = lane + 32 * j;
        l[j] = ch.temperature * (-0.5f * us[row * kUS + kk]);
        if (gauss) l[j] += vt[kk];
        if (k0 + kk >= K) l[j] = -CUDART_INF_F;
        mx = fmaxf(mx, l[j]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.0f;
      float* erow = ch.lg + ((long)t * n + row) * K + k0;   // the warp writes its query's 128 exponentials: 4 coalesced stores
#pragma unroll
      for (int j = 0; j < LPT; ++j) {
        const float e = expf(l[j] - mx);   // exp(-inf) = 0 for the padding classes (every tile holds at least one class)
        sum += e;
        if (k0 + lane + 32 * j < K) erow[lane + 32 * j] = e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) ch.stats_out[((long)t * n + row) * gridDim.x + blockIdx.x] = make_float2(mx, sum);
    }
  }
}

// After the last chained iteration: u [T,n,K] = e exp(m_tile - M) / S from the exponentials and tile statistics it left, and
// labels = arg-max_k u (first maximum).  One warp per query; same M and S (tile order) as the iteration kernel would form.
__global__ void __launch_bounds__(128)
finish_chain_kernel(const float* ex, const float2* __restrict__ stats, float* u, int* __restrict__ labels, int rows, int K,
                    int tiles, int kt) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float2* sp = stats + (long)row * tiles;
  float M = -CUDART_INF_F;
  for (int i = 0; i < tiles; ++i) M = fmaxf(M, sp[i].x);
  float S = 0.0f;
  for (int i = 0; i < tiles; ++i) S += sp[i].y * expf(sp[i].x - M);
  const float* e = ex + (long)row * K;
  float* out = u + (long)row * K;
  float best = -1.0f;
  int best_k = 0x7fffffff;
  for (int i = 0; i < tiles; ++i) {
    const float f = expf(sp[i].x - M) / S;
    const int k_hi = min(K, (i + 1) * kt);
    for (int k = i * kt + lane; k < k_hi; k += 32) {
      const float p = e[k] * f;
      if (p > best) {   // strict: the lowest k of this lane's classes wins ties
        best = p;
        best_k = k;
      }
      out[k] = p;
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
  if (lane == 0 && labels) labels[row] = best_k;
}

template <int MJ, int MN, int KT, int NT, bool TRI>
cudaError_t launch_iter_kt(const float* Z, int zs, const float* u, float* coef, float* wt, float* d2, int T, int n, int K,
                           int r, int mode, int want_d2, const ChainArgs* ch, cudaStream_t st) {
  constexpr int RQ = 16 * MJ, NQ = 16 * MN, ZP = RQ + 1, kUS = KT + 4, kWS = KT + 2;
  // samples, u tile, centroids, cluster sizes + row statistics and v of the chained form
  const size_t smem = sizeof(float) * ((size_t)NQ * ZP + 4 + (size_t)NQ * kUS + (size_t)RQ * kWS + KT + 2 * NQ + KT);
  static PerDeviceFlags attr_set;
  const int slot = current_device_slot();
  if (smem > 48 * 1024 && (slot < 0 || attr_set.v[slot].load(std::memory_order_acquire) == 0)) {
    cudaError_t e = cudaFuncSetAttribute(kproj_iter_kernel<MJ, MN, KT, false, NT, TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(kproj_iter_kernel<MJ, MN, KT, true, NT, TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (slot >= 0) attr_set.v[slot].store(1, std::memory_order_release);
  }
  const dim3 grid((K + KT - 1) / KT, T);
  if (ch) kproj_iter_kernel<MJ, MN, KT, true, NT, TRI><<<grid, NT, smem, st>>>(Z, zs, u, coef, wt, d2, n, K, r, mode, want_d2, *ch);
  else kproj_iter_kernel<MJ, MN, KT, false, NT, TRI><<<grid, NT, smem, st>>>(Z, zs, u, coef, wt, d2, n, K, r, mode, want_d2, ChainArgs{});
  note_launch();
  return cudaGetLastError();
}

// classes per CTA: 128; TCLIP_KM_TILE=64 selects the 64-class tile (3 CTAs per SM at RN50 shape instead of 2) — measured equal
// (0.172 vs 0.173 ms per iteration, profiles/r2_kmeans.md): the kernel is not limited by occupancy
int km_tile() {
  static const int tile = [] {
    const char* e = std::getenv("TCLIP_KM_TILE");
    return (e && std::atoi(e) == 64) ? 64 : 128;
  }();
  return tile;
}

// Soft k-means and EM-Gaussian run the chained form of the iteration kernel (ChainArgs); TCLIP_KM_CHAIN=0 keeps the separate
// assignment / column-sum launches per iteration (what hard k-means, whose logged criterion needs every u, always does).
// (Also tried: the class tiles of a task as one thread-block cluster, row statistics exchanged over distributed shared memory
// inside the launch — commit 2e92692, TCLIP_KM_FUSED there.  Same results, slower: 33 resident clusters of 8 x 107 KB CTAs,
// 4 quantised waves per 100 tasks, 224 us per iteration against 143 + 35; profiles/r2_kmeans.md.)
// TCLIP_KM_TRI=0: the iteration kernel treats the Cholesky factor as a dense matrix (measurement / cross-check)
bool kmeans_triangular() {
  static const bool on = [] {
    const char* e = std::getenv("TCLIP_KM_TRI");
    return !(e && std::atoi(e) == 0);
  }();
  return on;
}

bool kmeans_chained() {
  static const bool on = [] {
    const char* e = std::getenv("TCLIP_KM_CHAIN");
    return !(e && std::atoi(e) == 0);
  }();
  return on;
}

template <int MJ, int MN, bool TRI>
cudaError_t launch_iter(const float* Z, int zs, const float* u, float* coef, float* wt, float* d2, int T, int n, int K,
                        int r, int mode, int want_d2, const ChainArgs* ch, cudaStream_t st) {
  // (NT = 512, i.e. 32 instead of 16 resident warps per SM with half the register tile each, measures the same: loop 2.563 vs
  // 2.567 ms per batch — like the 64-class tile, and like scalar instead of packed multiply-adds: the kernel is bound by the
  // rate at which the FMA pipe takes three-register-operand instructions, not by latency; profiles/r2_kmeans.md)
  if (km_tile() == 128) return launch_iter_kt<MJ, MN, 128, 256, TRI>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
  return launch_iter_kt<MJ, MN, 64, 256, TRI>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
}

// r <= n always (r = min(n, D)): coordinate blocks MJ <= sample blocks MN
template <int MN>
cudaError_t launch_iter_mj(int mj, const float* Z, int zs, const float* u, float* coef, float* wt, float* d2, int T, int n,
                           int K, int r, int mode, int want_d2, const ChainArgs* ch, cudaStream_t st) {
#define TCLIP_KM_CASE(J) \
  if constexpr (J <= MN)  \
    if (mj == J) return launch_iter<J, MN, false>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
  TCLIP_KM_CASE(1) TCLIP_KM_CASE(2) TCLIP_KM_CASE(3) TCLIP_KM_CASE(4) TCLIP_KM_CASE(5) TCLIP_KM_CASE(6)
#undef TCLIP_KM_CASE
  return cudaErrorInvalidValue;
}

// tri: Z is the Cholesky buffer (lower triangular rows of zs = 16 ceil(n / 16) floats, r == n)
cudaError_t iterate(const float* Z, int zs, const float* u, float* coef, float* wt, float* d2, int T, int n, int K, int r,
                    int mode, int want_d2, const ChainArgs* ch, bool tri, cudaStream_t st) {
  const int mj = (r + 15) / 16, mn = (n + 15) / 16;
  if (mn < 1 || mn > 6 || mj > mn) return cudaErrorInvalidValue;
  if (tri) {
    if (mj != mn) return cudaErrorInvalidValue;
    switch (mn) {
      case 1: return launch_iter<1, 1, true>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
      case 2: return launch_iter<2, 2, true>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
      case 3: return launch_iter<3, 3, true>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
      case 4: return launch_iter<4, 4, true>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
      case 5: return launch_iter<5, 5, true>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
      default: return launch_iter<6, 6, true>(Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
    }
  }
  switch (mn) {
    case 1: return launch_iter_mj<1>(mj, Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
    case 2: return launch_iter_mj<2>(mj, Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
    case 3: return launch_iter_mj<3>(mj, Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
    case 4: return launch_iter_mj<4>(mj, Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
    case 5: return launch_iter_mj<5>(mj, Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
    default: return launch_iter_mj<6>(mj, Z, zs, u, coef, wt, d2, T, n, K, r, mode, want_d2, ch, st);
  }
}

// w[t,k,d] = sum_n coef[t,n,k] x[t,n,d]: 64 (k) x 64 (d) tile per CTA, the samples staged through shared memory
__global__ void __launch_bounds__(256)
expand_kernel(const float* __restrict__ coef, const float* __restrict__ x, float* __restrict__ w, int n, int K, int D) {
  constexpr int kTile = 64, kStage = 16;
  __shared__ float cs_[kStage][kTile + 4];
  __shared__ float xs[kStage][kTile + 4];
  const int t = blockIdx.z;
  const int k0 = blockIdx.y * kTile, d0 = blockIdx.x * kTile;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* cb = coef + (long)t * n * K;
  const float* xb = x + (long)t * n * D;
  float acc[4][4] = {};
  for (int n0 = 0; n0 < n; n0 += kStage) {
    for (int i = threadIdx.x; i < kStage * kTile; i += 256) {
      const int rr = i / kTile, c = i % kTile;
      const int nn = n0 + rr;
      cs_[rr][c] = (nn < n && k0 + c < K) ? cb[(long)nn * K + k0 + c] : 0.0f;
      xs[rr][c] = (nn < n && d0 + c < D) ? xb[(long)nn * D + d0 + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kStage; ++rr) {
      float cc[4], xx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cc[i] = cs_[rr][ty * 4 + i];
        xx[i] = xs[rr][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(cc[i], xx[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + tx * 4 + j;
      if (d < D) w[((long)t * K + k) * D + d] = acc[i][j];
    }
  }
}

inline size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

// the iteration kernel holds all n samples of a task (and min(n, D) coordinates) in shared memory
bool kmeans_sample_coordinates(int n, int D) { return n <= kMaxR && D >= 1; }

// rq: row pitch of the sample-coordinate arrays
static int coord_pitch(int n, int D) { return 16 * ((std::min(n, D) + 15) / 16); }

size_t kmeans_run_workspace_bytes(const KMeansRun& p) {
  const size_t T = p.T, n = p.n, K = p.K, D = p.D;
  size_t b = 0;
  b += align_up(sizeof(float) * T * n * K);              // d2
  const size_t its = (size_t)std::max(p.iters, 1);
  b += align_up(sizeof(float) * T * its);                // task norms (criterion)
  if (p.method == 2) b += align_up(sizeof(double) * T * n * its);   // per-query terms of the logged criteria
  if (p.method == 1) b += align_up(sizeof(float) * T * K);       // colsum scratch
  if (kmeans_sample_coordinates(p.n, p.D)) {
    const size_t rq = coord_pitch(p.n, p.D);
    if (p.method != 2) b += 2 * align_up(sizeof(float2) * T * n * ((K + km_tile() - 1) / km_tile()));   // row statistics
    b += align_up(sizeof(float) * T * K * rq);           // wt
    if (D > n) {
      b += align_up(sizeof(double) * T * kGramSplits * n * n);   // G (partial sums)
      b += align_up(sizeof(float) * T * n * rq);         // Z
    }
  } else if (!p.w) {
    b += align_up(sizeof(float) * T * K * D);            // w of the feature-space fallback
  }
  return b;
}

cudaError_t kmeans_expand_centroids(const float* coef, const float* x, float* w, int T, int n, int K, int D,
                                    cudaStream_t st) {
  expand_kernel<<<dim3((D + 63) / 64, (K + 63) / 64, T), 256, 0, st>>>(coef, x, w, n, K, D);
  note_launch();
  return cudaGetLastError();
}

#define KM_TRY(call)                       \
  do {                                     \
    cudaError_t e__ = (call);              \
    if (e__ != cudaSuccess) return e__;    \
  } while (0)

cudaError_t kmeans_run(const KMeansRun& p, void* workspace, cudaStream_t st) {
  const int T = p.T, n = p.n, K = p.K, D = p.D;
  char* base = static_cast<char*>(workspace);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* r = base + off;
    off += align_up(bytes);
    return r;
  };
  float* d2 = static_cast<float*>(take(sizeof(float) * (size_t)T * n * K));
  const size_t its = (size_t)std::max(p.iters, 1);
  float* task_norm = static_cast<float*>(take(sizeof(float) * (size_t)T * its));
  double* row_sq = p.method == 2 ? static_cast<double*>(take(sizeof(double) * (size_t)T * n * its)) : nullptr;
  float* colsum = p.method == 1 ? static_cast<float*>(take(sizeof(float) * (size_t)T * K)) : nullptr;
  const bool coords = kmeans_sample_coordinates(n, D);
  const bool chained = coords && p.method != 2 && kmeans_chained() && p.iters > 0;
  const float* Z = p.x;
  int zs = D, r = D;
  float *wt = nullptr, *w = p.w;
  float2* stats[2] = {nullptr, nullptr};
  bool tri = false;   // Z is a Cholesky factor
  if (coords) {
    const int rq = coord_pitch(n, D);
    wt = static_cast<float*>(take(sizeof(float) * (size_t)T * K * rq));
    if (p.method != 2) {
      const size_t sb = sizeof(float2) * (size_t)T * n * ((K + km_tile() - 1) / km_tile());
      stats[0] = static_cast<float2*>(take(sb));
      stats[1] = static_cast<float2*>(take(sb));
    }
    if (D > n) {
      double* G = static_cast<double*>(take(sizeof(double) * (size_t)T * kGramSplits * n * n));
      float* Zc = static_cast<float*>(take(sizeof(float) * (size_t)T * n * rq));
      gram_kernel<<<dim3(T, kGramSplits), 256, 0, st>>>(p.x, G, n, D);
      const size_t smem = sizeof(double) * (size_t)n * n;
      if (smem > 48 * 1024) KM_TRY(cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      chol_kernel<<<T, 256, smem, st>>>(G, Zc, n, rq);
      note_launch(2);
      Z = Zc;
      zs = rq;
      r = n;
      tri = kmeans_triangular();
    }
  } else if (!w) {
    w = static_cast<float*>(take(sizeof(float) * (size_t)T * K * D));
  }
  if (p.method == 1) KM_TRY(cudaMemsetAsync(p.v, 0, sizeof(float) * (size_t)T * K, st));
  // w_init (soft k-means, EM-Gaussian); hard k-means has none (hard_kmeans.py:186).  In sample coordinates it is folded into
  // the first iteration: w_init and the first w_update form the centroids of the SAME u, and where the masked update keeps
  // the old centroid (cluster size <= eps) the old one is w_init's unmasked value — so the first iteration simply runs
  // unmasked (mode 0) and leaves bit for bit what the two launches would.
  const bool fold_init = coords && p.method != 2 && p.iters > 0;
  if (p.method != 2 && !fold_init) {
    if (coords) KM_TRY(iterate(Z, zs, p.u, p.coef, wt, d2, T, n, K, r, 0, 0, nullptr, tri, st));
    else KM_TRY(kmeans_centroids(p.u, p.x, w, T, n, K, D, 0, st));
  }
  if (p.iter_events && p.iter_events[0]) KM_TRY(cudaEventRecord((cudaEvent_t)p.iter_events[0], st));
  for (int it = 0; it < p.iters; ++it) {
    const int mode = p.method == 2 ? 2 : ((fold_init && it == 0) ? 0 : 1);
    if (chained) {
      // M-step of the u the previous launch's logits describe, distances, logits and their row statistics: one launch
      const ChainArgs ch{it == 0 ? nullptr : stats[(it + 1) & 1], stats[it & 1], d2, p.method == 1 ? p.v : nullptr,
                         p.temperature, p.lambd, p.method};
      KM_TRY(iterate(Z, zs, p.u, p.coef, wt, d2, T, n, K, r, mode, 1, &ch, tri, st));
    } else {
      if (coords) {
        KM_TRY(iterate(Z, zs, p.u, p.coef, wt, d2, T, n, K, r, mode, 1, nullptr, tri, st));
      } else {
        KM_TRY(kmeans_centroids(p.u, p.x, w, T, n, K, D, mode == 2 ? 0 : 1, st));
        KM_TRY(kmeans_sqdist(p.x, w, d2, T, n, K, D, st));
      }
      if (p.method == 2) {
        // u_update in place; the per-query terms of || u_old - u || are taken as the old row is replaced (no copy of u is
        // kept), the criteria of all iterations are reduced after the loop
        KM_TRY(kmeans_assign_hard_tracked(d2, p.u, p.labels, row_sq + (size_t)it * T * n, T, n, K, st));
      } else {
        KM_TRY(kmeans_assign(d2, p.v, nullptr, p.temperature, p.lambd, p.u, p.labels, T, n, K, p.method, st));
        if (p.method == 1) KM_TRY(colsum_v(p.u, colsum, p.v, nullptr, T, n, K, st));   // v_update after u_update (em_gaussian.py:212-218)
      }
    }
    if (p.iter_events && p.iter_events[it + 1]) KM_TRY(cudaEventRecord((cudaEvent_t)p.iter_events[it + 1], st));
  }
  if (chained) {
    // u, labels [and v] of the last iteration, from the exponentials and tile statistics it left
    const int tiles = (K + km_tile() - 1) / km_tile();
    finish_chain_kernel<<<(T * n + 3) / 4, 128, 0, st>>>(d2, stats[(p.iters - 1) & 1], p.u, p.labels, T * n, K, tiles, km_tile());
    note_launch();
    if (p.method == 1) KM_TRY(colsum_v(p.u, colsum, p.v, nullptr, T, n, K, st));
  }
  if (p.method == 2) KM_TRY(kmeans_hard_criterions(row_sq, task_norm, p.criterions, p.iters, T, n, st));
  // the reference copies u_old after the update: the logged criterion of the soft variants is identically 0
  if (p.method != 2) KM_TRY(cudaMemsetAsync(p.criterions, 0, sizeof(float) * (size_t)std::max(p.iters, 1), st));
  if (coords && p.w) KM_TRY(kmeans_expand_centroids(p.coef, p.x, p.w, T, n, K, D, st));
  return cudaGetLastError();
}

}  // namespace tclip
