// dirichlet_mm.cu — the Dirichlet M-step: majorise-minimise fixed point on alpha (sm_100a).
//
// Replaces `curvature` + `update_alpha` of the reference (src/methods/zero_shot/em_dirichlet.py:153-177; identical
// copies in zero_shot/hard_em_dirichlet.py:153-177 and few_shot/{em,hard_em}_dirichlet.py:123-147), which is 99.6 %
// of the reference's run time.
//
// Layout: alpha, y are [rows, D] float32 row-major, rows = n_task * n_class, D = feature dim (<= 1024).
// One warp owns one row: element d lives in lane d % 32, register pair (d / 32) / 2, so a row of D <= 1024 floats stays
// in registers for a whole chunk of MM iterations (-y, read-only, in the warp's slice of shared memory); HBM sees 12 B
// per element per chunk (read alpha, read y, write alpha): the kernel is bound by the SM issue rate, not by memory
// (DESIGN.md §3.1).  The grid is persistent (<= kMMMinBlocks CTAs per SM), warps stride over the rows.
//
// The reference's early exit is *batch-global*: at l in {50, 100, ...} (l > 0) it stops iff
// ||a_new - a||^2 / ||a||^2 < 1e-11 over the whole [n_task, K, D] tensor.  The iteration loop is therefore cut into
// chunks that end exactly at those l: a chunk kernel emits per-CTA partial sums of the two norms for its last
// iteration, the last CTA to finish (atomic ticket) folds them in a fixed order (deterministic) and raises a
// device-side `done` flag that later chunk kernels test on entry.  No host synchronisation anywhere in the M-step.
//
// Two more forms of the same arithmetic:
//   * free-running rows (FR = true; the empty clusters of the skip-dead schedule): ignore the exit flag, store each
//     row's own criterion terms per check point, and stop iterating a row as soon as its fp32 trajectory is *proven*
//     periodic (bit-exact state comparison) — the remaining terms then follow by periodic extension;
//   * mm_spec_kernel: one row per CTA and the whole M-step in one launch for the few live rows of that schedule.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cstdlib>
#include <utility>

#include "tclip_kernels.cuh"
#include "tclip_math.cuh"

namespace tclip {

namespace {

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of the NP register pairs of one lane as a balanced tree of packed adds (fp32).
template <int NP>
__device__ __forceinline__ float lane_tree_sum(const float2 (&a)[NP], float2 tail_mask) {
  float2 t[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) t[j] = a[j];
  t[NP - 1] = f2mul(t[NP - 1], tail_mask);
#pragma unroll
  for (int w = 1; w < NP; w <<= 1) {
#pragma unroll
    for (int j = 0; j + w < NP; j += 2 * w) t[j] = f2add(t[j], t[j + w]);
  }
  return t[0].x + t[0].y;
}

// What a chunk kernel needs besides the row data.
struct ChunkArgs {
  const float* alpha_in;
  float* alpha_out;
  const float* y;
  const int* row_list;     // optional indirection
  const int* n_rows_dev;   // optional device-side row count
  int n_rows_host;
  int D;
  int n_iters;             // MM iterations of this chunk
  int has_check;           // the last iteration of this chunk is a check point
  int iters_cum;           // iterations executed once this chunk is done
  float tol;
  double2* partials;       // [gridDim.x]
  MMState* state;
  const double2* extra;    // optional: criterion terms of rows that are not iterated (cached dead rows)
  const int* split_gate;   // optional: {n_rows, cap}; n_rows <= cap => mm_spec_kernel runs the M-step instead
  // free-running rows only
  double2* row_cache;      // [n_checks][rows_total]
  int n_checks;
  int check_idx;
  long rows_total;
  int* frozen;
  float* snap;
  int snap_age;
  int snap_write;
  unsigned long long* work_ctr;
  int fr_chunks;           // free-running rows: all chunks of the M-step in ONE launch (the state stays in registers);
  int fr_chunk_iters;      // chunk 0 runs n_iters iterations, every later chunk fr_chunk_iters
};

constexpr int kSnapEvery = 3;  // snapshot fallback: detects chunk-periods 1..3 (iteration periods dividing 50, 100, 150)

// The last CTA of a chunk to finish folds the per-CTA partials (fixed order => reproducible), applies the reference's
// test `criterion < tol` (false for NaN, so a NaN never stops the loop) and keeps the executed-iteration count.
template <int THREADS>
__device__ __forceinline__ void finish_chunk(const ChunkArgs& g, double2 cta_terms, double2* red /* [THREADS] smem */) {
  __shared__ int is_last;
  if (threadIdx.x == 0) {
    if (g.has_check) g.partials[blockIdx.x] = cta_terms;
    __threadfence();
    const unsigned ticket = atomicAdd(&g.state->ticket, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double2 acc = make_double2(0.0, 0.0);
  if (g.has_check) {
    for (int i = threadIdx.x; i < (int)gridDim.x; i += THREADS) {
      const double2 p = g.partials[i];
      acc.x += p.x;
      acc.y += p.y;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = THREADS / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
      red[threadIdx.x].x += red[threadIdx.x + w].x;
      red[threadIdx.x].y += red[threadIdx.x + w].y;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    g.state->ticket = 0u;
    g.state->iters_done = g.iters_cum;
    if (g.has_check) {
      double num = red[0].x, den = red[0].y;
      if (g.extra) {
        num += g.extra->x;
        den += g.extra->y;
      }
      g.state->last_num = num;
      g.state->last_den = den;
      // the reference forms the ratio of two float32 squared norms; do the comparison on the float32 ratio too
      const float crit = (float)num / (float)den;
      if (crit < g.tol) g.state->done = 1;
    }
  }
}

// NP = ceil(D / 64) register pairs per lane: pair j holds elements d = (2j) * 32 + lane and (2j + 1) * 32 + lane, so
// every global access is a coalesced 128-byte row segment.  Only the last pair can hold padding (masked in the sums).
// All add / mul / fma work is issued as packed FFMA2 / FMUL2 / FADD2 (tclip_math.cuh: mm_update_pair).
template <int NP, bool FR>
__global__ void __launch_bounds__(kMMThreads, kMMMinBlocks)
mm_chunk_kernel(const ChunkArgs g) {
  if (!FR && g.state->done) return;  // an earlier chunk met the batch-global criterion: the M-step is over
  if (g.split_gate && g.split_gate[0] <= g.split_gate[1]) return;  // few rows: mm_spec_kernel runs this M-step
  // [warps][NP][32] pairs of -y, and for free-running rows a second slice: the state six iterations before the chunk end
  extern __shared__ float2 smem2[];
  __shared__ double2 red[kMMThreads];
  // Rows that are not free-running take psi(s) from the expansion around an anchor total (tclip_math.cuh:
  // row_psi_anchored): all lanes of a warp hold the same values, so the anchor lives once per warp in shared memory.  The
  // free-running rows keep the stateless evaluation: their periodicity proofs need the update to be a function of the
  // row state alone.
  __shared__ PsiAnchor anchors[kMMThreads / 32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int kWarps = kMMThreads / 32;
  const int n_rows = g.n_rows_dev ? *g.n_rows_dev : g.n_rows_host;
  const int D = g.D;
  float2* ny = smem2 + (size_t)warp * NP * 32 + lane;
  float2* a0s = smem2 + (size_t)(kWarps + warp) * NP * 32 + lane;  // FR only
  const int dx = (2 * NP - 2) * 32 + lane, dy = (2 * NP - 1) * 32 + lane;  // elements of the last pair
  const bool ok_x = dx < D, ok_y = dy < D;
  const float2 tail_mask = make_float2(ok_x ? 1.0f : 0.0f, ok_y ? 1.0f : 0.0f);
  double dsq_cta = 0.0, asq_cta = 0.0;  // per-lane partial sums over the rows of this warp (non-FR)

  for (int slot = blockIdx.x * kWarps + warp; slot < n_rows; slot += gridDim.x * kWarps) {
    const long row = g.row_list ? g.row_list[slot] : slot;
    if (FR && g.frozen[row]) continue;  // proven periodic earlier: its remaining check terms are already cached
    const float* ain = g.alpha_in + row * D;
    const float* yin = g.y + row * D;
    float* aout = g.alpha_out + row * D;

    float2 a[NP];
#pragma unroll
    for (int j = 0; j < NP - 1; ++j) {
      a[j] = make_float2(ain[(2 * j) * 32 + lane], ain[(2 * j + 1) * 32 + lane]);
      ny[j * 32] = make_float2(-__ldg(yin + (2 * j) * 32 + lane), -__ldg(yin + (2 * j + 1) * 32 + lane));
    }
    // padding lanes iterate on a harmless dummy (a = 1, y = -1)
    a[NP - 1] = make_float2(ok_x ? ain[dx] : 1.0f, ok_y ? ain[dy] : 1.0f);
    ny[(NP - 1) * 32] = make_float2(ok_x ? -__ldg(yin + dx) : 1.0f, ok_y ? -__ldg(yin + dy) : 1.0f);

    // Free-running rows walk through all their chunks here (they never look at the batch-global exit flag, so nothing
    // has to be decided between chunks); the other rows run the one chunk of this launch.
    const int n_chunks = FR ? g.fr_chunks : 1;
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int n_iters = (FR && chunk > 0) ? g.fr_chunk_iters : g.n_iters;
    const int check_idx = FR ? chunk : g.check_idx;
    const int snap_age = FR ? (chunk == 0 ? 0 : ((chunk - 1) % kSnapEvery) + 1) : 0;
    const bool snap_write = FR && (chunk % kSnapEvery == 0);
    bool froze = false;
    // Free-running rows: the last six iterations of the chunk form a window W[0..6]; W[6] == W[0] (bit for bit) proves
    // the trajectory periodic with a period dividing 6, and the terms of window updates 2, 4 and 6 are then the terms
    // of every later check point (their distance to this one is a multiple of 50 == 2 mod 6 iterations).
    // (the phase bookkeeping below assumes check points 2 mod 6 iterations apart: 50, the reference's spacing, is)
    const bool window = FR && n_iters >= 8 && (g.fr_chunk_iters % 6 == 2);
    const int n_plain = window ? n_iters - 6 : n_iters - 1;
    // Early proof: every six iterations the state is compared with the one six iterations earlier (kept in the warp's
    // shared-memory slice).  Equal => periodic from there on, so the window is run right away instead of at the chunk end;
    // its distance `delta` to the chunk end is kept even, so that the check points still fall on window updates 2, 4, 6.
    int delta = 0;
    bool at_fixed_point = false;
    const int probe0 = 6 + (n_iters & 1);
    double s = warp_sum_f64((double)lane_tree_sum<NP>(a, tail_mask));
    if (!FR) {
      __syncwarp();                     // the previous row of this warp is done with its anchor
      if (lane == 0) psi_anchor_reset(anchors[warp]);
      __syncwarp();
    }
    // one vote per row and iteration decides whether any element needs the small-a Taylor form (rare: fixed points sit
    // above 1/35 and only transients dip below 1/16); the running minimum is taken while the new values are produced
    float amin = fminf(a[0].x, a[0].y);
#pragma unroll
    for (int j = 1; j < NP; ++j) amin = fminf(amin, fminf(a[j].x, a[j].y));
    for (int it = 0; it < n_plain; ++it) {
      if (FR && window && it >= probe0 && (it - probe0) % 6 == 0) {
        bool same = it > probe0;
        if (same) {
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const float2 w0 = a0s[j * 32];
            // (padding lanes of the last pair iterate on a dummy that converges much later: they are not part of the row)
            if (j < NP - 1) same &= (w0.x == a[j].x) & (w0.y == a[j].y);
            else same &= ((w0.x == a[j].x) | !ok_x) & ((w0.y == a[j].y) | !ok_y);
          }
          same = __all_sync(0xffffffffu, same);
        }
        if (same) {
          delta = n_plain - it;
          break;
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) a0s[j * 32] = a[j];
      }
      RowPsi rp;
      if (FR) {
        rp = row_psi(s);
      } else {
        PsiAnchor an = anchors[warp];        // broadcast read: the 32 lanes hold the same row total, hence the same anchor
        const double anchored_at = an.s;
        rp = row_psi_anchored(s, an);
        if (an.s != anchored_at) {           // re-anchored (rare, warp-uniform): one lane publishes it, once every lane has read
          __syncwarp();
          if (lane == 0) anchors[warp] = an;
        }
        __syncwarp();
      }
      const bool any_small = __any_sync(0xffffffffu, amin < kSmallA);
      amin = 3.0e38f;
      bool fixed_point = FR && window;  // free-running rows: did this update leave every element unchanged?
      if (!any_small) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float2 an = mm_update_pair<0>(a[j], ny[j * 32], rp);
          if (FR) {
            if (j < NP - 1) fixed_point &= (an.x == a[j].x) & (an.y == a[j].y);
            else fixed_point &= ((an.x == a[j].x) | !ok_x) & ((an.y == a[j].y) | !ok_y);
          }
          a[j] = an;
          amin = fminf(amin, fminf(a[j].x, a[j].y));
        }
      } else {
        fixed_point = false;
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          a[j] = mm_update_pair<1>(a[j], ny[j * 32], rp);
          amin = fminf(amin, fminf(a[j].x, a[j].y));
        }
      }
      s = warp_sum_f64((double)lane_tree_sum<NP>(a, tail_mask));
      // An exact fixed point (measured: every empty cluster's row reaches one after 12-14 iterations,
      // profiles/r1_dead_rows_periodicity.txt): every later update is the identity, so the terms of every later check
      // are 0 and the row's squared norm; no window needed.
      if (FR && __all_sync(0xffffffffu, fixed_point)) {
        at_fixed_point = true;
        delta = n_iters - (it + 1);  // iterations not executed (work accounting; the phase is irrelevant here)
        break;
      }
    }
    if (FR && window && !at_fixed_point) {
#pragma unroll
      for (int j = 0; j < NP; ++j) a0s[j * 32] = a[j];
    }
    float2 d2 = make_float2(0.0f, 0.0f), a2 = make_float2(0.0f, 0.0f);  // terms of the last update
    if (FR && at_fixed_point) {
      // what every later check iteration would compute: a_new - a = 0 exactly, and the squared norm of this state,
      // accumulated in the same order as in the loop below
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float2 ao = (j == NP - 1) ? f2mul(a[j], tail_mask) : a[j];
        a2 = f2fma(ao, ao, a2);
      }
    }
    float2 d2_k2 = d2, a2_k2 = a2, d2_k4 = d2, a2_k4 = a2;               // FR: window updates 2 and 4
    const int n_tail = (FR && at_fixed_point) ? 0 : (window ? 6 : 1);
    for (int k = 1; k <= n_tail; ++k) {
      const RowPsi rp = row_psi(s);
      d2 = make_float2(0.0f, 0.0f);
      a2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float2 an = mm_update_pair(a[j], ny[j * 32], rp);
        float2 df = f2add(an, make_float2(-a[j].x, -a[j].y));
        float2 ao = a[j];
        if (j == NP - 1) {
          df = f2mul(df, tail_mask);
          ao = f2mul(ao, tail_mask);
        }
        d2 = f2fma(df, df, d2);
        a2 = f2fma(ao, ao, a2);
        a[j] = an;
      }
      if (FR && k == 2) {
        d2_k2 = d2;
        a2_k2 = a2;
      }
      if (FR && k == 4) {
        d2_k4 = d2;
        a2_k4 = a2;
      }
      if (k < n_tail) s = warp_sum_f64((double)lane_tree_sum<NP>(a, tail_mask));
    }
    if (!FR) {  // (the alpha of a free-running row is discarded: only its criterion terms are kept)
#pragma unroll
      for (int j = 0; j < NP - 1; ++j) {
        aout[(2 * j) * 32 + lane] = a[j].x;
        aout[(2 * j + 1) * 32 + lane] = a[j].y;
      }
      if (ok_x) aout[dx] = a[NP - 1].x;
      if (ok_y) aout[dy] = a[NP - 1].y;
    }

    if (!FR) {
      dsq_cta += (double)(d2.x + d2.y);
      asq_cta += (double)(a2.x + a2.y);
    } else {
      if (g.work_ctr && lane == 0) atomicAdd(g.work_ctr, (unsigned long long)(n_iters - delta));  // row-iterations executed
      // (a) period dividing 6 iterations, proven inside this chunk
      bool same6 = window;
      if (window && !at_fixed_point) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float2 w0 = a0s[j * 32];
          if (j < NP - 1) same6 &= (w0.x == a[j].x) & (w0.y == a[j].y);
          else same6 &= ((w0.x == a[j].x) | !ok_x) & ((w0.y == a[j].y) | !ok_y);
        }
        same6 = __all_sync(0xffffffffu, same6);
      }
      // (b) otherwise: period of 1..3 chunks, proven against the snapshot taken `snap_age` chunks ago.  The map "state at a
      // chunk end -> state at the next chunk end" is a fixed deterministic function of the row (y = const, 50 iterations).
      int period_chunks = 0;
      if (!same6) {
        float* sn = g.snap + row * D;
        bool same = snap_age > 0;
        if (snap_age > 0) {
#pragma unroll
          for (int j = 0; j < NP - 1; ++j)
            same &= (sn[(2 * j) * 32 + lane] == a[j].x) & (sn[(2 * j + 1) * 32 + lane] == a[j].y);
          if (ok_x) same &= sn[dx] == a[NP - 1].x;
          if (ok_y) same &= sn[dy] == a[NP - 1].y;
          same = __all_sync(0xffffffffu, same);
        }
        if (same) {
          period_chunks = snap_age;
        } else if (snap_write) {
#pragma unroll
          for (int j = 0; j < NP - 1; ++j) {
            sn[(2 * j) * 32 + lane] = a[j].x;
            sn[(2 * j + 1) * 32 + lane] = a[j].y;
          }
          if (ok_x) sn[dx] = a[NP - 1].x;
          if (ok_y) sn[dy] = a[NP - 1].y;
        }
      }
      const double t6x = warp_sum_f64((double)(d2.x + d2.y)), t6y = warp_sum_f64((double)(a2.x + a2.y));
      const double t2x = warp_sum_f64((double)(d2_k2.x + d2_k2.y)), t2y = warp_sum_f64((double)(a2_k2.x + a2_k2.y));
      const double t4x = warp_sum_f64((double)(d2_k4.x + d2_k4.y)), t4y = warp_sum_f64((double)(a2_k4.x + a2_k4.y));
      if (lane == 0) {
        double2* rc = g.row_cache + row;  // [n_checks][rows_total] so that the per-check sums read coalesced
        const long rs = g.rows_total;
        if (!same6) rc[check_idx * rs] = make_double2(t6x, t6y);
        if (same6) {
          // check j is (delta + 50 (j - check_idx)) iterations after the window's last update
          for (int j = check_idx; j < g.n_checks; ++j) {
            const int ph = (delta + 2 * (j - check_idx)) % 6;
            rc[j * rs] = ph == 0 ? make_double2(t6x, t6y) : (ph == 2 ? make_double2(t2x, t2y) : make_double2(t4x, t4y));
          }
          g.frozen[row] = 6;
        } else if (period_chunks > 0) {
          // state_end(c) == state_end(c - period)  =>  terms(j) == terms(j - period) for every later check j
          for (int j = check_idx + 1; j < g.n_checks; ++j) rc[j * rs] = rc[(j - period_chunks) * rs];
          g.frozen[row] = period_chunks;
        }
      }
      froze = same6 || period_chunks > 0;  // warp-uniform
    }
    if (FR && froze) break;
    }  // chunks
  }

  if (!FR) {
    __shared__ double2 wred[kWarps];
    const double ds = warp_sum_f64(dsq_cta), as = warp_sum_f64(asq_cta);
    if (lane == 0) wred[warp] = make_double2(ds, as);
    __syncthreads();
    double2 acc = make_double2(0.0, 0.0);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        acc.x += wred[w].x;
        acc.y += wred[w].y;
      }
    }
    finish_chunk<kMMThreads>(g, acc, red);
  }
}

// Few-rows form (the live clusters of the skip-dead schedule: a few hundred rows per batch): one row per CTA, its
// pairs dealt out to W warps so that the serial chain of one MM iteration is NPW pairs long instead of D/64.  Row total:
// warp partials through shared memory (ping-pong slots, one __syncthreads per iteration), summed as a tree by every thread.
//
// The whole M-step runs in ONE launch, speculatively: rows are independent except through the batch-global exit test, so
// every row simply iterates all iter_mm times and leaves, for each check point, its criterion terms and a snapshot of its
// state right after the check iteration (the state the reference would return if it stopped there).
// mm_spec_resolve_kernel then evaluates the checks in order with the terms of ALL rows (+ the cached dead rows) and, in
// the rare case that one fires, mm_spec_apply_kernel restores the snapshot of that check.  Same arithmetic, same result,
// no launch per 50 iterations.  Runs iff split_gate[0] <= split_gate[1].
constexpr int kSpecLeanMinBlocks = 6;   // CTAs per SM the register-lean few-rows kernel is compiled for (80 registers)
struct SpecArgs {
  const float* alpha_in;
  float* alpha_out;
  const float* y;
  const int* row_list;
  const int* n_rows_dev;
  const int* split_gate;
  int D;
  int iter_mm;
  int check_every;
  int n_checks;
  int cap;              // rows the scratch is sized for (= grid size)
  double2* terms;       // [n_checks][cap]
  float* snap;          // [n_checks][cap][D]
  const double2* extra; // optional [n_checks]
  float tol;
  MMState* state;       // iters_done out; `fired` check index in state->done_check
  unsigned long long* work_ctr;  // optional: += row-iterations executed
  int4* probe;          // optional [cap]: {iterations executed, iteration of the fixed point or -1, iteration at which a
                        // longer cycle was first detected or -1, its period} (measurement builds of the kernel only)
};

// PROBE = true is the statistics build (tclip_dirichlet_problem.spec_probe): same arithmetic and results, plus per row the
// iteration at which an update first returned every element bit for bit (an exact fixed point: the update is a
// deterministic function of the row state, so every later update would be the identity) and Brent's cycle search on the
// row state.  Measured on the bench workload (profiles/r2_spec_fixed_points.md): no live row ever reaches a fixed point;
// the singleton clusters (2 of 3 live rows) diverge and are never periodic, the large clusters dither in cycles of
// 60..480 iterations (elements oscillate by an ulp with periods 2..8, the row period is their least common multiple).
// So, unlike the empty clusters (mm_chunk_kernel<FR>), live rows offer no exact early stop worth its cost: carrying the
// per-iteration vote in the product kernel made it 14 % slower, and it is therefore compiled into this build only.
template <int W, int NPW, bool PIPE, bool PROBE>
// (two-phase form at D > 768: compiled for ONE CTA per SM — 184 registers, two CTAs resident — which measures 0.99 ms per tail
// iteration against 1.12 ms for the 168-register build that fits three; with a few hundred rows two per SM are enough)
__global__ void __launch_bounds__(32 * W, (W == 4 && NPW == 4) ? (PIPE ? 1 : kSpecLeanMinBlocks) : 0)
mm_spec_kernel(const SpecArgs g) {
  if (!(g.split_gate[0] <= g.split_gate[1])) return;
  if ((int)blockIdx.x >= *g.n_rows_dev) return;  // CTA-uniform
  __shared__ double part[2][W];
  __shared__ int flag[2][W];
  __shared__ double2 wred[W];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int D = g.D;
  const long row = g.row_list[blockIdx.x];
  const float* ain = g.alpha_in + row * D;
  const float* yin = g.y + row * D;
  float* aout = g.alpha_out + row * D;

  float2 a[NPW], ny[NPW], mask[NPW];
  int dxs[NPW], dys[NPW];
#pragma unroll
  for (int j = 0; j < NPW; ++j) {
    const int jp = warp * NPW + j;
    dxs[j] = (2 * jp) * 32 + lane;
    dys[j] = (2 * jp + 1) * 32 + lane;
    const bool okx = dxs[j] < D, oky = dys[j] < D;
    mask[j] = make_float2(okx ? 1.0f : 0.0f, oky ? 1.0f : 0.0f);
    a[j] = make_float2(okx ? ain[dxs[j]] : 1.0f, oky ? ain[dys[j]] : 1.0f);
    ny[j] = make_float2(okx ? -__ldg(yin + dxs[j]) : 1.0f, oky ? -__ldg(yin + dys[j]) : 1.0f);
  }
  // Software pipeline: the part of the next update that needs the pair only (mm_update_pre: both Stirling series, ln X,
  // the reciprocals) is issued together with the row-total shuffles, so it runs in the shadow of the reduction, the CTA
  // barrier and psi(s); what remains on the serial path after psi(s) is mm_update_post (~15 dependent operations).
  PairPre pre[NPW];
  int all_flags = 0;  // OR over the warps of the flags handed to the last row_total call
  auto row_total = [&](int parity, int my_flags) -> double {
    float2 t = f2mul(a[0], mask[0]);
#pragma unroll
    for (int j = 1; j < NPW; ++j) t = f2fma(a[j], mask[j], t);
    if (PIPE) {
#pragma unroll
      for (int j = 0; j < NPW; ++j) pre[j] = mm_update_pre(a[j]);
    }
    const double ws = warp_sum_f64((double)(t.x + t.y));
    if (lane == 0) part[parity][warp] = ws;
    if constexpr (PROBE) {
      const int wf = __reduce_or_sync(0xffffffffu, my_flags);
      if (lane == 0) flag[parity][warp] = wf;
    }
    __syncthreads();
    double pw[W];  // balanced tree: the total is on the serial path of every iteration
#pragma unroll
    for (int w = 0; w < W; ++w) pw[w] = part[parity][w];
    if constexpr (PROBE) {
      int f = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) f |= flag[parity][w];
      all_flags = f;
    }
#pragma unroll
    for (int st = 1; st < W; st <<= 1) {
#pragma unroll
      for (int w = 0; w + st < W; w += 2 * st) pw[w] += pw[w + st];
    }
    return pw[0];
  };
  double s = row_total(0, 0);
  int parity = 1;
  int next_check = g.check_every > 0 ? g.check_every : 0x7fffffff, c = 0;
  PsiAnchor anchor;  // psi(s) by expansion around an earlier row total: the float64 logarithm leaves the serial path
  psi_anchor_reset(anchor);
  int executed = g.iter_mm, fixed_at = -1;
  // measurement only: Brent's cycle search on the row state (reference state refreshed at powers of two)
  float2 ref[PROBE ? NPW : 1];
  int ref_iter = -1, cyc_at = -1, cyc_period = 0;
  for (int l = 0; l < g.iter_mm; ++l) {
    const RowPsi rp = row_psi_anchored(s, anchor);
    float2 an[NPW];
#pragma unroll
    for (int j = 0; j < NPW; ++j) an[j] = PIPE ? mm_update_post(pre[j], a[j], ny[j], rp) : mm_update_pair(a[j], ny[j], rp);
    if (l == next_check && c < g.n_checks) {
      // check iteration: its criterion terms and a snapshot of the new state
      float2 d2 = make_float2(0.0f, 0.0f), a2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int j = 0; j < NPW; ++j) {
        const float2 df = f2mul(f2add(an[j], make_float2(-a[j].x, -a[j].y)), mask[j]);
        const float2 ao = f2mul(a[j], mask[j]);
        d2 = f2fma(df, df, d2);
        a2 = f2fma(ao, ao, a2);
      }
      const double dsq = warp_sum_f64((double)(d2.x + d2.y)), asq = warp_sum_f64((double)(a2.x + a2.y));
      if (lane == 0) wred[warp] = make_double2(dsq, asq);
      float* sn = g.snap + ((long)c * g.cap + blockIdx.x) * D;
#pragma unroll
      for (int j = 0; j < NPW; ++j) {
        if (dxs[j] < D) sn[dxs[j]] = an[j].x;
        if (dys[j] < D) sn[dys[j]] = an[j].y;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double2 acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < W; ++w) {
          acc.x += wred[w].x;
          acc.y += wred[w].y;
        }
        g.terms[(long)c * g.cap + blockIdx.x] = acc;
      }
      ++c;
      next_check += g.check_every;
    }
    // bit 0: this update changed an element of the row (padding lanes iterate on a dummy and do not count);
    // bit 1: the new state differs from the reference state of the cycle search
    int my_flags = 0;
    if constexpr (PROBE) {
#pragma unroll
      for (int j = 0; j < NPW; ++j) {
        my_flags |= (int)((an[j].x != a[j].x) & (mask[j].x != 0.0f)) | (int)((an[j].y != a[j].y) & (mask[j].y != 0.0f));
        if (ref_iter >= 0)
          my_flags |= ((int)((an[j].x != ref[j].x) & (mask[j].x != 0.0f)) | (int)((an[j].y != ref[j].y) & (mask[j].y != 0.0f))) << 1;
      }
    }
#pragma unroll
    for (int j = 0; j < NPW; ++j) a[j] = an[j];
    if (l + 1 < g.iter_mm) {
      s = row_total(parity, my_flags);
      parity ^= 1;
      if constexpr (PROBE) {
        if (cyc_at < 0 && ref_iter >= 0 && !(all_flags & 2) && (all_flags & 1)) {
          cyc_at = l + 1;
          cyc_period = l + 1 - ref_iter;
        }
        if (fixed_at < 0 && !(all_flags & 1)) fixed_at = l + 1;
        if (cyc_at < 0 && l + 1 >= 8 && ((l + 1) & l) == 0) {  // state after l + 1 updates becomes the reference
#pragma unroll
          for (int j = 0; j < NPW; ++j) ref[j] = a[j];
          ref_iter = l + 1;
        }
      }
    }
  }
  if (threadIdx.x == 0) {
    if (g.work_ctr) atomicAdd(g.work_ctr, (unsigned long long)executed);
    if (PROBE && g.probe) g.probe[blockIdx.x] = make_int4(executed, fixed_at, cyc_at, cyc_period);
  }
#pragma unroll
  for (int j = 0; j < NPW; ++j) {
    if (dxs[j] < D) aout[dxs[j]] = a[j].x;
    if (dys[j] < D) aout[dys[j]] = a[j].y;
  }
}

// (Tried in round 2 and removed: deciding on the device which form runs and launching the up to 20 chunk kernels from there —
// CUDA dynamic parallelism, tail-launch stream — instead of enqueueing 20 launches that return at once when this kernel is
// selected (64 us of a tail iteration).  Correct, but device-side launches of these grids are far slower than the host's:
// outer iteration 0 went 9.3 -> 14.0 ms, a tail iteration 0.99 -> 1.10 ms, a batch with 1900 live rows 114 -> 178 ms.)

// One CTA: the checks in order, each over the terms of all speculated rows (fixed summation order) plus the cached dead
// rows; the first one below tol fires.  state->iters_done and state->fired are what the reference would have ended with.
__global__ void __launch_bounds__(1024)
mm_spec_resolve_kernel(const SpecArgs g) {
  if (!(g.split_gate[0] <= g.split_gate[1])) return;
  __shared__ double2 tot[32];
  __shared__ int fired;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  if (threadIdx.x == 0) fired = -1;
  __syncthreads();
  const int n_rows = *g.n_rows_dev;
  // one warp per check point sums that check's terms (lane-strided, then a shuffle tree: a fixed order); thread 0 then
  // walks the checks of the round in order
  for (int c0 = 0; c0 < g.n_checks; c0 += n_warps) {
    const int c = c0 + warp;
    if (c < g.n_checks) {
      double2 acc = make_double2(0.0, 0.0);
      for (int i = lane; i < n_rows; i += 32) {
        const double2 p = g.terms[(long)c * g.cap + i];
        acc.x += p.x;
        acc.y += p.y;
      }
      acc.x = warp_sum_f64(acc.x);
      acc.y = warp_sum_f64(acc.y);
      if (lane == 0) tot[warp] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 0; w < n_warps && c0 + w < g.n_checks; ++w) {
        double num = tot[w].x, den = tot[w].y;
        if (g.extra) {
          num += g.extra[c0 + w].x;
          den += g.extra[c0 + w].y;
        }
        g.state->last_num = num;
        g.state->last_den = den;
        const float crit = (float)num / (float)den;
        if (crit < g.tol) {
          fired = c0 + w;
          break;
        }
      }
    }
    __syncthreads();
    if (fired >= 0) break;
  }
  if (threadIdx.x == 0) {
    g.state->fired = fired;
    g.state->done = fired >= 0 ? 1 : 0;
    g.state->iters_done = fired >= 0 ? (fired + 1) * g.check_every + 1 : g.iter_mm;
  }
}

// Only if a check fired: every speculated row goes back to its state right after that check iteration.
__global__ void __launch_bounds__(256)
mm_spec_apply_kernel(const SpecArgs g) {
  if (!(g.split_gate[0] <= g.split_gate[1])) return;
  const int fired = g.state->fired;
  if (fired < 0 || (int)blockIdx.x >= *g.n_rows_dev) return;
  const long row = g.row_list[blockIdx.x];
  const float* sn = g.snap + ((long)fired * g.cap + blockIdx.x) * g.D;
  float* out = g.alpha_out + row * g.D;
  for (int d = threadIdx.x; d < g.D; d += 256) out[d] = sn[d];
}

__global__ void mm_reset_kernel(MMState* state) {
  state->done = 0;
  state->iters_done = 0;
  state->ticket = 0u;
  state->fired = -1;
  state->last_num = 0.0;
  state->last_den = 0.0;
}

// (Capping the residency of this kernel at 3 or 4 CTAs per SM — so that a few-rows kernel of another batch in flight,
// tclip_b200.pipeline, finds free registers next to it — was measured and does not pay: 2.42–2.48 k tasks/s against 2.53 k
// uncapped with 4 batches in flight, and the kernel alone drops from 0.60 to 0.57–0.59 of the FMA rate;
// gpurun_out/bs.json runs of scripts/gpu_streams.sh, commit message of this change.)
template <int NP, bool FR>
void launch_chunk(const ChunkArgs& g, int n_blocks, cudaStream_t st) {
  const size_t smem = (size_t)(kMMThreads / 32) * NP * 32 * sizeof(float2) * (FR ? 2 : 1);
  mm_chunk_kernel<NP, FR><<<n_blocks, kMMThreads, smem, st>>>(g);
}

template <int W, int NPW, bool PIPE = true>
void launch_spec(const SpecArgs& g, cudaStream_t st) {
  if (g.probe) mm_spec_kernel<W, NPW, PIPE, true><<<g.cap, 32 * W, 0, st>>>(g);
  else mm_spec_kernel<W, NPW, PIPE, false><<<g.cap, 32 * W, 0, st>>>(g);
}

using SpecFn = void (*)(const SpecArgs&, cudaStream_t);

using ChunkFn = void (*)(const ChunkArgs&, int, cudaStream_t);

template <bool FR, int... I>
constexpr auto make_table(std::integer_sequence<int, I...>) {
  return std::array<ChunkFn, sizeof...(I)>{&launch_chunk<I + 1, FR>...};
}

// NP = ceil(D / 64) pairs dealt to W = min(4, NP) warps, NPW = ceil(NP / W) pairs each.  The kernel is issue-bound on the
// SMs that hold two rows, and every warp repeats the row total and psi(s): 4 warps x 4 pairs measured 1.1 ms per
// 1000-iteration M-step of 224 rows, 8 x 2 1.3 ms, 16 x 1 1.8 ms, 2 x 8 1.4 ms (profiles/r1_spec_kernel.md; measured
// through an environment knob that commit 565e2b8 still has, removed since).  Round 2, with 4 batches in flight (where the
// SMs' issue rate, not the dependent chain, is the limit): 2 x 8 (12 % fewer instructions per row-iteration, 255 registers)
// 2929 tasks/s against 3108 for 4 x 4 — fewer, fatter warps lose there too.
// `lean`: the caller keeps several batches in flight.  The two-phase (software-pipelined) update holds 88 registers of
// look-ahead state per thread (168 in total): three CTAs per SM, i.e. the tails of only two batches fit on the GPU at once and
// those of the others queue behind them.  The plain update (bit-identical results, tests/test_math_host.py) needs 80
// registers, six CTAs per SM: 1.7 % slower alone, 9 % more tasks/s with 8 batches in flight (profiles/r2_in_flight.md).
SpecFn spec_fn(int np, bool lean) {
  if (lean && np > 12) return &launch_spec<4, 4, false>;
  switch (np) {
    case 1: return &launch_spec<1, 1>;
    case 2: return &launch_spec<1, 2>;   // D <= 128: one warp, no cross-warp traffic on the serial path
    case 3: return &launch_spec<3, 1>;
    case 4: return &launch_spec<4, 1>;
    case 5: case 6: case 7: case 8: return &launch_spec<4, 2>;
    case 9: case 10: case 11: case 12: return &launch_spec<4, 3>;
    default: return &launch_spec<4, 4>;
  }
}

// SM count of the CURRENT device (cached per device index; a failed query is not cached)
int sm_count() {
  static PerDeviceFlags cache;
  const int slot = current_device_slot();
  if (slot >= 0) {
    const int c = cache.v[slot].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
    return 148;  // B200
  if (slot >= 0) cache.v[slot].store(v, std::memory_order_relaxed);
  return v;
}

}  // namespace

int mm_max_dim() { return 32 * kMMMaxSlots; }

// One CTA per 4 rows, capped: warps stride over the rows.  `persistent` (the live-rows launch of the skip-dead schedule,
// whose real row count is only known on the device and is usually tiny) caps at one wave so that an almost empty launch
// costs nothing; otherwise 8 waves, which leaves the load balancing of a full batch to the hardware scheduler (a
// single static wave measured 10 % slower on 75 000 rows).
int mm_num_blocks(int n_rows, bool persistent) {
  const int need = (n_rows + (kMMThreads / 32) - 1) / (kMMThreads / 32);
  const int cap = sm_count() * kMMMinBlocks * (persistent ? 1 : 8);
  return need < cap ? need : cap;
}

// Enqueue one full M-step (<= iter_mm MM iterations with the batch-global early exit) on `st`.
cudaError_t mm_run(MMLaunch p, int iter_mm, int check_every, float tol, const double2* extra_checks, cudaStream_t st) {
  static constexpr auto table = make_table<false>(std::make_integer_sequence<int, kMMMaxSlots / 2>{});
  static constexpr auto table_fr = make_table<true>(std::make_integer_sequence<int, kMMMaxSlots / 2>{});
  if (p.D < 1 || p.D > mm_max_dim()) return cudaErrorInvalidValue;
  const int np = (p.D + 63) / 64;
  const bool free_run = p.row_cache != nullptr;
  const ChunkFn fn = free_run ? table_fr[np - 1] : table[np - 1];
  if (!free_run) {
    mm_reset_kernel<<<1, 1, 0, st>>>(p.state);
    note_launch();
  }
  int start = 0, check_idx = 0;
  while (start < iter_mm) {
    // the chunk ends at the next l with l > 0, l % check_every == 0, or at the last iteration
    int end = iter_mm - 1;
    int has_check = 0;
    if (check_every > 0) {
      const int from = start > 1 ? start : 1;
      const int cand = ((from + check_every - 1) / check_every) * check_every;  // first check point >= start
      if (cand <= iter_mm - 1) {
        end = cand;
        has_check = 1;
      }
    }
    if (free_run && (!has_check || check_idx > 0)) break;  // dead rows: one launch walks all chunks (no tail chunk:
                                                           // their alpha is discarded, only check terms matter)
    ChunkArgs g{};
    g.alpha_in = (start == 0) ? p.alpha_in : p.alpha_out;
    g.alpha_out = p.alpha_out;
    g.y = p.y;
    g.row_list = p.row_list;
    g.n_rows_dev = p.n_rows_dev;
    g.n_rows_host = p.n_rows;
    g.D = p.D;
    g.n_iters = end - start + 1;
    g.has_check = has_check;
    g.iters_cum = end + 1;
    g.tol = tol;
    g.partials = p.partials;
    g.state = p.state;
    g.extra = (has_check && extra_checks) ? extra_checks + check_idx : nullptr;
    g.split_gate = p.split_gate;
    if (free_run) {
      g.row_cache = p.row_cache;
      g.n_checks = p.n_checks;
      g.check_idx = check_idx;
      g.rows_total = p.rows_total;
      g.frozen = p.frozen;
      g.snap = p.snap;
      // periodicity snapshots every kSnapEvery chunks, compared at every chunk end in between (derived in the kernel)
      g.work_ctr = p.work_ctr;
      g.fr_chunks = p.n_checks;
      g.fr_chunk_iters = check_every;
    }
    fn(g, p.n_blocks, st);
    note_launch();
    if (has_check) ++check_idx;
    start = end + 1;
  }
  if (p.split_gate) {
    // the few-rows alternative of the same M-step (the chunk kernels above returned at once if it is selected)
    SpecArgs g{};
    g.alpha_in = p.alpha_in;
    g.alpha_out = p.alpha_out;
    g.y = p.y;
    g.row_list = p.row_list;
    g.n_rows_dev = p.n_rows_dev;
    g.split_gate = p.split_gate;
    g.D = p.D;
    g.iter_mm = iter_mm;
    g.check_every = check_every;
    g.n_checks = check_every > 0 ? (iter_mm - 1) / check_every : 0;
    g.cap = p.split_cap;
    g.terms = p.spec_terms;
    g.snap = p.spec_snap;
    g.work_ctr = p.work_ctr;
    g.probe = p.spec_probe;
    g.extra = extra_checks;
    g.tol = tol;
    g.state = p.state;
    spec_fn(np, p.spec_lean)(g, st);
    mm_spec_resolve_kernel<<<1, 32 * std::max(1, std::min(g.n_checks, 32)), 0, st>>>(g);
    mm_spec_apply_kernel<<<g.cap, 256, 0, st>>>(g);
    note_launch(3);
  }
  return cudaGetLastError();
}

}  // namespace tclip
