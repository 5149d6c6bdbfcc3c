// dirichlet_mm.cu — the Dirichlet M-step: majorise-minimise fixed point on alpha (sm_100a).
//
// Replaces `curvature` + `update_alpha` of the reference (src/methods/zero_shot/em_dirichlet.py:153-177; identical
// copies in zero_shot/hard_em_dirichlet.py:153-177 and few_shot/{em,hard_em}_dirichlet.py:123-147), which is 99.6 %
// of the reference's run time.
//
// Layout: alpha, y are [rows, D] float32 row-major, rows = n_task * n_class, D = feature dim (<= 1024).
// One warp owns one row: element d lives in lane d % 32, register slot d / 32, so a row of D <= 1024 floats stays
// in registers for a whole chunk of MM iterations; HBM sees 12 B per element per chunk (read alpha, read y, write
// alpha), i.e. the kernel is FP32/MUFU-issue bound, not memory bound (DESIGN.md "MM kernel").
//
// The reference's early exit is *batch-global*: at l in {50, 100, ...} (l > 0) it stops iff
// ||a_new - a||^2 / ||a||^2 < 1e-11 over the whole [n_task, K, D] tensor.  The iteration loop is therefore cut into
// chunks that end exactly at those l: a chunk kernel emits per-block partial sums of the two norms for its last
// iteration, a 1-block decide kernel folds them in a fixed order (deterministic) and raises a device-side `done`
// flag that later chunk kernels test on entry.  No host synchronisation anywhere in the M-step.
#include <cuda_runtime.h>

#include <array>
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

// NP = ceil(D / 64) register pairs per lane: pair j holds elements d = (2j) * 32 + lane and (2j + 1) * 32 + lane, so
// every global access is a coalesced 128-byte row segment.  Only the last pair can hold padding (masked in the sums).
// All add / mul / fma work is issued as packed FFMA2 / FMUL2 / FADD2 (tclip_math.cuh: mm_update_pair).
template <int NP>
__global__ void __launch_bounds__(kMMThreads, kMMMinBlocks)
mm_chunk_kernel(const float* alpha_in, float* alpha_out, const float* __restrict__ y,
                const int* __restrict__ row_list, const int* __restrict__ n_rows_dev, int n_rows_host, int D,
                int n_iters, int emit_check, double2* __restrict__ partials, const MMState* __restrict__ state,
                double2* __restrict__ row_cache, int n_checks, int check_idx, int* __restrict__ frozen,
                float* __restrict__ snap, int snap_age, int snap_write, unsigned long long* __restrict__ work_ctr,
                long cache_stride) {
  if (state->done) return;  // an earlier chunk met the batch-global criterion: the M-step is over
  extern __shared__ float2 ny_smem[];  // [warps per CTA][NP][32] pairs of -y
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_rows = n_rows_dev ? *n_rows_dev : n_rows_host;
  const int slot = blockIdx.x * (kMMThreads / 32) + warp;
  bool active = slot < n_rows;
  double dsq = 0.0, asq = 0.0;
  long row = 0;
  int period = 0;  // free-running rows: > 0 once this chunk proved the row's trajectory periodic (in chunks)

  if (active) {
    row = row_list ? row_list[slot] : slot;
    if (frozen && frozen[row]) active = false;  // proven periodic earlier: its remaining check terms are already cached
  }
  if (active) {
    const float* ain = alpha_in + row * D;
    const float* yin = y + row * D;
    float* aout = alpha_out + row * D;
    const int dx = (2 * NP - 2) * 32 + lane, dy = (2 * NP - 1) * 32 + lane;  // elements of the last pair
    const bool ok_x = dx < D, ok_y = dy < D;
    const float2 tail_mask = make_float2(ok_x ? 1.0f : 0.0f, ok_y ? 1.0f : 0.0f);

    // alpha stays in registers; -y (read-only, one LDS.64 per pair and iteration) lives in this warp's slice of shared
    // memory, which keeps the kernel at <= 80 registers, i.e. 6 instead of 4 warps per scheduler to overlap the FMA and
    // MUFU pipes (ncu: math-pipe-throttle was the top stall at 4 warps, profiles/r1_mm_chunk_packed.md)
    float2* ny = ny_smem + (size_t)warp * NP * 32 + lane;
    float2 a[NP];
#pragma unroll
    for (int j = 0; j < NP - 1; ++j) {
      a[j] = make_float2(ain[(2 * j) * 32 + lane], ain[(2 * j + 1) * 32 + lane]);
      ny[j * 32] = make_float2(-__ldg(yin + (2 * j) * 32 + lane), -__ldg(yin + (2 * j + 1) * 32 + lane));
    }
    // padding lanes iterate on a harmless dummy (a = 1, y = -1)
    a[NP - 1] = make_float2(ok_x ? ain[dx] : 1.0f, ok_y ? ain[dy] : 1.0f);
    ny[(NP - 1) * 32] = make_float2(ok_x ? -__ldg(yin + dx) : 1.0f, ok_y ? -__ldg(yin + dy) : 1.0f);

    double s = warp_sum_f64((double)lane_tree_sum<NP>(a, tail_mask));
    for (int it = 0; it < n_iters - 1; ++it) {
      const RowPsi rp = row_psi(s);
#pragma unroll
      for (int j = 0; j < NP; ++j) a[j] = mm_update_pair(a[j], ny[j * 32], rp);
      s = warp_sum_f64((double)lane_tree_sum<NP>(a, tail_mask));
    }
    {  // last iteration of the chunk: also the one the criterion is evaluated on
      const RowPsi rp = row_psi(s);
      float2 d2 = make_float2(0.0f, 0.0f), a2 = make_float2(0.0f, 0.0f);
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
      dsq = (double)(d2.x + d2.y);
      asq = (double)(a2.x + a2.y);
    }
#pragma unroll
    for (int j = 0; j < NP - 1; ++j) {
      aout[(2 * j) * 32 + lane] = a[j].x;
      aout[(2 * j + 1) * 32 + lane] = a[j].y;
    }
    if (ok_x) aout[dx] = a[NP - 1].x;
    if (ok_y) aout[dy] = a[NP - 1].y;

    if (work_ctr && lane == 0) atomicAdd(work_ctr, (unsigned long long)n_iters);  // row-iterations actually executed
    if (frozen) {
      // Free-running (dead-cluster) rows: the map "state at a chunk end -> state at the next chunk end" is a fixed
      // deterministic function of the row (y = const, 50 iterations), so a chunk-end state that equals, bit for bit, the
      // snapshot taken `snap_age` chunks ago proves the trajectory periodic with that period: every later check term
      // repeats and the row need not be iterated again.  (fp32 MM trajectories settle on such orbits after ~50-200
      // iterations; the reference would keep recomputing the same numbers.)
      float* sn = snap + row * D;
      bool same = snap_age > 0;
      if (snap_age > 0) {
#pragma unroll
        for (int j = 0; j < NP - 1; ++j) {
          same &= (sn[(2 * j) * 32 + lane] == a[j].x) & (sn[(2 * j + 1) * 32 + lane] == a[j].y);
        }
        if (ok_x) same &= sn[dx] == a[NP - 1].x;
        if (ok_y) same &= sn[dy] == a[NP - 1].y;
        same = __all_sync(0xffffffffu, same);
      }
      if (same) {
        period = snap_age;
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
  }

  if (emit_check == 2) {  // free-running rows: each row keeps its own terms
    dsq = warp_sum_f64(dsq);
    asq = warp_sum_f64(asq);
    if (active && lane == 0) {
      double2* rc = row_cache + row;  // [n_checks][rows_total] so that the per-check sums read coalesced
      const long rs = cache_stride;
      rc[check_idx * rs] = make_double2(dsq, asq);
      if (period > 0) {
        // state_end(c) == state_end(c - period)  =>  terms(j) == terms(j - period) for every later check j
        for (int j = check_idx + 1; j < n_checks; ++j) rc[j * rs] = rc[(j - period) * rs];
        frozen[row] = period;
      }
    }
  } else if (emit_check) {
    __shared__ double2 red[kMMThreads / 32];
    dsq = warp_sum_f64(dsq);
    asq = warp_sum_f64(asq);
    if (lane == 0) red[warp] = make_double2(dsq, asq);
    __syncthreads();
    if (threadIdx.x == 0) {
      double2 acc = red[0];
#pragma unroll
      for (int w = 1; w < kMMThreads / 32; ++w) {
        acc.x += red[w].x;
        acc.y += red[w].y;
      }
      partials[blockIdx.x] = acc;
    }
  }
}

// Folds the per-block partials of the chunk that just ran (fixed order => bit-reproducible), applies the reference's
// test `criterion < tol` (false for NaN, so a NaN never stops the loop) and keeps the executed-iteration count.
__global__ void __launch_bounds__(256)
mm_decide_kernel(const double2* __restrict__ partials, int n_partials, const double2* __restrict__ extra,
                 int has_check, int iters_cum, float tol, MMState* state) {
  if (state->done) return;
  __shared__ double2 red[256];
  double2 acc = make_double2(0.0, 0.0);
  if (has_check) {
    for (int i = threadIdx.x; i < n_partials; i += 256) {
      const double2 p = partials[i];
      acc.x += p.x;
      acc.y += p.y;
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
  if (threadIdx.x == 0) {
    state->iters_done = iters_cum;
    if (has_check) {
      double num = red[0].x, den = red[0].y;
      if (extra) {  // norm contributions of rows that are not iterated in this launch (cached dead rows)
        num += extra->x;
        den += extra->y;
      }
      state->last_num = num;
      state->last_den = den;
      // the reference forms the ratio of two float32 squared norms; do the comparison on the float32 ratio too
      const float crit = (float)num / (float)den;
      if (crit < tol) state->done = 1;
    }
  }
}

__global__ void mm_reset_kernel(MMState* state) {
  state->done = 0;
  state->iters_done = 0;
  state->last_num = 0.0;
  state->last_den = 0.0;
}

template <int NP>
void launch_chunk(const MMLaunch& p, int n_iters, int emit_check, int check_idx, int snap_age, int snap_write,
                  cudaStream_t st) {
  mm_chunk_kernel<NP><<<p.n_blocks, kMMThreads, (size_t)(kMMThreads / 32) * NP * 32 * sizeof(float2), st>>>(p.alpha_in, p.alpha_out, p.y, p.row_list, p.n_rows_dev,
                                                         p.n_rows, p.D, n_iters, emit_check, p.partials, p.state,
                                                         p.row_cache, p.n_checks, check_idx, p.frozen, p.snap,
                                                         snap_age, snap_write, p.work_ctr, (long)p.rows_total);
}

constexpr int kSnapEvery = 3;  // detects chunk-periods 1..3, i.e. iteration periods dividing 50, 100 or 150

using ChunkFn = void (*)(const MMLaunch&, int, int, int, int, int, cudaStream_t);

template <int... I>
constexpr auto make_table(std::integer_sequence<int, I...>) {
  return std::array<ChunkFn, sizeof...(I)>{&launch_chunk<I + 1>...};
}

}  // namespace

int mm_max_dim() { return 32 * kMMMaxSlots; }

int mm_num_blocks(int n_rows) { return (n_rows + (kMMThreads / 32) - 1) / (kMMThreads / 32); }

// Enqueue one full M-step (<= iter_mm MM iterations with the batch-global early exit) on `st`.
cudaError_t mm_run(MMLaunch p, int iter_mm, int check_every, float tol, const double2* extra_checks, cudaStream_t st) {
  static constexpr auto table = make_table(std::make_integer_sequence<int, kMMMaxSlots / 2>{});
  if (p.D < 1 || p.D > mm_max_dim()) return cudaErrorInvalidValue;
  const int np = (p.D + 63) / 64;
  const ChunkFn fn = table[np - 1];
  mm_reset_kernel<<<1, 1, 0, st>>>(p.state);
  note_launch();
  const float* first_in = p.alpha_in;
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
    MMLaunch q = p;
    q.alpha_in = (start == 0) ? first_in : p.alpha_out;
    const bool free_run = p.row_cache != nullptr;
    if (free_run && !has_check) break;  // dead rows: alpha is discarded, only check terms matter -> no tail chunk
    // periodicity snapshots of free-running rows every kSnapEvery chunks, compared at every chunk end in between
    const int snap_age = (free_run && p.frozen) ? (check_idx == 0 ? 0 : ((check_idx - 1) % kSnapEvery) + 1) : 0;
    const int snap_write = (free_run && p.frozen && check_idx % kSnapEvery == 0) ? 1 : 0;
    fn(q, end - start + 1, has_check ? (free_run ? 2 : 1) : 0, check_idx, snap_age, snap_write, st);
    const double2* extra = (has_check && extra_checks) ? extra_checks + check_idx : nullptr;
    mm_decide_kernel<<<1, 256, 0, st>>>(p.partials, p.n_blocks, extra, free_run ? 0 : has_check, end + 1, tol,
                                        p.state);
    note_launch(2);
    if (has_check) ++check_idx;
    start = end + 1;
  }
  return cudaGetLastError();
}

}  // namespace tclip
