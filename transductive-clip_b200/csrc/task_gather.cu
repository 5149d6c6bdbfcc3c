// task_gather.cu — device-side construction of a task batch from the cached feature matrix.
//
// Replaces, for one `run_task` batch, the indexing `all_features_query[indices, :]`, `all_labels_query[indices]` per
// task of the evaluator (src/eval_zero_shot.py:158-163) and the per-key `torch.cat(...).view(n_task, n_samples, -1)` of
// `Tasks_Generator_zero_shot.generate_tasks` (src/task_generator_zero_shot.py:36-65): the cached features [N, F] and
// labels [N] stay resident in HBM, the sampler's index lists ([T, n] int64, drawn on the host with the reference's own
// random calls) are the only thing that crosses PCIe (T*n*8 bytes instead of T*n*F*4), and one kernel writes
// x_q [T, n, F] and y_q [T, n, 1].  One warp per sample row, 16-byte accesses when the row pitch allows.
#include <cuda_runtime.h>

#include "tclip_kernels.cuh"

namespace tclip {

namespace {

__global__ void __launch_bounds__(256)
gather_tasks_kernel(const float* __restrict__ features, const long long* __restrict__ labels,
                    const long long* __restrict__ idx, float* __restrict__ x_q, long long* __restrict__ y_q,
                    long long n_rows, long long count, int F, int vec4, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= count) return;
  long long r = idx[m];
  const bool ok = r >= 0 && r < n_rows;
  if (!ok) {  // an index outside the feature matrix: count it (the caller raises), write zeros
    if (lane == 0 && bad) atomicAdd(bad, 1);
    r = 0;
  }
  const float* src = features + r * F;
  float* dst = x_q + m * F;
  if (vec4) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < F / 4; i += 32) d4[i] = ok ? s4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int i = lane; i < F; i += 32) dst[i] = ok ? src[i] : 0.0f;
  }
  if (lane == 0 && y_q) y_q[m] = ok ? labels[r] : -1;
}

}  // namespace

cudaError_t gather_tasks(const float* features, const long long* labels, const long long* idx, float* x_q,
                         long long* y_q, long long n_rows, long long count, int F, int* bad, cudaStream_t st) {
  const bool aligned = ((reinterpret_cast<unsigned long long>(features) | reinterpret_cast<unsigned long long>(x_q)) & 15ull) == 0;
  const int vec4 = (F % 4 == 0 && aligned) ? 1 : 0;
  const long long blocks = (count + 7) / 8;
  gather_tasks_kernel<<<(unsigned)blocks, 256, 0, st>>>(features, labels, idx, x_q, y_q, n_rows, count, F, vec4, bad);
  note_launch();
  return cudaGetLastError();
}

}  // namespace tclip

// Few-shot form: Tasks_Generator_few_shot.get_task (src/task_generator_few_shot.py:27-58) also relabels every task —
// `unique_labels = flip(unique(labels_support, sorted=False))`, label y -> its position in that list, and with softmax
// features the columns are re-ordered the same way (`data[:, unique_labels]`).  The per-task column list (col_perm [T, U])
// and the label map (label_map [T, n_labels], 0 where the reference leaves its zeros) are built on the host with the
// reference's own torch calls; this kernel applies them while gathering: x_out[t, m, j] = features[idx[t, m], col_perm[t, j]],
// y_out[t, m] = label_map[t, labels[idx[t, m]]].
namespace tclip {

namespace {

__global__ void __launch_bounds__(256)
gather_tasks_remap_kernel(const float* __restrict__ features, const long long* __restrict__ labels,
                          const long long* __restrict__ idx, const long long* __restrict__ col_perm,
                          const long long* __restrict__ label_map, float* __restrict__ x_out,
                          long long* __restrict__ y_out, long long n_rows, long long count, int per_task, int F, int U,
                          int n_labels, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= count) return;
  const long long t = m / per_task;
  long long r = idx[m];
  bool ok = r >= 0 && r < n_rows;
  if (!ok) r = 0;
  const float* src = features + r * F;
  const long long* perm = col_perm + t * U;
  float* dst = x_out + m * U;
  for (int j = lane; j < U; j += 32) {
    const long long c = perm[j];
    const bool okc = c >= 0 && c < F;
    ok &= okc;
    dst[j] = (ok && okc) ? src[c] : 0.0f;
  }
  long long lab = ok ? labels[r] : -1;
  const bool okl = lab >= 0 && lab < n_labels;
  if (lane == 0 && y_out) y_out[m] = okl ? label_map[t * n_labels + lab] : -1;
  ok = __all_sync(0xffffffffu, ok) && okl;
  if (!ok && lane == 0 && bad) atomicAdd(bad, 1);
}

}  // namespace

cudaError_t gather_tasks_remap(const float* features, const long long* labels, const long long* idx,
                               const long long* col_perm, const long long* label_map, float* x_out, long long* y_out,
                               long long n_rows, long long count, int per_task, int F, int U, int n_labels, int* bad,
                               cudaStream_t st) {
  const long long blocks = (count + 7) / 8;
  gather_tasks_remap_kernel<<<(unsigned)blocks, 256, 0, st>>>(features, labels, idx, col_perm, label_map, x_out, y_out,
                                                             n_rows, count, per_task, F, U, n_labels, bad);
  note_launch();
  return cudaGetLastError();
}

}  // namespace tclip
