// contraction_tc.cu — the E-step contraction  l3[t,n,k] = sum_d logz[t,n,d] * (alpha[t,k,d] - 1)  on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA), in split precision (3 x TF32) with
// the running sum kept OUTSIDE the tensor core in round-to-nearest fp32.
//
// Replaces the product `torch.log(query + eps) * (alpha - 1)` summed over d in get_logits of the reference
// (src/methods/zero_shot/em_dirichlet.py:37-38; few_shot/em_dirichlet.py:36-38), a [n x D] . [D x K] GEMM per task.
//
// Why this shape of kernel:
//   * both operands are K-major as they lie in HBM (logz [T,n,D], alpha [T,K,D]; the contraction runs over the innermost
//     index d), so TMA boxes of 128 rows x 32 floats land directly in the canonical 128-byte-swizzled UMMA layout;
//   * one CTA owns a 128 (queries, 75 used) x 128 (classes) output tile of one task and walks D in blocks of 32;
//   * split precision: x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi)  (round to nearest, unbiased), and
//     a b ~= a_lo b_hi + a_hi b_lo + a_hi b_hi  (the dropped a_lo b_lo is 2^-22 relative); `alpha - 1` is formed in fp32
//     first, exactly as the reference does.  Four warps do this element-wise on the tiles where TMA left them (the swizzle
//     is irrelevant to an element-wise map) and hand them to the MMA thread through the async proxy fence;
//   * the tensor core truncates when it adds into its fp32 accumulator (measured: profiles/r1_contraction_tc.md), which
//     over 375 dependent accumulations of same-signed terms (log z < 0, alpha > 0) is a bias of ~1e-5 relative — the
//     logits are ~1e7 and neighbouring classes differ by O(1).  So every D-block is accumulated from zero into its own
//     TMEM buffer (12 MMAs: the two small cross terms first) and four more warps drain that buffer with tcgen05.ld and
//     add it to fp32 registers with round-to-nearest adds — the Ootomo-Yokota scheme on TMEM.
//
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-5 hi/lo split, warps 6-9
// accumulate + store.  Pipelines: full[s] (TMA -> split), ready[s] (split -> MMA), empty[s] (MMA done -> TMA),
// acc_full[b] (MMA done -> drain), acc_empty[b] (drain -> MMA).  All waits are bounded and trap instead of hanging.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "tclip_kernels.cuh"

namespace tclip {

namespace {

constexpr int kTileM = 128;          // queries per tile (rows >= n are zero-filled by TMA)
constexpr int kTileN = 128;          // classes per tile
constexpr int kBlockK = 32;          // floats per D-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;            // tf32: 32 bytes per MMA
constexpr int kStages = 3;
constexpr int kAccBufs = 2;          // TMEM accumulator buffers of kTileN columns
constexpr int kTileBytes = kTileM * kBlockK * 4;          // 16 KB
constexpr int kStageBytes = 4 * kTileBytes;               // A(hi), A_lo, B(hi), B_lo
constexpr int kThreads = 320;
constexpr int kSplitThreads = 128;
constexpr int kDrainThreads = 128;
constexpr uint32_t kTmemCols = kAccBufs * kTileN;         // 256: power of two >= 32
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded: a protocol error traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive columns: thread i of the warp receives row (lane base + i), columns [col, col + 32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// Round to the nearest TF32 (10 explicit mantissa bits), ties away from zero — what cvt.rna.tf32.f32 returns for finite
// inputs, in two integer instructions (the PTX cvt expands to a ~6-instruction NaN/Inf-safe sequence, which made the
// split warps the bottleneck of the kernel; log z and alpha - 1 are finite and far from the overflow threshold).
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// K-major operand tile, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor:
// start >> 4 at [0,14), LBO at [16,30) (unused for swizzled K-major, 1), SBO >> 4 at [32,46), version 1 at [46,48),
// layout SWIZZLE_128B = 2 at [61,64)).  Stepping along K inside the swizzle row = advancing the start address.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: C = F32 (1 at [4,6)), A = B = TF32 (2 at [7,10), [10,13)), both K-major, N >> 3 at [17,23),
// M >> 4 at [24,29)
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

// kExternalAcc = true is the product; false keeps the whole sum in one TMEM accumulator (the textbook GEMM loop) and
// exists only so that the truncation of the tensor core's accumulate can be measured against it (tests, profiles/).
template <bool kExternalAcc>
__global__ void __launch_bounds__(kThreads, 1)
logits_tc_kernel(const __grid_constant__ CUtensorMap map_logz, const __grid_constant__ CUtensorMap map_alpha,
                 float* __restrict__ l3, int n, int K, int D, const int* __restrict__ gate, float b_shift, int b_shared,
                 const TcEpilogue ep) {
  if (gate != nullptr && !(gate[0] > gate[1])) return;  // the row-wise kernels take this E-step (skip-dead schedule)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bars = base + kStages * kStageBytes;
  const uint32_t full0 = bars, ready0 = bars + 8 * kStages, empty0 = bars + 16 * kStages;
  const uint32_t accfull0 = bars + 24 * kStages, accempty0 = accfull0 + 8 * kAccBufs;
  const uint32_t tmem_slot = accempty0 + 8 * kAccBufs;
  uint8_t* const gen_base = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to the same place

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid: (class tiles, row tiles of the A operand, tasks); b_shared: one B matrix for all tasks (task coordinate 0)
  const int t = blockIdx.z, m0 = blockIdx.y * kTileM, k0 = blockIdx.x * kTileN;
  const int tb = b_shared ? 0 : t;
  const int n_kb = (D + kBlockK - 1) / kBlockK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(ready0 + 8 * s, kSplitThreads);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int b = 0; b < kAccBufs; ++b) {
      mbar_init(accfull0 + 8 * b, 1);
      mbar_init(accempty0 + 8 * b, kDrainThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t use = kb / kStages;
        mbar_wait(empty0 + 8 * s, (use & 1) ^ 1);
        const uint32_t st = base + s * kStageBytes;
        mbar_arrive_expect_tx(full0 + 8 * s, 2 * kTileBytes);
        tma_load_3d(st, &map_logz, kb * kBlockK, m0, t, full0 + 8 * s);
        tma_load_3d(st + 2 * kTileBytes, &map_alpha, kb * kBlockK, k0, tb, full0 + 8 * s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      for (int kb = 0; kb < n_kb; ++kb) {
        const int s = kb % kStages, b = kExternalAcc ? kb % kAccBufs : 0;
        const uint32_t use = kb / kStages, acc_use = kb / kAccBufs;
        if (kExternalAcc) mbar_wait(accempty0 + 8 * b, (acc_use & 1) ^ 1);
        mbar_wait(ready0 + 8 * s, use & 1);
        tc_fence_after();
        const uint32_t st = base + s * kStageBytes;
        const uint32_t a_hi = st, a_lo = st + kTileBytes, b_hi = st + 2 * kTileBytes, b_lo = st + 3 * kTileBytes;
        const uint32_t d = tmem_base + b * kTileN;
        // the two small cross terms first, from zero; then the leading term
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_tf32(d, umma_desc(a_lo + k * kUmmaK * 4), umma_desc(b_hi + k * kUmmaK * 4), kInstrDesc,
                    (k > 0 || (!kExternalAcc && kb > 0)) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_tf32(d, umma_desc(a_hi + k * kUmmaK * 4), umma_desc(b_lo + k * kUmmaK * 4), kInstrDesc, 1);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)
          umma_tf32(d, umma_desc(a_hi + k * kUmmaK * 4), umma_desc(b_hi + k * kUmmaK * 4), kInstrDesc, 1);
        tc_commit(empty0 + 8 * s);      // the stage may be refilled once these MMAs have read it
        if (kExternalAcc || kb == n_kb - 1) tc_commit(accfull0 + 8 * b);  // the partial product is complete
      }
    }
  } else if (warp < 6) {
    // ===== split warps: raw tile -> hi (in place) + lo, element-wise =====
    const int tid = threadIdx.x - 64;
    for (int kb = 0; kb < n_kb; ++kb) {
      const int s = kb % kStages;
      const uint32_t use = kb / kStages;
      mbar_wait(full0 + 8 * s, use & 1);
      float4* a_hi = reinterpret_cast<float4*>(gen_base + s * kStageBytes);
      float4* a_lo = a_hi + kTileBytes / 16;
      float4* b_hi = a_hi + 2 * (kTileBytes / 16);
      float4* b_lo = a_hi + 3 * (kTileBytes / 16);
#pragma unroll 4
      for (int i = tid; i < kTileBytes / 16; i += kSplitThreads) {
        const float4 x = a_hi[i];
        float4 h, l;
        h.x = rn_tf32(x.x); h.y = rn_tf32(x.y); h.z = rn_tf32(x.z); h.w = rn_tf32(x.w);
        l.x = rn_tf32(x.x - h.x); l.y = rn_tf32(x.y - h.y); l.z = rn_tf32(x.z - h.z); l.w = rn_tf32(x.w - h.w);
        a_hi[i] = h;
        a_lo[i] = l;
      }
#pragma unroll 4
      for (int i = tid; i < kTileBytes / 16; i += kSplitThreads) {
        float4 x = b_hi[i];
        x.x -= b_shift; x.y -= b_shift; x.z -= b_shift; x.w -= b_shift;   // alpha - 1 in fp32, as the reference forms it (shift 0: plain product)
        float4 h, l;
        h.x = rn_tf32(x.x); h.y = rn_tf32(x.y); h.z = rn_tf32(x.z); h.w = rn_tf32(x.w);
        l.x = rn_tf32(x.x - h.x); l.y = rn_tf32(x.y - h.y); l.z = rn_tf32(x.z - h.z); l.w = rn_tf32(x.w - h.w);
        b_hi[i] = h;
        b_lo[i] = l;
      }
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(ready0 + 8 * s);
    }
  } else {
    // ===== accumulate warps: drain every D-block's partial product, sum in round-to-nearest fp32, store =====
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;           // row of the A operand (query index)
    float acc[kTileN];
#pragma unroll
    for (int j = 0; j < kTileN; ++j) acc[j] = 0.0f;
    for (int kb = kExternalAcc ? 0 : n_kb - 1; kb < n_kb; ++kb) {
      const int b = kExternalAcc ? kb % kAccBufs : 0;
      const uint32_t acc_use = kExternalAcc ? kb / kAccBufs : 0;
      mbar_wait(accfull0 + 8 * b, acc_use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * kTileN;
#pragma unroll
      for (int c = 0; c < kTileN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c * 32 + j] += v[j];
      }
      tc_fence_before();
      mbar_arrive(accempty0 + 8 * b);
    }
    if (row < n && ep.mode != 0) {
      // moments epilogue (rows are classes here): the arithmetic of moments_kernel (dirichlet_estep.cu)
      const long r = (long)t * n + row;
      const float cs = ep.colsum[r];
      if (ep.mode == 1) {          // zero-shot: sum u log z / max(sum u, eps), -10 for an empty cluster (em_dirichlet.py:219-222)
        const bool lv = cs > 1e-15f;
        const float den = fmaxf(cs, 1e-15f);
#pragma unroll
        for (int j = 0; j < kTileN; ++j) acc[j] = lv ? acc[j] / den : -10.0f;
      } else {                     // few-shot: (1 / (count_s + sum u)) * (support_sum + sum u log z) (few_shot/em_dirichlet.py:196-200)
        const float f = 1.0f / (ep.support_count[r] + cs);
        const float* ss = ep.support_sum + r * K + k0;
#pragma unroll
        for (int j = 0; j < kTileN; ++j) acc[j] = (k0 + j < K) ? f * (ss[j] + acc[j]) : 0.0f;
      }
    }
    if (row < n) {
      float* out = l3 + ((long)t * n + row) * K + k0;
      if (k0 + kTileN <= K && (K & 3) == 0) {
#pragma unroll
        for (int j = 0; j < kTileN; j += 4)
          *reinterpret_cast<float4*>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < kTileN; ++j)
          if (k0 + j < K) out[j] = acc[j];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

// [tasks][rows][D] float32 row-major -> boxes of 32 floats x 128 rows x 1 task, 128-byte swizzle, zero fill outside
bool make_map(CUtensorMap* m, const float* p, int tasks, int rows, int D) {
  EncodeFn enc = encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)rows, (cuuint64_t)tasks};
  const cuuint64_t strides[2] = {(cuuint64_t)D * 4, (cuuint64_t)rows * D * 4};
  const cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileM, 1};
  const cuuint32_t elem[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), dims, strides, box, elem,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool logits_tc_supported(int n, int K, int D) {
  // TMA: 16-byte aligned row pitch; one M tile of 128 queries
  return n >= 1 && n <= kTileM && K >= 1 && D >= 4 && (D % 4) == 0;
}

// C[t, m, k] = sum_d A[t, m, d] * (B[tb, k, d] - b_shift), tb = t or 0 (b_tasks == 1): 3 x TF32 on tcgen05, fp32
// round-to-nearest running sum outside the tensor core.  A [T, M, D], B [b_tasks, N, D], C [T, M, N].
cudaError_t gemm_nt_tc(const float* a, const float* b, float* c, int T, int M, int N, int D, int b_tasks, float b_shift,
                       const int* gate, bool accumulate_in_tmem, cudaStream_t st, const TcEpilogue* epilogue) {
  const TcEpilogue ep = epilogue ? *epilogue : TcEpilogue{};
  if (M < 1 || N < 1 || D < 4 || (D % 4) != 0 || (b_tasks != 1 && b_tasks != T)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) return cudaErrorMisalignedAddress;
  // cudaFuncSetAttribute is per device: opt in once per device of this process (idempotent, so a race between two host
  // threads only repeats the call)
  static PerDeviceFlags attr_set;
  const int slot = current_device_slot();
  if (slot < 0 || attr_set.v[slot].load(std::memory_order_acquire) == 0) {
    cudaError_t e = cudaFuncSetAttribute(logits_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(logits_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    if (slot >= 0) attr_set.v[slot].store(1, std::memory_order_release);
  }
  CUtensorMap ma, mb;
  if (!make_map(&ma, a, T, M, D) || !make_map(&mb, b, b_tasks, N, D)) return cudaErrorInvalidValue;
  const int m_tiles = (M + kTileM - 1) / kTileM;
  if (m_tiles > 65535 || T > 65535) return cudaErrorInvalidValue;
  dim3 grid((N + kTileN - 1) / kTileN, m_tiles, T);
  if (accumulate_in_tmem)
    logits_tc_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, c, M, N, D, gate, b_shift, b_tasks == 1 && T > 1, ep);
  else
    logits_tc_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, c, M, N, D, gate, b_shift, b_tasks == 1 && T > 1, ep);
  note_launch();
  return cudaGetLastError();
}

cudaError_t logits_tc(const float* logz, const float* alpha, float* l3, int T, int n, int K, int D, const int* gate,
                      bool accumulate_in_tmem, cudaStream_t st) {
  if (!logits_tc_supported(n, K, D)) return cudaErrorInvalidValue;
  return gemm_nt_tc(logz, alpha, l3, T, n, K, D, T, 1.0f, gate, accumulate_in_tmem, st);
}

}  // namespace tclip
