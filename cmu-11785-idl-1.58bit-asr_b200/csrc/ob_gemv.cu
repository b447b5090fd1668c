// ob_gemv.cu - the small-batch ("GEMV-like") regime of the quantised linear forward: M <= 64 token rows.
//
//   y[m, n] = (sum_k q[m, k] * Q[n, k]) * alpha_eff / s[m] + b[n]          (quant.py:126 with int8 activations)
//
// With a handful of token rows the layer is bound by streaming the packed 2-bit weights (N*K/4 bytes) once, not by the
// tensor cores, and a 128-row UMMA tile would be >= 50 % padding (plus TMEM allocation, tensor-map fetch and a TMA-store
// epilogue for a few KB of output).  Here the packed words go from HBM/L2 straight into registers, are expanded with the
// same PRMT table as the tensor-core path (16 int8 codes per 32-bit word) and contracted with DP4A (exact int32).
//
//   * The activation codes q [M, K] (<= 128 KB) are staged once per CTA in shared memory, rows padded by 16 bytes so that
//     the 16-byte loads of 8 consecutive rows fall into distinct banks.
//   * A warp owns whole output features.  Lane = (token row m, k-split ks): MP = min(32, pow2 >= M) lanes enumerate the
//     rows (a lane also takes row m + 32 when M > 32), the remaining 32 / MP lanes split the K/16 weight words of the
//     feature.  M >= 32: every weight word is one broadcast load and no cross-lane reduction exists at all;
//     M = 1: the 32 lanes read 128 contiguous bytes of the weight row - a classic GEMV.  The split partials are combined
//     with log2(32 / MP) shuffles per feature.
//   * The dequantisation epilogue is the tensor-core kernel's expression (same bits); outputs go through a [64 x 8] shared
//     tile so that every global store is a full 32-byte (fp32) / 16-byte (bf16) row segment.
#include "ob_common.cuh"

namespace ob {

constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kGemvRowsPerCta = 8;      // output features per CTA, one per warp
constexpr int kGemvMaxM = 64;
constexpr int kGemvMaxSmem = 200 * 1024;

// MP: lanes that enumerate token rows (power of two, <= 32); TWO: lanes also own row m + 32
template <int MP, int TWO, int OUT_BF16>
__global__ void __launch_bounds__(kGemvThreads)
gemv_tern_i8_kernel(const int8_t* __restrict__ q, const float* __restrict__ scale, const uint8_t* __restrict__ packed,
                    const float* __restrict__ alpha, int alpha_mode, const float* __restrict__ bias, int M, int N, int K,
                    void* __restrict__ y) {
  pdl_entry();
  extern __shared__ uint8_t gemv_smem[];
  __shared__ float out_s[kGemvMaxM][kGemvRowsPerCta];
  constexpr int KS = 32 / MP;                                   // lanes splitting the contraction
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = lane & (MP - 1), ks = lane / MP;
  const int pitch = K + 16;                                     // bytes per staged row
  const int words = K >> 4;                                     // 16 codes per packed word
  // ---- stage q: 16-byte chunks, coalesced
  {
    const int chunks_per_row = K >> 4;
    const int total = M * chunks_per_row;
    for (int i = threadIdx.x; i < total; i += kGemvThreads) {
      const int r = i / chunks_per_row, c = i - r * chunks_per_row;
      const int4 v = __ldg(reinterpret_cast<const int4*>(q + static_cast<size_t>(r) * K) + c);
      *reinterpret_cast<int4*>(gemv_smem + r * pitch + (c << 4)) = v;
    }
  }
  __syncthreads();
  const int n_local = warp;
  const int n0 = blockIdx.x * kGemvRowsPerCta;
  const int n = n0 + n_local;                                   // N % 8 == 0: always in range
  const uint32_t* wrow = reinterpret_cast<const uint32_t*>(packed) + static_cast<size_t>(n) * words;
  const bool row0_ok = m < M, row1_ok = TWO && (m + 32 < M);
  const uint8_t* q0 = gemv_smem + (row0_ok ? m : 0) * pitch;
  const uint8_t* q1 = gemv_smem + (row1_ok ? m + 32 : 0) * pitch;
  int acc0 = 0, acc1 = 0;
#pragma unroll 4
  for (int w = ks; w < words; w += KS) {
    const uint4 c = expand_word_i8(__ldg(wrow + w));            // codes 16w .. 16w+15 of feature n
    const int4 a = *reinterpret_cast<const int4*>(q0 + (w << 4));
    acc0 = __dp4a(a.x, static_cast<int>(c.x), acc0);
    acc0 = __dp4a(a.y, static_cast<int>(c.y), acc0);
    acc0 = __dp4a(a.z, static_cast<int>(c.z), acc0);
    acc0 = __dp4a(a.w, static_cast<int>(c.w), acc0);
    if (TWO) {
      const int4 b = *reinterpret_cast<const int4*>(q1 + (w << 4));
      acc1 = __dp4a(b.x, static_cast<int>(c.x), acc1);
      acc1 = __dp4a(b.y, static_cast<int>(c.y), acc1);
      acc1 = __dp4a(b.z, static_cast<int>(c.z), acc1);
      acc1 = __dp4a(b.w, static_cast<int>(c.w), acc1);
    }
  }
#pragma unroll
  for (int o = 16; o >= MP; o >>= 1) {                          // combine the k-split partials (none when MP == 32)
    acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
    if (TWO) acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
  }
  const float a_eff = load_alpha_eff(alpha, alpha_mode);
  const float b_n = bias != nullptr ? __ldg(bias + n) : 0.f;
  if (ks == 0) {
    if (row0_ok)                                                // |acc| <= 128 K < 2^24: exact conversion
      out_s[m][n_local] = fmaf(static_cast<float>(acc0), __fdiv_rn(a_eff, __ldg(scale + m)), b_n);
    if (row1_ok)
      out_s[m + 32][n_local] = fmaf(static_cast<float>(acc1), __fdiv_rn(a_eff, __ldg(scale + m + 32)), b_n);
  }
  __syncthreads();
  if (OUT_BF16) {
    for (int r = threadIdx.x; r < M; r += kGemvThreads) {
      __nv_bfloat162 p[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(out_s[r][2 * j], out_s[r][2 * j + 1]);
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(y) + static_cast<size_t>(r) * N + n0) =
          *reinterpret_cast<const uint4*>(p);
    }
  } else {
    for (int t = threadIdx.x; t < 2 * M; t += kGemvThreads) {
      const int r = t >> 1, h = t & 1;
      *reinterpret_cast<float4*>(static_cast<float*>(y) + static_cast<size_t>(r) * N + n0 + 4 * h) =
          make_float4(out_s[r][4 * h], out_s[r][4 * h + 1], out_s[r][4 * h + 2], out_s[r][4 * h + 3]);
    }
  }
}

// largest M served by this kernel for a given K.  Measured against the 128-row tcgen05 tile (bench.py `small_batch`): the
// DP4A kernel wins while the staged activations are small - M * K <= 32 K codes (M = 64 at K <= 512, M = 8 at K = 2048:
// 2.2-4.8 us vs 4.7-9.7 us) - and loses beyond (M = 64, K = 1024: 8.7 vs 7.3 us), where the tensor cores take over.
int small_m_limit(int K) {
  const int by_work = 32768 / K;
  return by_work < kGemvMaxM ? by_work : kGemvMaxM;
}
// what the kernel CAN serve (tests / measurements force it with ob_debug_set): activations staged in shared memory
int small_m_capacity(int K) {
  const int by_smem = kGemvMaxSmem / (K + 16);
  return by_smem < kGemvMaxM ? by_smem : kGemvMaxM;
}

template <int MP, int TWO>
static int launch_gemv_variant(const int8_t* q, const float* scale, const uint8_t* packed, const float* alpha, int alpha_mode,
                               const float* bias, int M, int N, int K, void* y, int out_bf16, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(M) * (K + 16);
  const dim3 grid(N / kGemvRowsPerCta);
  if (out_bf16) {
    auto kern = gemv_tern_i8_kernel<MP, TWO, 1>;
    static bool attr_set = false;
    if (!attr_set) { OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem)); attr_set = true; }
    launch_k((kern), dim3(grid), dim3(kGemvThreads), smem, st, q, scale, packed, alpha, alpha_mode, bias, M, N, K, y);
  } else {
    auto kern = gemv_tern_i8_kernel<MP, TWO, 0>;
    static bool attr_set = false;
    if (!attr_set) { OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemvMaxSmem)); attr_set = true; }
    launch_k((kern), dim3(grid), dim3(kGemvThreads), smem, st, q, scale, packed, alpha, alpha_mode, bias, M, N, K, y);
  }
  OB_LAUNCH_CHECK("gemv_tern_i8_kernel");
  return OB_OK;
}

int launch_gemv_tern_i8(const int8_t* q, const float* scale, const uint8_t* packed, const float* alpha, int alpha_mode,
                        const float* bias, int M, int N, int K, void* y, int out_bf16, cudaStream_t st) {
#define OB_GEMV_ARGS q, scale, packed, alpha, alpha_mode, bias, M, N, K, y, out_bf16, st
  if (M > 32) return launch_gemv_variant<32, 1>(OB_GEMV_ARGS);
  if (M > 16) return launch_gemv_variant<32, 0>(OB_GEMV_ARGS);
  if (M > 8) return launch_gemv_variant<16, 0>(OB_GEMV_ARGS);
  if (M > 4) return launch_gemv_variant<8, 0>(OB_GEMV_ARGS);
  if (M > 2) return launch_gemv_variant<4, 0>(OB_GEMV_ARGS);
  if (M > 1) return launch_gemv_variant<2, 0>(OB_GEMV_ARGS);
  return launch_gemv_variant<1, 0>(OB_GEMV_ARGS);
#undef OB_GEMV_ARGS
}

}  // namespace ob
