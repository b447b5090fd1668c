// ob_gemv.cu - the small-batch ("GEMV-like") regime of the quantised linear forward: M <= 64 token rows.
//
//   y[m, n] = (sum_k q[m, k] * Q[n, k]) * alpha_eff / s[m] + b[n]          (quant.py:126 with int8 activations)
//
// With a handful of token rows the layer is bound by streaming the packed 2-bit weights (N*K/4 bytes) once, not by the
// tensor cores, and a 128-row UMMA tile would be >= 50 % padding.  Here the packed words are read straight from HBM/L2
// into registers, expanded with the same PRMT table as the tensor-core path (16 int8 codes per 32-bit word) and
// contracted with DP4A (exact int32); nothing is staged in shared memory except the [M x 8] output slice of a CTA.
//
//   CTA = 4 warps = 8 output features (one per half-warp), all M rows.  Lane l of a half-warp owns the 16-code words
//   l, l+16, l+32, ... of its weight row: one 4-byte weight load feeds 16 codes x M rows; the activation codes of a row
//   are read as one 16-byte vector per word (L1/L2 resident: M*K <= 128 KB).  Partial sums are reduced over the 16 lanes
//   with shuffles, then the dequantisation epilogue is the tensor-core kernel's (same expression -> same bits).
#include "ob_common.cuh"

namespace ob {

constexpr int kGemvThreads = 128;
constexpr int kGemvRowsPerCta = 8;      // output features per CTA
constexpr int kGemvMChunk = 16;         // token rows accumulated in registers at a time
constexpr int kGemvMaxM = 64;

template <int OUT_BF16>
__global__ void __launch_bounds__(kGemvThreads)
gemv_tern_i8_kernel(const int8_t* __restrict__ q, const float* __restrict__ scale, const uint8_t* __restrict__ packed,
                    const float* __restrict__ alpha, int alpha_mode, const float* __restrict__ bias, int M, int N, int K,
                    void* __restrict__ y) {
  __shared__ float out_s[kGemvMaxM][kGemvRowsPerCta];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l16 = lane & 15;
  const int n_local = warp * 2 + (lane >> 4);
  const int n0 = blockIdx.x * kGemvRowsPerCta;
  const int n = n0 + n_local;                                   // N % 8 == 0: always in range
  const int words = K >> 4;                                     // 16 codes per packed word
  const uint32_t* wrow = reinterpret_cast<const uint32_t*>(packed) + static_cast<size_t>(n) * words;
  const float a_eff = load_alpha_eff(alpha, alpha_mode);
  const float b_n = bias != nullptr ? __ldg(bias + n) : 0.f;

  for (int m0 = 0; m0 < M; m0 += kGemvMChunk) {
    int acc[kGemvMChunk];
#pragma unroll
    for (int i = 0; i < kGemvMChunk; ++i) acc[i] = 0;
#pragma unroll 4
    for (int w = l16; w < words; w += 16) {
      const uint4 c = expand_word_i8(__ldg(wrow + w));          // codes 16w .. 16w+15 of feature n
      const int8_t* qcol = q + (static_cast<size_t>(w) << 4);
#pragma unroll
      for (int i = 0; i < kGemvMChunk; ++i) {
        if (m0 + i < M) {                                       // warp-uniform
          const int4 a = __ldg(reinterpret_cast<const int4*>(qcol + static_cast<size_t>(m0 + i) * K));
          int s = __dp4a(a.x, static_cast<int>(c.x), acc[i]);
          s = __dp4a(a.y, static_cast<int>(c.y), s);
          s = __dp4a(a.z, static_cast<int>(c.z), s);
          acc[i] = __dp4a(a.w, static_cast<int>(c.w), s);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kGemvMChunk; ++i) {
      if (m0 + i < M) {
        int s = acc[i];
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (l16 == 0) {
          const float factor = __fdiv_rn(a_eff, __ldg(scale + m0 + i));
          out_s[m0 + i][n_local] = fmaf(static_cast<float>(s), factor, b_n);     // |s| <= 128 K < 2^24: exact conversion
        }
      }
    }
  }
  __syncthreads();
  if (OUT_BF16) {
    for (int m = threadIdx.x; m < M; m += kGemvThreads) {
      __nv_bfloat162 p[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(out_s[m][2 * j], out_s[m][2 * j + 1]);
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(y) + static_cast<size_t>(m) * N + n0) =
          *reinterpret_cast<const uint4*>(p);
    }
  } else {
    for (int t = threadIdx.x; t < 2 * M; t += kGemvThreads) {
      const int m = t >> 1, h = t & 1;
      *reinterpret_cast<float4*>(static_cast<float*>(y) + static_cast<size_t>(m) * N + n0 + 4 * h) =
          make_float4(out_s[m][4 * h], out_s[m][4 * h + 1], out_s[m][4 * h + 2], out_s[m][4 * h + 3]);
    }
  }
}

int small_m_limit() { return kGemvMaxM; }

int launch_gemv_tern_i8(const int8_t* q, const float* scale, const uint8_t* packed, const float* alpha, int alpha_mode,
                        const float* bias, int M, int N, int K, void* y, int out_bf16, cudaStream_t st) {
  const dim3 grid(N / kGemvRowsPerCta);
  if (out_bf16)
    gemv_tern_i8_kernel<1><<<grid, kGemvThreads, 0, st>>>(q, scale, packed, alpha, alpha_mode, bias, M, N, K, y);
  else
    gemv_tern_i8_kernel<0><<<grid, kGemvThreads, 0, st>>>(q, scale, packed, alpha, alpha_mode, bias, M, N, K, y);
  OB_LAUNCH_CHECK("gemv_tern_i8_kernel");
  return OB_OK;
}

}  // namespace ob
