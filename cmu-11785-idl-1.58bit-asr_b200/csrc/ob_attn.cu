// ob_attn.cu - the element-wise chain between the attention matmuls of the reference's MHSA (conformer.py:118-128):
//   scores = (ac + rel_shift(bd)) / sqrt(d_head) -> masked_fill(-inf) -> softmax -> nan_to_num -> dropout
// fused into one forward and one backward kernel (nine / twelve torch kernels over [B,H,T,T] fp32 otherwise).
// One warp per (b, h, i) row held in registers; the relative shift (conformer.py:96-103: pad one column, view as
// [T+1, T], drop the first row) is pure index arithmetic: element (i, j) reads P[f / (T+1)][f % (T+1)] with
// f = T + i*T + j, P = [0 | bd].  SURVEY.md section 8f ranks 1/3 (callers around the routed projections).
#include "ob_common.cuh"

namespace ob {

__device__ __forceinline__ float wmax_(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum_(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NJ>
struct RowBits { using type = uint32_t; };                       // one flag per element a lane owns
template <>
struct RowBits<64> { using type = unsigned long long; };

// dropout keep flags of the NJ elements a lane owns in one row: element u is lane u % 8 of the Philox block with counter
// ((row * 32 + lane) * 8 + u / 8, offset)
template <int NJ>
__device__ __forceinline__ typename RowBits<NJ>::type rng_keep_bits(const DropRng& rng, int64_t row, int lane) {
  using bits_t = typename RowBits<NJ>::type;
  bits_t bits = 0;
#pragma unroll
  for (int g = 0; g < (NJ + 7) / 8; ++g)
    bits |= static_cast<bits_t>(philox_keep8((static_cast<unsigned long long>(row) * 32 + lane) * 8 + g, rng)) << (8 * g);
  return bits;
}

// NJ = ceil(T / 32) elements per lane
template <int NJ>
__global__ void __launch_bounds__(256, 3) relattn_softmax_fwd_kernel(const float* __restrict__ ac, const float* __restrict__ bd,
                                                                  const uint8_t* __restrict__ mask /* [B,T,T] */,
                                                                  const uint8_t* __restrict__ keep /* [B,H,T,T] | null */,
                                                                  float inv_keep, DropRng rng, float scale, int B, int H, int T, int ld,
                                                                  float* __restrict__ y, float* __restrict__ attn_d) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)B * H * T;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const int i = (int)(row % T);
    const int64_t bh = row / T;
    const int b = (int)(bh / H);
    const float* ac_r = ac + row * ld;                           // ac, bd, y, attn_d: rows of T floats, pitch ld
    const uint8_t* m_r = mask + ((int64_t)b * T + i) * T;
    const uint8_t* k_r = keep != nullptr ? keep + row * T : m_r;
    const int f0 = T + i * T;
    const int r0 = f0 / (T + 1), c0 = f0 - r0 * (T + 1);
    // element j of the shifted row is P[r0][c0 + j] or, past the end of that row of [0 | bd], P[r0 + 1][c0 + j - (T+1)];
    // relative to bd row r0 both are one linear offset: c0 + j - 1 (same row) or c0 + j - 1 + ld - (T+1) (next row)
    const float* bd_r0 = bd + (bh * (int64_t)T + r0) * ld;
    const int wrap_off = ld - T - 1;
    float s[NJ];
    float mx = -INFINITY;
    float av[NJ], bv[NJ];
    using bits_t = typename RowBits<NJ>::type;
    bits_t mbits = 0, kbits = 0, zbits = 0;                     // mask / keep / zero-column flags, one bit per element
#pragma unroll
    for (int u = 0; u < NJ; ++u) {
      const int j = min(lane + 32 * u, T - 1);
      const int t = c0 + j;
      const int wrapped = t >= T + 1;
      const uint32_t mk = __ldg(m_r + j);
      const uint32_t kp = keep != nullptr ? __ldg(k_r + j) : 1u;  // no explicit mask: the bits come from the RNG below
      av[u] = __ldg(ac_r + j);
      bv[u] = __ldg(bd_r0 + max(t - 1 + (wrapped ? wrap_off : 0), 0));
      zbits |= static_cast<bits_t>(t == 0 || t == T + 1) << u;   // the zero column of [0 | bd]
      mbits |= static_cast<bits_t>(mk != 0u) << u;
      kbits |= static_cast<bits_t>(kp != 0u) << u;
    }
#pragma unroll
    for (int u = 0; u < NJ; ++u) {
      s[u] = -INFINITY;
      if (lane + 32 * u < T && ((mbits >> u) & 1u)) {
        s[u] = (av[u] + (((zbits >> u) & 1u) ? 0.f : bv[u])) * scale;
        mx = fmaxf(mx, s[u]);
      }
    }
    mx = wmax_(mx);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < NJ; ++u) {
      s[u] = (s[u] == -INFINITY) ? 0.f : __expf(s[u] - mx);    // SFU exponential (2 ulp); masked terms are exactly 0
      sum += s[u];
    }
    sum = wsum_(sum);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;               // nan_to_num(nan = 0) of the reference (conformer.py:127)
    if (keep == nullptr && rng.threshold != 0u) kbits = rng_keep_bits<NJ>(rng, row, lane);
#pragma unroll
    for (int u = 0; u < NJ; ++u) {
      const int j = lane + 32 * u;
      if (j < T) {
        const float p = s[u] * inv;
        y[row * ld + j] = p;
        if (attn_d != nullptr) attn_d[row * ld + j] = ((kbits >> u) & 1u) ? p * inv_keep : 0.f;
      }
    }
  }
}

template <int NJ>
__global__ void __launch_bounds__(256, 3) relattn_softmax_bwd_kernel(const float* __restrict__ gd, const float* __restrict__ y,
                                                                  const uint8_t* __restrict__ keep, float inv_keep,
                                                                  DropRng rng, float scale, int B, int H, int T, int ld,
                                                                  float* __restrict__ d_ac, float* __restrict__ d_bd) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)B * H * T;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < rows; row += nwarps) {
    const int i = (int)(row % T);
    const int64_t bh = row / T;
    float* dbd_bh = d_bd + bh * (int64_t)T * ld;
    const int f0 = T + i * T;
    const int r0 = f0 / (T + 1), c0 = f0 - r0 * (T + 1);
    float g[NJ], p[NJ];
    uint8_t kp[NJ];
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < NJ; ++u) {                                // all loads first, clamped index, no predicate
      const int j = min(lane + 32 * u, T - 1);
      p[u] = __ldg(y + row * ld + j);
      g[u] = __ldg(gd + row * ld + j);
      kp[u] = keep != nullptr ? __ldg(keep + row * T + j) : (uint8_t)1;
    }
    const bool use_rng = keep == nullptr && rng.threshold != 0u;
    const typename RowBits<NJ>::type rbits = use_rng ? rng_keep_bits<NJ>(rng, row, lane) : 0;
    const float gscale = (keep != nullptr || use_rng) ? inv_keep : 1.0f;
#pragma unroll
    for (int u = 0; u < NJ; ++u) {
      if (lane + 32 * u < T) {
        const bool kept = use_rng ? ((rbits >> u) & 1u) != 0u : kp[u] != 0;
        g[u] = kept ? g[u] * gscale : 0.f;
        dot += g[u] * p[u];
      } else {
        g[u] = 0.f;
        p[u] = 0.f;
      }
    }
    dot = wsum_(dot);
#pragma unroll
    for (int u = 0; u < NJ; ++u) {
      const int j = lane + 32 * u;
      if (j < T) {
        const float ds = p[u] * (g[u] - dot) * scale;
        d_ac[row * ld + j] = ds;
        int c = c0 + j, r = r0;
        if (c >= T + 1) { c -= T + 1; r += 1; }
        if (c > 0) dbd_bh[(int64_t)r * ld + (c - 1)] = ds;        // c == 0 is the padded column: no gradient
      }
    }
    if (i == 0)                                                   // positions of [0 | bd] in front of the dropped row
      for (int j = lane; j < T - 1; j += 32) dbd_bh[j] = 0.f;
  }
}

}  // namespace ob

using namespace ob;

static int attn_blocks(int64_t rows) {
  const int64_t want = (rows + 7) / 8;
  return (int)(want < 148 * 16 ? want : 148 * 16);
}

#define OB_ATTN_DISPATCH(KERNEL, ...)                                         \
  do {                                                                        \
    const int nj = (T + 31) / 32;                                             \
    if (nj <= 8) launch_k((KERNEL<8>), dim3(blocks), dim3(256), 0, st, __VA_ARGS__);              \
    else if (nj <= 13) launch_k((KERNEL<13>), dim3(blocks), dim3(256), 0, st, __VA_ARGS__);       \
    else if (nj <= 16) launch_k((KERNEL<16>), dim3(blocks), dim3(256), 0, st, __VA_ARGS__);       \
    else if (nj <= 32) launch_k((KERNEL<32>), dim3(blocks), dim3(256), 0, st, __VA_ARGS__);       \
    else launch_k((KERNEL<64>), dim3(blocks), dim3(256), 0, st, __VA_ARGS__);                     \
  } while (0)

extern "C" int ob_relattn_softmax_fwd(const float* ac, const float* bd, const uint8_t* mask, const uint8_t* keep,
                                      float inv_keep, uint64_t seed, uint64_t offset, uint32_t drop_threshold, float scale,
                                      int B, int H, int T, int ld, float* y, float* attn_d, ob_stream_t stream) {
  OB_REQUIRE(ac && bd && mask && y, "ob_relattn_softmax_fwd: null pointer");
  OB_REQUIRE((keep != nullptr || drop_threshold != 0u) == (attn_d != nullptr),
             "ob_relattn_softmax_fwd: attn_d is written exactly when dropout (mask or RNG) is on");
  const DropRng rng = {seed, offset, drop_threshold};
  OB_REQUIRE(drop_threshold < 65536u, "ob_relattn_softmax_fwd: drop_threshold (%u) is a 16-bit value", drop_threshold);
  OB_REQUIRE(B > 0 && H > 0 && T > 0 && T <= 2048, "ob_relattn_softmax_fwd: need 0 < T <= 2048 (T=%d)", T);
  OB_REQUIRE(ld >= T, "ob_relattn_softmax_fwd: row pitch ld (%d) must be >= T (%d)", ld, T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = attn_blocks((int64_t)B * H * T);
  OB_ATTN_DISPATCH(relattn_softmax_fwd_kernel, ac, bd, mask, keep, inv_keep, rng, scale, B, H, T, ld, y, attn_d);
  OB_LAUNCH_CHECK("relattn_softmax_fwd_kernel");
  return OB_OK;
}

extern "C" int ob_relattn_softmax_bwd(const float* gd, const float* y, const uint8_t* keep, float inv_keep, uint64_t seed,
                                      uint64_t offset, uint32_t drop_threshold, float scale, int B, int H, int T, int ld,
                                      float* d_ac, float* d_bd, ob_stream_t stream) {
  OB_REQUIRE(gd && y && d_ac && d_bd, "ob_relattn_softmax_bwd: null pointer");
  const DropRng rng = {seed, offset, drop_threshold};
  OB_REQUIRE(drop_threshold < 65536u, "ob_relattn_softmax_bwd: drop_threshold (%u) is a 16-bit value", drop_threshold);
  OB_REQUIRE(B > 0 && H > 0 && T > 0 && T <= 2048, "ob_relattn_softmax_bwd: need 0 < T <= 2048 (T=%d)", T);
  OB_REQUIRE(ld >= T, "ob_relattn_softmax_bwd: row pitch ld (%d) must be >= T (%d)", ld, T);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = attn_blocks((int64_t)B * H * T);
  OB_ATTN_DISPATCH(relattn_softmax_bwd_kernel, gd, y, keep, inv_keep, rng, scale, B, H, T, ld, d_ac, d_bd);
  OB_LAUNCH_CHECK("relattn_softmax_bwd_kernel");
  return OB_OK;
}
