// ob_conv.cu - the middle of the reference's convolution module (conformer.py:141-167) in the [B, T, C] layout of the
// rest of the encoder:   a = pw1(ln(x))  ->  g = GLU(a)  ->  d = depthwise_conv1d(g)  ->  y = BatchNorm(d)  ->  s = swish(y).
// The two 1x1 convolutions around it are plain matrix products over the channel axis (ob_gemm_f32); everything between
// them is HBM-bound stencil / element-wise / reduction work, done here in as few passes as the data dependencies allow:
//
//   glu_dwconv_fwd      a -> d (+ per-tile BatchNorm partials: sum and centred sum of squares, Chan-combined in fp64)
//   bn_swish_fwd        d -> s
//   bn_swish_bwd_reduce g_s, d -> per-row-block sums of g_y and g_y * xhat        (g_y = g_s * swish'(y))
//   bn_swish_bwd_apply  g_s, d -> g_d
//   dwconv_bwd_data_glu g_d, a -> g_a (depthwise transposed convolution + GLU backward)
//   dwconv_bwd_weight   g_d, a -> per-tile partials of g_w and g_bias, reduced in fixed order
//
// BatchNorm uses batch statistics over all B*T frames including padding, biased variance, exactly as
// nn.BatchNorm1d(track_running_stats=False) does in the reference (conformer.py:148), in training and in eval mode.
// The depthwise kernel has ks <= 31 odd taps (reference default 31); shorter kernels are centred in the 31-tap window.
#include "ob_common.cuh"

namespace ob {

constexpr int kCvT = 64;                  // frames per tile
constexpr int kCvC = 64;                  // channels per tile
constexpr int kCvHalo = 15;
constexpr int kCvTaps = 31;
constexpr int kCvRows = kCvT + 2 * kCvHalo;   // 94
constexpr int kCvPer = kCvT / 4;          // outputs per thread (4 time groups of 64 channels = 256 threads)
constexpr int kCvWin = kCvPer + kCvTaps - 1;

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// centred 31-tap window of channel c; flip = 1 for the transposed convolution
__device__ __forceinline__ void load_taps(const float* __restrict__ w, int c, int ks, bool flip, float (&wr)[kCvTaps]) {
  const int off = (kCvTaps - ks) / 2;
#pragma unroll
  for (int j = 0; j < kCvTaps; ++j) {
    const int jj = flip ? kCvTaps - 1 - j : j;
    wr[j] = (jj >= off && jj < off + ks) ? __ldg(w + c * ks + jj - off) : 0.f;
  }
}

// tile of g = GLU(a) with halo into shared memory: rows t0-15 .. t0+78, zero outside the utterance.  128-bit loads: a
// thread owns 4 channels of a row (16 threads per 256-byte row half), all loads of a pass are issued before they are used
__device__ __forceinline__ void load_glu_tile(const float* __restrict__ a, int b, int T, int C, int t0, int c0, float* tile) {
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;          // 16 rows per pass
  constexpr int kPasses = (kCvRows + 15) / 16;
  float4 v1[kPasses], v2[kPasses];
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const int r = r0 + 16 * p, t = t0 - kCvHalo + r;
    v1[p] = v2[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < kCvRows && t >= 0 && t < T) {
      const float* row = a + (static_cast<int64_t>(b) * T + t) * (2 * C) + c0 + c4;
      v1[p] = __ldg(reinterpret_cast<const float4*>(row));
      v2[p] = __ldg(reinterpret_cast<const float4*>(row + C));
    }
  }
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const int r = r0 + 16 * p;
    if (r < kCvRows)
      *reinterpret_cast<float4*>(tile + r * kCvC + c4) = make_float4(v1[p].x * sigmoid_f(v2[p].x), v1[p].y * sigmoid_f(v2[p].y),
                                                                     v1[p].z * sigmoid_f(v2[p].z), v1[p].w * sigmoid_f(v2[p].w));
  }
}

__device__ __forceinline__ void load_plain_tile(const float* __restrict__ x, int b, int T, int C, int t0, int c0, float* tile) {
  const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;
  constexpr int kPasses = (kCvRows + 15) / 16;
  float4 v[kPasses];
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const int r = r0 + 16 * p, t = t0 - kCvHalo + r;
    v[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < kCvRows && t >= 0 && t < T) v[p] = __ldg(reinterpret_cast<const float4*>(x + (static_cast<int64_t>(b) * T + t) * C + c0 + c4));
  }
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const int r = r0 + 16 * p;
    if (r < kCvRows) *reinterpret_cast<float4*>(tile + r * kCvC + c4) = v[p];
  }
}

// grid: x = b * tblocks + tb, y = channel tile.  part: [B * tblocks][2][C] = (sum, centred sum of squares) per tile.
__global__ void __launch_bounds__(256) glu_dwconv_fwd_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                             const float* __restrict__ bias, int B, int T, int C, int ks,
                                                             float* __restrict__ d, float* __restrict__ part) {
  pdl_entry();
  __shared__ __align__(16) float tile[kCvRows * kCvC];
  __shared__ float red[4][kCvC];
  const int tblocks = (T + kCvT - 1) / kCvT;
  const int b = blockIdx.x / tblocks, tb = blockIdx.x % tblocks;
  const int t0 = tb * kCvT, c0 = blockIdx.y * kCvC;
  const int c = threadIdx.x & 63, tg = threadIdx.x >> 6;
  load_glu_tile(a, b, T, C, t0, c0, tile);
  float wr[kCvTaps];
  load_taps(w, c0 + c, ks, false, wr);
  const float bv = bias != nullptr ? __ldg(bias + c0 + c) : 0.f;
  __syncthreads();
  float win[kCvWin];
#pragma unroll
  for (int i = 0; i < kCvWin; ++i) win[i] = tile[(tg * kCvPer + i) * kCvC + c];
  float out[kCvPer];
  float sum = 0.f;
#pragma unroll
  for (int o = 0; o < kCvPer; ++o) {
    float acc = bv;
#pragma unroll
    for (int j = 0; j < kCvTaps; ++j) acc = fmaf(wr[j], win[o + j], acc);
    out[o] = acc;
    const int t = t0 + tg * kCvPer + o;
    if (t < T) {
      d[(static_cast<int64_t>(b) * T + t) * C + c0 + c] = acc;
      sum += acc;
    }
  }
  // BatchNorm partials of this tile: sum, then the sum of squares centred at the tile mean (stable to combine)
  const int n_tile = min(kCvT, T - t0);
  red[tg][c] = sum;
  __syncthreads();
  const float tsum = red[0][c] + red[1][c] + red[2][c] + red[3][c];
  const float tmean = tsum / static_cast<float>(n_tile);
  float m2 = 0.f;
#pragma unroll
  for (int o = 0; o < kCvPer; ++o)
    if (t0 + tg * kCvPer + o < T) m2 = fmaf(out[o] - tmean, out[o] - tmean, m2);
  __syncthreads();
  red[tg][c] = m2;
  __syncthreads();
  if (tg == 0) {
    float* p = part + static_cast<int64_t>(blockIdx.x) * 2 * C + c0 + c;
    p[0] = tsum;
    p[C] = red[0][c] + red[1][c] + red[2][c] + red[3][c];
  }
}

// Combination of the tile partials (sum, centred sum of squares) in fp64, fixed order: 16 thread groups each fold every
// 16th tile, then the 16 results are folded.  grid: C / 64 blocks of 1024 threads.
constexpr int kFinGroups = 16;

// (16 channels x 64 tile groups per block: C / 16 blocks instead of C / 64 - with four blocks the kernel was a 20 us chain of
// dependent fp64 adds on the critical path between the depthwise convolution and the normalisation)
constexpr int kStatCh = 16, kStatGroups = 64;

__global__ void __launch_bounds__(kStatCh * kStatGroups) bn_stats_finalize_kernel(const float* __restrict__ part, int B, int T, int C,
                                                                          float eps, float* __restrict__ mean,
                                                                          float* __restrict__ rstd) {
  pdl_entry();
  // total M2 = sum_tiles [ M2_tile + n_tile * (mean_tile - mean)^2 ]  (the pairwise Chan update, summed in closed form)
  __shared__ double sh[kStatGroups][kStatCh];
  const int cl = threadIdx.x & (kStatCh - 1), grp = threadIdx.x / kStatCh;
  const int c = blockIdx.x * kStatCh + cl;
  const int tblocks = (T + kCvT - 1) / kCvT, nblk = B * tblocks;
  const double n_tail = static_cast<double>(T - (tblocks - 1) * kCvT);
  double s = 0.0;
#pragma unroll 4
  for (int blk = grp; blk < nblk; blk += kStatGroups) s += part[static_cast<int64_t>(blk) * 2 * C + c];
  sh[grp][cl] = s;
  __syncthreads();
  double tot = 0.0;
  for (int g = 0; g < kStatGroups; ++g) tot += sh[g][cl];
  const double n = static_cast<double>(B) * T, mu = tot / n;
  __syncthreads();
  double m2 = 0.0;
#pragma unroll 4
  for (int blk = grp; blk < nblk; blk += kStatGroups) {
    const bool tail = (blk % tblocks) == tblocks - 1;
    const double nb = tail ? n_tail : static_cast<double>(kCvT);
    const double dm = part[static_cast<int64_t>(blk) * 2 * C + c] * (tail ? 1.0 / n_tail : 1.0 / kCvT) - mu;
    m2 += part[static_cast<int64_t>(blk) * 2 * C + C + c] + nb * dm * dm;
  }
  sh[grp][cl] = m2;
  __syncthreads();
  if (grp == 0) {
    double v = 0.0;
    for (int g = 0; g < kStatGroups; ++g) v += sh[g][cl];
    mean[c] = static_cast<float>(mu);
    rstd[c] = static_cast<float>(1.0 / sqrt(v / n + static_cast<double>(eps)));
  }
}

// out[i] = sum over blocks of part[blk * stride + i] for 64 consecutive items per CTA, fp64, fixed order (same grouping)
__device__ __forceinline__ double grouped_block_sum(const float* __restrict__ part, int nblocks, int64_t stride, int item,
                                                    double (*sh)[64]) {
  const int cl = threadIdx.x & 63, grp = threadIdx.x >> 6;
  double acc = 0.0;
#pragma unroll 4
  for (int blk = grp; blk < nblocks; blk += kFinGroups) acc += part[static_cast<int64_t>(blk) * stride + item];
  sh[grp][cl] = acc;
  __syncthreads();
  if (grp == 0)
    for (int g = 1; g < kFinGroups; ++g) acc += sh[g][cl];
  return acc;
}

// s = swish(gamma * (d - mean) * rstd + beta);  C % 4 == 0
__global__ void __launch_bounds__(256) bn_swish_fwd_kernel(const float* __restrict__ d, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, int64_t n4, int C4,
                                                           float* __restrict__ s) {
  pdl_entry();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int c4 = static_cast<int>(i % C4);
    const float4 x = __ldg(reinterpret_cast<const float4*>(d) + i);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c4), rs = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 y;
    y.x = fmaf((x.x - mu.x) * rs.x, ga.x, be.x);
    y.y = fmaf((x.y - mu.y) * rs.y, ga.y, be.y);
    y.z = fmaf((x.z - mu.z) * rs.z, ga.z, be.z);
    y.w = fmaf((x.w - mu.w) * rs.w, ga.w, be.w);
    reinterpret_cast<float4*>(s)[i] = make_float4(y.x * sigmoid_f(y.x), y.y * sigmoid_f(y.y), y.z * sigmoid_f(y.z),
                                                  y.w * sigmoid_f(y.w));
  }
}

__device__ __forceinline__ float swish_grad(float y) {
  const float sg = sigmoid_f(y);
  return sg + y * sg * (1.0f - sg);
}

constexpr int kBnRows = 128;              // rows per reduction block

// partial sums over a block of rows: part[blk][0][c] = sum g_y, part[blk][1][c] = sum g_y * xhat
__global__ void __launch_bounds__(256) bn_swish_bwd_reduce_kernel(const float* __restrict__ gs, const float* __restrict__ d,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  int64_t M, int C, float* __restrict__ part) {
  pdl_entry();
  // a thread owns 4 channels (128-bit loads) of every 16th row of the block; fixed-order fold over the 16 row groups
  __shared__ float4 red[2][16][16];
  const int c4 = threadIdx.x & 15, rg = threadIdx.x >> 4;
  const int cc = blockIdx.y * kCvC + c4 * 4;
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean + cc)), rs = __ldg(reinterpret_cast<const float4*>(rstd + cc));
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + cc)), be = __ldg(reinterpret_cast<const float4*>(beta + cc));
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kBnRows;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int r = rg; r < kBnRows; r += 16) {
    const int64_t row = r0 + r;
    if (row < M) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(d + row * C + cc));
      const float4 g = __ldg(reinterpret_cast<const float4*>(gs + row * C + cc));
      float xh, gy;
      xh = (x.x - mu.x) * rs.x, gy = g.x * swish_grad(fmaf(xh, ga.x, be.x)), s1.x += gy, s2.x = fmaf(gy, xh, s2.x);
      xh = (x.y - mu.y) * rs.y, gy = g.y * swish_grad(fmaf(xh, ga.y, be.y)), s1.y += gy, s2.y = fmaf(gy, xh, s2.y);
      xh = (x.z - mu.z) * rs.z, gy = g.z * swish_grad(fmaf(xh, ga.z, be.z)), s1.z += gy, s2.z = fmaf(gy, xh, s2.z);
      xh = (x.w - mu.w) * rs.w, gy = g.w * swish_grad(fmaf(xh, ga.w, be.w)), s1.w += gy, s2.w = fmaf(gy, xh, s2.w);
    }
  }
  red[0][rg][c4] = s1;
  red[1][rg][c4] = s2;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int which = threadIdx.x >> 4;
    float4 acc = red[which][0][c4];
    for (int g = 1; g < 16; ++g) {
      const float4 v = red[which][g][c4];
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    *reinterpret_cast<float4*>(part + (static_cast<int64_t>(blockIdx.x) * 2 + which) * C + cc) = acc;
  }
}

// sums[0][c] = sum g_y (= g_beta), sums[1][c] = sum g_y * xhat (= g_gamma).  grid: 2C / 64 blocks of 1024 threads
__global__ void __launch_bounds__(64 * kFinGroups) bn_bwd_finalize_kernel(const float* __restrict__ part, int nblocks, int C,
                                                                        float* __restrict__ sums) {
  pdl_entry();
  __shared__ double sh[kFinGroups][64];
  const int item = blockIdx.x * 64 + (threadIdx.x & 63);
  const double acc = grouped_block_sum(part, nblocks, 2 * static_cast<int64_t>(C), item, sh);
  if (threadIdx.x < 64) sums[item] = static_cast<float>(acc);
}

// g_d = gamma * rstd * (g_y - sum(g_y) / M - xhat * sum(g_y * xhat) / M)
__global__ void __launch_bounds__(256) bn_swish_bwd_apply_kernel(const float* __restrict__ gs, const float* __restrict__ d,
                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const float* __restrict__ sums, float inv_m, int64_t n4, int C4,
                                                                 float* __restrict__ gd) {
  pdl_entry();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int c4 = static_cast<int>(i % C4);
    const float4 x = __ldg(reinterpret_cast<const float4*>(d) + i), g = __ldg(reinterpret_cast<const float4*>(gs) + i);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c4), rs = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(sums) + c4), s2 = __ldg(reinterpret_cast<const float4*>(sums) + C4 + c4);
    const float xs[4] = {x.x, x.y, x.z, x.w}, gv[4] = {g.x, g.y, g.z, g.w}, mus[4] = {mu.x, mu.y, mu.z, mu.w};
    const float rss[4] = {rs.x, rs.y, rs.z, rs.w}, gas[4] = {ga.x, ga.y, ga.z, ga.w}, bes[4] = {be.x, be.y, be.z, be.w};
    const float s1s[4] = {s1.x, s1.y, s1.z, s1.w}, s2s[4] = {s2.x, s2.y, s2.z, s2.w};
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float xh = (xs[u] - mus[u]) * rss[u];
      const float gy = gv[u] * swish_grad(fmaf(xh, gas[u], bes[u]));
      o[u] = gas[u] * rss[u] * (gy - s1s[u] * inv_m - xh * s2s[u] * inv_m);
    }
    reinterpret_cast<float4*>(gd)[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// g_g = transposed depthwise convolution of g_d, then the GLU backward into both halves of g_a.  The a tile the GLU backward
// needs (both halves, 2 x [64 x 64]) is fetched with the g_d tile in one round of 128-bit loads and kept in shared memory, so
// the stencil is followed by shared-memory reads instead of a second, late round of global loads (the kernel is latency-bound:
// ~100 registers allow two blocks per SM).
constexpr int kCvBwdSmemBytes = (kCvRows * kCvC + 2 * kCvT * kCvC) * 4;

__global__ void __launch_bounds__(256) dwconv_bwd_data_glu_kernel(const float* __restrict__ gd, const float* __restrict__ a,
                                                                  const float* __restrict__ w, int B, int T, int C, int ks,
                                                                  float* __restrict__ ga) {
  pdl_entry();
  extern __shared__ __align__(16) float cv_smem[];
  float* tile = cv_smem;                              // [94][64] g_d with halo
  float* a_tile = cv_smem + kCvRows * kCvC;           // [2][64][64]: a1 (values), a2 (gates)
  const int tblocks = (T + kCvT - 1) / kCvT;
  const int b = blockIdx.x / tblocks, tb = blockIdx.x % tblocks;
  const int t0 = tb * kCvT, c0 = blockIdx.y * kCvC;
  const int c = threadIdx.x & 63, tg = threadIdx.x >> 6;
  {                                                   // a tile: 64 rows x (16 + 16) float4, all loads issued before the g_d tile's
    const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;
    float4 v1[4], v2[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int t = t0 + r0 + 16 * p;
      v1[p] = v2[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < T) {
        const float* row = a + (static_cast<int64_t>(b) * T + t) * (2 * C) + c0 + c4;
        v1[p] = __ldg(reinterpret_cast<const float4*>(row));
        v2[p] = __ldg(reinterpret_cast<const float4*>(row + C));
      }
    }
    load_plain_tile(gd, b, T, C, t0, c0, tile);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      *reinterpret_cast<float4*>(a_tile + (r0 + 16 * p) * kCvC + c4) = v1[p];
      *reinterpret_cast<float4*>(a_tile + kCvT * kCvC + (r0 + 16 * p) * kCvC + c4) = v2[p];
    }
  }
  float wr[kCvTaps];
  load_taps(w, c0 + c, ks, true, wr);
  __syncthreads();
  float win[kCvWin];
#pragma unroll
  for (int i = 0; i < kCvWin; ++i) win[i] = tile[(tg * kCvPer + i) * kCvC + c];
#pragma unroll
  for (int o = 0; o < kCvPer; ++o) {
    const int tl = tg * kCvPer + o, t = t0 + tl;
    if (t >= T) break;
    float gg = 0.f;
#pragma unroll
    for (int j = 0; j < kCvTaps; ++j) gg = fmaf(wr[j], win[o + j], gg);
    const int64_t idx = (static_cast<int64_t>(b) * T + t) * (2 * C) + c0 + c;
    const float a1 = a_tile[tl * kCvC + c], sg = sigmoid_f(a_tile[kCvT * kCvC + tl * kCvC + c]);
    ga[idx] = gg * sg;
    ga[idx + C] = gg * a1 * sg * (1.0f - sg);
  }
}

// per-tile partials of g_w[c][j] = sum_t g_d[t] * g[t + j - 15] and g_bias[c] = sum_t g_d[t]:
// part: [B * tblocks][32][C], slot 31 = bias.  grid: x = b * tblocks + tb, y = channel tile.
__global__ void __launch_bounds__(256) dwconv_bwd_weight_kernel(const float* __restrict__ gd, const float* __restrict__ a, int B,
                                                                int T, int C, float* __restrict__ part) {
  pdl_entry();
  __shared__ __align__(16) float buf[4 * 32 * kCvC];               // the GLU tile (94 x 64) first, then the cross-group reduction
  const int tblocks = (T + kCvT - 1) / kCvT;
  const int b = blockIdx.x / tblocks, tb = blockIdx.x % tblocks;
  const int t0 = tb * kCvT, c0 = blockIdx.y * kCvC;
  const int c = threadIdx.x & 63, tg = threadIdx.x >> 6;
  float gv[kCvPer];
#pragma unroll
  for (int o = 0; o < kCvPer; ++o) {
    const int t = t0 + tg * kCvPer + o;
    gv[o] = t < T ? __ldg(gd + (static_cast<int64_t>(b) * T + t) * C + c0 + c) : 0.f;
  }
  load_glu_tile(a, b, T, C, t0, c0, buf);
  __syncthreads();
  float win[kCvWin];
#pragma unroll
  for (int i = 0; i < kCvWin; ++i) win[i] = buf[(tg * kCvPer + i) * kCvC + c];
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll
  for (int o = 0; o < kCvPer; ++o) {
#pragma unroll
    for (int j = 0; j < kCvTaps; ++j) acc[j] = fmaf(gv[o], win[o + j], acc[j]);
    acc[31] += gv[o];
  }
  __syncthreads();                                   // everyone has its window: the tile memory is free
#pragma unroll
  for (int j = 0; j < 32; ++j) buf[(tg * 32 + j) * kCvC + c] = acc[j];
  __syncthreads();
  for (int j = tg; j < 32; j += 4) {
    const float v = buf[(0 * 32 + j) * kCvC + c] + buf[(1 * 32 + j) * kCvC + c] + buf[(2 * 32 + j) * kCvC + c] +
                    buf[(3 * 32 + j) * kCvC + c];
    part[(static_cast<int64_t>(blockIdx.x) * 32 + j) * C + c0 + c] = v;
  }
}

// g_w[c][j] (ks centred taps) and g_bias[c] from the per-tile partials.  grid: 32 * C / 64 blocks of 1024 threads
__global__ void __launch_bounds__(64 * kFinGroups) dwconv_bwd_weight_finalize_kernel(const float* __restrict__ part, int nblocks,
                                                                                   int C, int ks, float* __restrict__ gw,
                                                                                   float* __restrict__ gbias) {
  pdl_entry();
  __shared__ double sh[kFinGroups][64];
  const int item = blockIdx.x * 64 + (threadIdx.x & 63);      // slot * C + c
  const double acc = grouped_block_sum(part, nblocks, 32 * static_cast<int64_t>(C), item, sh);
  if (threadIdx.x >= 64) return;
  const int slot = item / C, c = item % C;
  const int off = (kCvTaps - ks) / 2;
  if (slot == 31) {
    if (gbias != nullptr) gbias[c] = static_cast<float>(acc);
  } else if (slot >= off && slot < off + ks) {
    gw[c * ks + slot - off] = static_cast<float>(acc);
  }
}

// ---- glue of the attention core (conformer.py:113-117): q + pos_bias_u and q + pos_bias_v in one pass, and in the
// backward g_q = g_qu + g_qw together with the column sums that are the gradients of the two biases ----
__global__ void __launch_bounds__(256) add_bias2_kernel(const float* __restrict__ q, const float* __restrict__ u,
                                                        const float* __restrict__ w, int64_t n4, int W4, float* __restrict__ qu,
                                                        float* __restrict__ qw) {
  pdl_entry();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int c4 = static_cast<int>(i % W4);
    const float4 x = __ldg(reinterpret_cast<const float4*>(q) + i);
    const float4 a = __ldg(reinterpret_cast<const float4*>(u) + c4), b = __ldg(reinterpret_cast<const float4*>(w) + c4);
    reinterpret_cast<float4*>(qu)[i] = make_float4(x.x + a.x, x.y + a.y, x.z + a.z, x.w + a.w);
    reinterpret_cast<float4*>(qw)[i] = make_float4(x.x + b.x, x.y + b.y, x.z + b.z, x.w + b.w);
  }
}

// g = ga + gb; part[blk][0][c] = sum of ga, part[blk][1][c] = sum of gb over the block's rows
__global__ void __launch_bounds__(256) add_colsum2_kernel(const float* __restrict__ ga, const float* __restrict__ gb, int64_t M,
                                                          int W, float* __restrict__ g, float* __restrict__ part) {
  pdl_entry();
  __shared__ float red[2][4][kCvC];
  const int c = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int cc = blockIdx.y * kCvC + c;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kBnRows;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
  for (int r = rg; r < kBnRows; r += 4) {
    const int64_t row = r0 + r;
    if (row < M) {
      const float a = __ldg(ga + row * W + cc), b = __ldg(gb + row * W + cc);
      g[row * W + cc] = a + b;
      s1 += a;
      s2 += b;
    }
  }
  red[0][rg][c] = s1;
  red[1][rg][c] = s2;
  __syncthreads();
  if (rg < 2) part[(static_cast<int64_t>(blockIdx.x) * 2 + rg) * W + cc] = red[rg][0][c] + red[rg][1][c] + red[rg][2][c] + red[rg][3][c];
}

// ---- module tail (conformer.py:45, 133-138, 163-167): out = x + scale * dropout(y) * rowmask, one pass; the dropout bits
// come from the Philox stream (8 consecutive elements = the 8 lanes of block e / 8) and are regenerated in the backward ----
template <bool BWD>
__global__ void __launch_bounds__(256) residual_dropout_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                               const float* __restrict__ rowmask, float scale, float inv_keep,
                                                               DropRng rng, int64_t n8, int C8, float* __restrict__ out) {
  pdl_entry();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float rm = rowmask != nullptr ? __ldg(rowmask + i / C8) : 1.0f;
    const uint32_t kb = rng.threshold != 0u ? philox_keep8(static_cast<unsigned long long>(i), rng) : 0xFFu;
    const float f = scale * rm * (rng.threshold != 0u ? inv_keep : 1.0f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(y) + 2 * i + h);
      float4 o;
      o.x = (kb >> (4 * h + 0)) & 1u ? __fmul_rn(v.x, f) : 0.f;      // explicit mul / add: same bits as the GEMM's fused tail
      o.y = (kb >> (4 * h + 1)) & 1u ? __fmul_rn(v.y, f) : 0.f;
      o.z = (kb >> (4 * h + 2)) & 1u ? __fmul_rn(v.z, f) : 0.f;
      o.w = (kb >> (4 * h + 3)) & 1u ? __fmul_rn(v.w, f) : 0.f;
      if (!BWD) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + h);
        o.x = __fadd_rn(o.x, r.x), o.y = __fadd_rn(o.y, r.y), o.z = __fadd_rn(o.z, r.z), o.w = __fadd_rn(o.w, r.w);
      }
      reinterpret_cast<float4*>(out)[2 * i + h] = o;
    }
  }
}

static int conv_tblocks(int T) { return (T + kCvT - 1) / kCvT; }
static int ew_blocks(int64_t n4) {
  const int64_t want = (n4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  return static_cast<int>(want < cap ? want : cap);
}

}  // namespace ob

using namespace ob;

#define OB_CONV_SHAPE(name)                                                                                            \
  OB_REQUIRE(B > 0 && T > 0 && C > 0 && C % 64 == 0, name ": need B, T > 0 and C a positive multiple of 64 (C=%d)", C); \
  OB_REQUIRE(ks >= 1 && ks <= 31 && (ks & 1), name ": kernel size must be odd and <= 31 (ks=%d)", ks)

extern "C" size_t ob_convmod_workspace_bytes(int B, int T, int C) {
  if (B <= 0 || T <= 0 || C <= 0) return 0;
  const size_t tiles = static_cast<size_t>(B) * conv_tblocks(T);
  const size_t rows = (static_cast<size_t>(B) * T + kBnRows - 1) / kBnRows;
  const size_t a = tiles * 32 * C * sizeof(float);             // weight-gradient partials (largest user)
  const size_t b = rows * 2 * C * sizeof(float);
  return (a > b ? a : b) + 256;
}

// `groups` > 1: the batch is a stack of `groups` independent passes (B / groups utterances each, contiguous): BatchNorm
// statistics - and therefore the two-pass backward - are per pass; the stencil kernels do not care.
extern "C" int ob_glu_dwconv_bn_fwd(const float* a, const float* w, const float* bias, int B, int T, int C, int ks, float eps,
                                    int groups, float* d, float* mean, float* rstd, void* ws, ob_stream_t stream) {
  OB_REQUIRE(a && w && d && mean && rstd && ws, "ob_glu_dwconv_bn_fwd: null pointer");
  OB_CONV_SHAPE("ob_glu_dwconv_bn_fwd");
  OB_REQUIRE(groups >= 1 && B % groups == 0, "ob_glu_dwconv_bn_fwd: B (%d) must be a multiple of groups (%d)", B, groups);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(ws);
  dim3 grid(B * conv_tblocks(T), C / kCvC);
  launch_k((glu_dwconv_fwd_kernel), dim3(grid), dim3(256), 0, st, a, w, bias, B, T, C, ks, d, part);
  OB_LAUNCH_CHECK("glu_dwconv_fwd_kernel");
  const int Bg = B / groups;
  for (int g = 0; g < groups; ++g) {                     // mean, rstd: [groups][C]
    launch_k((bn_stats_finalize_kernel), dim3(C / kStatCh), dim3(kStatCh * kStatGroups), 0, st, part + static_cast<size_t>(g) * Bg * conv_tblocks(T) * 2 * C, Bg, T,
                                                               C, eps, mean + g * C, rstd + g * C);
    OB_LAUNCH_CHECK("bn_stats_finalize_kernel");
  }
  return OB_OK;
}

extern "C" int ob_bn_swish_fwd(const float* d, const float* mean, const float* rstd, const float* gamma, const float* beta,
                               int64_t M, int C, int groups, float* s, ob_stream_t stream) {
  OB_REQUIRE(d && mean && rstd && gamma && beta && s, "ob_bn_swish_fwd: null pointer");
  OB_REQUIRE(M > 0 && C > 0 && C % 64 == 0, "ob_bn_swish_fwd: need M > 0 and C a positive multiple of 64 (C=%d)", C);
  OB_REQUIRE(groups >= 1 && M % groups == 0, "ob_bn_swish_fwd: M must be a multiple of groups (%d)", groups);
  const int64_t Mg = M / groups, n4 = Mg * C / 4;
  for (int g = 0; g < groups; ++g) {
    launch_k((bn_swish_fwd_kernel), dim3(ew_blocks(n4)), dim3(256), 0, static_cast<cudaStream_t>(stream), d + g * Mg * C, mean + g * C, rstd + g * C,
                                                                                      gamma, beta, n4, C / 4, s + g * Mg * C);
    OB_LAUNCH_CHECK("bn_swish_fwd_kernel");
  }
  return OB_OK;
}

// g_gamma_beta: [groups][2][C], per group (sum g_y = g_beta, sum g_y * xhat = g_gamma); the parameter gradients are the
// sums over the groups
extern "C" int ob_bn_swish_bwd(const float* gs, const float* d, const float* mean, const float* rstd, const float* gamma,
                               const float* beta, int64_t M, int C, int groups, float* gd, float* g_gamma_beta, void* ws,
                               ob_stream_t stream) {
  OB_REQUIRE(gs && d && mean && rstd && gamma && beta && gd && g_gamma_beta && ws, "ob_bn_swish_bwd: null pointer");
  OB_REQUIRE(M > 0 && C > 0 && C % 64 == 0, "ob_bn_swish_bwd: need M > 0 and C a positive multiple of 64 (C=%d)", C);
  OB_REQUIRE(groups >= 1 && M % groups == 0, "ob_bn_swish_bwd: M must be a multiple of groups (%d)", groups);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(ws);
  const int64_t Mg = M / groups, n4 = Mg * C / 4;
  const int rblocks = static_cast<int>((Mg + kBnRows - 1) / kBnRows);
  for (int g = 0; g < groups; ++g) {
    const float *gs_g = gs + g * Mg * C, *d_g = d + g * Mg * C, *mean_g = mean + g * C, *rstd_g = rstd + g * C;
    float* sums_g = g_gamma_beta + static_cast<size_t>(g) * 2 * C;
    launch_k((bn_swish_bwd_reduce_kernel), dim3(dim3(rblocks, C / kCvC)), dim3(256), 0, st, gs_g, d_g, mean_g, rstd_g, gamma, beta, Mg, C, part);
    OB_LAUNCH_CHECK("bn_swish_bwd_reduce_kernel");
    launch_k((bn_bwd_finalize_kernel), dim3(2 * C / 64), dim3(64 * kFinGroups), 0, st, part, rblocks, C, sums_g);
    OB_LAUNCH_CHECK("bn_bwd_finalize_kernel");
    launch_k((bn_swish_bwd_apply_kernel), dim3(ew_blocks(n4)), dim3(256), 0, st, gs_g, d_g, mean_g, rstd_g, gamma, beta, sums_g,
                                                            1.0f / static_cast<float>(Mg), n4, C / 4, gd + g * Mg * C);
    OB_LAUNCH_CHECK("bn_swish_bwd_apply_kernel");
  }
  return OB_OK;
}

extern "C" int ob_glu_dwconv_bwd(const float* gd, const float* a, const float* w, int B, int T, int C, int ks, float* ga,
                                 float* gw, float* gbias, void* ws, ob_stream_t stream) {
  OB_REQUIRE(gd && a && w && ga && gw && ws, "ob_glu_dwconv_bwd: null pointer");
  OB_CONV_SHAPE("ob_glu_dwconv_bwd");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(ws);
  dim3 grid(B * conv_tblocks(T), C / kCvC);
  static bool attr_set = false;
  if (!attr_set) {
    OB_CUDA(cudaFuncSetAttribute(dwconv_bwd_data_glu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCvBwdSmemBytes));
    attr_set = true;
  }
  launch_k((dwconv_bwd_data_glu_kernel), dim3(grid), dim3(256), kCvBwdSmemBytes, st, gd, a, w, B, T, C, ks, ga);
  OB_LAUNCH_CHECK("dwconv_bwd_data_glu_kernel");
  launch_k((dwconv_bwd_weight_kernel), dim3(grid), dim3(256), 0, st, gd, a, B, T, C, part);
  OB_LAUNCH_CHECK("dwconv_bwd_weight_kernel");
  launch_k((dwconv_bwd_weight_finalize_kernel), dim3(32 * C / 64), dim3(64 * kFinGroups), 0, st, part, B * conv_tblocks(T), C, ks, gw, gbias);
  OB_LAUNCH_CHECK("dwconv_bwd_weight_finalize_kernel");
  return OB_OK;
}

extern "C" int ob_add_bias2(const float* q, const float* u, const float* w, int64_t M, int W, float* qu, float* qw,
                            ob_stream_t stream) {
  OB_REQUIRE(q && u && w && qu && qw, "ob_add_bias2: null pointer");
  OB_REQUIRE(M > 0 && W > 0 && W % 4 == 0, "ob_add_bias2: need M > 0 and W a positive multiple of 4 (W=%d)", W);
  const int64_t n4 = M * W / 4;
  launch_k((add_bias2_kernel), dim3(ew_blocks(n4)), dim3(256), 0, static_cast<cudaStream_t>(stream), q, u, w, n4, W / 4, qu, qw);
  OB_LAUNCH_CHECK("add_bias2_kernel");
  return OB_OK;
}

extern "C" size_t ob_add_colsum2_workspace_bytes(int64_t M, int W) {
  if (M <= 0 || W <= 0) return 0;
  return static_cast<size_t>((M + kBnRows - 1) / kBnRows) * 2 * W * sizeof(float) + 256;
}

extern "C" int ob_add_colsum2(const float* ga, const float* gb, int64_t M, int W, float* g, float* sums, void* ws,
                              ob_stream_t stream) {
  OB_REQUIRE(ga && gb && g && sums && ws, "ob_add_colsum2: null pointer");
  OB_REQUIRE(M > 0 && W > 0 && W % 64 == 0, "ob_add_colsum2: need M > 0 and W a positive multiple of 64 (W=%d)", W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(ws);
  const int rblocks = static_cast<int>((M + kBnRows - 1) / kBnRows);
  launch_k((add_colsum2_kernel), dim3(dim3(rblocks, W / kCvC)), dim3(256), 0, st, ga, gb, M, W, g, part);
  OB_LAUNCH_CHECK("add_colsum2_kernel");
  launch_k((bn_bwd_finalize_kernel), dim3(2 * W / 64), dim3(64 * kFinGroups), 0, st, part, rblocks, W, sums);   // sums[0] = colsum(ga), [1] = colsum(gb)
  OB_LAUNCH_CHECK("bn_bwd_finalize_kernel");
  return OB_OK;
}

extern "C" int ob_residual_dropout_fwd(const float* x, const float* y, const float* rowmask, float scale, float inv_keep,
                                       uint64_t seed, uint64_t offset, uint32_t drop_threshold, int64_t M, int C, float* out,
                                       ob_stream_t stream) {
  OB_REQUIRE(x && y && out, "ob_residual_dropout_fwd: null pointer");
  OB_REQUIRE(M > 0 && C > 0 && C % 8 == 0, "ob_residual_dropout_fwd: need M > 0 and C a positive multiple of 8 (C=%d)", C);
  OB_REQUIRE(drop_threshold < 65536u, "ob_residual_dropout_fwd: drop_threshold (%u) is a 16-bit value", drop_threshold);
  const DropRng rng = {seed, offset, drop_threshold};
  const int64_t n8 = M * C / 8;
  launch_k((residual_dropout_kernel<false>), dim3(ew_blocks(n8)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, y, rowmask, scale, inv_keep, rng,
                                                                                                n8, C / 8, out);
  OB_LAUNCH_CHECK("residual_dropout_kernel");
  return OB_OK;
}

extern "C" int ob_residual_dropout_bwd(const float* g, const float* rowmask, float scale, float inv_keep, uint64_t seed,
                                       uint64_t offset, uint32_t drop_threshold, int64_t M, int C, float* gy,
                                       ob_stream_t stream) {
  OB_REQUIRE(g && gy, "ob_residual_dropout_bwd: null pointer");
  OB_REQUIRE(M > 0 && C > 0 && C % 8 == 0, "ob_residual_dropout_bwd: need M > 0 and C a positive multiple of 8 (C=%d)", C);
  OB_REQUIRE(drop_threshold < 65536u, "ob_residual_dropout_bwd: drop_threshold (%u) is a 16-bit value", drop_threshold);
  const DropRng rng = {seed, offset, drop_threshold};
  const int64_t n8 = M * C / 8;
  launch_k((residual_dropout_kernel<true>), dim3(ew_blocks(n8)), dim3(256), 0, static_cast<cudaStream_t>(stream), nullptr, g, rowmask, scale, inv_keep,
                                                                                               rng, n8, C / 8, gy);
  OB_LAUNCH_CHECK("residual_dropout_kernel");
  return OB_OK;
}
