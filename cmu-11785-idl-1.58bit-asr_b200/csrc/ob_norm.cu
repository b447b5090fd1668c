// ob_norm.cu - LayerNorm in front of the routed projections (conformer.py:35, 109).
// One warp per token row held in registers (C = 128 V), fp32 throughout; the backward fuses dx with per-block
// partial sums of d-gamma / d-beta, reduced in fixed order (deterministic).  HBM-bound streaming kernels.
#include "ob_common.cuh"

namespace ob {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int V>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float eps, int64_t M, int C,
                                                     float* __restrict__ y, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[V], b[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    g[j] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
    b[j] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  for (int64_t row = warp0; row < M; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = __ldg(xr + lane + 32 * j);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = wsum(s) * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
      ss += (a * a + c * c) + (d * d + e * e);
    }
    const float rstd = 1.0f / sqrtf(wsum(ss) * inv_c + eps);
    float4* yr = reinterpret_cast<float4*>(y + row * C);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 o;
      o.x = (v[j].x - mean) * rstd * g[j].x + b[j].x;
      o.y = (v[j].y - mean) * rstd * g[j].y + b[j].y;
      o.z = (v[j].z - mean) * rstd * g[j].z + b[j].z;
      o.w = (v[j].w - mean) * rstd * g[j].w + b[j].w;
      yr[lane + 32 * j] = o;
    }
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

// LayerNorm fused with the per-token absmax int8 quantiser of the routed projection(s) that consume it (conformer.py:35-36,
// 109-112): the normalised row never leaves the registers - int8 codes, the scale and the
// (mean, rstd) the backward needs are all that is written (5 bytes per element instead of 13 for LayerNorm + quantiser).
// Same expressions as ln_fwd_kernel and act_quant_reg_kernel, so the codes equal act_quant(layer_norm(x)) bit for bit.
template <int V>
__global__ void __launch_bounds__(256) ln_quant_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, int64_t M, int C,
                                                           int8_t* __restrict__ q, float* __restrict__ scale,
                                                           float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[V], b[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    g[j] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
    b[j] = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * j);
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  // the next row of this warp is loaded while the current one goes through its three dependent warp reductions
  constexpr bool kPre = V <= 4;                         // up to 512 columns; wider rows have enough loads in flight per lane
  float4 nx[V];
  if (kPre && warp0 < M) {
#pragma unroll
    for (int j = 0; j < V; ++j) nx[j] = __ldg(reinterpret_cast<const float4*>(x + warp0 * C) + lane + 32 * j);
  }
  for (int64_t row = warp0; row < M; row += nwarps) {
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = kPre ? nx[j] : __ldg(reinterpret_cast<const float4*>(x + row * C) + lane + 32 * j);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    if (kPre && row + nwarps < M) {
      const float4* xn = reinterpret_cast<const float4*>(x + (row + nwarps) * C);
#pragma unroll
      for (int j = 0; j < V; ++j) nx[j] = __ldg(xn + lane + 32 * j);
    }
    const float mean = wsum(s) * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
      ss += (a * a + c * c) + (d * d + e * e);
    }
    const float rstd = 1.0f / sqrtf(wsum(ss) * inv_c + eps);
    uint32_t amax = 0u;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 o;
      o.x = (v[j].x - mean) * rstd * g[j].x + b[j].x;
      o.y = (v[j].y - mean) * rstd * g[j].y + b[j].y;
      o.z = (v[j].z - mean) * rstd * g[j].z + b[j].z;
      o.w = (v[j].w - mean) * rstd * g[j].w + b[j].w;
      v[j] = o;
      amax = amax_bits4(amax, o);
    }
    const float sc = act_scale_from_amax(__uint_as_float(warp_max_bits(amax)));
    uint32_t* qr = reinterpret_cast<uint32_t*>(q + row * C);
#pragma unroll
    for (int j = 0; j < V; ++j) qr[lane + 32 * j] = quant4(v[j], sc);
    if (lane == 0) {
      scale[row] = sc;
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

constexpr int kLnBwdBlocks = 296;      // persistent grid (2 per SM); fixed so the parameter-gradient sum order is fixed

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) [+ resid],  g = dy * gamma;  partial d-gamma += dy * xhat, d-beta += dy.
// dy = dy0 (+ dy1 + dy2): the gradients of up to three projections that read the same normalised tensor (q, k, v) are
// summed on load; `resid` is the gradient that reached the module's input along the residual path, added on store - both
// spare a separate element-wise kernel over [M, C].
template <int V, int NDY, int RESID>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dy1,
                                                     const float* __restrict__ dy2, const float* __restrict__ x,
                                                     const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                     const float* __restrict__ gamma, const float* __restrict__ resid, int64_t M,
                                                     int C, float* __restrict__ dx, float* __restrict__ part /* [2][blocks][C] */) {
  pdl_entry();
  __shared__ float4 red[8][32 * V];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g[V], dg[V], db[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    g[j] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * j);
    dg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float inv_c = 1.0f / static_cast<float>(C);
  for (int64_t row = warp0; row < M; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    const float4* dr = reinterpret_cast<const float4*>(dy + row * C);
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    float4 xh[V], gg[V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float4 xv = __ldg(xr + lane + 32 * j);
      float4 dv = __ldg(dr + lane + 32 * j);
      if (NDY >= 2) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(dy1 + row * C) + lane + 32 * j);
        dv.x += t.x; dv.y += t.y; dv.z += t.z; dv.w += t.w;
      }
      if (NDY >= 3) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(dy2 + row * C) + lane + 32 * j);
        dv.x += t.x; dv.y += t.y; dv.z += t.z; dv.w += t.w;
      }
      xh[j] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      gg[j] = make_float4(dv.x * g[j].x, dv.y * g[j].y, dv.z * g[j].z, dv.w * g[j].w);
      s1 += (gg[j].x + gg[j].y) + (gg[j].z + gg[j].w);
      s2 += (gg[j].x * xh[j].x + gg[j].y * xh[j].y) + (gg[j].z * xh[j].z + gg[j].w * xh[j].w);
      dg[j].x += dv.x * xh[j].x; dg[j].y += dv.y * xh[j].y; dg[j].z += dv.z * xh[j].z; dg[j].w += dv.w * xh[j].w;
      db[j].x += dv.x; db[j].y += dv.y; db[j].z += dv.z; db[j].w += dv.w;
    }
    const float m1 = wsum(s1) * inv_c, m2 = wsum(s2) * inv_c;
    float4* oxr = reinterpret_cast<float4*>(dx + row * C);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 o;
      o.x = rstd * (gg[j].x - m1 - xh[j].x * m2);
      o.y = rstd * (gg[j].y - m1 - xh[j].y * m2);
      o.z = rstd * (gg[j].z - m1 - xh[j].z * m2);
      o.w = rstd * (gg[j].w - m1 - xh[j].w * m2);
      if (RESID) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(resid + row * C) + lane + 32 * j);
        o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
      }
      oxr[lane + 32 * j] = o;
    }
  }
  // block partials: warps combined in fixed order
  float* part_g = part + (int64_t)blockIdx.x * C;
  float* part_b = part + ((int64_t)gridDim.x + blockIdx.x) * C;
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int j = 0; j < V; ++j) red[wid][lane + 32 * j] = which == 0 ? dg[j] : db[j];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * V; i += 256) {
      float4 a = red[0][i];
#pragma unroll
      for (int w = 1; w < 8; ++w) {
        const float4 o = red[w][i];
        a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
      }
      reinterpret_cast<float4*>(which == 0 ? part_g : part_b)[i] = a;
    }
    __syncthreads();
  }
}

// [2][nblocks][C] partials -> d-gamma, d-beta: 32 columns x 8 row groups per block, fixed order
__global__ void __launch_bounds__(256) ln_param_grad_kernel(const float* __restrict__ part, int nblocks, int C,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_entry();
  __shared__ float red[8][33];
  const int which = blockIdx.y;
  const float* p = part + (int64_t)which * nblocks * C;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  // eight loads in flight per thread (the kernel is a chain of L2 round trips: with two it took 9.3 us for 1184 partial rows)
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < C) {
    int b = ry;
    for (; b + 56 < nblocks; b += 64) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += __ldg(p + (int64_t)(b + 8 * u) * C + c);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (b + 8 * u < nblocks) a[u] += __ldg(p + (int64_t)(b + 8 * u) * C + c);
  }
  red[ry][cx] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][cx];
    (which == 0 ? dgamma : dbeta)[c] = t;
  }
}

// Column sums of a row-major [M, N] fp32 matrix (bias gradients of the non-routed linears: dY.sum(0)), deterministic:
// a fixed grid of blocks each reduces a contiguous range of rows into one partial row, ln_param_grad_kernel adds the
// partial rows in fixed order.  torch's column reduction reaches ~0.8 TB/s on [76608, 256]; this streams at HBM speed.
constexpr int kColsumBlocks = 592;

__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, int64_t M, int N, int rows_per_block,
                                                             float* __restrict__ part /* [blocks][N] */) {
  pdl_entry();
  __shared__ float4 red[256];
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  for (int c0 = 0; c0 < N; c0 += 1024) {
    const int cw = min(1024, N - c0);
    const int tpr = cw >> 2;                       // threads per row (one float4 each)
    const int rpi = 256 / tpr;                     // rows per pass
    const int rg = threadIdx.x / tpr;
    const int c = c0 + (threadIdx.x - rg * tpr) * 4;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    if (rg < rpi) {
      int64_t r = r0 + rg;
      for (; r + rpi < r1; r += 2 * rpi) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + r * N + c));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (r + rpi) * N + c));
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
      }
      if (r < r1) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + r * N + c));
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      }
      a0.x += a1.x; a0.y += a1.y; a0.z += a1.z; a0.w += a1.w;
    }
    red[threadIdx.x] = a0;
    __syncthreads();
    if (rg == 0) {
      for (int g = 1; g < rpi; ++g) {
        const float4 o = red[g * tpr + threadIdx.x];
        a0.x += o.x; a0.y += o.y; a0.z += o.z; a0.w += o.w;
      }
      *reinterpret_cast<float4*>(part + (int64_t)blockIdx.x * N + c) = a0;
    }
    __syncthreads();
  }
}

}  // namespace ob

using namespace ob;

static int ln_blocks(int64_t M) {
  const int64_t want = (M + 7) / 8;
  return (int)(want < 148 * 8 ? want : 148 * 8);
}

extern "C" int ob_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, int64_t M, int C,
                                float* y, float* mean, float* rstd, ob_stream_t stream) {
  OB_REQUIRE(x && gamma && beta && y && mean && rstd && M > 0, "ob_layernorm_fwd: null pointer or M <= 0");
  OB_REQUIRE(C == 128 || C == 256 || C == 512 || C == 1024, "ob_layernorm_fwd: C (%d) must be 128, 256, 512 or 1024", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = ln_blocks(M);
  switch (C) {
    case 128:  launch_k((ln_fwd_kernel<1>), dim3(blocks), dim3(256), 0, st, x, gamma, beta, eps, M, C, y, mean, rstd); break;
    case 256:  launch_k((ln_fwd_kernel<2>), dim3(blocks), dim3(256), 0, st, x, gamma, beta, eps, M, C, y, mean, rstd); break;
    case 512:  launch_k((ln_fwd_kernel<4>), dim3(blocks), dim3(256), 0, st, x, gamma, beta, eps, M, C, y, mean, rstd); break;
    default:   launch_k((ln_fwd_kernel<8>), dim3(blocks), dim3(256), 0, st, x, gamma, beta, eps, M, C, y, mean, rstd); break;
  }
  OB_LAUNCH_CHECK("ln_fwd_kernel");
  return OB_OK;
}

extern "C" size_t ob_layernorm_bwd_workspace_bytes(int C) { return (size_t)2 * kLnBwdBlocks * C * sizeof(float); }

static int launch_ln_bwd(const float* dy, const float* dy1, const float* dy2, const float* x, const float* mean, const float* rstd,
                         const float* gamma, const float* resid, int64_t M, int C, float* dx, float* dgamma, float* dbeta, void* ws,
                         cudaStream_t st) {
  const int64_t want = (M + 7) / 8;
  const int blocks = (int)(want < kLnBwdBlocks ? want : kLnBwdBlocks);
  float* part = static_cast<float*>(ws);
  const int ndy = dy2 != nullptr ? 3 : (dy1 != nullptr ? 2 : 1);
#define OB_LN_BWD(V)                                                                                                          \
  do {                                                                                                                        \
    if (resid != nullptr) {                                                                                                   \
      if (ndy == 1) launch_k((ln_bwd_kernel<V, 1, 1>), dim3(blocks), dim3(256), 0, st, dy, dy1, dy2, x, mean, rstd, gamma, resid, M, C, dx, part); \
      else if (ndy == 2) launch_k((ln_bwd_kernel<V, 2, 1>), dim3(blocks), dim3(256), 0, st, dy, dy1, dy2, x, mean, rstd, gamma, resid, M, C, dx, part); \
      else launch_k((ln_bwd_kernel<V, 3, 1>), dim3(blocks), dim3(256), 0, st, dy, dy1, dy2, x, mean, rstd, gamma, resid, M, C, dx, part);          \
    } else {                                                                                                                  \
      if (ndy == 1) launch_k((ln_bwd_kernel<V, 1, 0>), dim3(blocks), dim3(256), 0, st, dy, dy1, dy2, x, mean, rstd, gamma, resid, M, C, dx, part); \
      else if (ndy == 2) launch_k((ln_bwd_kernel<V, 2, 0>), dim3(blocks), dim3(256), 0, st, dy, dy1, dy2, x, mean, rstd, gamma, resid, M, C, dx, part); \
      else launch_k((ln_bwd_kernel<V, 3, 0>), dim3(blocks), dim3(256), 0, st, dy, dy1, dy2, x, mean, rstd, gamma, resid, M, C, dx, part);          \
    }                                                                                                                         \
  } while (0)
  switch (C) {
    case 128:  OB_LN_BWD(1); break;
    case 256:  OB_LN_BWD(2); break;
    case 512:  OB_LN_BWD(4); break;
    default:   OB_LN_BWD(8); break;
  }
#undef OB_LN_BWD
  OB_LAUNCH_CHECK("ln_bwd_kernel");
  dim3 grid((C + 31) / 32, 2);
  launch_k((ln_param_grad_kernel), dim3(grid), dim3(256), 0, st, part, blocks, C, dgamma, dbeta);
  OB_LAUNCH_CHECK("ln_param_grad_kernel");
  return OB_OK;
}

extern "C" int ob_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                int64_t M, int C, float* dx, float* dgamma, float* dbeta, void* ws, ob_stream_t stream) {
  OB_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && ws && M > 0,
             "ob_layernorm_bwd: null pointer or M <= 0");
  OB_REQUIRE(C == 128 || C == 256 || C == 512 || C == 1024, "ob_layernorm_bwd: C (%d) must be 128, 256, 512 or 1024", C);
  return launch_ln_bwd(dy, nullptr, nullptr, x, mean, rstd, gamma, nullptr, M, C, dx, dgamma, dbeta, ws,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int ob_layernorm_bwd3(const float* dy0, const float* dy1, const float* dy2, const float* x, const float* mean,
                                 const float* rstd, const float* gamma, const float* resid, int64_t M, int C, float* dx,
                                 float* dgamma, float* dbeta, void* ws, ob_stream_t stream) {
  OB_REQUIRE(dy0 && x && mean && rstd && gamma && dx && dgamma && dbeta && ws && M > 0,
             "ob_layernorm_bwd3: null pointer or M <= 0");
  OB_REQUIRE(dy1 != nullptr || dy2 == nullptr, "ob_layernorm_bwd3: dy2 given without dy1");
  OB_REQUIRE(C == 128 || C == 256 || C == 512 || C == 1024, "ob_layernorm_bwd3: C (%d) must be 128, 256, 512 or 1024", C);
  return launch_ln_bwd(dy0, dy1, dy2, x, mean, rstd, gamma, resid, M, C, dx, dgamma, dbeta, ws, static_cast<cudaStream_t>(stream));
}

extern "C" int ob_layernorm_quant_fwd(const float* x, const float* gamma, const float* beta, float eps, int64_t M, int C,
                                      int8_t* q, float* scale, float* mean, float* rstd, ob_stream_t stream) {
  OB_REQUIRE(x && gamma && beta && q && scale && mean && rstd && M > 0, "ob_layernorm_quant_fwd: null pointer or M <= 0");
  OB_REQUIRE(C == 128 || C == 256 || C == 512 || C == 1024, "ob_layernorm_quant_fwd: C (%d) must be 128, 256, 512 or 1024", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = ln_blocks(M);
  switch (C) {
    case 128:  launch_k((ln_quant_fwd_kernel<1>), dim3(blocks), dim3(256), 0, st, x, gamma, beta, eps, M, C, q, scale, mean, rstd); break;
    case 256:  launch_k((ln_quant_fwd_kernel<2>), dim3(blocks), dim3(256), 0, st, x, gamma, beta, eps, M, C, q, scale, mean, rstd); break;
    case 512:  launch_k((ln_quant_fwd_kernel<4>), dim3(blocks), dim3(256), 0, st, x, gamma, beta, eps, M, C, q, scale, mean, rstd); break;
    default:   launch_k((ln_quant_fwd_kernel<8>), dim3(blocks), dim3(256), 0, st, x, gamma, beta, eps, M, C, q, scale, mean, rstd); break;
  }
  OB_LAUNCH_CHECK("ln_quant_fwd_kernel");
  return OB_OK;
}

static int colsum_blocks(int64_t M) {
  const int64_t want = (M + 31) / 32;
  return (int)(want < kColsumBlocks ? want : kColsumBlocks);
}

extern "C" size_t ob_colsum_workspace_bytes(int64_t M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return (size_t)colsum_blocks(M) * N * sizeof(float);
}

extern "C" int ob_colsum(const float* x, int64_t M, int N, float* out, void* ws, ob_stream_t stream) {
  OB_REQUIRE(x && out && ws && M > 0 && N > 0 && N % 4 == 0, "ob_colsum: null pointer, M <= 0 or N (%d) not a multiple of 4", N);
  OB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "ob_colsum: x must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = colsum_blocks(M);
  const int rows_per_block = (int)((M + blocks - 1) / blocks);
  float* part = static_cast<float*>(ws);
  launch_k((colsum_partial_kernel), dim3(blocks), dim3(256), 0, st, x, M, N, rows_per_block, part);
  OB_LAUNCH_CHECK("colsum_partial_kernel");
  const int used = (int)((M + rows_per_block - 1) / rows_per_block);          // blocks that own at least one row
  launch_k((ln_param_grad_kernel), dim3(dim3((N + 31) / 32, 1)), dim3(256), 0, st, part, used, N, out, out);
  OB_LAUNCH_CHECK("ln_param_grad_kernel(colsum)");
  return OB_OK;
}
