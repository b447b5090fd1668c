// ob_frontend.cu - first layer of the reference's subsampling front-end (conformer.py:177-181): Conv2d(1, C, 3, stride 2)
// + bias + ReLU over the [B, T, F] feature map, forward and backward, in the channels-last layout cuDNN's tensor-core
// kernels want for the second convolution.
//
// With one input channel the layer is a 9-tap stencil that expands every output position into C channels: 2 GB of
// output for the training batch (64 x 799 x 39 x 256 fp32) against 20 MB of input, i.e. purely write-bound.  torch runs it
// as an implicit-GEMM convolution + a separate bias add + a separate ReLU (three passes over the 2 GB), and in the backward
// as ReLU-backward + bias reduction + a weight-gradient convolution.  Here: one pass each way; the backward recomputes the
// ReLU mask from the 9 inputs instead of reading the activations.
//   forward : a warp owns an output position (1 KB contiguous in NHWC), a lane 8 channels (two 128-bit stores)
//   backward: same mapping, 10 x 8 accumulators per lane (9 taps + bias), fixed-order block and grid reduction
#include "ob_common.cuh"

namespace ob {

constexpr int kC1Lane = 8;                       // channels per lane (C = 256)
constexpr int kC1Blocks = 148 * 4;

__device__ __forceinline__ void conv1_load_params(const float* __restrict__ w, const float* __restrict__ bias, int lane,
                                                  float (&wr)[9][kC1Lane], float (&br)[kC1Lane]) {
#pragma unroll
  for (int k = 0; k < kC1Lane; ++k) {
    const int c = (k < 4 ? 0 : 128) + lane * 4 + (k & 3);         // two coalesced 128-bit chunks per lane
    br[k] = bias != nullptr ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][k] = __ldg(w + c * 9 + t);
  }
}

__device__ __forceinline__ void conv1_load_inputs(const float* __restrict__ x, int64_t pos, int T, int F, int T1, int F1,
                                                  float (&v)[9]) {
  const int64_t per = static_cast<int64_t>(T1) * F1;
  const int b = static_cast<int>(pos / per);
  const int r = static_cast<int>(pos - b * per);
  const int i = r / F1, j = r - i * F1;
  const float* xp = x + (static_cast<int64_t>(b) * T + 2 * i) * F + 2 * j;
#pragma unroll
  for (int di = 0; di < 3; ++di)
#pragma unroll
    for (int dj = 0; dj < 3; ++dj) v[di * 3 + dj] = __ldg(xp + di * F + dj);   // same address in every lane: broadcast
}

__global__ void __launch_bounds__(256) conv1_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, int B, int T, int F, int T1, int F1,
                                                             float* __restrict__ y) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  float wr[9][kC1Lane], br[kC1Lane];
  conv1_load_params(w, bias, lane, wr, br);
  const int64_t total = static_cast<int64_t>(B) * T1 * F1;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t pos = warp0; pos < total; pos += nwarps) {
    float v[9];
    conv1_load_inputs(x, pos, T, F, T1, F1, v);
    float o[kC1Lane];
#pragma unroll
    for (int k = 0; k < kC1Lane; ++k) {
      float acc = br[k];
#pragma unroll
      for (int t = 0; t < 9; ++t) acc = fmaf(wr[t][k], v[t], acc);
      o[k] = fmaxf(acc, 0.f);
    }
    float* yp = y + pos * 256 + lane * 4;
    *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(yp + 128) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// part: [gridDim.x][10][256]: rows 0..8 = sum over positions of g' * x_tap, row 9 = sum of g'  (g' = g where pre > 0)
__global__ void __launch_bounds__(256) conv1_relu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                             const float* __restrict__ w, const float* __restrict__ bias, int B,
                                                             int T, int F, int T1, int F1, float* __restrict__ part) {
  pdl_entry();
  __shared__ float red[10][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float wr[9][kC1Lane], br[kC1Lane];
  conv1_load_params(w, bias, lane, wr, br);
  float acc[10][kC1Lane];
#pragma unroll
  for (int t = 0; t < 10; ++t)
#pragma unroll
    for (int k = 0; k < kC1Lane; ++k) acc[t][k] = 0.f;
  const int64_t total = static_cast<int64_t>(B) * T1 * F1;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  // two positions per iteration: four 128-bit loads of g in flight per lane (the kernel streams 2 GB at one CTA per SM)
  for (int64_t pos = warp0; pos < total; pos += 2 * nwarps) {
    const int64_t pos2 = pos + nwarps;
    const bool has2 = pos2 < total;
    const float* gp = g + pos * 256 + lane * 4;
    const float* gq = g + (has2 ? pos2 : pos) * 256 + lane * 4;
    const float4 ga0 = __ldg(reinterpret_cast<const float4*>(gp)), ga1 = __ldg(reinterpret_cast<const float4*>(gp + 128));
    const float4 gb0 = __ldg(reinterpret_cast<const float4*>(gq)), gb1 = __ldg(reinterpret_cast<const float4*>(gq + 128));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !has2) break;
      float v[9];
      conv1_load_inputs(x, h == 0 ? pos : pos2, T, F, T1, F1, v);
      const float4 g0 = h == 0 ? ga0 : gb0, g1 = h == 0 ? ga1 : gb1;
      const float gs[kC1Lane] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int k = 0; k < kC1Lane; ++k) {
        float pre = br[k];
#pragma unroll
        for (int t = 0; t < 9; ++t) pre = fmaf(wr[t][k], v[t], pre);
        const float m = pre > 0.f ? gs[k] : 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[t][k] = fmaf(m, v[t], acc[t][k]);
        acc[9][k] += m;
      }
    }
  }
  // fixed-order fold over the 8 warps of the block
  for (int wsel = 0; wsel < 8; ++wsel) {
    if (warp == wsel) {
#pragma unroll
      for (int t = 0; t < 10; ++t)
#pragma unroll
        for (int k = 0; k < kC1Lane; ++k) {
          const int c = (k < 4 ? 0 : 128) + lane * 4 + (k & 3);
          red[t][c] = (wsel == 0 ? 0.f : red[t][c]) + acc[t][k];
        }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 10 * 256; i += 256) part[static_cast<int64_t>(blockIdx.x) * 2560 + i] = red[i / 256][i % 256];
}

// gw[c][tap], gb[c] = sums of the block partials, fp64, fixed order.  grid: 2560 / 64 blocks of 1024 threads
__global__ void __launch_bounds__(1024) conv1_bwd_finalize_kernel(const float* __restrict__ part, int nblocks, float* __restrict__ gw,
                                                                  float* __restrict__ gb) {
  pdl_entry();
  __shared__ double sh[16][64];
  const int cl = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int item = blockIdx.x * 64 + cl;                      // tap * 256 + c
  double acc = 0.0;
#pragma unroll 4
  for (int blk = grp; blk < nblocks; blk += 16) acc += part[static_cast<int64_t>(blk) * 2560 + item];
  sh[grp][cl] = acc;
  __syncthreads();
  if (grp == 0) {
    for (int gi = 1; gi < 16; ++gi) acc += sh[gi][cl];
    const int tap = item / 256, c = item % 256;
    if (tap == 9) {
      if (gb != nullptr) gb[c] = static_cast<float>(acc);
    } else {
      gw[c * 9 + tap] = static_cast<float>(acc);
    }
  }
}

}  // namespace ob

using namespace ob;

#define OB_CONV1_SHAPE(name)                                                                                          \
  OB_REQUIRE(B > 0 && T >= 3 && F >= 3, name ": need B > 0 and a feature map of at least 3 x 3 (T=%d, F=%d)", T, F);   \
  OB_REQUIRE(C == 256, name ": the kernel is specialised for C = 256 output channels (C=%d)", C)

extern "C" size_t ob_conv1_relu_workspace_bytes(void) { return static_cast<size_t>(kC1Blocks) * 2560 * sizeof(float); }

extern "C" int ob_conv1_relu_fwd(const float* x, const float* w, const float* bias, int B, int T, int F, int C, float* y,
                                 ob_stream_t stream) {
  OB_REQUIRE(x && w && y, "ob_conv1_relu_fwd: null pointer");
  OB_CONV1_SHAPE("ob_conv1_relu_fwd");
  const int T1 = (T - 3) / 2 + 1, F1 = (F - 3) / 2 + 1;
  const int64_t warps = static_cast<int64_t>(B) * T1 * F1;
  const int64_t want = (warps + 7) / 8;
  const int blocks = static_cast<int>(want < kC1Blocks ? want : kC1Blocks);
  launch_k((conv1_relu_fwd_kernel), dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), x, w, bias, B, T, F, T1, F1, y);
  OB_LAUNCH_CHECK("conv1_relu_fwd_kernel");
  return OB_OK;
}

extern "C" int ob_conv1_relu_bwd(const float* g, const float* x, const float* w, const float* bias, int B, int T, int F, int C,
                                 float* gw, float* gb, void* ws, ob_stream_t stream) {
  OB_REQUIRE(g && x && w && gw && ws, "ob_conv1_relu_bwd: null pointer");
  OB_CONV1_SHAPE("ob_conv1_relu_bwd");
  const int T1 = (T - 3) / 2 + 1, F1 = (F - 3) / 2 + 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(ws);
  const int64_t warps = static_cast<int64_t>(B) * T1 * F1;
  const int64_t want = (warps + 7) / 8;
  const int blocks = static_cast<int>(want < kC1Blocks ? want : kC1Blocks);
  launch_k((conv1_relu_bwd_kernel), dim3(blocks), dim3(256), 0, st, g, x, w, bias, B, T, F, T1, F1, part);
  OB_LAUNCH_CHECK("conv1_relu_bwd_kernel");
  launch_k((conv1_bwd_finalize_kernel), dim3(2560 / 64), dim3(1024), 0, st, part, blocks, gw, gb);
  OB_LAUNCH_CHECK("conv1_bwd_finalize_kernel");
  return OB_OK;
}
