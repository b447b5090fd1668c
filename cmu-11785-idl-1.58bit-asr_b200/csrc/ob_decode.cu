// ob_decode.cu - greedy CTC decoding on the device (BASELINE configs[4]).
// Semantics of the reference's ctc_greedy_decode (onebit_asr/metrics.py:51-60): argmax over the vocabulary per frame
// (first maximal index on ties, like torch.argmax), then drop blanks and collapse repeats.  Two streaming kernels:
// one warp per frame for the argmax (HBM-bound: the logits are read exactly once with 128-bit loads), one block per
// utterance for the order-preserving compaction (ballot + prefix scan).
#include "ob_common.cuh"

namespace ob {

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

__device__ __forceinline__ void take(float v, int i, float& best, int& bi) {
  if (v > best || (v == best && i < bi)) { best = v; bi = i; }
}

template <typename T>
__global__ void __launch_bounds__(256) frame_argmax_kernel(const T* __restrict__ logits, int B, int Tn, int V,
                                                           const int32_t* __restrict__ lens, int32_t* __restrict__ pred) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t frames = (int64_t)B * Tn;
  for (int64_t f = warp0; f < frames; f += nwarps) {
    const int b = (int)(f / Tn), t = (int)(f % Tn);
    if (t >= lens[b]) {                                    // padded frame: nothing to read
      if (lane == 0) pred[f] = -1;
      continue;
    }
    const T* row = logits + f * V;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    if (sizeof(T) == 4 && (V & 3) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
      const float4* r4 = reinterpret_cast<const float4*>(row);
      for (int i = lane; i < V / 4; i += 32) {
        const float4 v = __ldg(r4 + i);
        take(v.x, 4 * i, best, bi);
        take(v.y, 4 * i + 1, best, bi);
        take(v.z, 4 * i + 2, best, bi);
        take(v.w, 4 * i + 3, best, bi);
      }
    } else {
      for (int i = lane; i < V; i += 32) take(to_f<T>(row[i]), i, best, bi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      take(ov, oi, best, bi);
    }
    if (lane == 0) pred[f] = bi == 0x7fffffff ? 0 : bi;
  }
}

// one block per utterance: keep[t] = pred[t] != blank && pred[t] != pred[t-1]; compact in order
__global__ void __launch_bounds__(256) ctc_collapse_kernel(const int32_t* __restrict__ pred, int Tn,
                                                           const int32_t* __restrict__ lens, int blank_id,
                                                           int32_t* __restrict__ out_tokens, int32_t* __restrict__ out_lens) {
  pdl_entry();
  __shared__ int warp_cnt[8];
  __shared__ int base_s;
  const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int32_t* p = pred + (int64_t)b * Tn;
  int32_t* out = out_tokens + (int64_t)b * Tn;
  const int len = min(lens[b], Tn);
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < len; t0 += 256) {
    const int t = t0 + threadIdx.x;
    int tok = -1;
    bool keep = false;
    if (t < len) {
      tok = p[t];
      const int prev = t > 0 ? p[t - 1] : -2;
      keep = tok != blank_id && tok != prev;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < wid; ++w) off += warp_cnt[w];
    off += __popc(m & ((1u << lane) - 1u));
    if (keep) out[off] = tok;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
      base_s += tot;
    }
    __syncthreads();
  }
  const int n = base_s;
  for (int t = n + threadIdx.x; t < Tn; t += 256) out[t] = -1;
  if (threadIdx.x == 0) out_lens[b] = n;
}

}  // namespace ob

using namespace ob;

extern "C" size_t ob_ctc_decode_workspace_bytes(int B, int T) {
  return B > 0 && T > 0 ? (size_t)B * T * sizeof(int32_t) : 0;
}

extern "C" int ob_ctc_greedy_decode(const void* logits, int dtype, int B, int T, int V, const int32_t* lens, int blank_id,
                                    int32_t* out_tokens, int32_t* out_lens, void* ws, ob_stream_t stream) {
  OB_REQUIRE(logits && lens && out_tokens && out_lens && ws, "ob_ctc_greedy_decode: null pointer");
  OB_REQUIRE(B > 0 && T > 0 && V > 0, "ob_ctc_greedy_decode: bad shape B=%d T=%d V=%d", B, T, V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* pred = static_cast<int32_t*>(ws);
  const int64_t frames = (int64_t)B * T;
  const int64_t want = (frames + 7) / 8;
  const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
  if (dtype == OB_F32)
    launch_k((frame_argmax_kernel<float>), dim3(blocks), dim3(256), 0, st, static_cast<const float*>(logits), B, T, V, lens, pred);
  else if (dtype == OB_BF16)
    launch_k((frame_argmax_kernel<__nv_bfloat16>), dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(logits), B, T, V, lens, pred);
  else
    OB_REQUIRE(false, "ob_ctc_greedy_decode: unknown dtype tag %d", dtype);
  OB_LAUNCH_CHECK("frame_argmax_kernel");
  launch_k((ctc_collapse_kernel), dim3(B), dim3(256), 0, st, pred, T, lens, blank_id, out_tokens, out_lens);
  OB_LAUNCH_CHECK("ctc_collapse_kernel");
  return OB_OK;
}
