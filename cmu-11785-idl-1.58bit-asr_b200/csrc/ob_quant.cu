// ob_quant.cu - the HBM-bound kernels of the quantised linear layer:
//   weight absmean, weight quantise + 2-bit pack (+ transposed copy), dense quantise, STE backward,
//   code unpack, per-token absmax int8 activation quantiser, backward prep (bf16 casts + column sums).
// All of them are streaming kernels: 128-bit loads/stores, warp-shuffle reductions, grids sized from
// the SM count.  Reference semantics: onebit_asr/quant.py (line numbers on each kernel).
#include "ob_common.cuh"

namespace ob {

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Deterministic block sum (fixed tree); result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* smem /* THREADS/32 floats */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    r = lane < THREADS / 32 ? smem[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// mean |W|   (alpha initialisation, quant.py:111-113)
// ---------------------------------------------------------------------------------------------
constexpr int kAbsmeanBlocks = 296;   // 2 per SM on a 148-SM part; fixed so the sum order is fixed

__global__ void __launch_bounds__(256) absmean_partial_kernel(const float* __restrict__ W, int64_t n,
                                                              float* __restrict__ partials) {
  pdl_entry();
  __shared__ float red[8];
  float acc = 0.f;
  const int64_t n4 = n >> 2;
  const float4* W4 = reinterpret_cast<const float4*>(W);
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 v = __ldg(W4 + i);
    acc += (fabsf(v.x) + fabsf(v.y)) + (fabsf(v.z) + fabsf(v.w));
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += 256) acc += fabsf(W[i]);
  float s = block_sum<256>(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(512) absmean_final_kernel(const float* __restrict__ partials, int nparts, int64_t n,
                                                            float* __restrict__ out) {
  pdl_entry();
  __shared__ float red[16];
  float v = threadIdx.x < nparts ? partials[threadIdx.x] : 0.f;
  float s = block_sum<512>(v, red);
  if (threadIdx.x == 0) out[0] = s / static_cast<float>(n);
}

// ---------------------------------------------------------------------------------------------
// weight quantise + pack  (quant.py:49-60).  One thread per packed 32-bit word (16 weights).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weight_pack_rows_kernel(const float* __restrict__ W, const float* __restrict__ alpha,
                                                               int alpha_mode, int64_t nwords, int bitwidth,
                                                               uint32_t* __restrict__ packed) {
  pdl_entry();
  const float a_eff = load_alpha_eff(alpha, alpha_mode);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
    const float4* src = reinterpret_cast<const float4*>(W + i * 16);
    uint32_t word = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float4 f = __ldg(src + v);
      word |= code_field(f.x, a_eff, bitwidth) << field_pos_i8(4 * v + 0);
      word |= code_field(f.y, a_eff, bitwidth) << field_pos_i8(4 * v + 1);
      word |= code_field(f.z, a_eff, bitwidth) << field_pos_i8(4 * v + 2);
      word |= code_field(f.w, a_eff, bitwidth) << field_pos_i8(4 * v + 3);
    }
    packed[i] = word;
  }
}

// Tile kernel (N % 64 == 0, K % 64 == 0): one read of W produces both layouts.  A block quantises a
// 64(n) x 64(k) tile into shared memory; 256 threads then build one row word (OB_ORDER_I8, 16 consecutive k
// of one n) and one transposed word (OB_ORDER_BF16, 16 consecutive n of one k) each.
__global__ void __launch_bounds__(256) weight_pack_tile_kernel(const float* __restrict__ W,
                                                               const float* __restrict__ alpha, int alpha_mode,
                                                               int N, int K, int bitwidth,
                                                               uint32_t* __restrict__ packed,
                                                               uint32_t* __restrict__ packed_t) {
  pdl_entry();
  __shared__ uint8_t fields[64][65];
  const float a_eff = load_alpha_eff(alpha, alpha_mode);
  const int n0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * 16; i += 256) {       // 64 rows x 16 float4, 256 B contiguous per 16 lanes
    const int r = i >> 4, c4 = i & 15;
    float4 f = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(n0 + r) * K + k0) + c4);
    fields[r][c4 * 4 + 0] = (uint8_t)code_field(f.x, a_eff, bitwidth);
    fields[r][c4 * 4 + 1] = (uint8_t)code_field(f.y, a_eff, bitwidth);
    fields[r][c4 * 4 + 2] = (uint8_t)code_field(f.z, a_eff, bitwidth);
    fields[r][c4 * 4 + 3] = (uint8_t)code_field(f.w, a_eff, bitwidth);
  }
  __syncthreads();
  const int r = threadIdx.x >> 2, g = threadIdx.x & 3;
  uint32_t word = 0;
#pragma unroll
  for (int t = 0; t < 16; ++t) word |= (uint32_t)fields[r][g * 16 + t] << field_pos_i8(t);
  packed[(int64_t)(n0 + r) * (K / 16) + k0 / 16 + g] = word;
  if (packed_t != nullptr) {                               // here r plays the role of k
    uint32_t wt = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) wt |= (uint32_t)fields[g * 16 + t][r] << field_pos_bf16(t);
    packed_t[(int64_t)(k0 + r) * (N / 16) + n0 / 16 + g] = wt;
  }
}

// Every routed layer of a model, both bitwidths, in ONE launch (the co-training step of train.py:83-103 needs the 2-bit and the
// 1-bit codes of all 108 layers once per optimiser step: 216 launches of the kernel above otherwise).  W is read once per
// tile and quantised both ways (the two code sets share W / alpha_eff and differ only in the 0.5 threshold).  Blocks map to
// (layer, 64 x 64 tile) through the descriptors' running tile offsets.
__global__ void __launch_bounds__(256) weight_pack_multi_kernel(const ob_pack_desc* __restrict__ descs, int count, int alpha_mode) {
  pdl_entry();
  __shared__ uint8_t fields[2][64][65];            // [0]: 2-bit codes, [1]: 1-bit codes
  int lo = 0, hi = count - 1;                      // last descriptor with tile0 <= blockIdx.x
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].tile0 <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const ob_pack_desc d = descs[lo];
  const int t = blockIdx.x - d.tile0, tiles_k = d.K / 64;
  const int n0 = (t / tiles_k) * 64, k0 = (t % tiles_k) * 64;
  const float a_eff = load_alpha_eff(d.alpha, alpha_mode);
  for (int i = threadIdx.x; i < 64 * 16; i += 256) {
    const int r = i >> 4, c4 = i & 15;
    const float4 f = __ldg(reinterpret_cast<const float4*>(d.W + (int64_t)(n0 + r) * d.K + k0) + c4);
    const float w[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      fields[0][r][c4 * 4 + u] = (uint8_t)code_field(w[u], a_eff, 2);
      fields[1][r][c4 * 4 + u] = (uint8_t)code_field(w[u], a_eff, 1);
    }
  }
  __syncthreads();
  const int r = threadIdx.x >> 2, g = threadIdx.x & 3;
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    uint32_t* packed = reinterpret_cast<uint32_t*>(b == 0 ? d.packed2 : d.packed1);
    uint32_t* packed_t = reinterpret_cast<uint32_t*>(b == 0 ? d.packed2_t : d.packed1_t);
    if (packed != nullptr) {
      uint32_t word = 0;
#pragma unroll
      for (int tt = 0; tt < 16; ++tt) word |= (uint32_t)fields[b][r][g * 16 + tt] << field_pos_i8(tt);
      packed[(int64_t)(n0 + r) * (d.K / 16) + k0 / 16 + g] = word;
    }
    if (packed_t != nullptr) {
      uint32_t wt = 0;
#pragma unroll
      for (int tt = 0; tt < 16; ++tt) wt |= (uint32_t)fields[b][g * 16 + tt][r] << field_pos_bf16(tt);
      packed_t[(int64_t)(k0 + r) * (d.N / 16) + n0 / 16 + g] = wt;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dense W_hat = alpha_eff * Q   (quantize_weight, quant.py:45-70)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float dense_code(float w, float a_eff, int bitwidth) {
  uint32_t f = code_field(w, a_eff, bitwidth);
  float q = f == 0u ? 0.0f : (f == 2u ? -1.0f : 1.0f);
  return __fmul_rn(a_eff, q);                                  // quant.py:68
}

__global__ void __launch_bounds__(256) weight_dense_kernel(const float* __restrict__ W, const float* __restrict__ alpha,
                                                           int alpha_mode, int64_t n, int bitwidth,
                                                           float* __restrict__ out) {
  pdl_entry();
  const float a_eff = load_alpha_eff(alpha, alpha_mode);
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    float4 f = __ldg(reinterpret_cast<const float4*>(W) + i);
    float4 o = make_float4(dense_code(f.x, a_eff, bitwidth), dense_code(f.y, a_eff, bitwidth),
                           dense_code(f.z, a_eff, bitwidth), dense_code(f.w, a_eff, bitwidth));
    reinterpret_cast<float4*>(out)[i] = o;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += 256) out[i] = dense_code(W[i], a_eff, bitwidth);
}

// ---------------------------------------------------------------------------------------------
// STE backward pieces (quant.py:80-92), shared by the dense path and the grad_W finalizer
// ---------------------------------------------------------------------------------------------
// returns masked gradient; adds g*term to acc
__device__ __forceinline__ float ste_elem(float g, float w, float a_eff, int bitwidth, float& acc) {
  const float wa = __fdiv_rn(w, a_eff);
  const float mag = fabsf(wa);
  const float sgn = wa > 0.f ? 1.f : (wa < 0.f ? -1.f : 0.f);
  float term;
  if (mag < 1.0f) {                                                     // strict, quant.py:87
    const float proj = (bitwidth == 2) ? (mag >= 0.5f ? sgn : 0.f) : sgn;   // quant.py:88
    term = __fadd_rn(-wa, proj);
  } else {
    term = sgn;                                                         // quant.py:89
  }
  acc = __fmaf_rn(g, term, acc);
  return mag <= 1.0f ? g : 0.f;                                         // quant.py:81-82
}

constexpr int kSteBlockElems = 4096;   // elements per block (256 threads x 4 float4)

// g_parts: [splits][n] partial gradients summed in split order (splits == 1: plain upstream gradient)
__global__ void __launch_bounds__(256) ste_backward_kernel(const float* __restrict__ g_parts, int splits,
                                                           const float* __restrict__ W, const float* __restrict__ alpha,
                                                           int alpha_mode, int64_t n, int bitwidth,
                                                           float* __restrict__ grad_W, float* __restrict__ alpha_parts) {
  pdl_entry();
  __shared__ float red[8];
  const float a_eff = load_alpha_eff(alpha, alpha_mode);
  float acc = 0.f;
  const int64_t base = (int64_t)blockIdx.x * kSteBlockElems;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int64_t i = base + (int64_t)(it * 256 + threadIdx.x) * 4;
    if (i + 3 < n) {
      float4 g = __ldg(reinterpret_cast<const float4*>(g_parts + i));
      for (int s = 1; s < splits; ++s) {
        float4 p = __ldg(reinterpret_cast<const float4*>(g_parts + (int64_t)s * n + i));
        g.x += p.x; g.y += p.y; g.z += p.z; g.w += p.w;
      }
      float4 w = __ldg(reinterpret_cast<const float4*>(W + i));
      float4 o;
      o.x = ste_elem(g.x, w.x, a_eff, bitwidth, acc);
      o.y = ste_elem(g.y, w.y, a_eff, bitwidth, acc);
      o.z = ste_elem(g.z, w.z, a_eff, bitwidth, acc);
      o.w = ste_elem(g.w, w.w, a_eff, bitwidth, acc);
      *reinterpret_cast<float4*>(grad_W + i) = o;
    } else {
      for (int64_t j = i; j < n; ++j) {
        float g = g_parts[j];
        for (int s = 1; s < splits; ++s) g += g_parts[(int64_t)s * n + j];
        grad_W[j] = ste_elem(g, W[j], a_eff, bitwidth, acc);
      }
    }
  }
  float s = block_sum<256>(acc, red);
  if (threadIdx.x == 0) alpha_parts[blockIdx.x] = s;
}

constexpr int kTailRowChunks = 8;   // the column-sum partial rows are reduced by 8 blocks per 32-column group

// grad_bias from the per-row-block column sums of dY: block (bx = 32-column group, by = row chunk) reduces its chunk;
// the last chunk to finish (ticket) adds the 8 chunk sums in fixed order.  chunk_part: [kTailRowChunks][N] floats,
// tickets: one int per column group, zero on entry, reset on exit.
__device__ __forceinline__ void bias_tail_block(int bx, int by, const float* __restrict__ colsum, int n_col_blocks, int N,
                                                float* __restrict__ grad_bias, float* __restrict__ chunk_part,
                                                int* __restrict__ tickets) {
  __shared__ float col_red[8][33];
  __shared__ int is_last;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = bx * 32 + cx;
  const int rows_per_chunk = (n_col_blocks + kTailRowChunks - 1) / kTailRowChunks;
  const int b0 = by * rows_per_chunk, b1 = min(n_col_blocks, b0 + rows_per_chunk);
  float a0 = 0.f, a1 = 0.f;
  if (c < N) {
    int b = b0 + ry;
    for (; b + 8 < b1; b += 16) {
      a0 += colsum[(int64_t)b * N + c];
      a1 += colsum[(int64_t)(b + 8) * N + c];
    }
    if (b < b1) a0 += colsum[(int64_t)b * N + c];
  }
  col_red[ry][cx] = a0 + a1;
  __syncthreads();
  if (ry == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += col_red[j][cx];
    chunk_part[(int64_t)by * N + c] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&tickets[bx], 1) == kTailRowChunks - 1;
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (ry == 0 && c < N) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < kTailRowChunks; ++j) t += __ldcg(&chunk_part[(int64_t)j * N + c]);
      grad_bias[c] = t;
    }
    if (threadIdx.x == 0) tickets[bx] = 0;
  }
}

// sum of the per-block alpha partials in fixed order, sign(alpha) chained in OB_ALPHA_RAW mode (quant.py:124)
__device__ __forceinline__ void alpha_final(const float* __restrict__ alpha_parts, int n_parts, const float* __restrict__ alpha,
                                            int alpha_mode, float* __restrict__ grad_alpha, float* red) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_parts; i += 256) acc += __ldcg(alpha_parts + i);
  float s = block_sum<256>(acc, red);
  if (threadIdx.x == 0) {
    if (alpha_mode == OB_ALPHA_RAW) {
      const float a = __ldg(alpha);
      s = a > 0.f ? s : (a < 0.f ? -s : 0.f);
    }
    grad_alpha[0] = s;
  }
}

// grad_W finaliser for many token splits: 256 elements per block; thread (eg, sl) sums splits sl, sl+4, ... of one
// float4, the four split lanes are combined through shared memory in fixed order (deterministic), then the STE.
constexpr int kFinBlockElems = 256;

// Blocks [0, fin_blocks) finalise 256 elements each; the last of them to finish (ticket) reduces the alpha partials.
// Blocks >= fin_blocks reduce the grad_bias column sums.  One launch for the whole tail of the backward.
// tail_ws: [kTailRowChunks][N] floats, then 1 + ceil(N/32) int tickets (zero on entry, reset on exit).
// Two groups of splits: [0, splits_a) are partials of token rows quantised at bw_a, [splits_a, splits) at bw_b (the stacked
// co-training passes; one group when splits_a == splits).  grad_W is the masked sum of both groups (the STE mask does not
// depend on the bitwidth), the alpha term is taken per group (quant.py:86-91 depends on the codes).
__global__ void __launch_bounds__(256) dw_finalize_kernel(const float* __restrict__ g_parts, int splits_a, int splits,
                                                          const float* __restrict__ W, const float* __restrict__ alpha,
                                                          int alpha_mode, int64_t n, int bw_a, int bw_b,
                                                          float* __restrict__ grad_W, float* __restrict__ alpha_parts,
                                                          int fin_blocks, float* __restrict__ grad_alpha,
                                                          const float* __restrict__ colsum, int n_col_blocks, int N,
                                                          float* __restrict__ grad_bias, float* __restrict__ tail_ws) {
  pdl_entry();
  __shared__ float4 part[3][64];
  __shared__ float4 part_b[3][64];
  __shared__ float red[8];
  __shared__ int last_fin;
  int* tickets = reinterpret_cast<int*>(tail_ws + (int64_t)kTailRowChunks * N);
  if (static_cast<int>(blockIdx.x) >= fin_blocks) {
    const int t = blockIdx.x - fin_blocks;
    bias_tail_block(t / kTailRowChunks, t % kTailRowChunks, colsum, n_col_blocks, N, grad_bias, tail_ws, tickets + 1);
    return;
  }
  const int eg = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const int64_t i = (int64_t)blockIdx.x * kFinBlockElems + eg * 4;      // n % 256 == 0 (N, K multiples of 64)
  // splits s0 + sl, s0 + sl + 4, ... < s1 of this thread's float4, two accumulators in flight
  auto lane_sum = [&](int s0, int s1) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    int s = s0 + sl;
    for (; s + 4 < s1; s += 8) {
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(g_parts + (int64_t)s * n + i));
      const float4 p1 = __ldg(reinterpret_cast<const float4*>(g_parts + (int64_t)(s + 4) * n + i));
      a0.x += p0.x; a0.y += p0.y; a0.z += p0.z; a0.w += p0.w;
      a1.x += p1.x; a1.y += p1.y; a1.z += p1.z; a1.w += p1.w;
    }
    if (s < s1) {
      const float4 p0 = __ldg(reinterpret_cast<const float4*>(g_parts + (int64_t)s * n + i));
      a0.x += p0.x; a0.y += p0.y; a0.z += p0.z; a0.w += p0.w;
    }
    a0.x += a1.x; a0.y += a1.y; a0.z += a1.z; a0.w += a1.w;
    return a0;
  };
  const bool two = splits_a < splits;                                   // block-uniform
  // the latent weights and alpha of the STE are requested before the partial sums (one memory round trip instead of two)
  float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
  float a_eff = 1.f;
  if (sl == 0) {
    w = __ldg(reinterpret_cast<const float4*>(W + i));
    a_eff = load_alpha_eff(alpha, alpha_mode);
  }
  float4 a0 = lane_sum(0, splits_a), b0 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (two) b0 = lane_sum(splits_a, splits);
  if (sl > 0) {
    part[sl - 1][eg] = a0;
    if (two) part_b[sl - 1][eg] = b0;
  }
  __syncthreads();
  float acc = 0.f;
  if (sl == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4 o = part[j][eg];
      a0.x += o.x; a0.y += o.y; a0.z += o.z; a0.w += o.w;
    }
    float4 o;
    o.x = ste_elem(a0.x, w.x, a_eff, bw_a, acc);
    o.y = ste_elem(a0.y, w.y, a_eff, bw_a, acc);
    o.z = ste_elem(a0.z, w.z, a_eff, bw_a, acc);
    o.w = ste_elem(a0.w, w.w, a_eff, bw_a, acc);
    if (two) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4 ob2 = part_b[j][eg];
        b0.x += ob2.x; b0.y += ob2.y; b0.z += ob2.z; b0.w += ob2.w;
      }
      o.x += ste_elem(b0.x, w.x, a_eff, bw_b, acc);
      o.y += ste_elem(b0.y, w.y, a_eff, bw_b, acc);
      o.z += ste_elem(b0.z, w.z, a_eff, bw_b, acc);
      o.w += ste_elem(b0.w, w.w, a_eff, bw_b, acc);
    }
    *reinterpret_cast<float4*>(grad_W + i) = o;
  }
  const float tot = block_sum<256>(acc, red);
  if (threadIdx.x == 0) {
    alpha_parts[blockIdx.x] = tot;
    __threadfence();
    last_fin = atomicAdd(&tickets[0], 1) == fin_blocks - 1;
  }
  __syncthreads();
  if (last_fin) {
    __threadfence();
    alpha_final(alpha_parts, fin_blocks, alpha, alpha_mode, grad_alpha, red);
    if (threadIdx.x == 0) tickets[0] = 0;
  }
}

// stand-alone tail (dense quantize_weight backward, or few splits): block 0 reduces the alpha partials, blocks >= 1
// the grad_bias column sums.  tail_ws as above (tickets start one int later to share the layout).
__global__ void __launch_bounds__(256) bwd_tail_kernel(const float* __restrict__ alpha_parts, int n_alpha_parts,
                                                       const float* __restrict__ alpha, int alpha_mode,
                                                       float* __restrict__ grad_alpha, const float* __restrict__ colsum,
                                                       int n_col_blocks, int N, float* __restrict__ grad_bias,
                                                       float* __restrict__ tail_ws) {
  pdl_entry();
  __shared__ float red[8];
  if (blockIdx.x == 0) {
    if (grad_alpha != nullptr) alpha_final(alpha_parts, n_alpha_parts, alpha, alpha_mode, grad_alpha, red);
    return;
  }
  if (grad_bias == nullptr) return;
  int* tickets = reinterpret_cast<int*>(tail_ws + (int64_t)kTailRowChunks * N);
  const int t = blockIdx.x - 1;
  bias_tail_block(t / kTailRowChunks, t % kTailRowChunks, colsum, n_col_blocks, N, grad_bias, tail_ws, tickets + 1);
}

// ---------------------------------------------------------------------------------------------
// unpack (tests / export)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t* __restrict__ packed, int64_t nwords, int order,
                                                     int8_t* __restrict__ codes) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * 256) {
    const uint32_t w = __ldg(packed + i);
    int8_t out[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const uint32_t f = (w >> (order == OB_ORDER_I8 ? field_pos_i8(t) : field_pos_bf16(t))) & 3u;
      out[t] = f == 2u ? -1 : (f == 3u ? 1 : 0);
    }
    *reinterpret_cast<int4*>(codes + i * 16) = *reinterpret_cast<int4*>(out);
  }
}

// ---------------------------------------------------------------------------------------------
// per-token absmax int8 activation quantiser.  One warp per row; the row lives in registers
// (V float4 per lane, K = 128*V) so x is read from HBM exactly once.
// ---------------------------------------------------------------------------------------------
// K == 128*V: lane l holds float4 index l + 32*j, j < V (coalesced 512-byte warp loads)
template <typename T, int V>
__global__ void __launch_bounds__(256) act_quant_reg_kernel(const T* __restrict__ x, int64_t M, int K,
                                                            int8_t* __restrict__ q, float* __restrict__ scale) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // the next row of this warp is in flight while the current one is reduced, quantised and stored
  // (rows of up to 1024 elements: a 2048-wide row already keeps 16 loads per lane in flight and would double to 150 registers)
  constexpr bool kPre = V <= 8;
  float4 nx[V];
  if (kPre && warp0 < M) {
#pragma unroll
    for (int j = 0; j < V; ++j) nx[j] = load4<T>(x + warp0 * K + (lane + 32 * j) * 4);
  }
  for (int64_t row = warp0; row < M; row += nwarps) {
    float4 v[V];
    uint32_t amax = 0u;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = kPre ? nx[j] : load4<T>(x + row * K + (lane + 32 * j) * 4);
      amax = amax_bits4(amax, v[j]);
    }
    if (kPre && row + nwarps < M) {
      const T* xn = x + (row + nwarps) * K;
#pragma unroll
      for (int j = 0; j < V; ++j) nx[j] = load4<T>(xn + (lane + 32 * j) * 4);
    }
    const float s = act_scale_from_amax(__uint_as_float(warp_max_bits(amax)));
    uint32_t* qr = reinterpret_cast<uint32_t*>(q + row * K);
#pragma unroll
    for (int j = 0; j < V; ++j) qr[lane + 32 * j] = quant4(v[j], s);
    if (lane == 0) scale[row] = s;
  }
}

// generic K (multiple of 16): two passes over the row, the second one hits L1/L2
template <typename T>
__global__ void __launch_bounds__(256) act_quant_generic_kernel(const T* __restrict__ x, int64_t M, int K,
                                                                int8_t* __restrict__ q, float* __restrict__ scale) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int k4 = K >> 2;
  for (int64_t row = warp0; row < M; row += nwarps) {
    const T* xr = x + row * K;
    uint32_t amax = 0u;
    for (int i = lane; i < k4; i += 32) amax = amax_bits4(amax, load4<T>(xr + i * 4));
    const float s = act_scale_from_amax(__uint_as_float(warp_max_bits(amax)));
    uint32_t* qr = reinterpret_cast<uint32_t*>(q + row * K);
    for (int i = lane; i < k4; i += 32) qr[i] = quant4(load4<T>(xr + i * 4), s);
    if (lane == 0) scale[row] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// fused FFN mid-section (conformer.py:36-39): z = dropout(swish(h)), then the activation quantiser of lin2.
// One read of h (+ the keep mask), one write of int8 codes: replaces sigmoid, mul, dropout and act-quant kernels.
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(256) swish_drop_quant_kernel(const float* __restrict__ h, const uint8_t* __restrict__ keep,
                                                               float inv_keep, DropRng rng, int64_t M, int K,
                                                               int8_t* __restrict__ q, float* __restrict__ scale) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // Rows of up to 512 columns prefetch the warp's next row.  At the FFN width (1024 columns, V = 8) the prefetch was measured
  // SLOWER (55.8 vs 49.4 us at 25536 rows): 96 registers halve the occupancy, and the long dependent chains of Philox and the
  // swish need the warps more than the loads need the head start.
  constexpr bool kPre = V <= 4;
  float4 nx[V];
  if (kPre && warp0 < M) {
#pragma unroll
    for (int j = 0; j < V; ++j) nx[j] = __ldg(reinterpret_cast<const float4*>(h + warp0 * K + (lane + 32 * j) * 4));
  }
  for (int64_t row = warp0; row < M; row += nwarps) {
    float4 v[V];
    uint32_t kb[V];                                     // keep flags of float4 j in bits 0..3
    uint32_t amax = 0u;
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = kPre ? nx[j] : __ldg(reinterpret_cast<const float4*>(h + row * K + (lane + 32 * j) * 4));
    if (kPre && row + nwarps < M) {
      const float* hn = h + (row + nwarps) * K;
#pragma unroll
      for (int j = 0; j < V; ++j) nx[j] = __ldg(reinterpret_cast<const float4*>(hn + (lane + 32 * j) * 4));
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      kb[j] = 0xFu;
      if (keep != nullptr) {
        const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(keep + row * K + (lane + 32 * j) * 4));
        kb[j] = (m.x ? 1u : 0u) | (m.y ? 2u : 0u) | (m.z ? 4u : 0u) | (m.w ? 8u : 0u);
      }
    }
    if (keep == nullptr && rng.threshold != 0u) {
      // float4 f of the flat [M, K] tensor uses lanes 4*((f >> 5) & 1) .. +3 of the block with counter f & ~32: one
      // Philox call serves this thread's float4s j and j + 1
#pragma unroll
      for (int j = 0; j < V; j += 2) {
        const unsigned long long f = static_cast<unsigned long long>(row) * (K >> 2) + lane + 32 * j;
        const uint32_t b8 = philox_keep8(f, rng);
        kb[j] = b8 & 0xFu;
        kb[j + 1] = b8 >> 4;
      }
    }
    const float ik = (keep != nullptr || rng.threshold != 0u) ? inv_keep : 1.0f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 t = v[j];
      t.x = (kb[j] & 1u) ? swish_f(t.x) * ik : 0.f;
      t.y = (kb[j] & 2u) ? swish_f(t.y) * ik : 0.f;
      t.z = (kb[j] & 4u) ? swish_f(t.z) * ik : 0.f;
      t.w = (kb[j] & 8u) ? swish_f(t.w) * ik : 0.f;
      v[j] = t;
      amax = amax_bits4(amax, t);
    }
    const float s = act_scale_from_amax(__uint_as_float(warp_max_bits(amax)));
    uint32_t* qr = reinterpret_cast<uint32_t*>(q + row * K);
#pragma unroll
    for (int j = 0; j < V; ++j) qr[lane + 32 * j] = quant4(v[j], s);
    if (lane == 0) scale[row] = s;
  }
}

// The same for 1024-column rows (the FFN width) with TWO warps per row: 16 instead of 32 values per lane keep the kernel at
// ~40 registers (75 % instead of 50 % occupancy) and double the independent Philox / swish chains per SM, which is what
// this kernel is short of (its loads are a quarter of a row's time).  The two half-row maxima meet in shared memory behind a
// 64-thread named barrier; max is order-independent, so codes and scales are bit-identical to the one-warp kernel.
template <int WPR>       // warps per row: 2 or 4
__global__ void __launch_bounds__(256) swish_drop_quant_k1024_kernel(const float* __restrict__ h, const uint8_t* __restrict__ keep,
                                                                     float inv_keep, DropRng rng, int64_t M,
                                                                     int8_t* __restrict__ q, float* __restrict__ scale) {
  pdl_entry();
  constexpr int K = 1024, V = 8 / WPR, RPB = 8 / WPR;   // float4 per lane, rows per block and iteration
  __shared__ uint32_t half_max[2][RPB][WPR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = warp / WPR, half = warp % WPR;
  const float ik = (keep != nullptr || rng.threshold != 0u) ? inv_keep : 1.0f;
  int it = 0;
  for (int64_t row = (int64_t)blockIdx.x * RPB + pair; row < M; row += (int64_t)gridDim.x * RPB, it ^= 1) {
    const int f0 = half * (32 * V) + lane;              // float4 index within the row of this lane's first value group
    float4 v[V];
    uint32_t kb[V];
    uint32_t amax = 0u;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = __ldg(reinterpret_cast<const float4*>(h + row * K) + f0 + 32 * j);
      kb[j] = 0xFu;
      if (keep != nullptr) {
        const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(keep + row * K) + f0 + 32 * j);
        kb[j] = (m.x ? 1u : 0u) | (m.y ? 2u : 0u) | (m.z ? 4u : 0u) | (m.w ? 8u : 0u);
      }
    }
    if (keep == nullptr && rng.threshold != 0u) {
#pragma unroll
      for (int j = 0; j < V; j += 2) {                  // float4s f and f + 32 share one Philox block (see above)
        const unsigned long long f = static_cast<unsigned long long>(row) * (K >> 2) + f0 + 32 * j;
        const uint32_t b8 = philox_keep8(f, rng);
        kb[j] = b8 & 0xFu;
        kb[j + 1] = b8 >> 4;
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 t = v[j];
      t.x = (kb[j] & 1u) ? swish_f(t.x) * ik : 0.f;
      t.y = (kb[j] & 2u) ? swish_f(t.y) * ik : 0.f;
      t.z = (kb[j] & 4u) ? swish_f(t.z) * ik : 0.f;
      t.w = (kb[j] & 8u) ? swish_f(t.w) * ik : 0.f;
      v[j] = t;
      amax = amax_bits4(amax, t);
    }
    amax = warp_max_bits(amax);
    if (lane == 0) half_max[it][pair][half] = amax;
    named_bar_sync(1 + pair, 32 * WPR);                 // the warps of this row (slots alternate: no second barrier needed)
    uint32_t row_max = half_max[it][pair][0];
#pragma unroll
    for (int u = 1; u < WPR; ++u) row_max = max(row_max, half_max[it][pair][u]);
    const float s = act_scale_from_amax(__uint_as_float(row_max));
    uint32_t* qr = reinterpret_cast<uint32_t*>(q + row * K);
#pragma unroll
    for (int j = 0; j < V; ++j) qr[f0 + 32 * j] = quant4(v[j], s);
    if (half == 0 && lane == 0) scale[row] = s;
  }
}

// g_h = g_z * keep * inv_keep * swish'(h),  swish'(h) = sig + h * sig * (1 - sig).  A thread owns the float4 pair
// (f, f + 32) that shares one Philox block (see swish_drop_quant_kernel); a warp still reads 512 contiguous bytes per load.
__global__ void __launch_bounds__(256) swish_drop_bwd_kernel(const float* __restrict__ gz, const float* __restrict__ h,
                                                             const uint8_t* __restrict__ keep, float inv_keep, DropRng rng,
                                                             int64_t npairs, float* __restrict__ gh) {
  pdl_entry();
  const float ik = (keep != nullptr || rng.threshold != 0u) ? inv_keep : 1.0f;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * 256) {
    const int64_t f0 = ((p >> 5) << 6) + (p & 31);
    float4 g[2], x[2];
    uint32_t kb = 0xFFu;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      g[c] = __ldg(reinterpret_cast<const float4*>(gz) + f0 + 32 * c);
      x[c] = __ldg(reinterpret_cast<const float4*>(h) + f0 + 32 * c);
    }
    if (keep != nullptr) {
      const uchar4 m0 = __ldg(reinterpret_cast<const uchar4*>(keep) + f0), m1 = __ldg(reinterpret_cast<const uchar4*>(keep) + f0 + 32);
      kb = (m0.x ? 1u : 0u) | (m0.y ? 2u : 0u) | (m0.z ? 4u : 0u) | (m0.w ? 8u : 0u) | (m1.x ? 16u : 0u) | (m1.y ? 32u : 0u) |
           (m1.z ? 64u : 0u) | (m1.w ? 128u : 0u);
    } else if (rng.threshold != 0u) {
      kb = philox_keep8(static_cast<unsigned long long>(f0), rng);
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float xs[4] = {x[c].x, x[c].y, x[c].z, x[c].w}, gs[4] = {g[c].x, g[c].y, g[c].z, g[c].w};
      float o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float sg = __fdividef(1.0f, 1.0f + __expf(-xs[u]));
        o[u] = ((kb >> (4 * c + u)) & 1u) ? gs[u] * ik * (sg + xs[u] * sg * (1.0f - sg)) : 0.f;
      }
      reinterpret_cast<float4*>(gh)[f0 + 32 * c] = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward prep: dys = bf16(g / s_m), qb = bf16(q), column sums of g per row block, where g is the layer's upstream
// gradient - read as is (mode 0) or produced on the fly from the gradient of the op that FOLLOWS the layer in the module:
//   mode 1  module tail  out = x + scale * dropout(y) * frame_mask  (conformer.py:41-45, 133-138):
//           g = dOut * scale * frame_mask[row] * keep / (1 - p)         (same Philox lanes as the forward epilogue)
//   mode 2  FFN mid-section  z = dropout(swish(h)) feeding lin2 (conformer.py:36-39), g is lin1's upstream gradient:
//           g = dZ * keep / (1 - p) * swish'(h)                          (same lanes as swish_drop_quant_kernel)
// so neither the tail's nor the activation's backward exists as a separate pass over [M, N] fp32.
// ---------------------------------------------------------------------------------------------
constexpr int kPrepRows = 32;
constexpr int kPrepChunk = 2048;            // columns per pass: 256 threads x 8 columns

struct PrepOp {
  const float* rowmask;     // mode 1: float validity per row (global row index), or nullptr
  const float* h;           // mode 2: pre-activation [M, N] fp32
  float factor;             // mode 1: scale / (1 - p);  mode 2: 1 / (1 - p)   (p = 0 when the stream is off)
  DropRng rng;              // threshold 0 -> no dropout
  long long row_base;       // global index of row 0 of this call (rows of one tensor processed in several calls)
};

template <typename T, int MODE>
__global__ void __launch_bounds__(256) bwd_prep_kernel(const T* __restrict__ dY, const float* __restrict__ scale,
                                                       const int8_t* __restrict__ q, int M, int N, int K,
                                                       __nv_bfloat16* __restrict__ dys, __nv_bfloat16* __restrict__ qb,
                                                       float* __restrict__ colsum, PrepOp op) {
  pdl_entry();
  __shared__ float inv_s[kPrepRows];
  __shared__ float rowf[kPrepRows];
  __shared__ float4 red[2][256];
  const int r0 = blockIdx.x * kPrepRows;
  const int rows = min(kPrepRows, M - r0);
  if (threadIdx.x < kPrepRows) {
    const bool ok = threadIdx.x < rows;
    inv_s[threadIdx.x] = ok ? __frcp_rn(__ldg(scale + r0 + threadIdx.x)) : 0.f;
    float f = op.factor;
    if (MODE == 1 && ok && op.rowmask != nullptr) f *= __ldg(op.rowmask + op.row_base + r0 + threadIdx.x);
    rowf[threadIdx.x] = f;
  }
  __syncthreads();
  // a thread owns two float4 of a row (8 columns); tpr threads cover a row of the chunk, rpi = 256 / tpr row groups walk
  // down the rows and their column sums are combined in fixed order
  for (int c0 = 0; c0 < N; c0 += kPrepChunk) {
    const int cw = min(kPrepChunk, N - c0);
    const int tpr = cw >> 3;
    const int rpi = 256 / tpr;
    const int rg = threadIdx.x / tpr;
    const int j = threadIdx.x - rg * tpr;
    int ca, cb;
    if (MODE == 2) {            // float4 pair (f, f + 32): the two halves of one Philox block of the FFN mid-section's stream
      ca = c0 + 4 * (((j >> 5) << 6) + (j & 31));
      cb = ca + 128;
    } else {
      ca = c0 + 8 * j;
      cb = ca + 4;
    }
    float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = acc_a;
    if (rg < rpi) {
      // U rows per step: all their 16-byte loads are issued before the first one is consumed (memory-level parallelism)
      constexpr int U = MODE == 2 ? 2 : 4;
      for (int rb = rg; rb < rows; rb += U * rpi) {
        float4 va_[U], vb_[U], ha_[U], hb_[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = min(rb + u * rpi, rows - 1);               // clamped: the tail re-reads the last row, result unused
          const int64_t off = (int64_t)(r0 + r) * N;
          va_[u] = load4<T>(dY + off + ca);
          vb_[u] = load4<T>(dY + off + cb);
          if (MODE == 2) {
            ha_[u] = __ldg(reinterpret_cast<const float4*>(op.h + off + ca));
            hb_[u] = __ldg(reinterpret_cast<const float4*>(op.h + off + cb));
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = rb + u * rpi;
          if (r >= rows) break;
          const int64_t off = (int64_t)(r0 + r) * N;
          float4 va = va_[u], vb = vb_[u];
          if (MODE == 1) {
            const float f = rowf[r];
            uint32_t kb = 0xFFu;
            if (op.rng.threshold != 0u)
              kb = philox_keep8(static_cast<unsigned long long>(((op.row_base + r0 + r) * N + ca) >> 3), op.rng);
            va.x = (kb & 1u) ? va.x * f : 0.f;   va.y = (kb & 2u) ? va.y * f : 0.f;
            va.z = (kb & 4u) ? va.z * f : 0.f;   va.w = (kb & 8u) ? va.w * f : 0.f;
            vb.x = (kb & 16u) ? vb.x * f : 0.f;  vb.y = (kb & 32u) ? vb.y * f : 0.f;
            vb.z = (kb & 64u) ? vb.z * f : 0.f;  vb.w = (kb & 128u) ? vb.w * f : 0.f;
          } else if (MODE == 2) {
            const float4 ha = ha_[u], hb = hb_[u];
            uint32_t kb = 0xFFu;
            if (op.rng.threshold != 0u)
              kb = philox_keep8(static_cast<unsigned long long>(((op.row_base + r0 + r) * N + ca) >> 2), op.rng);
            const float ik = op.factor;
            const float hs[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
            const float gs[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float sg = __fdividef(1.0f, 1.0f + __expf(-hs[e]));
              o[e] = ((kb >> e) & 1u) ? gs[e] * ik * (sg + hs[e] * sg * (1.0f - sg)) : 0.f;
            }
            va = make_float4(o[0], o[1], o[2], o[3]);
            vb = make_float4(o[4], o[5], o[6], o[7]);
          }
          acc_a.x += va.x; acc_a.y += va.y; acc_a.z += va.z; acc_a.w += va.w;
          acc_b.x += vb.x; acc_b.y += vb.y; acc_b.z += vb.z; acc_b.w += vb.w;
          const float is = inv_s[r];
          *reinterpret_cast<uint2*>(dys + off + ca) = pack_bf16x4(va.x * is, va.y * is, va.z * is, va.w * is);
          *reinterpret_cast<uint2*>(dys + off + cb) = pack_bf16x4(vb.x * is, vb.y * is, vb.z * is, vb.w * is);
        }
      }
    }
    if (colsum != nullptr) {
      red[0][threadIdx.x] = acc_a;
      red[1][threadIdx.x] = acc_b;
      __syncthreads();
      if (rg == 0) {
        for (int g = 1; g < rpi; ++g) {
          const float4 oa = red[0][g * tpr + threadIdx.x], ob2 = red[1][g * tpr + threadIdx.x];
          acc_a.x += oa.x; acc_a.y += oa.y; acc_a.z += oa.z; acc_a.w += oa.w;
          acc_b.x += ob2.x; acc_b.y += ob2.y; acc_b.z += ob2.z; acc_b.w += ob2.w;
        }
        *reinterpret_cast<float4*>(colsum + (int64_t)blockIdx.x * N + ca) = acc_a;
        *reinterpret_cast<float4*>(colsum + (int64_t)blockIdx.x * N + cb) = acc_b;
      }
      __syncthreads();
    }
  }
  // q -> bf16 (exact): each thread converts 8 codes per step
  if (qb != nullptr) {
    const int64_t total8 = (int64_t)rows * K / 8;
    const int8_t* qsrc = q + (int64_t)r0 * K;
    __nv_bfloat16* qdst = qb + (int64_t)r0 * K;
    for (int64_t i = threadIdx.x; i < total8; i += 256) {
      uint2 raw = __ldg(reinterpret_cast<const uint2*>(qsrc) + i);
      const int8_t* b = reinterpret_cast<const int8_t*>(&raw);
      uint4 o;
      uint2 lo = pack_bf16x4((float)b[0], (float)b[1], (float)b[2], (float)b[3]);
      uint2 hi = pack_bf16x4((float)b[4], (float)b[5], (float)b[6], (float)b[7]);
      o.x = lo.x; o.y = lo.y; o.z = hi.x; o.w = hi.y;
      reinterpret_cast<uint4*>(qdst)[i] = o;
    }
  }
}

}  // namespace ob

// =============================================================================================
// C ABI
// =============================================================================================
using namespace ob;

extern "C" size_t ob_absmean_workspace_bytes(void) { return kAbsmeanBlocks * sizeof(float); }

extern "C" int ob_weight_absmean(const float* W, int64_t n, float* out, void* ws, ob_stream_t stream) {
  OB_REQUIRE(W && out && ws && n > 0, "ob_weight_absmean: null pointer or n <= 0");
  OB_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0, "ob_weight_absmean: W must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  launch_k((absmean_partial_kernel), dim3(kAbsmeanBlocks), dim3(256), 0, st, W, n, static_cast<float*>(ws));
  OB_LAUNCH_CHECK("absmean_partial_kernel");
  launch_k((absmean_final_kernel), dim3(1), dim3(512), 0, st, static_cast<float*>(ws), kAbsmeanBlocks, n, out);
  OB_LAUNCH_CHECK("absmean_final_kernel");
  return OB_OK;
}

extern "C" int ob_weight_quant_pack(const float* W, const float* alpha, int alpha_mode, int N, int K, int bitwidth,
                                    uint8_t* packed_i8, uint8_t* packed_t, ob_stream_t stream) {
  OB_REQUIRE(W && alpha && packed_i8, "ob_weight_quant_pack: null pointer");
  OB_REQUIRE(bitwidth == 1 || bitwidth == 2, "bitwidth must be one of {1,2,32}");
  OB_REQUIRE(N > 0 && K > 0 && K % 16 == 0, "ob_weight_quant_pack: K (%d) must be a positive multiple of 16", K);
  OB_REQUIRE(packed_t == nullptr || (N % 64 == 0 && K % 64 == 0),
             "ob_weight_quant_pack: the transposed copy needs N %% 64 == 0 and K %% 64 == 0 (N=%d K=%d)", N, K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N % 64 == 0 && K % 64 == 0) {
    dim3 grid(K / 64, N / 64);
    launch_k((weight_pack_tile_kernel), dim3(grid), dim3(256), 0, st, W, alpha, alpha_mode, N, K, bitwidth,
                                                 reinterpret_cast<uint32_t*>(packed_i8),
                                                 reinterpret_cast<uint32_t*>(packed_t));
    OB_LAUNCH_CHECK("weight_pack_tile_kernel");
    return OB_OK;
  }
  const int64_t nwords = (int64_t)N * K / 16;
  const int64_t want = (nwords + 255) / 256;
  const int blocks = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
  launch_k((weight_pack_rows_kernel), dim3(blocks), dim3(256), 0, st, W, alpha, alpha_mode, nwords, bitwidth,
                                                 reinterpret_cast<uint32_t*>(packed_i8));
  OB_LAUNCH_CHECK("weight_pack_rows_kernel");
  return OB_OK;
}

extern "C" int ob_weight_quant_pack_multi(const ob_pack_desc* descs_dev, int count, int total_tiles, int alpha_mode,
                                          ob_stream_t stream) {
  OB_REQUIRE(descs_dev && count > 0 && total_tiles > 0, "ob_weight_quant_pack_multi: null descriptor table or empty");
  OB_REQUIRE(alpha_mode == OB_ALPHA_RAW || alpha_mode == OB_ALPHA_EFF, "ob_weight_quant_pack_multi: unknown alpha mode %d", alpha_mode);
  launch_k((weight_pack_multi_kernel), dim3(total_tiles), dim3(256), 0, static_cast<cudaStream_t>(stream), descs_dev, count, alpha_mode);
  OB_LAUNCH_CHECK("weight_pack_multi_kernel");
  return OB_OK;
}

extern "C" int ob_weight_quant_dense(const float* W, const float* alpha, int alpha_mode, int64_t n, int bitwidth,
                                     float* w_hat, ob_stream_t stream) {
  OB_REQUIRE(W && alpha && w_hat && n > 0, "ob_weight_quant_dense: null pointer or n <= 0");
  OB_REQUIRE(bitwidth == 1 || bitwidth == 2, "bitwidth must be one of {1,2,32}");
  const int64_t want = (n / 4 + 255) / 256 + 1;
  const int blocks = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
  launch_k((weight_dense_kernel), dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), W, alpha, alpha_mode, n, bitwidth, w_hat);
  OB_LAUNCH_CHECK("weight_dense_kernel");
  return OB_OK;
}

static int ste_blocks(int64_t n) { return (int)((n + kSteBlockElems - 1) / kSteBlockElems); }

extern "C" size_t ob_ste_workspace_bytes(int64_t n) {
  const size_t fin = (size_t)((n + kFinBlockElems - 1) / kFinBlockElems);
  const size_t ste = (size_t)ste_blocks(n);
  return (fin > ste ? fin : ste) * sizeof(float);
}
// extra floats the grad_bias tail needs behind the alpha partials
static size_t tail_workspace_bytes(int N) { return ((size_t)kTailRowChunks * N + (size_t)(N + 31) / 32 + 9) * sizeof(float); }
namespace ob { size_t bwd_tail_workspace_bytes(int N) { return tail_workspace_bytes(N); } }

namespace ob {
// shared with ob_gemm.cu (grad_W finalizer)
int launch_ste_and_tail(const float* g_parts, int splits, const float* W, const float* alpha, int alpha_mode, int64_t n,
                        int bitwidth, float* grad_W, float* grad_alpha, float* alpha_parts, const float* colsum,
                        int n_col_blocks, int N, float* grad_bias, cudaStream_t st) {
  const int col_groups = (grad_bias != nullptr) ? (N + 31) / 32 : 0;
  const bool fused = splits >= 4 && n % kFinBlockElems == 0;
  const int blocks = fused ? (int)(n / kFinBlockElems) : ste_blocks(n);
  float* tail_ws = alpha_parts + blocks;                           // [kTailRowChunks][N] floats + (1 + col_groups) tickets
  if (fused || col_groups > 0)
    OB_CUDA(cudaMemsetAsync(tail_ws + (size_t)kTailRowChunks * N, 0, (size_t)(1 + col_groups) * sizeof(int), st));
  if (fused) {
    launch_k((dw_finalize_kernel), dim3(blocks + col_groups * kTailRowChunks), dim3(256), 0, st, 
        g_parts, splits, splits, W, alpha, alpha_mode, n, bitwidth, bitwidth, grad_W, alpha_parts, blocks, grad_alpha, colsum,
        n_col_blocks, N, grad_bias, tail_ws);
    OB_LAUNCH_CHECK("dw_finalize_kernel");
    return OB_OK;
  }
  launch_k((ste_backward_kernel), dim3(blocks), dim3(256), 0, st, g_parts, splits, W, alpha, alpha_mode, n, bitwidth, grad_W, alpha_parts);
  OB_LAUNCH_CHECK("ste_backward_kernel");
  launch_k((bwd_tail_kernel), dim3(1 + col_groups * kTailRowChunks), dim3(256), 0, st, alpha_parts, blocks, alpha, alpha_mode, grad_alpha,
                                                                   colsum, n_col_blocks, N, grad_bias, tail_ws);
  OB_LAUNCH_CHECK("bwd_tail_kernel");
  return OB_OK;
}

// the fused finaliser over two groups of splits (any split count; n % 256 == 0 holds for N, K multiples of 64)
// Where the finaliser's tickets live inside the workspace (behind the alpha partials and the column-sum chunk rows) and how
// many there are: the grad_W GEMM kernel zeroes them itself (one thread, before its main loop), which saves a memset node per layer.
int* dw_finalize_tickets(float* alpha_parts, int64_t n, int N, int with_bias, int* count) {
  const int col_groups = with_bias ? (N + 31) / 32 : 0;
  *count = 1 + col_groups;
  return reinterpret_cast<int*>(alpha_parts + n / kFinBlockElems + (size_t)kTailRowChunks * N);
}

int launch_dw_finalize_groups(const float* g_parts, int splits_a, int splits, const float* W, const float* alpha, int alpha_mode,
                              int64_t n, int bw_a, int bw_b, float* grad_W, float* grad_alpha, float* alpha_parts,
                              const float* colsum, int n_col_blocks, int N, float* grad_bias, cudaStream_t st) {
  if (n % kFinBlockElems != 0) {
    set_error("grad_W finaliser: N * K must be a multiple of %d", kFinBlockElems);
    return OB_ERR_ARG;
  }
  if (splits_a == 0) { splits_a = splits; bw_a = bw_b; }           // only the second group has rows
  const int col_groups = (grad_bias != nullptr) ? (N + 31) / 32 : 0;
  const int blocks = (int)(n / kFinBlockElems);
  float* tail_ws = alpha_parts + blocks;
  (void)col_groups;                                  // tickets zeroed by dw_pair_kernel (dw_finalize_tickets)
  launch_k((dw_finalize_kernel), dim3(blocks + col_groups * kTailRowChunks), dim3(256), 0, st, 
      g_parts, splits_a, splits, W, alpha, alpha_mode, n, bw_a, bw_b, grad_W, alpha_parts, blocks, grad_alpha, colsum,
      n_col_blocks, N, grad_bias, tail_ws);
  OB_LAUNCH_CHECK("dw_finalize_kernel");
  return OB_OK;
}
}  // namespace ob

extern "C" int ob_weight_ste_backward(const float* g, const float* W, const float* alpha, int alpha_mode, int64_t n,
                                      int bitwidth, float* grad_W, float* grad_alpha, void* ws, ob_stream_t stream) {
  OB_REQUIRE(g && W && alpha && grad_W && grad_alpha && ws && n > 0, "ob_weight_ste_backward: null pointer or n <= 0");
  OB_REQUIRE(bitwidth == 1 || bitwidth == 2, "bitwidth must be one of {1,2,32}");
  OB_REQUIRE(((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(grad_W)) & 15) == 0,
             "ob_weight_ste_backward: pointers must be 16-byte aligned");
  return launch_ste_and_tail(g, 1, W, alpha, alpha_mode, n, bitwidth, grad_W, grad_alpha, static_cast<float*>(ws),
                             nullptr, 0, 0, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int ob_unpack_codes(const uint8_t* packed, int R, int C, int order, int8_t* codes, ob_stream_t stream) {
  OB_REQUIRE(packed && codes && R > 0 && C > 0 && C % 16 == 0, "ob_unpack_codes: bad arguments");
  OB_REQUIRE(order == OB_ORDER_I8 || order == OB_ORDER_BF16, "ob_unpack_codes: unknown order %d", order);
  const int64_t nwords = (int64_t)R * C / 16;
  const int blocks = (int)((nwords + 255) / 256 < 4096 ? (nwords + 255) / 256 : 4096);
  launch_k((unpack_kernel), dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const uint32_t*>(packed), nwords,
                                                                      order, codes);
  OB_LAUNCH_CHECK("unpack_kernel");
  return OB_OK;
}

template <typename T>
static int launch_act_quant(const T* x, int64_t M, int K, int8_t* q, float* scale, cudaStream_t st) {
  const int64_t want = (M + 7) / 8;                               // 8 warps (rows) per block
  const int blocks = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
  switch (K) {
    case 128:  launch_k((act_quant_reg_kernel<T, 1>), dim3(blocks), dim3(256), 0, st, x, M, K, q, scale); break;
    case 256:  launch_k((act_quant_reg_kernel<T, 2>), dim3(blocks), dim3(256), 0, st, x, M, K, q, scale); break;
    case 512:  launch_k((act_quant_reg_kernel<T, 4>), dim3(blocks), dim3(256), 0, st, x, M, K, q, scale); break;
    case 1024: launch_k((act_quant_reg_kernel<T, 8>), dim3(blocks), dim3(256), 0, st, x, M, K, q, scale); break;
    case 2048: launch_k((act_quant_reg_kernel<T, 16>), dim3(blocks), dim3(256), 0, st, x, M, K, q, scale); break;
    default:   launch_k((act_quant_generic_kernel<T>), dim3(blocks), dim3(256), 0, st, x, M, K, q, scale); break;
  }
  OB_LAUNCH_CHECK("act_quant kernel");
  return OB_OK;
}

extern "C" int ob_act_quant_i8(const void* x, int x_dtype, int64_t M, int K, int8_t* q, float* scale,
                               ob_stream_t stream) {
  OB_REQUIRE(x && q && scale, "ob_act_quant_i8: null pointer");
  OB_REQUIRE(M > 0 && K > 0 && K % 16 == 0, "ob_act_quant_i8: K (%d) must be a positive multiple of 16", K);
  OB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0,
             "ob_act_quant_i8: x and q must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_dtype == OB_F32) return launch_act_quant(static_cast<const float*>(x), M, K, q, scale, st);
  if (x_dtype == OB_BF16) return launch_act_quant(static_cast<const __nv_bfloat16*>(x), M, K, q, scale, st);
  OB_REQUIRE(false, "ob_act_quant_i8: unknown dtype tag %d", x_dtype);
}

extern "C" int ob_bwd_colsum_blocks(int M) { return (M + kPrepRows - 1) / kPrepRows; }

extern "C" int ob_bwd_prep(const void* dY, int dy_dtype, const float* scale, const int8_t* q, int M, int N, int K,
                           void* dys_bf16, void* qb_bf16, float* colsum, ob_stream_t stream) {
  OB_REQUIRE(dY && scale && dys_bf16, "ob_bwd_prep: null pointer");
  OB_REQUIRE(qb_bf16 == nullptr || q != nullptr, "ob_bwd_prep: qb requested without q");
  OB_REQUIRE(M > 0 && N > 0 && N % 8 == 0 && K % 8 == 0, "ob_bwd_prep: need N %% 8 == 0 and K %% 8 == 0 (N=%d K=%d)", N, K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = ob_bwd_colsum_blocks(M);
  const PrepOp op = {nullptr, nullptr, 1.0f, {0ull, 0ull, 0u}, 0ll};
  if (dy_dtype == OB_F32)
    launch_k((bwd_prep_kernel<float, 0>), dim3(blocks), dim3(256), 0, st, static_cast<const float*>(dY), scale, q, M, N, K,
                                                      static_cast<__nv_bfloat16*>(dys_bf16),
                                                      static_cast<__nv_bfloat16*>(qb_bf16), colsum, op);
  else if (dy_dtype == OB_BF16)
    launch_k((bwd_prep_kernel<__nv_bfloat16, 0>), dim3(blocks), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(dY), scale, q, M, N, K,
                                                              static_cast<__nv_bfloat16*>(dys_bf16),
                                                              static_cast<__nv_bfloat16*>(qb_bf16), colsum, op);
  else
    OB_REQUIRE(false, "ob_bwd_prep: unknown dtype tag %d", dy_dtype);
  OB_LAUNCH_CHECK("bwd_prep_kernel");
  return OB_OK;
}

extern "C" int ob_bwd_prep_fused(const float* g_next, int mode, const float* rowmask, const float* h, float factor,
                                 uint64_t seed, uint64_t offset, uint32_t drop_threshold, int64_t row_base,
                                 const float* scale, const int8_t* q, int M, int N, int K, void* dys_bf16, void* qb_bf16,
                                 float* colsum, ob_stream_t stream) {
  OB_REQUIRE(g_next && scale && dys_bf16, "ob_bwd_prep_fused: null pointer");
  OB_REQUIRE(mode == OB_PREP_TAIL || mode == OB_PREP_SWISH, "ob_bwd_prep_fused: mode must be OB_PREP_TAIL or OB_PREP_SWISH (%d)", mode);
  OB_REQUIRE(mode != OB_PREP_SWISH || h != nullptr, "ob_bwd_prep_fused: OB_PREP_SWISH needs the pre-activation h");
  OB_REQUIRE(qb_bf16 == nullptr || q != nullptr, "ob_bwd_prep_fused: qb requested without q");
  OB_REQUIRE(drop_threshold < 65536u, "ob_bwd_prep_fused: drop_threshold (%u) is a 16-bit value", drop_threshold);
  OB_REQUIRE(M > 0 && N > 0 && K % 8 == 0 && row_base >= 0, "ob_bwd_prep_fused: bad sizes (M=%d N=%d K=%d)", M, N, K);
  OB_REQUIRE(mode == OB_PREP_SWISH ? N % 256 == 0 : N % 8 == 0,
             "ob_bwd_prep_fused: N (%d) must be a multiple of 8 (tail) / 256 (swish)", N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = ob_bwd_colsum_blocks(M);
  const PrepOp op = {rowmask, h, factor, {seed, offset, drop_threshold}, static_cast<long long>(row_base)};
  if (mode == OB_PREP_TAIL)
    launch_k((bwd_prep_kernel<float, 1>), dim3(blocks), dim3(256), 0, st, g_next, scale, q, M, N, K, static_cast<__nv_bfloat16*>(dys_bf16),
                                                      static_cast<__nv_bfloat16*>(qb_bf16), colsum, op);
  else
    launch_k((bwd_prep_kernel<float, 2>), dim3(blocks), dim3(256), 0, st, g_next, scale, q, M, N, K, static_cast<__nv_bfloat16*>(dys_bf16),
                                                      static_cast<__nv_bfloat16*>(qb_bf16), colsum, op);
  OB_LAUNCH_CHECK("bwd_prep_kernel(fused)");
  return OB_OK;
}

extern "C" int ob_swish_drop_quant(const float* h, const uint8_t* keep, float inv_keep, uint64_t seed, uint64_t offset,
                                   uint32_t drop_threshold, int64_t M, int K, int8_t* q, float* scale, ob_stream_t stream) {
  OB_REQUIRE(drop_threshold < 65536u, "ob_swish_drop_quant: drop_threshold (%u) is a 16-bit value", drop_threshold);
  const DropRng rng = {seed, offset, drop_threshold};
  OB_REQUIRE(h && q && scale && M > 0, "ob_swish_drop_quant: null pointer or M <= 0");
  OB_REQUIRE(K == 256 || K == 512 || K == 1024 || K == 2048, "ob_swish_drop_quant: K (%d) must be 256, 512, 1024 or 2048", K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t want = (M + 7) / 8;
  const int blocks = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
  switch (K) {
    case 256:  launch_k((swish_drop_quant_kernel<2>), dim3(blocks), dim3(256), 0, st, h, keep, inv_keep, rng, M, K, q, scale); break;
    case 512:  launch_k((swish_drop_quant_kernel<4>), dim3(blocks), dim3(256), 0, st, h, keep, inv_keep, rng, M, K, q, scale); break;
    case 1024: {
      // two warps per row: 138.6 -> 120 us at 76608 rows; four warps per row (100 % occupancy) measured 138 us again
      const int64_t want4 = (M + 3) / 4;
      const int blocks4 = (int)(want4 < (int64_t)num_sms() * 12 ? want4 : (int64_t)num_sms() * 12);
      launch_k((swish_drop_quant_k1024_kernel<2>), dim3(blocks4), dim3(256), 0, st, h, keep, inv_keep, rng, M, q, scale);
      break;
    }
    default:   launch_k((swish_drop_quant_kernel<16>), dim3(blocks), dim3(256), 0, st, h, keep, inv_keep, rng, M, K, q, scale); break;
  }
  OB_LAUNCH_CHECK("swish_drop_quant_kernel");
  return OB_OK;
}

extern "C" int ob_swish_drop_bwd(const float* gz, const float* h, const uint8_t* keep, float inv_keep, uint64_t seed,
                                 uint64_t offset, uint32_t drop_threshold, int64_t n, float* gh, ob_stream_t stream) {
  OB_REQUIRE(gz && h && gh && n > 0 && n % 256 == 0, "ob_swish_drop_bwd: null pointer or n not a positive multiple of 256");
  OB_REQUIRE(drop_threshold < 65536u, "ob_swish_drop_bwd: drop_threshold (%u) is a 16-bit value", drop_threshold);
  const DropRng rng = {seed, offset, drop_threshold};
  const int64_t npairs = n / 8;
  const int64_t want = (npairs + 255) / 256;
  const int blocks = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  launch_k((swish_drop_bwd_kernel), dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), gz, h, keep, inv_keep, rng, npairs, gh);
  OB_LAUNCH_CHECK("swish_drop_bwd_kernel");
  return OB_OK;
}
