// ob_common.cuh - shared device helpers: error plumbing, PTX wrappers for mbarrier / TMA / tcgen05
// on sm_100a, and the 2-bit code encode/decode used by every kernel.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "onebit.h"

namespace ob {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local last error, no exceptions across the C ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_device();   // OB_OK when the current device is sm_100
void count_launch();  // bumps the library-wide kernel launch counter (ob_launch_count)
int sm_count();       // SMs of the current device (ob_debug_set can cap it)
int pdl_enabled();    // 1: kernels are launched with programmatic stream serialization (ob_debug_set key 13 turns it off)
void set_pdl(int on);

// Every kernel of the library goes through this launcher.  With programmatic dependent launch the grid may be scheduled
// while the previous kernel of the stream is still draining: its blocks run up to their griddepcontrol.wait (pdl_entry /
// pdl_wait, executed before the first access to global memory) and continue once the predecessor has completed and its
// writes are visible, so the launch latency and the block prologues overlap the predecessor's tail.
static inline cudaLaunchAttribute pdl_attribute() {
  cudaLaunchAttribute a;
  a.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a.val.programmaticStreamSerializationAllowed = pdl_enabled();
  return a;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cuTensorMapEncodeTiled resolved at run time (no link-time dependency on libcuda); nullptr when unavailable
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn();

#define OB_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      ::ob::set_error(__VA_ARGS__);           \
      return OB_ERR_ARG;                      \
    }                                         \
  } while (0)

#define OB_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::ob::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                      __LINE__);                                                           \
      return OB_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define OB_LAUNCH_CHECK(name)                                                              \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      ::ob::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));            \
      return OB_ERR_CUDA;                                                                  \
    }                                                                                      \
    ::ob::count_launch();                                                                  \
  } while (0)

// ---------------------------------------------------------------------------------------------
// quantiser arithmetic shared by all kernels (bit-exact with the reference's ATen ops)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float load_alpha_eff(const float* alpha, int alpha_mode) {
  float a = __ldg(alpha);
  return alpha_mode == OB_ALPHA_RAW ? __fadd_rn(fabsf(a), 1e-8f) : a;   // quant.py:124
}

// field encoding: 0b00 = 0, 0b10 = -1, 0b11 = +1
__device__ __forceinline__ uint32_t code_field(float w, float a_eff, int bitwidth) {
  float wa = __fdiv_rn(w, a_eff);                       // quant.py:49 (IEEE division, never rcp*mul)
  float mag = fminf(fabsf(wa), 1.0f);                   // |clamp(wa,-1,1)|            quant.py:50
  uint32_t neg = wa < 0.0f;                             // sign(-0.0) == 0 -> not negative
  if (bitwidth == 1) return neg ? 2u : 3u;              // sign, zeros -> +1           quant.py:52-55
  if (mag < 0.5f) return 0u;                            // strict threshold            quant.py:60
  return neg ? 2u : 3u;
}

__device__ __forceinline__ int field_pos_i8(int t) { return 4 * (t & 3) + 2 * ((t >> 2) & 1) + 16 * (t >> 3); }
__device__ __forceinline__ int field_pos_bf16(int t) { return 8 * (t & 1) + 2 * ((t >> 1) & 3) + 16 * (t >> 3); }

// One packed word (OB_ORDER_I8) -> 16 int8 codes (4 words).  PRMT table: idx0,1 -> 0, idx2 -> -1, idx3 -> +1.
__device__ __forceinline__ uint4 expand_word_i8(uint32_t w) {
  constexpr uint32_t kLut = 0x01FF0000u;
  uint4 o;
  o.x = __byte_perm(kLut, 0u, w & 0x3333u);
  o.y = __byte_perm(kLut, 0u, (w >> 2) & 0x3333u);
  o.z = __byte_perm(kLut, 0u, (w >> 16) & 0x3333u);
  o.w = __byte_perm(kLut, 0u, (w >> 18) & 0x3333u);
  return o;
}

// One packed word (OB_ORDER_BF16) -> 16 bf16 codes (8 words, two 16-byte chunks).
// hi-byte table (selector f):   idx0,1 -> 0x00, idx2 -> 0xBF, idx3 -> 0x3F
// lo-byte table (selector f|4): idx4,5 -> 0x00, idx6,7 -> 0x80
__device__ __forceinline__ void expand_word_bf16(uint32_t w, uint4& c0, uint4& c1) {
  constexpr uint32_t kHi = 0x3FBF0000u, kLo = 0x80800000u;
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t v = (w >> (2 * (j & 3) + 16 * (j >> 2))) & 0x0303u;   // fields 2j (bit 0) and 2j+1 (bit 8)
    uint32_t sel = v * 0x11u + 0x0404u;                            // nibbles (f0|4, f0, f1|4, f1)
    o[j] = __byte_perm(kHi, kLo, sel);
  }
  c0 = make_uint4(o[0], o[1], o[2], o[3]);
  c1 = make_uint4(o[4], o[5], o[6], o[7]);
}

// ---------------------------------------------------------------------------------------------
// counter-based dropout masks: Philox4x32-10 (Salmon et al., SC'11).  One call yields four 32-bit words for the counter
// (group index, call offset) under the key `seed`, read as eight 16-bit lanes (lane k = half k % 2 of word k / 2, low half
// first); an element is kept when its lane is >= threshold (= p * 2^16, so p is realised to 2^-17 and the caller rescales
// by the exact 65536 / (65536 - threshold)).  The backward regenerates the same lanes instead of reading a stored mask.
// ---------------------------------------------------------------------------------------------
struct DropRng {
  unsigned long long seed;
  unsigned long long offset;
  uint32_t threshold;          // 16-bit; 0 -> dropout off
};

__device__ __forceinline__ uint4 philox4x32_10(unsigned long long group, unsigned long long offset, unsigned long long seed) {
  uint32_t c0 = static_cast<uint32_t>(group), c1 = static_cast<uint32_t>(group >> 32);
  uint32_t c2 = static_cast<uint32_t>(offset), c3 = static_cast<uint32_t>(offset >> 32);
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// keep flags of the eight lanes of one block: bit k set = lane k kept
__device__ __forceinline__ uint32_t philox_keep8(unsigned long long group, const DropRng& rng) {
  const uint4 w = philox4x32_10(group, rng.offset, rng.seed);
  const uint32_t t = rng.threshold;
  uint32_t bits = 0;
  bits |= ((w.x & 0xFFFFu) >= t ? 1u : 0u) | ((w.x >> 16) >= t ? 2u : 0u);
  bits |= ((w.y & 0xFFFFu) >= t ? 4u : 0u) | ((w.y >> 16) >= t ? 8u : 0u);
  bits |= ((w.z & 0xFFFFu) >= t ? 16u : 0u) | ((w.z >> 16) >= t ? 32u : 0u);
  bits |= ((w.w & 0xFFFFu) >= t ? 64u : 0u) | ((w.w >> 16) >= t ? 128u : 0u);
  return bits;
}


// ---------------------------------------------------------------------------------------------
// warp reductions, vector loads, the int8 activation quantiser's arithmetic, bf16 packing, swish
// (shared by ob_quant.cu, ob_norm.cu and the fused prologue / epilogue kernels)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x), b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

// Row absmax accumulated on the BIT PATTERNS of |v| (unsigned integer max): order-preserving for non-negative floats, and
// a NaN (pattern > 0x7f800000) or Inf in the row comes out on top instead of being dropped as fmaxf would.
__device__ __forceinline__ uint32_t amax_bits4(uint32_t acc, float4 v) {
  acc = max(acc, __float_as_uint(fabsf(v.x)));
  acc = max(acc, __float_as_uint(fabsf(v.y)));
  acc = max(acc, __float_as_uint(fabsf(v.z)));
  return max(acc, __float_as_uint(fabsf(v.w)));
}
__device__ __forceinline__ uint32_t warp_max_bits(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float act_scale_from_amax(float amax) {
  // torch evaluates `127.0 / t` as reciprocal(t) * 127 (two roundings); pinned bit-exactly by the oracle.
  // A non-finite row (NaN or Inf absmax) gets a NaN scale, so the GEMM epilogue's alpha / s turns the whole output row
  // into NaN: divergence stays visible instead of being clamped to -128 codes.
  if (!(amax <= 3.402823466e38f)) return __int_as_float(0x7fc00000);
  return __fmul_rn(__frcp_rn(fmaxf(amax, 1e-5f)), 127.0f);
}
// q = clamp(rint(v * s), -128, 127) for four values, packed little-endian.  Clamping to the (integer) bounds first and
// then adding 1.5 * 2^23 rounds to the nearest integer, ties to even, exactly like rint() - the sum's low mantissa
// byte IS the int8 two's-complement code - and it stays on the full-rate FP pipes (FRND and F2I run on the
// quarter-rate XU pipe, which ncu showed ~50-60 % busy in these kernels).
__device__ __forceinline__ uint32_t quant4(float4 v, float s) {
  constexpr float kMagic = 12582912.0f;
  const uint32_t a = __float_as_uint(fminf(fmaxf(__fmul_rn(v.x, s), -128.f), 127.f) + kMagic);
  const uint32_t b = __float_as_uint(fminf(fmaxf(__fmul_rn(v.y, s), -128.f), 127.f) + kMagic);
  const uint32_t c = __float_as_uint(fminf(fmaxf(__fmul_rn(v.z, s), -128.f), 127.f) + kMagic);
  const uint32_t d = __float_as_uint(fminf(fmaxf(__fmul_rn(v.w, s), -128.f), 127.f) + kMagic);
  return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// swish with the SFU exponential (ex2.approx, ~2 ulp) and an IEEE reciprocal: the full-precision expf made this
// kernel ALU-bound (26 M exponentials per launch at the FFN width); the result differs from torch's sigmoid*x in
// the last bits only, which moves an int8 code by at most one step on a ~1e-5 fraction of elements
// swish(h) = h * sigmoid(h) on the SFU (ex2 + rcp, ~2 ulp); exp(-h) = inf for h < -88 gives h * 0 = -0
__device__ __forceinline__ float swish_f(float h) { return h * __fdividef(1.0f, 1.0f + __expf(-h)); }

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&lo);
  r.y = *reinterpret_cast<uint32_t*>(&hi);
  return r;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------
// programmatic dependent launch: block until the preceding kernel of the stream has completed (no-op without the launch
// attribute); then let the NEXT kernel's blocks be scheduled as soon as resources free up
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_entry() {
  pdl_wait();
  pdl_launch_dependents();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 2-D tiled TMA load: coordinates are (c0 = innermost element index, c1 = row index)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c0,
                                             int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// 4-D tiled TMA (c0 innermost); used by the batched fp32 GEMM: (column, row, inner batch, outer batch)
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1,
                                             int32_t c2, int32_t c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// element-wise fp32 add of the staged tile into global memory (performed at L2)
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* smem_src, int32_t c0, int32_t c1,
                                                  int32_t c2, int32_t c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- explicit shared-state-space accesses and named barriers ----
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: arrives on the mbarrier once all previously issued MMAs of this thread retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem];  int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// fp32 operands read as tf32 (low 13 mantissa bits ignored) -> fp32
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem], fp32 operands read as tf32: A = 128 lanes (rows) x 8 consecutive 32-bit columns (K-major)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- thread-block cluster / CTA-pair (cta_group::2) variants ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// arrive on an mbarrier addressed in the shared::cluster window (possibly in the peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (.release.cta) semantics on purpose: a cluster-scope release compiles to MEMBAR.ALL.GPU per arrive.  The
  // data this arrive publishes is read by the local async proxy (tensor core) and was made visible to it by the
  // preceding fence.proxy.async; the remote thread only needs the count.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes are credited to the mbarrier at
// `mbar_cluster_addr` (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t mbar_cluster_addr,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit from the leader CTA, arriving on the same barrier offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- UMMA descriptors ----
// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.
//   K-major  operand: rows of 128 B (one swizzle span of the contraction axis), 8-row groups SBO apart.
//   MN-major operand: atoms of [contraction rows][128 B of the M/N axis]; SBO = 8-row group stride along the
//                     contraction axis, LBO = stride between 128-byte atoms along M/N.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;    // descriptor version (Blackwell)
  d |= 2ull << 61;    // SWIZZLE_128B
  return d;
}

// Same with an explicit layout type: 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks XORed with
// row & 3 - the only swizzled layout of MN-major 32-bit operands; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

// Instruction descriptor (32-bit): c_format [4,6), a_format [7,10), b_format [10,13), a_major 15, b_major 16,
// n>>3 at [17,23), m>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(uint32_t c_fmt, uint32_t a_fmt, uint32_t b_fmt, uint32_t a_mn_major,
                                                  uint32_t b_mn_major, uint32_t m, uint32_t n) {
  return (c_fmt << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) |
         ((m >> 4) << 24);
}
constexpr uint32_t kCFmtF32 = 1, kCFmtS32 = 2, kFmtBF16 = 1, kFmtS8 = 1, kFmtTF32 = 2;

}  // namespace ob
