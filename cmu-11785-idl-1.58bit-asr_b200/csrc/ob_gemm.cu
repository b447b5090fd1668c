// ob_gemm.cu - the tensor-core kernels of the quantised linear layer (sm_100a, tcgen05 + TMEM + TMA).
//
//   gemm_expand_kernel<kFwdI8>   y  = (q . Q^T) * alpha/s + b     int8 x ternary, kind::i8, S32 accumulators
//   gemm_expand_kernel<kDxBf16>  dx = (dys . Q) * alpha*s         bf16 x ternary, kind::f16, F32 accumulators
//        A tile: TMA, SWIZZLE_128B, K-major.  B tile: packed 2-bit codes land in shared memory by TMA, eight
//        expander warps rewrite them as int8 / bf16 in the UMMA K-major SWIZZLE_128B layout (one PRMT per output
//        word), fence.proxy.async, then the single MMA thread consumes them.  Persistent CTAs (optionally CTA pairs,
//        cta_group::2), static tile schedule (n fastest so co-resident CTAs share A rows in L2), double-buffered
//        TMEM accumulators, eight epilogue warps storing through swizzled shared memory with TMA.
//   dw_kernel                     dW_hat partials = dys^T . qb     bf16, both operands MN-major straight from the
//        row-major [tokens, features] tensors (no transposes in HBM), split over tokens; the STE mask, the
//        alpha reduction and grad_bias are fused into the finalizer (ob_quant.cu).
//
// Reference semantics: F.linear / LinearBackward behind onebit_asr/quant.py:126 and _QuantizeSTE.backward :72-92.
#include "ob_common.cuh"

namespace ob {

size_t bwd_tail_workspace_bytes(int N);
int launch_ste_and_tail(const float* g_parts, int splits, const float* W, const float* alpha, int alpha_mode, int64_t n,
                        int bitwidth, float* grad_W, float* grad_alpha, float* alpha_parts, const float* colsum,
                        int n_col_blocks, int N, float* grad_bias, cudaStream_t st);
int launch_dw_finalize_groups(const float* g_parts, int splits_a, int splits, const float* W, const float* alpha, int alpha_mode,
                              int64_t n, int bw_a, int bw_b, float* grad_W, float* grad_alpha, float* alpha_parts,
                              const float* colsum, int n_col_blocks, int N, float* grad_bias, cudaStream_t st);
int* dw_finalize_tickets(float* alpha_parts, int64_t n, int N, int with_bias, int* count);

// ---------------------------------------------------------------------------------------------
// debug / tuning knobs (ob_debug_set)
// ---------------------------------------------------------------------------------------------
enum DebugKey { kDbgSwapLboSbo = 1, kDbgForceBlockN = 2, kDbgForceSplits = 3, kDbgMaxCtas = 4, kDbgKernelFlags = 5,
                kDbgF32SplitMode = 6, kDbgF32Epilogue = 7, kDbgF32Pair = 8, kDbgSmallM = 9, kDbgF32Atm = 11, kDbgPdl = 13, kDbgTailTma = 14 };
int small_m_limit(int K);              // ob_gemv.cu
int small_m_capacity(int K);
int launch_gemv_tern_i8(const int8_t* q, const float* scale, const uint8_t* packed, const float* alpha, int alpha_mode,
                        const float* bias, int M, int N, int K, void* y, int out_bf16, cudaStream_t st);
static int g_dbg_small_m = 0;          // 0: small M takes the DP4A kernel where it wins (ob_gemv.cu), 1: always tcgen05, 2: DP4A whenever it can
void f32_gemm_debug(int split_mode);   // ob_gemm_f32.cu
void f32_gemm_debug_epilogue(int mode);
void f32_gemm_debug_pair(int on);
void f32_gemm_debug_atm(int on);
static int g_dbg_kernel_flags = 0;   // bit0 skip TMA stores, bit1 skip epilogue math/STS, bit2 skip expansion (timing experiments)
static int g_dbg_swap_lbo_sbo = 0;
static int g_dbg_force_block_n = 0;
static int g_dbg_force_splits = 0;
static int g_dbg_max_ctas = 0;
static int g_dbg_tail_tma = 1;        // fused tail: residual by TMA (ob_debug_set key 14, 0 = per-lane row loads)

static int g_sms = 0;
int sm_count() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_dbg_max_ctas > 0 ? g_dbg_max_ctas : g_sms;
}

// ---------------------------------------------------------------------------------------------
// tensor maps (driver entry point resolved at run time: no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major tensor [outer, inner] with `row_bytes` pitch; box = [box_outer, box_inner]
static int make_map(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                    uint64_t row_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not found");
    return OB_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    cudaFree(nullptr);      // thread without a current context (no runtime call yet): bind the primary context, retry
    r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (inner=%llu outer=%llu pitch=%llu box=%ux%u)", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_bytes, box_inner, box_outer);
    return OB_ERR_CUDA;
  }
  return OB_OK;
}

// ---------------------------------------------------------------------------------------------
// A x expand(B) GEMM
// ---------------------------------------------------------------------------------------------
enum GemmMode { kFwdI8 = 0, kDxBf16 = 1 };

constexpr int kBlockM = 128;                               // rows per CTA (UMMA M = 128 x CTAS)
constexpr int kATileBytes = kBlockM * 128;                 // 128 rows x one 128-byte swizzle span
constexpr int kExpandWarps = 8;
constexpr int kExpandThreads = kExpandWarps * 32;
constexpr int kEpiWarp0 = 4 + kExpandWarps;                // first epilogue warp (multiple of 4: TMEM lane quarters)
constexpr int kEpiWarps = 8;                               // two warps per TMEM lane quarter, each takes half of the columns
constexpr int kGemmThreads = (kEpiWarp0 + kEpiWarps) * 32; // 20 warps, see roles below
constexpr int kStageOutBytes = 32 * 128;                   // one epilogue chunk: 32 rows x 128 B

// CTAS = 2: a CTA pair (cluster of 2, cta_group::2) works on a [256 x BLOCK_N] tile; each CTA loads its own 128
// rows of A and expands its own half (BLOCK_N/2 rows) of B, so the expansion work and the shared-memory operand
// traffic per MMA are halved.
template <int MODE, int BLOCK_N, int STAGES, int CTAS, int OUT_BUFS, int RES = 0>
struct GemmSmem {
  static constexpr int kPackedRowBytes = MODE == kFwdI8 ? 32 : 16;   // 128 int8 / 64 bf16 codes per k-block
  static constexpr int kRowsB = BLOCK_N / CTAS;                      // B rows expanded by this CTA
  static constexpr int kBTileBytes = kRowsB * 128;
  static constexpr int kBpTileBytes = kRowsB * kPackedRowBytes;
  static constexpr int kOffA = 0;
  static constexpr int kOffB = kOffA + STAGES * kATileBytes;
  static constexpr int kOffOut = kOffB + STAGES * kBTileBytes;       // 8 warps x OUT_BUFS buffers x 4 KB, 1024-aligned
  static constexpr int kOffRes = kOffOut + kEpiWarps * OUT_BUFS * kStageOutBytes;   // RES: one residual chunk per epilogue warp
  static constexpr int kOffBp = kOffRes + (RES ? kEpiWarps * kStageOutBytes : 0);
  static constexpr int kOffBias = kOffBp + STAGES * kBpTileBytes;    // [2][BLOCK_N] floats
  static constexpr int kOffBar = kOffBias + 2 * BLOCK_N * 4;
  static constexpr int kNumBars = 4 * STAGES + 4 + (RES ? kEpiWarps : 0);
  static constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
  static constexpr int kBytes = kOffTmemSlot + 16;
  static constexpr int kDynBytes = kBytes + 1024;                    // slack for the 1024-byte alignment
};

// Optional module tail fused into the forward epilogue (TAIL = 1, fp32 output only): the layer's output y never reaches HBM,
//   out = resid + (keep ? y * factor * rowmask[row] : 0)      = x + scale * dropout(y) * frame_mask   (conformer.py:41-45, 133-138)
// with the keep bits of the library's Philox stream (element index (row_base + row) * N + col, eight 16-bit lanes per
// block of 8 elements - the same lanes residual_dropout_kernel and the backward prep use).
struct TailArgs {
  const float* resid;       // [M, N] fp32, laid out like the output
  const float* rowmask;     // float validity per GLOBAL row (row_base + row), or nullptr
  float factor;             // scale / (1 - p)  (scale when dropout is off)
  DropRng rng;              // threshold 0 -> no dropout
  long long row_base;
};

// Warp roles: 0 = TMA producer, 1 = MMA issuer (leader CTA only), 2 = TMEM allocator, 3 = idle,
// 4..11 = expanders, 12..19 = epilogue (warp % 4 selects the TMEM lane quarter, (warp-12)/4 the column half).
// OUT_BF16: output element type (0 = fp32, 1 = bf16); the epilogue moves 128 bytes of a row per chunk.
// OUT_BUFS: staging buffers per epilogue warp; the TMA stores of up to OUT_BUFS-1 earlier chunks stay in flight
// (the store-read latency, not the instruction count, bounds the output rate with only two).
// TAIL = 2: the residual chunk ([32 rows x 128 B], the geometry of the output box) arrives by TMA in a per-warp staging buffer
// instead of by per-lane row loads - a lane owns a ROW of the accumulator, so its own loads touch 32 different 128-byte lines per
// instruction and thrash the little L1 that is left beside the operand ring.  The next chunk's residual is requested as soon
// as this one sits in registers, the first one of a tile before the wait for the accumulator.  One staging buffer (OUT_BUFS = 1).
template <int MODE, int BLOCK_N, int STAGES, int OUT_BF16, int CTAS, int OUT_BUFS, int TAIL>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_expand_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bp,
                   const __grid_constant__ CUtensorMap map_out, const float* __restrict__ row_scale,
                   const float* __restrict__ alpha, int alpha_mode, const float* __restrict__ bias, int M, int NC,
                   int KC, int dbg, const TailArgs tail, const __grid_constant__ CUtensorMap map_res) {
  static_assert(TAIL == 0 || (MODE == kFwdI8 && OUT_BF16 == 0), "the fused module tail exists for the fp32 forward only");
  static_assert(TAIL != 2 || OUT_BUFS == 1, "the TMA-fed tail walks one chunk at a time");
  using L = GemmSmem<MODE, BLOCK_N, STAGES, CTAS, OUT_BUFS, TAIL == 2>;
  constexpr int kElemsPerKBlock = MODE == kFwdI8 ? 128 : 64;
  constexpr int kChunkCols = OUT_BF16 ? 64 : 32;          // output columns per 128-byte chunk
  constexpr int kTileM = kBlockM * CTAS;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
  constexpr uint32_t kIdesc = MODE == kFwdI8 ? make_idesc(kCFmtS32, kFmtS8, kFmtS8, 0, 0, kTileM, BLOCK_N)
                                             : make_idesc(kCFmtF32, kFmtBF16, kFmtBF16, 0, 0, kTileM, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  // keep the shared state space visible to the compiler: offset arithmetic on the original pointer
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* a_full_bar = bars;                    // A tiles landed (pair mode: both CTAs' tiles, leader's barrier)
  uint64_t* bp_full_bar = bars + STAGES;          // this CTA's packed B tile landed
  uint64_t* bready_bar = bars + 2 * STAGES;       // expanders wrote the B tile (pair mode: both CTAs', leader's barrier)
  uint64_t* empty_bar = bars + 3 * STAGES;        // MMAs that read the stage retired
  uint64_t* tmem_full_bar = bars + 4 * STAGES;    // [2] accumulator ready for the epilogue
  uint64_t* tmem_empty_bar = bars + 4 * STAGES + 2;   // [2] epilogue drained the accumulator (leader's barrier)
  uint64_t* res_bar = bars + 4 * STAGES + 4;          // TAIL = 2: [kEpiWarps] residual chunk landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int m_tiles = (M + kTileM - 1) / kTileM;
  const int n_blocks = (NC + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_blocks;
  const int num_kb = (KC + kElemsPerKBlock - 1) / kElemsPerKBlock;
  const int first_tile = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_bp);
    tma_prefetch_desc(&map_out);
    if (TAIL == 2) tma_prefetch_desc(&map_res);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&a_full_bar[s], 1);
      mbar_init(&bp_full_bar[s], 1);
      mbar_init(&bready_bar[s], kExpandWarps * CTAS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kEpiWarps * CTAS);
    }
    if (TAIL == 2)
      for (int s = 0; s < kEpiWarps; ++s) mbar_init(&res_bar[s], 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    if (CTAS == 2) { tmem_alloc_pair(tmem_slot, kTmemCols); tmem_relinquish_pair(); }
    else           { tmem_alloc(tmem_slot, kTmemCols);      tmem_relinquish(); }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_entry();       // everything above touched no global memory: it overlaps the previous kernel's tail

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
        const int a_row = m_blk * kTileM + rank * kBlockM;
        const int b_row = n_blk * BLOCK_N + rank * L::kRowsB;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&bp_full_bar[stage], L::kBpTileBytes);
          tma_load_2d(smem + L::kOffBp + stage * L::kBpTileBytes, &map_bp, &bp_full_bar[stage],
                      kb * L::kPackedRowBytes, b_row);
          if (CTAS == 2) {
            if (rank == 0) mbar_expect_tx(&a_full_bar[stage], 2 * kATileBytes);
            tma_load_2d_pair(smem + L::kOffA + stage * kATileBytes, &map_a, mapa_u32(smem_u32(&a_full_bar[stage]), 0),
                             kb * kElemsPerKBlock, a_row);
          } else {
            mbar_expect_tx(&a_full_bar[stage], kATileBytes);
            tma_load_2d(smem + L::kOffA + stage * kATileBytes, &map_a, &a_full_bar[stage], kb * kElemsPerKBlock, a_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (one thread of the leader CTA) ------------------------------
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        if (CTAS == 2) mbar_wait_cluster(&tmem_empty_bar[as], aphase ^ 1); else mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&a_full_bar[stage], phase);
          if (CTAS == 2) mbar_wait_cluster(&bready_bar[stage], phase); else mbar_wait(&bready_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(sbase + L::kOffA + stage * kATileBytes, 0, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(sbase + L::kOffB + stage * L::kBTileBytes, 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) {        // 4 MMAs of 32 contraction bytes each; +32 B = +2 in the address field
            const uint32_t acc = (kb | k) != 0;
            if (MODE == kFwdI8) {
              if (CTAS == 2) umma_i8_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, kIdesc, acc);
              else           umma_i8(d_tmem, a_desc + 2 * k, b_desc + 2 * k, kIdesc, acc);
            } else {
              if (CTAS == 2) umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, kIdesc, acc);
              else           umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, kIdesc, acc);
            }
          }
          if (CTAS == 2) {
            umma_commit_pair(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_pair(&tmem_full_bar[as]);
          } else {
            umma_commit(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit(&tmem_full_bar[as]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < kEpiWarp0) {
    // ------------------------------ expanders: packed 2-bit -> operand tile ------------------------------
    // Thread te owns word column c of rows r0 + k*rows_per_pass: the swizzled chunk offset is loop-invariant.
    const int te = threadIdx.x - 128;
    uint32_t stage = 0, phase = 0;
    uint32_t dst_off;
    if (MODE == kFwdI8) {        // 8 words per row; word c -> 16-byte chunk c
      const int r = te >> 3, c = te & 7;
      dst_off = r * 128 + ((c ^ (r & 7)) << 4);
    } else {                     // 4 words per row; word c -> chunks 2c, 2c+1
      const int r = te >> 2, c = te & 3;
      dst_off = r * 128 + (((2 * c) ^ (r & 7)) << 4);
    }
    const uint32_t bready_addr0 = CTAS == 2 ? mapa_u32(smem_u32(&bready_bar[0]), 0) : smem_u32(&bready_bar[0]);
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&bp_full_bar[stage], phase);
        const uint32_t src = sbase + L::kOffBp + stage * L::kBpTileBytes + te * 4;
        const uint32_t dst = sbase + L::kOffB + stage * L::kBTileBytes + dst_off;
        if (dbg & 4) {
        } else if (MODE == kFwdI8) {
          constexpr int kWords = L::kRowsB * 8;                     // rows advance by 32 per pass
          constexpr int kIters = (kWords + kExpandThreads - 1) / kExpandThreads;
          uint32_t w[kIters];
#pragma unroll
          for (int i = 0; i < kIters; ++i)
            if (kWords % kExpandThreads == 0 || te + i * kExpandThreads < kWords) w[i] = lds32(src + i * kExpandThreads * 4);
#pragma unroll
          for (int i = 0; i < kIters; ++i)
            if (kWords % kExpandThreads == 0 || te + i * kExpandThreads < kWords)
              sts128(dst + i * 32 * 128, expand_word_i8(w[i]));
        } else {
          constexpr int kWords = L::kRowsB * 4;                     // rows advance by 64 per pass
          constexpr int kIters = (kWords + kExpandThreads - 1) / kExpandThreads;
          uint32_t w[kIters];
#pragma unroll
          for (int i = 0; i < kIters; ++i)
            if (kWords % kExpandThreads == 0 || te + i * kExpandThreads < kWords) w[i] = lds32(src + i * kExpandThreads * 4);
#pragma unroll
          for (int i = 0; i < kIters; ++i)
            if (kWords % kExpandThreads == 0 || te + i * kExpandThreads < kWords) {
              uint4 c0, c1;
              expand_word_bf16(w[i], c0, c1);
              sts128(dst + i * 64 * 128, c0);
              sts128((dst + i * 64 * 128) ^ 16u, c1);
            }
        }
        fence_proxy_async_smem();                // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 2) mbar_arrive_cluster(bready_addr0 + stage * 8); else mbar_arrive(&bready_bar[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------ epilogue: TMEM -> registers -> swizzled smem -> TMA store ------------------
    const int ew = warp - kEpiWarp0;                        // 0..7
    const int e = ew & 3, half = ew >> 2;
    const int et = threadIdx.x - kEpiWarp0 * 32;            // 0..255
    const float a_eff = load_alpha_eff(alpha, alpha_mode);
    const uint32_t out_buf = sbase + L::kOffOut + ew * OUT_BUFS * kStageOutBytes;   // [NG groups][G chunks][4 KB]
    const uint32_t out_row = lane * 128;
    const uint32_t swz = (lane & 7) << 4;
    const uint32_t tmem_empty_addr0 = CTAS == 2 ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : smem_u32(&tmem_empty_bar[0]);
    uint32_t buf = 0;
    uint32_t res_phase = 0;                                  // TAIL = 2: parity of this warp's residual barrier
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int row0 = m_blk * kTileM + rank * kBlockM + e * 32;
      const int row = row0 + lane;
      float factor = 0.f, tail_f = 0.f;
      if (row < M) {
        const float s = __ldg(row_scale + row);
        factor = MODE == kFwdI8 ? __fdiv_rn(a_eff, s) : a_eff * s;
        if (TAIL) tail_f = tail.rowmask != nullptr ? tail.factor * __ldg(tail.rowmask + tail.row_base + row) : tail.factor;
      }
      // bias slice of this tile -> smem (double-buffered by accumulator stage; the named barrier below orders it)
      float* bias_s = reinterpret_cast<float*>(smem + L::kOffBias) + as * BLOCK_N;
      for (int j = et; j < BLOCK_N; j += kEpiWarps * 32) {
        const int col = n_blk * BLOCK_N + j;
        bias_s[j] = (bias != nullptr && col < NC) ? __ldg(bias + col) : 0.f;
      }
      named_bar_sync(1, kEpiWarps * 32);
      const uint32_t bias_addr = sbase + L::kOffBias + as * BLOCK_N * 4;
      // chunks are staged and stored in groups of G (one fence / commit per group, NG groups in flight); the two
      // warps of a lane quarter split the tile's chunks in halves
      constexpr int NG = OUT_BUFS >= 2 ? 2 : 1;
      constexpr int G = OUT_BUFS / NG;
      constexpr int kChunks = BLOCK_N / kChunkCols;
      constexpr int kPerHalf = (kChunks + 1) / 2;
      const int c_begin = half * kPerHalf, c_end = min(kChunks, c_begin + kPerHalf);
      if (TAIL == 2 && lane == 0 && row0 < M && n_blk * BLOCK_N + c_begin * kChunkCols < NC) {
        mbar_expect_tx(&res_bar[ew], kStageOutBytes);         // the tile's first residual chunk flies while the MMAs finish
        tma_load_2d(smem + L::kOffRes + ew * kStageOutBytes, &map_res, &res_bar[ew], n_blk * BLOCK_N + c_begin * kChunkCols, row0);
      }
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += G) {
        if (n_blk * BLOCK_N + c0 * kChunkCols >= NC) break;  // warp-uniform
        // the staging group we are about to overwrite must have been read by its TMA stores
        if (lane == 0 && !(dbg & 16)) tma_store_wait_read<NG - 1>();
        __syncwarp();
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int c = c0 + g;
          const int col0 = n_blk * BLOCK_N + c * kChunkCols;
          if (c >= c_end || col0 >= NC) break;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(e * 32) << 16) + as * BLOCK_N + c * kChunkCols;
          const uint32_t obuf = out_buf + (buf * G + g) * kStageOutBytes + out_row;
#pragma unroll
          for (int h = 0; h < (OUT_BF16 ? 2 : 1); ++h) {
            // fused tail: the residual row segment (8 x 16 B, straight from HBM) is requested BEFORE the accumulator is read
            // from TMEM, so its latency overlaps the tcgen05.ld and the previous chunk's stores
            float4 res[TAIL ? 8 : 1];
            if (TAIL == 2) {
              if (row0 < M) {                                      // warp-uniform: rows beyond M have no chunk in flight
                mbar_wait(&res_bar[ew], res_phase);
                res_phase ^= 1;
                const uint32_t rbuf = sbase + L::kOffRes + ew * kStageOutBytes + out_row;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) res[j4] = lds128f(rbuf + ((j4 << 4) ^ swz));
                fence_proxy_async_smem();                          // these reads before the next TMA write into the buffer
                __syncwarp();
                const int ncol0 = col0 + kChunkCols;
                if (lane == 0 && c + 1 < c_end && ncol0 < NC) {
                  mbar_expect_tx(&res_bar[ew], kStageOutBytes);
                  tma_load_2d(smem + L::kOffRes + ew * kStageOutBytes, &map_res, &res_bar[ew], ncol0, row0);
                }
              } else {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) res[j4] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            } else if (TAIL) {
              const float* rrow = tail.resid + static_cast<int64_t>(row) * NC + col0;
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4)
                res[j4] = row < M ? __ldg(reinterpret_cast<const float4*>(rrow) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t r[32];
            tmem_ld_32x32(taddr + h * 32, r);
            // the chunk's bias slice is fetched while the accumulator is on its way (the shared-memory accessors are volatile:
            // inside the store loop every load would wait behind the previous store)
            float4 bvec[TAIL ? 1 : 8];                              // (the tail variants hold the residual chunk in those registers)
            if (!TAIL) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) bvec[j4] = lds128f(bias_addr + (c * kChunkCols + h * 32 + j4 * 4) * 4);
            }
            tmem_ld_wait();
            if (dbg & 2) continue;
            if (TAIL) {
              // 32 columns = four Philox blocks of 8 elements
              const int64_t e0 = (tail.row_base + row) * static_cast<int64_t>(NC) + col0;
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                const uint32_t kb = tail.rng.threshold != 0u
                                        ? philox_keep8(static_cast<unsigned long long>((e0 >> 3) + j8), tail.rng) : 0xFFu;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int j4 = 2 * j8 + u;
                  const float4 b = lds128f(bias_addr + (c * kChunkCols + j4 * 4) * 4);
                  const float4 rr = res[j4];
                  const uint32_t k4 = kb >> (4 * u);
                  float v0 = fmaf(__uint_as_float(r[4 * j4 + 0] + 0x4B400000u) - 12582912.0f, factor, b.x);
                  float v1 = fmaf(__uint_as_float(r[4 * j4 + 1] + 0x4B400000u) - 12582912.0f, factor, b.y);
                  float v2 = fmaf(__uint_as_float(r[4 * j4 + 2] + 0x4B400000u) - 12582912.0f, factor, b.z);
                  float v3 = fmaf(__uint_as_float(r[4 * j4 + 3] + 0x4B400000u) - 12582912.0f, factor, b.w);
                  // explicit mul / add (no contraction): the same bits as residual_dropout_kernel on a stored y
                  v0 = __fadd_rn((k4 & 1u) ? __fmul_rn(v0, tail_f) : 0.f, rr.x);
                  v1 = __fadd_rn((k4 & 2u) ? __fmul_rn(v1, tail_f) : 0.f, rr.y);
                  v2 = __fadd_rn((k4 & 4u) ? __fmul_rn(v2, tail_f) : 0.f, rr.z);
                  v3 = __fadd_rn((k4 & 8u) ? __fmul_rn(v3, tail_f) : 0.f, rr.w);
                  sts128(obuf + ((j4 << 4) ^ swz), make_uint4(__float_as_uint(v0), __float_as_uint(v1),
                                                              __float_as_uint(v2), __float_as_uint(v3)));
                }
              }
              continue;
            }
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b = bvec[TAIL ? 0 : j4];
              float v0, v1, v2, v3;
              if (MODE == kFwdI8) {
                // exact int32 -> fp32 for |acc| < 2^22 (|acc| <= 128 K, K <= 32768) with two full-rate ops
                // instead of the quarter-rate I2F: bits(1.5 * 2^23 + acc) = 0x4B400000 + acc
                v0 = fmaf(__uint_as_float(r[4 * j4 + 0] + 0x4B400000u) - 12582912.0f, factor, b.x);
                v1 = fmaf(__uint_as_float(r[4 * j4 + 1] + 0x4B400000u) - 12582912.0f, factor, b.y);
                v2 = fmaf(__uint_as_float(r[4 * j4 + 2] + 0x4B400000u) - 12582912.0f, factor, b.z);
                v3 = fmaf(__uint_as_float(r[4 * j4 + 3] + 0x4B400000u) - 12582912.0f, factor, b.w);
              } else {
                v0 = fmaf(__uint_as_float(r[4 * j4 + 0]), factor, b.x);
                v1 = fmaf(__uint_as_float(r[4 * j4 + 1]), factor, b.y);
                v2 = fmaf(__uint_as_float(r[4 * j4 + 2]), factor, b.z);
                v3 = fmaf(__uint_as_float(r[4 * j4 + 3]), factor, b.w);
              }
              if (OUT_BF16) {
                // two float4 groups make one 16-byte chunk of 8 bf16: stash the even group, emit on the odd one
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v0, v1), p1 = __floats2bfloat162_rn(v2, v3);
                r[4 * j4 + 0] = *reinterpret_cast<uint32_t*>(&p0);
                r[4 * j4 + 1] = *reinterpret_cast<uint32_t*>(&p1);
                if (j4 & 1) {
                  const int chunk = h * 4 + (j4 >> 1);
                  sts128(obuf + ((chunk << 4) ^ swz),
                         make_uint4(r[4 * (j4 - 1) + 0], r[4 * (j4 - 1) + 1], r[4 * j4 + 0], r[4 * j4 + 1]));
                }
              } else {
                sts128(obuf + ((j4 << 4) ^ swz), make_uint4(__float_as_uint(v0), __float_as_uint(v1),
                                                            __float_as_uint(v2), __float_as_uint(v3)));
              }
            }
          }
        }
        if (!(dbg & 8)) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int col0 = n_blk * BLOCK_N + (c0 + g) * kChunkCols;
            if (c0 + g < c_end && col0 < NC && !(dbg & 1))
              tma_store_2d(&map_out, smem + L::kOffOut + ((ew * NG + buf) * G + g) * kStageOutBytes, col0, row0);
          }
          tma_store_commit();
        }
        buf = (buf + 1 == NG) ? 0 : buf + 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(tmem_empty_addr0 + as * 8); else mbar_arrive(&tmem_empty_bar[as]);
      }
    }
    if (lane == 0) tma_store_wait_all<0>();                  // global writes complete before the CTA retires
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

__device__ __forceinline__ void store_row_chunk_f32(float* dst, const float (&v)[32], int valid_cols) {
  if (valid_cols >= 32) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < valid_cols) dst[j] = v[j];
  }
}

// ---------------------------------------------------------------------------------------------
// dW_hat partials = dys^T . qb   (both operands MN-major from row-major [tokens, features] tensors)
// ---------------------------------------------------------------------------------------------
constexpr int kDwTokBlock = 64;                // tokens per pipeline stage (4 MMAs of 16)
constexpr int kDwAtomBytes = kDwTokBlock * 128;   // [64 tokens][64 features] bf16
constexpr int kDwThreads = 256;

template <int BLOCK_N, int STAGES>
struct DwSmem {
  static constexpr int kATileBytes = 2 * kDwAtomBytes;                 // 128 output rows (layer N)
  static constexpr int kBTileBytes = (BLOCK_N / 64) * kDwAtomBytes;    // BLOCK_N output columns (layer K)
  static constexpr int kOffA = 0;
  static constexpr int kOffB = STAGES * kATileBytes;
  static constexpr int kOffBar = kOffB + STAGES * kBTileBytes;
  static constexpr int kNumBars = 2 * STAGES + 1;
  static constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
  static constexpr int kBytes = kOffTmemSlot + 16;
  static constexpr int kDynBytes = kBytes + 1024;
};

// grid: x = (n_tile * k_tiles + k_tile), y = split.  Warp roles: 0 producer, 1 MMA, 2 TMEM alloc, 4..7 epilogue.
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kDwThreads, 1)
dw_kernel(const __grid_constant__ CUtensorMap map_dys, const __grid_constant__ CUtensorMap map_qb,
          float* __restrict__ partials, int M, int N, int K, int tb_per_split, uint32_t lbo, uint32_t sbo) {
  using L = DwSmem<BLOCK_N, STAGES>;
  constexpr uint32_t kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  constexpr uint32_t kIdesc = make_idesc(kCFmtF32, kFmtBF16, kFmtBF16, 1, 1, 128, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = full_bar + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_tiles = (K + BLOCK_N - 1) / BLOCK_N;
  const int n0 = (blockIdx.x / k_tiles) * 128, k0 = (blockIdx.x % k_tiles) * BLOCK_N;
  const int split = blockIdx.y;
  const int num_tb = (M + kDwTokBlock - 1) / kDwTokBlock;
  const int tb0 = split * tb_per_split;
  const int tb1 = min(num_tb, tb0 + tb_per_split);
  const int nkb = tb1 - tb0;                       // >= 1 by construction of the split count

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dys);
    tma_prefetch_desc(&map_qb);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_entry();       // everything above touched no global memory: it overlaps the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int t0 = (tb0 + kb) * kDwTokBlock;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], L::kATileBytes + L::kBTileBytes);
        uint8_t* a = smem + L::kOffA + stage * L::kATileBytes;
        uint8_t* b = smem + L::kOffB + stage * L::kBTileBytes;
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_2d(a + j * kDwAtomBytes, &map_dys, &full_bar[stage], n0 + 64 * j, t0);
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          tma_load_2d(b + j * kDwAtomBytes, &map_qb, &full_bar[stage], k0 + 64 * j, t0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem + L::kOffA + stage * L::kATileBytes), lbo, sbo);
        const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + L::kOffB + stage * L::kBTileBytes), lbo, sbo);
#pragma unroll
        for (int k = 0; k < kDwTokBlock / 16; ++k)     // 16 tokens = 16 rows x 128 B = 2048 B -> +128 in the address field
          umma_f16(tmem_base, a_desc + 128 * k, b_desc + 128 * k, kIdesc, (kb | k) != 0);
        umma_commit(&empty_bar[stage]);
        if (kb == nkb - 1) umma_commit(done_bar);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int n = n0 + e * 32 + lane;
    float* dst_row = partials + (static_cast<int64_t>(split) * N + n) * K;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(e * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int col0 = k0 + c * 32;
      if (n < N && col0 < K) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        store_row_chunk_f32(dst_row + col0, v, K - col0);
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}


// ---------------------------------------------------------------------------------------------
// dW_hat partials = dys^T . q   on CTA pairs, with the int8 activation codes converted to bf16 in shared memory
// ---------------------------------------------------------------------------------------------
// A pair (cluster of 2, cta_group::2) owns a [256 layer-N rows x 256 layer-K columns] tile of one token split.  Each CTA
// loads its own 128 N-rows of dys (two [64 tokens][64 n] bf16 atoms by TMA, MN-major as in dw_kernel) and its own half
// of the tile's K-columns of q as INT8 ([64 tokens][128 features], 8 KB, un-swizzled); eight converter warps rewrite the
// codes as bf16 (exact) in the MN-major SWIZZLE_128B atom layout.  Against dw_kernel the operand bytes a CTA pulls from
// L2 per token drop from 768 to 384 (the kernel is bound by that stream), and the bf16 copy of q no longer exists in HBM.
// Signalling as in the CTA-pair fp32 GEMM: each CTA's TMA completes on its own `full` barrier, the converter warps of
// both CTAs arrive on the leader's `ready` barrier (the peer's dys tile is covered: its converters waited for the
// barrier that tile completes on), tcgen05.commit multicasts `empty` / `done` to both CTAs.
constexpr int kDw2Threads = 384;               // warps: 0 producer, 1 MMA (leader), 2 TMEM, 3 idle, 4..11 converters + epilogue
constexpr int kDw2ConvWarps = 8;

template <int STAGES>
struct Dw2Smem {
  static constexpr int kATileBytes = 2 * kDwAtomBytes;        // this CTA's 128 N-rows: two [64 tok][64 n] atoms
  static constexpr int kBTileBytes = 2 * kDwAtomBytes;        // this CTA's 128 K-columns, converted
  static constexpr int kQTileBytes = kDwTokBlock * 128;       // [64 tok][128 features] int8
  static constexpr int kOffA = 0;
  static constexpr int kOffB = STAGES * kATileBytes;
  static constexpr int kOffQ = kOffB + STAGES * kBTileBytes;
  static constexpr int kOffBar = kOffQ + STAGES * kQTileBytes;
  static constexpr int kNumBars = 3 * STAGES + 1;
  static constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
  static constexpr int kBytes = kOffTmemSlot + 16;
  static constexpr int kDynBytes = kBytes + 1024;
};

// eight int8 codes -> eight bf16 (exact): u = code + 128 placed in the mantissa of 2^23, minus (2^23 + 128)
__device__ __forceinline__ uint4 i8x8_to_bf16x8(uint2 raw) {
  const uint32_t w0 = raw.x ^ 0x80808080u, w1 = raw.y ^ 0x80808080u;
  float f[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    f[k] = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7540u | k)) - 8388736.0f;
    f[4 + k] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7540u | k)) - 8388736.0f;
  }
  uint4 o;
  __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
  o.x = *reinterpret_cast<uint32_t*>(&p0);
  o.y = *reinterpret_cast<uint32_t*>(&p1);
  o.z = *reinterpret_cast<uint32_t*>(&p2);
  o.w = *reinterpret_cast<uint32_t*>(&p3);
  return o;
}

__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

// Two token groups per launch (the 2-bit and the 1-bit rows of a stacked batch, each with its own tensor maps so that a
// group's last token block is zero-filled instead of running into the other group): splits [0, splits_a) walk group A,
// the rest group B; the finaliser sums the two ranges of partials separately (the alpha term depends on the bitwidth).
// grid: x = 2 * (n_tile * k_tiles + k_tile) + cluster rank, y = split
template <int STAGES>
__global__ void __launch_bounds__(kDw2Threads, 1)
dw_pair_kernel(const __grid_constant__ CUtensorMap map_dys_a, const __grid_constant__ CUtensorMap map_q_a,
               const __grid_constant__ CUtensorMap map_dys_b, const __grid_constant__ CUtensorMap map_q_b,
               float* __restrict__ partials, int M_a, int M_b, int splits_a, int N, int K, int tb_per_split,
               int* __restrict__ fin_tickets, int n_fin_tickets) {
  using L = Dw2Smem<STAGES>;
  constexpr uint32_t kTmemCols = 256;
  constexpr uint32_t kIdesc = make_idesc(kCFmtF32, kFmtBF16, kFmtBF16, 1, 1, 256, 256);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* ready_bar = full_bar + STAGES;         // leader's: converter warps of both CTAs
  uint64_t* empty_bar = full_bar + 2 * STAGES;
  uint64_t* done_bar = full_bar + 3 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int k_tiles = (K + 255) / 256;
  const int tile = blockIdx.x >> 1;
  const int n0 = (tile / k_tiles) * 256 + static_cast<int>(rank) * 128;      // this CTA's rows of the output tile
  const int kt0 = (tile % k_tiles) * 256;                                     // the tile's columns
  const int kq0 = kt0 + static_cast<int>(rank) * 128;                         // the half this CTA converts
  const int split = blockIdx.y;
  const bool in_a = split < splits_a;
  const CUtensorMap* map_dys = in_a ? &map_dys_a : &map_dys_b;
  const CUtensorMap* map_q = in_a ? &map_q_a : &map_q_b;
  const int num_tb = ((in_a ? M_a : M_b) + kDwTokBlock - 1) / kDwTokBlock;
  const int tb0 = (in_a ? split : split - splits_a) * tb_per_split;
  const int nkb = min(num_tb, tb0 + tb_per_split) - tb0;      // >= 1 by construction of the split counts

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(map_dys);
    tma_prefetch_desc(map_q);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], kDw2ConvWarps * 2);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_entry();       // everything above touched no global memory: it overlaps the previous kernel's tail
  // the finaliser that follows this grid counts its blocks in these tickets: zeroed here instead of by a memset node
  if (blockIdx.x == 0 && blockIdx.y == 0 && warp == 3 && lane < n_fin_tickets) {
    for (int i = lane; i < n_fin_tickets; i += 32) fin_tickets[i] = 0;
  }

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int t0 = (tb0 + kb) * kDwTokBlock;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], L::kATileBytes + L::kQTileBytes);
        uint8_t* a = smem + L::kOffA + stage * L::kATileBytes;
#pragma unroll
        for (int j = 0; j < 2; ++j) tma_load_2d(a + j * kDwAtomBytes, map_dys, &full_bar[stage], n0 + 64 * j, t0);
        tma_load_2d(smem + L::kOffQ + stage * L::kQTileBytes, map_q, &full_bar[stage], kq0, t0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_cluster(&ready_bar[stage], phase);
        tc_fence_after();
        const uint64_t a_desc = make_smem_desc_sw128(sbase + L::kOffA + stage * L::kATileBytes, kDwAtomBytes, 1024);
        const uint64_t b_desc = make_smem_desc_sw128(sbase + L::kOffB + stage * L::kBTileBytes, kDwAtomBytes, 1024);
#pragma unroll
        for (int k = 0; k < kDwTokBlock / 16; ++k)     // 16 tokens = 2048 B -> +128 in the address field
          umma_f16_pair(tmem_base, a_desc + 128 * k, b_desc + 128 * k, kIdesc, (kb | k) != 0);
        umma_commit_pair(&empty_bar[stage]);
        if (kb == nkb - 1) umma_commit_pair(done_bar);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ---- converters: [64 tok][128 features] int8 -> two [64 tok][64 features] bf16 atoms (16-byte chunk c of token row t
    //      lands at chunk c ^ (t & 7)); 16 lanes read one token row (128 B), 8 lanes write one atom row ----
    const int tc = threadIdx.x - 128;
    const uint32_t ready_addr0 = mapa_u32(smem_u32(&ready_bar[0]), 0);
    uint32_t stage = 0, phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      const uint32_t src = sbase + L::kOffQ + stage * L::kQTileBytes;
      const uint32_t dst = sbase + L::kOffB + stage * L::kBTileBytes;
      uint2 raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) raw[i] = lds64(src + (tc + 256 * i) * 8);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int task = tc + 256 * i, t = task >> 4, cc = task & 15;
        sts128(dst + (cc >> 3) * kDwAtomBytes + t * 128 + (((cc & 7) ^ (t & 7)) << 4), i8x8_to_bf16x8(raw[i]));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ready_addr0 + stage * 8);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // ---- epilogue: this CTA's 128 rows x 256 columns of the partial product ----
    const int e = warp & 3, half = (warp - 4) >> 2;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int n = n0 + e * 32 + lane;
    float* dst_row = partials + (static_cast<int64_t>(split) * N + n) * K;
#pragma unroll 1
    for (int c = half * 4; c < half * 4 + 4; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(e * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int col0 = kt0 + c * 32;
      if (n < N && col0 < K) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        store_row_chunk_f32(dst_row + col0, v, K - col0);
      }
    }
    tc_fence_before();
  }

  tc_fence_before();
  cluster_sync_all();                 // the peer's shared memory is read until the last MMA retires
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
struct GemmCfg { int ctas, block_n; };

// Candidate tilings, widest first.  The CTA-pair kernel (cta_group::2, [256 x BLOCK_N] tiles) halves the per-CTA
// expansion work and operand traffic; the single-CTA tilings fill the machine on small problems.
static GemmCfg pick_gemm_cfg(int M, int NC) {
  static const GemmCfg cands[5] = {{2, 256}, {2, 128}, {1, 256}, {1, 128}, {1, 64}};
  if (g_dbg_force_block_n > 0) {
    GemmCfg c = {g_dbg_force_block_n >= 1000 ? 2 : 1, g_dbg_force_block_n % 1000};
    if (c.block_n == 64 || c.block_n == 128 || c.block_n == 256) {
      if (c.ctas == 2 && c.block_n == 64) c.ctas = 1;
      return c;
    }
  }
  // cost model: waves x (columns per tile) / relative tensor throughput of the tiling (measured at K = 2048);
  // every tiling gives a CTA 128 rows, so the per-wave cost is proportional to BLOCK_N.
  static const double kRelTput[5] = {1.0, 0.65, 0.85, 0.55, 0.30};
  GemmCfg best = {1, 64};
  double best_cost = 1e30;
  for (int i = 0; i < 5; ++i) {
    const GemmCfg c = cands[i];
    if (c.block_n > 64 && NC <= c.block_n / 2) continue;             // do not pad a narrow output to a wide tile
    const int tile_m = kBlockM * c.ctas;
    const int units = sm_count() / c.ctas;                           // concurrently running tiles
    const int tiles = ((M + tile_m - 1) / tile_m) * ((NC + c.block_n - 1) / c.block_n);
    const int waves = (tiles + units - 1) / units;
    const double cost = static_cast<double>(waves) * c.block_n / kRelTput[i];
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

template <int MODE, int BLOCK_N, int STAGES, int OUT_BF16, int CTAS, int OUT_BUFS, int TAIL>
static int launch_gemm_expand(const CUtensorMap& map_a, const CUtensorMap& map_bp, const CUtensorMap& map_out,
                              const float* row_scale, const float* alpha, int alpha_mode, const float* bias, int M,
                              int NC, int KC, cudaStream_t st, const TailArgs& tail) {
  using L = GemmSmem<MODE, BLOCK_N, STAGES, CTAS, OUT_BUFS, TAIL == 2>;
  CUtensorMap map_res = map_out;                             // TAIL = 2: the residual has the geometry of the output
  if (TAIL == 2) {
    const int rc = make_map(&map_res, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, tail.resid, NC, M, (uint64_t)NC * 4, 32, 32,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != OB_OK) return rc;
  }
  static_assert(L::kDynBytes <= 232448, "shared memory budget exceeded");
  auto kern = gemm_expand_kernel<MODE, BLOCK_N, STAGES, OUT_BF16, CTAS, OUT_BUFS, TAIL>;
  static bool attr_set = false;
  if (!attr_set) {
    OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynBytes));
    attr_set = true;
  }
  const int tile_m = kBlockM * CTAS;
  const int tiles = ((M + tile_m - 1) / tile_m) * ((NC + BLOCK_N - 1) / BLOCK_N);
  const int units = sm_count() / CTAS;
  const int grid = (tiles < units ? tiles : units) * CTAS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = L::kDynBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1] = pdl_attribute();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  OB_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_bp, map_out, row_scale, alpha, alpha_mode, bias, M, NC, KC,
                             g_dbg_kernel_flags, tail, map_res));
  count_launch();
  return OB_OK;
}

static const TailArgs kNoTail = {nullptr, nullptr, 1.0f, {0ull, 0ull, 0u}, 0ll};

template <int MODE, int OUT_BF16, int TAIL = 0>
static int dispatch_gemm_expand(const void* a, const uint8_t* packed, const float* row_scale, const float* alpha,
                                int alpha_mode, const float* bias, void* out, int M, int NC, int KC, cudaStream_t st,
                                const TailArgs& tail = kNoTail) {
  const GemmCfg cfg = pick_gemm_cfg(M, NC);
  const int bn = cfg.block_n;
  CUtensorMap map_a, map_bp, map_out;
  int rc;
  if (MODE == kFwdI8) {
    rc = make_map(&map_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, a, KC, M, (uint64_t)KC, 128, kBlockM, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != OB_OK) return rc;
    rc = make_map(&map_bp, CU_TENSOR_MAP_DATA_TYPE_UINT8, packed, KC / 4, NC, (uint64_t)KC / 4, 32, bn / cfg.ctas,
                  CU_TENSOR_MAP_SWIZZLE_NONE);
  } else {
    rc = make_map(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, a, KC, M, (uint64_t)KC * 2, 64, kBlockM,
                  CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != OB_OK) return rc;
    rc = make_map(&map_bp, CU_TENSOR_MAP_DATA_TYPE_UINT8, packed, KC / 4, NC, (uint64_t)KC / 4, 16, bn / cfg.ctas,
                  CU_TENSOR_MAP_SWIZZLE_NONE);
  }
  if (rc != OB_OK) return rc;
  // output: 32-row x 128-byte boxes written by the epilogue warps (TMA clips the M and N tails)
  if (OUT_BF16)
    rc = make_map(&map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, out, NC, M, (uint64_t)NC * 2, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  else
    rc = make_map(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, out, NC, M, (uint64_t)NC * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != OB_OK) return rc;
#define OB_GEMM_ARGS map_a, map_bp, map_out, row_scale, alpha, alpha_mode, bias, M, NC, KC, st, tail
  if (cfg.ctas == 2) {
    if (bn == 256) {
      // the fused module tail takes its residual by TMA on the 256-wide pair tiles (one staging buffer, four operand stages)
      constexpr int kTail2 = TAIL ? 2 : 0;
      if (TAIL && g_dbg_tail_tma) return launch_gemm_expand<MODE, 256, 4, OUT_BF16, 2, 1, kTail2>(OB_GEMM_ARGS);
      // long contractions want the deeper operand ring; short ones are output-bound and want double-buffered staging
      // (four staging buffers + two operand stages at K <= 256 measured no better: forward equal, grad_x 15-35 % slower)
      if (KC >= 1024) return launch_gemm_expand<MODE, 256, 5, OUT_BF16, 2, 1, TAIL>(OB_GEMM_ARGS);
      return launch_gemm_expand<MODE, 256, 4, OUT_BF16, 2, 2, TAIL>(OB_GEMM_ARGS);
    }
    return launch_gemm_expand<MODE, 128, 6, OUT_BF16, 2, 2, TAIL>(OB_GEMM_ARGS);
  }
  switch (bn) {
    case 256: return launch_gemm_expand<MODE, 256, 3, OUT_BF16, 1, 1, TAIL>(OB_GEMM_ARGS);
    case 128: return launch_gemm_expand<MODE, 128, 4, OUT_BF16, 1, 2, TAIL>(OB_GEMM_ARGS);
    default:  return launch_gemm_expand<MODE, 64, 6, OUT_BF16, 1, 2, TAIL>(OB_GEMM_ARGS);
  }
#undef OB_GEMM_ARGS
}

struct DwPlan {
  int block_n, k_tiles, n_tiles, splits, tb_per_split;
};

static DwPlan plan_dw(int M, int N, int K) {
  DwPlan p;
  p.block_n = K >= 256 ? 256 : (K >= 128 ? 128 : 64);
  p.k_tiles = (K + p.block_n - 1) / p.block_n;
  p.n_tiles = (N + 127) / 128;
  const int num_tb = (M + kDwTokBlock - 1) / kDwTokBlock;
  int want = sm_count() / (p.k_tiles * p.n_tiles);
  if (g_dbg_force_splits > 0) want = g_dbg_force_splits;
  if (want < 1) want = 1;
  if (want > num_tb) want = num_tb;
  p.tb_per_split = (num_tb + want - 1) / want;
  p.splits = (num_tb + p.tb_per_split - 1) / p.tb_per_split;   // no empty split
  return p;
}

// pair kernel: [256 x 256] output tiles, token splits over the CTA pairs; token rows [0, rows_a) and [rows_a, M) are split
// separately with a common number of token blocks per split (no split spans both groups)
struct DwPairPlan {
  int k_tiles, n_tiles, splits_a, splits, tb_per_split;
};

static DwPairPlan plan_dw_pair(int M, int rows_a, int N, int K) {
  DwPairPlan p;
  p.k_tiles = (K + 255) / 256;
  p.n_tiles = (N + 255) / 256;
  const int tb_a = (rows_a + kDwTokBlock - 1) / kDwTokBlock, tb_b = (M - rows_a + kDwTokBlock - 1) / kDwTokBlock;
  int want = (sm_count() / 2) / (p.k_tiles * p.n_tiles);
  if (g_dbg_force_splits > 0) want = g_dbg_force_splits;
  if (want < 1) want = 1;
  if (want > tb_a + tb_b) want = tb_a + tb_b;
  p.tb_per_split = (tb_a + tb_b + want - 1) / want;
  p.splits_a = (tb_a + p.tb_per_split - 1) / p.tb_per_split;
  p.splits = p.splits_a + (tb_b + p.tb_per_split - 1) / p.tb_per_split;
  return p;
}
// upper bound of plan_dw_pair(...).splits over every position of the group boundary: ceil(a/t) + ceil(b/t) <= want + 2
static int max_pair_splits(int M, int N, int K) {
  (void)M;
  int want = (sm_count() / 2) / (((K + 255) / 256) * ((N + 255) / 256));
  if (g_dbg_force_splits > 0) want = g_dbg_force_splits;
  return (want < 1 ? 1 : want) + 2;
}

static int launch_dw_pair(const CUtensorMap& map_dys_a, const CUtensorMap& map_q_a, const CUtensorMap& map_dys_b,
                          const CUtensorMap& map_q_b, float* partials, int M, int rows_a, int N, int K, const DwPairPlan& p,
                          int* fin_tickets, int n_fin_tickets, cudaStream_t st) {
  constexpr int kStages = 5;
  using L = Dw2Smem<kStages>;
  static_assert(L::kDynBytes <= 232448, "shared memory budget exceeded");
  auto kern = dw_pair_kernel<kStages>;
  static bool attr_set = false;
  if (!attr_set) {
    OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynBytes));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * p.n_tiles * p.k_tiles, p.splits);
  cfg.blockDim = dim3(kDw2Threads);
  cfg.dynamicSmemBytes = L::kDynBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1] = pdl_attribute();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  OB_CUDA(cudaLaunchKernelEx(&cfg, kern, map_dys_a, map_q_a, map_dys_b, map_q_b, partials, rows_a, M - rows_a, p.splits_a, N,
                             K, p.tb_per_split, fin_tickets, n_fin_tickets));
  count_launch();
  return OB_OK;
}

template <int BLOCK_N, int STAGES>
static int launch_dw(const CUtensorMap& map_dys, const CUtensorMap& map_qb, float* partials, int M, int N, int K,
                     const DwPlan& p, cudaStream_t st) {
  using L = DwSmem<BLOCK_N, STAGES>;
  static_assert(L::kDynBytes <= 232448, "shared memory budget exceeded");
  auto kern = dw_kernel<BLOCK_N, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynBytes));
    attr_set = true;
  }
  // MN-major SWIZZLE_128B: SBO = 8 contraction rows x 128 B, LBO = one [64 x 64] atom
  uint32_t lbo = kDwAtomBytes, sbo = 1024;
  if (g_dbg_swap_lbo_sbo) { uint32_t t = lbo; lbo = sbo; sbo = t; }
  dim3 grid(p.n_tiles * p.k_tiles, p.splits);
  launch_k((kern), dim3(grid), dim3(kDwThreads), L::kDynBytes, st, map_dys, map_qb, partials, M, N, K, p.tb_per_split, lbo, sbo);
  OB_LAUNCH_CHECK("dw_kernel");
  return OB_OK;
}

}  // namespace ob

// =============================================================================================
// C ABI
// =============================================================================================
using namespace ob;

extern "C" int ob_debug_set(int key, int value) {
  switch (key) {
    case kDbgSwapLboSbo: g_dbg_swap_lbo_sbo = value; return OB_OK;
    case kDbgForceBlockN: g_dbg_force_block_n = value; return OB_OK;
    case kDbgForceSplits: g_dbg_force_splits = value; return OB_OK;
    case kDbgMaxCtas: g_dbg_max_ctas = value; return OB_OK;
    case kDbgKernelFlags: g_dbg_kernel_flags = value; return OB_OK;
    case kDbgF32SplitMode: f32_gemm_debug(value); return OB_OK;
    case kDbgF32Epilogue: f32_gemm_debug_epilogue(value); return OB_OK;
    case kDbgF32Pair: f32_gemm_debug_pair(value); return OB_OK;
    case kDbgSmallM: g_dbg_small_m = value; return OB_OK;
    case kDbgF32Atm: f32_gemm_debug_atm(value); return OB_OK;
    case kDbgPdl: set_pdl(value); return OB_OK;
    case kDbgTailTma: g_dbg_tail_tma = value; return OB_OK;
    default: set_error("ob_debug_set: unknown key %d", key); return OB_ERR_ARG;
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int ob_gemm_tern_i8_fwd(const int8_t* q, const float* scale, const uint8_t* packed_i8, const float* alpha,
                                   int alpha_mode, const float* bias, int M, int N, int K, void* y, int y_dtype,
                                   ob_stream_t stream) {
  OB_REQUIRE(q && scale && packed_i8 && alpha && y, "ob_gemm_tern_i8_fwd: null pointer");
  OB_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 64 == 0 && K <= 32768,
             "ob_gemm_tern_i8_fwd: need K %% 64 == 0, K <= 32768 and N %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
  OB_REQUIRE(aligned16(q) && aligned16(packed_i8) && aligned16(y), "ob_gemm_tern_i8_fwd: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != OB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OB_REQUIRE(y_dtype == OB_F32 || y_dtype == OB_BF16, "ob_gemm_tern_i8_fwd: unknown dtype tag %d", y_dtype);
  // small-batch regime: weight-streaming DP4A kernel instead of a mostly empty 128-row UMMA tile (same epilogue bits)
  if (g_dbg_small_m != 1 && M <= (g_dbg_small_m == 2 ? small_m_capacity(K) : small_m_limit(K)))
    return launch_gemv_tern_i8(q, scale, packed_i8, alpha, alpha_mode, bias, M, N, K, y, y_dtype == OB_BF16, st);
  if (y_dtype == OB_F32)
    return dispatch_gemm_expand<kFwdI8, 0>(q, packed_i8, scale, alpha, alpha_mode, bias, y, M, N, K, st);
  if (y_dtype == OB_BF16)
    return dispatch_gemm_expand<kFwdI8, 1>(q, packed_i8, scale, alpha, alpha_mode, bias, y, M, N, K, st);
  OB_REQUIRE(false, "ob_gemm_tern_i8_fwd: unknown dtype tag %d", y_dtype);
}

extern "C" int ob_gemm_tern_i8_fwd_tail(const int8_t* q, const float* scale, const uint8_t* packed_i8, const float* alpha,
                                        int alpha_mode, const float* bias, int M, int N, int K, const float* resid,
                                        const float* rowmask, float tail_factor, uint64_t seed, uint64_t offset,
                                        uint32_t drop_threshold, int64_t row_base, float* out, ob_stream_t stream) {
  OB_REQUIRE(q && scale && packed_i8 && alpha && resid && out, "ob_gemm_tern_i8_fwd_tail: null pointer");
  OB_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 64 == 0 && K <= 32768 && row_base >= 0,
             "ob_gemm_tern_i8_fwd_tail: need K %% 64 == 0, K <= 32768 and N %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
  OB_REQUIRE(drop_threshold < 65536u, "ob_gemm_tern_i8_fwd_tail: drop_threshold (%u) is a 16-bit value", drop_threshold);
  OB_REQUIRE(aligned16(q) && aligned16(packed_i8) && aligned16(out) && aligned16(resid),
             "ob_gemm_tern_i8_fwd_tail: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != OB_OK) return rc;
  const TailArgs tail = {resid, rowmask, tail_factor, {seed, offset, drop_threshold}, static_cast<long long>(row_base)};
  return dispatch_gemm_expand<kFwdI8, 0, 1>(q, packed_i8, scale, alpha, alpha_mode, bias, out, M, N, K,
                                            static_cast<cudaStream_t>(stream), tail);
}

extern "C" int ob_bwd_dx(const void* dys_bf16, const float* scale, const uint8_t* packed_t, const float* alpha,
                         int alpha_mode, int M, int N, int K, void* dx, int dx_dtype, ob_stream_t stream) {
  OB_REQUIRE(dys_bf16 && scale && packed_t && alpha && dx, "ob_bwd_dx: null pointer");
  OB_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 64 == 0,
             "ob_bwd_dx: need K %% 64 == 0 and N %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
  OB_REQUIRE(aligned16(dys_bf16) && aligned16(packed_t) && aligned16(dx), "ob_bwd_dx: pointers must be 16-byte aligned");
  int rc = check_device();
  if (rc != OB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // contraction runs over the layer's N; the output has K columns
  if (dx_dtype == OB_F32)
    return dispatch_gemm_expand<kDxBf16, 0>(dys_bf16, packed_t, scale, alpha, alpha_mode, nullptr, dx, M, K, N, st);
  if (dx_dtype == OB_BF16)
    return dispatch_gemm_expand<kDxBf16, 1>(dys_bf16, packed_t, scale, alpha, alpha_mode, nullptr, dx, M, K, N, st);
  OB_REQUIRE(false, "ob_bwd_dx: unknown dtype tag %d", dx_dtype);
}

extern "C" size_t ob_bwd_dw_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const DwPlan p = plan_dw(M, N, K);
  const int s2 = max_pair_splits(M, N, K);
  const size_t partials = (size_t)(p.splits > s2 ? p.splits : s2) * N * K * sizeof(float);
  return partials + ob_ste_workspace_bytes((int64_t)N * K) + bwd_tail_workspace_bytes(N) + 256;
}

extern "C" int ob_bwd_dw(const void* dys_bf16, const void* qb_bf16, const float* colsum, const float* W,
                         const float* alpha, int alpha_mode, int bitwidth, int M, int N, int K, float* grad_W,
                         float* grad_alpha, float* grad_bias, void* ws, size_t ws_bytes, ob_stream_t stream) {
  OB_REQUIRE(dys_bf16 && qb_bf16 && W && alpha && grad_W && grad_alpha && ws, "ob_bwd_dw: null pointer");
  OB_REQUIRE((grad_bias == nullptr) || (colsum != nullptr), "ob_bwd_dw: grad_bias requested without colsum partials");
  OB_REQUIRE(bitwidth == 1 || bitwidth == 2, "bitwidth must be one of {1,2,32}");
  OB_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 64 == 0,
             "ob_bwd_dw: need K %% 64 == 0 and N %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
  OB_REQUIRE(aligned16(dys_bf16) && aligned16(qb_bf16) && aligned16(W) && aligned16(grad_W) && aligned16(ws),
             "ob_bwd_dw: pointers must be 16-byte aligned");
  if (ws_bytes < ob_bwd_dw_workspace_bytes(M, N, K)) {
    set_error("ob_bwd_dw: workspace too small (%zu < %zu)", ws_bytes, ob_bwd_dw_workspace_bytes(M, N, K));
    return OB_ERR_WORKSPACE;
  }
  int rc = check_device();
  if (rc != OB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DwPlan p = plan_dw(M, N, K);
  CUtensorMap map_dys, map_qb;
  rc = make_map(&map_dys, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dys_bf16, N, M, (uint64_t)N * 2, 64, kDwTokBlock,
                CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != OB_OK) return rc;
  rc = make_map(&map_qb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, qb_bf16, K, M, (uint64_t)K * 2, 64, kDwTokBlock,
                CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != OB_OK) return rc;
  float* partials = static_cast<float*>(ws);
  float* alpha_parts = partials + (size_t)p.splits * N * K;
  switch (p.block_n) {
    case 256: rc = launch_dw<256, 4>(map_dys, map_qb, partials, M, N, K, p, st); break;
    case 128: rc = launch_dw<128, 6>(map_dys, map_qb, partials, M, N, K, p, st); break;
    default:  rc = launch_dw<64, 8>(map_dys, map_qb, partials, M, N, K, p, st); break;
  }
  if (rc != OB_OK) return rc;
  return launch_ste_and_tail(partials, p.splits, W, alpha, alpha_mode, (int64_t)N * K, bitwidth, grad_W, grad_alpha,
                             alpha_parts, colsum, ob_bwd_colsum_blocks(M), N, grad_bias, st);
}

// rows [0, rows2) of dys / q belong to the 2-bit pass, rows [rows2, M) to the 1-bit pass (rows2 = M or 0: one group)
static int bwd_dw_q8_groups(const char* who, const void* dys_bf16, const int8_t* q, const float* colsum, const float* W,
                            const float* alpha, int alpha_mode, int rows2, int M, int N, int K, float* grad_W,
                            float* grad_alpha, float* grad_bias, void* ws, size_t ws_bytes, ob_stream_t stream) {
  OB_REQUIRE(dys_bf16 && q && W && alpha && grad_W && grad_alpha && ws, "%s: null pointer", who);
  OB_REQUIRE((grad_bias == nullptr) || (colsum != nullptr), "%s: grad_bias requested without colsum partials", who);
  OB_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 64 == 0, "%s: need K %% 64 == 0 and N %% 64 == 0 (M=%d N=%d K=%d)",
             who, M, N, K);
  OB_REQUIRE(rows2 >= 0 && rows2 <= M, "%s: rows2 (%d) outside [0, M = %d]", who, rows2, M);
  OB_REQUIRE(aligned16(dys_bf16) && aligned16(q) && aligned16(W) && aligned16(grad_W) && aligned16(ws),
             "%s: pointers must be 16-byte aligned", who);
  if (ws_bytes < ob_bwd_dw_workspace_bytes(M, N, K)) {
    set_error("%s: workspace too small (%zu < %zu)", who, ws_bytes, ob_bwd_dw_workspace_bytes(M, N, K));
    return OB_ERR_WORKSPACE;
  }
  int rc = check_device();
  if (rc != OB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DwPairPlan p = plan_dw_pair(M, rows2, N, K);
  // an empty group gets the other group's maps (never dereferenced: it has no splits)
  const int ra = rows2 > 0 ? rows2 : M, rb = M - rows2 > 0 ? M - rows2 : M;
  const int off_b = M - rows2 > 0 ? rows2 : 0;
  const uint8_t* dys_b = static_cast<const uint8_t*>(dys_bf16) + (size_t)off_b * N * 2;
  CUtensorMap map_dys_a, map_q_a, map_dys_b, map_q_b;
  rc = make_map(&map_dys_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dys_bf16, N, ra, (uint64_t)N * 2, 64, kDwTokBlock,
                CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != OB_OK) return rc;
  rc = make_map(&map_q_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, q, K, ra, (uint64_t)K, 128, kDwTokBlock, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != OB_OK) return rc;
  rc = make_map(&map_dys_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dys_b, N, rb, (uint64_t)N * 2, 64, kDwTokBlock,
                CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != OB_OK) return rc;
  rc = make_map(&map_q_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, q + (size_t)off_b * K, K, rb, (uint64_t)K, 128, kDwTokBlock,
                CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != OB_OK) return rc;
  float* partials = static_cast<float*>(ws);
  float* alpha_parts = partials + (size_t)p.splits * N * K;
  int n_tickets = 0;
  int* tickets = dw_finalize_tickets(alpha_parts, (int64_t)N * K, N, grad_bias != nullptr, &n_tickets);
  rc = launch_dw_pair(map_dys_a, map_q_a, map_dys_b, map_q_b, partials, M, rows2, N, K, p, tickets, n_tickets, st);
  if (rc != OB_OK) return rc;
  return launch_dw_finalize_groups(partials, p.splits_a, p.splits, W, alpha, alpha_mode, (int64_t)N * K, 2, 1, grad_W, grad_alpha,
                                   alpha_parts, colsum, ob_bwd_colsum_blocks(M), N, grad_bias, st);
}

extern "C" int ob_bwd_dw_q8(const void* dys_bf16, const int8_t* q, const float* colsum, const float* W,
                            const float* alpha, int alpha_mode, int bitwidth, int M, int N, int K, float* grad_W,
                            float* grad_alpha, float* grad_bias, void* ws, size_t ws_bytes, ob_stream_t stream) {
  OB_REQUIRE(bitwidth == 1 || bitwidth == 2, "bitwidth must be one of {1,2,32}");
  return bwd_dw_q8_groups("ob_bwd_dw_q8", dys_bf16, q, colsum, W, alpha, alpha_mode, bitwidth == 2 ? M : 0, M, N, K, grad_W,
                          grad_alpha, grad_bias, ws, ws_bytes, stream);
}

extern "C" int ob_bwd_dw_q8_groups(const void* dys_bf16, const int8_t* q, const float* colsum, const float* W,
                                   const float* alpha, int alpha_mode, int rows2, int M, int N, int K, float* grad_W,
                                   float* grad_alpha, float* grad_bias, void* ws, size_t ws_bytes, ob_stream_t stream) {
  return bwd_dw_q8_groups("ob_bwd_dw_q8_groups", dys_bf16, q, colsum, W, alpha, alpha_mode, rows2, M, N, K, grad_W, grad_alpha,
                          grad_bias, ws, ws_bytes, stream);
}
