// ob_gemm_f32.cu - batched fp32 GEMM on the tcgen05 tensor cores with fp32-level accuracy ("3xTF32").
//
//   D[b](m, n) (+)= scale * sum_k A[b](m, k) * B[b](n, k) + bias[n]          fp32 in, fp32 out, fp32 accumulation in TMEM
//
// The non-routed matmuls around the quantised layer (the attention products of
// conformer.py:113-129, the vocabulary projections, the 1x1 convolutions) are fp32 in the reference.  The tensor cores
// read fp32 operands as tf32 (10 explicit mantissa bits), so each operand tile is split in shared memory into
//   hi = round_to_tf32(x)   and   lo = x - hi   (exact in fp32, |lo| <= 2^-11 |x|)
// by eight splitter warps, and the single MMA thread issues  lo.hi + hi.lo + hi.hi  per k-step: the dropped terms are
// O(2^-22) relative, i.e. the result carries fp32-level error instead of the 2^-11 of a plain tf32 product.
// PASSES = 1 is the plain tf32 product (no split), kept for comparison.
//
// Structure (same skeleton as gemm_expand_kernel): persistent CTAs, static tile schedule over (batch, m, n) with n
// fastest; warp 0 TMA producer (4-D tensor maps: column, row, inner batch, outer batch - strided views such as the
// [B, T, H, d] projections are read in place, broadcast operands have a zero batch stride), warp 1 MMA issuer, warp 2
// TMEM allocator, warps 4-11 splitters, warps 12-19 epilogue (TMEM -> registers -> per-warp swizzled staging -> TMA box
// store; coalesced row stores / vector atomic adds from the same staging buffer for split-K partials and accumulation).  Operands may be K-major (contraction axis contiguous) or MN-major (the M / N
// axis contiguous), so transposed products need no transposes in HBM.  Edges: TMA zero-fills out-of-range loads and
// clips out-of-range stores, so M, N, K are arbitrary (leading dimensions must be multiples of 4 elements).
#include "ob_common.cuh"

namespace ob {

constexpr int kFTileM = 128;
constexpr int kFKBlock = 32;                       // fp32 elements per k-block: 128 B = one swizzle span
constexpr int kFATileBytes = kFTileM * 128;
constexpr int kFSplitWarps = 8;
constexpr int kFSplitThreads = kFSplitWarps * 32;
constexpr int kFEpiWarp0 = 4 + kFSplitWarps;
constexpr int kFEpiWarps = 8;
constexpr int kFThreads = (kFEpiWarp0 + kFEpiWarps) * 32;
constexpr int kFStageOutBytes = 32 * 128;          // one epilogue chunk: 32 rows x 32 fp32

// CTAS = 2: a CTA pair (cluster of 2, cta_group::2) works on a [256 x BLOCK_N] tile.  Each CTA loads and splits its own
// 128 rows of A and its own half (BLOCK_N / 2 rows) of B, so per output element the splitter traffic and the tensor
// core's operand reads of B are halved - the shared-memory pipe is what limits the deep shapes.
// ATM = 1 ("A through TMEM", single CTA, PASSES = 3): the splitter warps read the raw A tile from shared memory once and
// write its hi / lo parts into tensor memory (128 lanes = rows, 32 + 32 columns per stage); the MMAs take A from TMEM, so the
// three passes of a k-step no longer read A from shared memory and no lo copy of A is written there.  This is what binds
// the attention products (a [T x T] operand streamed against a 64-wide one): 144 KB -> 80 KB of shared-memory traffic per
// k-block.  The A tile needs no UMMA layout any more: a transposed (MN-major) A is read from un-swizzled [32 k][32 m] boxes.
template <int BLOCK_N, int STAGES, int PASSES, int CTAS = 1, int ATM = 0>
struct F32Smem {
  static constexpr int kRowsB = BLOCK_N / CTAS;                        // B rows held by this CTA
  static constexpr int kBTileBytes = kRowsB * 128;
  static constexpr int kHiBytes = kFATileBytes + kBTileBytes;          // [A_hi | B_hi], then [A_lo | B_lo] behind it
  // ATM: a stage of the (deep) raw ring is just the TMA-written [A | B]; what the splitters produce - A hi / lo in tensor
  // memory, B lo in shared memory - lives in a short second ring of kProc stages, so that the bytes in flight from HBM / L2
  // are not limited by the space the split copies take (these products are bound by exactly that latency x depth).
  static constexpr int kProc = ATM ? (BLOCK_N <= 64 ? 3 : 2) : 0;
  static constexpr int kStageBytes = ATM ? kHiBytes : (PASSES == 3 ? 2 : 1) * kHiBytes;
  static constexpr int kOffBLo = kHiBytes + kFATileBytes;              // B_lo within a stage (not ATM)
  static constexpr int kOffBLoRing = STAGES * kStageBytes;             // ATM: kProc x B_lo
  static constexpr int kOffOut = kOffBLoRing + kProc * kBTileBytes;    // 8 epilogue warps x 4 KB
  static constexpr int kOffBar = kOffOut + kFEpiWarps * kFStageOutBytes;
  static constexpr int kNumBars = 3 * STAGES + 4 + 2 * kProc;          // full, ready, empty, 4 accumulator barriers, ATM: pready, pfree
  static constexpr int kOffTmemSlot = kOffBar + kNumBars * 8;
  static constexpr int kBytes = kOffTmemSlot + 16;
  static constexpr int kDynBytes = kBytes + 1024;
  static constexpr int kTmemACol = 2 * BLOCK_N;                        // ATM: processed stage s holds A_hi at +64 s, A_lo at +64 s + 32
};

struct F32Params {
  const float* bias;      // [N] or null
  float scale;
  int accumulate;         // 1: D += ... (vector atomic adds)
  int M, N, K;
  int nb0, nb1;           // batch = nb0 * nb1 (outer, inner)
  int a_b0, a_b1, b_b0, b_b1;   // 1 = the operand has this batch axis, 0 = broadcast (coordinate 0)
  int split_mode;         // 0: hi = rna(x) rewritten in place; 1: hi tile left raw (tensor core truncates), lo = x - trunc(x)
  int k_splits, kb_per_split;   // split-K (batch == 1 only): split s stores its partial product in part[s] (pitch ldp),
  float* part;                  // summed in fixed order by splitk_reduce_kernel
  int64_t ldp;
  float* D;                     // output, row pitch ldd, batch strides in elements
  int64_t ldd, d_bs0, d_bs1;
  int tma_out;                  // 1: plain stores of full chunks go out as TMA box stores from the staging buffer
  int d_b0, d_b1;               // batch coordinates of the output map
};

template <int A_MN, int B_MN, int BLOCK_N, int STAGES, int PASSES, int CTAS = 1, int ATM = 0>
__global__ void __launch_bounds__(kFThreads, 1)
f32_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_d, const F32Params p) {
  using L = F32Smem<BLOCK_N, STAGES, PASSES, CTAS, ATM>;
  static_assert(CTAS == 1 || (CTAS == 2 && PASSES == 3), "the CTA-pair variant is built for the split product only");
  static_assert(!ATM || (CTAS == 1 && PASSES == 3 && 2 * BLOCK_N + 64 * L::kProc <= 512),
                "A through TMEM: single CTA, split product, accumulators + A stages within the 512 columns");
  constexpr int kTileM = kFTileM * CTAS;
  constexpr uint32_t kTmemCols = ATM ? 512 : (2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N);
  constexpr uint32_t kIdesc = make_idesc(kCFmtF32, kFmtTF32, kFmtTF32, ATM ? 0 : A_MN, B_MN, kTileM, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* full_bar = bars;                      // this CTA's operand tiles landed (TMA)
  uint64_t* ready_bar = bars + STAGES;            // splitters wrote hi / lo (pair: both CTAs' splitters, leader's barrier)
  uint64_t* empty_bar = bars + 2 * STAGES;        // MMAs that read the stage retired (pair: multicast to both CTAs)
  uint64_t* tmem_full_bar = bars + 3 * STAGES;    // [2]
  uint64_t* tmem_empty_bar = bars + 3 * STAGES + 2;   // (pair: both CTAs' epilogue warps, leader's barrier)
  uint64_t* pready_bar = bars + 3 * STAGES + 4;       // ATM: splitters filled processed stage (TMEM A hi / lo, B lo)
  uint64_t* pfree_bar = pready_bar + L::kProc;        // ATM: the MMAs that read it retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int m_tiles = (p.M + kTileM - 1) / kTileM;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles_mn = m_tiles * n_tiles;
  const int tiles_per_batch = tiles_mn * p.k_splits;               // split index outermost within a batch item
  const int num_tiles = tiles_per_batch * p.nb0 * p.nb1;
  const int total_kb = (p.K + kFKBlock - 1) / kFKBlock;
  const int first_tile = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;   // the CTAs of a pair walk the same tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_d);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], kFSplitWarps * CTAS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kFEpiWarps * CTAS);
    }
    for (int s = 0; s < L::kProc; ++s) {
      mbar_init(&pready_bar[s], kFSplitWarps);
      mbar_init(&pfree_bar[s], 1);
    }
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
        const int batch = tile / tiles_per_batch, ts = tile - batch * tiles_per_batch;
        const int split = ts / tiles_mn, t = ts - split * tiles_mn;
        const int m0 = (t / n_tiles) * kTileM + rank * kFTileM, n0 = (t % n_tiles) * BLOCK_N + rank * L::kRowsB;
        const int b0 = batch / p.nb1, b1 = batch - b0 * p.nb1;
        const int kb0 = split * p.kb_per_split, num_kb = min(p.kb_per_split, total_kb - kb0);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int k0 = (kb0 + kb) * kFKBlock;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], L::kHiBytes);
          uint8_t* a = smem + stage * L::kStageBytes;
          uint8_t* b = a + kFATileBytes;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < kFTileM / 32; ++j)       // atoms of [32 k-rows][32 m] = 4 KB
              tma_load_4d(a + j * 4096, &map_a, &full_bar[stage], m0 + 32 * j, k0, b1 * p.a_b1, b0 * p.a_b0);
          } else {
            tma_load_4d(a, &map_a, &full_bar[stage], k0, m0, b1 * p.a_b1, b0 * p.a_b0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < L::kRowsB / 32; ++j)
              tma_load_4d(b + j * 4096, &map_b, &full_bar[stage], n0 + 32 * j, k0, b1 * p.b_b1, b0 * p.b_b0);
          } else {
            tma_load_4d(b, &map_b, &full_bar[stage], k0, n0, b1 * p.b_b1, b0 * p.b_b0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0 && rank == 0) {                 // pair: one thread of the leader CTA issues for both
      uint32_t stage = 0, phase = 0;
      uint32_t ps = 0, pphase = 0;                // ATM: cursor over the processed ring
      int it = 0;
      // K-major: SWIZZLE_128B, 8-row groups 1024 B apart.  MN-major fp32: SWIZZLE_128B_BASE32B, LBO = stride between the
      // 128-byte column blocks along M/N (one [32 k x 32 mn] TMA box), SBO = 4 contraction rows of 128 B
      constexpr uint32_t kLboA = A_MN ? 4096 : 0, kLboB = B_MN ? 4096 : 0;
      constexpr uint32_t kSboA = A_MN ? 512 : 1024, kSboB = B_MN ? 512 : 1024;
      constexpr uint32_t kLayA = A_MN ? 1 : 2, kLayB = B_MN ? 1 : 2;
      constexpr uint32_t kStepA = A_MN ? 64 : 2, kStepB = B_MN ? 64 : 2;   // 8 contraction elements, in 16-byte units
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
        const uint32_t as = it & 1, aphase = (it >> 1) & 1;
        if (CTAS == 2) mbar_wait_cluster(&tmem_empty_bar[as], aphase ^ 1); else mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        const int split = (tile % tiles_per_batch) / tiles_mn;
        const int num_kb = min(p.kb_per_split, total_kb - split * p.kb_per_split);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (ATM) {
            mbar_wait(&pready_bar[ps], pphase);
            tc_fence_after();
            const uint32_t b_addr = sbase + stage * L::kStageBytes + kFATileBytes;
            const uint64_t b_hi = make_smem_desc(b_addr, kLboB, kSboB, kLayB);
            const uint64_t b_lo = make_smem_desc(sbase + L::kOffBLoRing + ps * L::kBTileBytes, kLboB, kSboB, kLayB);
            const uint32_t a_hi_t = tmem_base + L::kTmemACol + ps * 64, a_lo_t = a_hi_t + 32;
#pragma unroll
            for (int k = 0; k < kFKBlock / 8; ++k) {
              umma_tf32_ts(d_tmem, a_lo_t + 8 * k, b_hi + kStepB * k, kIdesc, (kb | k) != 0);
              umma_tf32_ts(d_tmem, a_hi_t + 8 * k, b_lo + kStepB * k, kIdesc, 1);
              umma_tf32_ts(d_tmem, a_hi_t + 8 * k, b_hi + kStepB * k, kIdesc, 1);
            }
            umma_commit(&empty_bar[stage]);
            umma_commit(&pfree_bar[ps]);
            if (kb == num_kb - 1) umma_commit(&tmem_full_bar[as]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            if (++ps == L::kProc) { ps = 0; pphase ^= 1; }
            continue;
          }
          // pair: the peer's TMA-written hi tile is covered by its splitters, which waited for it before arriving here
          if (CTAS == 2) mbar_wait_cluster(&ready_bar[stage], phase);
          else           mbar_wait(PASSES == 3 ? &ready_bar[stage] : &full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = sbase + stage * L::kStageBytes, b_addr = a_addr + kFATileBytes;
          const uint64_t a_hi = make_smem_desc(a_addr, kLboA, kSboA, kLayA);
          const uint64_t b_hi = make_smem_desc(b_addr, kLboB, kSboB, kLayB);
          const uint64_t a_lo = make_smem_desc(a_addr + L::kHiBytes, kLboA, kSboA, kLayA);
          const uint64_t b_lo = make_smem_desc(a_addr + L::kOffBLo, kLboB, kSboB, kLayB);
#pragma unroll
          for (int k = 0; k < kFKBlock / 8; ++k) {
            const uint32_t acc = (kb | k) != 0;
            if (CTAS == 2) {
              umma_tf32_pair(d_tmem, a_lo + kStepA * k, b_hi + kStepB * k, kIdesc, acc);
              umma_tf32_pair(d_tmem, a_hi + kStepA * k, b_lo + kStepB * k, kIdesc, 1);
              umma_tf32_pair(d_tmem, a_hi + kStepA * k, b_hi + kStepB * k, kIdesc, 1);
            } else if (PASSES == 3) {
              umma_tf32(d_tmem, a_lo + kStepA * k, b_hi + kStepB * k, kIdesc, acc);
              umma_tf32(d_tmem, a_hi + kStepA * k, b_lo + kStepB * k, kIdesc, 1);
              umma_tf32(d_tmem, a_hi + kStepA * k, b_hi + kStepB * k, kIdesc, 1);
            } else {
              umma_tf32(d_tmem, a_hi + kStepA * k, b_hi + kStepB * k, kIdesc, acc);
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
  } else if (warp >= 4 && warp < kFEpiWarp0) {
    // ------------------------------ splitters: x -> (hi, lo), element-wise, layout-agnostic ------------------------------
    if (ATM) {
      // A: thread = (row, half of the k-block); raw words are the hi operand (the tensor core truncates), lo = x - trunc(x)
      const int te = threadIdx.x - 128;
      const int quarter = warp & 3, khalf = (warp - 4) >> 2;       // TMEM lane quarter of this warp, columns [16 khalf, +16)
      const int row = quarter * 32 + lane;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + L::kTmemACol + 16 * khalf;
      constexpr int kBChunks = L::kBTileBytes / 16;
      static_assert(kBChunks % kFSplitThreads == 0, "B tile must divide over the splitter threads");
      uint32_t stage = 0, phase = 0, ps = 0, pphase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int split = (tile % tiles_per_batch) / tiles_mn;
        const int num_kb = min(p.kb_per_split, total_kb - split * p.kb_per_split);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          const uint32_t a_smem = sbase + stage * L::kStageBytes;
          uint32_t x[16], l[16];
          if (A_MN) {                                  // un-swizzled atoms [32 k][32 m]: lanes read 128 contiguous bytes
            const uint32_t src = a_smem + quarter * 4096 + lane * 4 + (16 * khalf) * 128;
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = lds32(src + j * 128);
          } else {                                     // K-major rows of 128 B, 16-byte chunk c at c ^ (row & 7)
            const uint32_t src = a_smem + row * 128;
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const uint4 v = lds128(src + (((4 * khalf + j4) ^ (row & 7)) << 4));
              x[4 * j4] = v.x, x[4 * j4 + 1] = v.y, x[4 * j4 + 2] = v.z, x[4 * j4 + 3] = v.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j)
            l[j] = __float_as_uint(__uint_as_float(x[j]) - __uint_as_float(x[j] & 0xFFFFE000u));
          mbar_wait(&pfree_bar[ps], pphase ^ 1);       // the MMAs that read this processed stage have retired
          tc_fence_after();
          tmem_st_32x16(t_row + ps * 64, x);
          tmem_st_32x16(t_row + ps * 64 + 32, l);
          // B: the raw tile is the hi operand, lo goes to the processed ring (element-wise, layout-agnostic)
          const uint32_t bh = a_smem + kFATileBytes + te * 16, bl = sbase + L::kOffBLoRing + ps * L::kBTileBytes + te * 16;
#pragma unroll
          for (int i = 0; i < kBChunks / kFSplitThreads; ++i) {
            const uint4 v = lds128(bh + i * kFSplitThreads * 16);
            const uint32_t xs[4] = {v.x, v.y, v.z, v.w};
            uint32_t ls[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) ls[j] = __float_as_uint(__uint_as_float(xs[j]) - __uint_as_float(xs[j] & 0xFFFFE000u));
            sts128(bl + i * kFSplitThreads * 16, make_uint4(ls[0], ls[1], ls[2], ls[3]));
          }
          tmem_st_wait();
          tc_fence_before();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pready_bar[ps]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++ps == L::kProc) { ps = 0; pphase ^= 1; }
        }
      }
    } else if (PASSES == 3) {
      const int te = threadIdx.x - 128;
      uint32_t stage = 0, phase = 0;
      constexpr int kChunks = L::kHiBytes / 16;
      constexpr int kIters = kChunks / kFSplitThreads;
      static_assert(kChunks % (2 * kFSplitThreads) == 0, "tile bytes must divide over the splitter threads");
      const uint32_t ready_addr0 = CTAS == 2 ? mapa_u32(smem_u32(&ready_bar[0]), 0) : smem_u32(&ready_bar[0]);
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int split = (tile % tiles_per_batch) / tiles_mn;
        const int num_kb = min(p.kb_per_split, total_kb - split * p.kb_per_split);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          const uint32_t hi = sbase + stage * L::kStageBytes + te * 16;
          const uint32_t lo = hi + L::kHiBytes;
#pragma unroll
          for (int i0 = 0; i0 < kIters; i0 += 2) {
            uint4 v[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) v[i] = lds128(hi + (i0 + i) * kFSplitThreads * 16);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint32_t x[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, h[4], l[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (p.split_mode == 0) {
                  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h[j]) : "f"(__uint_as_float(x[j])));
                } else {
                  h[j] = x[j] & 0xFFFFE000u;
                }
                l[j] = __float_as_uint(__uint_as_float(x[j]) - __uint_as_float(h[j]));
              }
              if (p.split_mode == 0) sts128(hi + (i0 + i) * kFSplitThreads * 16, make_uint4(h[0], h[1], h[2], h[3]));
              sts128(lo + (i0 + i) * kFSplitThreads * 16, make_uint4(l[0], l[1], l[2], l[3]));
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 2) mbar_arrive_cluster(ready_addr0 + stage * 8); else mbar_arrive(&ready_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= kFEpiWarp0) {
    // ------------------------------ epilogue: TMEM -> registers -> smem transpose -> coalesced global stores ------------
    // A warp owns 32 rows x 32 columns per chunk (lane = row in TMEM); it stages the chunk in its private 4 KB buffer
    // (row-swizzled, conflict-free both ways) and writes it back four rows per instruction, 128 contiguous bytes per row.
    // Synchronous within the warp, so one buffer suffices and any number of stores stays in flight.
    const int ew = warp - kFEpiWarp0;
    const int e = ew & 3, half = ew >> 2;
    const uint32_t obuf = sbase + L::kOffOut + ew * kFStageOutBytes;
    const uint32_t wr_row = lane * 128, wr_swz = (lane & 7) << 4;
    const int rd_r = lane >> 3, rd_c = lane & 7;
    const uint32_t tmem_empty_addr0 = CTAS == 2 ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : smem_u32(&tmem_empty_bar[0]);
    int it = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++it) {
      const int batch = tile / tiles_per_batch, ts = tile - batch * tiles_per_batch;
      const int split = ts / tiles_mn, t = ts - split * tiles_mn;
      const int m0 = (t / n_tiles) * kTileM + rank * kFTileM, n0 = (t % n_tiles) * BLOCK_N;
      const int b0 = batch / p.nb1, b1 = batch - b0 * p.nb1;
      const uint32_t as = it & 1, aphase = (it >> 1) & 1;
      const int row0 = m0 + e * 32;
      const bool to_part = p.k_splits > 1;                    // partial product of one K chunk: bias and D come later
      const float* bias = to_part ? nullptr : p.bias;         // warp-uniform broadcast loads below (L1-resident)
      float* d_tile = to_part ? p.part + static_cast<int64_t>(split) * p.M * p.ldp : p.D + b0 * p.d_bs0 + b1 * p.d_bs1;
      const int64_t ldd = to_part ? p.ldp : p.ldd;
      const bool add = p.accumulate && !to_part;
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      constexpr int kChunks = BLOCK_N / 32;
      constexpr int kPerHalf = (kChunks + 1) / 2;
      const int c_begin = half * kPerHalf, c_end = min(kChunks, c_begin + kPerHalf);
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N || row0 >= p.M) break;               // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(e * 32) << 16) + as * BLOCK_N + c * 32, r);
        tmem_ld_wait();
        const bool tma_chunk = p.tma_out && !to_part && !add;      // warp-uniform
        if (tma_chunk) {
          if (lane == 0) tma_store_wait_read<0>();                  // the previous box store has read the staging buffer
          __syncwarp();
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias != nullptr) {
            const int col = col0 + 4 * j4;
            if (col + 3 < p.N) {
              b = __ldg(reinterpret_cast<const float4*>(bias + col));
            } else {
              if (col < p.N) b.x = __ldg(bias + col);
              if (col + 1 < p.N) b.y = __ldg(bias + col + 1);
              if (col + 2 < p.N) b.z = __ldg(bias + col + 2);
            }
          }
          const float v0 = fmaf(__uint_as_float(r[4 * j4 + 0]), p.scale, b.x);
          const float v1 = fmaf(__uint_as_float(r[4 * j4 + 1]), p.scale, b.y);
          const float v2 = fmaf(__uint_as_float(r[4 * j4 + 2]), p.scale, b.z);
          const float v3 = fmaf(__uint_as_float(r[4 * j4 + 3]), p.scale, b.w);
          sts128(obuf + wr_row + ((j4 << 4) ^ wr_swz), make_uint4(__float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2),
                                                                  __float_as_uint(v3)));
        }
        if (tma_chunk) {                                            // the box store clips the M / N tails itself
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&map_d, smem + L::kOffOut + ew * kFStageOutBytes, col0, row0, b1 * p.d_b1, b0 * p.d_b0);
            tma_store_commit();
          }
          continue;
        }
        __syncwarp();
        const int col = col0 + 4 * rd_c;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int rr = q * 4 + rd_r;
          const float4 v = lds128f(obuf + rr * 128 + ((rd_c ^ (rr & 7)) << 4));
          const int row = row0 + rr;
          if (row < p.M && col < p.N) {
            float* dst = d_tile + static_cast<int64_t>(row) * ldd + col;
            if (col + 3 < p.N) {
              if (add) atomicAdd(reinterpret_cast<float4*>(dst), v);
              else     *reinterpret_cast<float4*>(dst) = v;
            } else {
              const float vs[3] = {v.x, v.y, v.z};
              for (int u = 0; u < 3; ++u)
                if (col + u < p.N) {
                  if (add) atomicAdd(dst + u, vs[u]);
                  else     dst[u] = vs[u];
                }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(tmem_empty_addr0 + as * 8); else mbar_arrive(&tmem_empty_bar[as]);
      }
    }
    if (p.tma_out && lane == 0) tma_store_wait_all<0>();          // global writes complete before the CTA retires
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();     // pair: the peer's shared memory is read until the last MMA retires
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// D = (accumulate ? D : 0) + bias + sum_s part[s], fixed order; one thread per 4 columns
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N, int64_t ldp,
                                                            const float* __restrict__ bias, int accumulate, float* __restrict__ D,
                                                            int64_t ldd) {
  pdl_entry();
  const int n4 = static_cast<int>(ldp / 4);
  const int64_t total = static_cast<int64_t>(M) * n4;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int row = static_cast<int>(i / n4), col = static_cast<int>(i % n4) * 4;
    if (col >= N) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(part + (static_cast<int64_t>(s) * M + row) * ldp + col));
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    float* dst = D + static_cast<int64_t>(row) * ldd + col;
    const float vs[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (col + u < N) {
        float v = vs[u] + (bias != nullptr ? __ldg(bias + col + u) : 0.f);
        if (accumulate) v += dst[u];
        dst[u] = v;
      }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct F32Plan {
  int block_n, stages, k_splits, kb_per_split;
  int pair;                 // 1: CTA pairs on [256 x 256] tiles (cta_group::2)
};
static int g_f32_pair = 1;           // ob_debug_set key 8; measured: bitwise equal to single CTAs, 13-29 % faster (profiles/r01_f32gemm_pair_ab.jsonl)
constexpr int kSplitKChunk = 1024;

// Tile configuration and split-K decision.  Split-K: deep contractions with few output tiles (weight gradients, the
// front-end projection).  Chunks of <= kSplitKChunk keep the in-tensor-core accumulation chains short (the fp32
// accumulator truncates, which biases long chains) and fill the SMs; the chunk products are summed in fixed order.
static F32Plan plan_f32(int M, int N, int K, int batch) {
  F32Plan pl;
  if (N <= 64) pl.block_n = 64, pl.stages = 4;
  else if (N > 128 && K > 128) pl.block_n = 256, pl.stages = 2;
  else pl.block_n = 128, pl.stages = 3;
  // pairs pay where the kernel is shared-memory bound (deep contraction, wide output) and M fills the 256-row tiles
  pl.pair = g_f32_pair && pl.block_n == 256 && M >= 2 * kFTileM;
  const int tile_m = pl.pair ? 2 * kFTileM : kFTileM;
  const int64_t tiles_mn = (int64_t)((M + tile_m - 1) / tile_m) * ((N + pl.block_n - 1) / pl.block_n);
  const int total_kb = (K + kFKBlock - 1) / kFKBlock;
  pl.k_splits = 1, pl.kb_per_split = total_kb;
  if (batch == 1 && K >= 2 * kSplitKChunk && tiles_mn * (pl.pair ? 4 : 2) <= sm_count() * 4) {
    pl.kb_per_split = kSplitKChunk / kFKBlock;
    pl.k_splits = (total_kb + pl.kb_per_split - 1) / pl.kb_per_split;
  }
  return pl;
}
static int64_t part_pitch(int N) { return (N + 3) / 4 * 4; }

struct F32Operand {
  const void* ptr;
  int mn_major;
  int64_t ld, bs0, bs1;     // elements
};

// 4-D map over (contiguous axis, strided axis, inner batch, outer batch)
static int make_map4(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, int64_t ld, int nb1, int64_t bs1,
                     int nb0, int64_t bs0, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not found");
    return OB_ERR_CUDA;
  }
  const bool has1 = nb1 > 1 && bs1 != 0, has0 = nb0 > 1 && bs0 != 0;
  cuuint64_t gdim[4] = {inner, outer, has1 ? (cuuint64_t)nb1 : 1, has0 ? (cuuint64_t)nb0 : 1};
  cuuint64_t gstride[3] = {(cuuint64_t)ld * 4, has1 ? (cuuint64_t)bs1 * 4 : (cuuint64_t)ld * 4,
                           has0 ? (cuuint64_t)bs0 * 4 : (cuuint64_t)ld * 4};
  cuuint32_t box[4] = {box_inner, box_outer, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // (L2 promotion none / 128 B / 256 B measured identical on the attention products: tools/gpu_f32atm.py)
  const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
    // a thread that has made no runtime call yet (e.g. an autograd worker) has no current context for the driver API:
    // bind the primary context of the current device and retry
    cudaFree(nullptr);
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gdim, gstride, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D fp32) failed with %d (inner=%llu outer=%llu ld=%lld nb=%dx%d bs=%lld,%lld box=%ux%u)",
              (int)r, (unsigned long long)inner, (unsigned long long)outer, (long long)ld, nb0, nb1, (long long)bs0,
              (long long)bs1, box_inner, box_outer);
    return OB_ERR_CUDA;
  }
  return OB_OK;
}

static int g_f32_atm = 1;            // A through TMEM for the single-CTA split products (ob_debug_set key 11)
void f32_gemm_debug_atm(int on) { g_f32_atm = on; }
static int g_f32_split_mode = 1;     // the tensor core truncates fp32 -> tf32 (measured), so the raw tile is the hi part
static int g_f32_epilogue = 0;       // 0 auto, 1 direct row stores, 2 TMA box stores
void f32_gemm_debug_epilogue(int mode) { g_f32_epilogue = mode; }
void f32_gemm_debug_pair(int on) { g_f32_pair = on; }
void f32_gemm_debug(int split_mode) { g_f32_split_mode = split_mode; }

template <int A_MN, int B_MN, int BLOCK_N, int STAGES, int PASSES, int CTAS = 1, int ATM = 0>
static int launch_f32(const F32Operand& A, const F32Operand& B, float* D, int64_t ldd, int64_t d_bs0, int64_t d_bs1,
                      F32Params p, const F32Plan& pl, cudaStream_t st) {
  using L = F32Smem<BLOCK_N, STAGES, PASSES, CTAS, ATM>;
  static_assert(L::kDynBytes <= 232448, "shared memory budget exceeded");
  auto kern = f32_gemm_kernel<A_MN, B_MN, BLOCK_N, STAGES, PASSES, CTAS, ATM>;
  static bool attr_set = false;
  if (!attr_set) {
    OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynBytes));
    attr_set = true;
  }
  CUtensorMap map_a, map_b, map_d;
  int rc;
  constexpr CUtensorMapSwizzle kSwK = CU_TENSOR_MAP_SWIZZLE_128B, kSwMN = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  if (A_MN) rc = make_map4(&map_a, A.ptr, p.M, p.K, A.ld, p.nb1, A.bs1, p.nb0, A.bs0, 32, 32,
                           ATM ? CU_TENSOR_MAP_SWIZZLE_NONE : kSwMN);   // ATM: read by the splitter warps, not by the tensor core
  else      rc = make_map4(&map_a, A.ptr, p.K, p.M, A.ld, p.nb1, A.bs1, p.nb0, A.bs0, 32, kFTileM, kSwK);
  if (rc != OB_OK) return rc;
  if (B_MN) rc = make_map4(&map_b, B.ptr, p.N, p.K, B.ld, p.nb1, B.bs1, p.nb0, B.bs0, 32, 32, kSwMN);
  else      rc = make_map4(&map_b, B.ptr, p.K, p.N, B.ld, p.nb1, B.bs1, p.nb0, B.bs0, 32, L::kRowsB, kSwK);
  if (rc != OB_OK) return rc;
  rc = make_map4(&map_d, D, p.N, p.M, ldd, p.nb1, d_bs1, p.nb0, d_bs0, 32, 32, kSwK);
  if (rc != OB_OK) return rc;
  p.d_b0 = (p.nb0 > 1 && d_bs0 != 0), p.d_b1 = (p.nb1 > 1 && d_bs1 != 0);
  // box stores from the staging buffer beat the direct row stores on every measured shape (scores 92 -> 70 us, vocabulary
  // 334 -> 318 us); the direct path remains for accumulation, split-K partials and as a debug alternative
  p.tma_out = g_f32_epilogue != 1;
  p.a_b0 = (p.nb0 > 1 && A.bs0 != 0), p.a_b1 = (p.nb1 > 1 && A.bs1 != 0);
  p.b_b0 = (p.nb0 > 1 && B.bs0 != 0), p.b_b1 = (p.nb1 > 1 && B.bs1 != 0);
  p.D = D, p.ldd = ldd, p.d_bs0 = p.nb0 > 1 ? d_bs0 : 0, p.d_bs1 = p.nb1 > 1 ? d_bs1 : 0;
  p.split_mode = g_f32_split_mode;
  constexpr int kTileM = kFTileM * CTAS;
  const int64_t tiles_mn = (int64_t)((p.M + kTileM - 1) / kTileM) * ((p.N + BLOCK_N - 1) / BLOCK_N);
  p.k_splits = pl.k_splits, p.kb_per_split = pl.kb_per_split;
  const int64_t tiles = tiles_mn * p.k_splits * p.nb0 * p.nb1;
  int groups = sm_count() / CTAS;                  // persistent: one CTA (or CTA pair) per SM (pair of SMs)
  if (tiles < groups) groups = (int)tiles;
  if (CTAS == 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups * 2), cfg.blockDim = dim3(kFThreads), cfg.dynamicSmemBytes = L::kDynBytes, cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    attr[1] = pdl_attribute();
    cfg.attrs = attr, cfg.numAttrs = 2;
    OB_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_d, p));
  } else {
    launch_k((kern), dim3(groups), dim3(kFThreads), L::kDynBytes, st, map_a, map_b, map_d, p);
  }
  OB_LAUNCH_CHECK("f32_gemm_kernel");
  if (p.k_splits > 1) {
    const int64_t work = (int64_t)p.M * (p.ldp / 4);
    const int64_t want = (work + 255) / 256, cap = (int64_t)sm_count() * 16;
    launch_k((splitk_reduce_kernel), dim3((int)(want < cap ? want : cap)), dim3(256), 0, st, p.part, p.k_splits, p.M, p.N, p.ldp, p.bias,
                                                                         p.accumulate, D, ldd);
    OB_LAUNCH_CHECK("splitk_reduce_kernel");
  }
  return OB_OK;
}

template <int A_MN, int B_MN, int PASSES>
static int dispatch_f32_cfg(const F32Operand& A, const F32Operand& B, float* D, int64_t ldd, int64_t d_bs0, int64_t d_bs1,
                            const F32Params& p, const F32Plan& pl, cudaStream_t st) {
  if (PASSES == 3 && pl.pair) return launch_f32<A_MN, B_MN, 256, 3, 3, 2>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
  // The 128-wide single-CTA split products (attention scores and their gradients) keep A in tensor memory: 191 -> 171 us at
  // 192 x 4 x 399 x 399.  The 64-wide ones (probs . v and the transposed products) gain nothing from it - with a third of the
  // MMAs and no split at all they still take 125 us (tools/gpu_f32atm.py) - and stay on the shared-memory operand path.
  // (The truncating split only: cvt.rna would need the hi tile rewritten.)  g_f32_atm = 2 forces the 64-wide variant (tests).
  if (PASSES == 3 && g_f32_atm && g_f32_split_mode == 1) {
    if (pl.block_n == 128) return launch_f32<A_MN, B_MN, 128, 5, 3, 1, 1>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
    if (pl.block_n == 64 && g_f32_atm == 2) return launch_f32<A_MN, B_MN, 64, 7, 3, 1, 1>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
  }
  if (pl.block_n == 64) return launch_f32<A_MN, B_MN, 64, 4, PASSES>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
  if (pl.block_n == 256) return launch_f32<A_MN, B_MN, 256, 2, PASSES>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
  return launch_f32<A_MN, B_MN, 128, 3, PASSES>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
}

template <int PASSES>
static int dispatch_f32(const F32Operand& A, const F32Operand& B, float* D, int64_t ldd, int64_t d_bs0, int64_t d_bs1,
                        const F32Params& p, const F32Plan& pl, cudaStream_t st) {
  if (!A.mn_major && !B.mn_major) return dispatch_f32_cfg<0, 0, PASSES>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
  if (!A.mn_major && B.mn_major) return dispatch_f32_cfg<0, 1, PASSES>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
  if (A.mn_major && !B.mn_major) return dispatch_f32_cfg<1, 0, PASSES>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
  return dispatch_f32_cfg<1, 1, PASSES>(A, B, D, ldd, d_bs0, d_bs1, p, pl, st);
}

}  // namespace ob

using namespace ob;

static bool f32_ok(const void* ptr, int64_t ld, int64_t bs0, int64_t bs1) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld > 0 && ld % 4 == 0 && bs0 % 4 == 0 && bs1 % 4 == 0 && bs0 >= 0 &&
         bs1 >= 0;
}

extern "C" size_t ob_gemm_f32_workspace_bytes(int M, int N, int K, int nb0, int nb1) {
  if (M <= 0 || N <= 0 || K <= 0 || nb0 <= 0 || nb1 <= 0) return 0;
  const F32Plan pl = plan_f32(M, N, K, nb0 * nb1);
  return pl.k_splits > 1 ? static_cast<size_t>(pl.k_splits) * M * part_pitch(N) * sizeof(float) : 0;
}

extern "C" int ob_gemm_f32(const float* A, int a_mn_major, int64_t lda, int64_t a_bs0, int64_t a_bs1, const float* B,
                           int b_mn_major, int64_t ldb, int64_t b_bs0, int64_t b_bs1, float* D, int64_t ldd, int64_t d_bs0,
                           int64_t d_bs1, const float* bias, float scale, int accumulate, int M, int N, int K, int nb0,
                           int nb1, int passes, void* ws, size_t ws_bytes, ob_stream_t stream) {
  OB_REQUIRE(A && B && D, "ob_gemm_f32: null pointer");
  OB_REQUIRE(M > 0 && N > 0 && K > 0 && nb0 > 0 && nb1 > 0, "ob_gemm_f32: M, N, K and the batch counts must be positive");
  OB_REQUIRE(passes == 1 || passes == 3, "ob_gemm_f32: passes must be 1 (tf32) or 3 (fp32-level split)");
  OB_REQUIRE(f32_ok(A, lda, a_bs0, a_bs1) && f32_ok(B, ldb, b_bs0, b_bs1) && f32_ok(D, ldd, d_bs0, d_bs1),
             "ob_gemm_f32: pointers must be 16-byte aligned, leading dimensions and batch strides multiples of 4 elements");
  OB_REQUIRE(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K) && ldd >= N,
             "ob_gemm_f32: leading dimension smaller than the contiguous extent");
  OB_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "ob_gemm_f32: bias must be 16-byte aligned");
  OB_REQUIRE(accumulate || ((nb0 == 1 || d_bs0 != 0) && (nb1 == 1 || d_bs1 != 0)),
             "ob_gemm_f32: the output can be shared by a batch axis (stride 0) only when accumulating (sum over that axis)");
  int rc = check_device();
  if (rc != OB_OK) return rc;
  F32Operand a = {A, a_mn_major != 0, lda, a_bs0, a_bs1}, b = {B, b_mn_major != 0, ldb, b_bs0, b_bs1};
  F32Params p = {};
  p.bias = bias, p.scale = scale, p.accumulate = accumulate != 0;
  p.M = M, p.N = N, p.K = K, p.nb0 = nb0, p.nb1 = nb1;
  const F32Plan pl = plan_f32(M, N, K, nb0 * nb1);
  if (pl.k_splits > 1) {
    const size_t need = static_cast<size_t>(pl.k_splits) * M * part_pitch(N) * sizeof(float);
    if (ws == nullptr || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 15) != 0) {
      set_error("ob_gemm_f32: this shape splits K %d ways and needs a 16-byte aligned workspace of %zu bytes (got %zu)",
                pl.k_splits, need, ws_bytes);
      return OB_ERR_WORKSPACE;
    }
    p.part = static_cast<float*>(ws), p.ldp = part_pitch(N);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return passes == 3 ? dispatch_f32<3>(a, b, D, ldd, d_bs0, d_bs1, p, pl, st)
                     : dispatch_f32<1>(a, b, D, ldd, d_bs0, d_bs1, p, pl, st);
}
