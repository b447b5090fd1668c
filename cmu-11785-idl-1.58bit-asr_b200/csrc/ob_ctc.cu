// ob_ctc.cu - CTC loss of the training step, forward and backward, straight from the logits.
//
// Reference semantics (onebit_asr/losses.py:41-47):
//     log_probs = F.log_softmax(ctc_logits, dim=-1).transpose(0, 1)
//     nn.CTCLoss(blank, zero_infinity=True)(log_probs, tokens, feat_lens, token_lens)          # reduction 'mean'
// i.e. loss = mean_b( nll_b / max(L_b, 1) ), infinite nll_b (no valid alignment) counted as 0 with zero gradient.
//
// torch runs this as log_softmax (read + write [B,T,V]), a [B,T,V] -> [T,B,V] copy, alpha, beta, a collect kernel that
// writes a [T,B,V] gradient, an exp pass and the log_softmax backward: about ten passes over a 0.5 GB tensor per loss.
// Here the [B,T,V] logits are read twice and the gradient written once:
//   ctc_row_lse_kernel     lse[b,t] = logsumexp_v x[b,t,v]                 one block per frame, online max/sum, 128-bit loads
//   ctc_alpha_beta_kernel  forward and backward variables in the log domain, one block per (utterance, direction),
//                          one thread per state; log-probabilities are gathered as x[b,t,label] - lse[b,t], four frames
//                          prefetched ahead of the serial recursion; writes alpha, beta [B,T,S] and nll[b]
//   ctc_mean_kernel        loss = mean_b(nll_b / max(L_b,1)), fixed order
//   ctc_grad_kernel        dx[b,t,v] = g_b (softmax(x)[b,t,v] - occupancy[b,t,v]), occupancy summed per label in state
//                          order (deterministic), zeros beyond the input length; one block per frame
// All reductions have a fixed order: the loss and its gradient are run-to-run deterministic.
#include <cfloat>

#include "ob_common.cuh"

namespace ob {

constexpr int kCtcRowThreads = 256;
constexpr int kCtcPrefetch = 4;               // frames of gathered log-probabilities in flight in the recursion

__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// (max, sum of exp(x - max)) pairs combine associatively; fixed combine order below
__device__ __forceinline__ void lse_combine(float& m, float& s, float om, float os) {
  const float nm = fmaxf(m, om);
  if (nm == -INFINITY) { m = nm; s = 0.f; return; }
  s = s * expf(m - nm) + os * expf(om - nm);
  m = nm;
}

__global__ void __launch_bounds__(kCtcRowThreads)
ctc_row_lse_kernel(const float* __restrict__ x, int64_t ld, const int64_t* __restrict__ in_lens, int Tn, int V,
                   float* __restrict__ lse) {
  pdl_entry();
  const int64_t row = blockIdx.x;
  const int b = static_cast<int>(row / Tn), t = static_cast<int>(row % Tn);
  if (t >= in_lens[b]) {                                   // padded frame: never read
    if (threadIdx.x == 0) lse[row] = 0.f;
    return;
  }
  const float* xr = x + row * ld;
  float m = -INFINITY, s = 0.f;
  const int v4 = ((reinterpret_cast<uintptr_t>(xr) & 15) == 0) ? V / 4 : 0;
  for (int i = threadIdx.x; i < v4; i += kCtcRowThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + i);
    const float lm = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    if (lm > m) { s *= expf(m - lm); m = lm; }            // expf(-inf) = 0 on the first element
    if (m != -INFINITY) s += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
  }
  for (int i = 4 * v4 + threadIdx.x; i < V; i += kCtcRowThreads) {
    const float v = __ldg(xr + i);
    if (v > m) { s *= expf(m - v); m = v; }
    if (m != -INFINITY) s += expf(v - m);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    lse_combine(m, s, om, os);
  }
  __shared__ float wm[kCtcRowThreads / 32], wsum[kCtcRowThreads / 32];
  if ((threadIdx.x & 31) == 0) { wm[threadIdx.x >> 5] = m; wsum[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M = wm[0], S = wsum[0];
    for (int w = 1; w < kCtcRowThreads / 32; ++w) lse_combine(M, S, wm[w], wsum[w]);
    lse[row] = M + logf(S);
  }
}

// grid 2B: blocks [0, B) run the forward recursion of utterance b, blocks [B, 2B) the backward one.  blockDim >= S_max.
// vars [B, Tn, Sp] (Sp = state pitch).  Shared: two ping-pong rows of blockDim + 4 floats (two -inf pads each side).
__global__ void ctc_alpha_beta_kernel(const float* __restrict__ x, int64_t ld, const float* __restrict__ lse,
                                      const int64_t* __restrict__ in_lens, const int64_t* __restrict__ targets,
                                      int64_t tgt_ld, const int64_t* __restrict__ tgt_lens, int B, int Tn, int V, int Lmax,
                                      int blank, int Sp, float* __restrict__ alpha, float* __restrict__ beta,
                                      float* __restrict__ nll) {
  pdl_entry();
  extern __shared__ float sh[];
  const bool backward = blockIdx.x >= B;
  const int b = backward ? blockIdx.x - B : blockIdx.x;
  const int s = threadIdx.x;
  const int Tb = static_cast<int>(min(in_lens[b], static_cast<int64_t>(Tn)));
  const int Lb = static_cast<int>(min(max(tgt_lens[b], static_cast<int64_t>(0)), static_cast<int64_t>(Lmax)));
  const int S = 2 * Lb + 1;
  const int pitch = blockDim.x + 4;
  float* buf0 = sh + 2;                                    // buf[-2], buf[-1], buf[S], buf[S+1] stay -inf
  float* buf1 = sh + pitch + 2;
  for (int i = threadIdx.x; i < 2 * pitch; i += blockDim.x) sh[i] = -INFINITY;

  const bool live = s < S;
  int label = blank;
  bool skip = false;                                       // forward: s may be entered from s-2; backward: s may go to s+2
  if (live && (s & 1)) {
    const int j = s >> 1;
    label = static_cast<int>(targets[b * tgt_ld + j]);
    const int other = backward ? j + 1 : j - 1;
    if (other >= 0 && other < Lb) skip = static_cast<int>(targets[b * tgt_ld + other]) != label;
  }
  label = min(max(label, 0), V - 1);
  float* vars = (backward ? beta : alpha) + static_cast<int64_t>(b) * Tn * Sp;
  if (Tb <= 0) {                                           // no frames: only the empty target aligns
    if (!backward && s == 0) nll[b] = Lb == 0 ? 0.f : INFINITY;
    return;
  }
  const float* xb = x + static_cast<int64_t>(b) * Tn * ld + label;
  const float* lb = lse + static_cast<int64_t>(b) * Tn;
  __syncthreads();

  // frame order: forward 0 .. Tb-1, backward Tb-1 .. 0
  const int step = backward ? -1 : 1;
  const int t_first = backward ? Tb - 1 : 0;
  float pre_x[kCtcPrefetch], pre_l[kCtcPrefetch];          // raw loads; subtracted where they are consumed
#pragma unroll
  for (int i = 0; i < kCtcPrefetch; ++i) {
    const int t = t_first + step * i;
    const bool in = live && i < Tb;
    pre_x[i] = in ? __ldg(xb + static_cast<int64_t>(t) * ld) : 0.f;
    pre_l[i] = in ? __ldg(lb + t) : 0.f;
  }
  float* cur = buf0;
  float* nxt = buf1;
  for (int n0 = 0; n0 < Tb; n0 += kCtcPrefetch) {
    float cur_x[kCtcPrefetch], cur_l[kCtcPrefetch];
#pragma unroll
    for (int i = 0; i < kCtcPrefetch; ++i) cur_x[i] = pre_x[i], cur_l[i] = pre_l[i];
#pragma unroll
    for (int i = 0; i < kCtcPrefetch; ++i) {               // gathers of the next group fly during this group's recursion
      const int n = n0 + kCtcPrefetch + i;
      const int t = t_first + step * n;
      const bool in = live && n < Tb;
      pre_x[i] = in ? __ldg(xb + static_cast<int64_t>(t) * ld) : 0.f;
      pre_l[i] = in ? __ldg(lb + t) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kCtcPrefetch; ++i) {
      const int n = n0 + i;
      if (n >= Tb) break;                                  // block-uniform
      const int t = t_first + step * n;
      float v = -INFINITY;
      if (live) {
        if (n == 0) {
          // forward: states 0, 1 start; backward: states S-1, S-2 end
          const bool edge = backward ? (s >= S - 2) : (s <= 1);
          v = edge ? cur_x[i] - cur_l[i] : -INFINITY;
        } else {
          const float a = cur[s];
          const float c1 = backward ? cur[s + 1] : cur[s - 1];
          const float c2 = skip ? (backward ? cur[s + 2] : cur[s - 2]) : -INFINITY;
          v = lse3(a, c1, c2) + (cur_x[i] - cur_l[i]);
        }
        nxt[s] = v;
        vars[static_cast<int64_t>(t) * Sp + s] = v;
      }
      __syncthreads();
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
  }
  if (!backward) {
    // nll = -logsumexp(alpha[Tb-1][S-1], alpha[Tb-1][S-2]); `cur` holds the last frame
    if (s == 0) {
      const float a = cur[S - 1], c = S > 1 ? cur[S - 2] : -INFINITY;
      nll[b] = -lse3(a, c, -INFINITY);
    }
  }
}

// loss = (1/B) sum_b (isinf(nll_b) ? 0 : nll_b / max(L_b, 1)); one warp, fixed order
__global__ void __launch_bounds__(32)
ctc_mean_kernel(const float* __restrict__ nll, const int64_t* __restrict__ tgt_lens, int B, int Lmax, float* __restrict__ loss) {
  pdl_entry();
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 32) {
    const float v = nll[b];
    const int64_t L = min(max(tgt_lens[b], static_cast<int64_t>(1)), static_cast<int64_t>(max(Lmax, 1)));
    if (v != INFINITY) acc += v / static_cast<float>(L);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) loss[0] = acc / static_cast<float>(B);
}

// One block per frame.  Shared: occupancy per state [Sp], labels [Lmax].
__global__ void __launch_bounds__(kCtcRowThreads)
ctc_grad_kernel(const float* __restrict__ x, int64_t ld, const float* __restrict__ lse, const int64_t* __restrict__ in_lens,
                const int64_t* __restrict__ targets, int64_t tgt_ld, const int64_t* __restrict__ tgt_lens, int B, int Tn,
                int V, int Lmax, int blank, int Sp, const float* __restrict__ alpha, const float* __restrict__ beta,
                const float* __restrict__ nll, const float* __restrict__ grad_out, float* __restrict__ dx, int64_t ldg) {
  pdl_entry();
  extern __shared__ float sh[];
  float* occ = sh;                                         // [Sp]
  int* lab = reinterpret_cast<int*>(sh + Sp);              // [Lmax]
  const int64_t row = blockIdx.x;
  const int b = static_cast<int>(row / Tn), t = static_cast<int>(row % Tn);
  const int Tb = static_cast<int>(min(in_lens[b], static_cast<int64_t>(Tn)));
  const int Lb = static_cast<int>(min(max(tgt_lens[b], static_cast<int64_t>(0)), static_cast<int64_t>(Lmax)));
  const float nll_b = nll[b];
  float* dr = dx + row * ldg;
  const bool vec = (reinterpret_cast<uintptr_t>(dr) & 15) == 0 && (reinterpret_cast<uintptr_t>(x + row * ld) & 15) == 0;
  const int v4 = vec ? V / 4 : 0;
  if (t >= Tb || nll_b == INFINITY) {                      // beyond the input, or no valid alignment (zero_infinity)
    for (int i = threadIdx.x; i < v4; i += kCtcRowThreads) reinterpret_cast<float4*>(dr)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 4 * v4 + threadIdx.x; i < V; i += kCtcRowThreads) dr[i] = 0.f;
    return;
  }
  const int S = 2 * Lb + 1;
  const float g = grad_out[0] / (static_cast<float>(B) * static_cast<float>(max(Lb, 1)));
  const float* xr = x + row * ld;
  const float l = lse[row];
  // dense part: g * softmax
  for (int i = threadIdx.x; i < v4; i += kCtcRowThreads) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + i);
    reinterpret_cast<float4*>(dr)[i] = make_float4(g * expf(v.x - l), g * expf(v.y - l), g * expf(v.z - l), g * expf(v.w - l));
  }
  for (int i = 4 * v4 + threadIdx.x; i < V; i += kCtcRowThreads) dr[i] = g * expf(__ldg(xr + i) - l);
  // state occupancies exp(alpha + beta - lp + nll)
  for (int j = threadIdx.x; j < Lb; j += kCtcRowThreads) lab[j] = min(max(static_cast<int>(targets[b * tgt_ld + j]), 0), V - 1);
  __syncthreads();
  const float* ar = alpha + (static_cast<int64_t>(b) * Tn + t) * Sp;
  const float* br = beta + (static_cast<int64_t>(b) * Tn + t) * Sp;
  for (int s = threadIdx.x; s < S; s += kCtcRowThreads) {
    const int label = (s & 1) ? lab[s >> 1] : min(max(blank, 0), V - 1);
    const float ab = ar[s] + br[s];
    occ[s] = ab == -INFINITY ? 0.f : expf(ab - (__ldg(xr + label) - l) + nll_b);
  }
  __syncthreads();                                         // also orders the dense stores above before the updates below
  // sparse part, summed per label in state order: the blank by thread 0, each label by its first occurrence
  for (int j = threadIdx.x; j <= Lb; j += kCtcRowThreads) {
    if (j == Lb) {
      float sum = 0.f;
      for (int s = 0; s < S; s += 2) sum += occ[s];
      const int label = min(max(blank, 0), V - 1);
      bool shared_with_target = false;                     // a target equal to the blank id is malformed; keep the sum exact anyway
      for (int k = 0; k < Lb; ++k) shared_with_target |= lab[k] == label;
      if (!shared_with_target) dr[label] -= g * sum;
      else atomicAdd(dr + label, -g * sum);
    } else {
      const int label = lab[j];
      bool first = true;
      for (int k = 0; k < j; ++k) first &= lab[k] != label;
      if (first) {
        float sum = 0.f;
        for (int k = j; k < Lb; ++k)
          if (lab[k] == label) sum += occ[2 * k + 1];
        if (label != min(max(blank, 0), V - 1)) dr[label] -= g * sum;
        else atomicAdd(dr + label, -g * sum);
      }
    }
  }
}

}  // namespace ob

using namespace ob;

static int ctc_block_threads(int Lmax) { return ((2 * Lmax + 1) + 31) / 32 * 32; }

extern "C" int ob_ctc_state_pitch(int Lmax) { return Lmax >= 0 ? (2 * Lmax + 1 + 3) / 4 * 4 : 0; }

static int ctc_check(const void* logits, int64_t ld, int B, int T, int V, int Lmax, int blank, int64_t tgt_ld) {
  OB_REQUIRE(logits != nullptr, "ob_ctc_loss: null logits");
  OB_REQUIRE(B > 0 && T > 0 && V > 0 && Lmax >= 0, "ob_ctc_loss: bad shape B=%d T=%d V=%d Lmax=%d", B, T, V, Lmax);
  OB_REQUIRE(ld >= V && tgt_ld >= Lmax, "ob_ctc_loss: row pitch smaller than the row (ld=%lld V=%d, tgt_ld=%lld Lmax=%d)",
             (long long)ld, V, (long long)tgt_ld, Lmax);
  OB_REQUIRE(blank >= 0 && blank < V, "ob_ctc_loss: blank id %d outside the vocabulary of %d", blank, V);
  OB_REQUIRE(ctc_block_threads(Lmax) <= 1024, "ob_ctc_loss: targets longer than 511 labels are not supported (Lmax=%d)", Lmax);
  return OB_OK;
}

extern "C" int ob_ctc_loss_fwd(const float* logits, int64_t ld, const int64_t* in_lens, const int64_t* targets, int64_t tgt_ld,
                               const int64_t* tgt_lens, int B, int T, int V, int Lmax, int blank, float* lse, float* alpha,
                               float* beta, float* nll, float* loss, ob_stream_t stream) {
  int rc = ctc_check(logits, ld, B, T, V, Lmax, blank, tgt_ld);
  if (rc != OB_OK) return rc;
  OB_REQUIRE(in_lens && tgt_lens && (targets || Lmax == 0) && lse && alpha && beta && nll && loss, "ob_ctc_loss_fwd: null pointer");
  rc = check_device();
  if (rc != OB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Sp = ob_ctc_state_pitch(Lmax);
  launch_k((ctc_row_lse_kernel), dim3(B * T), dim3(kCtcRowThreads), 0, st, logits, ld, in_lens, T, V, lse);
  OB_LAUNCH_CHECK("ctc_row_lse_kernel");
  const int threads = ctc_block_threads(Lmax);
  const size_t smem = 2 * static_cast<size_t>(threads + 4) * sizeof(float);
  launch_k((ctc_alpha_beta_kernel), dim3(2 * B), dim3(threads), smem, st, logits, ld, lse, in_lens, targets, tgt_ld, tgt_lens, B, T, V, Lmax, blank,
                                                      Sp, alpha, beta, nll);
  OB_LAUNCH_CHECK("ctc_alpha_beta_kernel");
  launch_k((ctc_mean_kernel), dim3(1), dim3(32), 0, st, nll, tgt_lens, B, Lmax, loss);
  OB_LAUNCH_CHECK("ctc_mean_kernel");
  return OB_OK;
}

extern "C" int ob_ctc_loss_bwd(const float* logits, int64_t ld, const int64_t* in_lens, const int64_t* targets, int64_t tgt_ld,
                               const int64_t* tgt_lens, int B, int T, int V, int Lmax, int blank, const float* lse,
                               const float* alpha, const float* beta, const float* nll, const float* grad_out, float* grad_logits,
                               int64_t ldg, ob_stream_t stream) {
  int rc = ctc_check(logits, ld, B, T, V, Lmax, blank, tgt_ld);
  if (rc != OB_OK) return rc;
  OB_REQUIRE(in_lens && tgt_lens && (targets || Lmax == 0) && lse && alpha && beta && nll && grad_out && grad_logits,
             "ob_ctc_loss_bwd: null pointer");
  OB_REQUIRE(ldg >= V, "ob_ctc_loss_bwd: gradient row pitch %lld smaller than V=%d", (long long)ldg, V);
  rc = check_device();
  if (rc != OB_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Sp = ob_ctc_state_pitch(Lmax);
  const size_t smem = static_cast<size_t>(Sp) * sizeof(float) + static_cast<size_t>(Lmax > 0 ? Lmax : 1) * sizeof(int);
  launch_k((ctc_grad_kernel), dim3(B * T), dim3(kCtcRowThreads), smem, st, logits, ld, lse, in_lens, targets, tgt_ld, tgt_lens, B, T, V, Lmax, blank,
                                                       Sp, alpha, beta, nll, grad_out, grad_logits, ldg);
  OB_LAUNCH_CHECK("ctc_grad_kernel");
  return OB_OK;
}
