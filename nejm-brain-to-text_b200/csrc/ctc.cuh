// Fused log-softmax + CTC loss + gradient wrt logits, one CTA per trial.
//
// Reference: rnn_trainer.py:538-545 -- torch.nn.CTCLoss(blank=0, reduction='none',
// zero_infinity=False) applied to logits.log_softmax(2), then torch.mean over the batch.
// alpha/beta follow Graves et al. 2006 with both including the emission at t, so that
//   dL/dlogit[t,c] = softmax[t,c] - exp(logsumexp_{s: l'_s = c}(alpha_t(s)+beta_t(s)) - lp[t,c] - ll)
// (the form ATen's ctc_loss backward uses).  Label expansion, lengths and the skip rule are integer
// logic and must match exactly.
//
// Numerics: the recursions run in fp32 log space, but every frame's alpha (beta) row is shifted so that its
// maximum is 0 and the shifts are accumulated in double.  Without this the fp32 rounding of values of magnitude
// ~T*log(C) accumulates to ~3e-5 absolute gradient error at T = 40; with it the gradient agrees with a float64
// evaluation to ~1e-6.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cuda_bf16.h>

namespace b2t {

constexpr int CTC_THREADS = 128;

struct CtcParams {
  const float* logits;      // row (t*Bpad + b), pitch ldl floats
  int ldl, Bpad, T, C, blank;
  const int* labels;        // [B][Smax]
  int Smax;
  const int* in_len;        // [B]
  const int* tgt_len;       // [B]
  float* alpha;             // scratch [B][T][Lmax], Lmax = 2*Smax+1 (row-normalised alpha)
  float* beta;              // scratch, same shape (row-normalised beta; used by the CTA-per-trial parallel kernel)
  float* loss;              // [B]
  float* dlogits;           // fp32, same layout as logits (nullable => loss only)
  __nv_bfloat16* dlogits_bf16;  // bf16 copy for the tensor-core GEMMs (nullable)
  float* dbias;             // [C] atomicAdd of sum_t,b dlogits (nullable)
  float grad_scale;         // 1 / global batch (torch.mean)
};

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + log1pf(expf(fminf(a, b) - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, b), c);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// max over the block of per-thread values; result broadcast to every thread (one __syncthreads pair)
__device__ __forceinline__ float ctc_block_max(float v, float* sred) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sred[0];
#pragma unroll
  for (int i = 1; i < CTC_THREADS / 32; ++i) r = fmaxf(r, sred[i]);
  __syncthreads();
  return r;
}

// dynamic smem: shiftA[T] (double) | lp[T*C] | ext[Lmax] (int) | rowA[Lmax] | rowB[Lmax] | ab[Lmax]
__global__ void __launch_bounds__(CTC_THREADS)
ctc_loss_grad_kernel(const CtcParams p) {
  extern __shared__ double ctc_smem_d[];
  const int b = blockIdx.x;
  const int Lmax = 2 * p.Smax + 1;
  double* shiftA = ctc_smem_d;                               // cumulative alpha shift up to and including frame t
  float* lp = reinterpret_cast<float*>(shiftA + p.T);
  int* ext = reinterpret_cast<int*>(lp + (size_t)p.T * p.C);
  float* rowA = reinterpret_cast<float*>(ext + Lmax);
  float* rowB = rowA + Lmax;
  float* ab = rowB + Lmax;
  __shared__ float sred[CTC_THREADS / 32];
  __shared__ double s_ll;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int Tb = p.in_len[b];
  Tb = Tb < 0 ? 0 : (Tb > p.T ? p.T : Tb);
  const int S = p.tgt_len[b];
  const int L = 2 * S + 1;
  const float NINF = -CUDART_INF_F;

  // ---- extended label sequence (integer logic)
  for (int s = tid; s < L; s += CTC_THREADS) ext[s] = (s & 1) ? p.labels[(size_t)b * p.Smax + (s >> 1)] : p.blank;

  // ---- log-softmax per frame: one warp per frame
  for (int t = warp; t < Tb; t += CTC_THREADS / 32) {
    const float* row = p.logits + ((size_t)t * p.Bpad + b) * p.ldl;
    float m = NINF;
    for (int c = lane; c < p.C; c += 32) m = fmaxf(m, row[c]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int c = lane; c < p.C; c += 32) sum += expf(row[c] - m);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float lz = m + logf(sum);
    for (int c = lane; c < p.C; c += 32) lp[t * p.C + c] = row[c] - lz;
  }
  __syncthreads();

  float* alpha = p.alpha + (size_t)b * p.T * Lmax;
  // ---- alpha recursion (rows shifted to max 0; shifts accumulated in double)
  float* prev = rowA;
  float* cur = rowB;
  double accA = 0.0;
  for (int t = 0; t < Tb; ++t) {
    float vmax = NINF;
    for (int s = tid; s < L; s += CTC_THREADS) {
      float a;
      if (t == 0) {
        a = NINF;
        if (s == 0) a = lp[p.blank];
        else if (s == 1) a = lp[ext[1]];
      } else {
        const float a0 = prev[s];
        const float a1 = s >= 1 ? prev[s - 1] : NINF;
        const float a2 = (s >= 2 && ext[s] != p.blank && ext[s] != ext[s - 2]) ? prev[s - 2] : NINF;
        a = lse3(a0, a1, a2) + lp[t * p.C + ext[s]];
      }
      cur[s] = a;
      vmax = fmaxf(vmax, a);
    }
    const float m = ctc_block_max(vmax, sred);               // includes the barriers that publish cur[]
    const float sh = (m == NINF) ? 0.f : m;
    for (int s = tid; s < L; s += CTC_THREADS) {
      const float a = cur[s] - sh;
      cur[s] = a;
      alpha[(size_t)t * Lmax + s] = a;
    }
    accA += (double)sh;
    if (tid == 0) shiftA[t] = accA;
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  if (tid == 0) {
    double ll;
    if (Tb > 0) ll = (double)lse2(prev[L - 1], L > 1 ? prev[L - 2] : NINF) + accA;
    else ll = (S == 0) ? 0.0 : -(double)CUDART_INF_F;
    s_ll = ll;
    p.loss[b] = (float)(-ll);
  }
  __syncthreads();
  if (!p.dlogits && !p.dlogits_bf16) return;
  const double ll = s_ll;

  // ---- beta recursion + gradient, t descending (same normalisation)
  float dbacc = 0.f;   // thread c < C accumulates sum_t dlogits[t][c]
  float* bprev = rowA;
  float* bcur = rowB;
  double accB = 0.0;
  for (int t = p.T - 1; t >= 0; --t) {
    float* drow = p.dlogits ? p.dlogits + ((size_t)t * p.Bpad + b) * p.ldl : nullptr;
    __nv_bfloat16* drow16 = p.dlogits_bf16 ? p.dlogits_bf16 + ((size_t)t * p.Bpad + b) * p.ldl : nullptr;
    if (t >= Tb) {     // frames beyond the input length get zero gradient
      for (int c = tid; c < p.ldl; c += CTC_THREADS) {
        if (drow) drow[c] = 0.f;
        if (drow16) drow16[c] = __float2bfloat16_rn(0.f);
      }
      continue;
    }
    float vmax = NINF;
    for (int s = tid; s < L; s += CTC_THREADS) {
      float v;
      if (t == Tb - 1) {
        v = (s == L - 1 || s == L - 2) ? lp[t * p.C + ext[s]] : NINF;
      } else {
        const float b0 = bprev[s];
        const float b1 = s + 1 < L ? bprev[s + 1] : NINF;
        const float b2 = (s + 2 < L && ext[s + 2] != p.blank && ext[s + 2] != ext[s]) ? bprev[s + 2] : NINF;
        v = lse3(b0, b1, b2) + lp[t * p.C + ext[s]];
      }
      bcur[s] = v;
      vmax = fmaxf(vmax, v);
    }
    const float m = ctc_block_max(vmax, sred);
    const float sh = (m == NINF) ? 0.f : m;
    accB += (double)sh;
    // log of the path mass through (t, s) relative to the total: alpha + beta - ll with the big terms combined in double
    const float base = (float)(shiftA[t] + accB - ll);
    for (int s = tid; s < L; s += CTC_THREADS) {
      const float v = bcur[s] - sh;
      bcur[s] = v;
      ab[s] = alpha[(size_t)t * Lmax + s] + v + base;
    }
    __syncthreads();
    for (int c = tid; c < p.ldl; c += CTC_THREADS) {
      float g = 0.f;
      if (c < p.C) {
        float mm = NINF;
        for (int s = (c == p.blank ? 0 : 1); s < L; s += 2)
          if (ext[s] == c) mm = fmaxf(mm, ab[s]);
        float occ = 0.f;
        if (mm != NINF) {
          float sum = 0.f;
          for (int s = (c == p.blank ? 0 : 1); s < L; s += 2)
            if (ext[s] == c) sum += expf(ab[s] - mm);
          occ = expf(mm + logf(sum) - lp[t * p.C + c]);
        }
        g = (expf(lp[t * p.C + c]) - occ) * p.grad_scale;
        dbacc += g;
      }
      if (drow) drow[c] = g;
      if (drow16) drow16[c] = __float2bfloat16_rn(g);
    }
    __syncthreads();
    float* tmp = bprev; bprev = bcur; bcur = tmp;
  }
  if (p.dbias && tid < p.C) atomicAdd(p.dbias + tid, dbacc);
}

// ---------------------------------------------------------------------------------------------------------
// Warp-per-trial variant for label sequences with 2*Smax+1 <= 32*CPL (CPL <= 4): no block barriers at all.
// Lane i owns the CPL consecutive lattice cells s in [i*CPL, (i+1)*CPL); neighbours come from warp shuffles; the
// per-class occupancy of a frame is accumulated in linear space (a posterior, <= 1) with shared-memory atomics.
// Same arithmetic as the block kernel above (row-normalised fp32 recursions, shifts in double).
template <int CPL>
__global__ void __launch_bounds__(32)
ctc_loss_grad_warp_kernel(const CtcParams p) {
  extern __shared__ double ctc_smem_d[];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int Lmax = 2 * p.Smax + 1;
  double* shiftA = ctc_smem_d;
  float* lp = reinterpret_cast<float*>(shiftA + p.T);
  float* occ = lp + (size_t)p.T * p.C;                        // [C]
  int Tb = p.in_len[b];
  Tb = Tb < 0 ? 0 : (Tb > p.T ? p.T : Tb);
  const int S = p.tgt_len[b];
  const int L = 2 * S + 1;
  const float NINF = -CUDART_INF_F;
  const unsigned FULL = 0xffffffffu;

  int ext[CPL];
  bool skip[CPL], skipf[CPL];                                 // may come from s-2 / may go to s+2
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int s = lane * CPL + j;
    auto lab = [&](int q) { return (q & 1) ? p.labels[(size_t)b * p.Smax + (q >> 1)] : p.blank; };
    ext[j] = s < L ? lab(s) : p.blank;
    skip[j] = s < L && s >= 2 && ext[j] != p.blank && ext[j] != lab(s - 2);
    skipf[j] = s + 2 < L && lab(s + 2) != p.blank && lab(s + 2) != ext[j];
  }
  for (int t = 0; t < Tb; ++t) {                              // log-softmax rows (2 classes per lane for C <= 64)
    const float* row = p.logits + ((size_t)t * p.Bpad + b) * p.ldl;
    float m = NINF;
    for (int c = lane; c < p.C; c += 32) m = fmaxf(m, row[c]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    float sum = 0.f;
    for (int c = lane; c < p.C; c += 32) sum += expf(row[c] - m);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    const float lz = m + logf(sum);
    for (int c = lane; c < p.C; c += 32) lp[t * p.C + c] = row[c] - lz;
  }
  __syncwarp();

  float* alpha = p.alpha + (size_t)b * p.T * Lmax;
  float a[CPL];
  double accA = 0.0;
  for (int t = 0; t < Tb; ++t) {
    float na[CPL];
    if (t == 0) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int s = lane * CPL + j;
        na[j] = (s == 0) ? lp[p.blank] : ((s == 1 && s < L) ? lp[ext[j]] : NINF);
      }
    } else {
      float l1 = __shfl_up_sync(FULL, a[CPL - 1], 1);
      float l2 = CPL >= 2 ? __shfl_up_sync(FULL, a[CPL >= 2 ? CPL - 2 : 0], 1) : __shfl_up_sync(FULL, a[0], 2);
      if (lane == 0) { l1 = NINF; l2 = NINF; }
      if (CPL == 1 && lane == 1) l2 = NINF;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int s = lane * CPL + j;
        const float a1 = j >= 1 ? a[j >= 1 ? j - 1 : 0] : l1;
        const float a2 = j >= 2 ? a[j >= 2 ? j - 2 : 0] : (j == 1 ? l1 : l2);
        na[j] = s < L ? lse3(a[j], a1, skip[j] ? a2 : NINF) + lp[t * p.C + ext[j]] : NINF;
      }
    }
    float m = na[0];
#pragma unroll
    for (int j = 1; j < CPL; ++j) m = fmaxf(m, na[j]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    const float sh = (m == NINF) ? 0.f : m;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      a[j] = na[j] - sh;
      const int s = lane * CPL + j;
      if (s < L) alpha[(size_t)t * Lmax + s] = a[j];
    }
    accA += (double)sh;
    if (lane == 0) shiftA[t] = accA;
  }
  double ll;
  {
    // alpha_T(L-1) and alpha_T(L-2) live in some lanes; fetch them with shuffles
    float last = NINF, last2 = NINF;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int s = lane * CPL + j;
      if (s == L - 1) last = a[j];
      if (s == L - 2) last2 = a[j];
    }
    for (int o = 16; o; o >>= 1) { last = fmaxf(last, __shfl_xor_sync(FULL, last, o)); last2 = fmaxf(last2, __shfl_xor_sync(FULL, last2, o)); }
    if (Tb > 0) ll = (double)lse2(last, last2) + accA;
    else ll = (S == 0) ? 0.0 : -(double)CUDART_INF_F;
  }
  if (lane == 0) p.loss[b] = (float)(-ll);
  if (!p.dlogits && !p.dlogits_bf16) return;
  __syncwarp();

  float dbacc[2] = {0.f, 0.f};
  float bt[CPL];
  double accB = 0.0;
  for (int t = p.T - 1; t >= 0; --t) {
    float* drow = p.dlogits ? p.dlogits + ((size_t)t * p.Bpad + b) * p.ldl : nullptr;
    __nv_bfloat16* drow16 = p.dlogits_bf16 ? p.dlogits_bf16 + ((size_t)t * p.Bpad + b) * p.ldl : nullptr;
    if (t >= Tb) {
      for (int c = lane; c < p.ldl; c += 32) {
        if (drow) drow[c] = 0.f;
        if (drow16) drow16[c] = __float2bfloat16_rn(0.f);
      }
      continue;
    }
    float nb[CPL];
    if (t == Tb - 1) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int s = lane * CPL + j;
        nb[j] = (s < L && (s == L - 1 || s == L - 2)) ? lp[t * p.C + ext[j]] : NINF;
      }
    } else {
      float r1 = __shfl_down_sync(FULL, bt[0], 1);
      float r2 = CPL >= 2 ? __shfl_down_sync(FULL, bt[CPL >= 2 ? 1 : 0], 1) : __shfl_down_sync(FULL, bt[0], 2);
      if (lane == 31) { r1 = NINF; r2 = NINF; }
      if (CPL == 1 && lane == 30) r2 = NINF;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int s = lane * CPL + j;
        const float b1 = j + 1 < CPL ? bt[j + 1 < CPL ? j + 1 : 0] : r1;
        const float b2 = j + 2 < CPL ? bt[j + 2 < CPL ? j + 2 : 0] : (j + 1 < CPL ? r1 : r2);
        nb[j] = s < L ? lse3(bt[j], s + 1 < L ? b1 : NINF, skipf[j] ? b2 : NINF) + lp[t * p.C + ext[j]] : NINF;
      }
    }
    float m = nb[0];
#pragma unroll
    for (int j = 1; j < CPL; ++j) m = fmaxf(m, nb[j]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    const float sh = (m == NINF) ? 0.f : m;
    accB += (double)sh;
    const float base = (float)(shiftA[t] + accB - ll);
    for (int c = lane; c < p.C; c += 32) occ[c] = 0.f;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      bt[j] = nb[j] - sh;
      const int s = lane * CPL + j;
      if (s < L) {
        const float ab = alpha[(size_t)t * Lmax + s] + bt[j] + base;       // log of the (doubly emitted) path mass through (t, s)
        const float w = expf(ab - lp[t * p.C + ext[j]]);
        if (w > 0.f) atomicAdd(&occ[ext[j]], w);
      }
    }
    __syncwarp();
    int k = 0;
    for (int c = lane; c < p.ldl; c += 32, ++k) {
      float g = 0.f;
      if (c < p.C) {
        g = (expf(lp[t * p.C + c]) - occ[c]) * p.grad_scale;
        if (k < 2) dbacc[k] += g;
      }
      if (drow) drow[c] = g;
      if (drow16) drow16[c] = __float2bfloat16_rn(g);
    }
    __syncwarp();
  }
  if (p.dbias) {
    int k = 0;
    for (int c = lane; c < p.C && k < 2; c += 32, ++k) atomicAdd(p.dbias + c, dbacc[k]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// CTA-per-trial variant for 2*Smax+1 <= 32*CPL: the three phases of the warp kernel, de-serialised.
//   1. log-softmax: one warp per frame, all warps;
//   2. warp 0 runs the alpha recursion (t ascending) WHILE warp 1 runs the beta recursion (t descending); both keep the
//      lattice row in registers (neighbours by shuffle) and stream the normalised rows to L2 scratch;
//   3. gradient: one warp per frame, all warps, from alpha + beta + the accumulated shifts.
// The serial chain is max(alpha, beta) steps instead of log-softmax + alpha + (beta + gradient) steps, and neither
// recursion has memory loads, atomics or stores-with-consumers on its dependency chain.  Arithmetic is the same as in
// the kernels above (row-normalised fp32 recursions, shifts in double).
constexpr int CTC_PAR_WARPS = 8;

template <int CPL>
__global__ void __launch_bounds__(32 * CTC_PAR_WARPS)
ctc_loss_grad_par_kernel(const CtcParams p) {
  extern __shared__ double ctc_smem_d[];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Lmax = 2 * p.Smax + 1;
  double* shiftA = ctc_smem_d;                                // [T] cumulative alpha shift up to and including frame t
  double* shiftB = shiftA + p.T;                              // [T] cumulative beta shift from the last frame down to t
  float* lp = reinterpret_cast<float*>(shiftB + p.T);         // [T][C]
  float* occ_all = lp + (size_t)p.T * p.C;                    // [warps][C]
  float* dbsum = occ_all + CTC_PAR_WARPS * p.C;               // [C]
  __shared__ double s_ll;
  int Tb = p.in_len[b];
  Tb = Tb < 0 ? 0 : (Tb > p.T ? p.T : Tb);
  const int S = p.tgt_len[b];
  const int L = 2 * S + 1;
  const float NINF = -CUDART_INF_F;
  const unsigned FULL = 0xffffffffu;

  int ext[CPL];
  bool skip[CPL], skipf[CPL];                                 // may come from s-2 / may go to s+2
#pragma unroll
  for (int j = 0; j < CPL; ++j) {
    const int s = lane * CPL + j;
    auto lab = [&](int q) { return (q & 1) ? p.labels[(size_t)b * p.Smax + (q >> 1)] : p.blank; };
    ext[j] = s < L ? lab(s) : p.blank;
    skip[j] = s < L && s >= 2 && ext[j] != p.blank && ext[j] != lab(s - 2);
    skipf[j] = s + 2 < L && lab(s + 2) != p.blank && lab(s + 2) != ext[j];
  }
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) dbsum[c] = 0.f;

  // ---- 1. log-softmax rows
  for (int t = warp; t < Tb; t += CTC_PAR_WARPS) {
    const float* row = p.logits + ((size_t)t * p.Bpad + b) * p.ldl;
    float m = NINF;
    for (int c = lane; c < p.C; c += 32) m = fmaxf(m, row[c]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
    float sum = 0.f;
    for (int c = lane; c < p.C; c += 32) sum += expf(row[c] - m);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    const float lz = m + logf(sum);
    for (int c = lane; c < p.C; c += 32) lp[t * p.C + c] = row[c] - lz;
  }
  __syncthreads();

  // ---- 2. the two recursions, concurrently
  float* alpha = p.alpha + (size_t)b * p.T * Lmax;
  float* beta = p.beta + (size_t)b * p.T * Lmax;
  if (warp == 0) {
    float a[CPL];
    double accA = 0.0;
    for (int t = 0; t < Tb; ++t) {
      float na[CPL];
      if (t == 0) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int s = lane * CPL + j;
          na[j] = (s == 0) ? lp[p.blank] : ((s == 1 && s < L) ? lp[ext[j]] : NINF);
        }
      } else {
        float l1 = __shfl_up_sync(FULL, a[CPL - 1], 1);
        float l2 = CPL >= 2 ? __shfl_up_sync(FULL, a[CPL >= 2 ? CPL - 2 : 0], 1) : __shfl_up_sync(FULL, a[0], 2);
        if (lane == 0) { l1 = NINF; l2 = NINF; }
        if (CPL == 1 && lane == 1) l2 = NINF;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int s = lane * CPL + j;
          const float a1 = j >= 1 ? a[j >= 1 ? j - 1 : 0] : l1;
          const float a2 = j >= 2 ? a[j >= 2 ? j - 2 : 0] : (j == 1 ? l1 : l2);
          na[j] = s < L ? lse3(a[j], a1, skip[j] ? a2 : NINF) + lp[t * p.C + ext[j]] : NINF;
        }
      }
      float m = na[0];
#pragma unroll
      for (int j = 1; j < CPL; ++j) m = fmaxf(m, na[j]);
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
      const float sh = (m == NINF) ? 0.f : m;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        a[j] = na[j] - sh;
        const int s = lane * CPL + j;
        if (s < L) alpha[(size_t)t * Lmax + s] = a[j];
      }
      accA += (double)sh;
      if (lane == 0) shiftA[t] = accA;
    }
    float last = NINF, last2 = NINF;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int s = lane * CPL + j;
      if (Tb > 0 && s == L - 1) last = a[j];
      if (Tb > 0 && s == L - 2) last2 = a[j];
    }
    for (int o = 16; o; o >>= 1) { last = fmaxf(last, __shfl_xor_sync(FULL, last, o)); last2 = fmaxf(last2, __shfl_xor_sync(FULL, last2, o)); }
    double ll;
    if (Tb > 0) ll = (double)lse2(last, last2) + accA;
    else ll = (S == 0) ? 0.0 : -(double)CUDART_INF_F;
    if (lane == 0) { s_ll = ll; p.loss[b] = (float)(-ll); }
  } else if (warp == 1 && (p.dlogits || p.dlogits_bf16)) {
    float bt[CPL];
    double accB = 0.0;
    for (int t = Tb - 1; t >= 0; --t) {
      float nb[CPL];
      if (t == Tb - 1) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int s = lane * CPL + j;
          nb[j] = (s < L && (s == L - 1 || s == L - 2)) ? lp[t * p.C + ext[j]] : NINF;
        }
      } else {
        float r1 = __shfl_down_sync(FULL, bt[0], 1);
        float r2 = CPL >= 2 ? __shfl_down_sync(FULL, bt[CPL >= 2 ? 1 : 0], 1) : __shfl_down_sync(FULL, bt[0], 2);
        if (lane == 31) { r1 = NINF; r2 = NINF; }
        if (CPL == 1 && lane == 30) r2 = NINF;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int s = lane * CPL + j;
          const float b1 = j + 1 < CPL ? bt[j + 1 < CPL ? j + 1 : 0] : r1;
          const float b2 = j + 2 < CPL ? bt[j + 2 < CPL ? j + 2 : 0] : (j + 1 < CPL ? r1 : r2);
          nb[j] = s < L ? lse3(bt[j], s + 1 < L ? b1 : NINF, skipf[j] ? b2 : NINF) + lp[t * p.C + ext[j]] : NINF;
        }
      }
      float m = nb[0];
#pragma unroll
      for (int j = 1; j < CPL; ++j) m = fmaxf(m, nb[j]);
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
      const float sh = (m == NINF) ? 0.f : m;
      accB += (double)sh;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        bt[j] = nb[j] - sh;
        const int s = lane * CPL + j;
        if (s < L) beta[(size_t)t * Lmax + s] = bt[j];
      }
      if (lane == 0) shiftB[t] = accB;
    }
  }
  __syncthreads();                                            // also makes this CTA's alpha/beta rows (global) visible to all its warps
  if (!p.dlogits && !p.dlogits_bf16) return;
  const double ll = s_ll;

  // ---- 3. gradient rows
  float dbacc[2] = {0.f, 0.f};
  float* occ = occ_all + warp * p.C;
  for (int t = warp; t < p.T; t += CTC_PAR_WARPS) {
    float* drow = p.dlogits ? p.dlogits + ((size_t)t * p.Bpad + b) * p.ldl : nullptr;
    __nv_bfloat16* drow16 = p.dlogits_bf16 ? p.dlogits_bf16 + ((size_t)t * p.Bpad + b) * p.ldl : nullptr;
    if (t >= Tb) {                                            // frames beyond the input length get zero gradient
      for (int c = lane; c < p.ldl; c += 32) {
        if (drow) drow[c] = 0.f;
        if (drow16) drow16[c] = __float2bfloat16_rn(0.f);
      }
      continue;
    }
    const float base = (float)(shiftA[t] + shiftB[t] - ll);
    for (int c = lane; c < p.C; c += 32) occ[c] = 0.f;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      const int s = lane * CPL + j;
      if (s < L) {
        const float ab = alpha[(size_t)t * Lmax + s] + beta[(size_t)t * Lmax + s] + base;   // log of the (doubly emitted) path mass through (t, s)
        const float w = expf(ab - lp[t * p.C + ext[j]]);
        if (w > 0.f) atomicAdd(&occ[ext[j]], w);
      }
    }
    __syncwarp();
    int k = 0;
    for (int c = lane; c < p.ldl; c += 32, ++k) {
      float g = 0.f;
      if (c < p.C) {
        g = (expf(lp[t * p.C + c]) - occ[c]) * p.grad_scale;
        if (k < 2) dbacc[k] += g;
      }
      if (drow) drow[c] = g;
      if (drow16) drow16[c] = __float2bfloat16_rn(g);
    }
    __syncwarp();
  }
  if (p.dbias) {
    int k = 0;
    for (int c = lane; c < p.C && k < 2; c += 32, ++k) atomicAdd(&dbsum[c], dbacc[k]);
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) atomicAdd(p.dbias + c, dbsum[c]);
  }
}

inline size_t ctc_par_smem_bytes(int T, int C) { return (size_t)2 * T * sizeof(double) + ((size_t)T * C + (CTC_PAR_WARPS + 1) * C) * sizeof(float) + 16; }

inline size_t ctc_warp_smem_bytes(int T, int C) { return (size_t)T * sizeof(double) + ((size_t)T * C + C) * sizeof(float) + 16; }

inline size_t ctc_smem_bytes(int T, int C, int Smax) {
  const int Lmax = 2 * Smax + 1;
  return (size_t)T * sizeof(double) + ((size_t)T * C + 4 * (size_t)Lmax) * sizeof(float) + 16;
}

}  // namespace b2t
