// Bandwidth-bound kernels around the GEMMs: input augmentation + Gaussian smoothing, patch
// materialisation (fallback), patch fold + day-layer backward, transposes, reductions,
// greedy CTC decode + edit distance.
#pragma once
#include "sm100.cuh"

namespace b2t {

// ---------------------------------------------------------------------------------------------
// transform_data (rnn_trainer.py:436-484) + gauss_smooth (data_augmentations.py:6-37), fused.
//   xin[b,t,d] = x[b,t+cut,d] + white_std * N(0,1)[b,t+cut,d] + offset_std * N(0,1)[b,d]      (train)
//   out[b,t,d] = sum_j taps[j] * xin[b, t + j - left, d]      zero outside [0, T_len)
// 'same': left = (ntaps-1)/2, T_out = T_len;   'valid': left = 0, T_out = T_len - ntaps + 1.
// The noise is added to the padded tensor (pad frames of short trials receive noise too), exactly
// as the reference does.  Noise comes from Philox (Box-Muller) unless explicit arrays are given.
struct PreParams {
  const float* x;          // [B][T_in][D]
  __nv_bfloat16* out;      // [Bpad][T_alloc][D]; rows t >= T_out and trials >= B are zero-filled
  float* out_f32;          // when non-null: fp32 output [B][T_alloc][D] instead (stand-alone gauss_smooth)
  int B, Bpad, T_in, T_alloc, D;
  int cut;                 // frames dropped from the front (random_cut)
  int ntaps, valid;        // ntaps == 0 => no smoothing
  float taps[16];          // right-aligned: taps[16 - ntaps + j] = tap j, leading zeros ({..,0,1} when ntaps == 0)
  float white_std, offset_std;
  const float* white;      // optional explicit N(0,1) draws [B][T_in][D]
  const float* offset;     // optional explicit N(0,1) draws [B][D]
  int use_philox;          // draw noise on device when the explicit arrays are null and std > 0
  unsigned long long seed, rng_offset;
};

__device__ __forceinline__ void box_muller4(uint4 r, float (&n)[4]) {
  const float u0 = (r.x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
  const float u1 = (r.y >> 8) * (1.0f / 16777216.0f);
  const float u2 = (r.z >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f);
  const float u3 = (r.w >> 8) * (1.0f / 16777216.0f);
  const float m0 = sqrtf(-2.0f * __logf(u0)), m1 = sqrtf(-2.0f * __logf(u2));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  n[0] = m0 * c; n[1] = m0 * s;
  __sincosf(6.283185307179586f * u3, &s, &c);
  n[2] = m1 * c; n[3] = m1 * s;
}

constexpr int PRE_TT = 16;   // outputs per thread along time

// grid: (ceil(T_alloc / PRE_TT), Bpad), block: D/4 threads; thread owns 4 consecutive channels.
__global__ void pre_smooth_kernel(const PreParams p) {
  const int b = blockIdx.y;
  const int d = threadIdx.x * 4;
  const int t0 = blockIdx.x * PRE_TT;
  if (d >= p.D) return;
  __nv_bfloat16* out = p.out + ((size_t)b * p.T_alloc + t0) * p.D + d;
  float* out32 = p.out_f32 ? p.out_f32 + ((size_t)b * p.T_alloc + t0) * p.D + d : nullptr;
  const int T_len = p.T_in - p.cut;
  const int nt = p.ntaps > 0 ? p.ntaps : 1;
  const int left = (p.ntaps > 0 && !p.valid) ? (p.ntaps - 1) / 2 : 0;
  const int T_out = (p.ntaps > 0 && p.valid) ? T_len - p.ntaps + 1 : T_len;
  if (b >= p.B) {
    if (!out32)
      for (int i = 0; i < PRE_TT && t0 + i < p.T_alloc; ++i) *reinterpret_cast<uint2*>(out + (size_t)i * p.D) = make_uint2(0, 0);
    return;
  }
  float off[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.offset_std > 0.f) {
    if (p.offset) {
      const float4 o = *reinterpret_cast<const float4*>(p.offset + (size_t)b * p.D + d);
      off[0] = o.x; off[1] = o.y; off[2] = o.z; off[3] = o.w;
    } else if (p.use_philox) {
      const unsigned long long c = ((unsigned long long)b * p.D + d) / 4 + p.rng_offset;
      box_muller4(philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x0ff5u, 0), make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32))), off);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) off[i] *= p.offset_std;
  }
  float win[16][4];   // sliding window of the last `nt` inputs
#pragma unroll
  for (int i = 0; i < 16; ++i) win[i][0] = win[i][1] = win[i][2] = win[i][3] = 0.f;
  // inputs needed: xin[t0 - left .. t0 + PRE_TT - 1 - left + nt - 1]
  for (int i = 0; i < PRE_TT + nt - 1; ++i) {
    const int tl = t0 - left + i;                       // index into the (cut) sequence
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (tl >= 0 && tl < T_len) {
      const size_t e = ((size_t)b * p.T_in + tl + p.cut) * p.D + d;
      const float4 xv = *reinterpret_cast<const float4*>(p.x + e);
      v[0] = xv.x; v[1] = xv.y; v[2] = xv.z; v[3] = xv.w;
      if (p.white_std > 0.f) {
        float nz[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.white) {
          const float4 w = *reinterpret_cast<const float4*>(p.white + e);
          nz[0] = w.x; nz[1] = w.y; nz[2] = w.z; nz[3] = w.w;
        } else if (p.use_philox) {
          const unsigned long long c = e / 4 + p.rng_offset;
          box_muller4(philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x0a15u, 0), make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32))), nz);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += p.white_std * nz[k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] += off[k];
    }
    // shift window (static indexing keeps it in registers)
#pragma unroll
    for (int w = 0; w < 15; ++w) {
#pragma unroll
      for (int k = 0; k < 4; ++k) win[w][k] = win[w + 1][k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) win[15][k] = v[k];
    const int to = i - (nt - 1);                        // output index within the tile
    if (to >= 0 && t0 + to < p.T_alloc) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (t0 + to < T_out) {
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          const float tp = p.taps[w];                    // right-aligned: taps[16 - nt + j] = tap j
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] += tp * win[w][k];
        }
      }
      if (out32) *reinterpret_cast<float4*>(out32 + (size_t)to * p.D) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      else st_bf16x4(out + (size_t)to * p.D, acc[0], acc[1], acc[2], acc[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fallback materialisation of the patch view:  xu[(t',b)][p*D+d] = xd[b][t'*stride + p][d].
__global__ void unfold_kernel(const __nv_bfloat16* __restrict__ xd, __nv_bfloat16* __restrict__ xu, int Bpad, int T_alloc, int D, int Tp,
                              int patch, int stride) {
  const size_t n8 = (size_t)Tp * Bpad * patch * D / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * 8;
    const int k = (int)(e % ((size_t)patch * D));
    const size_t m = e / ((size_t)patch * D);
    const int b = (int)(m % Bpad), tp = (int)(m / Bpad);
    const int pp = k / D, d = k % D;
    reinterpret_cast<uint4*>(xu)[i] = *reinterpret_cast<const uint4*>(xd + ((size_t)b * T_alloc + (size_t)tp * stride + pp) * D + d);
  }
}

// ---------------------------------------------------------------------------------------------
// Patch fold + day-layer backward (rnn_model.py:95-119 reversed):
//   dxd[b,t,d]  = sum_{p : (t-p) % stride == 0, 0 <= (t-p)/stride < Tp} dXu[((t-p)/stride, b)][p*D + d]
//   dpre[b,t,d] = dxd * dropmask/keep * (1 - |y|)^2          y = softsign output (pre-dropout)
//   dbias_day[day[b]][d] += sum_t dpre[b,t,d]
struct FoldParams {
  const __nv_bfloat16* dxu;   // [Tp][Bpad][patch*D]
  const __nv_bfloat16* xd;    // [Bpad][T_alloc][D] (post-dropout day-layer output)
  __nv_bfloat16* dpre;        // [Bpad][T_alloc][D]
  float* dbias_day;           // [n_days][bias_pitch] (atomicAdd)
  int bias_pitch;
  const int* day_idx;         // [B]
  int B, Bpad, T_alloc, T_valid, D, Tp, patch, stride;
  float keep;
  unsigned long long seed, rng_offset;
};
constexpr int FOLD_TT = 16;
// grid (ceil(T_alloc/FOLD_TT), Bpad), block D/4
__global__ void fold_dpre_kernel(const FoldParams p) {
  const int b = blockIdx.y, d = threadIdx.x * 4, t0 = blockIdx.x * FOLD_TT;
  if (d >= p.D) return;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const float inv_keep = 1.0f / p.keep;
  for (int i = 0; i < FOLD_TT; ++i) {
    const int t = t0 + i;
    if (t >= p.T_alloc) break;
    const size_t e = ((size_t)b * p.T_alloc + t) * p.D + d;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (b < p.B && t < p.T_valid) {
      for (int pp = t % p.stride; pp < p.patch; pp += p.stride) {
        const int tp = (t - pp) / p.stride;
        if (t - pp < 0 || tp >= p.Tp) continue;
        float v[4];
        ld_bf16x4(p.dxu + ((size_t)tp * p.Bpad + b) * p.patch * p.D + (size_t)pp * p.D + d, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) g[k] += v[k];
      }
      float y[4];
      ld_bf16x4(p.xd + e, y);
      uint32_t rr[4] = {0, 0, 0, 0};
      if (p.keep < 1.0f) {
        const unsigned long long c = (e >> 2) + p.rng_offset;
        const uint4 rnd = philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x0da1u, 0), make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
        rr[0] = rnd.x; rr[1] = rnd.y; rr[2] = rnd.z; rr[3] = rnd.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float yy = y[k];
        float m = 1.0f;
        if (p.keep < 1.0f) {
          const bool kept = u32_to_unit(rr[k]) < p.keep;
          m = kept ? inv_keep : 0.0f;
          yy = yy * p.keep;                 // undo the 1/keep scaling to recover the softsign output
        }
        const float s = 1.0f - fabsf(yy);
        g[k] = g[k] * m * s * s;
        bsum[k] += g[k];
      }
    }
    st_bf16x4(p.dpre + e, g[0], g[1], g[2], g[3]);
  }
  if (b < p.B && p.dbias_day) {
    float* dst = p.dbias_day + (size_t)p.day_idx[b] * p.bias_pitch + d;
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(dst + k, bsum[k]);
  }
}

// ---------------------------------------------------------------------------------------------
// bf16 transpose [R][C] -> [C][R] (W_hh -> W_hh^T after each optimizer step)
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = in[(size_t)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) out[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

// fp32 -> bf16 flat cast
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __float2bfloat16_rn(in[i]);
}

// initial state: hseq slot 0 (bf16) and the fp32 register seed, from h0 (broadcast) or explicit states
__global__ void init_state_kernel(const float* __restrict__ h0, const float* __restrict__ states, int B, int Bpad, int H,
                                  __nv_bfloat16* __restrict__ slot0, float* __restrict__ h_init) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bpad * H) return;
  const int b = i / H, j = i % H;
  const float v = (states && b < B) ? states[(size_t)b * H + j] : h0[j];
  slot0[i] = __float2bfloat16_rn(v);
  h_init[i] = v;
}

// g_h0[j] += sum_{b < B} dh0[b][j]
__global__ void reduce_dh0_kernel(const float* __restrict__ dh0, int B, int H, float* __restrict__ g) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += dh0[(size_t)b * H + j];
  atomicAdd(g + j, s);
}

// logits [T][Bpad][ldl] -> out [B][T][C]
__global__ void gather_logits_kernel(const float* __restrict__ lg, int T, int B, int Bpad, int ldl, int C, float* __restrict__ out) {
  const size_t n = (size_t)B * T * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % T);
    const int b = (int)(i / ((size_t)C * T));
    out[i] = lg[((size_t)t * Bpad + b) * ldl + c];
  }
}
// inverse (for feeding externally computed dlogits [B][T][C] into the backward pass)
__global__ void scatter_dlogits_kernel(const float* __restrict__ in, int T, int B, int Bpad, int ldl, int C, float* __restrict__ out32,
                                       __nv_bfloat16* __restrict__ out16) {
  const size_t n = (size_t)T * Bpad * ldl;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ldl);
    const int b = (int)((i / ldl) % Bpad);
    const int t = (int)(i / ((size_t)ldl * Bpad));
    const float v = (b < B && c < C) ? in[((size_t)b * T + t) * C + c] : 0.f;
    if (out32) out32[i] = v;
    out16[i] = __float2bfloat16_rn(v);
  }
}
__global__ void colsum_kernel(const float* __restrict__ in, size_t rows, int ld, int C, float* __restrict__ out) {
  const int c = threadIdx.x;
  if (c >= C) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // four independent loads in flight (the column walk is latency-bound)
  for (size_t r = (size_t)blockIdx.x * 4; r < rows; r += (size_t)gridDim.x * 4) {
    s0 += in[r * ld + c];
    if (r + 1 < rows) s1 += in[(r + 1) * ld + c];
    if (r + 2 < rows) s2 += in[(r + 2) * ld + c];
    if (r + 3 < rows) s3 += in[(r + 3) * ld + c];
  }
  atomicAdd(out + c, (s0 + s1) + (s2 + s3));
}

// ---------------------------------------------------------------------------------------------
// Greedy CTC decode + Levenshtein distance (rnn_trainer.py:724-736): integer work, bit-exact.
// One CTA per trial; thread 0 does the serial collapse and the DP (validation-only path).
struct GreedyParams {
  const float* logits;   // [T][Bpad][ldl]
  int T, Bpad, ldl, C;
  const int* in_len;     // [B]
  const int* labels;     // [B][Smax]
  const int* tgt_len;    // [B]
  int Smax;
  int* decoded;          // [B][T] collapsed ids, -1 padded
  int* dec_len;          // [B]
  int* edit;             // [B]
  int* scratch;          // [B][2*(Smax+1)]
};
__global__ void greedy_edit_kernel(const GreedyParams p) {
  extern __shared__ int g_arg[];   // [T]
  const int b = blockIdx.x;
  int Tb = p.in_len[b];
  Tb = Tb < 0 ? 0 : (Tb > p.T ? p.T : Tb);
  for (int t = threadIdx.x; t < Tb; t += blockDim.x) {
    const float* row = p.logits + ((size_t)t * p.Bpad + b) * p.ldl;
    int best = 0;
    float bv = row[0];
    for (int c = 1; c < p.C; ++c)
      if (row[c] > bv) { bv = row[c]; best = c; }   // first maximum wins, as torch.argmax
    g_arg[t] = best;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  int* dec = p.decoded + (size_t)b * p.T;
  int n = 0, prev = -1;
  for (int t = 0; t < Tb; ++t) {
    const int a = g_arg[t];
    if (a != prev && a != 0) dec[n++] = a;
    prev = a;
  }
  for (int t = n; t < p.T; ++t) dec[t] = -1;
  p.dec_len[b] = n;
  const int S = p.tgt_len[b];
  const int* lab = p.labels + (size_t)b * p.Smax;
  int* r0 = p.scratch + (size_t)b * 2 * (p.Smax + 1);
  int* r1 = r0 + p.Smax + 1;
  for (int j = 0; j <= S; ++j) r0[j] = j;
  for (int i = 1; i <= n; ++i) {
    r1[0] = i;
    for (int j = 1; j <= S; ++j) {
      const int sub = r0[j - 1] + (dec[i - 1] != lab[j - 1]);
      const int del = r0[j] + 1, ins = r1[j - 1] + 1;
      r1[j] = min(sub, min(del, ins));
    }
    int* tmp = r0; r0 = r1; r1 = tmp;
  }
  p.edit[b] = r0[S];
}

}  // namespace b2t
