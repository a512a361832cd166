// Persistent GRU recurrence kernels for sm_100a (one launch per layer and direction).
//
// Reference semantics: torch.nn.GRU as used at rnn_model.py:65-72,126 (gate order r,z,n):
//   r = sig(gx_r + W_hr h + b_hr)   z = sig(gx_z + W_hz h + b_hz)
//   n = tanh(gx_n + r * (W_hn h + b_hn))      h' = (1-z) * n + z * h
// gx = x W_ih^T + b_ih is produced for all time steps by the tcgen05 GEMM (gemm.cuh).
//
// Decomposition.  The batch is cut into groups of 16 trials and the hidden units into slices of 32.
// CTA (slice s, group g) keeps the 96 rows {r,z,n} x 32 units of W_hh (bf16, 147 KB for H=768)
// resident in shared memory for the whole sequence and, per time step, computes
//     D[128 (96 used), 16] = W_slice[128, H] * h_{t-1}[16, H]^T
// with H/16 tcgen05.mma (M=128, N=16, K=16) into TMEM.  h_{t-1} is fetched by TMA as the K-major
// B operand (H/64 boxes of 16 rows x 128 B).  The three gate row-blocks land in TMEM lane
// quarters 0,1,2; the epilogue warps move them through a 6 KB smem exchange so that one thread
// owns (trial, 4 units) with all three gates, applies the gate math in fp32 (the hidden state itself
// is carried in fp32 registers across steps), and writes h_t (bf16) for the next step plus the
// activations BPTT needs.  Trials are independent, so only the H/32 CTAs of one batch group
// synchronise per step, through a release/acquire counter in global memory.
#pragma once
#include "sm100.cuh"

namespace b2t {

constexpr int REC_BG = 16;        // trials per batch group (UMMA N)
constexpr int REC_US = 32;        // hidden units per CTA
constexpr int REC_THREADS = 192;  // warp0 TMA, warp1 MMA, warps 2..5 epilogue
constexpr int REC_XPAD = 20;      // exchange row pitch (floats)

struct RecFwdParams {
  int H, T, Bpad;                 // hidden size, time steps, padded batch (multiple of 16)
  int n_slices;                   // H / 32
  const float* gx;                // [T][Bpad][3H] fp32, includes b_ih
  const float* bhh;               // [3H]
  __nv_bfloat16* hseq;            // [(T+1)][Bpad][H]; slot 0 = initial state, slot t+1 = h_t
  const float* h_init;            // [Bpad][H] fp32 initial state (register carry)
  float* h_final;                 // [Bpad][H] fp32 (nullable)
  __nv_bfloat16* hdrop;           // [T][Bpad][H] dropout(h_t) for the next layer (nullable => not written)
  __nv_bfloat16 *R, *Z, *Nn, *HN; // [T][Bpad][H] stash for BPTT (nullable when not training)
  int* done;                      // [n_groups][T] arrival counters, zeroed before launch
  float keep;                     // dropout keep prob for hdrop
  unsigned long long seed, rng_offset;
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void wait_counter(const int* ctr, int target) {
  uint32_t spins = 0;
  while (ld_acquire_gpu(ctr) < target) {
    if (++spins > (1u << 26)) __trap();
  }
}

// dropout decision for element (row m, unit j) of a [*, H] activation; same function in fwd and bwd.
__device__ __forceinline__ uint4 rec_dropout_bits(unsigned long long seed, unsigned long long offset, unsigned long long elem4) {
  const unsigned long long c = elem4 + offset;
  return philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x6a7eu, 0), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// tmap_w : W_hh bf16 [3H][H]   dims (H, 3H)          box (64, 32)
// tmap_h : hseq bf16          dims (H, (T+1)*Bpad)  box (64, 16)
__global__ void __launch_bounds__(REC_THREADS, 1)
gru_rec_fwd_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_h, const RecFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = p.H / 64;                               // contraction chunks
  uint8_t* sW = smem;                                    // KC x 12 KB (96 rows x 128 B)
  uint8_t* sH = sW + KC * 12288;                         // KC x 2 KB (16 rows x 128 B); also absorbs the M=128 over-read of sW
  float* sX = reinterpret_cast<float*>(sH + (KC * 2048 > 4096 ? KC * 2048 : 4096));   // [3][32][REC_XPAD]
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(sX + 3 * 32 * REC_XPAD);
  uint64_t* bar_h = bar_w + 1;                           // [KC] (<= 16)
  uint64_t* bar_d = bar_h + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_d + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.n_slices, grp = blockIdx.x / p.n_slices;
  const int j0 = slice * REC_US, b0 = grp * REC_BG;
  int* done = p.done + (size_t)grp * p.T;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_h);
    mbar_init(bar_w, 1);
    for (int c = 0; c < KC; ++c) mbar_init(&bar_h[c], 1);
    mbar_init(bar_d, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<32>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // resident weights: per chunk three 32-row boxes (r, z, n rows of this unit slice)
      mbar_arrive_expect_tx(bar_w, KC * 12288);
      for (int c = 0; c < KC; ++c)
        for (int g = 0; g < 3; ++g) tma_load_2d(sW + c * 12288 + g * 4096, &tmap_w, bar_w, c * 64, g * p.H + j0);
      for (int t = 0; t < p.T; ++t) {
        if (t > 0) {
          wait_counter(&done[t - 1], p.n_slices);
          fence_proxy_async_all();             // order the acquired generic-proxy writes before async-proxy reads
        }
        for (int c = 0; c < KC; ++c) {
          mbar_arrive_expect_tx(&bar_h[c], 2048);
          tma_load_2d(sH + c * 2048, &tmap_h, &bar_h[c], c * 64, t * p.Bpad + b0);   // slot t = h_{t-1}
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, REC_BG, 0, 0);
      mbar_wait(bar_w, 0);
      for (int t = 0; t < p.T; ++t) {
        for (int c = 0; c < KC; ++c) {
          mbar_wait(&bar_h[c], t & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(sW + c * 12288), sb = smem_u32(sH + c * 2048);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_d, umma_smem_desc(sa + k * 32, 16, 1024), umma_smem_desc(sb + k * 32, 16, 1024), idesc, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_d);
      }
    }
  } else {
    // ---------------- epilogue: 128 threads; thread e owns trial b0 + e/8 and units j0 + 4*(e%8) .. +3
    const int e = threadIdx.x - 64;
    const int q = warp & 3;                    // TMEM lane quarter: 0 -> r rows, 1 -> z, 2 -> n, 3 -> unused
    const int bl = e >> 3, u0 = (e & 7) * 4;
    const int b = b0 + bl, j = j0 + u0;
    float h[4], bh[3][4];
    {
      const float4 hv = *reinterpret_cast<const float4*>(p.h_init + (size_t)b * p.H + j);
      h[0] = hv.x; h[1] = hv.y; h[2] = hv.z; h[3] = hv.w;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const float4 bv = *reinterpret_cast<const float4*>(p.bhh + g * p.H + j);
        bh[g][0] = bv.x; bh[g][1] = bv.y; bh[g][2] = bv.z; bh[g][3] = bv.w;
      }
    }
    const bool train = p.R != nullptr;
    const float inv_keep = 1.0f / p.keep;
    for (int t = 0; t < p.T; ++t) {
      const size_t row = (size_t)t * p.Bpad + b;
      // prefetch the input projection while the MMA runs
      float4 gxv[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) gxv[g] = __ldg(reinterpret_cast<const float4*>(p.gx + row * 3 * p.H + g * p.H + j));

      mbar_wait(bar_d, t & 1);
      tc_fence_after();
      if (q < 3) {
        uint32_t v[16];
        tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(sX + (q * 32 + lane) * REC_XPAD);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
      tc_fence_before();
      epi_bar_sync();
      float hn[4], r[4], z[4], n[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ar = sX[(0 * 32 + u0 + i) * REC_XPAD + bl];
        const float az = sX[(1 * 32 + u0 + i) * REC_XPAD + bl];
        const float an = sX[(2 * 32 + u0 + i) * REC_XPAD + bl];
        const float gr = (&gxv[0].x)[i], gz = (&gxv[1].x)[i], gn = (&gxv[2].x)[i];
        r[i] = sigmoid_f(gr + ar + bh[0][i]);
        z[i] = sigmoid_f(gz + az + bh[1][i]);
        hn[i] = an + bh[2][i];
        n[i] = tanh_f(gn + r[i] * hn[i]);
        h[i] = (1.0f - z[i]) * n[i] + z[i] * h[i];
      }
      const size_t off = row * p.H + j;
      st_bf16x4(p.hseq + ((size_t)(t + 1) * p.Bpad + b) * p.H + j, h[0], h[1], h[2], h[3]);
      if (train) {
        st_bf16x4(p.R + off, r[0], r[1], r[2], r[3]);
        st_bf16x4(p.Z + off, z[0], z[1], z[2], z[3]);
        st_bf16x4(p.Nn + off, n[0], n[1], n[2], n[3]);
        st_bf16x4(p.HN + off, hn[0], hn[1], hn[2], hn[3]);
      }
      if (p.hdrop) {
        float d[4] = {h[0], h[1], h[2], h[3]};
        if (p.keep < 1.0f) {
          const uint4 rnd = rec_dropout_bits(p.seed, p.rng_offset, off >> 2);
          const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) d[i] = (u32_to_unit(rr[i]) < p.keep) ? d[i] * inv_keep : 0.0f;
        }
        st_bf16x4(p.hdrop + off, d[0], d[1], d[2], d[3]);
      }
      // publish h_t: generic stores -> (proxy fence, gpu fence) -> all epilogue threads done -> one release
      fence_proxy_async_all();
      __threadfence();
      epi_bar_sync();
      if (e == 0) red_release_add(&done[t], 1);
    }
    if (p.h_final) *reinterpret_cast<float4*>(p.h_final + (size_t)b * p.H + j) = make_float4(h[0], h[1], h[2], h[3]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_d);
  }
}

// --------------------------------------------------------------------------------------------
// Backward recurrence (BPTT).  Per step t (descending), with dh_t the total gradient wrt h_t:
//   dn = dh*(1-z)  dz = dh*(h_{t-1}-n)  dn_pre = dn*(1-n^2)  dz_pre = dz*z*(1-z)
//   dr_pre = dn_pre*hn*r*(1-r)
//   dGx_t = [dr_pre, dz_pre, dn_pre]        dGh_t = [dr_pre, dz_pre, dn_pre*r]
//   dh_{t-1} = dh*z + dGh_t W_hh + dY_{t-1}
// CTA (slice, group) owns 32 hidden units and keeps W_hh^T[32 units][3H] (147 KB) resident as the
// A operand; the B operand is dGh_t[16 trials][3H] (K-major), fetched by TMA once every CTA of the
// batch group has published its 96 columns of it.  dW_ih/dW_hh are GEMMs over dGx/dGh afterwards;
// the bias gradients are accumulated here in registers and reduced with one atomicAdd per thread.
struct RecBwdParams {
  int H, T, Bpad, n_slices;
  const float* dY;                  // [T][Bpad][H] fp32 gradient wrt this layer's (dropped) output
  const __nv_bfloat16* hseq;        // [(T+1)][Bpad][H]
  const __nv_bfloat16 *R, *Z, *Nn, *HN;
  __nv_bfloat16* dGx;               // [T][Bpad][3H]
  __nv_bfloat16* dGh;               // [T][Bpad][3H]
  float* dbih;                      // [3H] (atomicAdd)
  float* dbhh;                      // [3H] (atomicAdd)
  float* dh0;                       // [Bpad][H] gradient wrt the initial state (written)
  int* done;                        // [n_groups][T]
  int n_valid;                      // trials < n_valid contribute (pad trials are masked out)
  float keep;                       // dropout applied to this layer's output in forward (1 => none)
  unsigned long long seed, rng_offset;
};

// tmap_wt : W_hh^T bf16 [H][3H]  dims (3H, H)        box (64, 32)
// tmap_g  : dGh bf16            dims (3H, T*Bpad)   box (64, 16)
__global__ void __launch_bounds__(REC_THREADS, 1)
gru_rec_bwd_kernel(const __grid_constant__ CUtensorMap tmap_wt, const __grid_constant__ CUtensorMap tmap_g, const RecBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = 3 * p.H / 64;                           // contraction chunks over the 3H gate columns
  uint8_t* sW = smem;                                    // KC x 4 KB (32 rows x 128 B)
  uint8_t* sG = sW + KC * 4096;                          // KC x 2 KB; also absorbs the 12 KB M=128 over-read
  float* sX = reinterpret_cast<float*>(sG + (KC * 2048 > 12288 ? KC * 2048 : 12288));  // [32][REC_XPAD]
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(sX + 32 * REC_XPAD);
  uint64_t* bar_g = bar_w + 1;                           // [KC] (<= 48)
  uint64_t* bar_d = bar_g + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_d + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.n_slices, grp = blockIdx.x / p.n_slices;
  const int j0 = slice * REC_US, b0 = grp * REC_BG;
  int* done = p.done + (size_t)grp * p.T;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_wt);
    tma_prefetch_desc(&tmap_g);
    mbar_init(bar_w, 1);
    for (int c = 0; c < KC; ++c) mbar_init(&bar_g[c], 1);
    mbar_init(bar_d, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<32>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // Steps are indexed s = 0..T-1 for t = T-1-s.  MMA s (s >= 1) consumes dGh_{t+1} and feeds dh_t.
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_w, KC * 4096);
      for (int c = 0; c < KC; ++c) tma_load_2d(sW + c * 4096, &tmap_wt, bar_w, c * 64, j0);
      for (int s = 1; s <= p.T; ++s) {                   // s == T: extra product for the initial-state gradient
        const int tsrc = p.T - s;                        // dGh_{t+1}
        wait_counter(&done[tsrc], p.n_slices);
        fence_proxy_async_all();
        for (int c = 0; c < KC; ++c) {
          mbar_arrive_expect_tx(&bar_g[c], 2048);
          tma_load_2d(sG + c * 2048, &tmap_g, &bar_g[c], c * 64, tsrc * p.Bpad + b0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, REC_BG, 0, 0);
      mbar_wait(bar_w, 0);
      for (int s = 1; s <= p.T; ++s) {
        for (int c = 0; c < KC; ++c) {
          mbar_wait(&bar_g[c], (s - 1) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(sW + c * 4096), sb = smem_u32(sG + c * 2048);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_d, umma_smem_desc(sa + k * 32, 16, 1024), umma_smem_desc(sb + k * 32, 16, 1024), idesc, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_d);
      }
    }
  } else {
    const int e = threadIdx.x - 64;
    const int q = warp & 3;
    const int bl = e >> 3, u0 = (e & 7) * 4;
    const int b = b0 + bl, j = j0 + u0;
    const bool valid = b < p.n_valid;
    const float inv_keep = 1.0f / p.keep;
    float carry[4] = {0.f, 0.f, 0.f, 0.f};               // dh_{t+1} * z_{t+1}
    float accx[3][4], acch[4];                           // bias-gradient partial sums (dGh differs only in n)
#pragma unroll
    for (int i = 0; i < 4; ++i) { accx[0][i] = accx[1][i] = accx[2][i] = 0.f; acch[i] = 0.f; }
    for (int s = 0; s < p.T; ++s) {
      const int t = p.T - 1 - s;
      const size_t row = (size_t)t * p.Bpad + b;
      const size_t off = row * p.H + j;
      float r[4], z[4], n[4], hn[4], hp[4];
      ld_bf16x4(p.R + off, r); ld_bf16x4(p.Z + off, z); ld_bf16x4(p.Nn + off, n); ld_bf16x4(p.HN + off, hn);
      ld_bf16x4(p.hseq + off, hp);                       // slot t = h_{t-1}
      float4 dyv = __ldg(reinterpret_cast<const float4*>(p.dY + off));
      float dy[4] = {dyv.x, dyv.y, dyv.z, dyv.w};
      if (p.keep < 1.0f) {
        const uint4 rnd = rec_dropout_bits(p.seed, p.rng_offset, off >> 2);
        const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) dy[i] = (u32_to_unit(rr[i]) < p.keep) ? dy[i] * inv_keep : 0.0f;
      }
      float dh[4];
      if (s > 0) {
        mbar_wait(bar_d, (s - 1) & 1);
        tc_fence_after();
        if (q == 0) {
          uint32_t v[16];
          tmem_ld16(tmem_d, v);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(sX + lane * REC_XPAD);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
        tc_fence_before();
        epi_bar_sync();
#pragma unroll
        for (int i = 0; i < 4; ++i) dh[i] = carry[i] + sX[(u0 + i) * REC_XPAD + bl] + dy[i];
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) dh[i] = dy[i];
      }
      float gr[4], gz[4], gn[4], gnh[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = valid ? dh[i] : 0.0f;
        const float dn = d * (1.0f - z[i]);
        const float dz = d * (hp[i] - n[i]);
        gn[i] = dn * (1.0f - n[i] * n[i]);
        gz[i] = dz * z[i] * (1.0f - z[i]);
        gr[i] = gn[i] * hn[i] * r[i] * (1.0f - r[i]);
        gnh[i] = gn[i] * r[i];
        carry[i] = d * z[i];
        accx[0][i] += gr[i]; accx[1][i] += gz[i]; accx[2][i] += gn[i]; acch[i] += gnh[i];
      }
      const size_t goff = row * 3 * p.H + j;
      st_bf16x4(p.dGx + goff, gr[0], gr[1], gr[2], gr[3]);
      st_bf16x4(p.dGx + goff + p.H, gz[0], gz[1], gz[2], gz[3]);
      st_bf16x4(p.dGx + goff + 2 * p.H, gn[0], gn[1], gn[2], gn[3]);
      st_bf16x4(p.dGh + goff, gr[0], gr[1], gr[2], gr[3]);
      st_bf16x4(p.dGh + goff + p.H, gz[0], gz[1], gz[2], gz[3]);
      st_bf16x4(p.dGh + goff + 2 * p.H, gnh[0], gnh[1], gnh[2], gnh[3]);
      fence_proxy_async_all();
      __threadfence();
      epi_bar_sync();
      if (e == 0) red_release_add(&done[t], 1);
    }
    // gradient wrt the initial state: dh_{-1} = dh_0 * z_0 + dGh_0 W_hh (the extra product s == T)
    {
      mbar_wait(bar_d, (p.T - 1) & 1);
      tc_fence_after();
      if (q == 0) {
        uint32_t v[16];
        tmem_ld16(tmem_d, v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(sX + lane * REC_XPAD);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
      tc_fence_before();
      epi_bar_sync();
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = valid ? carry[i] + sX[(u0 + i) * REC_XPAD + bl] : 0.0f;
      *reinterpret_cast<float4*>(p.dh0 + (size_t)b * p.H + j) = make_float4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(p.dbih + j + i, accx[0][i]);
      atomicAdd(p.dbih + p.H + j + i, accx[1][i]);
      atomicAdd(p.dbih + 2 * p.H + j + i, accx[2][i]);
      atomicAdd(p.dbhh + j + i, accx[0][i]);
      atomicAdd(p.dbhh + p.H + j + i, accx[1][i]);
      atomicAdd(p.dbhh + 2 * p.H + j + i, acch[i]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<32>(tmem_d);
  }
}

}  // namespace b2t
