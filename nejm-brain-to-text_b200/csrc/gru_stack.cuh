// Whole-stack persistent GRU recurrence for sm_100a: ONE cooperative launch per direction runs every layer and every
// time step (the chunked per-layer launches of gru_rec.cuh remain as the fall-back for shapes that do not fit).
//
// Reference semantics: torch.nn.GRU(num_layers=L, dropout=p) as used at rnn_model.py:65-72,126 and its autograd backward.
//
// Why.  The per-(layer, time chunk) launches of gru_rec.cuh run a wave-front of (chunks + L - 1) stages, each a third of the
// sequence long plus launch / weight-fill overhead: 231 serial recurrence steps for T' = 97, L = 5.  Here all L layers are
// resident at once and layer l trails layer l-1 by only the few steps its input projection needs, so the serial chain is
// T' + (L-1) * lag steps.  To make L layers fit on 148 SMs every CTA hosts NSUB = 2 independent batch groups of BG trials
// (L * H/32 CTAs for a batch of 2 * BG): the two groups alternate on the CTA's tensor core, loader warps and epilogue
// warps, so that the L2 hand-over latency of one group (publish -> poll -> stage) is covered by the MMAs of the other.
//
// Layer coupling (forward).  Layer l >= 1 needs gx_l[t] = W_ih_l x_l[t] + b_ih with x_l[t] = dropout(h_{l-1}[t]).  That GEMM
// runs on the SMs the recurrence leaves free, as ONE gated launch of the tcgen05 GEMM per layer (gemm.cuh): its TMA
// producer waits, per 128-row tile (= two time steps at 64 trials), on the progress counter prog_{l-1}[t] that the
// epilogue warps of layer l-1 bump after their h_t / dropout(h_t) stores (release), and its epilogue bumps done_l[tile]
// after storing gx; the epilogue warps of layer l wait for done_l[tile] (acquire) before they read gx_l[t].
// Backward is the mirror image: layer l needs dY_l[t] = dGx_{l+1}[t] W_ih_{l+1}, produced tile by tile (time descending)
// by a gated data-gradient GEMM that follows prog_{l+1}[t].
//
// Inside a layer the per-step exchange is unchanged from gru_rec.cuh / gru_rec_bwd2.cuh: the data is the signal
// (sentinel-filled hseq / dGh polled by loader warps, generation-tagged fp32 partial sums).
#pragma once
#include "gru_rec.cuh"
#include "gru_rec_bwd2.cuh"

namespace b2t {

constexpr int STACK_MAX_LAYERS = 8;

struct StackFwdLayer {
  const float* gx;                // [T][Bpad][3H] fp32 incl. b_ih (layer 0: complete before launch; layer >= 1: gated GEMM)
  const float* bhh;               // [3H]
  const __nv_bfloat16* whh;       // [3H][H]
  __nv_bfloat16* hseq;            // [(T+1)][Bpad][H], slot 0 = initial state, the rest sentinel-filled
  float* h_state;                 // [Bpad][H] fp32 initial state in, final state out
  __nv_bfloat16* hdrop;           // [T][Bpad][H] dropout(h_t) for the next layer (nullable)
  __nv_bfloat16 *R, *Z, *Nn, *HN; // BPTT stash (nullable)
  const int* gx_done;             // per 128-row tile of gx: completion count of the gated input GEMM (null: no gating)
  int* prog;                      // [T]: += 1 per epilogue warp, CTA and batch group once step t's outputs are stored (null: no consumer)
  int gx_need;                    // value gx_done[tile] reaches when the tile is complete
  float keep;                     // dropout keep probability of hdrop
  unsigned long long rng_offset;
  unsigned char* kmask;           // [T][Bpad][H/4] keep bits of hdrop (bit i = unit 4k+i kept), re-read by the backward kernel (nullable)
};

struct StackFwdParams {
  int H, Bpad, T, n_slices, n_layers, n_cgroups;   // n_cgroups: CTA-level batch groups (each CTA hosts NSUB groups of BG trials)
  int poll_delay;
  int trace_cta;                  // CTA whose group 0 writes the cycle trace
  unsigned long long seed;
  long long* trace;
  int use_tma;                    // stage h_{t-1} with TMA tensor loads (after the probe) and verify it in shared memory, instead of polled 16-byte loads
  StackFwdLayer lay[STACK_MAX_LAYERS];
  CUtensorMap tm_h[STACK_MAX_LAYERS];   // hseq of every layer as [(T+1)*Bpad rows][H], box {64, BG}, SWIZZLE_128B
};

template <int BG, int NSUB> struct StackCfg {
  using Base = RecCfg<BG>;
  static constexpr int kThreads = Base::kFwdThreads + 32;      // + one warp whose lane 0 publishes the layer's progress
  static constexpr int bwd_threads(int esets) { return kThreads + (esets - 1) * Base::kEpiThreads; }
  static constexpr int fwd_threads(int lsets) { return kThreads + (lsets - 1) * Base::kLoadThreads; }
  static constexpr size_t fwd_smem_bytes(int H) {
    return (size_t)NSUB * ((size_t)2 * (H / 64) * BG * 128 + (size_t)3 * 32 * Base::kXPitch * 4) + (size_t)(NSUB * 32 + 4 * NSUB) * 8 + 64 + 1024;
  }
  static constexpr size_t bwd_smem_bytes(int H) {
    return (size_t)(3 * H / 4 / 64) * (128 * 128 + (size_t)NSUB * BG * 128) + (size_t)(NSUB * 16 + 3 * NSUB) * 8 + 64 + 1024;
  }
};

// Bounded acquire-wait on a gpu-scope counter by lane 0 of the calling warp; the other lanes are released by the warp barrier
// (their later loads are ordered behind lane 0's acquire through it).
__device__ __forceinline__ void warp_wait_ge(const int* ctr, int need, int lane) {
  if (lane == 0) {
    uint32_t spins = 0;
    while (ld_acquire_gpu(ctr) < need) {
      __nanosleep(32);
      if (++spins > (1u << 26)) __trap();
    }
  }
  __syncwarp();
}
__device__ __forceinline__ bool warp_test_ge(const int* ctr, int need, int lane) {
  int ok = 0;
  if (lane == 0) ok = ld_acquire_gpu(ctr) >= need;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
// Release: every lane's earlier global stores become visible at gpu scope before the counter moves.
__device__ __forceinline__ void warp_signal(int* ctr, int lane) {
  __threadfence();
  __syncwarp();
  if (lane == 0) red_release_add(ctr, 1);
}
// CTA-local monotonic counter in shared memory (no phase parity to alias when the waiter falls behind)
__device__ __forceinline__ void smem_release_inc(uint32_t* ctr) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(ctr)) : "memory");
}
__device__ __forceinline__ uint32_t smem_acquire_ld(const uint32_t* ctr) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(ctr)) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldcg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// Probe before staging.  A full polling round moves the whole operand (48 KB per CTA at BG = 32) through L2 whether or not the
// peers have published; with every layer resident that wasted traffic starves the gated GEMMs.  So the loader threads first
// poll one 16-byte unit per (producer CTA, producer epilogue warp[, gate]) -- a few KB per round -- and only when every probe
// has been seen (named barrier over the loader warps) do they stage the operand, which still checks every word.
template <int NLOAD_THREADS, typename AddrF>
__device__ __forceinline__ void probe_until_published(int n_probe, int lt, AddrF addr, int bar_id = 2) {
  for (int i = lt; i < n_probe; i += NLOAD_THREADS) {
    const uint8_t* a = addr(i);
    uint32_t spins = 0;
    while (has_sentinel(ld_l2_v4(a))) {
      __nanosleep(40);
      if (++spins > REC_MAX_SPINS) __trap();
    }
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NLOAD_THREADS) : "memory");
}

// A load ptxas may not hoist out of a polling loop (it does hoist the weak ld.global.cg when the loop holds nothing else).
__device__ __forceinline__ uint4 ld_strong_v4(const void* ptr) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ long long stk_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// cross-layer timing (profiling aid, with the cycle trace): the first CTA of every layer stamps %globaltimer when it has stored
// steps 0, 1, T/2 and T-1 -> trace[2*T*8 + dir*64 + layer*8 + k]
#define STK_LAYER_STAMP(dir, first_cta, step)                                                                          \
  do {                                                                                                                 \
    if (p.trace != nullptr && (first_cta)) {                                                                           \
      const int _k = (step) == 0 ? 1 : (step) == 1 ? 2 : (step) == p.T / 2 ? 3 : (step) == p.T - 1 ? 4 : -1;           \
      if (_k > 0) p.trace[(size_t)2 * p.T * 8 + (dir) * 64 + layer * 8 + _k] = stk_globaltimer();                      \
    }                                                                                                                  \
  } while (0)

// ------------------------------------------------------------------------------------------------ forward
// LSETS = 2 (with NSUB = 2): each batch group gets its own set of loader warps, so the staging of one group's h_{t-1} no longer
// waits behind the staging of the other group's (measured with one set: "h_t stored -> next chunk staged" 5 300 cycles, of
// which 2 800 are the other group's staging).
template <int BG, int NSUB, int LSETS = 1>
__global__ void __launch_bounds__(RecCfg<BG>::kFwdThreads + 32 + (LSETS - 1) * RecCfg<BG>::kLoadThreads, 1)
gru_stack_fwd_kernel(const __grid_constant__ StackFwdParams p) {
  static_assert(LSETS == 1 || (LSETS == 2 && NSUB == 2), "one loader set, or one per batch group");
  using Cfg = RecCfg<BG>;
  constexpr int NTHREADS = Cfg::kFwdThreads + 32 + (LSETS - 1) * Cfg::kLoadThreads;
  constexpr int XP = Cfg::kXPitch;
  constexpr int CHUNK_BYTES = BG * 128;
  constexpr int UNITS = BG * 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = p.H / 64;
  const size_t SH_SUB = (size_t)2 * KC * CHUNK_BYTES;      // bytes of one group's double-buffered operand
  uint8_t* sH = smem;                                       // [NSUB][2][KC] chunks
  float* sX = reinterpret_cast<float*>(sH + NSUB * SH_SUB); // [NSUB][3][32][XP]
  uint64_t* bar_h = reinterpret_cast<uint64_t*>(sX + NSUB * 3 * 32 * XP);   // [NSUB][2][16]
  uint64_t* bar_d = bar_h + NSUB * 32;                      // [NSUB]
  uint64_t* bar_s = bar_d + NSUB;                           // [NSUB]
  uint64_t* bar_t = bar_s + NSUB;                           // [NSUB] TMA staging of a group's operand
  uint32_t* cnt_p = reinterpret_cast<uint32_t*>(bar_t + NSUB);   // [NSUB] epilogue warps that have stored a step's outputs (monotonic)
  uint32_t* tmem_slot = cnt_p + 2 * NSUB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_layer = p.n_slices * p.n_cgroups;
  const int layer = blockIdx.x / per_layer, rr = blockIdx.x % per_layer;
  const int slice = rr % p.n_slices, cgrp = rr / p.n_slices;
  const StackFwdLayer& L = p.lay[layer];
  const int j0 = slice * REC_US;
  const int a_cols = p.H / 2;
  const bool is_signaller = warp == NTHREADS / 32 - 1;
  const bool is_loader = warp == 0 || (warp >= 2 + Cfg::kEpiWarps && !is_signaller);
  // loader set 0 = warp 0 + the kLoadWarps - 1 warps behind the epilogue warps; set 1 (LSETS == 2) = the kLoadWarps warps after those
  const int lset = (LSETS == 2 && warp >= 2 + Cfg::kEpiWarps + Cfg::kLoadWarps - 1) ? 1 : 0;
  const bool tracing = p.trace != nullptr && (int)blockIdx.x == p.trace_cta;
#define STK_TRACE(step, slot) do { if (tracing) p.trace[(step) * 8 + (slot)] = clock64(); } while (0)

  if (threadIdx.x == 0 && p.trace != nullptr && slice == 0 && cgrp == 0) p.trace[(size_t)2 * p.T * 8 + layer * 8] = stk_globaltimer();
  if (threadIdx.x == 0) {
    for (int c = 0; c < NSUB * 32; ++c) mbar_init(&bar_h[c], Cfg::kLoadWarps);
    for (int s = 0; s < NSUB; ++s) { mbar_init(&bar_d[s], 1); mbar_init(&bar_s[s], Cfg::kEpiWarps); mbar_init(&bar_t[s], 1); cnt_p[s] = 0u; }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<REC_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base + a_cols;               // accumulators: NSUB x BG columns behind the A operand

  // ---- one-time: W_hh slice of this layer -> TMEM (lane 32q+l = gate q, unit j0+l; quarter 3 zero)
  if (warp >= 2 && warp < 6) {
    const int q = warp & 3;
    const uint4* src = reinterpret_cast<const uint4*>(L.whh + ((size_t)(q < 3 ? q : 0) * p.H + j0 + lane) * p.H);
    for (int w0 = 0; w0 < a_cols; w0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (q < 3) u = __ldg(src + w0 / 4 + i);
        v[4 * i] = u.x; v[4 * i + 1] = u.y; v[4 * i + 2] = u.z; v[4 * i + 3] = u.w;
      }
      tmem_st16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + w0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (is_loader) {
    // ---------------- loaders: for every (t, group) pair in order, poll h_{t-1} of the group out of hseq and stage it
    const int lw = warp == 0 ? 0 : (lset == 1 ? warp - (2 + Cfg::kEpiWarps + Cfg::kLoadWarps - 1) : warp - (1 + Cfg::kEpiWarps));
    const int lt = lw * 32 + lane;
    constexpr int UPT = Cfg::kUnitsPerThread;
    constexpr int ROWS_PER_PASS = Cfg::kLoadThreads / 8;
    constexpr int PG = UPT == 1 ? B2T_FWD_POLL_GROUP : 6;
    const bool active = lt < UNITS / UPT;
    const int row = lt >> 3, seg = lt & 7;
    const uint32_t soff = row * 128 + ((seg ^ (row & 7)) << 4);
    const size_t g_unit = (size_t)ROWS_PER_PASS * p.H * sizeof(__nv_bfloat16);
    for (int t = 0; t < p.T; ++t) {
      const int buf = t & 1;
#pragma unroll
      for (int sub = 0; sub < NSUB; ++sub) {
        if (LSETS == 2 && sub != lset) continue;              // this set serves one batch group
        const int b0 = (cgrp * NSUB + sub) * BG;
        if (t > 0) {
          mbar_wait(&bar_s[sub], (uint32_t)(t - 1) & 1u);     // this CTA has stored its own slice of h_{t-1}: the peers do so about now
          if (p.poll_delay > 0) {
            const long long t0 = clock64();
            while (clock64() - t0 < p.poll_delay) {}
          }
        }
        if (t > 0) {   // probe (producer slice, producer epilogue warp): its lanes 0..1 own units 0..7 of trial 4 * warp
          const __nv_bfloat16* hrow = L.hseq + ((size_t)t * p.Bpad + b0) * p.H;
          probe_until_published<Cfg::kLoadThreads>(p.n_slices * Cfg::kEpiWarps, lt, [&](int i) {
            return reinterpret_cast<const uint8_t*>(hrow + (size_t)(4 * (i % Cfg::kEpiWarps)) * p.H + (i / Cfg::kEpiWarps) * REC_US);
          }, 2 + lset);
        }
        const uint8_t* g = reinterpret_cast<const uint8_t*>(L.hseq + ((size_t)t * p.Bpad + b0 + row) * p.H) + seg * 16;   // slot t = h_{t-1}
        uint8_t* sdst = sH + sub * SH_SUB + (size_t)buf * KC * CHUNK_BYTES + soff;
        uint64_t* bars_sub = &bar_h[(sub * 2 + buf) * 16];
        if (p.use_tma && UPT == 1) {
          // The probe has seen every producer warp's store of this step; fetch the operand with KC tensor loads (they land in the
          // swizzled layout the MMA reads) and verify it in shared memory: a word that is still the sentinel -- a store of the
          // same instruction that became visible later than the probed one -- sends this warp through the polled path below.
          if (lt == 0) {
            asm volatile("fence.proxy.async.global;" ::: "memory");   // the producers' generic-proxy stores -> async-proxy reads
            mbar_arrive_expect_tx(&bar_t[sub], (uint32_t)(KC * CHUNK_BYTES));
            uint8_t* dst0 = sH + sub * SH_SUB + (size_t)buf * KC * CHUNK_BYTES;
            for (int c = 0; c < KC; ++c) tma_load_2d(dst0 + (size_t)c * CHUNK_BYTES, &p.tm_h[layer], &bar_t[sub], c * 64, t * p.Bpad + b0);
          }
          mbar_wait(&bar_t[sub], (uint32_t)t & 1u);
          bool ok = true;
          if (active)
            for (int c = 0; c < KC; ++c) ok = ok && !has_sentinel(*reinterpret_cast<const uint4*>(sdst + (size_t)c * CHUNK_BYTES));
          if (__all_sync(0xffffffffu, ok)) {
            for (int c = 0; c < KC; ++c)
              if (lane == 0) mbar_arrive(&bars_sub[c]);
            if (lt == 0 && sub == 0) { STK_TRACE(t, 0); STK_TRACE(t, 1); }
            continue;
          }
        }
#pragma unroll
        for (int gi = 0; gi < (16 + PG - 1) / PG; ++gi) {    // chunk groups unrolled at compile time (a run-time loop around the poll registers spills)
          const int c0 = gi * PG;
          if (c0 < KC) {
            uint64_t* bars = bars_sub + c0;
            const uint8_t* gg = g + c0 * 128;
            poll_and_stage<PG, UPT>(KC - c0 < PG ? KC - c0 : PG, active, [&](int c) { return gg + c * 128; }, g_unit,
                                    sdst + (size_t)c0 * CHUNK_BYTES, CHUNK_BYTES, ROWS_PER_PASS * 128, [&](int c) {
                                      if (lane == 0) mbar_arrive(&bars[c]);
                                      if (lt == 0 && sub == 0 && c0 + c == 0) STK_TRACE(t, 0);
                                    });
          }
        }
        if (lt == 0 && sub == 0) STK_TRACE(t, 1);
      }
    }
  } else if (is_signaller) {
    if (lane == 0 && L.prog) {
      // ---------------- progress signaller: once every epilogue warp has stored a step's outputs (cnt_p), make them visible at
      // gpu scope and bump the layer's progress counter.  The membar costs ~1 us when stores are still in flight; here it
      // delays nobody (in the epilogue warps it sat in front of every step of both groups).
      for (int t = 0; t < p.T; ++t) {
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          uint32_t spins = 0;
          while (smem_acquire_ld(&cnt_p[sub]) < (uint32_t)(t + 1) * Cfg::kEpiWarps) {
            __nanosleep(40);
            if (++spins > (1u << 27)) __trap();
          }
          __threadfence();
          red_release_add(L.prog + t, 1);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BG, 0, 0);
      for (int t = 0; t < p.T; ++t) {
        const int buf = t & 1;
        const uint32_t par = (uint32_t)(t >> 1) & 1u;
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          const uint32_t sbase = smem_u32(sH + sub * SH_SUB + (size_t)buf * KC * CHUNK_BYTES);
          for (int c = 0; c < KC; ++c) {
            mbar_wait(&bar_h[(sub * 2 + buf) * 16 + c], par);
            if (c == 0 && sub == 0) STK_TRACE(t, 2);
            tc_fence_after();
            const uint32_t sb = sbase + c * CHUNK_BYTES;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_ts(tmem_d + sub * BG, tmem_base + (c * 4 + k) * 8, umma_smem_desc(sb + k * 32, 16, 1024), idesc, (c | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_d[sub]);
          if (sub == 0) STK_TRACE(t, 3);
        }
      }
    }
  } else {
    // ---------------- epilogue: thread e owns trial (group base) + e/8 and units j0 + 4*(e%8) .. +3, for both groups in turn
    const int e = threadIdx.x - 64;
    const int ew = e >> 5;
    const int q = warp & 3;
    const int chalf = ew >> 2;
    const int bl = e >> 3, u0 = (e & 7) * 4;
    const int j = j0 + u0;
    float h[NSUB][4], bh[3][4];
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
      const int b = (cgrp * NSUB + sub) * BG + bl;
      const float4 hv = *reinterpret_cast<const float4*>(L.h_state + (size_t)b * p.H + j);
      h[sub][0] = hv.x; h[sub][1] = hv.y; h[sub][2] = hv.z; h[sub][3] = hv.w;
    }
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 bv = *reinterpret_cast<const float4*>(L.bhh + g * p.H + j);
      bh[g][0] = bv.x; bh[g][1] = bv.y; bh[g][2] = bv.z; bh[g][3] = bv.w;
    }
    const bool train = L.R != nullptr;
    const float inv_keep = 1.0f / L.keep;
    int ready_tile = -1;                                     // gx tiles up to here are known complete
    for (int t = 0; t < p.T; ++t) {
#pragma unroll
      for (int sub = 0; sub < NSUB; ++sub) {
        const int b0 = (cgrp * NSUB + sub) * BG;
        const int b = b0 + bl;
        const size_t row = (size_t)t * p.Bpad + b;
        if (L.gx_done) {                                     // the input projection of this step's rows must have landed
          const int tile = (int)(((size_t)t * p.Bpad + b0) / 128);   // BG divides 128 and Bpad: the group's rows lie in one tile
          if (tile > ready_tile) { warp_wait_ge(L.gx_done + tile, L.gx_need, lane); ready_tile = tile; }
        }
        float4 gxv[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) gxv[g] = ldcg_f4(L.gx + row * 3 * p.H + g * p.H + j);

        mbar_wait(&bar_d[sub], (uint32_t)t & 1u);
        if (e == 0 && sub == 0) STK_TRACE(t, 4);
        tc_fence_after();
        float* sXs = sX + sub * 3 * 32 * XP;
        if (q < 3) {
          uint32_t v[16];
          tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + sub * BG + chalf * 16, v);
          tmem_ld_wait();
          float4* dst = reinterpret_cast<float4*>(sXs + (q * 32 + lane) * XP + chalf * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
        tc_fence_before();
        epi_bar_sync<Cfg::kEpiThreads>();                    // gates exchanged (sX of a group is rewritten only after the group's next MMA, which needs every thread's h_t)
        float hn[4], r[4], z[4], n[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ar = sXs[(0 * 32 + u0 + i) * XP + bl];
          const float az = sXs[(1 * 32 + u0 + i) * XP + bl];
          const float an = sXs[(2 * 32 + u0 + i) * XP + bl];
          const float gr = (&gxv[0].x)[i], gz = (&gxv[1].x)[i], gn = (&gxv[2].x)[i];
          r[i] = sigmoid_f(gr + ar + bh[0][i]);
          z[i] = sigmoid_f(gz + az + bh[1][i]);
          hn[i] = an + bh[2][i];
          n[i] = tanh_f(gn + r[i] * hn[i]);
          h[sub][i] = (1.0f - z[i]) * n[i] + z[i] * h[sub][i];
        }
        const size_t off = row * p.H + j;
        {
          __nv_bfloat162 lo = __floats2bfloat162_rn(h[sub][0], h[sub][1]), hi = __floats2bfloat162_rn(h[sub][2], h[sub][3]);
          st_relaxed_v2(L.hseq + ((size_t)(t + 1) * p.Bpad + b) * p.H + j, *reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_s[sub]);
        if (e == 0 && sub == 0) STK_TRACE(t, 5);
        if (e == 0 && sub == NSUB - 1) STK_LAYER_STAMP(0, slice == 0 && cgrp == 0, t);
        if (train) {
          st_bf16x4(L.R + off, r[0], r[1], r[2], r[3]);
          st_bf16x4(L.Z + off, z[0], z[1], z[2], z[3]);
          st_bf16x4(L.Nn + off, n[0], n[1], n[2], n[3]);
          st_bf16x4(L.HN + off, hn[0], hn[1], hn[2], hn[3]);
        }
        if (L.hdrop) {
          float d[4] = {h[sub][0], h[sub][1], h[sub][2], h[sub][3]};
          if (L.keep < 1.0f) {
            const uint4 rnd = rec_dropout_bits(p.seed, L.rng_offset, off >> 2);
            const uint32_t rr4[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
            uint32_t kb = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const bool kept = u32_to_unit(rr4[i]) < L.keep;
              d[i] = kept ? d[i] * inv_keep : 0.0f;
              kb |= (kept ? 1u : 0u) << i;
            }
            if (L.kmask) L.kmask[off >> 2] = (unsigned char)kb;   // backward re-reads the mask instead of re-drawing it (the Philox rounds sat on its critical path)
          }
          st_bf16x4(L.hdrop + off, d[0], d[1], d[2], d[3]);
        }
        if (L.prog) {                                        // outputs of (t, group) stored by this warp: hand over to the signaller thread
          __syncwarp();
          if (lane == 0) smem_release_inc(&cnt_p[sub]);
        }
      }
    }
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
      const int b = (cgrp * NSUB + sub) * BG + bl;
      *reinterpret_cast<float4*>(L.h_state + (size_t)b * p.H + j) = make_float4(h[sub][0], h[sub][1], h[sub][2], h[sub][3]);
    }
  }
#undef STK_TRACE

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<REC_TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ backward
// Two-dimensional decomposition of gru_rec_bwd2.cuh (H multiple of 256), all layers in one launch, NSUB batch groups per CTA.
struct StackBwdLayer {
  const float* dY;                // [T][Bpad][H] fp32 gradient wrt this layer's (dropped) output (top layer: complete before launch)
  const __nv_bfloat16* hseq;
  const __nv_bfloat16 *R, *Z, *Nn, *HN;
  const __nv_bfloat16* whh;
  __nv_bfloat16* dGx;             // [T][Bpad][3H]
  __nv_bfloat16* dGh;             // [T][Bpad][3H], sentinel-filled: doubles as the exchange medium of the dG all-gather
  float* part;                    // [2][groups][H/128][4][4][BG*32] tagged fp32 partial sums
  float *dbih, *dbhh;             // [3H] (atomicAdd)
  float* dh_state;                // [Bpad][H]: gradient wrt the initial state on exit
  int dy_polled;                  // dY is produced beside this kernel by the gated data-gradient GEMM of the layer above: sentinel-filled, polled
  int* prog;                      // [T]: += 1 per epilogue warp, CTA and batch group once dGx[t] is stored (null: no consumer)
  int gen_base;                   // generation of this launch's first step (tag = (gen >> 1) & 3, buffer = gen & 1)
  float keep;                     // dropout that forward applied to this layer's output (1 => none)
  unsigned long long rng_offset;
  const unsigned char* kmask;     // keep bits written by the forward stack kernel (null: regenerate them from the Philox stream)
};

struct StackBwdParams {
  int H, Bpad, T, n_layers, n_cgroups, n_valid;
  int poll_delay;
  int trace_cta;
  unsigned long long seed;
  long long* trace;
  StackBwdLayer lay[STACK_MAX_LAYERS];
};

// ESETS = 2 (with NSUB = 2): each batch group gets its own set of epilogue warps.  With one set the four phases of a step --
// A(s,0) B(s-1,1) B(s,0) A(s,1) -- queue up behind each other on the same warps and the step is bound by their sum (measured:
// "accumulator done -> partials reduced" 9 000 cycles of which most is waiting for the warps, not for the peers).
// LSPLIT = 1 (with NSUB = 2): the 8 loader warps work as two sets of 4, one per batch group (two 16-byte units per thread and chunk),
// so that a group's dG is staged the moment it is published instead of behind the other group's staging.
template <int BG, int NSUB, int ESETS = 1, int LSPLIT = 0>
__global__ void __launch_bounds__(RecCfg<BG>::kFwdThreads + 32 + (ESETS - 1) * RecCfg<BG>::kEpiThreads, 1)
gru_stack_bwd_kernel(const __grid_constant__ StackBwdParams p) {
  static_assert(ESETS == 1 || (ESETS == 2 && NSUB == 2), "one epilogue set, or one per batch group");
  static_assert(LSPLIT == 0 || (NSUB == 2 && BG == 32), "split loader sets are written for two groups of 32 trials");
  constexpr int LOAD_WARPS = LSPLIT ? RecCfg<BG>::kLoadWarps / 2 : RecCfg<BG>::kLoadWarps;   // per set
  constexpr int LOAD_THREADS = 32 * LOAD_WARPS;
  using Cfg = RecCfg<BG>;
  constexpr int CHUNK_BYTES = BG * 128;
  constexpr int A_CHUNK = 128 * 128;
  constexpr int UNITS = BG * 8;
  constexpr int NTHREADS = Cfg::kFwdThreads + 32 + (ESETS - 1) * Cfg::kEpiThreads;
  constexpr int EPI_WARPS_ALL = ESETS * Cfg::kEpiWarps;     // warps 2 .. 2 + EPI_WARPS_ALL - 1
  constexpr int NSL = ESETS == 2 ? 1 : NSUB;                // batch groups whose state one epilogue thread carries
  constexpr uint32_t TMEM_COLS = NSUB * BG < 32 ? 32 : NSUB * BG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KQ = p.H / 4;
  const int CPG = KQ / 64;
  const int KC = 3 * CPG;                                  // <= 9
  const int MB = p.H / 128;
  uint8_t* sA = smem;                                      // [KC][128][128 B]
  uint8_t* sB = sA + (size_t)KC * A_CHUNK;                 // [NSUB][KC][BG][128 B]
  const size_t SB_SUB = (size_t)KC * CHUNK_BYTES;
  uint64_t* bar_h = reinterpret_cast<uint64_t*>(sB + NSUB * SB_SUB);   // [NSUB][16]
  uint64_t* bar_d = bar_h + NSUB * 16;                     // [NSUB]
  uint64_t* bar_s = bar_d + NSUB;                          // [NSUB]
  uint32_t* cnt_p = reinterpret_cast<uint32_t*>(bar_s + NSUB);   // [NSUB] epilogue warps that have stored dGx of a step (monotonic)
  uint32_t* tmem_slot = cnt_p + 2 * NSUB;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_group = MB * 4;
  const int per_layer = per_group * p.n_cgroups;
  const int layer = blockIdx.x / per_layer, rl = blockIdx.x % per_layer;
  const int cgrp = rl / per_group, rr = rl % per_group;
  const int mb = rr >> 2, kq = rr & 3;
  const int NG = p.n_cgroups * NSUB;                       // batch groups of the layer
  const StackBwdLayer& L = p.lay[layer];
  const int j0 = mb * 128 + kq * REC_US;
  const bool is_signaller = warp == NTHREADS / 32 - 1;
  const bool is_loader = warp == 0 || (warp >= 2 + EPI_WARPS_ALL && !is_signaller);
  const bool tracing = p.trace != nullptr && (int)blockIdx.x == p.trace_cta;
#define STK_TRACE(step, slot) do { if (tracing) p.trace[((size_t)p.T + (step)) * 8 + (slot)] = clock64(); } while (0)

  if (threadIdx.x == 0 && p.trace != nullptr && rl == 0) p.trace[(size_t)2 * p.T * 8 + 64 + layer * 8] = stk_globaltimer();
  if (threadIdx.x == 0) {
    for (int c = 0; c < NSUB * 16; ++c) mbar_init(&bar_h[c], LOAD_WARPS);
    for (int s = 0; s < NSUB; ++s) { mbar_init(&bar_d[s], 1); mbar_init(&bar_s[s], Cfg::kEpiWarps); cnt_p[s] = 0u; }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);

  // ---- one-time: A operand = W_hh^T block (rows: output units mb*128.., columns: gate g of units kq*KQ..), K-major, swizzled
  {
    const int KK = 3 * KQ, NP = KK / 2;
    for (int it = threadIdx.x; it < 16 * NP; it += NTHREADS) {
      const int pr = it % NP, k8 = it / NP;
      const int kk = 2 * pr, g = kk / KQ, uu = kk - g * KQ;
      const __nv_bfloat16* src = L.whh + ((size_t)g * p.H + kq * KQ + uu) * p.H + mb * 128 + k8 * 8;
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(src));
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(src + p.H));
      const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w}, hh[4] = {hi.x, hi.y, hi.z, hi.w};
      uint8_t* base = sA + (size_t)(kk >> 6) * A_CHUNK + (kk & 7) * 2;
      const int ku = (kk & 63) >> 3;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t a = (l[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu, b = (hh[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
        const int kr = k8 * 8 + i;
        *reinterpret_cast<uint32_t*>(base + kr * 128 + ((ku ^ (kr & 7)) << 4)) = a | (b << 16);
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  // Steps are indexed s = 0..T-1 for t = T-1-s.
  if (is_loader) {
    const int lw_all = warp == 0 ? 0 : warp - (1 + EPI_WARPS_ALL);          // 0 .. kLoadWarps - 1
    const int lset = LSPLIT ? lw_all / LOAD_WARPS : 0;
    const int lw = LSPLIT ? lw_all % LOAD_WARPS : lw_all;
    const int lt = lw * 32 + lane;
    constexpr int UPT = (UNITS + LOAD_THREADS - 1) / LOAD_THREADS;
    constexpr int PC = UPT == 1 ? 9 : 5;
    constexpr int ROWS_PER_PASS = LOAD_THREADS / 8;
    const bool active = lt < UNITS / UPT;
    const int row = lt >> 3, seg = lt & 7;
    const uint32_t soff = row * 128 + ((seg ^ (row & 7)) << 4);
    const size_t g_unit = (size_t)ROWS_PER_PASS * 3 * p.H * sizeof(__nv_bfloat16);
    for (int s = 0; s < p.T; ++s) {
      const int t = p.T - 1 - s;
#pragma unroll
      for (int sub = 0; sub < NSUB; ++sub) {
        if (LSPLIT && sub != lset) continue;                   // this set serves one batch group
        const int b0 = (cgrp * NSUB + sub) * BG;
        // start when this CTA's own epilogue has published dG_t of the group (also orders the staging after MMA(s-1, group),
        // which read the same single buffer: dG of step s is published only after the accumulator of step s-1 was drained)
        mbar_wait(&bar_s[sub], (uint32_t)s & 1u);
        if (p.poll_delay > 0) {
          const long long t0 = clock64();
          while (clock64() - t0 < p.poll_delay) {}
        }
        {   // probe (producer CTA of this contraction quarter, its epilogue warp, gate)
          const __nv_bfloat16* grow = L.dGh + ((size_t)t * p.Bpad + b0) * 3 * p.H + kq * KQ;
          const int npc = KQ / REC_US;
          probe_until_published<LOAD_THREADS>(npc * Cfg::kEpiWarps * 3, lt, [&](int i) {
            const int gate = i % 3, w = (i / 3) % Cfg::kEpiWarps, pc = i / (3 * Cfg::kEpiWarps);
            return reinterpret_cast<const uint8_t*>(grow + (size_t)(4 * w) * 3 * p.H + (size_t)gate * p.H + pc * REC_US);
          }, 2 + lset);
        }
        const uint8_t* g = reinterpret_cast<const uint8_t*>(L.dGh + ((size_t)t * p.Bpad + b0 + row) * 3 * p.H + kq * KQ) + seg * 16;
        auto chunk_addr = [&](int cc) {
          const int gate = cc / CPG, ci = cc - gate * CPG;
          return g + ((size_t)gate * p.H + ci * 64) * sizeof(__nv_bfloat16);
        };
        uint8_t* sdst = sB + sub * SB_SUB + soff;
        uint64_t* bars = &bar_h[sub * 16];
        if constexpr (PC >= 9) {
          poll_and_stage<PC, UPT>(KC, active, chunk_addr, g_unit, sdst, CHUNK_BYTES, ROWS_PER_PASS * 128, [&](int c) {
            if (lane == 0) mbar_arrive(&bars[c]);
            if (lt == 0 && sub == 0 && c == 0) STK_TRACE(s, 0);
          });
        } else {
          for (int c0 = 0; c0 < KC; c0 += PC) {
            poll_and_stage<PC, UPT>(KC - c0 < PC ? KC - c0 : PC, active, [&](int c) { return chunk_addr(c0 + c); }, g_unit,
                                    sdst + (size_t)c0 * CHUNK_BYTES, CHUNK_BYTES, ROWS_PER_PASS * 128, [&](int c) {
                                      if (lane == 0) mbar_arrive(&bars[c0 + c]);
                                    });
          }
        }
        if (lt == 0 && sub == 0) STK_TRACE(s, 1);
      }
    }
  } else if (is_signaller) {
    if (lane == 0 && L.prog) {
      // ---------------- progress signaller: dGx of (step, group) stored by every epilogue warp -> visible -> counter (see forward)
      for (int s = 0; s < p.T; ++s) {
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          uint32_t spins = 0;
          while (smem_acquire_ld(&cnt_p[sub]) < (uint32_t)(s + 1) * Cfg::kEpiWarps) {
            __nanosleep(40);
            if (++spins > (1u << 27)) __trap();
          }
          __threadfence();
          red_release_add(L.prog + (p.T - 1 - s), 1);
          if (p.trace) p.trace[(size_t)2 * p.T * 8 + 128 + (size_t)blockIdx.x * 8 + 5] = 0xC000 | (s << 4) | sub;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BG, 0, 0);
      const uint32_t sa = smem_u32(sA);
      for (int s = 0; s < p.T; ++s) {
        const uint32_t par = (uint32_t)s & 1u;
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          const uint32_t sb = smem_u32(sB + sub * SB_SUB);
          mbar_wait(&bar_s[sub], par);                     // accumulator of step s-1 drained by every epilogue warp
          for (int c = 0; c < KC; ++c) {
            mbar_wait(&bar_h[sub * 16 + c], par);
            if (c == 0 && sub == 0) STK_TRACE(s, 2);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_d + sub * BG, umma_smem_desc(sa + c * A_CHUNK + k * 32, 16, 1024), umma_smem_desc(sb + c * CHUNK_BYTES + k * 32, 16, 1024),
                        idesc, (c | k) != 0 ? 1u : 0u);
          }
          umma_commit(&bar_d[sub]);
          if (sub == 0) STK_TRACE(s, 3);
        }
      }
    }
  } else {
    // ---------------- epilogue.  Per (step, group): phase A = reduce partials of the previous step, gate math, publish dG_t;
    // phase B = drain the accumulator of the group's MMA into the four partial blocks.  The phases of the two groups are
    // interleaved A(s,0) B(s-1,1) B(s,0) A(s,1).
    const int eset = (warp - 2) / Cfg::kEpiWarps;          // 0 when there is one set
    const int e = threadIdx.x - 64 - eset * Cfg::kEpiThreads;
    const int ew = e >> 5;
    const int q = warp & 3;
    const int chalf = ew >> 2;
    const int bl = e >> 3, u0 = (e & 7) * 4;
    const int j = j0 + u0;
    const float inv_keep = 1.0f / L.keep;
    auto SI = [](int sub) { return ESETS == 2 ? 0 : sub; };   // slot of a group's state in this thread's arrays
    float carry[NSL][4];
    float accx[3][4], acch[4];                             // bias-gradient sums: both groups add into the same units
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      accx[0][i] = accx[1][i] = accx[2][i] = 0.f; acch[i] = 0.f;
#pragma unroll
      for (int sub = 0; sub < NSL; ++sub) carry[sub][i] = 0.f;
    }

    auto block_of = [&](int grp, int gen, int dest, int src) {
      return L.part + (((((size_t)(gen & 1) * NG + grp) * MB + mb) * 4 + dest) * 4 + src) * (BG * 32);
    };
    auto reduce_partials = [&](int grp, int gen, float (&P)[4]) {
      const uint32_t tag = (uint32_t)(gen >> 1) & 3u;
      const float* base = block_of(grp, gen, kq, 0) + (ew * 32 + lane) * 4;
      uint4 v[4];
      uint32_t pending = 0xFu, spins = 0;
      while (true) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if ((pending >> i) & 1u) v[i] = ld_relaxed_v4(base + (size_t)i * (BG * 32));
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (((pending >> i) & 1u) && tags_match(v[i], tag)) pending &= ~(1u << i);
        if (!pending) break;
        if (++spins > (1u << 22)) {
          if (p.trace) { long long* d = p.trace + (size_t)2 * p.T * 8 + 128 + (size_t)blockIdx.x * 8; d[0] = 0xD2; d[1] = layer; d[2] = grp; d[3] = gen; d[4] = pending; __threadfence_system(); }
          __trap();
        }
      }
      float G[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        G[0] += __uint_as_float(v[i].x & ~3u); G[1] += __uint_as_float(v[i].y & ~3u);
        G[2] += __uint_as_float(v[i].z & ~3u); G[3] += __uint_as_float(v[i].w & ~3u);
      }
      const int k = lane >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sl = 4 * (lane & 7) + i;
        const float a0 = __shfl_sync(0xffffffffu, G[0], sl), a1 = __shfl_sync(0xffffffffu, G[1], sl);
        const float a2 = __shfl_sync(0xffffffffu, G[2], sl), a3 = __shfl_sync(0xffffffffu, G[3], sl);
        P[i] = k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
      }
    };
    struct Stash { uint2 r, z, n, hn, hp; float4 dy; };
    auto unpack = [](const uint2& u, float (&f)[4]) {
      const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
      f[0] = __low2float(lo); f[1] = __high2float(lo); f[2] = __low2float(hi); f[3] = __high2float(hi);
    };
    auto pack2 = [](float a, float c) {
      __nv_bfloat162 v = __floats2bfloat162_rn(a, c);
      return *reinterpret_cast<uint32_t*>(&v);
    };
    // forward-pass stash of step t (always there) ...
    auto load_fwd_stash = [&](int b, int t, Stash& st) {
      const size_t off = ((size_t)t * p.Bpad + b) * p.H + j;
      st.r = __ldg(reinterpret_cast<const uint2*>(L.R + off)); st.z = __ldg(reinterpret_cast<const uint2*>(L.Z + off));
      st.n = __ldg(reinterpret_cast<const uint2*>(L.Nn + off)); st.hn = __ldg(reinterpret_cast<const uint2*>(L.HN + off));
      st.hp = *reinterpret_cast<const uint2*>(L.hseq + off);             // slot t = h_{t-1}
    };
    // ... and the gradient from the layer above, which a gated GEMM may still be producing (see phase A)
    auto load_dy = [&](int b, int t, Stash& st) {
      st.dy = ldcg_f4(L.dY + ((size_t)t * p.Bpad + b) * p.H + j);
    };
    // inter-layer dropout mask of (row, units j..j+3), regenerated from the forward pass's Philox stream: bit i = unit i kept.
    // Drawn one step ahead (behind the publish of the current step) so that the Philox rounds are off the step's critical path.
    auto keep_bits = [&](int b, int t) -> uint32_t {
      if (!(L.keep < 1.0f)) return 0xFu;
      const size_t off = ((size_t)t * p.Bpad + b) * p.H + j;
      if (L.kmask) return (uint32_t)__ldg(L.kmask + (off >> 2));
      const uint4 rnd = rec_dropout_bits(p.seed, L.rng_offset, off >> 2);
      return (u32_to_unit(rnd.x) < L.keep ? 1u : 0u) | (u32_to_unit(rnd.y) < L.keep ? 2u : 0u) | (u32_to_unit(rnd.z) < L.keep ? 4u : 0u) |
             (u32_to_unit(rnd.w) < L.keep ? 8u : 0u);
    };
    Stash cur[NSL];
    uint32_t kbits[NSL];
#pragma unroll
    for (int si = 0; si < NSL; ++si) {
      const int b = (cgrp * NSUB + (ESETS == 2 ? eset : si)) * BG + bl;
      load_fwd_stash(b, p.T - 1, cur[si]);
      load_dy(b, p.T - 1, cur[si]);
      kbits[si] = keep_bits(b, p.T - 1);
    }

    auto phase_a = [&](int s, int sub) {
      if (p.trace && e == 0) p.trace[(size_t)2 * p.T * 8 + 128 + (size_t)blockIdx.x * 8 + 6] = 0xA000 | (s << 4) | sub;
      const int t = p.T - 1 - s;
      const int grp = cgrp * NSUB + sub;
      const int b0 = grp * BG, b = b0 + bl;
      const bool valid = b < p.n_valid;
      const size_t row = (size_t)t * p.Bpad + b;
      const size_t off = row * p.H + j;
      // dY of this step was requested one step ago.  It comes from the layer above through a GEMM that runs beside this kernel:
      // the buffer is sentinel-filled before the step and the data is its own signal (one L2 round trip instead of a flag
      // acquire followed by the load; a 16-byte store lands whole)
      if (L.dy_polled) {
        uint32_t spins = 0;
        while (__float_as_uint(cur[SI(sub)].dy.x) == 0xFFFFFFFFu || __float_as_uint(cur[SI(sub)].dy.w) == 0xFFFFFFFFu) {
          const uint4 v = ld_strong_v4(L.dY + ((size_t)t * p.Bpad + b) * p.H + j);
          cur[SI(sub)].dy = make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
          if (++spins > (1u << 22)) {
            if (p.trace) { long long* d = p.trace + (size_t)2 * p.T * 8 + 128 + (size_t)blockIdx.x * 8; d[0] = 0xD1; d[1] = layer; d[2] = t; d[3] = b; d[4] = j; d[5] = s; __threadfence_system(); }
            __trap();
          }
        }
      }
      float dmask[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) dmask[i] = (L.keep < 1.0f) ? (((kbits[SI(sub)] >> i) & 1u) ? inv_keep : 0.0f) : 1.0f;
      float P[4] = {0.f, 0.f, 0.f, 0.f};
      if (s > 0) {
        if (e == 0 && sub == 0) STK_TRACE(s, 4);
        reduce_partials(grp, L.gen_base + s - 1, P);
        if (e == 0 && sub == 0) STK_TRACE(s, 5);
      }
      float r[4], z[4], n[4], hn[4], hp[4];
      unpack(cur[SI(sub)].r, r); unpack(cur[SI(sub)].z, z); unpack(cur[SI(sub)].n, n); unpack(cur[SI(sub)].hn, hn); unpack(cur[SI(sub)].hp, hp);
      const float dy[4] = {cur[SI(sub)].dy.x * dmask[0], cur[SI(sub)].dy.y * dmask[1], cur[SI(sub)].dy.z * dmask[2], cur[SI(sub)].dy.w * dmask[3]};
      float gr[4], gz[4], gn[4], gnh[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float dh = s > 0 ? carry[SI(sub)][i] + P[i] + dy[i] : dy[i];
        const float d = valid ? dh : 0.0f;
        const float dn = d * (1.0f - z[i]);
        const float dz = d * (hp[i] - n[i]);
        gn[i] = dn * (1.0f - n[i] * n[i]);
        gz[i] = dz * z[i] * (1.0f - z[i]);
        gr[i] = gn[i] * hn[i] * r[i] * (1.0f - r[i]);
        gnh[i] = gn[i] * r[i];
        carry[SI(sub)][i] = d * z[i];
        accx[0][i] += gr[i]; accx[1][i] += gz[i]; accx[2][i] += gn[i]; acch[i] += gnh[i];
      }
      // publish dGh_t: the data is the signal (polled by the loaders of the group's CTAs)
      const size_t goff = row * 3 * p.H + j;
      st_relaxed_v2(L.dGh + goff, pack2(gr[0], gr[1]), pack2(gr[2], gr[3]));
      st_relaxed_v2(L.dGh + goff + p.H, pack2(gz[0], gz[1]), pack2(gz[2], gz[3]));
      st_relaxed_v2(L.dGh + goff + 2 * p.H, pack2(gnh[0], gnh[1]), pack2(gnh[2], gnh[3]));
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_s[sub]);
      if (e == 0 && sub == 0) STK_TRACE(s, 6);
      if (e == 0 && sub == NSUB - 1) STK_LAYER_STAMP(1, rl == 0, s);
      st_bf16x4(L.dGx + goff, gr[0], gr[1], gr[2], gr[3]);
      st_bf16x4(L.dGx + goff + p.H, gz[0], gz[1], gz[2], gz[3]);
      st_bf16x4(L.dGx + goff + 2 * p.H, gn[0], gn[1], gn[2], gn[3]);
      if (L.prog) {                                          // dGx of (s, group) stored by this warp: hand over to the signaller thread
        __syncwarp();
        if (lane == 0) smem_release_inc(&cnt_p[sub]);
      }
      // next step's operands: forward stash, dY (possibly still the sentinel: re-polled when it is needed) and the dropout mask
      if (s + 1 < p.T) {
        load_fwd_stash(b, t - 1, cur[SI(sub)]);
        load_dy(b, t - 1, cur[SI(sub)]);
        kbits[SI(sub)] = keep_bits(b, t - 1);
      }
    };
    auto phase_b = [&](int s, int sub) {
      if (p.trace && e == 0) p.trace[(size_t)2 * p.T * 8 + 128 + (size_t)blockIdx.x * 8 + 7] = 0xB000 | (s << 4) | sub;
      const int grp = cgrp * NSUB + sub;
      mbar_wait(&bar_d[sub], (uint32_t)s & 1u);
      if (e == 0 && sub == 0) STK_TRACE(s, 7);
      tc_fence_after();
      const int gen = L.gen_base + s;
      const uint32_t tag = (uint32_t)(gen >> 1) & 3u;
      uint32_t v[16];
      tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + sub * BG + chalf * 16, v);
      tmem_ld_wait();
      float* dst = block_of(grp, gen, q, kq) + ((chalf * 4) * 32 + lane) * 4;
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4)
        st_relaxed_v4(dst + g4 * 128, (v[4 * g4] & ~3u) | tag, (v[4 * g4 + 1] & ~3u) | tag, (v[4 * g4 + 2] & ~3u) | tag, (v[4 * g4 + 3] & ~3u) | tag);
      tc_fence_before();
    };

    for (int s = 0; s < p.T; ++s) {
      if constexpr (NSUB == 1) {
        phase_a(s, 0);
        phase_b(s, 0);
      } else if constexpr (ESETS == 2) {
        phase_a(s, eset);
        phase_b(s, eset);
      } else {
        // In the steady state the two groups run half a period apart: A(s,0) and B(s-1,1) are due together, then B(s,0) and A(s,1).
        // (A phase right behind the B phase it depends on -- B(s-1,1) A(s,1) -- would stall for a full L2 hand-over.)
        phase_a(s, 0);
        if (s > 0) phase_b(s - 1, 1);
        phase_b(s, 0);
        phase_a(s, 1);
      }
    }
    if constexpr (NSUB == 2 && ESETS == 1) phase_b(p.T - 1, 1);
    // recurrent gradient wrt the initial state: dh_{-1} = dh_0 * z_0 + dGh_0 W_hh
#pragma unroll
    for (int si = 0; si < NSL; ++si) {
      const int sub = ESETS == 2 ? eset : si;
      const int grp = cgrp * NSUB + sub;
      const int b = grp * BG + bl;
      const bool valid = b < p.n_valid;
      float P[4];
      reduce_partials(grp, L.gen_base + p.T - 1, P);
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = valid ? carry[si][i] + P[i] : 0.0f;
      *reinterpret_cast<float4*>(L.dh_state + (size_t)b * p.H + j) = make_float4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(L.dbih + j + i, accx[0][i]);
      atomicAdd(L.dbih + p.H + j + i, accx[1][i]);
      atomicAdd(L.dbih + 2 * p.H + j + i, accx[2][i]);
      atomicAdd(L.dbhh + j + i, accx[0][i]);
      atomicAdd(L.dbhh + p.H + j + i, accx[1][i]);
      atomicAdd(L.dbhh + 2 * p.H + j + i, acch[i]);
    }
  }
#undef STK_TRACE

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_d);
  }
}

}  // namespace b2t
