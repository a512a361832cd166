// Persistent GRU recurrence kernels for sm_100a (one launch per layer, direction and time chunk).
//
// Reference semantics: torch.nn.GRU as used at rnn_model.py:65-72,126 (gate order r,z,n):
//   r = sig(gx_r + W_hr h + b_hr)   z = sig(gx_z + W_hz h + b_hz)
//   n = tanh(gx_n + r * (W_hn h + b_hn))      h' = (1-z) * n + z * h
// gx = x W_ih^T + b_ih is produced for all time steps by the tcgen05 GEMM (gemm.cuh).
//
// Decomposition.  The batch is cut into groups of BG trials (16 or 32) and the hidden units into slices of 32.
// CTA (slice s, group g) owns the 96 gate rows {r,z,n} x 32 units of W_hh for the whole chunk.
// The weights live in TENSOR MEMORY as the A operand of tcgen05.mma (lane = gate row, 32-bit column =
// two packed bf16 of K), so a time step never re-reads them from shared memory.
//
//   forward : D[128 (96 used), BG] = W_slice[128, H] (TMEM) * h_{t-1}[BG, H]^T (smem, via TMA)
//             The three gate row-blocks land in TMEM lane quarters 0,1,2; the epilogue warps move them
//             through a small smem exchange so that one thread owns (trial, 4 units) with all gates, does
//             the gate math in fp32 (h itself is carried in fp32 registers) and writes h_t (bf16).
//   backward: the same CTA owns dG_t[BG, 96] (its own 32 units x 3 gates) and keeps W_slice^T in TMEM:
//             P[H, BG] = W_slice^T[H, 96] (TMEM, H/128 row blocks) * dG_t[BG, 96]^T (smem, written
//             locally).  P is this CTA's partial of dh_{t-1} for ALL units; the H/32 CTAs of a batch
//             group exchange fp32 partials through L2 (reduce-scatter) once per step.
//
// Measured on B200 (profiles/r1_mma_dispatch_microbench.md): a tcgen05.mma with N <= 64 costs a fixed ~50 cycles,
// so the step is bound by the NUMBER of MMAs (48 forward, 36 backward) and by the exchange round trip through
// L2, not by N.  BG = 32 therefore costs the same per step as BG = 16 but needs half the CTAs (48 for H = 768,
// B = 64), which lets the engine run the chunks of up to three layers concurrently (wave-front over layers and
// time chunks, see engine.cu).  A chunk [t_begin, t_end) carries its state in fp32 (h / dh) between launches.
//
// Per-step exchange.  Trials are independent, so only the H/32 CTAs of one batch group exchange data.  There is no
// flag and no fence on that path: the DATA is the signal.  Every 32-bit word is written by one relaxed gpu-scope
// store and polled with relaxed gpu-scope loads (each word is single-copy atomic, the consumer only uses words it
// has itself observed as valid, and what it does with them is data-dependent on those loads):
//   forward : hseq slots are pre-filled with the bf16 pair 0xFFFF'FFFF (a NaN pattern cvt.rn never produces) by the
//             engine; loader warps poll the BG x H block of h_{t-1} until no word is the sentinel and stage it in
//             shared memory (manual 128B swizzle) as the B operand.
//   backward: the two low mantissa bits of every fp32 partial carry a generation tag ((step >> 1) & 3; the buffer is
//             double-buffered on step & 1), so a stale word is never mistaken for a fresh one and nothing is reset.
// Cooperative launch guarantees that the CTAs polling each other are co-resident.
#pragma once
#include "sm100.cuh"

namespace b2t {

constexpr int REC_US = 32;        // hidden units per CTA
constexpr int REC_TMEM_COLS = 512;

template <int BG> struct RecCfg {
  static constexpr int kEpiWarps = BG / 4;            // 4 / 8 / 16 (BG = 16 / 32 / 64): one thread per (trial, 4 units)
  static constexpr int kEpiThreads = 32 * kEpiWarps;
  static constexpr int kLoadWarps = 8;                // forward: warp 0 + the last 7 warps poll/stage h_{t-1}
  static constexpr int kLoadThreads = 32 * kLoadWarps;
  static constexpr int kUnitsPerThread = (BG * 8 + kLoadThreads - 1) / kLoadThreads;   // 16 B units per loader thread and 64-column chunk (2 for BG = 64)
  static constexpr int kPollChunks = kUnitsPerThread == 1 ? 16 : 6;                    // chunks polled together (bounds the registers of a polling round)
  static constexpr int kFwdThreads = 64 + kEpiThreads + 32 * (kLoadWarps - 1);   // warp1 = MMA issuer + TMEM owner
  static constexpr int kBwdThreads = kEpiThreads;                               // epilogue threads per batch group (the kernel adds one MMA-issuing warp per group)
  static constexpr int kXPitch = BG + 4;              // exchange row pitch (floats)
  static constexpr size_t fwd_smem_bytes(int H) { return (size_t)2 * (H / 64) * BG * 128 + (size_t)3 * 32 * kXPitch * 4 + 512 + 1024; }
  // backward: W_slice^T as the K-major A operand (2 chunks of 64 kk) + dG_t as the B operand (2 chunks) + barriers
  static constexpr size_t bwd_smem_bytes(int H, int nsub) { return (size_t)2 * ((H + 127) / 128) * 128 * 128 + (size_t)nsub * (2 * BG * 128 + 64) + 64 + 1024; }
};

struct RecFwdParams {
  int H, Bpad;                    // hidden size, padded batch (multiple of BG)
  int t_begin, t_end;             // time steps of this chunk
  int n_slices;                   // H / 32
  const float* gx;                // [T][Bpad][3H] fp32, includes b_ih
  const float* bhh;               // [3H]
  const __nv_bfloat16* whh;       // [3H][H] bf16
  __nv_bfloat16* hseq;            // [(T+1)][Bpad][H]; slot 0 = initial state, slot t+1 = h_t
  float* h_state;                 // [Bpad][H] fp32: h_{t_begin-1} on entry, h_{t_end-1} on exit (chunk carry / final state)
  __nv_bfloat16* hdrop;           // [T][Bpad][H] dropout(h_t) for the next layer (nullable => not written)
  __nv_bfloat16 *R, *Z, *Nn, *HN; // [T][Bpad][H] stash for BPTT (nullable when not training)
  int T;                          // total steps
  int poll_delay;                 // cycles between this CTA's own h_t store and the first polling round (tuning knob)
  float keep;                     // dropout keep prob for hdrop
  unsigned long long seed, rng_offset;
  long long* trace;               // optional [T][8] clock64 samples from CTA 0 (profiling aid)
};

#define REC_TRACE(step, slot) do { if (p.trace && blockIdx.x == 0) p.trace[(step) * 8 + (slot)] = clock64(); } while (0)
#define REC_TRACE_G0(step, slot) do { if (sub == 0) REC_TRACE(step, slot); } while (0)   // backward: group 0 of the CTA only

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
template <int NT> __device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

constexpr uint32_t REC_SENTINEL = 0xFFFFFFFFu;       // "not written yet" marker of an hseq word (two bf16)
constexpr uint32_t REC_MAX_SPINS = 1u << 24;           // bounded polling: a lost peer traps instead of hanging the GPU

#ifndef B2T_FWD_POLL_GROUP
#define B2T_FWD_POLL_GROUP 6      // chunks of h_{t-1} polled together by the forward loaders (16 = all of them in one round); measured on B200: 6 -> 4.24 ms per step, 16 -> 4.32, 4 -> 4.29, 3 -> 4.35
#endif
#ifndef B2T_POLL_RELAXED
#define B2T_POLL_RELAXED 0
#endif
// Poll load of the backward exchanges.  tools/ubench/l2_signal.cu: a ping-pong through L2 takes ~1200 cycles per round trip with
// ld.relaxed.gpu polls and far less with ld.global.cg polls (L2 only, never L1), so the cache-global flavour is the default.
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* ptr) {
  uint4 v;
#if B2T_POLL_RELAXED
  asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
#else
  asm volatile("ld.global.cg.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
#endif
  return v;
}
// Poll load for the forward exchange: cache-global (L2 only, never L1), so every execution reads the point of
// coherence; unlike ld.relaxed.gpu the 16-byte accesses of a warp are coalesced into full-line requests.
__device__ __forceinline__ uint4 ld_l2_v4(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.cg.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_v2(void* ptr, uint32_t a, uint32_t b) {
  asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1,%2};" ::"l"(ptr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_relaxed_v4(void* ptr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(ptr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ bool has_sentinel(const uint4& v) {
  return v.x == REC_SENTINEL || v.y == REC_SENTINEL || v.z == REC_SENTINEL || v.w == REC_SENTINEL;
}
__device__ __forceinline__ bool tags_match(const uint4& v, uint32_t tag) {
  return (((v.x ^ tag) | (v.y ^ tag) | (v.z ^ tag) | (v.w ^ tag)) & 3u) == 0u;
}

// A operand from TMEM, B operand from shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// dropout decision for element (row m, unit j) of a [*, H] activation; same function in fwd and bwd.
__device__ __forceinline__ uint4 rec_dropout_bits(unsigned long long seed, unsigned long long offset, unsigned long long elem4) {
  const unsigned long long c = elem4 + offset;
  return philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x6a7eu, 0), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// One group of up to PC chunks of a K-major B operand: poll the 16-byte units this thread owns (UPT per chunk, g_unit bytes
// apart in global memory, s_unit bytes apart in shared memory) until no word is the sentinel, stage them (the caller's sdst
// already carries the swizzled offset of the thread's first unit) and report each chunk once the whole warp has staged it.
// A polling round (re)loads every pending chunk with all loads in flight together.
template <int PC, int UPT, typename AddrF, typename DoneF>
__device__ __forceinline__ void poll_and_stage(int nc, bool active, AddrF gaddr, size_t g_unit, uint8_t* sdst, int s_chunk, int s_unit, DoneF done) {
  uint4 v[PC * UPT];
  uint32_t pending = (1u << nc) - 1u;                        // warp-uniform: chunks of this group not staged yet
  uint32_t spins = 0;
  while (pending) {
#pragma unroll
    for (int c = 0; c < PC; ++c)
      if (((pending >> c) & 1u) && active) {
        const uint8_t* src = gaddr(c);
#pragma unroll
        for (int u = 0; u < UPT; ++u) v[c * UPT + u] = ld_l2_v4(src + u * g_unit);
      }
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      if ((pending >> c) & 1u) {
        bool ok = true;
        if (active) {
#pragma unroll
          for (int u = 0; u < UPT; ++u) ok = ok && !has_sentinel(v[c * UPT + u]);
        }
        if (__all_sync(0xffffffffu, ok)) {
          if (active) {
#pragma unroll
            for (int u = 0; u < UPT; ++u) *reinterpret_cast<uint4*>(sdst + c * s_chunk + u * s_unit) = v[c * UPT + u];
            fence_proxy_async_smem();                        // generic smem write -> visible to the tensor-core (async) proxy
          }
          __syncwarp();
          done(c);
          pending &= ~(1u << c);
        }
      }
    }
    if (++spins > REC_MAX_SPINS) __trap();
  }
}

template <int BG>
__global__ void __launch_bounds__(RecCfg<BG>::kFwdThreads, 1)
gru_rec_fwd_kernel(const RecFwdParams p) {
  using Cfg = RecCfg<BG>;
  constexpr int XP = Cfg::kXPitch;
  constexpr int CHUNK_BYTES = BG * 128;                  // BG rows x 64 bf16
  constexpr int UNITS = BG * 8;                          // 16-byte units per chunk (<= kLoadThreads)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KC = p.H / 64;                               // 64-wide contraction chunks (<= 16)
  uint8_t* sH = smem;                                    // [2 buffers][KC] x CHUNK_BYTES, K-major, 128B swizzle
  float* sX = reinterpret_cast<float*>(sH + 2 * KC * CHUNK_BYTES);   // [3][32][XP]
  uint64_t* bar_h = reinterpret_cast<uint64_t*>(sX + 3 * 32 * XP);   // [2][16]
  uint64_t* bar_d = bar_h + 32;
  uint64_t* bar_s = bar_d + 1;                           // this CTA's epilogue warps have stored their part of h_t
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.n_slices, grp = blockIdx.x / p.n_slices;
  const int j0 = slice * REC_US, b0 = grp * BG;
  const int a_cols = p.H / 2;                            // TMEM columns of the A operand; accumulator follows
  const bool is_loader = warp == 0 || warp >= 2 + Cfg::kEpiWarps;

  if (threadIdx.x == 0) {
    for (int c = 0; c < 32; ++c) mbar_init(&bar_h[c], Cfg::kLoadWarps);
    mbar_init(bar_d, 1);
    mbar_init(bar_s, Cfg::kEpiWarps);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<REC_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base + a_cols;

  // ---- one-time: W_hh slice -> TMEM.  Lane 32q+l holds gate q, unit j0+l; quarter 3 is zero.
  if (warp >= 2 && warp < 6) {
    const int q = warp & 3;
    const uint4* src = reinterpret_cast<const uint4*>(p.whh + ((size_t)(q < 3 ? q : 0) * p.H + j0 + lane) * p.H);
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
    // ---------------- loaders: poll h_{t-1} (all units of this batch group) out of L2 and stage it as the B operand.
    // Thread lt owns the 16-byte unit (row = lt / 8, segment = lt % 8) of every 64-column chunk.  Double-buffered:
    // the peers' h_t can land while this CTA's MMA of step t still reads h_{t-1}.
    const int lw = warp == 0 ? 0 : warp - (1 + Cfg::kEpiWarps);
    const int lt = lw * 32 + lane;
    constexpr int UPT = Cfg::kUnitsPerThread, PC = Cfg::kPollChunks;
    constexpr int ROWS_PER_PASS = Cfg::kLoadThreads / 8;   // unit u of a thread is row (lt / 8) + u * ROWS_PER_PASS
    const bool active = lt < UNITS / UPT;
    const int row = lt >> 3, seg = lt & 7;
    const uint32_t soff = row * 128 + ((seg ^ (row & 7)) << 4);      // ROWS_PER_PASS is a multiple of 8: same swizzle phase for every unit
    const size_t g_unit = (size_t)ROWS_PER_PASS * p.H * sizeof(__nv_bfloat16);
    for (int t = p.t_begin; t < p.t_end; ++t) {
      const int step = t - p.t_begin, buf = step & 1;
      // The peers store their slices of h_{t-1} at about the time this CTA stores its own: start polling then
      // (polling earlier only burns L2 bandwidth and de-phases the rounds from the arrival of the data).
      if (step > 0) {
        mbar_wait(bar_s, (uint32_t)(step - 1) & 1u);
        if (p.poll_delay > 0) {
          const long long t0 = clock64();
          while (clock64() - t0 < p.poll_delay) {}
        }
      }
      const uint8_t* g = reinterpret_cast<const uint8_t*>(p.hseq + ((size_t)t * p.Bpad + b0 + row) * p.H) + seg * 16;   // slot t = h_{t-1}
      uint8_t* sdst = sH + (size_t)buf * KC * CHUNK_BYTES + soff;
      if constexpr (PC >= 16) {
#if B2T_FWD_POLL_GROUP >= 16
        // all chunks (KC <= 16) in one group: no group loop (a run-time loop around the register array costs spills)
        uint64_t* bars = &bar_h[buf * 16];
        poll_and_stage<PC, UPT>(KC, active, [&](int c) { return g + c * 128; }, g_unit, sdst, CHUNK_BYTES, ROWS_PER_PASS * 128, [&](int c) {
          if (lane == 0) mbar_arrive(&bars[c]);
          if (lt == 0 && c == 0) REC_TRACE(t, 0);          // first chunk of h_{t-1} staged
        });
#else
        // groups of B2T_FWD_POLL_GROUP chunks, unrolled at compile time: a polling round moves fewer bytes, so the first
        // chunks reach the MMA (the step's critical resource) earlier
        constexpr int PG = B2T_FWD_POLL_GROUP;
#pragma unroll
        for (int gi = 0; gi < (16 + PG - 1) / PG; ++gi) {
          const int c0 = gi * PG;
          if (c0 < KC) {
            uint64_t* bars = &bar_h[buf * 16 + c0];
            const uint8_t* gg = g + c0 * 128;
            poll_and_stage<PG, UPT>(KC - c0 < PG ? KC - c0 : PG, active, [&](int c) { return gg + c * 128; }, g_unit,
                                    sdst + (size_t)c0 * CHUNK_BYTES, CHUNK_BYTES, ROWS_PER_PASS * 128, [&](int c) {
                                      if (lane == 0) mbar_arrive(&bars[c]);
                                      if (lt == 0 && c0 + c == 0) REC_TRACE(t, 0);
                                    });
          }
        }
#endif
      } else {
        for (int c0 = 0; c0 < KC; c0 += PC) {                // chunk groups in the order the MMA consumes them
          const uint8_t* gg = g + c0 * 128;
          uint64_t* bars = &bar_h[buf * 16 + c0];
          poll_and_stage<PC, UPT>(KC - c0 < PC ? KC - c0 : PC, active, [&](int c) { return gg + c * 128; }, g_unit,
                                  sdst + (size_t)c0 * CHUNK_BYTES, CHUNK_BYTES, ROWS_PER_PASS * 128, [&](int c) {
                                    if (lane == 0) mbar_arrive(&bars[c]);
                                    if (lt == 0 && c0 + c == 0) REC_TRACE(t, 0);
                                  });
        }
      }
      if (lt == 0) REC_TRACE(t, 1);                        // all chunks staged
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BG, 0, 0);
      for (int t = p.t_begin; t < p.t_end; ++t) {
        const int step = t - p.t_begin, buf = step & 1;
        const uint32_t par = (uint32_t)(step >> 1) & 1u;
        for (int c = 0; c < KC; ++c) {
          mbar_wait(&bar_h[buf * 16 + c], par);
          if (c == 0) REC_TRACE(t, 2);         // first operand chunk landed
          tc_fence_after();
          const uint32_t sb = smem_u32(sH + ((size_t)buf * KC + c) * CHUNK_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ts(tmem_d, tmem_base + (c * 4 + k) * 8, umma_smem_desc(sb + k * 32, 16, 1024), idesc, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_d);
        REC_TRACE(t, 3);                       // all MMAs issued
      }
    }
  } else {
    // ---------------- epilogue: thread e owns trial b0 + e/8 and units j0 + 4*(e%8) .. +3
    const int e = threadIdx.x - 64;
    const int ew = e >> 5;                     // epilogue warp index
    const int q = warp & 3;                    // TMEM lane quarter: 0 -> r rows, 1 -> z, 2 -> n, 3 -> unused
    const int chalf = ew >> 2;                 // which 16 accumulator columns this warp moves (BG = 32 has two halves)
    const int bl = e >> 3, u0 = (e & 7) * 4;
    const int b = b0 + bl, j = j0 + u0;
    float h[4], bh[3][4];
    {
      const float4 hv = *reinterpret_cast<const float4*>(p.h_state + (size_t)b * p.H + j);
      h[0] = hv.x; h[1] = hv.y; h[2] = hv.z; h[3] = hv.w;
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const float4 bv = *reinterpret_cast<const float4*>(p.bhh + g * p.H + j);
        bh[g][0] = bv.x; bh[g][1] = bv.y; bh[g][2] = bv.z; bh[g][3] = bv.w;
      }
    }
    const bool train = p.R != nullptr;
    const float inv_keep = 1.0f / p.keep;
    for (int t = p.t_begin; t < p.t_end; ++t) {
      const size_t row = (size_t)t * p.Bpad + b;
      // prefetch the input projection while the MMA runs
      float4 gxv[3];
#pragma unroll
      for (int g = 0; g < 3; ++g) gxv[g] = __ldg(reinterpret_cast<const float4*>(p.gx + row * 3 * p.H + g * p.H + j));

      mbar_wait(bar_d, (uint32_t)(t - p.t_begin) & 1u);
      if (e == 0) REC_TRACE(t, 4);             // accumulator complete
      tc_fence_after();
      if (q < 3) {
        uint32_t v[16];
        tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + chalf * 16, v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(sX + (q * 32 + lane) * XP + chalf * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
      tc_fence_before();
      // The only CTA barrier of the step.  sX and the accumulator are not overwritten before every thread here has
      // stored its part of h_t: the next accumulator needs all of h_t, including this CTA's own slice.
      epi_bar_sync<Cfg::kEpiThreads>();
      if (e == 0) REC_TRACE(t, 7);             // gates exchanged
      float hn[4], r[4], z[4], n[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float ar = sX[(0 * 32 + u0 + i) * XP + bl];
        const float az = sX[(1 * 32 + u0 + i) * XP + bl];
        const float an = sX[(2 * 32 + u0 + i) * XP + bl];
        const float gr = (&gxv[0].x)[i], gz = (&gxv[1].x)[i], gn = (&gxv[2].x)[i];
        r[i] = sigmoid_f(gr + ar + bh[0][i]);
        z[i] = sigmoid_f(gz + az + bh[1][i]);
        hn[i] = an + bh[2][i];
        n[i] = tanh_f(gn + r[i] * hn[i]);
        h[i] = (1.0f - z[i]) * n[i] + z[i] * h[i];
      }
      const size_t off = row * p.H + j;
      {  // publish h_t: the data is the signal (see the header); the BPTT stash below is off the peers' critical path
        __nv_bfloat162 lo = __floats2bfloat162_rn(h[0], h[1]), hi = __floats2bfloat162_rn(h[2], h[3]);
        st_relaxed_v2(p.hseq + ((size_t)(t + 1) * p.Bpad + b) * p.H + j, *reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_s);
      if (e == 0) REC_TRACE(t, 5);             // h_t stored
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
    }
    *reinterpret_cast<float4*>(p.h_state + (size_t)b * p.H + j) = make_float4(h[0], h[1], h[2], h[3]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<REC_TMEM_COLS>(tmem_base);
  }
}

// --------------------------------------------------------------------------------------------
// Backward recurrence (BPTT).  Per step t (descending), with dh_t the total gradient wrt h_t:
//   dn = dh*(1-z)  dz = dh*(h_{t-1}-n)  dn_pre = dn*(1-n^2)  dz_pre = dz*z*(1-z)
//   dr_pre = dn_pre*hn*r*(1-r)
//   dGx_t = [dr_pre, dz_pre, dn_pre]        dGh_t = [dr_pre, dz_pre, dn_pre*r]
//   dh_{t-1} = dh*z + dGh_t W_hh + dY_{t-1}
// dW_ih/dW_hh are GEMMs over dGx/dGh afterwards; the bias gradients are accumulated here in registers.
struct RecBwdParams {
  int H, Bpad, n_slices;
  int t_begin, t_end;               // this chunk processes t = t_end-1 ... t_begin
  int T;                            // total steps
  int gen_base;                     // generation of this launch's first step (see the header): buffer = gen & 1, tag = (gen >> 1) & 3
  const float* dY;                  // [T][Bpad][H] fp32 gradient wrt this layer's (dropped) output
  const __nv_bfloat16* hseq;        // [(T+1)][Bpad][H]
  const __nv_bfloat16 *R, *Z, *Nn, *HN;
  const __nv_bfloat16* whh;         // [3H][H] bf16
  __nv_bfloat16* dGx;               // [T][Bpad][3H]
  __nv_bfloat16* dGh;               // [T][Bpad][3H]
  float* part;                      // [2][n_groups][n_slices(dest)][n_slices(src)][BG][32] fp32 partial sums of dh
  float* dbih;                      // [3H] (atomicAdd)
  float* dbhh;                      // [3H] (atomicAdd)
  float* dh_state;                  // [Bpad][H] fp32: recurrent part of dh_{t_end-1} on entry (ignored when first_chunk),
                                    //                  recurrent part of dh_{t_begin-1} on exit (= grad wrt the initial state at t_begin = 0)
  int first_chunk;                  // 1: t_end == T, no incoming recurrent gradient
  int n_valid;                      // trials < n_valid contribute (pad trials are masked out)
  int poll_delay;                   // gru_rec_bwd2_kernel: cycles between publishing dG_t and the first polling round
  float keep;                       // dropout applied to this layer's output in forward (1 => none)
  unsigned long long seed, rng_offset;
  long long* trace;                 // optional [T][8] clock64 samples from CTA 0
};

// NSUB independent batch groups of BG trials share one CTA (and one copy of W_slice^T): each group has its own epilogue
// warps, dG operand, accumulator columns, barriers and exchange blocks, and runs the step loop on its own.  A group's step is
// a chain of latencies (gather partials through L2 -> gate math -> MMAs -> drain -> publish); with two groups per CTA the
// chain of one overlaps the other's, which nearly doubles the trials a CTA advances per unit time.
template <int BG, int NSUB>
__global__ void __launch_bounds__((RecCfg<BG>::kBwdThreads + 32) * NSUB, 1)
gru_rec_bwd_kernel(const RecBwdParams p) {
  using Cfg = RecCfg<BG>;
  constexpr int CHUNK_BYTES = BG * 128;
  constexpr int NTHREADS = (Cfg::kBwdThreads + 32) * NSUB;  // per group: kBwdThreads epilogue threads, then (after all of those) one MMA-issuing warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int MB = (p.H + 127) / 128;                      // 128-row blocks of the output (all hidden units)
  const int A_CHUNK = MB * 128 * 128;                    // bytes of one 64-kk chunk of the A operand
  const bool is_issuer = threadIdx.x >= Cfg::kBwdThreads * NSUB;
  const int sub = is_issuer ? (threadIdx.x - Cfg::kBwdThreads * NSUB) >> 5 : threadIdx.x / Cfg::kBwdThreads;   // batch group within the CTA (warp-uniform)
  uint8_t* sA = smem;                                    // W_slice^T, K-major A operand: [2 chunks of 64 kk][MB*128 rows k][128 B], 128B swizzle
  uint8_t* sB_all = sA + 2 * A_CHUNK;                    // per group, 2 chunks: dG_t as K-major B operand [BG trials][128 kk] (kk = gate*32 + unit, 96 used)
  uint8_t* sB = sB_all + sub * 2 * CHUNK_BYTES;
  uint64_t* bar_all = reinterpret_cast<uint64_t*>(sB_all + NSUB * 2 * CHUNK_BYTES);
  uint64_t* bar_d = bar_all + sub * 8;                   // [8] per group: one per 128-row output block
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_all + NSUB * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NS = p.n_slices, NG = (gridDim.x / NS) * NSUB;
  const int slice = blockIdx.x % NS, grp = (blockIdx.x / NS) * NSUB + sub;
  const int j0 = slice * REC_US, b0 = grp * BG;
  const int nsteps = p.t_end - p.t_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 8 * NSUB; ++i) mbar_init(&bar_all[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<REC_TMEM_COLS>(tmem_slot);
  for (int i = threadIdx.x; i < NSUB * 2 * CHUNK_BYTES / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(sB_all)[i] = 0u;

  // ---- one-time: W_slice^T -> shared memory (the accumulators of all MB output blocks need the whole TMEM for BG = 64,
  //      and an A operand read from shared memory dispatches faster than one read from TMEM for N <= 64, see
  //      profiles/r1_mma_dispatch_microbench.md).  Row k (output unit), column kk = gate*32 + unit; one item = two
  //      adjacent kk (same gate, units u and u+1) x 8 consecutive k: two 16 B global loads, eight 4 B smem stores.
  //      Consecutive lanes take consecutive kk pairs: the stores of a warp fall into one 128 B row (conflict free).
  {
    const int K8 = (MB * 128) / 8;
    for (int it = threadIdx.x; it < K8 * 48; it += NTHREADS) {
      const int pr = it % 48, k8 = it / 48;
      const int kk = 2 * pr, k = k8 * 8;
      uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
      if (k < p.H) {
        const __nv_bfloat16* src = p.whh + ((size_t)(kk >> 5) * p.H + j0 + (kk & 31)) * p.H + k;
        lo = __ldg(reinterpret_cast<const uint4*>(src));
        hi = __ldg(reinterpret_cast<const uint4*>(src + p.H));      // kk+1 is the next unit of the same gate
      }
      const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w}, h[4] = {hi.x, hi.y, hi.z, hi.w};
      uint8_t* base = sA + (kk >> 6) * A_CHUNK + (kk & 7) * 2;
      const int ku = (kk & 63) >> 3;                         // 16 B unit of the row
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t a = (l[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu, b = (h[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
        const int kr = k + i;
        *reinterpret_cast<uint32_t*>(base + kr * 128 + ((ku ^ (kr & 7)) << 4)) = a | (b << 16);
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base + sub * MB * BG;       // this group's accumulators: MB blocks of BG columns

  // Steps are indexed s = 0..nsteps-1 for t = t_end-1-s.  The partials published at step s feed dh of step s+1.
  constexpr uint32_t idesc = umma_idesc_bf16(128, BG, 0, 0);
  if (is_issuer) {
    // ---------------- MMA issuer of this group: waits until the epilogue warps have written dG_t, then issues block by block
    // (the drain of block mb by the epilogue warps overlaps the MMAs of blocks mb+1..)
    const uint32_t sa = smem_u32(sA), sb = smem_u32(sB);
    for (int s = 0; s < nsteps; ++s) {
      asm volatile("bar.sync %0, %1;" ::"r"(1 + sub), "n"(Cfg::kEpiThreads + 32) : "memory");
      if (elect_one()) {
        if (sub == 0) REC_TRACE(s, 2);
        tc_fence_after();
        for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
          for (int ks = 0; ks < 6; ++ks) {               // K = 96 = 6 x 16
            const uint64_t adesc = umma_smem_desc(sa + (ks >> 2) * A_CHUNK + mb * (128 * 128) + (ks & 3) * 32, 16, 1024);
            const uint64_t bdesc = umma_smem_desc(sb + (ks >> 2) * CHUNK_BYTES + (ks & 3) * 32, 16, 1024);
            umma_bf16(tmem_d + mb * BG, adesc, bdesc, idesc, ks != 0 ? 1u : 0u);
          }
          umma_commit(&bar_d[mb]);
        }
        if (sub == 0) REC_TRACE(s, 3);
      }
      __syncwarp();
    }
  } else {
    const int e = threadIdx.x - sub * Cfg::kBwdThreads;
    const int ew = e >> 5;
    const int q = warp & 3;                                // TMEM lane quarter this warp may read (kEpiWarps is a multiple of 4)
    const int chalf = ew >> 2;
    const int bl = e >> 3, u0 = (e & 7) * 4;
    const int b = b0 + bl, j = j0 + u0;
    const bool valid = b < p.n_valid;
    const float inv_keep = 1.0f / p.keep;
    float carry[4] = {0.f, 0.f, 0.f, 0.f};               // dh_{t+1} * z_{t+1}
    float accx[3][4], acch[4];                           // bias-gradient partial sums (dGh differs only in n)
#pragma unroll
    for (int i = 0; i < 4; ++i) { accx[0][i] = accx[1][i] = accx[2][i] = 0.f; acch[i] = 0.f; }

    // Partial block layout (one per (buffer, group, destination slice, source slice)): 16-byte granules
    // [trial quad tq][unit 0..31][4 trials], so that the producer (one TMEM lane = one unit, consecutive trials in its
    // registers) and the consumer (warp ew = trial quad ew, lane = unit) both move 512 contiguous bytes per warp access.
    auto gather_partials = [&](int gen, float (&P)[4]) {
      // sum over the NS source CTAs; a word is valid once its two low bits carry the generation tag
      const uint32_t tag = (uint32_t)(gen >> 1) & 3u;
      const float* base = p.part + ((((size_t)(gen & 1) * NG + grp) * NS + slice) * NS) * (BG * 32) + (ew * 32 + lane) * 4;
      float G[4] = {0.f, 0.f, 0.f, 0.f};                  // (unit = lane, trials 4*ew .. 4*ew+3)
      constexpr int NB = NTHREADS > 256 ? 12 : 24;          // partial blocks polled together (register budget: 128 regs at 512 threads)
      for (int src0 = 0; src0 < NS; src0 += NB) {
        uint4 v[NB];
        uint32_t pending = 0;
#pragma unroll
        for (int i = 0; i < NB; ++i)
          if (src0 + i < NS) pending |= 1u << i;
        uint32_t spins = 0;
        while (true) {
#pragma unroll
          for (int i = 0; i < NB; ++i)
            if ((pending >> i) & 1u) v[i] = ld_relaxed_v4(base + (size_t)(src0 + i) * (BG * 32));
#pragma unroll
          for (int i = 0; i < NB; ++i)
            if (((pending >> i) & 1u) && tags_match(v[i], tag)) pending &= ~(1u << i);
          if (!pending) break;
          if (++spins > REC_MAX_SPINS) __trap();
        }
#pragma unroll
        for (int i = 0; i < NB; ++i)                       // fixed summation order: results do not depend on arrival order
          if (src0 + i < NS) {
            G[0] += __uint_as_float(v[i].x & ~3u); G[1] += __uint_as_float(v[i].y & ~3u);
            G[2] += __uint_as_float(v[i].z & ~3u); G[3] += __uint_as_float(v[i].w & ~3u);
          }
      }
      // transpose inside the warp: this thread needs trial k = lane/8 of units 4*(lane%8) + i
      const int k = lane >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sl = 4 * (lane & 7) + i;
        const float a0 = __shfl_sync(0xffffffffu, G[0], sl), a1 = __shfl_sync(0xffffffffu, G[1], sl);
        const float a2 = __shfl_sync(0xffffffffu, G[2], sl), a3 = __shfl_sync(0xffffffffu, G[3], sl);
        P[i] = k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
      }
    };

    // BPTT stash of one step, kept as raw words: the loads for step s+1 are issued while step s drains its accumulators,
    // so their latency never sits between two exchanges.
    struct Stash { uint2 r, z, n, hn, hp; float4 dy; };
    auto load_stash = [&](int t) {
      const size_t off = ((size_t)t * p.Bpad + b) * p.H + j;
      Stash st;
      st.r = __ldg(reinterpret_cast<const uint2*>(p.R + off)); st.z = __ldg(reinterpret_cast<const uint2*>(p.Z + off));
      st.n = __ldg(reinterpret_cast<const uint2*>(p.Nn + off)); st.hn = __ldg(reinterpret_cast<const uint2*>(p.HN + off));
      st.hp = *reinterpret_cast<const uint2*>(p.hseq + off);             // slot t = h_{t-1}
      st.dy = __ldg(reinterpret_cast<const float4*>(p.dY + off));
      return st;
    };
    auto unpack = [](const uint2& u, float (&f)[4]) {
      const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
      f[0] = __low2float(lo); f[1] = __high2float(lo); f[2] = __low2float(hi); f[3] = __high2float(hi);
    };
    Stash cur = load_stash(p.t_end - 1);

    for (int s = 0; s < nsteps; ++s) {
      const int t = p.t_end - 1 - s;
      const size_t row = (size_t)t * p.Bpad + b;
      const size_t off = row * p.H + j;
      float dmask[4] = {1.0f, 1.0f, 1.0f, 1.0f};         // dropout of this layer's output (same Philox stream as forward)
      if (p.keep < 1.0f) {
        const uint4 rnd = rec_dropout_bits(p.seed, p.rng_offset, off >> 2);
        const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) dmask[i] = (u32_to_unit(rr[i]) < p.keep) ? inv_keep : 0.0f;
      }
      float P[4] = {0.f, 0.f, 0.f, 0.f};
      if (s > 0) {
        if (e == 0) REC_TRACE_G0(s, 0);
        gather_partials(p.gen_base + s - 1, P);          // polls until the partials of step s-1 from every CTA of the group are there
        if (e == 0) REC_TRACE_G0(s, 1);
      }
      float r[4], z[4], n[4], hn[4], hp[4];
      unpack(cur.r, r); unpack(cur.z, z); unpack(cur.n, n); unpack(cur.hn, hn); unpack(cur.hp, hp);
      const float dy[4] = {cur.dy.x * dmask[0], cur.dy.y * dmask[1], cur.dy.z * dmask[2], cur.dy.w * dmask[3]};
      float dh[4];
      if (s > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dh[i] = carry[i] + P[i] + dy[i];
      } else if (!p.first_chunk) {                       // recurrent gradient handed over by the later chunk
        const float4 c4 = *reinterpret_cast<const float4*>(p.dh_state + (size_t)b * p.H + j);
        dh[0] = c4.x + dy[0]; dh[1] = c4.y + dy[1]; dh[2] = c4.z + dy[2]; dh[3] = c4.w + dy[3];
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
      // dGh_t -> shared memory as the K-major, 128B-swizzled B operand: row = trial, kk = gate*32 + unit
      {
        const float* gsel[3] = {gr, gz, gnh};
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const int kk = g * 32 + u0;                    // 4 consecutive kk, 8 bytes
          const int chunk = kk >> 6, kl = kk & 63;
          uint8_t* dst = sB + chunk * CHUNK_BYTES + bl * 128 + ((((kl >> 3) ^ (bl & 7)) & 7) << 4) + (kl & 7) * 2;
          __nv_bfloat162 lo = __floats2bfloat162_rn(gsel[g][0], gsel[g][1]), hi = __floats2bfloat162_rn(gsel[g][2], gsel[g][3]);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&lo);
          u.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(dst) = u;
        }
      }
      fence_proxy_async_smem();                          // generic smem writes -> visible to the tensor-core (async) proxy
      tc_fence_before();
      asm volatile("bar.arrive %0, %1;" ::"r"(1 + sub), "n"(Cfg::kEpiThreads + 32) : "memory");   // hand dG_t to the issuer warp
      if (s + 1 < nsteps) cur = load_stash(t - 1);       // in flight during the drain below
      // partial sums -> L2 (granule layout above): lane = output unit within its 32-unit destination slice
      const int gen = p.gen_base + s;
      const uint32_t tag = (uint32_t)(gen >> 1) & 3u;
      for (int mb = 0; mb < MB; ++mb) {
        mbar_wait(&bar_d[mb], s & 1);
        if (e == 0 && mb == 0) REC_TRACE_G0(s, 4);
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + mb * BG + chalf * 16, v);
        tmem_ld_wait();
        const int dest = mb * 4 + q;                     // destination slice = k / 32
        if (dest < NS) {
          float* dst = p.part + ((((size_t)(gen & 1) * NG + grp) * NS + dest) * NS + slice) * (BG * 32) + ((chalf * 4) * 32 + lane) * 4;
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4)
            st_relaxed_v4(dst + g4 * 128, (v[4 * g4] & ~3u) | tag, (v[4 * g4 + 1] & ~3u) | tag, (v[4 * g4 + 2] & ~3u) | tag, (v[4 * g4 + 3] & ~3u) | tag);
        }
      }
      // off the critical path: gate gradients for the weight-gradient GEMMs
      const size_t goff = row * 3 * p.H + j;
      st_bf16x4(p.dGh + goff, gr[0], gr[1], gr[2], gr[3]);
      st_bf16x4(p.dGh + goff + p.H, gz[0], gz[1], gz[2], gz[3]);
      st_bf16x4(p.dGh + goff + 2 * p.H, gnh[0], gnh[1], gnh[2], gnh[3]);
      st_bf16x4(p.dGx + goff, gr[0], gr[1], gr[2], gr[3]);
      st_bf16x4(p.dGx + goff + p.H, gz[0], gz[1], gz[2], gz[3]);
      st_bf16x4(p.dGx + goff + 2 * p.H, gn[0], gn[1], gn[2], gn[3]);

      if (e == 0) REC_TRACE_G0(s, 5);
      // No barrier here: sB is rewritten only after this thread has gathered the partials of step s from every CTA
      // (so MMA(s) is long complete), and MMA(s+1) is issued behind the named barrier above, at which every epilogue
      // warp arrives only after it has drained the accumulators of step s.
      tc_fence_before();
    }
    // recurrent gradient for the step before this chunk: dh_{t_begin-1} (rec) = dh_{t_begin} * z + dGh_{t_begin} W_hh
    {
      float P[4];
      gather_partials(p.gen_base + nsteps - 1, P);
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = valid ? carry[i] + P[i] : 0.0f;
      *reinterpret_cast<float4*>(p.dh_state + (size_t)b * p.H + j) = make_float4(o[0], o[1], o[2], o[3]);
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
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<REC_TMEM_COLS>(tmem_base);
  }
}

}  // namespace b2t
