// Backward GRU recurrence, two-dimensional decomposition (used when H is a multiple of 256).
//
//   dh_{t-1} = dh_t * z_t + dGh_t W_hh + dY_{t-1},   dGh_t = gate-math(dh_t, stash_t)          (see gru_rec.cuh)
//
// gru_rec_bwd_kernel splits the contraction 24 ways (every CTA contracts its own 96 gate columns against ALL output
// units) and pays for it with a 24-way reduce-scatter of fp32 partial sums: 98 KB written and 98 KB read per CTA and
// time step (BG = 32), which is what bounds that kernel (a B200 SM stores ~30 B/clk towards L2).
//
// Here the H/32 CTAs of a batch group form a (H/128) x 4 grid.  CTA (mb, kq)
//   * keeps W_hh^T[128 output units of block mb][gate columns of the units kq*H/4 .. (kq+1)*H/4) ] in shared memory
//     (128 x 3H/4 bf16 = 147 KB at H = 768) as the K-major A operand,
//   * polls dGh_t of those H/4 units x 3 gates (bf16, [BG][3H/4] = 36 KB) out of the dGh array itself, exactly like
//     the forward kernel polls h_{t-1} out of hseq: the array is pre-filled with a sentinel and the data is the signal,
//   * runs 3H/64 = 36 tcgen05.mma (M = 128, N = BG, K = 16) into a [128][BG] fp32 accumulator,
//   * sends the three quarters of it that belong to the other kq of its block to those CTAs (3 x 4 KB, tagged fp32 words
//     as in gru_rec_bwd_kernel) and receives 3 x 4 KB: a 4-way instead of a 24-way reduction,
//   * finalises dh for its own 32 units (mb*128 + kq*32 ..), does the gate math and stores dGh_t / dGx_t for them.
// Per CTA and step: 36 + 12 KB read, 12 + 12 KB written instead of 98 + 98 KB; the price is two exchanges per step
// (dG all-gather, 4-way partial reduction) instead of one.
//
// Deadlock freedom / ordering: every wait is on data of an earlier point of the same dependency chain; cooperative launch
// makes the CTAs co-resident; all polling loops are bounded (trap instead of hang).
#pragma once
#include "gru_rec.cuh"

namespace b2t {

// Optional cross-CTA timing of one step (profiling aid): every CTA writes %globaltimer at five points of step REC_SKEW_STEP
// behind the per-step cycle trace (slots [T*8 + cta*8 + k]).
constexpr int REC_SKEW_STEP = 40;
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define REC_SKEW(step, k) do { if (p.trace && (step) == REC_SKEW_STEP) p.trace[(size_t)p.T * 8 + blockIdx.x * 8 + (k)] = global_ns(); } while (0)

template <int BG> struct RecBwd2Cfg {
  using Base = RecCfg<BG>;
  static constexpr int kThreads = Base::kFwdThreads;   // warp 0 + last 7 warps: loaders; warp 1: MMA issuer / TMEM owner; then BG/4 epilogue warps
  static constexpr size_t smem_bytes(int H) {
    return (size_t)(3 * H / 4 / 64) * (128 * 128 + BG * 128) + (size_t)16 * 8 + 64 + 1024;
  }
};

template <int BG>
__global__ void __launch_bounds__(RecBwd2Cfg<BG>::kThreads, 1)
gru_rec_bwd2_kernel(const RecBwdParams p) {
  using Cfg = RecCfg<BG>;
  constexpr int CHUNK_BYTES = BG * 128;                  // BG trials x 64 bf16 of the contraction
  constexpr int A_CHUNK = 128 * 128;                     // 128 output units x 64 bf16
  constexpr int UNITS = BG * 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KQ = p.H / 4;                                // units of this CTA's contraction range
  const int CPG = KQ / 64;                               // 64-wide chunks per gate
  const int KC = 3 * CPG;                                // chunks of the contraction (<= 9)
  const int MB = p.H / 128;
  uint8_t* sA = smem;                                    // [KC][128 rows][128 B]
  uint8_t* sB = sA + (size_t)KC * A_CHUNK;               // [KC][BG rows][128 B]
  uint64_t* bar_h = reinterpret_cast<uint64_t*>(sB + (size_t)KC * CHUNK_BYTES);   // [16] chunk staged
  uint64_t* bar_d = bar_h + 16;                          // accumulator complete
  uint64_t* bar_s = bar_d + 1;                           // epilogue warps have published dG_t of the current step
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_group = MB * 4;
  const int grp = blockIdx.x / per_group, rr = blockIdx.x % per_group;
  const int mb = rr >> 2, kq = rr & 3;
  const int NG = gridDim.x / per_group;
  const int j0 = mb * 128 + kq * REC_US, b0 = grp * BG;  // the units / trials this CTA finalises
  const int nsteps = p.t_end - p.t_begin;
  const bool is_loader = warp == 0 || warp >= 2 + Cfg::kEpiWarps;

  if (threadIdx.x == 0) {
    for (int c = 0; c < 16; ++c) mbar_init(&bar_h[c], Cfg::kLoadWarps);
    mbar_init(bar_d, 1);
    mbar_init(bar_s, Cfg::kEpiWarps);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);

  // ---- one-time: A operand.  Row i = output unit mb*128 + i; column kk = g*KQ + uu is gate g of unit kq*KQ + uu, i.e.
  //      W_hh[g*H + kq*KQ + uu][mb*128 + i].  One item = two adjacent kk x 8 consecutive rows (two 16 B loads, eight 4 B stores).
  {
    const int KK = 3 * KQ, NP = KK / 2;
    for (int it = threadIdx.x; it < 16 * NP; it += RecBwd2Cfg<BG>::kThreads) {
      const int pr = it % NP, k8 = it / NP;
      const int kk = 2 * pr, g = kk / KQ, uu = kk - g * KQ;
      const __nv_bfloat16* src = p.whh + ((size_t)g * p.H + kq * KQ + uu) * p.H + mb * 128 + k8 * 8;
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(src));
      const uint4 hi = __ldg(reinterpret_cast<const uint4*>(src + p.H));     // uu + 1 (KQ is even: same gate)
      const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w}, h[4] = {hi.x, hi.y, hi.z, hi.w};
      uint8_t* base = sA + (size_t)(kk >> 6) * A_CHUNK + (kk & 7) * 2;
      const int ku = (kk & 63) >> 3;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t a = (l[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu, b = (h[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
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

  // Steps are indexed s = 0..nsteps-1 for t = t_end-1-s.
  if (is_loader) {
    // ---------------- loaders: poll dGh_t (this CTA's contraction range, all trials of the group) and stage it as the B operand.
    const int lw = warp == 0 ? 0 : warp - (1 + Cfg::kEpiWarps);
    const int lt = lw * 32 + lane;
    constexpr int UPT = Cfg::kUnitsPerThread;
    constexpr int PC = UPT == 1 ? 9 : 5;
    constexpr int ROWS_PER_PASS = Cfg::kLoadThreads / 8;
    const bool active = lt < UNITS / UPT;
    const int row = lt >> 3, seg = lt & 7;
    const uint32_t soff = row * 128 + ((seg ^ (row & 7)) << 4);
    const size_t g_unit = (size_t)ROWS_PER_PASS * 3 * p.H * sizeof(__nv_bfloat16);
    for (int s = 0; s < nsteps; ++s) {
      const int t = p.t_end - 1 - s;
      // Start polling when this CTA's own epilogue has published its dG_t: the peers do so at about the same time.  This
      // also orders the staging after MMA(s-1), which read the same (single) buffer: an epilogue warp publishes dG of step s
      // only after it has drained the accumulator of step s-1.
      mbar_wait(bar_s, (uint32_t)s & 1u);
      if (p.poll_delay > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < p.poll_delay) {}
      }
      const uint8_t* g = reinterpret_cast<const uint8_t*>(p.dGh + ((size_t)t * p.Bpad + b0 + row) * 3 * p.H + kq * KQ) + seg * 16;
      auto chunk_addr = [&](int cc) {
        const int gate = cc / CPG, ci = cc - gate * CPG;
        return g + ((size_t)gate * p.H + ci * 64) * sizeof(__nv_bfloat16);
      };
      if constexpr (PC >= 9) {                                          // all chunks (KC <= 9) in one group
        poll_and_stage<PC, UPT>(KC, active, chunk_addr, g_unit, sB + soff, CHUNK_BYTES, ROWS_PER_PASS * 128, [&](int c) {
          if (lane == 0) mbar_arrive(&bar_h[c]);
          if (lt == 0 && c == 0) REC_TRACE(s, 0);                      // first chunk of dG_t staged
        });
      } else {
        for (int c0 = 0; c0 < KC; c0 += PC) {
          poll_and_stage<PC, UPT>(KC - c0 < PC ? KC - c0 : PC, active, [&](int c) { return chunk_addr(c0 + c); }, g_unit,
                                  sB + (size_t)c0 * CHUNK_BYTES + soff, CHUNK_BYTES, ROWS_PER_PASS * 128, [&](int c) {
                                    if (lane == 0) mbar_arrive(&bar_h[c0 + c]);
                                    if (lt == 0 && c0 + c == 0) REC_TRACE(s, 0);
                                  });
        }
      }
      if (lt == 0) { REC_TRACE(s, 1); REC_SKEW(s, 1); }                 // all chunks staged
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BG, 0, 0);
      const uint32_t sa = smem_u32(sA), sb = smem_u32(sB);
      for (int s = 0; s < nsteps; ++s) {
        const uint32_t par = (uint32_t)s & 1u;
        // every epilogue warp has drained the accumulator of step s-1 (it publishes dG of step s only afterwards)
        mbar_wait(bar_s, par);
        for (int c = 0; c < KC; ++c) {
          mbar_wait(&bar_h[c], par);
          if (c == 0) REC_TRACE(s, 2);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_d, umma_smem_desc(sa + c * A_CHUNK + k * 32, 16, 1024), umma_smem_desc(sb + c * CHUNK_BYTES + k * 32, 16, 1024),
                      idesc, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar_d);
        REC_TRACE(s, 3); REC_SKEW(s, 2);
      }
    }
  } else {
    // ---------------- epilogue: thread e owns trial b0 + e/8 and units j0 + 4*(e%8) .. +3
    const int e = threadIdx.x - 64;
    const int ew = e >> 5;
    const int q = warp & 3;                    // TMEM lane quarter = destination kq of the rows this warp drains
    const int chalf = ew >> 2;                 // which 16 accumulator columns (trials) this warp drains
    const int bl = e >> 3, u0 = (e & 7) * 4;
    const int b = b0 + bl, j = j0 + u0;
    const bool valid = b < p.n_valid;
    const float inv_keep = 1.0f / p.keep;
    float carry[4] = {0.f, 0.f, 0.f, 0.f};
    float accx[3][4], acch[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { accx[0][i] = accx[1][i] = accx[2][i] = 0.f; acch[i] = 0.f; }

    // Partial blocks: part[buffer][group][mb][dest kq][src kq][BG*32 floats], granules [trial quad][unit][4 trials] (as in
    // gru_rec_bwd_kernel: producer lane = unit, consumer warp = trial quad).  The own quarter travels the same way.
    auto block_of = [&](int gen, int dest, int src) {
      return p.part + (((((size_t)(gen & 1) * NG + grp) * MB + mb) * 4 + dest) * 4 + src) * (BG * 32);
    };
    auto reduce_partials = [&](int gen, float (&P)[4]) {
      const uint32_t tag = (uint32_t)(gen >> 1) & 3u;
      const float* base = block_of(gen, kq, 0) + (ew * 32 + lane) * 4;
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
        if (++spins > REC_MAX_SPINS) __trap();
      }
      float G[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {                          // fixed summation order
        G[0] += __uint_as_float(v[i].x & ~3u); G[1] += __uint_as_float(v[i].y & ~3u);
        G[2] += __uint_as_float(v[i].z & ~3u); G[3] += __uint_as_float(v[i].w & ~3u);
      }
      const int k = lane >> 3;                               // this thread needs trial k of the quad, units 4*(lane%8) + i
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sl = 4 * (lane & 7) + i;
        const float a0 = __shfl_sync(0xffffffffu, G[0], sl), a1 = __shfl_sync(0xffffffffu, G[1], sl);
        const float a2 = __shfl_sync(0xffffffffu, G[2], sl), a3 = __shfl_sync(0xffffffffu, G[3], sl);
        P[i] = k == 0 ? a0 : (k == 1 ? a1 : (k == 2 ? a2 : a3));
      }
    };

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
    auto pack2 = [](float a, float c) {
      __nv_bfloat162 v = __floats2bfloat162_rn(a, c);
      return *reinterpret_cast<uint32_t*>(&v);
    };
    Stash cur = load_stash(p.t_end - 1);

    for (int s = 0; s < nsteps; ++s) {
      const int t = p.t_end - 1 - s;
      const size_t row = (size_t)t * p.Bpad + b;
      const size_t off = row * p.H + j;
      float dmask[4] = {1.0f, 1.0f, 1.0f, 1.0f};
      if (p.keep < 1.0f) {
        const uint4 rnd = rec_dropout_bits(p.seed, p.rng_offset, off >> 2);
        const uint32_t rr4[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) dmask[i] = (u32_to_unit(rr4[i]) < p.keep) ? inv_keep : 0.0f;
      }
      float P[4] = {0.f, 0.f, 0.f, 0.f};
      if (s > 0) {
        if (e == 0) REC_TRACE(s, 4);
        reduce_partials(p.gen_base + s - 1, P);
        if (e == 0) { REC_TRACE(s, 5); REC_SKEW(s, 4); }
      }
      float r[4], z[4], n[4], hn[4], hp[4];
      unpack(cur.r, r); unpack(cur.z, z); unpack(cur.n, n); unpack(cur.hn, hn); unpack(cur.hp, hp);
      const float dy[4] = {cur.dy.x * dmask[0], cur.dy.y * dmask[1], cur.dy.z * dmask[2], cur.dy.w * dmask[3]};
      float dh[4];
      if (s > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dh[i] = carry[i] + P[i] + dy[i];
      } else if (!p.first_chunk) {
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
      // publish dGh_t (the data is the signal: relaxed gpu-scope stores of whole words, polled by the loaders of the group)
      const size_t goff = row * 3 * p.H + j;
      st_relaxed_v2(p.dGh + goff, pack2(gr[0], gr[1]), pack2(gr[2], gr[3]));
      st_relaxed_v2(p.dGh + goff + p.H, pack2(gz[0], gz[1]), pack2(gz[2], gz[3]));
      st_relaxed_v2(p.dGh + goff + 2 * p.H, pack2(gnh[0], gnh[1]), pack2(gnh[2], gnh[3]));
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_s);
      if (e == 0) { REC_TRACE(s, 6); REC_SKEW(s, 0); }
      if (s + 1 < nsteps) cur = load_stash(t - 1);       // in flight during the MMA
      st_bf16x4(p.dGx + goff, gr[0], gr[1], gr[2], gr[3]);
      st_bf16x4(p.dGx + goff + p.H, gz[0], gz[1], gz[2], gz[3]);
      st_bf16x4(p.dGx + goff + 2 * p.H, gn[0], gn[1], gn[2], gn[3]);

      // accumulator -> four partial blocks (one per destination kq = TMEM lane quarter)
      mbar_wait(bar_d, (uint32_t)s & 1u);
      if (e == 0) { REC_TRACE(s, 7); REC_SKEW(s, 3); }
      tc_fence_after();
      {
        const int gen = p.gen_base + s;
        const uint32_t tag = (uint32_t)(gen >> 1) & 3u;
        uint32_t v[16];
        tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + chalf * 16, v);
        tmem_ld_wait();
        float* dst = block_of(gen, q, kq) + ((chalf * 4) * 32 + lane) * 4;
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4)
          st_relaxed_v4(dst + g4 * 128, (v[4 * g4] & ~3u) | tag, (v[4 * g4 + 1] & ~3u) | tag, (v[4 * g4 + 2] & ~3u) | tag, (v[4 * g4 + 3] & ~3u) | tag);
      }
      tc_fence_before();
    }
    // recurrent gradient for the step before this chunk
    {
      float P[4];
      reduce_partials(p.gen_base + nsteps - 1, P);
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<64>(*tmem_slot);
  }
}

}  // namespace b2t
