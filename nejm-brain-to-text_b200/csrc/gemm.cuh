// Persistent warp-specialised bf16 GEMM for sm_100a:  C[m,n] = sum_k A[m,k] * B[n,k]  (fp32 accumulate)
//
//   warp 0      : TMA producer (one elected lane)   global -> 6-stage smem ring (SWIZZLE_128B)
//   warp 1      : tcgen05.mma issuer (one lane); also owns the TMEM allocation (2 x 128 fp32 columns)
//   warps 2..5  : epilogue; warp w reads TMEM lanes 32*(w%4).. with tcgen05.ld and writes C
//
// Operands are described by 4-D tensor maps so that the same kernel serves
//   * plain row-major matrices (K-major operand: rows x K, K contiguous),
//   * the overlapping sliding-window "patch" view of the day-layer output (rows (t',b) of the
//     unfolded [B*T', 14*512] matrix are addressed with strides, never materialised; reference
//     rnn_model.py:106-119),
//   * transposed use of a row-major matrix (MN-major operand: the contraction index is the row index
//     of the stored array) for the weight-gradient and data-gradient GEMMs,
//   * batched problems with a per-problem index map (the 45 day-specific 512x512 layers,
//     rnn_model.py:95-98).
//
// Tile 128 x 128 x 64; one tcgen05.mma is 128 x 128 x 16.
#pragma once
#include "sm100.cuh"
#include "gemm_params.h"

namespace b2t {


constexpr int GEMM_BM = 128, GEMM_BK = 64;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;
// Tile width BN = 128 (6-stage ring) or 256 (4-stage ring).  The main loop is bound by what one SM can pull out of L2
// (~65 B/clk measured): a 128 x 256 tile needs 48 KB per 64-deep slice for twice the MMA work of a 128 x 128 tile's 32 KB.
template <int BN> struct GemmTile {
  static constexpr int kStages = BN == 256 ? 4 : 6;
  static constexpr int kBBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = GEMM_A_BYTES + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * 32 * 36 * 4 /*epilogue staging*/ + 4 * BN * 4 /*bias*/;
};
constexpr int GEMM_EPI_PITCH = 36;      // floats per row of an epilogue warp's 32 x 32 staging tile (16-byte aligned rows, conflict-free float4 access)
constexpr int GEMM_THREADS = 192;


template <typename OutT> __device__ __forceinline__ void store_vec4(OutT* dst, const float (&w)[4]);
template <> __device__ __forceinline__ void store_vec4<float>(float* dst, const float (&w)[4]) {
  *reinterpret_cast<float4*>(dst) = make_float4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ void store_vec4<__nv_bfloat16>(__nv_bfloat16* dst, const float (&w)[4]) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(w[0], w[1]), hi = __floats2bfloat162_rn(w[2], w[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(dst) = u;
}
template <typename OutT> __device__ __forceinline__ void store_one(OutT* dst, float w);
template <> __device__ __forceinline__ void store_one<float>(float* dst, float w) { *dst = w; }
template <> __device__ __forceinline__ void store_one<__nv_bfloat16>(__nv_bfloat16* dst, float w) { *dst = __float2bfloat16_rn(w); }

template <bool A_MN, bool B_MN, int EPI, typename OutT, int GEMM_BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
  constexpr int GEMM_STAGES = GemmTile<GEMM_BN>::kStages, GEMM_STAGE_BYTES = GemmTile<GEMM_BN>::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* tfull_bar = empty_bar + GEMM_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_m * p.tiles_n * p.nz;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * GEMM_BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tn = tile % p.tiles_n;
        const int tm0 = (tile / p.tiles_n) % p.tiles_m;
        const int tm = p.tm_reverse ? p.tiles_m - 1 - tm0 : tm0;
        const int z = tile / (p.tiles_n * p.tiles_m);
        const int zb = (p.z_map && p.zmap_b) ? p.z_map[z] : z;
        if (p.gate) {
          // rows of this tile exist once the producer has passed the tile's last (first, when walking backwards) time step
          int ts = p.tm_reverse ? (tm * GEMM_BM) / p.gate_rows_per_step : (tm * GEMM_BM + GEMM_BM - 1) / p.gate_rows_per_step;
          if (ts > p.gate_steps - 1) ts = p.gate_steps - 1;
          uint32_t spins = 0;
          while (ld_acquire_gpu(p.gate + ts) < p.gate_need) {
            __nanosleep(64);
            if (++spins > (1u << 26)) __trap();
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes of the producer -> TMA (async proxy) reads
        }
        for (int kc = 0; kc < p.k_iters; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * GEMM_STAGE_BYTES;
          uint8_t* sb = sa + GEMM_A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], GEMM_STAGE_BYTES);
          const int kb = kc % p.k_rin_blocks, ko = kc / p.k_rin_blocks;
          if (A_MN) {
            tma_load_4d(sa, &tmap_a, &full_bar[stage], tm * GEMM_BM, kb * p.k_bin, ko * (64 / p.k_bin), z);
            tma_load_4d(sa + 8192, &tmap_a, &full_bar[stage], tm * GEMM_BM + 64, kb * p.k_bin, ko * (64 / p.k_bin), z);
          } else {
            const int rb = tm % p.a_rin_blocks, ro = tm / p.a_rin_blocks;
            tma_load_4d(sa, &tmap_a, &full_bar[stage], kc * GEMM_BK, rb * p.a_bin, ro * (GEMM_BM / p.a_bin), z);
          }
          if (B_MN) {
#pragma unroll
            for (int nb = 0; nb < GEMM_BN / 64; ++nb)      // 64-column blocks, 8 KB apart (the MN-major descriptor's leading-dimension stride)
              tma_load_4d(sb + nb * 8192, &tmap_b, &full_bar[stage], tn * GEMM_BN + nb * 64, kb * p.k_bin, ko * (64 / p.k_bin), zb);
          } else {
            tma_load_4d(sb, &tmap_b, &full_bar[stage], kc * GEMM_BK, 0, tn * GEMM_BN, zb);
          }
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, GEMM_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
        for (int kc = 0; kc < p.k_iters; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * GEMM_STAGE_BYTES);
          const uint32_t sb = sa + GEMM_A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc(sa + k * 2048, 8192, 1024) : umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc(sb + k * 2048, 8192, 1024) : umma_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kc | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs have read it
          if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);        // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    // Phase 1 (thread = accumulator row, tcgen05.ld gives 32 consecutive columns of it): bias / activation / dropout, then the
    // 32 x 32 block of the warp goes through a padded shared-memory tile.  Phase 2 (8 lanes = 128 contiguous bytes of one output
    // row, 4 rows per instruction): coalesced stores.  Writing rows straight from phase 1 touches 32 different lines per store
    // instruction (one 16-byte piece each) and made K = 768 GEMMs epilogue-bound: 7.6 us per 128 x 128 tile against 1.6 us of MMAs.
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;             // row inside the 128-row tile (phase 1)
    float* stg = reinterpret_cast<float*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES + 256) + q * (32 * GEMM_EPI_PITCH);
    float* bias_s = reinterpret_cast<float*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES + 256) + 4 * 32 * GEMM_EPI_PITCH + q * GEMM_BN;
    const int sub_row = lane >> 3, cg = (lane & 7) * 4;   // phase 2: row within a group of four, first of this lane's four columns
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tn = tile % p.tiles_n;
      const int tm0 = (tile / p.tiles_n) % p.tiles_m;
      const int tm = p.tm_reverse ? p.tiles_m - 1 - tm0 : tm0;
      const int z = tile / (p.tiles_n * p.tiles_m);
      const int zb = p.z_map ? p.z_map[z] : z;
      auto out_row = [&](int rr, bool& ok) -> long long {   // output row (and validity) of tile row rr
        if (A_MN) {
          const long long m = (long long)tm * GEMM_BM + rr;
          ok = m < p.M;
          return m;
        }
        const int rb = tm % p.a_rin_blocks, ro = tm / p.a_rin_blocks;
        const int rin = rb * p.a_bin + (rr % p.a_bin);
        const int rout = ro * (GEMM_BM / p.a_bin) + (rr / p.a_bin);
        ok = rout < p.a_rout;
        return (long long)rout * p.a_rin + rin;
      };
      bool row_ok;
      const long long m = out_row(r, row_ok);
      const int zc = (EPI == EPI_ATOMIC) ? zb : z;
      OutT* cbase = reinterpret_cast<OutT*>(p.C) + (long long)zc * p.c_zstride;
      const float* bias = p.bias ? p.bias + (long long)zb * p.bias_zstride : nullptr;
      const bool vec_ok = (p.ldc % (16 / (int)sizeof(OutT))) == 0 && (reinterpret_cast<uintptr_t>(cbase) & 15) == 0;
      // phase-2 rows of this lane: tile row q*32 + 4*i + sub_row, i = 0..7
      long long m2[8];
      uint32_t ok2 = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        bool ok;
        m2[i] = out_row(q * 32 + 4 * i + sub_row, ok);
        ok2 |= (ok ? 1u : 0u) << i;
      }

      // this tile's bias row -> the warp's shared-memory slice while the MMAs still run (a per-chunk global load sat on the
      // epilogue's critical path: ~600 cycles of L2 latency for each of the 4-8 chunks of a tile)
      if (bias) {
#pragma unroll
        for (int i = 0; i < GEMM_BN / 32; ++i) {
          const int n = tn * GEMM_BN + i * 32 + lane;
          bias_s[i * 32 + lane] = n < p.N ? __ldg(bias + n) : 0.0f;
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      __syncwarp();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * GEMM_BN;
      // one 32-column chunk of the accumulator row block: bias, epilogue function, staging tile, coalesced stores
      auto process_chunk = [&](const int c0, uint32_t (&v)[32]) {
        const int n0 = tn * GEMM_BN + c0;
        const int ncols = p.N - n0;          // columns of this 32-chunk that exist
        if (ncols <= 0) return;              // (warp-uniform)
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        if (bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {       // broadcast reads of the prefetched row
            const float4 bv = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * i);
            f[4 * i] += bv.x; f[4 * i + 1] += bv.y; f[4 * i + 2] += bv.z; f[4 * i + 3] += bv.w;
          }
        }
        if (EPI == EPI_DAY && row_ok) {      // (rows beyond the trial's length are never stored: skip their Philox draws)
          // softsign (rnn_model.py:99) then inverted dropout (rnn_model.py:102-103)
          const float inv_keep = 1.0f / p.keep;
          const unsigned long long e0 = (unsigned long long)zc * p.c_zstride + (unsigned long long)m * p.ldc + n0;
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            uint4 rnd = make_uint4(0, 0, 0, 0);
            if (p.keep < 1.0f) {
              const unsigned long long c = (e0 >> 2) + i4 + p.rng_offset;
              rnd = philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x0da1u, 0), make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
            }
            const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float y = f[4 * i4 + j];
              y = __fdividef(y, 1.0f + fabsf(y));   // (2 ulp; the result is rounded to bf16)
              if (p.keep < 1.0f) y = (u32_to_unit(rr[j]) < p.keep) ? y * inv_keep : 0.0f;
              f[4 * i4 + j] = y;
            }
          }
        }
        // ---- through the warp's staging tile
        __syncwarp();                        // the previous chunk's phase-2 reads are done
        {
          float4* dst = reinterpret_cast<float4*>(stg + lane * GEMM_EPI_PITCH);
#pragma unroll
          for (int i = 0; i < 8; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        }
        __syncwarp();
        const int nc4 = ncols - cg;          // columns left from this lane's first column
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!((ok2 >> i) & 1u) || nc4 <= 0) continue;
          const float4 o = *reinterpret_cast<const float4*>(stg + (4 * i + sub_row) * GEMM_EPI_PITCH + cg);
          float w[4] = {o.x, o.y, o.z, o.w};
          OutT* dst = cbase + m2[i] * p.ldc + n0 + cg;
          if (EPI == EPI_ATOMIC) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nc4) atomicAdd(reinterpret_cast<float*>(dst) + k, w[k]);
          } else if (nc4 >= 4 && vec_ok) {
            if (EPI == EPI_ACCUM) {
              const float4 old = *reinterpret_cast<const float4*>(dst);
              w[0] += old.x; w[1] += old.y; w[2] += old.z; w[3] += old.w;
            }
            store_vec4<OutT>(dst, w);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nc4) {
                if (EPI == EPI_ACCUM) w[k] += reinterpret_cast<const float*>(dst)[k];
                store_one<OutT>(dst + k, w[k]);
              }
          }
        }
      };
      if constexpr (EPI == EPI_DAY) {
        // The softsign + Philox epilogue is ~1 400 instructions per chunk: unrolled over the chunks (and ping-pong buffered) the
        // kernel grew to 150 KB of code and the epilogue warps stalled on instruction fetch (ncu: 30 % "no instruction").  One
        // rolled loop over the chunks keeps the body resident.
        uint32_t v[32];
#pragma unroll 1
        for (int c0 = 0; c0 < GEMM_BN; c0 += 32) {
          tmem_ld32(tacc + c0, v);
          tmem_ld_wait32(v);
          if (c0 + 32 == GEMM_BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          process_chunk(c0, v);
        }
      } else {
        uint32_t va[32], vb[32];
        tmem_ld32(tacc, va);
#pragma unroll
        for (int ci = 0; ci < GEMM_BN / 32; ++ci) {
          const int c0 = ci * 32;
          uint32_t (&v)[32] = (ci & 1) ? vb : va;
          tmem_ld_wait32(v);
          if (ci + 1 < GEMM_BN / 32) tmem_ld32(tacc + c0 + 32, (ci & 1) ? va : vb);   // next chunk in flight during this one's stores
          if (c0 + 32 == GEMM_BN) {            // accumulator fully read: hand it back to the MMA issuer before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          }
          process_chunk(c0, v);
        }
      }
      if (p.done) {
        __threadfence();                     // this lane's stores of the tile are visible at gpu scope before the count below
        __syncwarp();
        if (lane == 0) red_release_add(p.done + tm, 1);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * GEMM_BN>(tmem_base);
  }
}

}  // namespace b2t
