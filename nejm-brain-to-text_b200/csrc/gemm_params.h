// Kernel-side parameter block of the tcgen05 GEMM (shared by host planner and device code).
#pragma once
namespace b2t {

enum { EPI_STORE = 0, EPI_DAY = 1, EPI_ATOMIC = 2, EPI_ACCUM = 3 };   // ACCUM: C += A*B (each tile owned by one CTA)

struct GemmParams {
  int M, N;                 // output extent per problem (M only used for MN-major A)
  int k_iters;              // 64-wide contraction chunks
  int tiles_m, tiles_n, nz;
  // K-major A: rows of a tile are (rin, rout) pairs, r = rout_l * a_bin + rin_l; output row = rout * a_rin + rin
  int a_bin, a_rin_blocks, a_rin, a_rout;
  // MN-major operands: contraction chunk kc covers rin block kc % k_rin_blocks, rout block kc / k_rin_blocks
  int k_bin, k_rin_blocks;
  void* C;
  long long ldc, c_zstride;
  const float* bias;        // per output column (nullable)
  long long bias_zstride;
  const int* z_map;         // nullable: problem z uses bias / (atomic) output index z_map[z]
  int zmap_b;               // ... and also the B-operand problem index
  float keep;               // EPI_DAY dropout keep probability (1 => no dropout)
  unsigned long long seed, rng_offset;
  // Gating against a concurrently running producer / consumer (the whole-stack recurrence kernels, gru_stack.cuh).  Rows of
  // a K-major A operand are time-major (row = t * gate_rows_per_step + trial).  Before loading M-tile tm the TMA producer
  // waits until gate[t] >= gate_need for the LAST time step the tile touches (first one when tm_reverse: the backward
  // recurrence walks time downwards); after an M-tile's output is stored every epilogue warp adds 1 to done[tm].
  const int* gate;          // nullable: per-time-step progress counters of the producer
  int gate_need;
  int gate_rows_per_step;   // Bpad
  int gate_steps;           // T'
  int* done;                // nullable: per-M-tile completion counters (consumer waits for 4 * tiles_n)
  int tm_reverse;           // walk the M-tiles from the last to the first
};

}  // namespace b2t
