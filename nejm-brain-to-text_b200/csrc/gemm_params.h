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
};

}  // namespace b2t
