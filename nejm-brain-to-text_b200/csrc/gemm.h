// Plan/launch interface of the tcgen05 GEMM (see gemm.cuh for the kernel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_params.h"

namespace b2t {

// Description of one (possibly batched) GEMM  C[z][m,n] = sum_k A[z][m,k] * B[zb][n,k].
struct GemmSpec {
  int a_mn = 0, b_mn = 0;     // 0: K-major (contraction contiguous), 1: MN-major (contraction = stored row index)
  int epi = 0;                // EPI_STORE / EPI_DAY / EPI_ATOMIC
  int out_bf16 = 0;
  long long M = 0, N = 0, K = 0;
  const void* A = nullptr;
  const void* B = nullptr;
  long long lda = 0, ldb = 0; // row strides (elements) of the stored arrays (plain case)
  // K-major A given as a strided "patch" view: rows are (rin, rout) pairs with these strides
  int a_rin = 0; long long a_rout = 0;
  long long a_rin_stride = 0, a_rout_stride = 0;
  // MN-major operands: the contraction rows are (k_rin, k_rout) pairs (0 => plain rows with lda/ldb)
  int k_rin = 0; long long k_rout = 0;
  long long b_rin_stride = 0, b_rout_stride = 0;
  // batching
  int nz = 1, nzb = 0;
  long long a_zstride = 0, b_zstride = 0;
  const int* z_map = nullptr;
  int zmap_b = 0;
  // output / epilogue
  void* C = nullptr;
  long long ldc = 0, c_zstride = 0;
  const float* bias = nullptr;
  long long bias_zstride = 0;
  float keep = 1.0f;
  unsigned long long seed = 0, rng_offset = 0;
  // gating (see gemm_params.h); max_ctas > 0 caps the grid (a gated GEMM runs beside a persistent recurrence kernel on spare SMs)
  const int* gate = nullptr;
  int gate_need = 0, gate_rows_per_step = 0, gate_steps = 0;
  int* done = nullptr;
  int tm_reverse = 0, max_ctas = 0;
  int bn = 0;                 // force the tile width (128 / 256); 0 = chosen by the planner
};

struct GemmPlan {
  CUtensorMap ta, tb;
  GemmParams p;
  int a_mn, b_mn, epi, out_bf16;
  int max_ctas, bn;
};

int gemm_plan_build(GemmPlan* pl, const GemmSpec& s);
cudaError_t gemm_run(const GemmPlan& pl, cudaStream_t st);
cudaError_t gemm_preload();   // load every kernel variant now (see gemm.cu)
int num_sms();

}  // namespace b2t
