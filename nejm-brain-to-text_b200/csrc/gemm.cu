// Host side of the tcgen05 GEMM: plan construction (tensor maps + tile decomposition) and launch.
#include "gemm.h"
#include "gemm.cuh"
#include "tmap.h"

namespace b2t {

static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <bool A_MN, bool B_MN, int EPI, typename OutT, int BN>
static cudaError_t launch_bn(const GemmPlan& pl, cudaStream_t st) {
  auto kern = gemm_bf16_kernel<A_MN, B_MN, EPI, OutT, BN>;
  constexpr int GEMM_SMEM_BYTES = GemmTile<BN>::kSmemBytes;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int total = pl.p.tiles_m * pl.p.tiles_n * pl.p.nz;
  if (total <= 0) return cudaSuccess;
  int grid = total < num_sms() ? total : num_sms();
  if (pl.max_ctas > 0 && grid > pl.max_ctas) grid = pl.max_ctas;
  kern<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(pl.ta, pl.tb, pl.p);
  return cudaGetLastError();
}
template <bool A_MN, bool B_MN, int EPI, typename OutT>
static cudaError_t launch_variant(const GemmPlan& pl, cudaStream_t st) {
  return pl.bn == 256 ? launch_bn<A_MN, B_MN, EPI, OutT, 256>(pl, st) : launch_bn<A_MN, B_MN, EPI, OutT, 128>(pl, st);
}

// With lazy module loading (the CUDA 12 default) the first launch of a kernel loads it, which can need a device-wide
// synchronisation: a gated GEMM launched for the first time while the persistent recurrence kernel is already waiting for its
// output would then dead-lock.  Touch every variant once, up front (cudaFuncSetAttribute loads the function).
template <bool A_MN, bool B_MN, int EPI, typename OutT>
static cudaError_t preload_variant() {
  cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<A_MN, B_MN, EPI, OutT, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmTile<128>::kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_bf16_kernel<A_MN, B_MN, EPI, OutT, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmTile<256>::kSmemBytes);
}
cudaError_t gemm_preload() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e;
  if ((e = preload_variant<false, false, EPI_STORE, float>()) != cudaSuccess) return e;
  if ((e = preload_variant<false, false, EPI_STORE, __nv_bfloat16>()) != cudaSuccess) return e;
  if ((e = preload_variant<false, true, EPI_STORE, float>()) != cudaSuccess) return e;
  if ((e = preload_variant<false, true, EPI_STORE, __nv_bfloat16>()) != cudaSuccess) return e;
  if ((e = preload_variant<false, true, EPI_DAY, __nv_bfloat16>()) != cudaSuccess) return e;
  if ((e = preload_variant<true, true, EPI_STORE, float>()) != cudaSuccess) return e;
  if ((e = preload_variant<true, true, EPI_ATOMIC, float>()) != cudaSuccess) return e;
  if ((e = preload_variant<true, true, EPI_ACCUM, float>()) != cudaSuccess) return e;
  done = true;
  return cudaSuccess;
}

cudaError_t gemm_run(const GemmPlan& pl, cudaStream_t st) {
  const int key = (pl.a_mn ? 1 : 0) | (pl.b_mn ? 2 : 0) | (pl.epi << 2) | (pl.out_bf16 ? 16 : 0);
  switch (key) {
    case 0 | (EPI_STORE << 2): return launch_variant<false, false, EPI_STORE, float>(pl, st);
    case 0 | (EPI_STORE << 2) | 16: return launch_variant<false, false, EPI_STORE, __nv_bfloat16>(pl, st);
    case 2 | (EPI_STORE << 2): return launch_variant<false, true, EPI_STORE, float>(pl, st);
    case 2 | (EPI_STORE << 2) | 16: return launch_variant<false, true, EPI_STORE, __nv_bfloat16>(pl, st);
    case 2 | (EPI_DAY << 2) | 16: return launch_variant<false, true, EPI_DAY, __nv_bfloat16>(pl, st);
    case 3 | (EPI_STORE << 2): return launch_variant<true, true, EPI_STORE, float>(pl, st);
    case 3 | (EPI_ATOMIC << 2): return launch_variant<true, true, EPI_ATOMIC, float>(pl, st);
    case 3 | (EPI_ACCUM << 2): return launch_variant<true, true, EPI_ACCUM, float>(pl, st);
    default: return cudaErrorInvalidValue;
  }
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// largest power of two <= cap that divides n
static int pow2_divisor(int n, int cap) {
  int b = 1;
  while (b * 2 <= cap && n % (b * 2) == 0) b *= 2;
  return b;
}

int gemm_plan_build(GemmPlan* pl, const GemmSpec& s) {
  memset(pl, 0, sizeof(*pl));
  pl->a_mn = s.a_mn; pl->b_mn = s.b_mn; pl->epi = s.epi; pl->out_bf16 = s.out_bf16;
  GemmParams& p = pl->p;
  p.N = s.N;
  p.nz = s.nz > 0 ? s.nz : 1;
  p.C = s.C; p.ldc = s.ldc; p.c_zstride = s.c_zstride;
  p.bias = s.bias; p.bias_zstride = s.bias_zstride; p.z_map = s.z_map; p.zmap_b = s.zmap_b;
  p.keep = s.keep > 0 ? s.keep : 1.0f; p.seed = s.seed; p.rng_offset = s.rng_offset;
  p.gate = s.gate; p.gate_need = s.gate_need; p.gate_rows_per_step = s.gate_rows_per_step > 0 ? s.gate_rows_per_step : 1;
  p.gate_steps = s.gate_steps; p.done = s.done; p.tm_reverse = s.tm_reverse;
  pl->max_ctas = s.max_ctas;
  if ((s.gate || s.done) && (s.a_mn || s.a_rin > 0 || p.nz != 1)) return -20;   // gating is defined for plain K-major A (time-major rows) only
  // tile width: 256 when the columns divide evenly and (after the M decomposition below) enough tiles remain to fill the chip
  int bn = 128;
  {
    const int force = getenv("B2T_GEMM_BN") ? atoi(getenv("B2T_GEMM_BN")) : 0;   // bring-up / test override
    const long long rows = s.a_mn ? s.M : (s.a_rin > 0 ? (long long)s.a_rin * s.a_rout : s.M);
    const long long tiles256 = (long long)ceil_div(rows, GEMM_BM) * ceil_div(s.N, 256) * p.nz;
    if (s.N % 256 == 0 && s.epi != EPI_ATOMIC && s.epi != EPI_DAY && (tiles256 >= num_sms() || s.gate || s.done)) bn = 256;   // (the day layer's Philox epilogue is slower when unrolled over 8 chunks: measured 258 vs 105 us)
    if (s.bn == 128 || (s.bn == 256 && s.N % 256 == 0)) bn = s.bn;
    if (force == 128 || (force == 256 && s.N % 256 == 0)) bn = force;
  }
  pl->bn = bn;
  p.tiles_n = ceil_div(s.N, bn);
  p.k_bin = 1; p.k_rin_blocks = 1; p.a_bin = 1; p.a_rin_blocks = 1; p.a_rin = 1; p.a_rout = s.M;

  // ---- contraction decomposition (MN-major operands index the contraction by stored rows)
  int k_rin = s.k_rin > 0 ? s.k_rin : 1;           // inner contraction-row dimension (e.g. batch)
  long long k_rout = s.k_rin > 0 ? s.k_rout : s.K; // outer contraction-row dimension (e.g. time)
  if (s.a_mn || s.b_mn) {
    p.k_bin = pow2_divisor(k_rin, 64);
    p.k_rin_blocks = k_rin / p.k_bin;
    p.k_iters = p.k_rin_blocks * ceil_div(k_rout, 64 / p.k_bin);
  }
  if (!s.a_mn || !s.b_mn) {
    const int ki = ceil_div(s.K, GEMM_BK);
    if ((s.a_mn || s.b_mn) && ki != p.k_iters) {
      // mixed majors: the K-major operand advances 64 contraction elements per chunk, so the
      // MN-major one must be a plain matrix (k_rin == 1).
      if (k_rin != 1) return -10;
    }
    p.k_iters = (s.a_mn || s.b_mn) ? p.k_iters : ki;
  }

  // ---- A operand
  uint64_t dims[4], str[3];
  uint32_t box[4];
  if (!s.a_mn) {
    // rows = (rin, rout) pairs; a plain matrix has rin == 1
    const bool patch = s.a_rin > 0;
    const int rin = patch ? s.a_rin : 1;
    const long long rout = patch ? s.a_rout : s.M;
    p.a_bin = pow2_divisor(rin, GEMM_BM);
    p.a_rin_blocks = rin / p.a_bin;
    p.a_rin = rin; p.a_rout = (int)rout;
    p.tiles_m = p.a_rin_blocks * ceil_div(rout, GEMM_BM / p.a_bin);
    p.M = (int)(rout * rin);
    dims[0] = s.K; dims[1] = rin; dims[2] = rout; dims[3] = p.nz;
    str[0] = patch ? s.a_rin_stride : s.lda;
    str[1] = patch ? s.a_rout_stride : s.lda;
    str[2] = s.a_zstride > 0 ? s.a_zstride : 8;
    box[0] = 64; box[1] = p.a_bin; box[2] = GEMM_BM / p.a_bin; box[3] = 1;
  } else {
    p.M = s.M;
    p.tiles_m = ceil_div(s.M, GEMM_BM);
    dims[0] = s.M; dims[1] = k_rin; dims[2] = k_rout; dims[3] = p.nz;
    str[0] = s.k_rin > 0 ? s.a_rin_stride : s.lda;
    str[1] = s.k_rin > 0 ? s.a_rout_stride : s.lda;
    str[2] = s.a_zstride > 0 ? s.a_zstride : 8;
    box[0] = 64; box[1] = p.k_bin; box[2] = 64 / p.k_bin; box[3] = 1;
  }
  if (make_tmap_bf16_4d(&pl->ta, s.A, dims, str, box)) return -1;

  // ---- B operand
  const int nzb = s.nzb > 0 ? s.nzb : p.nz;
  if (!s.b_mn) {
    dims[0] = s.K; dims[1] = 1; dims[2] = s.N; dims[3] = nzb;
    str[0] = s.ldb; str[1] = s.ldb; str[2] = s.b_zstride > 0 ? s.b_zstride : 8;
    box[0] = 64; box[1] = 1; box[2] = bn; box[3] = 1;
  } else {
    dims[0] = s.N; dims[1] = k_rin; dims[2] = k_rout; dims[3] = nzb;
    str[0] = s.k_rin > 0 ? s.b_rin_stride : s.ldb;
    str[1] = s.k_rin > 0 ? s.b_rout_stride : s.ldb;
    str[2] = s.b_zstride > 0 ? s.b_zstride : 8;
    box[0] = 64; box[1] = p.k_bin; box[2] = 64 / p.k_bin; box[3] = 1;
  }
  if (make_tmap_bf16_4d(&pl->tb, s.B, dims, str, box)) return -2;
  return 0;
}

}  // namespace b2t
