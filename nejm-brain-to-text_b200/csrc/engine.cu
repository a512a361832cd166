// Engine: owns the buffer carving, tensor-map plans and kernel sequencing of the GRU -> CTC
// training / inference step, and exports it through the C ABI in include/b2t_b200.h.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/b2t_b200.h"
#include "ctc.cuh"
#include "elementwise.cuh"
#include "gemm.h"
#include "gru_rec.cuh"
#include "gru_rec_bwd2.cuh"
#include "gru_stack.cuh"
#include "optim.cuh"
#include "tmap.h"

using namespace b2t;

// ------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static long long g_launches = 0;
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess) return fail(B2T_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)
#define LAUNCHED() (++g_launches, cudaGetLastError())

extern "C" const char* b2t_last_error(void) { return g_err; }
extern "C" int b2t_version(void) { return 100; }
extern "C" long long b2t_launch_count(void) { return g_launches; }

// ------------------------------------------------------------------------------------ layout
static inline long long r64(long long x) { return (x + 63) / 64 * 64; }
static inline int r16(int x) { return (x + 15) / 16 * 16; }

struct SegInfo {
  std::string name;
  long long offset, rows, cols;
  int group, day;
};

static std::vector<SegInfo> build_layout(const b2t_config& c, long long* total) {
  std::vector<SegInfo> v;
  long long off = 0;
  auto add = [&](const std::string& n, long long r, long long cc, int group, int day) {
    v.push_back({n, off, r, cc, group, day});
    off += r64(r * cc);
  };
  const int D = c.neural_dim, H = c.n_units;
  const int K0 = D * (c.patch_size > 0 ? c.patch_size : 1);
  for (int d = 0; d < c.n_days; ++d) add("day_weights." + std::to_string(d), D, D, 1, d);
  for (int d = 0; d < c.n_days; ++d) add("day_biases." + std::to_string(d), 1, D, 1, d);
  for (int l = 0; l < c.n_layers; ++l) {
    const std::string s = std::to_string(l);
    add("gru.weight_ih_l" + s, 3 * H, l == 0 ? K0 : H, 2, -1);
    add("gru.weight_hh_l" + s, 3 * H, H, 2, -1);
    add("gru.bias_ih_l" + s, 1, 3 * H, 0, -1);
    add("gru.bias_hh_l" + s, 1, 3 * H, 0, -1);
  }
  add("out.weight", c.n_classes, H, 2, -1);
  add("out.bias", 1, c.n_classes, 0, -1);
  add("h0", 1, H, 2, -1);
  if (total) *total = off;
  return v;
}

static int check_cfg(const b2t_config* c) {
  if (!c) return fail(B2T_ERR_ARG, "null config");
  if (c->n_units % 64 != 0 || c->n_units < 64 || c->n_units > 768)
    return fail(B2T_ERR_UNSUPPORTED, "n_units=%d: must be a multiple of 64 in [64,768] (W_hh slice must fit in shared memory)", c->n_units);
  if (c->neural_dim % 8 != 0 || c->neural_dim < 8 || c->neural_dim > 1024) return fail(B2T_ERR_UNSUPPORTED, "neural_dim=%d: must be a multiple of 8, <= 1024", c->neural_dim);
  if (c->n_classes < 2 || c->n_classes > 64) return fail(B2T_ERR_UNSUPPORTED, "n_classes=%d: must be in [2,64]", c->n_classes);
  if (c->n_layers < 1 || c->n_layers > 16 || c->n_days < 1) return fail(B2T_ERR_ARG, "bad n_layers/n_days");
  if (c->patch_size < 0 || (c->patch_size > 0 && c->patch_stride < 1)) return fail(B2T_ERR_ARG, "bad patch config");
  return 0;
}

extern "C" int b2t_param_segments(const b2t_config* cfg) {
  if (check_cfg(cfg)) return B2T_ERR_ARG;
  return (int)build_layout(*cfg, nullptr).size();
}
extern "C" int b2t_param_segment(const b2t_config* cfg, int index, char* name, int name_cap, long long* offset, long long* rows, long long* cols) {
  if (check_cfg(cfg)) return B2T_ERR_ARG;
  auto v = build_layout(*cfg, nullptr);
  if (index < 0 || index >= (int)v.size()) return fail(B2T_ERR_ARG, "segment index out of range");
  if (name && name_cap > 0) snprintf(name, name_cap, "%s", v[index].name.c_str());
  if (offset) *offset = v[index].offset;
  if (rows) *rows = v[index].rows;
  if (cols) *cols = v[index].cols;
  return 0;
}
extern "C" long long b2t_param_elems(const b2t_config* cfg) {
  if (check_cfg(cfg)) return B2T_ERR_ARG;
  long long t = 0;
  build_layout(*cfg, &t);
  return t;
}
extern "C" long long b2t_grad_elems(const b2t_config* cfg) {
  const long long t = b2t_param_elems(cfg);
  return t < 0 ? t : t + r64(cfg->n_days);
}

// ------------------------------------------------------------------------------------ engine
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

struct Carver {
  uint8_t* base;
  size_t off, cap;
  bool dry;
  template <typename T>
  T* take(size_t n) {
    off = (off + 1023) & ~size_t(1023);
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

constexpr int LDL = 64;        // pitch (elements) of logits / dlogits rows: 128 B in bf16, TMA friendly
constexpr int MAX_CHUNKS = 8;  // time chunks of the layer wave-front
constexpr size_t REC_SMEM_BYTES = 120 * 1024;   // recurrence kernels request at least this much shared memory: one CTA per SM (they own the TMEM)
constexpr int MAX_LANES = 5;   // concurrent recurrence launches (side streams)

struct LayerBuf {
  __nv_bfloat16 *hseq, *hdrop, *R, *Z, *Nn, *HN, *dGx, *dGh;
  unsigned char* kmask;         // keep bits of hdrop, one byte per 4 units (training)
  float *gx, *dY, *part, *h_state, *dh_state;
  int gen = 0;                                       // backward partial-exchange generation (mod 8), advanced per launch
};

struct b2t_engine {
  b2t_config cfg;
  int maxB, maxT, maxS, training;
  int D, H, L, C, K0, patch, stride;
  std::vector<SegInfo> segs;
  long long n_params;
  float *params, *grads, *m1, *m2;
  // carved buffers
  __nv_bfloat16* shadow;
  __nv_bfloat16 *xs, *xd, *xu, *dxu, *dpre, *dlog16;
  float *logits, *dlog32, *alpha, *beta;
  std::vector<LayerBuf> lay;
  int *steps, *greedy_scratch;
  int part_geom = -1;
  int poll_delay = 0, poll_delay_b = 0;
  float *sumsq, *stats;
  Segment* d_segs;
  ChunkRef* d_chunks;
  int n_chunks;
  float* touched;             // tail of the gradient buffer: [n_days]
  // current shape + plans
  int B = 0, Bpad = 0, T_in = 0, T_out = 0, Tp = 0, M = 0;
  int BG = 16, BGb = 16, NSUBb = 1, bwd2 = 0, n_tchunks = 1, n_lanes = 1, n_lanes_b = 1;   // forward / backward trials per CTA and concurrent launches
  int tc_begin[MAX_CHUNKS + 1];
  bool plans_ok = false, use_unfold_copy = false;
  GemmPlan p_day, p_head, p_dwout, p_dytop, p_daydw;
  std::vector<GemmPlan> p_dwih, p_dwhh, p_dwih0;   // p_dwih0: layer-0 dW_ih per time chunk (accumulating)
  int comm_sms = 0;                               // SMs the backward tail leaves to the collective (0 = none reserved)
  GemmPlan p_dwih_b, p_dwhh_b, p_dwih0_h[2];
  CUtensorMap tm_hseq[STACK_MAX_LAYERS], tm_dgh[STACK_MAX_LAYERS];   // operand views for the TMA staging of the stack kernels
  bool stk_tma = false;
  bool dwih0_split = false;                     // all layers' dW_ih (l >= 1) / dW_hh in one batched launch each (stack schedule)
  bool dw_batched = false;
  std::vector<std::vector<GemmPlan>> p_in, p_dx;     // [layer >= 1][chunk]
  // side streams / events of the wave-front
  cudaStream_t lane[MAX_LANES + 2] = {};   // [MAX_LANES] = bulk stream (layer-0 projections / data gradients), [MAX_LANES + 1] = second bulk stream (weight gradients)
  cudaEvent_t ev_start = nullptr, ev_lane_end[MAX_LANES + 2] = {}, ev_top = nullptr, ev_init = nullptr, ev_head = nullptr;
  std::vector<cudaEvent_t> ev_r, ev_dx;              // [layer * MAX_CHUNKS + chunk]
  cudaEvent_t ev_g0[MAX_CHUNKS] = {};                // layer-0 input projection chunks (issued ahead on the bulk stream)
  // state of the last forward
  bool have_fwd = false, have_dlogits = false, fwd_training = false;
  unsigned long long seed = 0;
  const int* day_idx = nullptr;
  bool states_given = false;
  long long* trace = nullptr;   // optional device buffer [2][T'][8] for the recurrence cycle trace
  // whole-stack persistent recurrence (gru_stack.cuh): one cooperative launch per direction, gated GEMMs beside it
  int stack = 0, stk_BG = 32, stk_NSUB = 2, stk_ncg = 1, stk_grid = 0, stk_need = 0, stk_gemm_ctas = 1;
  int* ctr = nullptr;           // [4][L][ctr_stride] progress / completion counters: fwd_prog, gx_done, bwd_prog, dy_done
  int ctr_stride = 0;
  cudaStream_t gstream[STACK_MAX_LAYERS] = {};
  cudaEvent_t ev_g[STACK_MAX_LAYERS] = {};
  // gradient buckets (contiguous ranges of the flat gradient buffer) with the event after which each is final: lets a data-parallel
  // caller start the all-reduce of a bucket while backward is still producing the others (b2t_grad_bucket*)
  std::vector<cudaEvent_t> ev_bucket;
  std::vector<std::pair<long long, long long>> bucket_range;   // [offset, count)
  std::vector<int> bucket_order;                                // buckets in the order they become final
};

// Optional timeline (B2T_TIMELINE=1): CUDA events around every task of a step, dumped by b2t_debug_dump_timeline.
struct TlRec { std::string name; int lane; cudaEvent_t a, b; };
static std::vector<TlRec> g_tl;
static bool g_tl_on = false;
struct TlScope {
  cudaStream_t st; size_t idx; bool on;
  TlScope(const char* name, int lane, cudaStream_t s) : st(s), idx(0), on(g_tl_on) {
    if (!on) return;
    TlRec r; r.name = name; r.lane = lane;
    cudaEventCreate(&r.a); cudaEventCreate(&r.b);
    cudaEventRecord(r.a, st);
    g_tl.push_back(r); idx = g_tl.size() - 1;
  }
  ~TlScope() { if (on) cudaEventRecord(g_tl[idx].b, st); }
};
extern "C" int b2t_debug_timeline(int enable) { g_tl_on = enable != 0; for (auto& r : g_tl) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } g_tl.clear(); return 0; }
extern "C" int b2t_debug_dump_timeline(char* buf, int cap) {
  cudaDeviceSynchronize();
  int n = 0;
  for (auto& r : g_tl) {
    float t0 = 0, t1 = 0;
    cudaEventElapsedTime(&t0, g_tl[0].a, r.a); cudaEventElapsedTime(&t1, g_tl[0].a, r.b);
    n += snprintf(buf + n, cap - n > 0 ? cap - n : 0, "%d %s %.1f %.1f\n", r.lane, r.name.c_str(), t0 * 1e3, t1 * 1e3);
    if (n >= cap) break;
  }
  return n;
}

static long long seg_off(const b2t_engine* e, const std::string& n) {
  for (auto& s : e->segs)
    if (s.name == n) return s.offset;
  return -1;
}

static size_t carve(b2t_engine* e, void* ws, size_t cap, bool dry) {
  Carver c{reinterpret_cast<uint8_t*>(ws), 0, cap, dry};
  const int Bp = r16(e->maxB), T = e->maxT, D = e->D, H = e->H, L = e->L;
  const int Tp = e->cfg.patch_size > 0 ? (T - e->patch) / e->stride + 1 : T;
  const int Tq = Tp > 0 ? Tp : 1;
  const size_t M = (size_t)Tq * Bp;
  const bool tr = e->training != 0;
  const int NS = H / 32, NG = Bp / 16;
  e->shadow = c.take<__nv_bfloat16>(e->n_params);
  e->xs = c.take<__nv_bfloat16>((size_t)Bp * T * D);
  e->xd = c.take<__nv_bfloat16>((size_t)Bp * T * D);
  e->xu = c.take<__nv_bfloat16>(M * e->K0);
  e->lay.resize(L);
  for (int l = 0; l < L; ++l) {
    LayerBuf& b = e->lay[l];
    b.gx = c.take<float>(M * 3 * H);
    b.hseq = c.take<__nv_bfloat16>((M + Bp) * H);
    __nv_bfloat16* hd = tr ? c.take<__nv_bfloat16>(M * H) : nullptr;   // reserved for every layer so that the per-layer blocks have one stride (batched dW GEMMs)
    b.hdrop = (l < L - 1) ? hd : nullptr;
    b.kmask = tr ? c.take<unsigned char>(M * H / 4) : nullptr;
    b.R = tr ? c.take<__nv_bfloat16>(M * H) : nullptr;
    b.Z = tr ? c.take<__nv_bfloat16>(M * H) : nullptr;
    b.Nn = tr ? c.take<__nv_bfloat16>(M * H) : nullptr;
    b.HN = tr ? c.take<__nv_bfloat16>(M * H) : nullptr;
    b.h_state = c.take<float>((size_t)Bp * H);
    b.dGx = b.dGh = nullptr; b.dY = b.part = b.dh_state = nullptr;
    if (tr) {
      b.dGx = c.take<__nv_bfloat16>(M * 3 * H);
      b.dGh = c.take<__nv_bfloat16>(M * 3 * H);
      b.dY = c.take<float>(M * H);
      b.part = c.take<float>((size_t)2 * NG * NS * NS * 16 * 32);
      b.dh_state = c.take<float>((size_t)Bp * H);
    }
  }
  e->ctr_stride = (int)r64((long long)Tq + (long long)(M + 127) / 128 + 2);
  e->ctr = c.take<int>((size_t)4 * L * e->ctr_stride);
  e->logits = c.take<float>(M * LDL);
  e->dlog32 = c.take<float>(M * LDL);
  e->dlog16 = c.take<__nv_bfloat16>(M * LDL);
  e->alpha = c.take<float>((size_t)e->maxB * Tq * (2 * e->maxS + 1));
  e->beta = c.take<float>((size_t)e->maxB * Tq * (2 * e->maxS + 1));
  e->greedy_scratch = c.take<int>((size_t)e->maxB * 2 * (e->maxS + 1));
  if (tr) {
    e->dxu = c.take<__nv_bfloat16>(M * e->K0);
    e->dpre = c.take<__nv_bfloat16>((size_t)Bp * T * D);
    e->sumsq = c.take<float>(4);
    e->stats = e->sumsq ? e->sumsq + 1 : nullptr;
    e->steps = c.take<int>(e->segs.size());
    e->d_segs = c.take<Segment>(e->segs.size());
    size_t nch = 0;
    for (auto& s : e->segs) nch += (size_t)((s.rows * s.cols + OPT_CHUNK - 1) / OPT_CHUNK);
    e->n_chunks = (int)nch;
    e->d_chunks = c.take<ChunkRef>(nch);
  }
  return c.off + 1024;
}

extern "C" long long b2t_workspace_bytes(const b2t_config* cfg, int max_batch, int max_T, int max_label_len, int training) {
  if (check_cfg(cfg)) return B2T_ERR_ARG;
  b2t_engine e;
  e.cfg = *cfg; e.maxB = max_batch; e.maxT = max_T; e.maxS = max_label_len > 0 ? max_label_len : 1; e.training = training;
  e.D = cfg->neural_dim; e.H = cfg->n_units; e.L = cfg->n_layers; e.C = cfg->n_classes;
  e.patch = cfg->patch_size > 0 ? cfg->patch_size : 1; e.stride = cfg->patch_size > 0 ? cfg->patch_stride : 1;
  e.K0 = e.D * e.patch;
  e.segs = build_layout(*cfg, &e.n_params);
  return (long long)carve(&e, nullptr, 0, true);
}

extern "C" void b2t_engine_destroy(b2t_engine* e) {
  if (!e) return;
  for (int i = 0; i <= MAX_LANES + 1; ++i) {
    if (e->lane[i]) cudaStreamDestroy(e->lane[i]);
    if (e->ev_lane_end[i]) cudaEventDestroy(e->ev_lane_end[i]);
  }
  for (cudaEvent_t ev : e->ev_bucket) if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < STACK_MAX_LAYERS; ++i) {
    if (e->gstream[i]) cudaStreamDestroy(e->gstream[i]);
    if (e->ev_g[i]) cudaEventDestroy(e->ev_g[i]);
  }
  if (e->ev_start) cudaEventDestroy(e->ev_start);
  if (e->ev_top) cudaEventDestroy(e->ev_top);
  if (e->ev_head) cudaEventDestroy(e->ev_head);
  if (e->ev_init) cudaEventDestroy(e->ev_init);
  for (cudaEvent_t ev : e->ev_r) if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->ev_dx) if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->ev_g0) if (ev) cudaEventDestroy(ev);
  delete e;
}

extern "C" b2t_engine* b2t_engine_create(const b2t_config* cfg, int max_batch, int max_T, int max_label_len, int training, float* params,
                                         float* grads, float* exp_avg, float* exp_avg_sq, void* workspace, long long workspace_bytes) {
  if (check_cfg(cfg)) return nullptr;
  if (!params || !workspace || max_batch < 1 || max_T < 1) { fail(B2T_ERR_ARG, "null buffer or bad sizes"); return nullptr; }
  if (training && (!grads || !exp_avg || !exp_avg_sq)) { fail(B2T_ERR_ARG, "training engine needs grads/exp_avg/exp_avg_sq"); return nullptr; }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    fail(B2T_ERR_CUDA, "no CUDA device"); return nullptr;
  }
  if (major != 10) { fail(B2T_ERR_UNSUPPORTED, "this library contains sm_100a code only (device is sm_%d*)", major); return nullptr; }
  const long long need = b2t_workspace_bytes(cfg, max_batch, max_T, max_label_len, training);
  if (workspace_bytes < need) { fail(B2T_ERR_WORKSPACE, "workspace too small: %lld < %lld", workspace_bytes, need); return nullptr; }
  b2t_engine* e = new b2t_engine();
  e->cfg = *cfg; e->maxB = max_batch; e->maxT = max_T; e->maxS = max_label_len > 0 ? max_label_len : 1; e->training = training;
  e->D = cfg->neural_dim; e->H = cfg->n_units; e->L = cfg->n_layers; e->C = cfg->n_classes;
  e->patch = cfg->patch_size > 0 ? cfg->patch_size : 1; e->stride = cfg->patch_size > 0 ? cfg->patch_stride : 1;
  e->K0 = e->D * e->patch;
  e->segs = build_layout(*cfg, &e->n_params);
  e->params = params; e->grads = grads; e->m1 = exp_avg; e->m2 = exp_avg_sq;
  carve(e, workspace, (size_t)workspace_bytes, false);
  e->touched = grads ? grads + e->n_params : nullptr;
  bool ok = cudaEventCreateWithFlags(&e->ev_start, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_top, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_head, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&e->ev_init, cudaEventDisableTiming) == cudaSuccess;
  // recurrence lanes outrank the bulk stream: when SMs free up, the latency-critical cooperative launches are placed first
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  for (int i = 0; i <= MAX_LANES + 1 && ok; ++i)
    ok = cudaStreamCreateWithPriority(&e->lane[i], cudaStreamNonBlocking, (i < MAX_LANES) != (env_int("B2T_BULK_PRIO", 0) != 0) ? prio_hi : prio_lo) == cudaSuccess &&
         cudaEventCreateWithFlags(&e->ev_lane_end[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < STACK_MAX_LAYERS && ok; ++i)
    ok = cudaStreamCreateWithPriority(&e->gstream[i], cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
         cudaEventCreateWithFlags(&e->ev_g[i], cudaEventDisableTiming) == cudaSuccess;
  e->ev_r.assign((size_t)e->L * MAX_CHUNKS, nullptr);
  e->ev_dx.assign((size_t)e->L * MAX_CHUNKS, nullptr);
  for (size_t i = 0; i < e->ev_r.size() && ok; ++i)
    ok = cudaEventCreateWithFlags(&e->ev_r[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&e->ev_dx[i], cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < MAX_CHUNKS && ok; ++i) ok = cudaEventCreateWithFlags(&e->ev_g0[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { fail(B2T_ERR_CUDA, "stream/event creation failed"); b2t_engine_destroy(e); return nullptr; }
  // kernels that wait on each other while running must all be loaded beforehand (lazy module loading would dead-lock them)
  {
    cudaError_t pe = gemm_preload();
    if (pe == cudaSuccess) pe = cudaFuncSetAttribute(gru_stack_fwd_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<32, 2>::fwd_smem_bytes(e->H)));
    if (pe == cudaSuccess) pe = cudaFuncSetAttribute(gru_stack_fwd_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<32, 1>::fwd_smem_bytes(e->H)));
    if (pe == cudaSuccess) pe = cudaFuncSetAttribute(gru_stack_fwd_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<16, 1>::fwd_smem_bytes(e->H)));
    if (pe == cudaSuccess) pe = cudaFuncSetAttribute(gru_stack_fwd_kernel<32, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<32, 2>::fwd_smem_bytes(e->H)));
    if (pe == cudaSuccess && e->H % 256 == 0) {
      pe = cudaFuncSetAttribute(gru_stack_bwd_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<32, 2>::bwd_smem_bytes(e->H)));
      if (pe == cudaSuccess) pe = cudaFuncSetAttribute(gru_stack_bwd_kernel<32, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<32, 2>::bwd_smem_bytes(e->H)));
      if (pe == cudaSuccess) pe = cudaFuncSetAttribute(gru_stack_bwd_kernel<32, 2, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<32, 2>::bwd_smem_bytes(e->H)));
      if (pe == cudaSuccess) pe = cudaFuncSetAttribute(gru_stack_bwd_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<32, 1>::bwd_smem_bytes(e->H)));
      if (pe == cudaSuccess) pe = cudaFuncSetAttribute(gru_stack_bwd_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(REC_SMEM_BYTES, StackCfg<16, 1>::bwd_smem_bytes(e->H)));
    }
    if (pe != cudaSuccess) { fail(B2T_ERR_CUDA, "kernel preload failed: %s", cudaGetErrorString(pe)); b2t_engine_destroy(e); return nullptr; }
  }
  if (training) {
    // buckets: 0 = day layers; 1 = layer-0 input weights, first half of the rows; 2 + l = rest of layer l (layer 0: W_hh + biases);
    // L + 2 = head, h0, touched flags; L + 3 = layer-0 input weights, second half of the rows
    const long long n_grad = e->n_params + r64(e->cfg.n_days);
    auto off = [&](const std::string& n) { return seg_off(e, n); };
    e->bucket_range.push_back({0, off("gru.weight_ih_l0")});
    const long long wih0_half = (long long)(3 * e->H / 2) * e->K0;   // rows [0, 3H/2) of W_ih0; the rest is bucket L + 3
    e->bucket_range.push_back({off("gru.weight_ih_l0"), wih0_half});
    for (int l = 0; l < e->L; ++l) {
      const long long b = l == 0 ? off("gru.weight_hh_l0") : off("gru.weight_ih_l" + std::to_string(l));
      const long long en = l + 1 < e->L ? off("gru.weight_ih_l" + std::to_string(l + 1)) : off("out.weight");
      e->bucket_range.push_back({b, en - b});
    }
    e->bucket_range.push_back({off("out.weight"), n_grad - off("out.weight")});
    e->bucket_range.push_back({off("gru.weight_ih_l0") + wih0_half, off("gru.weight_hh_l0") - off("gru.weight_ih_l0") - wih0_half});
    e->ev_bucket.assign(e->bucket_range.size(), nullptr);
    for (auto& ev : e->ev_bucket)
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { fail(B2T_ERR_CUDA, "event creation failed"); b2t_engine_destroy(e); return nullptr; }
    std::vector<Segment> hs;
    std::vector<ChunkRef> hc;
    for (size_t i = 0; i < e->segs.size(); ++i) {
      auto& s = e->segs[i];
      hs.push_back({s.offset, s.rows * s.cols, s.group, s.day});
      for (long long f = 0; f < s.rows * s.cols; f += OPT_CHUNK) hc.push_back({(int)i, (int)f});
    }
    if (cudaMemcpy(e->d_segs, hs.data(), hs.size() * sizeof(Segment), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(e->d_chunks, hc.data(), hc.size() * sizeof(ChunkRef), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(e->steps, 0, e->segs.size() * sizeof(int)) != cudaSuccess) {
      fail(B2T_ERR_CUDA, "engine setup copies failed");
      b2t_engine_destroy(e);
      return nullptr;
    }
  }
  return e;
}
extern "C" int* b2t_step_counters(b2t_engine* e) { return e ? e->steps : nullptr; }
extern "C" int b2t_debug_set_trace(b2t_engine* e, long long* buf) {
  if (!e) return fail(B2T_ERR_ARG, "null engine");
  e->trace = buf;
  return 0;
}

// Test hook: device pointer of an internal activation buffer of the last forward/backward (layouts in DESIGN.md section 2).
extern "C" int b2t_debug_buffer(b2t_engine* e, const char* name, int layer, void** ptr, long long* elems) {
  if (!e || !name || !ptr) return fail(B2T_ERR_ARG, "null argument");
  if (layer < 0 || layer >= e->L) return fail(B2T_ERR_ARG, "layer out of range");
  const std::string n(name);
  const LayerBuf& b = e->lay[layer];
  const long long Bp = e->Bpad, M = e->M, H = e->H;
  void* p = nullptr;
  long long cnt = 0;
  if (n == "xs") { p = e->xs; cnt = Bp * e->T_in * e->D; }
  else if (n == "xd") { p = e->xd; cnt = Bp * e->T_in * e->D; }
  else if (n == "dpre") { p = e->dpre; cnt = Bp * e->T_in * e->D; }
  else if (n == "hseq") { p = b.hseq; cnt = (M + Bp) * H; }
  else if (n == "hdrop") { p = b.hdrop; cnt = M * H; }
  else if (n == "gx") { p = b.gx; cnt = M * 3 * H; }
  else if (n == "dGx") { p = b.dGx; cnt = M * 3 * H; }
  else if (n == "dGh") { p = b.dGh; cnt = M * 3 * H; }
  else if (n == "dY") { p = b.dY; cnt = M * H; }
  else if (n == "logits") { p = e->logits; cnt = M * LDL; }
  else return fail(B2T_ERR_ARG, "unknown buffer '%s'", name);
  if (!p) return fail(B2T_ERR_STATE, "buffer '%s' does not exist in this engine (inference-only, or last layer)", name);
  *ptr = p;
  if (elems) *elems = cnt;
  return 0;
}

extern "C" int b2t_refresh_weights(b2t_engine* e, void* stream) {
  if (!e) return fail(B2T_ERR_ARG, "null engine");
  cudaStream_t st = (cudaStream_t)stream;
  cast_bf16_kernel<<<num_sms() * 4, 256, 0, st>>>(e->params, e->shadow, (size_t)e->n_params);
  CK(LAUNCHED());
  return 0;
}

// ------------------------------------------------------------------------------------ plans
static int build_plans(b2t_engine* e) {
  e->dwih0_split = false;
  e->poll_delay = env_int("B2T_POLL_DELAY", 700);
  e->poll_delay_b = env_int("B2T_POLL_DELAY_BWD", 300);
  const int D = e->D, H = e->H, L = e->L, Bp = e->Bpad, Tp = e->Tp, M = e->M, K0 = e->K0, T = e->T_in;
  const bool tr = e->training != 0;
  // Data parallel: the gradient all-reduce of the early buckets runs while the tail of backward (persistent GEMMs, one CTA per SM)
  // still computes; NCCL's CTAs only get SMs the GEMMs leave free, so the tail GEMMs can be told to leave some (b2t_set_comm_sms)
  const int tail_ctas = e->comm_sms > 0 ? std::max(num_sms() - e->comm_sms, 8) : 0;
  // ---- recurrence geometry: trials per CTA, concurrent launches, time chunks.  A tcgen05.mma costs the same for every
  //      N <= 64 (profiles/r1_mma_dispatch_microbench.md), so wider batch groups need fewer CTAs for the same MMA time.
  auto pick_bg = [&](const char* env) {
    // BG = 64 halves the CTAs but doubles what each CTA moves through L2 per step; measured on B200 the step then takes
    // 1.7-1.9x as long (the exchange is bound per SM), so 32 is the default and 64 stays an opt-in.
    int bg = (Bp % 64 == 0) ? 64 : (Bp % 32 == 0) ? 32 : 16;
    const int cap = env_int(env, 32);
    while (bg > 16 && bg > cap) bg /= 2;
    while (bg < 64 && Bp % (2 * bg) == 0 && (H / 32) * (Bp / bg) > num_sms()) bg *= 2;   // large batches: all CTAs must be co-resident
    return bg;
  };
  e->BG = pick_bg("B2T_REC_BG_FWD");
  e->BGb = pick_bg("B2T_REC_BG_BWD");
  // backward: a CTA can host NSUB independent batch groups (gru_rec.cuh).  Measured on B200: two groups of 16 take as long per
  // step as one group of 32 (the step is a chain of latencies, not of bytes), so one group per CTA stays the default.
  e->NSUBb = 1;
  {
    const int want = env_int("B2T_REC_NSUB_BWD", 0);
    if (want > 0) e->NSUBb = want;
    while (e->NSUBb > 1 && ((Bp / e->BGb) % e->NSUBb != 0 || !((e->BGb == 16 && (e->NSUBb == 2 || e->NSUBb == 4)) || (e->BGb == 32 && e->NSUBb == 2)))) e->NSUBb /= 2;
  }
  // backward kernel: the 2-D decomposition (4-way partial reduction + dG all-gather) needs H/4 to be a multiple of 64
  e->bwd2 = (H % 256 == 0 && e->BGb <= 32 && env_int("B2T_REC_BWD2", 0) != 0) ? 1 : 0;
  if (e->bwd2) e->NSUBb = 1;
  auto pick_lanes = [&](int bg, const char* env) {
    const int ctas = (H / 32) * (Bp / bg);
    int n = std::max(1, std::min(MAX_LANES, num_sms() / std::max(ctas, 1)));
    n = std::min(n, L);
    return std::max(1, std::min(n, env_int(env, MAX_LANES)));
  };
  e->n_lanes = pick_lanes(e->BG, "B2T_REC_LANES_FWD");
  e->n_lanes_b = pick_lanes(e->BGb * e->NSUBb, "B2T_REC_LANES_BWD");
  int nch = env_int("B2T_REC_CHUNKS", 0);
  if (nch <= 0) nch = (std::max(e->n_lanes, e->n_lanes_b) > 1 && Tp >= 48) ? (std::max(e->n_lanes, e->n_lanes_b) >= 5 ? 6 : 3) : 1;
  nch = std::max(1, std::min(std::min(nch, MAX_CHUNKS), Tp));
  // ---- whole-stack persistent recurrence: every layer resident at once, two batch groups per CTA when the group count is even
  {
    const int bg = (Bp % 32 == 0) ? 32 : 16;
    const int ng = Bp / bg;
    const int nsub = (ng % 2 == 0) ? 2 : 1;
    const int grid = L * (H / 32) * (ng / nsub);
    const size_t smem_f = bg == 32 ? (nsub == 2 ? StackCfg<32, 2>::fwd_smem_bytes(H) : StackCfg<32, 1>::fwd_smem_bytes(H))
                                   : StackCfg<16, 1>::fwd_smem_bytes(H);
    // Tools that serialise kernel execution (Nsight Compute's kernel replay, compute-sanitizer) would dead-lock kernels that wait
    // for each other while running (the persistent recurrence and its gated GEMMs): under them the stack schedule is off unless
    // B2T_STACK=1 forces it (single-layer models have no cross-kernel dependency and can be profiled with it).
    static const bool serialising_tool = getenv("NV_NSIGHT_INJECTION_PORT_BASE") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") ||
                                         getenv("NV_SANITIZER_INJECTION_PORT_BASE") ||
                                         (getenv("CUDA_LAUNCH_BLOCKING") && atoi(getenv("CUDA_LAUNCH_BLOCKING")) != 0);   // blocking launches serialise as well
    bool ok = env_int("B2T_STACK", serialising_tool ? 0 : 1) != 0 && L <= STACK_MAX_LAYERS && grid + std::max(L - 1, 0) <= num_sms() && smem_f <= 227 * 1024 &&
              128 % bg == 0 && (nsub * bg) <= 128 && (128 % (nsub * bg) == 0);
    if (tr) ok = ok && (H % 256 == 0);          // backward uses the two-dimensional decomposition (H/128 x 4 CTAs per batch group)
    e->stack = ok ? 1 : 0;
    if (ok) {
      e->stk_BG = bg; e->stk_NSUB = nsub; e->stk_ncg = ng / nsub; e->stk_grid = grid;
      e->stk_need = (H / 32) * ng;              // one signal per CTA and batch group of a layer
      e->stk_gemm_ctas = L > 1 ? std::max(1, (num_sms() - grid) / (L - 1)) : 1;
      nch = 1;                                  // no time chunking: the GEMMs beside the recurrence are gated per 128-row tile instead
    }
  }
  e->n_tchunks = nch;
  for (int c = 0; c <= nch; ++c) e->tc_begin[c] = (int)((long long)c * Tp / nch);
  e->p_dwih.assign(L, GemmPlan());
  e->p_dwhh.assign(L, GemmPlan());
  e->p_in.assign(L, std::vector<GemmPlan>(nch));
  e->p_dx.assign(L, std::vector<GemmPlan>(nch));
  int rc;
  {  // day layer: xd[b] = softsign(xs[b] @ W_day[day_b] + b_day[day_b]) (+dropout)      rnn_model.py:95-103
    GemmSpec s;
    s.a_mn = 0; s.b_mn = 1; s.epi = EPI_DAY; s.out_bf16 = 1;
    s.M = e->T_out; s.N = D; s.K = D;
    s.A = e->xs; s.lda = D; s.a_zstride = (long long)T * D; s.nz = e->B;
    s.B = e->shadow + seg_off(e, "day_weights.0"); s.ldb = D; s.b_zstride = (long long)D * D; s.nzb = e->cfg.n_days;
    s.z_map = e->day_idx; s.zmap_b = 1;
    s.C = e->xd; s.ldc = D; s.c_zstride = (long long)T * D;
    s.bias = e->params + seg_off(e, "day_biases.0"); s.bias_zstride = r64(D);
    s.keep = 1.0f;
    if ((rc = gemm_plan_build(&e->p_day, s))) return fail(B2T_ERR_CUDA, "day-layer plan failed (%d)", rc);
  }
  // layer-0 input projection, one GEMM per time chunk so that the layer-0 recurrence can start after the first one
  for (int c = 0; c < nch; ++c) {
    const int t0 = e->tc_begin[c], nt = e->tc_begin[c + 1] - e->tc_begin[c];
    GemmSpec s;
    s.a_mn = 0; s.b_mn = 0; s.epi = EPI_STORE; s.out_bf16 = 0;
    s.N = 3 * H; s.K = K0; s.ldb = K0; s.M = (long long)nt * Bp;
    s.B = e->shadow + seg_off(e, "gru.weight_ih_l0");
    s.C = e->lay[0].gx + (size_t)t0 * Bp * 3 * H; s.ldc = 3 * H;
    s.bias = e->params + seg_off(e, "gru.bias_ih_l0");
    rc = -1;
    if (!e->use_unfold_copy) {   // strided patch view of xd (rnn_model.py:106-119), never materialised
      s.A = e->xd + (size_t)t0 * e->stride * D; s.a_rin = Bp; s.a_rout = nt; s.a_rin_stride = (long long)T * D; s.a_rout_stride = (long long)e->stride * D;
      rc = gemm_plan_build(&e->p_in[0][c], s);
      if (rc) {
        fprintf(stderr, "b2t: strided patch tensor map rejected (%d); falling back to a materialised unfold\n", rc);
        e->use_unfold_copy = true;
      }
    }
    if (e->use_unfold_copy) {
      s.a_rin = 0; s.A = e->xu + (size_t)t0 * Bp * K0; s.lda = K0;
      if ((rc = gemm_plan_build(&e->p_in[0][c], s))) return fail(B2T_ERR_CUDA, "L0 input plan failed (%d)", rc);
    }
  }
  for (int l = 0; l < L; ++l) {
    const std::string sl = std::to_string(l);
    if (l > 0) {
      const __nv_bfloat16* in = e->lay[l - 1].hdrop ? e->lay[l - 1].hdrop : e->lay[l - 1].hseq + (size_t)Bp * H;
      for (int c = 0; c < nch; ++c) {      // per time chunk: rows [t0*Bp, t1*Bp)
        const long long r0 = (long long)e->tc_begin[c] * Bp, rows = (long long)(e->tc_begin[c + 1] - e->tc_begin[c]) * Bp;
        GemmSpec s;
        s.a_mn = 0; s.b_mn = 0; s.epi = EPI_STORE; s.out_bf16 = 0;
        s.M = rows; s.N = 3 * H; s.K = H; s.lda = H; s.ldb = H;
        s.A = in + r0 * H;
        s.B = e->shadow + seg_off(e, "gru.weight_ih_l" + sl);
        s.C = e->lay[l].gx + r0 * 3 * H; s.ldc = 3 * H;
        s.bias = e->params + seg_off(e, "gru.bias_ih_l" + sl);
        if (e->stack) {   // follows the recurrence of layer l-1 tile by tile and reports to the recurrence of layer l
          s.gate = e->ctr + (size_t)(0 * L + (l - 1)) * e->ctr_stride; s.gate_need = e->stk_need; s.gate_rows_per_step = Bp; s.gate_steps = Tp;
          s.bn = env_int("B2T_GATED_BN_FWD", 0);
          s.done = e->ctr + (size_t)(1 * L + l) * e->ctr_stride; s.max_ctas = e->stk_gemm_ctas;
        }
        if ((rc = gemm_plan_build(&e->p_in[l][c], s))) return fail(B2T_ERR_CUDA, "input plan %d/%d failed (%d)", l, c, rc);
      }
    }
  }
  {  // head: logits = top @ W_out^T + b_out                                       rnn_model.py:129
    GemmSpec s;
    s.a_mn = 0; s.b_mn = 0; s.epi = EPI_STORE; s.out_bf16 = 0;
    s.M = M; s.N = e->C; s.K = H;
    s.A = e->lay[L - 1].hseq + (size_t)Bp * H; s.lda = H;
    s.B = e->shadow + seg_off(e, "out.weight"); s.ldb = H;
    s.C = e->logits; s.ldc = LDL;
    s.bias = e->params + seg_off(e, "out.bias");
    if ((rc = gemm_plan_build(&e->p_head, s))) return fail(B2T_ERR_CUDA, "head plan failed (%d)", rc);
  }
  if (!tr) return 0;
  {  // dW_out[C][H] = dlogits^T top
    GemmSpec s;
    s.a_mn = 1; s.b_mn = 1; s.epi = EPI_STORE;
    s.M = e->C; s.N = H; s.K = M;
    s.A = e->dlog16; s.lda = LDL; s.B = e->lay[L - 1].hseq + (size_t)Bp * H; s.ldb = H;
    s.C = e->grads + seg_off(e, "out.weight"); s.ldc = H;
    if ((rc = gemm_plan_build(&e->p_dwout, s))) return fail(B2T_ERR_CUDA, "dW_out plan failed (%d)", rc);
  }
  {  // dY_top[M][H] = dlogits W_out
    GemmSpec s;
    s.a_mn = 0; s.b_mn = 1; s.epi = EPI_STORE;
    s.M = M; s.N = H; s.K = e->C;
    s.A = e->dlog16; s.lda = LDL; s.B = e->shadow + seg_off(e, "out.weight"); s.ldb = H;
    s.C = e->lay[L - 1].dY; s.ldc = H;
    if ((rc = gemm_plan_build(&e->p_dytop, s))) return fail(B2T_ERR_CUDA, "dY_top plan failed (%d)", rc);
  }
  for (int l = 0; l < L; ++l) {
    const std::string sl = std::to_string(l);
    if (l == 0) {   // dW_ih0 = sum over time chunks of dGx0[chunk]^T X_unf[chunk]: chunk GEMMs fill idle SMs while the last recurrences run
      e->p_dwih0.assign(nch, GemmPlan());
      for (int c = 0; c < nch; ++c) {
        const int t0 = e->tc_begin[c], nt = e->tc_begin[c + 1] - e->tc_begin[c];
        GemmSpec s;
        s.a_mn = 1; s.b_mn = 1; s.epi = (c == nch - 1) ? EPI_STORE : EPI_ACCUM;    // the LAST time chunk is computed first and initialises the buffer
        s.M = 3 * H; s.K = (long long)nt * Bp; s.N = K0; s.ldc = K0;
        s.A = e->lay[0].dGx + (size_t)t0 * Bp * 3 * H; s.lda = 3 * H;
        s.C = e->grads + seg_off(e, "gru.weight_ih_l0");
        if (!e->use_unfold_copy) {
          s.k_rin = Bp; s.k_rout = nt;
          s.a_rin_stride = 3 * H; s.a_rout_stride = (long long)Bp * 3 * H;
          s.B = e->xd + (size_t)t0 * e->stride * D; s.b_rin_stride = (long long)T * D; s.b_rout_stride = (long long)e->stride * D;
        } else {
          s.B = e->xu + (size_t)t0 * Bp * K0; s.ldb = K0;
        }
        s.max_ctas = tail_ctas;
        if ((rc = gemm_plan_build(&e->p_dwih0[c], s))) return fail(B2T_ERR_CUDA, "dW_ih0 plan %d failed (%d)", c, rc);
        if (nch == 1 && (3 * H / 2) % 128 == 0) {   // the same product as two row halves: each is a gradient bucket of its own, so the all-reduce of the first runs beside the GEMM of the second
          e->dwih0_split = true;
          for (int h = 0; h < 2; ++h) {
            GemmSpec hs = s;
            hs.M = 3 * H / 2;
            hs.A = (const __nv_bfloat16*)s.A + (size_t)h * (3 * H / 2);
            hs.C = (float*)s.C + (size_t)h * (3 * H / 2) * K0;
            if ((rc = gemm_plan_build(&e->p_dwih0_h[h], hs))) { e->dwih0_split = false; break; }
          }
        }
      }
    } else {  // dW_ih = dGx^T X
      GemmSpec s;
      s.a_mn = 1; s.b_mn = 1; s.epi = EPI_STORE;
      s.M = 3 * H; s.K = M;
      s.A = e->lay[l].dGx; s.lda = 3 * H;
      s.C = e->grads + seg_off(e, "gru.weight_ih_l" + sl);
      s.N = H; s.ldc = H; s.ldb = H;
      s.B = e->lay[l - 1].hdrop ? e->lay[l - 1].hdrop : e->lay[l - 1].hseq + (size_t)Bp * H;
      if ((rc = gemm_plan_build(&e->p_dwih[l], s))) return fail(B2T_ERR_CUDA, "dW_ih plan %d failed (%d)", l, rc);
    }
    {  // dW_hh = dGh^T H_prev   (H_prev = hseq slots 0..T-1)
      GemmSpec s;
      s.a_mn = 1; s.b_mn = 1; s.epi = EPI_STORE;
      s.M = 3 * H; s.N = H; s.K = M;
      s.A = e->lay[l].dGh; s.lda = 3 * H; s.B = e->lay[l].hseq; s.ldb = H;
      s.C = e->grads + seg_off(e, "gru.weight_hh_l" + sl); s.ldc = H;
      if ((rc = gemm_plan_build(&e->p_dwhh[l], s))) return fail(B2T_ERR_CUDA, "dW_hh plan %d failed (%d)", l, rc);
    }
    for (int c = 0; c < nch; ++c) {   // data gradient per time chunk: dX_unf (l == 0, folded afterwards) or dY_{l-1} = dGx_l W_ih_l
      const long long r0 = (long long)e->tc_begin[c] * Bp, rows = (long long)(e->tc_begin[c + 1] - e->tc_begin[c]) * Bp;
      GemmSpec s;
      s.a_mn = 0; s.b_mn = 1; s.epi = EPI_STORE;
      s.M = rows; s.K = 3 * H; s.A = e->lay[l].dGx + r0 * 3 * H; s.lda = 3 * H;
      s.B = e->shadow + seg_off(e, "gru.weight_ih_l" + sl);
      if (l == 0) { s.N = K0; s.ldb = K0; s.out_bf16 = 1; s.C = e->dxu + r0 * K0; s.ldc = K0; }
      else {
        s.N = H; s.ldb = H; s.out_bf16 = 0; s.C = e->lay[l - 1].dY + r0 * H; s.ldc = H;
        if (e->stack) {   // follows the backward recurrence of layer l (time descending) and reports to the one of layer l-1
          s.gate = e->ctr + (size_t)(2 * L + l) * e->ctr_stride; s.gate_need = e->stk_need; s.gate_rows_per_step = Bp; s.gate_steps = Tp;
          s.bn = env_int("B2T_GATED_BN_BWD", 128);   // six 128-wide tiles on the GEMM's 7 CTAs finish an M-tile in one round (9.5 us instead of 17: the next layer trails by that much less)
          if (env_int("B2T_DX_DONE", 0)) s.done = e->ctr + (size_t)(3 * L + (l - 1)) * e->ctr_stride;
          s.tm_reverse = 1; s.max_ctas = e->stk_gemm_ctas;   // (no completion counters: the consumer polls the sentinel-filled dY itself)
        }
      }
      if (l == 0) s.max_ctas = tail_ctas;
      if ((rc = gemm_plan_build(&e->p_dx[l][c], s))) return fail(B2T_ERR_CUDA, "dX plan %d/%d failed (%d)", l, c, rc);
    }
  }
  // All layers' dW_ih (l >= 1) and dW_hh have one shape (3H x H over K = T'*Bpad rows): one batched launch each keeps every SM
  // busy for ~3 waves of tiles instead of nine single-wave GEMMs on 108 of the SMs.  Needs one stride between the layers' buffers.
  e->dw_batched = false;
  if (e->stack && L >= 3 && env_int("B2T_DW_BATCHED", 1) != 0) {
    const long long za = e->lay[1].dGx - e->lay[0].dGx;
    const long long zc_ih = seg_off(e, "gru.weight_ih_l2") - seg_off(e, "gru.weight_ih_l1");
    const long long zc_hh = seg_off(e, "gru.weight_hh_l1") - seg_off(e, "gru.weight_hh_l0");
    bool uniform = za > 0 && zc_ih > 0 && zc_hh > 0 && za % 8 == 0;
    for (int l = 0; l + 1 < L && uniform; ++l) {
      const std::string a = std::to_string(l), b = std::to_string(l + 1);
      uniform = e->lay[l + 1].dGx - e->lay[l].dGx == za && e->lay[l + 1].dGh - e->lay[l].dGh == za && e->lay[l + 1].hseq - e->lay[l].hseq == za &&
                (l + 2 >= L || e->lay[l + 1].hdrop - e->lay[l].hdrop == za) && seg_off(e, "gru.weight_hh_l" + b) - seg_off(e, "gru.weight_hh_l" + a) == zc_hh &&
                (l == 0 || seg_off(e, "gru.weight_ih_l" + b) - seg_off(e, "gru.weight_ih_l" + a) == zc_ih);
    }
    if (uniform) {
      GemmSpec s;
      s.a_mn = 1; s.b_mn = 1; s.epi = EPI_STORE; s.bn = 128;
      s.M = 3 * H; s.N = H; s.K = M; s.lda = 3 * H; s.ldb = H; s.ldc = H;
      s.a_zstride = za; s.b_zstride = za; s.max_ctas = tail_ctas;
      GemmSpec ih = s;
      ih.nz = L - 1; ih.A = e->lay[1].dGx; ih.B = e->lay[0].hdrop; ih.C = e->grads + seg_off(e, "gru.weight_ih_l1"); ih.c_zstride = zc_ih;
      GemmSpec hh = s;
      hh.nz = L; hh.A = e->lay[0].dGh; hh.B = e->lay[0].hseq; hh.C = e->grads + seg_off(e, "gru.weight_hh_l0"); hh.c_zstride = zc_hh;
      if (gemm_plan_build(&e->p_dwih_b, ih) == 0 && gemm_plan_build(&e->p_dwhh_b, hh) == 0) e->dw_batched = true;
    }
  }
  e->stk_tma = false;
  // Opt-in (B2T_STACK_TMA=1): after the probe the forward loaders fetch h_{t-1} with TMA tensor loads and verify it in shared memory.
  // Measured slower (3.07 vs 3.00 ms per step): the MMA chain can only start when the whole operand has been verified, while the
  // polled path hands it over chunk by chunk.
  if (e->stack && env_int("B2T_STACK_TMA", 0) != 0) {
    bool ok = true;
    for (int l = 0; l < L && ok; ++l) {
      const uint64_t rows = (uint64_t)(Tp + 1) * Bp;
      uint64_t dims[4] = {(uint64_t)H, rows, 1, 1}, str[3] = {(uint64_t)H, (uint64_t)H * rows, (uint64_t)H * rows};
      uint32_t box[4] = {64, (uint32_t)e->stk_BG, 1, 1};
      ok = make_tmap_bf16_4d(&e->tm_hseq[l], e->lay[l].hseq, dims, str, box) == 0;
      if (ok && tr) {
        const uint64_t rg = (uint64_t)Tp * Bp;
        uint64_t d2[4] = {(uint64_t)3 * H, rg, 1, 1}, s2[3] = {(uint64_t)3 * H, (uint64_t)3 * H * rg, (uint64_t)3 * H * rg};
        ok = make_tmap_bf16_4d(&e->tm_dgh[l], e->lay[l].dGh, d2, s2, box) == 0;
      }
    }
    e->stk_tma = ok;
  }
  {  // dW_day[day_b] += xs[b]^T dpre[b]
    GemmSpec s;
    s.a_mn = 1; s.b_mn = 1; s.epi = EPI_ATOMIC;
    s.M = D; s.N = D; s.K = e->T_out;
    s.A = e->xs; s.lda = D; s.a_zstride = (long long)T * D; s.nz = e->B;
    s.B = e->dpre; s.ldb = D; s.b_zstride = (long long)T * D; s.nzb = e->B;
    s.z_map = e->day_idx; s.zmap_b = 0;
    s.C = e->grads + seg_off(e, "day_weights.0"); s.ldc = D; s.c_zstride = (long long)D * D;
    s.max_ctas = tail_ctas;
    if ((rc = gemm_plan_build(&e->p_daydw, s))) return fail(B2T_ERR_CUDA, "dW_day plan failed (%d)", rc);
  }
  return 0;
}

static int gauss_taps(float std, int size, float* taps16, int* ntaps) {
  // data_augmentations.py:19-24 -- impulse response of scipy gaussian_filter1d(sigma=std), > 0.01, renormalised
  const int radius = (int)(4.0 * std + 0.5);
  std::vector<double> w(2 * radius + 1);
  double sum = 0;
  for (int i = -radius; i <= radius; ++i) { w[i + radius] = exp(-0.5 / ((double)std * std) * i * i); sum += w[i + radius]; }
  std::vector<float> imp(size, 0.f);
  const int c = size / 2;
  for (int i = 0; i < 2 * radius + 1; ++i) {
    const int j = c - radius + i;
    if (j >= 0 && j < size) imp[j] += (float)(w[i] / sum);
  }
  std::vector<float> k;
  for (float v : imp) if (v > 0.01f) k.push_back(v);
  if (k.empty() || k.size() > 16) return -1;
  float s = 0;
  for (float v : k) s += v;
  for (int i = 0; i < 16; ++i) taps16[i] = 0.f;
  for (size_t i = 0; i < k.size(); ++i) taps16[16 - k.size() + i] = k[i] / s;
  *ntaps = (int)k.size();
  return 0;
}

extern "C" int b2t_set_comm_sms(b2t_engine* e, int n_sms) {
  if (!e || n_sms < 0 || n_sms >= num_sms()) return fail(B2T_ERR_ARG, "b2t_set_comm_sms: bad arguments");
  if (e->comm_sms != n_sms) { e->comm_sms = n_sms; e->plans_ok = false; }
  return 0;
}

extern "C" int b2t_output_frames(const b2t_config* cfg, int T, int smooth_mode, int ntaps, int cut) {
  int To = T - cut;
  if (smooth_mode == 2) To -= ntaps - 1;
  if (cfg->patch_size > 0) return To < cfg->patch_size ? 0 : (To - cfg->patch_size) / cfg->patch_stride + 1;
  return To;
}

// The recurrence kernels allocate all 512 TMEM columns, so two CTAs must never share an SM: requesting more than
// half of the shared memory forces one CTA per SM.  Cooperative launch guarantees that every CTA of the grid is
// co-resident (they spin on each other's flags).
template <int BG>
static cudaError_t launch_rec_fwd_t(const RecFwdParams& p, int grid, cudaStream_t st) {
  const size_t smem = std::max(REC_SMEM_BYTES, RecCfg<BG>::fwd_smem_bytes(p.H));
  cudaError_t err = cudaFuncSetAttribute(gru_rec_fwd_kernel<BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  void* args[1] = {(void*)&p};
  ++g_launches;
  return cudaLaunchCooperativeKernel((const void*)gru_rec_fwd_kernel<BG>, dim3(grid), dim3(RecCfg<BG>::kFwdThreads), args, smem, st);
}
template <int BG, int NSUB>
static cudaError_t launch_rec_bwd_t(const RecBwdParams& p, int grid, cudaStream_t st) {
  const size_t smem = std::max(REC_SMEM_BYTES, RecCfg<BG>::bwd_smem_bytes(p.H, NSUB));
  cudaError_t err = cudaFuncSetAttribute(gru_rec_bwd_kernel<BG, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  void* args[1] = {(void*)&p};
  ++g_launches;
  return cudaLaunchCooperativeKernel((const void*)gru_rec_bwd_kernel<BG, NSUB>, dim3(grid), dim3((RecCfg<BG>::kBwdThreads + 32) * NSUB), args, smem, st);
}
static cudaError_t launch_rec_fwd(int BG, const RecFwdParams& p, int grid, cudaStream_t st) {
  return BG == 64 ? launch_rec_fwd_t<64>(p, grid, st) : BG == 32 ? launch_rec_fwd_t<32>(p, grid, st) : launch_rec_fwd_t<16>(p, grid, st);
}
template <int BG>
static cudaError_t launch_rec_bwd2_t(const RecBwdParams& p, int grid, cudaStream_t st) {
  const size_t smem = std::max(REC_SMEM_BYTES, RecBwd2Cfg<BG>::smem_bytes(p.H));
  cudaError_t err = cudaFuncSetAttribute(gru_rec_bwd2_kernel<BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  void* args[1] = {(void*)&p};
  ++g_launches;
  return cudaLaunchCooperativeKernel((const void*)gru_rec_bwd2_kernel<BG>, dim3(grid), dim3(RecBwd2Cfg<BG>::kThreads), args, smem, st);
}
// two-dimensional decomposition (gru_rec_bwd2.cuh): grid = H/32 CTAs per batch group
static cudaError_t launch_rec_bwd2(int BG, const RecBwdParams& p, int grid, cudaStream_t st) {
  return BG == 32 ? launch_rec_bwd2_t<32>(p, grid, st) : launch_rec_bwd2_t<16>(p, grid, st);
}
// backward: `nsub` batch groups of BG trials per CTA (grid = slices x groups / nsub)
static cudaError_t launch_rec_bwd(int BG, int nsub, const RecBwdParams& p, int grid, cudaStream_t st) {
  if (BG == 16 && nsub == 2) return launch_rec_bwd_t<16, 2>(p, grid, st);
  if (BG == 16 && nsub == 4) return launch_rec_bwd_t<16, 4>(p, grid, st);
  if (BG == 32 && nsub == 2) return launch_rec_bwd_t<32, 2>(p, grid, st);
  if (nsub != 1) return cudaErrorInvalidValue;
  return BG == 64 ? launch_rec_bwd_t<64, 1>(p, grid, st) : BG == 32 ? launch_rec_bwd_t<32, 1>(p, grid, st) : launch_rec_bwd_t<16, 1>(p, grid, st);
}

template <int BG, int NSUB, int LSETS = 1>
static cudaError_t launch_stack_fwd_t(const StackFwdParams& p, int grid, cudaStream_t st) {
  const size_t smem = std::max(REC_SMEM_BYTES, StackCfg<BG, NSUB>::fwd_smem_bytes(p.H));
  cudaError_t err = cudaFuncSetAttribute(gru_stack_fwd_kernel<BG, NSUB, LSETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  void* args[1] = {(void*)&p};
  ++g_launches;
  return cudaLaunchCooperativeKernel((const void*)gru_stack_fwd_kernel<BG, NSUB, LSETS>, dim3(grid), dim3(StackCfg<BG, NSUB>::fwd_threads(LSETS)), args, smem, st);
}
template <int BG, int NSUB, int ESETS = 1, int LSPLIT = 0>
static cudaError_t launch_stack_bwd_t(const StackBwdParams& p, int grid, cudaStream_t st) {
  const size_t smem = std::max(REC_SMEM_BYTES, StackCfg<BG, NSUB>::bwd_smem_bytes(p.H));
  cudaError_t err = cudaFuncSetAttribute(gru_stack_bwd_kernel<BG, NSUB, ESETS, LSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  void* args[1] = {(void*)&p};
  ++g_launches;
  return cudaLaunchCooperativeKernel((const void*)gru_stack_bwd_kernel<BG, NSUB, ESETS, LSPLIT>, dim3(grid), dim3(StackCfg<BG, NSUB>::bwd_threads(ESETS)), args, smem, st);
}
static cudaError_t launch_stack_fwd(int BG, int NSUB, const StackFwdParams& p, int grid, cudaStream_t st) {
  static const int lsets = env_int("B2T_FWD_LSETS", 2);   // one set of loader warps per batch group (see gru_stack.cuh)
  if (BG == 32 && NSUB == 2) return lsets == 2 ? launch_stack_fwd_t<32, 2, 2>(p, grid, st) : launch_stack_fwd_t<32, 2>(p, grid, st);
  if (BG == 32) return launch_stack_fwd_t<32, 1>(p, grid, st);
  return NSUB == 1 ? launch_stack_fwd_t<16, 1>(p, grid, st) : cudaErrorInvalidValue;   // 16-trial groups only arise for an odd group count
}
static cudaError_t launch_stack_bwd(int BG, int NSUB, const StackBwdParams& p, int grid, cudaStream_t st) {
  static const int esets = env_int("B2T_BWD_ESETS", 2);   // one set of epilogue warps per batch group (see gru_stack.cuh)
  static const int lsplit = env_int("B2T_BWD_LSPLIT", 0);   // the loader warps as two half sets, one per batch group: measured slower (3.26 vs 2.94 ms per step: half the loaders stage a group twice as slowly), opt-in
  if (BG == 32 && NSUB == 2 && esets == 2 && lsplit) return launch_stack_bwd_t<32, 2, 2, 1>(p, grid, st);
  if (BG == 32 && NSUB == 2) return esets == 2 ? launch_stack_bwd_t<32, 2, 2>(p, grid, st) : launch_stack_bwd_t<32, 2>(p, grid, st);
  if (BG == 32) return launch_stack_bwd_t<32, 1>(p, grid, st);
  return NSUB == 1 ? launch_stack_bwd_t<16, 1>(p, grid, st) : cudaErrorInvalidValue;
}

// Gradient regions that backward accumulates into with atomics (biases, h0, day layers): cleared ahead of time.  GEMM-stored
// gradients are fully overwritten.  Untouched day segments are cleared too -- cheap, and keeps the gradient-norm reduction free
// of stale values.
static int zero_accumulated_grads(b2t_engine* e, cudaStream_t st) {
  const int H = e->H, L = e->L;
  CK(cudaMemsetAsync(e->grads + seg_off(e, "day_weights.0"), 0, (size_t)(seg_off(e, "gru.weight_ih_l0") - seg_off(e, "day_weights.0")) * sizeof(float), st));
  for (int l = 0; l < L; ++l)
    CK(cudaMemsetAsync(e->grads + seg_off(e, "gru.bias_ih_l" + std::to_string(l)), 0, (size_t)2 * r64(3 * H) * sizeof(float), st));
  CK(cudaMemsetAsync(e->grads + seg_off(e, "out.bias"), 0, (size_t)(r64(e->C) + r64(H)) * sizeof(float), st));
  return 0;
}

// ------------------------------------------------------------------------------------ forward
// Wave-front schedule.  Layer l runs on side stream ("lane") l % n_lanes; its time chunk c needs chunk c of the layer
// below (event) and its own chunk c-1 (stream order).  Tasks are issued diagonal by diagonal so that every lane's FIFO
// is in dependency order.  With 48-CTA launches (H = 768, B = 64, BG = 32) three lanes run concurrently on 144 SMs.
extern "C" int b2t_forward(b2t_engine* e, const b2t_forward_args* a, void* stream) {
  if (!e || !a || !a->x || !a->day_idx) return fail(B2T_ERR_ARG, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = e->D, H = e->H, L = e->L;
  if (a->B < 1 || a->B > e->maxB || a->T < 1 || a->T > e->maxT) return fail(B2T_ERR_ARG, "B=%d T=%d exceed engine limits (%d, %d)", a->B, a->T, e->maxB, e->maxT);
  if (a->training && !e->training) return fail(B2T_ERR_STATE, "engine was created for inference only");
  PreParams pp;
  memset(&pp, 0, sizeof(pp));
  int ntaps = 0;
  for (int i = 0; i < 16; ++i) pp.taps[i] = 0.f;
  pp.taps[15] = 1.f;
  if (a->smooth_mode != 0) {
    if (gauss_taps(a->smooth_std, a->smooth_size, pp.taps, &ntaps)) return fail(B2T_ERR_UNSUPPORTED, "smoothing kernel has more than 16 taps");
  }
  const int cut = a->training ? a->cut : 0;
  const int T_out = a->T - cut - (a->smooth_mode == 2 ? ntaps - 1 : 0);
  const int Tp = b2t_output_frames(&e->cfg, a->T, a->smooth_mode, ntaps, cut);
  if (Tp < 1) return fail(B2T_ERR_ARG, "input too short: T=%d gives no output frame", a->T);
  const int Bp = r16(a->B);
  if ((H / 32) * (Bp / ((Bp % 64 == 0) ? 64 : (Bp % 32 == 0) ? 32 : 16)) > num_sms())
    return fail(B2T_ERR_UNSUPPORTED, "batch %d needs more co-resident CTAs than the %d SMs; split the batch", a->B, num_sms());
  if (!e->plans_ok || e->B != a->B || e->T_in != a->T || e->Tp != Tp || e->T_out != T_out || e->day_idx != a->day_idx) {
    e->B = a->B; e->Bpad = Bp; e->T_in = a->T; e->T_out = T_out; e->Tp = Tp; e->M = Tp * Bp; e->day_idx = a->day_idx;
    e->plans_ok = false;
    CK(cudaMemsetAsync(e->xd, 0, (size_t)Bp * a->T * D * sizeof(__nv_bfloat16), st));   // pad trials must stay finite
    if (build_plans(e)) return B2T_ERR_CUDA;
    e->plans_ok = true;
  }
  e->have_fwd = false; e->have_dlogits = false;
  e->fwd_training = a->training != 0;
  e->seed = a->seed;
  e->states_given = a->states != nullptr;
  const float keep_in = (a->training && e->cfg.input_dropout > 0.f) ? 1.0f - e->cfg.input_dropout : 1.0f;
  const float keep_rnn = (a->training && e->cfg.rnn_dropout > 0.f) ? 1.0f - e->cfg.rnn_dropout : 1.0f;
  const int BG = e->BG, nch = e->n_tchunks, NL = e->n_lanes;
  const int n_groups = Bp / BG, grid = (H / 32) * n_groups;

  // 1. augmentation + smoothing -> xs (bf16); 2. day layer; 3. layer-0 input projection (whole sequence)
  // Side work that does not depend on this step's input runs on the second bulk stream while augmentation and day layer run:
  // sentinels + initial states of the recurrence and (training) the gradient regions that backward accumulates into.
  {
    cudaStream_t ss = e->lane[MAX_LANES + 1];
    CK(cudaEventRecord(e->ev_top, st));
    CK(cudaStreamWaitEvent(ss, e->ev_top, 0));
    for (int l = 0; l < L; ++l) {
      // "not written yet" sentinel for the recurrence's data-as-signal exchange (gru_rec.cuh); slot 0 is the initial state
      CK(cudaMemsetAsync(e->lay[l].hseq + (size_t)Bp * H, 0xFF, (size_t)Tp * Bp * H * sizeof(__nv_bfloat16), ss));
      init_state_kernel<<<(Bp * H + 255) / 256, 256, 0, ss>>>(e->params + seg_off(e, "h0"), a->states ? a->states + (size_t)l * a->B * H : nullptr,
                                                               a->B, Bp, H, e->lay[l].hseq, e->lay[l].h_state);
      CK(LAUNCHED());
    }
    if (a->training) { const int rc = zero_accumulated_grads(e, ss); if (rc) return rc; }
    if (e->stack) {
      CK(cudaMemsetAsync(e->ctr, 0, (size_t)4 * L * e->ctr_stride * sizeof(int), ss));
      if (a->training)   // the dGh arrays double as the exchange medium of the backward stack kernel: "not written yet" sentinel
        for (int l = 0; l < L; ++l) {
          CK(cudaMemsetAsync(e->lay[l].dGh, 0xFF, (size_t)e->M * 3 * H * sizeof(__nv_bfloat16), ss));
          if (l < L - 1) CK(cudaMemsetAsync(e->lay[l].dY, 0xFF, (size_t)e->M * H * sizeof(float), ss));   // polled by layer l's backward recurrence while the gated GEMM of layer l+1 fills it
        }
    }
    CK(cudaEventRecord(e->ev_init, ss));
  }
  pp.x = a->x; pp.out = e->xs; pp.out_f32 = nullptr; pp.B = a->B; pp.Bpad = Bp; pp.T_in = a->T; pp.T_alloc = a->T; pp.D = D;
  pp.cut = cut; pp.ntaps = ntaps; pp.valid = a->smooth_mode == 2;
  pp.white_std = a->training ? a->white_noise_std : 0.f;
  pp.offset_std = a->training ? a->offset_noise_std : 0.f;
  pp.white = a->white_noise; pp.offset = a->offset_noise; pp.use_philox = 1;
  pp.seed = a->seed; pp.rng_offset = 0;
  {
    dim3 g((a->T + PRE_TT - 1) / PRE_TT, Bp);
    TlScope tl("pre", 9, st);
    pre_smooth_kernel<<<g, D / 4, 0, st>>>(pp);
    CK(LAUNCHED());
  }
  e->p_day.p.keep = keep_in; e->p_day.p.seed = a->seed; e->p_day.p.rng_offset = 0;
  { TlScope tl("day", 9, st); CK(gemm_run(e->p_day, st)); ++g_launches; }
  if (e->use_unfold_copy) {
    unfold_kernel<<<num_sms() * 4, 256, 0, st>>>(e->xd, e->xu, Bp, a->T, D, Tp, e->patch, e->stride);
    CK(LAUNCHED());
  }
  const bool save = a->training != 0;
  if (e->stack) {
    // ---- whole-stack schedule: layer-0 input projection on the whole chip, then ONE persistent recurrence kernel for all layers
    //      with the input projections of layers >= 1 as gated GEMMs on the SMs it leaves free (gru_stack.cuh)
    CK(cudaStreamWaitEvent(st, e->ev_init, 0));
    { TlScope tl("G0", 9, st); CK(gemm_run(e->p_in[0][0], st)); ++g_launches; }
    CK(cudaEventRecord(e->ev_start, st));
    cudaStream_t rs = e->lane[0];
    CK(cudaStreamWaitEvent(rs, e->ev_start, 0));
    StackFwdParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.H = H; sp.Bpad = Bp; sp.T = Tp; sp.n_slices = H / 32; sp.n_layers = L; sp.n_cgroups = e->stk_ncg;
    sp.poll_delay = env_int("B2T_STACK_POLL_DELAY", 600);   /* cycles between "own slice stored" and the first probe: the peers cannot be visible sooner than an L2 round trip, and early probes only add traffic (measured: 0 -> 2.947, 400-800 -> 2.92-2.93, 1200 -> 2.947 ms per step; the same delay in the backward kernel costs 0.1 ms) */ sp.seed = a->seed; sp.trace = e->trace; sp.trace_cta = env_int("B2T_TRACE_CTA_FWD", 0);
    sp.use_tma = e->stk_tma ? 1 : 0;
    for (int l = 0; l < L; ++l) sp.tm_h[l] = e->tm_hseq[l];
    for (int l = 0; l < L; ++l) {
      const std::string sl = std::to_string(l);
      StackFwdLayer& y = sp.lay[l];
      y.gx = e->lay[l].gx; y.bhh = e->params + seg_off(e, "gru.bias_hh_l" + sl); y.whh = e->shadow + seg_off(e, "gru.weight_hh_l" + sl);
      y.hseq = e->lay[l].hseq; y.h_state = e->lay[l].h_state;
      y.hdrop = save ? e->lay[l].hdrop : nullptr;
      y.R = save ? e->lay[l].R : nullptr; y.Z = save ? e->lay[l].Z : nullptr; y.Nn = save ? e->lay[l].Nn : nullptr; y.HN = save ? e->lay[l].HN : nullptr;
      y.keep = keep_rnn; y.rng_offset = (unsigned long long)(l + 1) << 40;
      y.kmask = (save && y.hdrop) ? e->lay[l].kmask : nullptr;
      if (!save && e->lay[l].hdrop) { y.hdrop = e->lay[l].hdrop; y.keep = 1.0f; }   // eval through a training engine: the next layer reads hdrop
      y.gx_done = l > 0 ? e->ctr + (size_t)(1 * L + l) * e->ctr_stride : nullptr;
      y.gx_need = 4 * e->p_in[l][0].p.tiles_n;
      y.prog = l < L - 1 ? e->ctr + (size_t)(0 * L + l) * e->ctr_stride : nullptr;
    }
    { TlScope tl("SR", 0, rs); CK(launch_stack_fwd(e->stk_BG, e->stk_NSUB, sp, e->stk_grid, rs)); }
    CK(cudaEventRecord(e->ev_lane_end[0], rs));
    for (int l = 1; l < L; ++l) {
      cudaStream_t gs = e->gstream[l];
      CK(cudaStreamWaitEvent(gs, e->ev_start, 0));
      { TlScope tl(("G" + std::to_string(l)).c_str(), l, gs); CK(gemm_run(e->p_in[l][0], gs)); ++g_launches; }
      CK(cudaEventRecord(e->ev_g[l], gs));
      CK(cudaStreamWaitEvent(st, e->ev_g[l], 0));
    }
    CK(cudaStreamWaitEvent(st, e->ev_lane_end[0], 0));
  } else {
  CK(cudaEventRecord(e->ev_start, st));
  for (int i = 0; i <= MAX_LANES; ++i) {
    CK(cudaStreamWaitEvent(e->lane[i], e->ev_start, 0));
    CK(cudaStreamWaitEvent(e->lane[i], e->ev_init, 0));
  }
  CK(cudaStreamWaitEvent(st, e->ev_init, 0));
  // layer-0 input projection: all chunks up front on the bulk stream (they depend on no recurrence)
  for (int c = 0; c < nch; ++c) {
    TlScope tl(("G0." + std::to_string(c)).c_str(), 8, e->lane[MAX_LANES]);
    CK(gemm_run(e->p_in[0][c], e->lane[MAX_LANES])); ++g_launches;
    CK(cudaEventRecord(e->ev_g0[c], e->lane[MAX_LANES]));
  }

  // 4. GRU stack: wave-front over (layer, chunk).  Task (l, c) = input projection of chunk c (l > 0: needs chunk c of the layer
  //    below) + recurrence over chunk c (needs chunk c-1 of the same layer).  Tasks are list-scheduled onto the lane that
  //    frees first (estimated durations), dependencies are CUDA events, issue order is diagonal by diagonal.
  double lane_free[MAX_LANES] = {};
  std::vector<double> t_end((size_t)L * MAX_CHUNKS, 0.0);
  const double dur_g = 40.0, dur_r = 135.0;
  for (int d = 0; d < nch + L - 1; ++d) {
    for (int l = 0; l < L; ++l) {
      const int c = d - l;
      if (c < 0 || c >= nch) continue;
      double ready = 0.0;
      if (l > 0) ready = std::max(ready, t_end[(size_t)(l - 1) * MAX_CHUNKS + c]);
      if (c > 0) ready = std::max(ready, t_end[(size_t)l * MAX_CHUNKS + c - 1]);
      int li = 0;
      for (int i = 1; i < NL; ++i)
        if (std::max(lane_free[i], ready) < std::max(lane_free[li], ready) - 1e-9) li = i;
      const double start = std::max(lane_free[li], ready);
      lane_free[li] = t_end[(size_t)l * MAX_CHUNKS + c] = start + (l == 0 ? 0.0 : dur_g) + dur_r;
      cudaStream_t ls = e->lane[li];
      if (l > 0) CK(cudaStreamWaitEvent(ls, e->ev_r[(size_t)(l - 1) * MAX_CHUNKS + c], 0));
      if (c > 0) CK(cudaStreamWaitEvent(ls, e->ev_r[(size_t)l * MAX_CHUNKS + c - 1], 0));
      if (l == 0) CK(cudaStreamWaitEvent(ls, e->ev_g0[c], 0));
      else { TlScope tl(("G" + std::to_string(l) + "." + std::to_string(c)).c_str(), li, ls); CK(gemm_run(e->p_in[l][c], ls)); ++g_launches; }
      RecFwdParams rp;
      rp.H = H; rp.Bpad = Bp; rp.t_begin = e->tc_begin[c]; rp.t_end = e->tc_begin[c + 1]; rp.T = Tp; rp.n_slices = H / 32;
      rp.gx = e->lay[l].gx; rp.bhh = e->params + seg_off(e, "gru.bias_hh_l" + std::to_string(l));
      rp.whh = e->shadow + seg_off(e, "gru.weight_hh_l" + std::to_string(l));
      rp.hseq = e->lay[l].hseq; rp.h_state = e->lay[l].h_state;
      rp.hdrop = save ? e->lay[l].hdrop : nullptr;
      rp.R = save ? e->lay[l].R : nullptr; rp.Z = save ? e->lay[l].Z : nullptr;
      rp.Nn = save ? e->lay[l].Nn : nullptr; rp.HN = save ? e->lay[l].HN : nullptr;
      rp.poll_delay = e->poll_delay; rp.keep = keep_rnn; rp.seed = a->seed; rp.rng_offset = (unsigned long long)(l + 1) << 40;
      rp.trace = (l == 0) ? e->trace : nullptr;
      // eval-mode forward through a training engine: the next layer reads hdrop if it exists, so keep it in sync
      if (!save && e->lay[l].hdrop) { rp.hdrop = e->lay[l].hdrop; rp.keep = 1.0f; }
      { TlScope tl(("R" + std::to_string(l) + "." + std::to_string(c)).c_str(), li, ls); CK(launch_rec_fwd(BG, rp, grid, ls)); }
      CK(cudaEventRecord(e->ev_r[(size_t)l * MAX_CHUNKS + c], ls));
    }
  }
  // join: the user stream continues after the top layer's last chunk (which transitively follows everything else)
  for (int i = 0; i <= MAX_LANES; ++i) {
    if (i < MAX_LANES && i >= NL) continue;
    CK(cudaEventRecord(e->ev_lane_end[i], e->lane[i]));
    CK(cudaStreamWaitEvent(st, e->ev_lane_end[i], 0));
  }
  }   // !stack
  // 5. head
  { TlScope tl("head", 9, st); CK(gemm_run(e->p_head, st)); ++g_launches; }
  if (a->logits_out) {
    gather_logits_kernel<<<num_sms() * 2, 256, 0, st>>>(e->logits, Tp, a->B, Bp, LDL, e->C, a->logits_out);
    CK(LAUNCHED());
  }
  if (a->hidden_out) {
    for (int l = 0; l < L; ++l)
      CK(cudaMemcpyAsync(a->hidden_out + (size_t)l * a->B * H, e->lay[l].h_state, (size_t)a->B * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  e->have_fwd = true;
  return Tp;
}

// ------------------------------------------------------------------------------------ CTC
template <int CPL>
static int run_ctc_warp(const CtcParams& cp, int B, size_t smem, cudaStream_t st) {
  CK(cudaFuncSetAttribute(ctc_loss_grad_warp_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_loss_grad_warp_kernel<CPL><<<B, 32, smem, st>>>(cp);
  CK(LAUNCHED());
  return 0;
}

template <int CPL>
static int run_ctc_par(const CtcParams& cp, int B, size_t smem, cudaStream_t st) {
  CK(cudaFuncSetAttribute(ctc_loss_grad_par_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_loss_grad_par_kernel<CPL><<<B, 32 * CTC_PAR_WARPS, smem, st>>>(cp);
  CK(LAUNCHED());
  return 0;
}

static int run_ctc(const CtcParams& cp, int B, cudaStream_t st) {
  if (cp.C > CTC_THREADS || cp.ldl > CTC_THREADS) return fail(B2T_ERR_UNSUPPORTED, "too many classes");
  const int Lmax = 2 * cp.Smax + 1;
  const size_t psmem = ctc_par_smem_bytes(cp.T, cp.C);
  static const bool force_warp = getenv("B2T_CTC_WARP") != nullptr;      // A/B switch for profiling
  if (Lmax <= 128 && cp.C <= 64 && psmem <= 200 * 1024 && cp.beta && !force_warp) {   // CTA per trial, alpha and beta concurrently
    const int cpl = (Lmax + 31) / 32;
    switch (cpl) {
      case 1: return run_ctc_par<1>(cp, B, psmem, st);
      case 2: return run_ctc_par<2>(cp, B, psmem, st);
      case 3: return run_ctc_par<3>(cp, B, psmem, st);
      default: return run_ctc_par<4>(cp, B, psmem, st);
    }
  }
  const size_t wsmem = ctc_warp_smem_bytes(cp.T, cp.C);
  if (Lmax <= 128 && cp.C <= 64 && wsmem <= 200 * 1024) {     // warp-per-trial kernel: no block barriers
    const int cpl = (Lmax + 31) / 32;
    switch (cpl) {
      case 1: return run_ctc_warp<1>(cp, B, wsmem, st);
      case 2: return run_ctc_warp<2>(cp, B, wsmem, st);
      case 3: return run_ctc_warp<3>(cp, B, wsmem, st);
      default: return run_ctc_warp<4>(cp, B, wsmem, st);
    }
  }
  const size_t smem = ctc_smem_bytes(cp.T, cp.C, cp.Smax);
  if (smem > 200 * 1024) return fail(B2T_ERR_UNSUPPORTED, "CTC problem too large for shared memory (T=%d, S=%d)", cp.T, cp.Smax);
  CK(cudaFuncSetAttribute(ctc_loss_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_loss_grad_kernel<<<B, CTC_THREADS, smem, st>>>(cp);
  CK(LAUNCHED());
  return 0;
}

extern "C" int b2t_ctc_loss(b2t_engine* e, const int* labels, int Smax, const int* in_len, const int* tgt_len, float grad_scale,
                            float* loss_out, int want_grad, void* stream) {
  if (!e || !labels || !in_len || !tgt_len || !loss_out) return fail(B2T_ERR_ARG, "null argument");
  if (!e->have_fwd) return fail(B2T_ERR_STATE, "b2t_ctc_loss needs a preceding b2t_forward");
  if (Smax < 1 || Smax > e->maxS) return fail(B2T_ERR_ARG, "Smax=%d exceeds engine limit %d", Smax, e->maxS);
  cudaStream_t st = (cudaStream_t)stream;
  CtcParams cp;
  cp.logits = e->logits; cp.ldl = LDL; cp.Bpad = e->Bpad; cp.T = e->Tp; cp.C = e->C; cp.blank = 0;
  cp.labels = labels; cp.Smax = Smax; cp.in_len = in_len; cp.tgt_len = tgt_len;
  cp.alpha = e->alpha; cp.beta = e->beta; cp.loss = loss_out;
  cp.dlogits = want_grad ? e->dlog32 : nullptr;
  cp.dlogits_bf16 = want_grad ? e->dlog16 : nullptr;
  cp.dbias = nullptr;
  cp.grad_scale = grad_scale;
  if (want_grad) {
    if (!e->training) return fail(B2T_ERR_STATE, "inference engine cannot keep gradients");
    // pad trials (b >= B) are never visited by the kernel: clear them once
    CK(cudaMemsetAsync(e->dlog16, 0, (size_t)e->M * LDL * sizeof(__nv_bfloat16), st));
    CK(cudaMemsetAsync(e->dlog32, 0, (size_t)e->M * LDL * sizeof(float), st));
  }
  int rc;
  { TlScope tl("ctc", 9, st); rc = run_ctc(cp, e->B, st); }
  if (rc) return rc;
  if (want_grad) e->have_dlogits = true;
  return 0;
}

extern "C" long long b2t_ctc_workspace_bytes(int T, int B, int Smax) { return (long long)2 * T * B * (2 * Smax + 1) * 4 + 1024; }

extern "C" int b2t_ctc_loss_tbc(const float* logits_tbc, int T, int B, int C, const int* labels, int Smax, const int* in_len, const int* tgt_len,
                                float grad_scale, float* loss_out, float* dlogits_tbc, void* workspace, long long workspace_bytes, void* stream) {
  if (!logits_tbc || !labels || !in_len || !tgt_len || !loss_out || !workspace) return fail(B2T_ERR_ARG, "null argument");
  if (workspace_bytes < b2t_ctc_workspace_bytes(T, B, Smax)) return fail(B2T_ERR_WORKSPACE, "CTC workspace too small");
  CtcParams cp;
  cp.logits = logits_tbc; cp.ldl = C; cp.Bpad = B; cp.T = T; cp.C = C; cp.blank = 0;
  cp.labels = labels; cp.Smax = Smax; cp.in_len = in_len; cp.tgt_len = tgt_len;
  cp.alpha = reinterpret_cast<float*>(workspace); cp.beta = cp.alpha + (size_t)T * B * (2 * Smax + 1);
  cp.loss = loss_out; cp.dlogits = dlogits_tbc; cp.dlogits_bf16 = nullptr; cp.dbias = nullptr;
  cp.grad_scale = grad_scale;
  return run_ctc(cp, B, (cudaStream_t)stream);
}

extern "C" int b2t_set_dlogits(b2t_engine* e, const float* dlogits, void* stream) {
  if (!e || !dlogits) return fail(B2T_ERR_ARG, "null argument");
  if (!e->have_fwd || !e->training) return fail(B2T_ERR_STATE, "no training forward to attach gradients to");
  cudaStream_t st = (cudaStream_t)stream;
  scatter_dlogits_kernel<<<num_sms() * 2, 256, 0, st>>>(dlogits, e->Tp, e->B, e->Bpad, LDL, e->C, e->dlog32, e->dlog16);
  CK(LAUNCHED());
  e->have_dlogits = true;
  return 0;
}

// ------------------------------------------------------------------------------------ backward
extern "C" int b2t_backward(b2t_engine* e, void* stream) {
  if (!e) return fail(B2T_ERR_ARG, "null engine");
  if (!e->training || !e->have_fwd || !e->fwd_training) return fail(B2T_ERR_STATE, "b2t_backward needs a training-mode b2t_forward");
  if (!e->have_dlogits) return fail(B2T_ERR_STATE, "no dlogits: call b2t_ctc_loss(want_grad) or b2t_set_dlogits first");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = e->D, H = e->H, L = e->L, Bp = e->Bpad, Tp = e->Tp;
  const float keep_in = e->cfg.input_dropout > 0.f ? 1.0f - e->cfg.input_dropout : 1.0f;
  const float keep_rnn = e->cfg.rnn_dropout > 0.f ? 1.0f - e->cfg.rnn_dropout : 1.0f;
  const int BG = e->BGb, nch = e->n_tchunks, NL = e->n_lanes_b;
  const int n_groups = Bp / BG, grid = (H / 32) * n_groups / e->NSUBb;
  // (the regions that are accumulated with atomics were cleared on a side stream during forward: zero_accumulated_grads)
  CK(cudaMemsetAsync(e->touched, 0, r64(e->cfg.n_days) * sizeof(float), st));
  mark_days_kernel<<<(e->B + 127) / 128, 128, 0, st>>>(e->day_idx, e->B, e->touched);
  CK(LAUNCHED());
  if (e->bwd2 && !e->stack) {   // the dGh arrays double as the exchange medium of gru_rec_bwd2_kernel: "not written yet" sentinel
    for (int l = 0; l < L; ++l) CK(cudaMemsetAsync(e->lay[l].dGh, 0xFF, (size_t)e->M * 3 * H * sizeof(__nv_bfloat16), st));
  }
  {
    const int geom = e->stack ? 1000000 + (Bp * 64 + e->stk_BG) * 2 : (Bp * 64 + BG) * 2 + e->bwd2;
    if (e->part_geom != geom) {   // partial-exchange buffers: (re)start the generation tags (gru_rec.cuh) for this geometry
      for (int l = 0; l < L; ++l) {
        CK(cudaMemsetAsync(e->lay[l].part, 0xFF, (size_t)2 * Bp * (H / 32) * (H / 32) * 32 * sizeof(float), st));
        e->lay[l].gen = 0;
      }
      e->part_geom = geom;
    }
  }

  // head.  Only dY of the top layer is on the way to the backward recurrence; in the stack schedule the head's own gradients
  // (out.bias column sums, dW_out: six tiles) are issued behind it and run beside the recurrence's first steps.
  auto head_grads = [&]() -> int {
    colsum_kernel<<<256, 64, 0, st>>>(e->dlog32, (size_t)e->M, LDL, e->C, e->grads + seg_off(e, "out.bias"));
    CK(LAUNCHED());
    { TlScope tl("dWout", 9, st); CK(gemm_run(e->p_dwout, st)); ++g_launches; }
    return 0;
  };
  if (!e->stack) { if (int rc = head_grads()) return rc; }
  { TlScope tl("dYtop", 9, st); CK(gemm_run(e->p_dytop, st)); ++g_launches; }
  CK(cudaEventRecord(e->ev_top, st));
  if (e->stack) {
    if (int rc = head_grads()) return rc;
    CK(cudaEventRecord(e->ev_head, st));                      // (head gradients final: awaited before the head bucket's event below)
    // ---- whole-stack schedule: ONE persistent backward recurrence for all layers; the data-gradient GEMMs dY_{l-1} = dGx_l W_ih_l
    //      follow it as gated GEMMs (time descending) on the free SMs; weight gradients, layer-0 data gradient, fold and day layer after it
    cudaStream_t rs = e->lane[0], bs = e->lane[MAX_LANES], bw = e->lane[MAX_LANES + 1];
    CK(cudaStreamWaitEvent(rs, e->ev_top, 0));
    StackBwdParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.H = H; sp.Bpad = Bp; sp.T = Tp; sp.n_layers = L; sp.n_cgroups = e->stk_ncg; sp.n_valid = e->B;
    sp.poll_delay = env_int("B2T_STACK_POLL_DELAY_BWD", 0); sp.seed = e->seed; sp.trace = e->trace; sp.trace_cta = env_int("B2T_TRACE_CTA_BWD", (L - 1) * (H / 32) * e->stk_ncg);
    for (int l = 0; l < L; ++l) {
      const std::string sl = std::to_string(l);
      StackBwdLayer& y = sp.lay[l];
      y.dY = e->lay[l].dY; y.hseq = e->lay[l].hseq; y.R = e->lay[l].R; y.Z = e->lay[l].Z; y.Nn = e->lay[l].Nn; y.HN = e->lay[l].HN;
      y.whh = e->shadow + seg_off(e, "gru.weight_hh_l" + sl);
      y.dGx = e->lay[l].dGx; y.dGh = e->lay[l].dGh; y.part = e->lay[l].part;
      y.dbih = e->grads + seg_off(e, "gru.bias_ih_l" + sl); y.dbhh = e->grads + seg_off(e, "gru.bias_hh_l" + sl);
      y.dh_state = e->lay[l].dh_state;
      y.dy_polled = l < L - 1 ? 1 : 0;
      y.prog = l > 0 ? e->ctr + (size_t)(2 * L + l) * e->ctr_stride : nullptr;
      y.gen_base = e->lay[l].gen; e->lay[l].gen = (e->lay[l].gen + Tp) & 7;
      y.keep = (l < L - 1) ? keep_rnn : 1.0f;
      y.kmask = (l < L - 1 && env_int("B2T_BWD_KMASK", 1)) ? e->lay[l].kmask : nullptr;
      y.rng_offset = (unsigned long long)(l + 1) << 40;
    }
    { TlScope tl("SRB", 0, rs); CK(launch_stack_bwd(e->stk_BG, e->stk_NSUB, sp, e->stk_grid, rs)); }
    CK(cudaEventRecord(e->ev_r[0], rs));
    // (nothing else may be queued behind the recurrence on its stream before the gated GEMMs below are launched: streams share
    //  hardware queues, and work stuck behind the recurrence would hold back the GEMMs the recurrence is waiting for)
    for (int l = L - 1; l >= 1; --l) {
      cudaStream_t gs = e->gstream[l];
      CK(cudaStreamWaitEvent(gs, e->ev_top, 0));
      { TlScope tl(("DX" + std::to_string(l)).c_str(), l, gs); CK(gemm_run(e->p_dx[l][0], gs)); ++g_launches; }
      CK(cudaEventRecord(e->ev_g[l], gs));
      CK(cudaStreamWaitEvent(st, e->ev_g[l], 0));
    }
    CK(cudaStreamWaitEvent(bs, e->ev_r[0], 0));
    CK(cudaStreamWaitEvent(bw, e->ev_r[0], 0));
    if (!e->states_given) {   // gradient of the learned initial state: five tiny kernels, first on the weight-gradient stream so that its end is a GEMM
      for (int l = 0; l < L; ++l) {
        reduce_dh0_kernel<<<(H + 127) / 128, 128, 0, bw>>>(e->lay[l].dh_state, e->B, H, e->grads + seg_off(e, "h0"));
        CK(LAUNCHED());
      }
    }
    // chain on the first bulk stream: layer-0 data gradient -> patch fold -> day layer; everything else on the second one
    { TlScope tl("DX0", 8, bs); CK(gemm_run(e->p_dx[0][0], bs)); ++g_launches; }
    {
      FoldParams fp;
      fp.dxu = e->dxu; fp.xd = e->xd; fp.dpre = e->dpre; fp.dbias_day = e->grads + seg_off(e, "day_biases.0"); fp.bias_pitch = (int)r64(D);
      fp.day_idx = e->day_idx; fp.B = e->B; fp.Bpad = Bp; fp.T_alloc = e->T_in; fp.T_valid = e->T_out; fp.D = D; fp.Tp = Tp;
      fp.patch = e->patch; fp.stride = e->stride; fp.keep = keep_in; fp.seed = e->seed; fp.rng_offset = 0;
      dim3 g((e->T_in + FOLD_TT - 1) / FOLD_TT, Bp);
      { TlScope tl("fold", 8, bs); fold_dpre_kernel<<<g, D / 4, 0, bs>>>(fp); CK(LAUNCHED()); }
      { TlScope tl("daydW", 8, bs); CK(gemm_run(e->p_daydw, bs)); ++g_launches; }
      CK(cudaEventRecord(e->ev_bucket[0], bs));                 // day layers final
    }
    e->bucket_order.clear();
    // Weight gradients.  Order = the order in which the gradient buckets become final, chosen so that the LAST one is small (its
    // all-reduce cannot overlap anything): all upper-layer weights in two batched launches, then layer 0's input weights in two halves.
    auto upper_layers = [&]() -> int {
      if (e->dw_batched) {
        { TlScope tl("dWihB", 7, bw); CK(gemm_run(e->p_dwih_b, bw)); ++g_launches; }
        { TlScope tl("dWhhB", 7, bw); CK(gemm_run(e->p_dwhh_b, bw)); ++g_launches; }
      }
      for (int l = L - 1; l >= 0; --l) {
        const std::string sl = std::to_string(l);
        if (!e->dw_batched) {
          if (l > 0) { TlScope tl(("dWih" + sl).c_str(), 7, bw); CK(gemm_run(e->p_dwih[l], bw)); ++g_launches; }
          { TlScope tl(("dWhh" + sl).c_str(), 7, bw); CK(gemm_run(e->p_dwhh[l], bw)); ++g_launches; }
        }
        CK(cudaEventRecord(e->ev_bucket[2 + l], bw)); e->bucket_order.push_back(2 + l);    // (bias gradients were final when the recurrence ended)
      }
      CK(cudaStreamWaitEvent(bw, e->ev_head, 0));                                           // head gradients (issued beside the recurrence)
      CK(cudaEventRecord(e->ev_bucket[L + 2], bw)); e->bucket_order.push_back(L + 2);      // head, h0 (all layers), touched flags
      return 0;
    };
    if (e->dwih0_split && e->dw_batched && env_int("B2T_TAIL_SPLIT", 0)) {   // (measured on 2 GPUs: 3.77 ms against 3.74 ms for the plain order below -- the collective and the GEMMs share HBM, a finer overlap does not pay)
      if (int rc = upper_layers()) return rc;
      { TlScope tl("dWih0a", 7, bw); CK(gemm_run(e->p_dwih0_h[0], bw)); ++g_launches; }
      CK(cudaEventRecord(e->ev_bucket[1], bw)); e->bucket_order.push_back(1);
      e->bucket_order.push_back(0);                             // day layers: final when the DX0 -> fold -> day-dW chain ends
      { TlScope tl("dWih0b", 7, bw); CK(gemm_run(e->p_dwih0_h[1], bw)); ++g_launches; }
      CK(cudaEventRecord(e->ev_bucket[L + 3], bw)); e->bucket_order.push_back(L + 3);
    } else {
      { TlScope tl("dWih0", 7, bw); CK(gemm_run(e->p_dwih0[0], bw)); ++g_launches; }
      CK(cudaEventRecord(e->ev_bucket[1], bw)); e->bucket_order.push_back(1);
      CK(cudaEventRecord(e->ev_bucket[L + 3], bw)); e->bucket_order.push_back(L + 3);
      e->bucket_order.push_back(0);
      if (int rc = upper_layers()) return rc;
    }
    for (int i = MAX_LANES; i <= MAX_LANES + 1; ++i) {
      CK(cudaEventRecord(e->ev_lane_end[i], e->lane[i]));
      CK(cudaStreamWaitEvent(st, e->ev_lane_end[i], 0));
    }
    e->have_dlogits = false;
    return 0;
  }
  for (int i = 0; i < NL; ++i) CK(cudaStreamWaitEvent(e->lane[i], e->ev_top, 0));
  CK(cudaStreamWaitEvent(e->lane[MAX_LANES], e->ev_top, 0));
  CK(cudaStreamWaitEvent(e->lane[MAX_LANES + 1], e->ev_top, 0));

  // wave-front over (layer descending, time chunk descending); k counts chunks from the end of the sequence.
  // Task (l, c) = recurrence over chunk c (needs chunk c+1 of the same layer and dY_l[chunk c] from the layer above)
  // followed by the data-gradient GEMM for the layer below.  Same list scheduling as in forward.
  double lane_free[MAX_LANES] = {};
  std::vector<double> t_end((size_t)L * MAX_CHUNKS, 0.0);
  const double dur_rb = 215.0, dur_dx = 25.0;
  for (int d = 0; d < nch + L - 1; ++d) {
    for (int l = L - 1; l >= 0; --l) {
      const int k = d - (L - 1 - l);
      if (k < 0 || k >= nch) continue;
      const int c = nch - 1 - k;
      const std::string sl = std::to_string(l);
      double ready = 0.0;
      if (l < L - 1) ready = std::max(ready, t_end[(size_t)(l + 1) * MAX_CHUNKS + c]);
      if (c < nch - 1) ready = std::max(ready, t_end[(size_t)l * MAX_CHUNKS + c + 1]);
      int li = 0;
      for (int i = 1; i < NL; ++i)
        if (std::max(lane_free[i], ready) < std::max(lane_free[li], ready) - 1e-9) li = i;
      lane_free[li] = t_end[(size_t)l * MAX_CHUNKS + c] = std::max(lane_free[li], ready) + dur_rb + (l > 0 ? dur_dx : 0.0);
      cudaStream_t ls = e->lane[li];
      if (l < L - 1) CK(cudaStreamWaitEvent(ls, e->ev_dx[(size_t)(l + 1) * MAX_CHUNKS + c], 0));   // dY_l[chunk c] is ready
      if (c < nch - 1) CK(cudaStreamWaitEvent(ls, e->ev_dx[(size_t)l * MAX_CHUNKS + c + 1], 0));    // chunk c+1 of this layer is done
      RecBwdParams bp;
      bp.H = H; bp.Bpad = Bp; bp.n_slices = H / 32; bp.T = Tp;
      bp.t_begin = e->tc_begin[c]; bp.t_end = e->tc_begin[c + 1]; bp.first_chunk = (c == nch - 1);
      bp.dY = e->lay[l].dY;
      bp.hseq = e->lay[l].hseq; bp.R = e->lay[l].R; bp.Z = e->lay[l].Z; bp.Nn = e->lay[l].Nn; bp.HN = e->lay[l].HN;
      bp.whh = e->shadow + seg_off(e, "gru.weight_hh_l" + sl);
      bp.dGx = e->lay[l].dGx; bp.dGh = e->lay[l].dGh; bp.part = e->lay[l].part;
      bp.dbih = e->grads + seg_off(e, "gru.bias_ih_l" + sl); bp.dbhh = e->grads + seg_off(e, "gru.bias_hh_l" + sl);
      bp.dh_state = e->lay[l].dh_state;
      bp.n_valid = e->B;
      bp.gen_base = e->lay[l].gen; e->lay[l].gen = (e->lay[l].gen + (bp.t_end - bp.t_begin)) & 7;
      bp.keep = (l < L - 1) ? keep_rnn : 1.0f;
      bp.seed = e->seed; bp.rng_offset = (unsigned long long)(l + 1) << 40;
      bp.trace = (l == L - 1 && c == nch - 1 && e->trace) ? e->trace + (size_t)Tp * 8 : nullptr;
      bp.poll_delay = e->poll_delay_b;
      { TlScope tl(("RB" + sl + "." + std::to_string(c)).c_str(), li, ls); CK(e->bwd2 ? launch_rec_bwd2(BG, bp, grid, ls) : launch_rec_bwd(BG, e->NSUBb, bp, grid, ls)); }
      if (l > 0) { TlScope tl(("DX" + sl + "." + std::to_string(c)).c_str(), li, ls); CK(gemm_run(e->p_dx[l][c], ls)); ++g_launches; }
      CK(cudaEventRecord(e->ev_dx[(size_t)l * MAX_CHUNKS + c], ls));
      if (l == 0) {   // layer 0: data gradient (to be folded) and weight gradient of this chunk fill idle SMs on the bulk stream
        // the data gradient heads the chain dX -> fold -> day layer (first bulk stream); the weight gradient is off that chain
        cudaStream_t bs = e->lane[MAX_LANES], bw = e->lane[MAX_LANES + 1];
        CK(cudaStreamWaitEvent(bs, e->ev_dx[(size_t)c], 0));
        CK(cudaStreamWaitEvent(bw, e->ev_dx[(size_t)c], 0));
        { TlScope tl(("DX0." + std::to_string(c)).c_str(), 8, bs); CK(gemm_run(e->p_dx[0][c], bs)); ++g_launches; }
        { TlScope tl(("dWih0." + std::to_string(c)).c_str(), 7, bw); CK(gemm_run(e->p_dwih0[c], bw)); ++g_launches; }
      }
      if (c == 0) {   // the layer's recurrence is complete: weight gradients over the whole sequence, on the bulk stream
        cudaStream_t bs = e->lane[MAX_LANES + (l > 0 ? 1 : 0)];   // layer 0 ends the chain dX -> fold -> day layer: keep it in order on the first bulk stream
        cudaStream_t bw = e->lane[MAX_LANES + 1];                // weight gradients of the hidden-to-hidden matrices and h0
        CK(cudaEventRecord(e->ev_r[(size_t)l * MAX_CHUNKS], ls));
        if (l == 0) CK(cudaStreamWaitEvent(bw, e->ev_r[(size_t)l * MAX_CHUNKS], 0));
        CK(cudaStreamWaitEvent(bs, e->ev_r[(size_t)l * MAX_CHUNKS], 0));
        if (l > 0) { TlScope tl(("dWih" + sl).c_str(), 7, bs); CK(gemm_run(e->p_dwih[l], bs)); ++g_launches; }
        { TlScope tl(("dWhh" + sl).c_str(), 7, bw); CK(gemm_run(e->p_dwhh[l], bw)); ++g_launches; }
        if (!e->states_given) {
          reduce_dh0_kernel<<<(H + 127) / 128, 128, 0, bw>>>(e->lay[l].dh_state, e->B, H, e->grads + seg_off(e, "h0"));
          CK(LAUNCHED());
        }
        if (l == 0) {   // patch fold + day layer
          FoldParams fp;
          fp.dxu = e->dxu; fp.xd = e->xd; fp.dpre = e->dpre; fp.dbias_day = e->grads + seg_off(e, "day_biases.0"); fp.bias_pitch = (int)r64(D);
          fp.day_idx = e->day_idx; fp.B = e->B; fp.Bpad = Bp; fp.T_alloc = e->T_in; fp.T_valid = e->T_out; fp.D = D; fp.Tp = Tp;
          fp.patch = e->patch; fp.stride = e->stride; fp.keep = keep_in; fp.seed = e->seed; fp.rng_offset = 0;
          dim3 g((e->T_in + FOLD_TT - 1) / FOLD_TT, Bp);
          { TlScope tl("fold", 8, bs); fold_dpre_kernel<<<g, D / 4, 0, bs>>>(fp); CK(LAUNCHED()); }
          { TlScope tl("daydW", 8, bs); CK(gemm_run(e->p_daydw, bs)); ++g_launches; }
        }
      }
    }
  }
  for (int i = 0; i <= MAX_LANES + 1; ++i) {
    if (i < MAX_LANES && i >= NL) continue;
    CK(cudaEventRecord(e->ev_lane_end[i], e->lane[i]));
    CK(cudaStreamWaitEvent(st, e->ev_lane_end[i], 0));
  }
  e->bucket_order.clear();
  for (size_t k = 0; k < e->ev_bucket.size(); ++k) { CK(cudaEventRecord(e->ev_bucket[k], st)); e->bucket_order.push_back((int)k); }
  e->have_dlogits = false;
  return 0;
}

// ------------------------------------------------------------------------------------ gradient buckets (data parallelism)
extern "C" int b2t_grad_buckets(b2t_engine* e) { return e ? (int)e->bucket_range.size() : fail(B2T_ERR_ARG, "null engine"); }
// The i-th bucket to become final in the last b2t_backward: its element range in the gradient buffer.
extern "C" int b2t_grad_bucket(b2t_engine* e, int i, long long* offset, long long* count) {
  if (!e || i < 0 || i >= (int)e->bucket_order.size()) return fail(B2T_ERR_ARG, "bucket index out of range (call after b2t_backward)");
  const int k = e->bucket_order[i];
  if (offset) *offset = e->bucket_range[k].first;
  if (count) *count = e->bucket_range[k].second;
  return k;
}
// Make `stream` wait until the i-th bucket (same numbering) of the last b2t_backward is final.
extern "C" int b2t_grad_bucket_wait(b2t_engine* e, int i, void* stream) {
  if (!e || i < 0 || i >= (int)e->bucket_order.size()) return fail(B2T_ERR_ARG, "bucket index out of range (call after b2t_backward)");
  CK(cudaStreamWaitEvent((cudaStream_t)stream, e->ev_bucket[e->bucket_order[i]], 0));
  return 0;
}

// ------------------------------------------------------------------------------------ optimizer
extern "C" int b2t_optimizer_step(b2t_engine* e, const b2t_adamw_args* a, float* stats_out, void* stream) {
  if (!e || !a) return fail(B2T_ERR_ARG, "null argument");
  if (!e->training) return fail(B2T_ERR_STATE, "inference engine has no optimizer state");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(e->sumsq, 0, sizeof(float), st));
  sumsq_kernel<<<num_sms() * 4, 256, 0, st>>>(e->grads, (size_t)e->n_params, e->sumsq);
  CK(LAUNCHED());
  AdamParams ap;
  ap.p = e->params; ap.g = e->grads; ap.m = e->m1; ap.v = e->m2; ap.shadow = e->shadow;
  ap.segs = e->d_segs; ap.chunks = e->d_chunks; ap.step = e->steps; ap.day_touched = e->touched;
  ap.sumsq = e->sumsq; ap.stats = e->stats; ap.max_norm = a->max_grad_norm;
  for (int i = 0; i < 3; ++i) { ap.lr[i] = a->lr[i]; ap.wd[i] = a->weight_decay[i]; }
  ap.beta1 = a->beta1; ap.beta2 = a->beta2; ap.eps = a->eps;
  TlScope tl_adam("adamw", 9, st);
  clip_adamw_kernel<<<e->n_chunks, 256, 0, st>>>(ap);
  CK(LAUNCHED());
  bump_steps_kernel<<<((int)e->segs.size() + 127) / 128, 128, 0, st>>>(e->d_segs, (int)e->segs.size(), e->touched, e->steps, e->sumsq, a->max_grad_norm);
  CK(LAUNCHED());
  if (stats_out) CK(cudaMemcpyAsync(stats_out, e->stats, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ------------------------------------------------------------------------------------ greedy decode
extern "C" int b2t_greedy_edit(b2t_engine* e, const int* labels, int Smax, const int* in_len, const int* tgt_len, int* decoded, int* dec_len,
                               int* edit, void* stream) {
  if (!e || !labels || !in_len || !tgt_len || !decoded || !dec_len || !edit) return fail(B2T_ERR_ARG, "null argument");
  if (!e->have_fwd) return fail(B2T_ERR_STATE, "needs a preceding b2t_forward");
  if (Smax < 1 || Smax > e->maxS) return fail(B2T_ERR_ARG, "Smax out of range");
  GreedyParams gp;
  gp.logits = e->logits; gp.T = e->Tp; gp.Bpad = e->Bpad; gp.ldl = LDL; gp.C = e->C;
  gp.in_len = in_len; gp.labels = labels; gp.tgt_len = tgt_len; gp.Smax = Smax;
  gp.decoded = decoded; gp.dec_len = dec_len; gp.edit = edit; gp.scratch = e->greedy_scratch;
  greedy_edit_kernel<<<e->B, 128, e->Tp * sizeof(int), (cudaStream_t)stream>>>(gp);
  CK(LAUNCHED());
  return 0;
}

// ------------------------------------------------------------------------------------ stand-alone smoothing
extern "C" int b2t_gauss_smooth(const float* x, int B, int T, int D, float std, int size, int mode, float* out, void* stream) {
  if (!x || !out || B < 1 || T < 1 || D < 4 || D % 4 != 0 || D > 4096) return fail(B2T_ERR_ARG, "bad gauss_smooth arguments");
  if (mode != 1 && mode != 2) return fail(B2T_ERR_ARG, "mode must be 1 ('same') or 2 ('valid')");
  PreParams pp;
  memset(&pp, 0, sizeof(pp));
  int ntaps = 0;
  if (gauss_taps(std, size, pp.taps, &ntaps)) return fail(B2T_ERR_UNSUPPORTED, "smoothing kernel has more than 16 taps");
  const int T_out = mode == 2 ? T - ntaps + 1 : T;
  if (T_out < 1) return fail(B2T_ERR_ARG, "input shorter than the smoothing kernel");
  pp.x = x; pp.out = nullptr; pp.out_f32 = out; pp.B = B; pp.Bpad = B; pp.T_in = T; pp.T_alloc = T_out; pp.D = D;
  pp.ntaps = ntaps; pp.valid = mode == 2;
  dim3 grid((T_out + PRE_TT - 1) / PRE_TT, B);
  pre_smooth_kernel<<<grid, D / 4, 0, (cudaStream_t)stream>>>(pp);
  CK(LAUNCHED());
  return T_out;
}

// ------------------------------------------------------------------------------------ GEMM test hook
extern "C" int b2t_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int a_mn, int b_mn, int out_bf16, const float* bias,
                             void* stream) {
  GemmSpec s;
  s.a_mn = a_mn; s.b_mn = b_mn; s.epi = EPI_STORE; s.out_bf16 = out_bf16;
  s.M = M; s.N = N; s.K = K;
  s.A = A; s.lda = a_mn ? M : K;
  s.B = B; s.ldb = b_mn ? N : K;
  s.C = C; s.ldc = N; s.bias = bias;
  // bring-up knobs of the test hook: cap the grid like a gated GEMM beside the recurrence, and exercise the completion counters
  s.max_ctas = env_int("B2T_GEMM_MAXCTAS", 0);
  static int* dbg_done = nullptr;
  if (env_int("B2T_GEMM_DONE", 0) && !a_mn) {
    if (!dbg_done) { cudaMalloc(&dbg_done, 1 << 20); cudaMemset(dbg_done, 0, 1 << 20); }
    s.done = dbg_done;
  }
  GemmPlan pl;
  int rc = gemm_plan_build(&pl, s);
  if (rc) return fail(B2T_ERR_CUDA, "gemm plan failed (%d)", rc);
  CK(gemm_run(pl, (cudaStream_t)stream));
  ++g_launches;
  return 0;
}
