// LM-free CTC prefix beam search (the searcher BrainSpeechDecoder builds when no FST is given).
//
// Reference: language_model/runtime/core/decoder/ctc_prefix_beam_search.cc:44-136, PrefixScore .h:27-42,
// LogAdd utils/utils.cc:24-30.  Golden vector: ctc_prefix_beam_search_test.cc:18-59.
//
// The per-frame update is a chain of order-dependent float LogAdd merges over at most
// first_beam x second_beam (default 10 x 10) candidates, so there is no useful parallelism inside one utterance;
// the kernel therefore runs one utterance per thread and gets its throughput from decoding many utterances
// at once (the pipeline is embarrassingly parallel over trials).  Prefixes are interned as (parent, token) trie
// nodes so that prefix identity is an integer comparison.
#include <math.h>
#include <stdio.h>

#include <vector>

#include <cuda_runtime.h>

#include "../../include/b2t_b200.h"

namespace {

constexpr int PB_MAX_BEAM = 64;       // second_beam_size limit
constexpr int PB_MAX_TOPK = 64;       // first_beam_size limit
constexpr int PB_MAX_CAND = 704;      // >= second_beam * (first_beam + 1) (checked by the host entry point)
#define PB_NEG (-3.402823466e+38f)    // -kFloatMax

struct Hyp;
struct PbParams {
  const float* logp;   // [N][T][C]
  const int* lens;     // [N]
  int N, T, C, blank, first_beam, second_beam, max_len;
  // per-utterance scratch
  int* trie_parent; int* trie_token; int trie_cap;     // [N][trie_cap]
  int* times;          // [N][2 buffers][PB_MAX_CAND][2 (s, ns)][max_len]
  struct Hyp* hyps;    // [N][PB_MAX_BEAM + PB_MAX_CAND]
  // outputs
  int* out_ids; int* out_len; float* out_score; float* out_viterbi; int* out_times; int* out_n;
  int* status;
};

struct Hyp {
  int node, len, last;              // trie node of the prefix, its length, its last token (-1 when empty)
  float s, ns, v_s, v_ns, cur_token_prob;
};

__device__ inline float log_add(float x, float y) {
  if (x <= PB_NEG) return y;
  if (y <= PB_NEG) return x;
  const float m = fmaxf(x, y);
  return logf(expf(x - m) + expf(y - m)) + m;
}
__device__ inline float hyp_score(const Hyp& h) { return log_add(h.s, h.ns); }
__device__ inline float hyp_viterbi(const Hyp& h) { return h.v_s > h.v_ns ? h.v_s : h.v_ns; }

__global__ void prefix_beam_kernel(const PbParams p) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= p.N) return;
  const float* logp = p.logp + (size_t)u * p.T * p.C;
  const int Tn = min(max(p.lens[u], 0), p.T);
  int* tpar = p.trie_parent + (size_t)u * p.trie_cap;
  int* ttok = p.trie_token + (size_t)u * p.trie_cap;
  int ntrie = 1;                                        // node 0 = empty prefix
  tpar[0] = -1; ttok[0] = -1;
  const size_t tstride = (size_t)2 * p.max_len;          // per hypothesis: times_s | times_ns
  int* tbuf[2] = {p.times + (size_t)u * 2 * PB_MAX_CAND * tstride, p.times + ((size_t)u * 2 + 1) * PB_MAX_CAND * tstride};

  Hyp* cur = p.hyps + (size_t)u * (PB_MAX_BEAM + PB_MAX_CAND);
  Hyp* nxt = cur + PB_MAX_BEAM;
  int ncur = 1, cb = 0;
  cur[0] = Hyp{0, 0, -1, 0.0f, PB_NEG, 0.0f, 0.0f, PB_NEG};
  const int k = min(min(p.first_beam, p.C), PB_MAX_TOPK);
  bool overflow = false;

  for (int t = 0; t < Tn; ++t) {
    const float* row = logp + (size_t)t * p.C;
    // 1. first beam: top-k classes, descending value, lower index first among ties
    int top[PB_MAX_TOPK];
    for (int i = 0; i < k; ++i) {
      int best = -1;
      for (int c = 0; c < p.C; ++c) {
        bool used = false;
        for (int j = 0; j < i; ++j) used |= (top[j] == c);
        if (!used && (best < 0 || row[c] > row[best])) best = c;
      }
      top[i] = best;
    }
    // 2. token passing into nxt[] (linear-probe on (node) identity)
    int nn = 0;
    int* tc = tbuf[cb];
    int* tn = tbuf[cb ^ 1];
    auto find_or_add = [&](int node, int len, int last) -> int {
      for (int i = 0; i < nn; ++i)
        if (nxt[i].node == node) return i;
      if (nn >= PB_MAX_CAND) { overflow = true; return nn - 1; }
      nxt[nn] = Hyp{node, len, last, PB_NEG, PB_NEG, PB_NEG, PB_NEG, PB_NEG};
      return nn++;
    };
    auto child = [&](int node, int tok) -> int {
      for (int i = ntrie - 1; i > 0; --i)
        if (tpar[i] == node && ttok[i] == tok) return i;
      if (ntrie >= p.trie_cap) { overflow = true; return ntrie - 1; }
      tpar[ntrie] = node; ttok[ntrie] = tok;
      return ntrie++;
    };
    auto copy_times = [&](int* dst, const int* src, int n) { for (int i = 0; i < n && i < p.max_len; ++i) dst[i] = src[i]; };
    for (int i = 0; i < k; ++i) {
      const int id = top[i];
      const float prob = row[id];
      for (int h = 0; h < ncur; ++h) {
        const Hyp ps = cur[h];
        const int* ps_ts = tc + (size_t)h * tstride;
        const int* ps_tns = ps_ts + p.max_len;
        const int* ps_times = ps.v_s > ps.v_ns ? ps_ts : ps_tns;
        if (id == p.blank) {
          const int j = find_or_add(ps.node, ps.len, ps.last);
          nxt[j].s = log_add(nxt[j].s, hyp_score(ps) + prob);
          nxt[j].v_s = hyp_viterbi(ps) + prob;
          copy_times(tn + (size_t)j * tstride, ps_times, ps.len);
        } else if (ps.len > 0 && id == ps.last) {
          const int j1 = find_or_add(ps.node, ps.len, ps.last);
          nxt[j1].ns = log_add(nxt[j1].ns, ps.ns + prob);
          if (nxt[j1].v_ns < ps.v_ns + prob) {
            nxt[j1].v_ns = ps.v_ns + prob;
            if (nxt[j1].cur_token_prob < prob) {
              nxt[j1].cur_token_prob = prob;
              int* d = tn + (size_t)j1 * tstride + p.max_len;
              copy_times(d, ps_tns, ps.len);
              if (ps.len > 0 && ps.len <= p.max_len) d[ps.len - 1] = t;
            }
          }
          if (ps.len + 1 > p.max_len) { overflow = true; continue; }
          const int j2 = find_or_add(child(ps.node, id), ps.len + 1, id);
          nxt[j2].ns = log_add(nxt[j2].ns, ps.s + prob);
          if (nxt[j2].v_ns < ps.v_s + prob) {
            nxt[j2].v_ns = ps.v_s + prob;
            nxt[j2].cur_token_prob = prob;
            int* d = tn + (size_t)j2 * tstride + p.max_len;
            copy_times(d, ps_ts, ps.len);
            d[ps.len] = t;
          }
        } else {
          if (ps.len + 1 > p.max_len) { overflow = true; continue; }
          const int j = find_or_add(child(ps.node, id), ps.len + 1, id);
          nxt[j].ns = log_add(nxt[j].ns, hyp_score(ps) + prob);
          if (nxt[j].v_ns < hyp_viterbi(ps) + prob) {
            nxt[j].v_ns = hyp_viterbi(ps) + prob;
            nxt[j].cur_token_prob = prob;
            int* d = tn + (size_t)j * tstride + p.max_len;
            copy_times(d, ps_times, ps.len);
            d[ps.len] = t;
          }
        }
      }
    }
    // 3. second beam: keep the best second_beam by score (descending); selection sort keeps earlier candidates on ties
    const int keep = min(min(nn, p.second_beam), PB_MAX_BEAM);
    int order[PB_MAX_BEAM];
    for (int r = 0; r < keep; ++r) {
      int best = -1;
      float bs = 0.f;
      for (int i = 0; i < nn; ++i) {
        bool used = false;
        for (int j = 0; j < r; ++j) used |= (order[j] == i);
        if (used) continue;
        const float sc = hyp_score(nxt[i]);
        if (best < 0 || sc > bs) { best = i; bs = sc; }
      }
      order[r] = best;
    }
    // 4. new beam; compact the time vectors into the other buffer in beam order
    for (int r = 0; r < keep; ++r) cur[r] = nxt[order[r]];
    // tn currently holds times indexed by candidate; re-index into tc (free now) by beam position
    for (int r = 0; r < keep; ++r) {
      const int* src = tn + (size_t)order[r] * tstride;
      int* dst = tc + (size_t)r * tstride;
      for (int i = 0; i < (int)tstride; ++i) dst[i] = src[i];
    }
    ncur = keep;
    // tc now holds the current beam's times; keep cb unchanged
  }
  // outputs
  p.out_n[u] = ncur;
  const int* tc = tbuf[cb];
  for (int r = 0; r < ncur; ++r) {
    const Hyp& h = cur[r];
    const size_t o = (size_t)u * p.second_beam + r;
    p.out_len[o] = h.len;
    p.out_score[o] = hyp_score(h);
    p.out_viterbi[o] = hyp_viterbi(h);
    int node = h.node;
    for (int i = h.len - 1; i >= 0; --i) { p.out_ids[o * p.max_len + i] = ttok[node]; node = tpar[node]; }
    const int* ts = tc + (size_t)r * tstride;
    const int* src = h.v_s > h.v_ns ? ts : ts + p.max_len;
    for (int i = 0; i < h.len; ++i) p.out_times[o * p.max_len + i] = src[i];
  }
  if (overflow) p.status[u] = 1;
}

thread_local char g_perr[256] = "";

}  // namespace

extern "C" {

const char* b2t_prefix_last_error(void) { return g_perr; }

// logp: host [N][T][C] log-probabilities; lens: host [N].  Outputs (host): ids [N][second_beam][max_len], len / score /
// viterbi [N][second_beam], times [N][second_beam][max_len], n_hyp [N]; hypotheses are sorted best first.
int b2t_prefix_beam_search(const float* logp, const int* lens, int N, int T, int C, int blank, int first_beam, int second_beam, int max_len,
                           int* out_ids, int* out_len, float* out_score, float* out_viterbi, int* out_times, int* out_n) {
  if (!logp || !lens || !out_ids || !out_len || !out_score || !out_viterbi || !out_times || !out_n || N < 1 || T < 0 || C < 1 || max_len < 1) {
    snprintf(g_perr, sizeof(g_perr), "bad arguments");
    return B2T_ERR_ARG;
  }
  if (second_beam < 1 || second_beam > PB_MAX_BEAM || first_beam < 1 || first_beam > PB_MAX_TOPK) {
    snprintf(g_perr, sizeof(g_perr), "beam sizes must be in [1, %d]", PB_MAX_BEAM);
    return B2T_ERR_UNSUPPORTED;
  }
  PbParams p;
  p.N = N; p.T = T; p.C = C; p.blank = blank; p.first_beam = first_beam; p.second_beam = second_beam; p.max_len = max_len;
  p.trie_cap = 1 + T * second_beam * first_beam;
  float *d_logp, *d_score, *d_vit;
  int *d_lens, *d_tp, *d_tt, *d_times, *d_ids, *d_len, *d_otimes, *d_n, *d_status;
  Hyp* d_hyps;
  if (second_beam * (first_beam + 1) > PB_MAX_CAND) {
    snprintf(g_perr, sizeof(g_perr), "second_beam * (first_beam + 1) must be <= %d", PB_MAX_CAND);
    return B2T_ERR_UNSUPPORTED;
  }
  const size_t nb = (size_t)N * second_beam;
  const size_t times_elems = (size_t)N * 2 * PB_MAX_CAND * 2 * max_len;
  bool ok = cudaMalloc(&d_logp, (size_t)N * T * C * 4 + 4) == cudaSuccess && cudaMalloc(&d_lens, N * 4) == cudaSuccess &&
            cudaMalloc(&d_tp, (size_t)N * p.trie_cap * 4) == cudaSuccess && cudaMalloc(&d_tt, (size_t)N * p.trie_cap * 4) == cudaSuccess &&
            cudaMalloc(&d_times, times_elems * 4) == cudaSuccess && cudaMalloc(&d_ids, nb * max_len * 4) == cudaSuccess &&
            cudaMalloc(&d_len, nb * 4) == cudaSuccess && cudaMalloc(&d_score, nb * 4) == cudaSuccess && cudaMalloc(&d_vit, nb * 4) == cudaSuccess &&
            cudaMalloc(&d_otimes, nb * max_len * 4) == cudaSuccess && cudaMalloc(&d_n, N * 4) == cudaSuccess && cudaMalloc(&d_status, N * 4) == cudaSuccess &&
            cudaMalloc(&d_hyps, (size_t)N * (PB_MAX_BEAM + PB_MAX_CAND) * sizeof(Hyp)) == cudaSuccess;
  if (!ok) { snprintf(g_perr, sizeof(g_perr), "cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())); return B2T_ERR_CUDA; }
  cudaMemcpy(d_logp, logp, (size_t)N * T * C * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_lens, lens, N * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_status, 0, N * 4); cudaMemset(d_ids, 0, nb * max_len * 4); cudaMemset(d_otimes, 0, nb * max_len * 4);
  cudaMemset(d_len, 0, nb * 4); cudaMemset(d_score, 0, nb * 4); cudaMemset(d_vit, 0, nb * 4);
  p.hyps = d_hyps;
  p.logp = d_logp; p.lens = d_lens; p.trie_parent = d_tp; p.trie_token = d_tt; p.times = d_times;
  p.out_ids = d_ids; p.out_len = d_len; p.out_score = d_score; p.out_viterbi = d_vit; p.out_times = d_otimes; p.out_n = d_n; p.status = d_status;
  prefix_beam_kernel<<<(N + 31) / 32, 32>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  int rc = 0;
  if (e != cudaSuccess) { snprintf(g_perr, sizeof(g_perr), "prefix_beam_kernel: %s", cudaGetErrorString(e)); rc = B2T_ERR_CUDA; }
  std::vector<int> status(N);
  if (!rc) {
    cudaMemcpy(out_ids, d_ids, nb * max_len * 4, cudaMemcpyDeviceToHost); cudaMemcpy(out_len, d_len, nb * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out_score, d_score, nb * 4, cudaMemcpyDeviceToHost); cudaMemcpy(out_viterbi, d_vit, nb * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(out_times, d_otimes, nb * max_len * 4, cudaMemcpyDeviceToHost); cudaMemcpy(out_n, d_n, N * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(status.data(), d_status, N * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < N; ++i)
      if (status[i]) { snprintf(g_perr, sizeof(g_perr), "utterance %d: prefix longer than max_len=%d or candidate overflow", i, max_len); rc = B2T_ERR_WORKSPACE; }
  }
  void* ptrs[] = {d_logp, d_lens, d_tp, d_tt, d_times, d_ids, d_len, d_score, d_vit, d_otimes, d_n, d_status, d_hyps};
  for (void* q : ptrs) cudaFree(q);
  return rc;
}

}  // extern "C"
