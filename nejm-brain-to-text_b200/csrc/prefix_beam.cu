// LM-free CTC prefix beam search (the searcher BrainSpeechDecoder builds when no FST is given).
//
// Reference: language_model/runtime/core/decoder/ctc_prefix_beam_search.cc:44-136, PrefixScore .h:27-42,
// LogAdd utils/utils.cc:24-30.  Golden vector: ctc_prefix_beam_search_test.cc:18-59.
//
// One WARP per utterance.  The per-frame update is a chain of order-dependent merges (which candidate is created first
// decides ties in the second beam), so lane 0 walks the (token, hypothesis) pairs in the reference's order, but every step of
// that walk is O(1): prefixes are interned as (parent, token) trie nodes found through a hash table, the candidates of the
// frame are found through a second hash table keyed by trie node, and the per-hypothesis Viterbi time vectors are copied by
// all 32 lanes.  The first beam (top-k classes) and the second beam (top-n candidates, earlier candidate first among equal
// scores) are warp-wide arg-max selections.  Throughput comes from decoding many utterances at once.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <chrono>
#include <stdio.h>

#include <vector>

#include <cuda_runtime.h>

#include "../../include/b2t_b200.h"

namespace {

constexpr int PB_MAX_BEAM = 512;      // second_beam_size limit (BASELINE.json configs[4] sweeps the width to 500)
constexpr int PB_MAX_TOPK = 64;       // first_beam_size limit
constexpr int PB_MAX_BUCKETS = 1109;  // largest bucket count the emulated unordered_map can reach with <= 512 elements (13 -> ... -> 541), one step spare
constexpr int PB_SMEM_LIMIT = 227 * 1024 - 1024;   // dynamic part; the kernel has a few hundred bytes of static shared memory as well
// The per-utterance working set lives in dynamic shared memory sized from the beams (second_beam * (first_beam + 1) candidates of
// 32 bytes dominate: 3.5 KB at 10 x 10, 180 KB at 512 x 10); the host entry point rejects combinations beyond 227 KB.
#define PB_NEG (-3.402823466e+38f)    // -kFloatMax

struct Hyp;
struct PbParams {
  const float* logp;   // [N][T][C]
  const int* lens;     // [N]
  int N, T, C, blank, first_beam, second_beam, max_len;
  int cand_cap, chash_cap;                             // candidates per frame; candidate hash size (power of two)
  struct Hyp* nxt_g; unsigned short* chash_g;          // [N][cand_cap] / [N][chash_cap] in global memory when the candidates do not fit shared memory (else null)
  // per-utterance scratch
  int* trie_parent; int* trie_token; int trie_cap;     // [N][trie_cap]
  int* trie_hash; int trie_hash_cap;                   // [N][trie_hash_cap] (power of two) -> trie node or -1
  unsigned long long* trie_h64;                        // [N][trie_cap] PrefixHash of the node's prefix (ctc_prefix_beam_search.h:44-53)
  int* times;          // [N][2 buffers][cand_cap][2 (s, ns)][max_len]
  // outputs
  int* out_ids; int* out_len; float* out_score; float* out_viterbi; int* out_times; int* out_n;
  int* status;
};

struct Hyp {
  int node, len, last;              // trie node of the prefix, its length, its last token (-1 when empty)
  float s, ns, v_s, v_ns, cur_token_prob;
};

constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned short CH_EMPTY = 0xffffu;

__device__ inline float log_add(float x, float y) {
  if (x <= PB_NEG) return y;
  if (y <= PB_NEG) return x;
  const float m = fmaxf(x, y);
  return logf(expf(x - m) + expf(y - m)) + m;
}
__device__ inline float hyp_score(const Hyp& h) { return log_add(h.s, h.ns); }
__device__ inline float hyp_viterbi(const Hyp& h) { return h.v_s > h.v_ns ? h.v_s : h.v_ns; }
__device__ inline unsigned mix(unsigned a, unsigned b) {
  unsigned h = a * 0x9E3779B1u ^ (b + 0x85EBCA6Bu + (a << 6) + (a >> 2));
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
  return h;
}

// warp-wide arg-max of (value, index): larger value first, smaller index among equal values; idx < 0 = no candidate
__device__ inline void warp_argmax(float& v, int& idx) {
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(FULL, v, o);
    const int oi = __shfl_xor_sync(FULL, idx, o);
    if (oi >= 0 && (idx < 0 || ov > v || (ov == v && oi < idx))) { v = ov; idx = oi; }
  }
}

// Iteration order of the reference's cur_hyps_ (std::unordered_map<vector<int>, PrefixScore, PrefixHash>, libstdc++).
// The Viterbi time vectors depend on it: within a frame the first writer of a candidate keeps its times when a later writer
// brings a better Viterbi score at the same token probability (cur_token_prob guard, ctc_prefix_beam_search.cc:79-86), and the
// writers are visited in the map's order, not in score order.  libstdc++ keeps all nodes in one forward list; a node whose
// bucket is empty goes to the head of the list, otherwise to the beginning of its bucket's run; growing the table re-inserts
// the nodes in list order by the same rule; clear() keeps the bucket count (13 after the first insertion, then the next
// prime >= twice the old count in libstdc++'s table: 29, 59, 127, 257, 541).  Emulated serially for the hypotheses of a frame.
struct UMapOrder {
  int B, next_resize, head, n;
  int* nxt;                  // [second_beam] successor in the list (-1 = end)
  int* before;               // [PB_MAX_BUCKETS] per bucket: -2 = empty, -1 = the list head sentinel, >= 0 = node preceding the bucket's run
  unsigned long long* h;     // [second_beam] hash of the node's prefix
  __device__ void clear() {
    head = -1; n = 0;
    for (int b = 0; b < B; ++b) before[b] = -2;
  }
  __device__ void link(int node, int nb) {
    const int b = (int)(h[node] % (unsigned long long)nb);
    if (before[b] != -2) {
      const int pr = before[b];
      if (pr == -1) { nxt[node] = head; head = node; }
      else { nxt[node] = nxt[pr]; nxt[pr] = node; }
    } else {
      nxt[node] = head; head = node;
      if (nxt[node] >= 0) before[(int)(h[nxt[node]] % (unsigned long long)nb)] = node;
      before[b] = -1;
    }
  }
  __device__ void insert(int node, unsigned long long hv, int* scratch) {
    h[node] = hv;
    if (n + 1 > next_resize) {
      const int minb = max(n + 1, next_resize ? 0 : 11);
      if (minb >= B) {
        const int want = max(minb + 1, 2 * B);
        const int nb = want <= 13 ? 13 : want <= 29 ? 29 : want <= 59 ? 59 : want <= 127 ? 127 : want <= 257 ? 257 : want <= 541 ? 541 : 1109;
        int cnt = 0;                                       // re-insert in the current list order
        for (int q = head; q >= 0; q = nxt[q]) scratch[cnt++] = q;
        head = -1;
        for (int b = 0; b < nb; ++b) before[b] = -2;
        for (int i = 0; i < cnt; ++i) link(scratch[i], nb);
        B = nb; next_resize = nb;
      } else next_resize = B;
    }
    link(node, B);
    ++n;
  }
};

__global__ void __launch_bounds__(32) prefix_beam_kernel(const PbParams p) {
  // the state lane 0 walks with dependent accesses lives in (dynamic) shared memory; the trie, its hash and the time vectors stay
  // in global memory (one or two accesses per pair / copied by all lanes)
  extern __shared__ __align__(16) unsigned char pb_smem[];
  __shared__ int s_top[PB_MAX_TOPK];
  const int u = blockIdx.x, lane = threadIdx.x;
  // (wide first beams: second_beam x (first_beam + 1) candidates of 32 B exceed shared memory; they then live in global memory and the
  //  serial walk pays L2 latency per access -- a functional path for the sweep's corner, not a fast one)
  Hyp* s_nxt = p.nxt_g ? p.nxt_g + (size_t)u * p.cand_cap : reinterpret_cast<Hyp*>(pb_smem);
  Hyp* s_cur = p.nxt_g ? reinterpret_cast<Hyp*>(pb_smem) : s_nxt + p.cand_cap;
  unsigned long long* s_uh = reinterpret_cast<unsigned long long*>(s_cur + p.second_beam);
  int* s_order = reinterpret_cast<int*>(s_uh + p.second_beam);
  int* s_iter = s_order + p.second_beam;
  int* s_unext = s_iter + p.second_beam;
  int* s_scratch = s_unext + p.second_beam;
  int* s_ubefore = s_scratch + p.second_beam;
  unsigned short* s_chash = p.chash_g ? p.chash_g + (size_t)u * p.chash_cap : reinterpret_cast<unsigned short*>(s_ubefore + PB_MAX_BUCKETS);
  if (u >= p.N) return;
  const float* logp = p.logp + (size_t)u * p.T * p.C;
  const int Tn = min(max(p.lens[u], 0), p.T);
  int* tpar = p.trie_parent + (size_t)u * p.trie_cap;
  int* ttok = p.trie_token + (size_t)u * p.trie_cap;
  int* thash = p.trie_hash + (size_t)u * p.trie_hash_cap;
  unsigned short* chash = s_chash;
  const unsigned chmask = (unsigned)p.chash_cap - 1u;
  unsigned long long* th64 = p.trie_h64 + (size_t)u * p.trie_cap;
  int ntrie = 1;                                        // node 0 = empty prefix (lane 0's copy is the authoritative one)
  UMapOrder um;                                         // lane 0 only
  um.B = 1; um.next_resize = 0; um.head = -1; um.n = 0; um.nxt = s_unext; um.before = s_ubefore; um.h = s_uh;
  if (lane == 0) {
    tpar[0] = -1; ttok[0] = -1; th64[0] = 0ull;
    s_ubefore[0] = -2;
    um.insert(0, 0ull, s_scratch);                     // Reset(): cur_hyps_[empty prefix]
    s_iter[0] = 0;
  }
  for (int i = lane; i < p.trie_hash_cap; i += 32) thash[i] = -1;
  for (int i = lane; i < p.chash_cap; i += 32) chash[i] = CH_EMPTY;
  const size_t tstride = (size_t)2 * p.max_len;          // per hypothesis: times_s | times_ns
  int* tbuf[2] = {p.times + (size_t)u * 2 * p.cand_cap * tstride, p.times + ((size_t)u * 2 + 1) * p.cand_cap * tstride};

  Hyp* cur = s_cur;
  Hyp* nxt = s_nxt;
  int ncur = 1;
  if (lane == 0) cur[0] = Hyp{0, 0, -1, 0.0f, PB_NEG, 0.0f, 0.0f, PB_NEG};
  const int k = min(min(p.first_beam, p.C), PB_MAX_TOPK);
  bool overflow = false;
  __syncwarp();

  for (int t = 0; t < Tn; ++t) {
    const float* row = logp + (size_t)t * p.C;
    // 1. first beam: top-k classes, descending value, lower index first among ties (lane l owns classes l and l + 32)
    {
      float v0 = lane < p.C ? row[lane] : 0.f, v1 = lane + 32 < p.C ? row[lane + 32] : 0.f;
      bool u0 = lane < p.C, u1 = lane + 32 < p.C;       // still available
      for (int i = 0; i < k; ++i) {
        float v; int idx;
        if (u0 && (!u1 || v0 >= v1)) { v = v0; idx = lane; } else if (u1) { v = v1; idx = lane + 32; } else { v = 0.f; idx = -1; }
        warp_argmax(v, idx);
        if (idx == lane) u0 = false;
        if (idx == lane + 32) u1 = false;
        if (lane == 0) s_top[i] = idx;
      }
    }
    __syncwarp();
    // 2. token passing into nxt[]: lane 0 walks the pairs in the reference's order, all lanes copy the time vectors
    int nn = 0;
    int* tc = tbuf[0];
    int* tn = tbuf[1];
    for (int i = 0; i < k; ++i) {
      const int id = s_top[i];
      const float prob = row[id];
      for (int hh = 0; hh < ncur; ++hh) {
        const int h = s_iter[hh];                        // the reference's unordered_map iteration order (see UMapOrder)
        // up to two time-vector copies per pair: (dst, src, n, position that receives t or -1)
        long long c_dst[2] = {-1, -1}, c_src[2] = {0, 0};
        int c_n[2] = {0, 0}, c_set[2] = {-1, -1}, ncp = 0;
        if (lane == 0) {
          auto find_or_add = [&](int node, int len, int last) -> int {
            unsigned slot = mix((unsigned)node, 0x51u) & chmask;
            for (;;) {
              const unsigned short j = chash[slot];
              if (j == CH_EMPTY) break;
              if (nxt[j].node == node) return (int)j;
              slot = (slot + 1) & chmask;
            }
            if (nn >= p.cand_cap) { overflow = true; return nn - 1; }
            nxt[nn] = Hyp{node, len, last, PB_NEG, PB_NEG, PB_NEG, PB_NEG, PB_NEG};
            chash[slot] = (unsigned short)nn;
            return nn++;
          };
          auto child = [&](int node, int tok) -> int {
            unsigned slot = mix((unsigned)node, (unsigned)tok) & (unsigned)(p.trie_hash_cap - 1);
            for (;;) {
              const int j = thash[slot];
              if (j < 0) break;
              if (tpar[j] == node && ttok[j] == tok) return j;
              slot = (slot + 1) & (unsigned)(p.trie_hash_cap - 1);
            }
            if (ntrie >= p.trie_cap) { overflow = true; return ntrie - 1; }
            tpar[ntrie] = node; ttok[ntrie] = tok;
            th64[ntrie] = (unsigned long long)(long long)tok + 31ull * th64[node];
            thash[slot] = ntrie;
            return ntrie++;
          };
          auto want_copy = [&](int* dst, const int* src, int n, int set_pos) {
            c_dst[ncp] = dst - p.times; c_src[ncp] = src - p.times; c_n[ncp] = n; c_set[ncp] = set_pos; ++ncp;
          };
          const Hyp ps = cur[h];
          const int* ps_ts = tc + (size_t)h * tstride;
          const int* ps_tns = ps_ts + p.max_len;
          const int* ps_times = ps.v_s > ps.v_ns ? ps_ts : ps_tns;
          if (id == p.blank) {
            const int j = find_or_add(ps.node, ps.len, ps.last);
            nxt[j].s = log_add(nxt[j].s, hyp_score(ps) + prob);
            nxt[j].v_s = hyp_viterbi(ps) + prob;
            want_copy(tn + (size_t)j * tstride, ps_times, ps.len, -1);
          } else if (ps.len > 0 && id == ps.last) {
            const int j1 = find_or_add(ps.node, ps.len, ps.last);
            nxt[j1].ns = log_add(nxt[j1].ns, ps.ns + prob);
            if (nxt[j1].v_ns < ps.v_ns + prob) {
              nxt[j1].v_ns = ps.v_ns + prob;
              if (nxt[j1].cur_token_prob < prob) {
                nxt[j1].cur_token_prob = prob;
                want_copy(tn + (size_t)j1 * tstride + p.max_len, ps_tns, ps.len, (ps.len > 0 && ps.len <= p.max_len) ? ps.len - 1 : -1);
              }
            }
            if (ps.len + 1 > p.max_len) overflow = true;
            else {
              const int j2 = find_or_add(child(ps.node, id), ps.len + 1, id);
              nxt[j2].ns = log_add(nxt[j2].ns, ps.s + prob);
              if (nxt[j2].v_ns < ps.v_s + prob) {
                nxt[j2].v_ns = ps.v_s + prob;
                nxt[j2].cur_token_prob = prob;
                want_copy(tn + (size_t)j2 * tstride + p.max_len, ps_ts, ps.len, ps.len);
              }
            }
          } else {
            if (ps.len + 1 > p.max_len) overflow = true;
            else {
              const int j = find_or_add(child(ps.node, id), ps.len + 1, id);
              nxt[j].ns = log_add(nxt[j].ns, hyp_score(ps) + prob);
              if (nxt[j].v_ns < hyp_viterbi(ps) + prob) {
                nxt[j].v_ns = hyp_viterbi(ps) + prob;
                nxt[j].cur_token_prob = prob;
                want_copy(tn + (size_t)j * tstride + p.max_len, ps_times, ps.len, ps.len);
              }
            }
          }
        }
        ncp = __shfl_sync(FULL, ncp, 0);
        for (int c = 0; c < ncp; ++c) {                  // in order: a later copy may overwrite an earlier one's destination
          const long long dsto = __shfl_sync(FULL, c_dst[c], 0), srco = __shfl_sync(FULL, c_src[c], 0);
          const int n = __shfl_sync(FULL, c_n[c], 0), sp = __shfl_sync(FULL, c_set[c], 0);
          int* dst = p.times + dsto;
          const int* src = p.times + srco;
          const int lim = min(max(n, sp + 1), p.max_len);
          for (int q = lane; q < lim; q += 32) dst[q] = (q == sp) ? t : src[q];
          __syncwarp();
        }
      }
    }
    nn = __shfl_sync(FULL, nn, 0);
    __syncwarp();
    // 3. second beam: the best second_beam candidates by score, descending; the earlier candidate wins among equal scores
    const int keep = min(min(nn, p.second_beam), PB_MAX_BEAM);
    {
      // the candidates' scores are computed once and parked in cur_token_prob (dead after the token passing of this frame and
      // never read from a beam entry); a taken candidate is marked with -inf (scores are > -FLT_MAX).  Lane l owns l, l + 32, ...
      for (int i = lane; i < nn; i += 32) nxt[i].cur_token_prob = hyp_score(nxt[i]);
      __syncwarp();
      for (int r = 0; r < keep; ++r) {
        float bv = 0.f; int bi = -1;
        for (int i = lane; i < nn; i += 32) {
          const float sc = nxt[i].cur_token_prob;
          if (sc == -INFINITY) continue;
          if (bi < 0 || sc > bv) { bv = sc; bi = i; }
        }
        warp_argmax(bv, bi);
        if (bi >= 0 && (bi & 31) == lane) nxt[bi].cur_token_prob = -INFINITY;
        if (lane == 0) s_order[r] = bi;
      }
    }
    __syncwarp();
    // 4. new beam; compact the time vectors into the other buffer in beam order; clear the candidate hash
    for (int r = lane; r < keep; r += 32) cur[r] = nxt[s_order[r]];
    for (int r = 0; r < keep; ++r) {
      const int* src = tn + (size_t)s_order[r] * tstride;
      int* dst = tc + (size_t)r * tstride;
      for (int i = lane; i < (int)tstride; i += 32) dst[i] = src[i];
    }
    for (int i = lane; i < p.chash_cap; i += 32) chash[i] = CH_EMPTY;
    ncur = keep;
    __syncwarp();
    if (lane == 0) {                                     // cur_hyps_.clear(); cur_hyps_[prefix] = score in sorted order
      um.clear();
      for (int r = 0; r < keep; ++r) um.insert(r, th64[cur[r].node], s_scratch);
      int c = 0;
      for (int q = um.head; q >= 0; q = um.nxt[q]) s_iter[c++] = q;
    }
    __syncwarp();
  }
  // outputs
  if (lane == 0) p.out_n[u] = ncur;
  const int* tc = tbuf[0];
  for (int r = lane; r < ncur; r += 32) {
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
  overflow = __any_sync(FULL, overflow);
  if (overflow && lane == 0) p.status[u] = 1;
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
  const double t_begin = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  PbParams p;
  p.N = N; p.T = T; p.C = C; p.blank = blank; p.first_beam = first_beam; p.second_beam = second_beam; p.max_len = max_len;
  p.trie_cap = 1 + T * second_beam * first_beam;
  float *d_logp, *d_score, *d_vit;
  int *d_lens, *d_tp, *d_tt, *d_th, *d_times, *d_ids, *d_len, *d_otimes, *d_n, *d_status;
  p.trie_hash_cap = 1024;
  while (p.trie_hash_cap < 2 * p.trie_cap) p.trie_hash_cap *= 2;
  p.cand_cap = second_beam * (std::min(first_beam, C) + 1);
  p.chash_cap = 256;
  while (p.chash_cap < 2 * p.cand_cap && p.chash_cap < 8192) p.chash_cap *= 2;
  while (p.chash_cap * 3 < p.cand_cap * 4) p.chash_cap *= 2;          // load factor <= 0.75 when the 2x table would not fit
  const size_t smem_small = (size_t)second_beam * (sizeof(Hyp) + 8 + 4 * 4) + PB_MAX_BUCKETS * 4;
  size_t smem = (size_t)p.cand_cap * sizeof(Hyp) + smem_small + (size_t)p.chash_cap * 2;
  const bool cand_in_global = smem > (size_t)PB_SMEM_LIMIT;            // wide first beam x wide second beam
  if (cand_in_global) {
    smem = smem_small;
    while (p.chash_cap < 2 * p.cand_cap) p.chash_cap *= 2;
  }
  if (p.cand_cap >= 0xffff) {
    snprintf(g_perr, sizeof(g_perr), "second_beam %d x (first_beam %d + 1) candidates per frame exceed the 65534 the candidate hash can index", second_beam, first_beam);
    return B2T_ERR_UNSUPPORTED;
  }
  const size_t nb = (size_t)N * second_beam;
  const size_t times_elems = (size_t)N * 2 * p.cand_cap * 2 * max_len;
  // one cached device workspace per host thread, grown on demand (cudaMalloc / cudaFree per call cost more than the search)
  struct Ws { void* base = nullptr; size_t cap = 0; ~Ws() { /* released with the context */ } };
  static thread_local Ws ws;
  size_t off = 0;
  auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  const size_t o_logp = carve((size_t)N * T * C * 4 + 4), o_lens = carve((size_t)N * 4), o_tp = carve((size_t)N * p.trie_cap * 4),
               o_tt = carve((size_t)N * p.trie_cap * 4), o_th = carve((size_t)N * p.trie_hash_cap * 4), o_h64 = carve((size_t)N * p.trie_cap * 8), o_times = carve(times_elems * 4),
               o_ids = carve(nb * max_len * 4), o_len = carve(nb * 4), o_score = carve(nb * 4), o_vit = carve(nb * 4),
               o_otimes = carve(nb * max_len * 4), o_n = carve((size_t)N * 4), o_status = carve((size_t)N * 4),
               o_nxt = carve(cand_in_global ? (size_t)N * p.cand_cap * sizeof(Hyp) : 0), o_ch = carve(cand_in_global ? (size_t)N * p.chash_cap * 2 : 0);
  bool ok = true;
  if (off > ws.cap) {
    if (ws.base) cudaFree(ws.base);
    ws.base = nullptr; ws.cap = 0;
    ok = cudaMalloc(&ws.base, off) == cudaSuccess;
    if (ok) ws.cap = off;
  }
  if (ok) {
    uint8_t* b8 = reinterpret_cast<uint8_t*>(ws.base);
    d_logp = reinterpret_cast<float*>(b8 + o_logp); d_lens = reinterpret_cast<int*>(b8 + o_lens);
    d_tp = reinterpret_cast<int*>(b8 + o_tp); d_tt = reinterpret_cast<int*>(b8 + o_tt); d_th = reinterpret_cast<int*>(b8 + o_th);
    d_times = reinterpret_cast<int*>(b8 + o_times); d_ids = reinterpret_cast<int*>(b8 + o_ids); d_len = reinterpret_cast<int*>(b8 + o_len);
    d_score = reinterpret_cast<float*>(b8 + o_score); d_vit = reinterpret_cast<float*>(b8 + o_vit);
    d_otimes = reinterpret_cast<int*>(b8 + o_otimes); d_n = reinterpret_cast<int*>(b8 + o_n); d_status = reinterpret_cast<int*>(b8 + o_status);
  }
  if (!ok) { snprintf(g_perr, sizeof(g_perr), "cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError())); return B2T_ERR_CUDA; }
  cudaMemcpy(d_logp, logp, (size_t)N * T * C * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_lens, lens, N * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_status, 0, N * 4); cudaMemset(d_ids, 0, nb * max_len * 4); cudaMemset(d_otimes, 0, nb * max_len * 4);
  cudaMemset(d_len, 0, nb * 4); cudaMemset(d_score, 0, nb * 4); cudaMemset(d_vit, 0, nb * 4);
  p.logp = d_logp; p.lens = d_lens; p.trie_parent = d_tp; p.trie_token = d_tt; p.trie_hash = d_th; p.trie_h64 = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws.base) + o_h64); p.times = d_times;
  p.nxt_g = cand_in_global ? reinterpret_cast<Hyp*>(reinterpret_cast<uint8_t*>(ws.base) + o_nxt) : nullptr;
  p.chash_g = cand_in_global ? reinterpret_cast<unsigned short*>(reinterpret_cast<uint8_t*>(ws.base) + o_ch) : nullptr;
  p.out_ids = d_ids; p.out_len = d_len; p.out_score = d_score; p.out_viterbi = d_vit; p.out_times = d_otimes; p.out_n = d_n; p.status = d_status;
  static const bool timing = getenv("B2T_DECODER_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  cudaDeviceSynchronize();
  const double t_k0 = now();
  cudaError_t e = smem > 48 * 1024 ? cudaFuncSetAttribute(prefix_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) : cudaSuccess;
  if (e == cudaSuccess) {
    prefix_beam_kernel<<<N, 32, smem>>>(p);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (timing) fprintf(stderr, "b2t prefix beam: N=%d T=%d beams %dx%d trie_cap %d: setup %.2f ms, kernel %.2f ms\n", N, T, first_beam, second_beam, p.trie_cap, t_k0 - t_begin, now() - t_k0);
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
  return rc;
}

}  // extern "C"
