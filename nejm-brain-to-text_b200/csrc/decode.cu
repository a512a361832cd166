// GPU WFST token-passing decoder (the production n-gram CTC search) + host lattice post-processing, exported as
// b2t_decoder_* in include/b2t_b200.h.
//
// Reference path replaced (language_model/runtime/core):
//   decoder/ctc_wfst_beam_search.cc:70-160      frame feeder, blank-frame skipping, final n-best
//   kaldi/decoder/lattice-faster-decoder.cc     ProcessEmitting :723-824, ProcessNonemitting :839-909,
//                                               GetCutoff :651-720, FinalizeDecoding :632-647 (pruning :298-506)
//   kaldi/decoder/lattice-faster-online-decoder.cc:59-177   back-pointer best path
//   decoder/brain_speech_decoder.cc:103-137     Decode / UpdateResult;  utils/string.cc:121-146 ProcessBlank
//   server/x86/python/lm_decoder.cc:14-49       DecodeNumpy / DecodeNumpyLogProbs
//
// Design.  One CTA per utterance ("slot"); the whole frame loop runs inside one kernel launch, so a batch of
// utterances decodes concurrently on as many SMs with no host round trip per frame.  The graph (CSR, 16-byte arcs)
// is read-only and shared; per slot there is a dense state -> (best cost, winning arc) table updated with 64-bit
// atomicMin (recombination), a token pool and a link pool.  Per frame:
//   1. best cost (block min) and the Kaldi cutoff (beam / max_active / min_active; exact k-th smallest by radix select)
//   2. pass A over emitting arcs of the surviving tokens: exact next-frame best -> next_cutoff
//   3. pass B: arcs below next_cutoff become forward links; atomicMin recombines tokens per destination state
//   4. epsilon closure: relax input-epsilon arcs until no token improves, then emit the epsilon links
//   5. back-pointers = source of the winning link; table entries of the frame are cleared
// Kaldi tightens next_cutoff while it iterates (order dependent); we use its final value from the start, which
// yields exactly the tokens that can survive the following frame's beam ("decode parity").
// The irregular once-per-sentence work (final-cost pruning, best path, n-best) runs on the host from the
// token/link pools.
#include <math.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b2t_b200.h"

namespace {

thread_local char g_derr[512] = "";
int dfail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_derr, sizeof(g_derr), fmt, ap);
  va_end(ap);
  return code;
}
#define DCK(call)                                                                                                    \
  do {                                                                                                               \
    cudaError_t _e = (call);                                                                                         \
    if (_e != cudaSuccess) return dfail(B2T_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)

struct DArc { int il, ol; float w; int next; };
struct DLink { int src, dst, il, ol; float graph, ac; };   // src/dst are token-pool indices

constexpr unsigned long long EMPTY64 = ~0ull;
constexpr int DEC_THREADS = 512;

__host__ __device__ inline unsigned int fkey(float x) {
  unsigned int k;
#ifdef __CUDA_ARCH__
  k = __float_as_uint(x);
#else
  memcpy(&k, &x, 4);
#endif
  return (k & 0x80000000u) ? ~k : (k | 0x80000000u);
}
__device__ inline float funkey(unsigned int k) {
  k = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(k);
}

struct DecParams {
  const DArc* arcs;
  const long long* off;
  const unsigned char* has_eps;
  int nstates, start;
  float beam, beam_delta, acoustic_scale, length_penalty;
  int max_active, min_active;
  unsigned long long* best64;   // [slots][nstates]
  int* tokidx;                  // [slots][nstates]
  int* tok_state; float* tok_cost; int* tok_bp; unsigned char* tok_dirty;   // [slots][tok_cap]
  DLink* links;                 // [slots][link_cap]
  int* frame_tok_off;           // [slots][max_frames + 3]  token-pool offset of frame_plus_one f; [.. + 1] = end
  int* frame_link_off;          // [slots][max_frames + 3]
  float* cost_offsets;          // [slots][max_frames]
  int* counters;                // [slots][4]: 0 ntok, 1 nlink, 2 frames decoded, 3 status
  const float* logp;            // [slots][max_frames][C]
  const int* n_fed;             // [slots]
  const int* slot_ids;          // [gridDim.x] slot handled by each CTA
  int tok_cap, link_cap, max_frames, C;
  // strict serial-order mode (decode_strict.cuh); all per slot
  const long long* eps_off;     // compact CSR of the input-epsilon arcs, in arc order
  const struct EArc* eps_arcs;
  int* first_rank;              // [slots][nstates] insertion rank of the state's token in the frame being built
  int* eps_bp;                  // [slots][nstates] source state of the epsilon arc that last lowered the token's cost
  int* bucket_min;              // [slots][hcap]    rank of the first token of each hash bucket
  int* s_cum;                   // [slots][fc + 1]  exclusive prefix sum of out-degrees in list order
  float* s_premin;              // [slots][fc]      exclusive prefix min of the tokens' best emitting arc
  unsigned int *s_k0, *s_k1;    // [slots][fc]      radix-sort keys (ping-pong)
  int *s_v0, *s_v1;             // [slots][fc]      radix-sort values
  int* s_byrank;                // [slots][fc]      states of the new frame in insertion order
  int* s_queue;                 // [slots][qcap]    ProcessNonemitting's LIFO
  int* s_hash;                  // [slots]          HashList size
  int fc, qcap, hcap;
};

struct EArc { int next; float w; int ol; int orig; };

__device__ float block_min(float v, float* sred) {
  for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sred[0];
  for (int i = 1; i < DEC_THREADS / 32; ++i) r = fminf(r, sred[i]);
  __syncthreads();
  return r;
}

// exact k-th smallest (0-based) of cost[0..n) by 4 radix passes; every thread returns the value
__device__ float select_kth(const float* cost, int n, int k, unsigned int* hist, unsigned int* sh) {
  unsigned int prefix = 0, mask = 0;
  for (int pass = 3; pass >= 0; --pass) {
    for (int i = threadIdx.x; i < 256; i += DEC_THREADS) hist[i] = 0;
    __syncthreads();
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < n; i += DEC_THREADS) {
      const unsigned int key = fkey(cost[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int acc = 0, d = 0;
      for (; d < 256; ++d) {
        if (acc + hist[d] > (unsigned int)k) break;
        acc += hist[d];
      }
      sh[0] = d; sh[1] = acc;
    }
    __syncthreads();
    prefix |= sh[0] << shift;
    mask |= 255u << shift;
    k -= (int)sh[1];
    __syncthreads();
  }
  return funkey(prefix);
}

// Relax the input-epsilon arcs of the tokens [tb, *ntok) of one frame until no token improves (ProcessNonemitting),
// then emit the epsilon forward links.  New tokens are appended to the pool.
__device__ void eps_closure(const DecParams& p, unsigned long long* best64, int* tokidx, int* tok_state, unsigned char* dirty, DLink* links,
                            int* counters, int tb, float cutoff, int* s_flag) {
  // mark every token whose state has epsilon arcs
  for (int i = tb + threadIdx.x; i < counters[0]; i += DEC_THREADS) dirty[i] = p.has_eps[tok_state[i]];
  __syncthreads();
  for (int round = 0; round < 64; ++round) {
    if (threadIdx.x == 0) *s_flag = 0;
    __syncthreads();
    const int te = counters[0];
    __syncthreads();
    for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
      if (!dirty[i]) continue;
      dirty[i] = 0;
      const int st = tok_state[i];
      const float cur = funkey((unsigned int)(best64[st] >> 32));
      if (cur >= cutoff) continue;
      for (long long a = p.off[st]; a < p.off[st + 1]; ++a) {
        const DArc arc = p.arcs[a];
        if (arc.il != 0) continue;
        const float tot = __fadd_rn(cur, arc.w);
        if (tot < cutoff) {
          const unsigned long long key = ((unsigned long long)fkey(tot) << 32) | (unsigned int)a;
          const unsigned long long old = atomicMin(&best64[arc.next], key);
          if (old == EMPTY64) {
            const int slot = atomicAdd(&counters[0], 1);
            if (slot < p.tok_cap) {
              tok_state[slot] = arc.next;
              tokidx[arc.next] = slot;
              dirty[slot] = p.has_eps[arc.next];
              if (dirty[slot]) *s_flag = 1;
            } else {
              counters[3] = 1;
            }
          } else if (key < old && (unsigned int)(old >> 32) != (unsigned int)(key >> 32)) {
            // improved an existing token: its own epsilon arcs must be relaxed again
            if (p.has_eps[arc.next]) {
              // tokidx of a token created in this very round may not be visible yet; the creator marks it dirty itself
              const int j = tokidx[arc.next];
              if (j >= tb && j < p.tok_cap) dirty[j] = 1;
              *s_flag = 1;
            }
          }
        }
      }
    }
    __syncthreads();
    if (counters[0] > p.tok_cap) { if (threadIdx.x == 0) { counters[0] = p.tok_cap; counters[3] = 1; } }
    __syncthreads();
    if (*s_flag == 0) break;
    // tokens improved while tokidx was not yet visible: re-mark conservatively (cheap, frames are small)
    for (int i = tb + threadIdx.x; i < counters[0]; i += DEC_THREADS)
      if (p.has_eps[tok_state[i]]) dirty[i] = 1;
    __syncthreads();
  }
  // epsilon links from the converged costs
  const int te = counters[0];
  for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
    const int st = tok_state[i];
    if (!p.has_eps[st]) continue;
    const float cur = funkey((unsigned int)(best64[st] >> 32));
    if (cur >= cutoff) continue;
    for (long long a = p.off[st]; a < p.off[st + 1]; ++a) {
      const DArc arc = p.arcs[a];
      if (arc.il != 0) continue;
      const float tot = __fadd_rn(cur, arc.w);
      if (tot < cutoff) {
        const int li = atomicAdd(&counters[1], 1);
        if (li < p.link_cap) links[li] = DLink{i, tokidx[arc.next], 0, arc.ol, arc.w, 0.0f};
        else counters[3] = 2;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && counters[1] > p.link_cap) counters[1] = p.link_cap;
  __syncthreads();
}

// final costs of the frame's tokens, back-pointers (source of the winning link), and table clean-up
__device__ void finish_frame(const DecParams& p, unsigned long long* best64, int* tokidx, int* tok_state, float* tok_cost, int* tok_bp,
                             const DLink* links, const int* counters, int tb, int lb) {
  const int te = counters[0], le = counters[1];
  for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) tok_cost[i] = funkey((unsigned int)(best64[tok_state[i]] >> 32));
  __syncthreads();
  for (int l = lb + threadIdx.x; l < le; l += DEC_THREADS) {
    const DLink k = links[l];
    const float tot = k.il != 0 ? __fadd_rn(__fadd_rn(tok_cost[k.src], k.ac), k.graph) : __fadd_rn(tok_cost[k.src], k.graph);
    const unsigned long long b = best64[tok_state[k.dst]];
    if ((unsigned int)(b >> 32) == fkey(tot)) {
      // the winning arc id disambiguates equal-cost links deterministically
      const int st = tok_state[k.src];
      const long long a = (long long)(unsigned int)b;
      if (a >= p.off[st] && a < p.off[st + 1]) tok_bp[k.dst] = k.src;
    }
  }
  __syncthreads();
  for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
    const int st = tok_state[i];
    best64[st] = EMPTY64;
    tokidx[st] = -1;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(DEC_THREADS, 1)
wfst_decode_kernel(const DecParams p) {
  __shared__ float sred[DEC_THREADS / 32];
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sh[2];
  __shared__ int s_flag;
  __shared__ float s_prep;
  const int slot = p.slot_ids[blockIdx.x];
  unsigned long long* best64 = p.best64 + (size_t)slot * p.nstates;
  int* tokidx = p.tokidx + (size_t)slot * p.nstates;
  int* tok_state = p.tok_state + (size_t)slot * p.tok_cap;
  float* tok_cost = p.tok_cost + (size_t)slot * p.tok_cap;
  int* tok_bp = p.tok_bp + (size_t)slot * p.tok_cap;
  unsigned char* dirty = p.tok_dirty + (size_t)slot * p.tok_cap;
  DLink* links = p.links + (size_t)slot * p.link_cap;
  int* ftok = p.frame_tok_off + (size_t)slot * (p.max_frames + 3);
  int* flink = p.frame_link_off + (size_t)slot * (p.max_frames + 3);
  float* coff = p.cost_offsets + (size_t)slot * p.max_frames;
  int* counters = p.counters + slot * 4;
  const float* logp = p.logp + (size_t)slot * p.max_frames * p.C;
  const int n_fed = p.n_fed[slot];

  if (counters[2] < 0) {
    // ---- InitDecoding (lattice-faster-decoder.cc:58-75): start token + epsilon closure with cutoff = beam
    if (threadIdx.x == 0) {
      counters[0] = 1; counters[1] = 0;
      tok_state[0] = p.start; tok_bp[0] = -1;
      best64[p.start] = ((unsigned long long)fkey(0.0f) << 32);
      tokidx[p.start] = 0;
      ftok[0] = 0; flink[0] = 0;
    }
    __syncthreads();
    eps_closure(p, best64, tokidx, tok_state, dirty, links, counters, 0, p.beam, &s_flag);
    finish_frame(p, best64, tokidx, tok_state, tok_cost, tok_bp, links, counters, 0, 0);
    if (threadIdx.x == 0) { tok_bp[0] = -1; ftok[1] = counters[0]; flink[1] = counters[1]; counters[2] = 0; }
    __syncthreads();
  }

  for (int frame = counters[2]; frame < n_fed; ++frame) {
    if (counters[3] != 0) break;
    const int tb = ftok[frame], te = ftok[frame + 1], n = te - tb;
    const float* ll = logp + (size_t)frame * p.C;
    // ---- 1. best cost + cutoff (GetCutoff, :651-720)
    float v = INFINITY;
    for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) v = fminf(v, tok_cost[i]);
    const float best = block_min(v, sred);
    float cur_cutoff, adaptive_beam;
    {
      const float beam_cutoff = __fadd_rn(best, p.beam);
      float max_cut = INFINITY, min_cut = INFINITY;
      if (n > p.max_active) max_cut = select_kth(tok_cost + tb, n, p.max_active, hist, sh);
      if (max_cut < beam_cutoff) {
        adaptive_beam = __fadd_rn(__fsub_rn(max_cut, best), p.beam_delta);
        cur_cutoff = max_cut;
      } else {
        if (n > p.min_active) min_cut = p.min_active == 0 ? best : select_kth(tok_cost + tb, n, p.min_active, hist, sh);
        if (min_cut > beam_cutoff) {
          adaptive_beam = __fadd_rn(__fsub_rn(min_cut, best), p.beam_delta);
          cur_cutoff = min_cut;
        } else {
          adaptive_beam = p.beam;
          cur_cutoff = beam_cutoff;
        }
      }
    }
    const float cost_offset = n > 0 ? -best : 0.0f;
    // ---- prepass on the best token (first one holding the minimum), its own association of the sum (:756-774)
    if (threadIdx.x == 0) s_prep = INFINITY;
    __syncthreads();
    {
      int cand = 0x7fffffff;
      for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS)
        if (tok_cost[i] == best) { cand = i; break; }
      for (int o = 16; o; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
      __shared__ int s_cand[DEC_THREADS / 32];
      if ((threadIdx.x & 31) == 0) s_cand[threadIdx.x >> 5] = cand;
      __syncthreads();
      if (threadIdx.x == 0) {
        int b = s_cand[0];
        for (int i = 1; i < DEC_THREADS / 32; ++i) b = min(b, s_cand[i]);
        float nc = INFINITY;
        if (b != 0x7fffffff) {
          const int st = tok_state[b];
          const float tc = tok_cost[b];
          for (long long a = p.off[st]; a < p.off[st + 1]; ++a) {
            const DArc arc = p.arcs[a];
            if (arc.il == 0) continue;
            const float like = __fmul_rn(p.acoustic_scale, ll[arc.il - 1]);
            float nw = __fadd_rn(__fsub_rn(__fadd_rn(arc.w, cost_offset), like), tc);
            if (st != arc.next) nw = __fadd_rn(nw, p.length_penalty);
            nc = fminf(nc, __fadd_rn(nw, adaptive_beam));
          }
        }
        s_prep = nc;
      }
      __syncthreads();
    }
    // ---- 2. pass A: exact best cost of the next frame
    v = INFINITY;
    for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
      const float tc = tok_cost[i];
      if (!(tc <= cur_cutoff)) continue;
      const int st = tok_state[i];
      for (long long a = p.off[st]; a < p.off[st + 1]; ++a) {
        const DArc arc = p.arcs[a];
        if (arc.il == 0) continue;
        const float ac = __fsub_rn(cost_offset, __fmul_rn(p.acoustic_scale, ll[arc.il - 1]));
        const float g = st != arc.next ? __fadd_rn(arc.w, p.length_penalty) : arc.w;
        v = fminf(v, __fadd_rn(__fadd_rn(tc, ac), g));
      }
    }
    const float next_best = block_min(v, sred);
    const float next_cutoff = fminf(s_prep, __fadd_rn(next_best, adaptive_beam));
    // ---- 3. pass B: links + recombination
    const int tb_new = counters[0], lb_new = counters[1];
    __syncthreads();
    for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
      const float tc = tok_cost[i];
      if (!(tc <= cur_cutoff)) continue;
      const int st = tok_state[i];
      for (long long a = p.off[st]; a < p.off[st + 1]; ++a) {
        const DArc arc = p.arcs[a];
        if (arc.il == 0) continue;
        const float ac = __fsub_rn(cost_offset, __fmul_rn(p.acoustic_scale, ll[arc.il - 1]));
        const float g = st != arc.next ? __fadd_rn(arc.w, p.length_penalty) : arc.w;
        const float tot = __fadd_rn(__fadd_rn(tc, ac), g);
        if (tot >= next_cutoff) continue;
        const int li = atomicAdd(&counters[1], 1);
        if (li >= p.link_cap) { counters[3] = 2; continue; }
        const unsigned long long key = ((unsigned long long)fkey(tot) << 32) | (unsigned int)a;
        const unsigned long long old = atomicMin(&best64[arc.next], key);
        if (old == EMPTY64) {
          const int ns = atomicAdd(&counters[0], 1);
          if (ns < p.tok_cap) { tok_state[ns] = arc.next; tokidx[arc.next] = ns; }
          else counters[3] = 1;
        }
        links[li] = DLink{i, arc.next /* state for now */, arc.il, arc.ol, g, ac};
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { counters[0] = min(counters[0], p.tok_cap); counters[1] = min(counters[1], p.link_cap); }
    __syncthreads();
    // destination state -> token index
    for (int l = lb_new + threadIdx.x; l < counters[1]; l += DEC_THREADS) links[l].dst = tokidx[links[l].dst];
    __syncthreads();
    // ---- 4. epsilon closure + 5. finish
    if (threadIdx.x == 0) coff[frame] = cost_offset;
    eps_closure(p, best64, tokidx, tok_state, dirty, links, counters, tb_new, next_cutoff, &s_flag);
    // emitting links carry the frame's source costs; epsilon links the converged ones: tok_cost of sources of
    // emitting links is final already (previous frame)
    finish_frame(p, best64, tokidx, tok_state, tok_cost, tok_bp, links, counters, tb_new, lb_new);
    if (threadIdx.x == 0) { ftok[frame + 2] = counters[0]; flink[frame + 2] = counters[1]; counters[2] = frame + 1; }
    __syncthreads();
  }
}

#include "decode_strict.cuh"

// ------------------------------------------------------------------------------------------------ lattice pruning on the GPU
// FinalizeDecoding / PruneForwardLinksFinal / PruneForwardLinks / PruneTokensForFrame (lattice-faster-decoder.cc:632-647,
// 379-465, 298-373, 485-506) for a finished utterance: per-token extra cost (cost of the best complete path through the token
// minus the best path overall) by a backward sweep over the frame levels, links and tokens beyond lattice_beam dropped, the
// survivors compacted in place (stable, so creation order -- which the host n-best uses for ties -- is preserved).
// Only the survivors (a few per cent of the lattice at max_active = 7000) cross PCIe afterwards.
//
// Level f holds the tokens [ftok[f], ftok[f+1]).  Links are stored in creation order: range [flink[f], flink[f+1]) holds
// the epsilon links inside level f preceded by the emitting links from level f-1 into level f.
// The reference iterates extra costs from 0 upwards until nothing changes; the token lattice is acyclic, so the fixpoint is
// unique and is reached here from above (atomicMin on the bit pattern of non-negative floats).
struct PruneParams {
  const float* fin;              // [nstates] final costs (inf = not final)
  float lattice_beam;
  int* tok_state; float* tok_cost; DLink* links;      // per-slot pools (compacted in place)
  unsigned int* extra;           // [slots][tok_cap] scratch: extra cost bits
  int* newidx;                   // [slots][tok_cap] scratch
  const int* frame_tok_off; const int* frame_link_off; const int* counters;
  int* c_ftok;                   // [slots][max_frames + 3] compacted level offsets
  int* c_counts;                 // [slots][2] survivors: tokens, links
  const int* slot_ids;
  int tok_cap, link_cap, max_frames;
};

__device__ inline float link_extra(const float* cost, const unsigned int* extra, const DLink& k) {
  // extra[dst] + ((cost[src] + ac + graph) - cost[dst]), same association as the reference
  return __fadd_rn(__uint_as_float(extra[k.dst]), __fsub_rn(__fadd_rn(__fadd_rn(cost[k.src], k.ac), k.graph), cost[k.dst]));
}

// exclusive offsets of the set flags inside the block, in thread order; returns the block total
__device__ int block_rank(bool flag, int* rank, int* s_warp) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = __popc(m & ((1u << lane) - 1u));
  __syncthreads();
  if (lane == 0) s_warp[w] = __popc(m);
  __syncthreads();
  int base = 0, total = 0;
  for (int i = 0; i < DEC_THREADS / 32; ++i) { if (i < w) base += s_warp[i]; total += s_warp[i]; }
  *rank = base + r;
  return total;
}

__global__ void __launch_bounds__(DEC_THREADS, 1)
lattice_prune_kernel(const PruneParams p) {
  __shared__ float sred[DEC_THREADS / 32];
  __shared__ int s_warp[DEC_THREADS / 32];
  __shared__ int s_flag;
  const int slot = p.slot_ids[blockIdx.x];
  int* tok_state = p.tok_state + (size_t)slot * p.tok_cap;
  float* cost = p.tok_cost + (size_t)slot * p.tok_cap;
  DLink* links = p.links + (size_t)slot * p.link_cap;
  unsigned int* extra = p.extra + (size_t)slot * p.tok_cap;
  int* newidx = p.newidx + (size_t)slot * p.tok_cap;
  const int* ftok = p.frame_tok_off + (size_t)slot * (p.max_frames + 3);
  const int* flink = p.frame_link_off + (size_t)slot * (p.max_frames + 3);
  int* c_ftok = p.c_ftok + (size_t)slot * (p.max_frames + 3);
  const int F = p.counters[slot * 4 + 2] + 1;             // levels
  const float lb = p.lattice_beam;
  const unsigned int INF_BITS = 0x7f800000u;

  // ---- last level: final costs
  const int fb = ftok[F - 1], fe = ftok[F];
  float bf = INFINITY, bp = INFINITY;
  for (int i = fb + threadIdx.x; i < fe; i += DEC_THREADS) {
    const float fc = p.fin[tok_state[i]];
    bp = fminf(bp, cost[i]);
    bf = fminf(bf, __fadd_rn(cost[i], fc));
  }
  const float best_final = block_min(bf, sred), best_plain = block_min(bp, sred);
  const bool any_final = best_final != INFINITY;
  const float final_best = any_final ? best_final : best_plain;

  for (int f = F - 1; f >= 0; --f) {
    const int tb = ftok[f], te = ftok[f + 1];
    for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
      float e0 = INFINITY;
      if (f == F - 1) {
        const float fc = any_final ? p.fin[tok_state[i]] : 0.0f;
        e0 = __fsub_rn(__fadd_rn(cost[i], fc), final_best);
        if (!(e0 <= lb)) e0 = INFINITY;                    // te > lattice_beam -> inf (also catches NaN from inf - inf)
      }
      extra[i] = __float_as_uint(e0);
    }
    __syncthreads();
    // emitting links out of level f (their destinations, level f + 1, are final already): one pass
    if (f < F - 1) {
      for (int l = flink[f + 1] + threadIdx.x; l < flink[f + 2]; l += DEC_THREADS) {
        const DLink k = links[l];
        if (k.il == 0) continue;
        float lec = link_extra(cost, extra, k);
        if (!(lec <= lb)) continue;
        if (lec < 0.0f) lec = 0.0f;
        atomicMin(&extra[k.src], __float_as_uint(lec));
      }
    }
    // epsilon links inside level f: relax until nothing changes
    for (;;) {
      __syncthreads();
      if (threadIdx.x == 0) s_flag = 0;
      __syncthreads();
      for (int l = flink[f] + threadIdx.x; l < flink[f + 1]; l += DEC_THREADS) {
        const DLink k = links[l];
        if (k.il != 0) continue;
        float lec = link_extra(cost, extra, k);
        if (!(lec <= lb)) continue;
        if (lec < 0.0f) lec = 0.0f;
        const unsigned int nb = __float_as_uint(lec);
        if (atomicMin(&extra[k.src], nb) > nb) s_flag = 1;
      }
      __syncthreads();
      if (!s_flag) break;
    }
  }
  __syncthreads();

  // ---- stable compaction of the surviving tokens, level by level
  int run = 0;
  for (int f = 0; f < F; ++f) {
    if (threadIdx.x == 0) c_ftok[f] = run;
    for (int base = ftok[f]; base < ftok[f + 1]; base += DEC_THREADS) {
      const int i = base + threadIdx.x;
      const bool in = i < ftok[f + 1];
      const bool alive = in && extra[i] != INF_BITS;
      int st = 0; float c = 0.f;
      if (alive) { st = tok_state[i]; c = cost[i]; }
      int rank;
      const int total = block_rank(alive, &rank, s_warp);
      if (in) newidx[i] = alive ? run + rank : -1;
      if (alive) { tok_state[run + rank] = st; cost[run + rank] = c; }       // run + rank <= i: in place is safe (reads of the chunk are done)
      run += total;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) { c_ftok[F] = run; p.c_counts[slot * 2] = run; }
  __syncthreads();
  // ---- links: alive iff both ends survive and the link itself stays inside the beam.  The costs were compacted above, so the
  //      link test uses the new indices; extra is still indexed by the old ones.
  int lrun = 0;
  const int nl = flink[F];
  for (int base = 0; base < nl; base += DEC_THREADS) {
    const int l = base + threadIdx.x;
    bool alive = false;
    DLink k;
    if (l < nl) {
      k = links[l];
      const int ns = newidx[k.src], nd = newidx[k.dst];
      if (ns >= 0 && nd >= 0) {
        const float lec = __fadd_rn(__uint_as_float(extra[k.dst]), __fsub_rn(__fadd_rn(__fadd_rn(cost[ns], k.ac), k.graph), cost[nd]));
        alive = lec <= lb;
        k.src = ns; k.dst = nd;
      }
    }
    int rank;
    const int total = block_rank(alive, &rank, s_warp);
    if (alive) links[lrun + rank] = k;
    lrun += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) p.c_counts[slot * 2 + 1] = lrun;
}

// ------------------------------------------------------------------------------------------------ best path on the GPU
// Online GetBestPath (lattice-faster-online-decoder.cc:59-177): best token of the last level (with or without final costs),
// back-pointers to the start, the cheapest matching link at every hop, weights multiplied from the start state forward.
// One CTA per slot; only the word ids and two floats cross PCIe (the partial result after every Decode() chunk).
constexpr int BP_MAX_PATH = 4096;
struct BestPathParams {
  const float* fin;
  const int* tok_state; const float* tok_cost; const int* tok_bp; const DLink* links;
  const int* frame_tok_off; const int* frame_link_off; const float* cost_offsets; const int* counters;
  const int* slot_ids;
  int* out_words;          // [slots][BP_MAX_PATH]
  float* out_cost;         // [slots][4]: graph, acoustic, ok (1 / 0), number of words
  float* scratch;          // [slots][BP_MAX_PATH][2] (graph, acoustic) of the hops, last hop first
  int* scratch_ol;         // [slots][BP_MAX_PATH]
  int tok_cap, link_cap, max_frames, use_final;
};

__global__ void __launch_bounds__(256, 1)
best_path_kernel(const BestPathParams p) {
  __shared__ unsigned long long s_key[8];
  __shared__ unsigned long long s_best;
  __shared__ int s_any;
  const int slot = p.slot_ids[blockIdx.x];
  const int* tok_state = p.tok_state + (size_t)slot * p.tok_cap;
  const float* cost = p.tok_cost + (size_t)slot * p.tok_cap;
  const int* bp = p.tok_bp + (size_t)slot * p.tok_cap;
  const DLink* links = p.links + (size_t)slot * p.link_cap;
  const int* ftok = p.frame_tok_off + (size_t)slot * (p.max_frames + 3);
  const int* flink = p.frame_link_off + (size_t)slot * (p.max_frames + 3);
  const float* coff = p.cost_offsets + (size_t)slot * p.max_frames;
  float* hop = p.scratch + (size_t)slot * BP_MAX_PATH * 2;
  int* hop_ol = p.scratch_ol + (size_t)slot * BP_MAX_PATH;
  int* words = p.out_words + (size_t)slot * BP_MAX_PATH;
  float* out = p.out_cost + slot * 4;
  const int F = p.counters[slot * 4 + 2] + 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // block-wide minimum of 64-bit keys
  auto block_min_key = [&](unsigned long long k) {
    for (int o = 16; o; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, k, o); if (x < k) k = x; }
    __syncthreads();
    if (lane == 0) s_key[w] = k;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long b = s_key[0]; for (int i = 1; i < 8; ++i) if (s_key[i] < b) b = s_key[i]; s_best = b; }
    __syncthreads();
    return s_best;
  };
  const int fb = ftok[F - 1], fe = ftok[F];
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  if (p.use_final)
    for (int i = fb + threadIdx.x; i < fe; i += 256)
      if (p.fin[tok_state[i]] != INFINITY) s_any = 1;
  __syncthreads();
  const bool any_final = p.use_final && s_any;
  // best token: smallest cost, the LAST created among equal costs (the reference walks the list in reverse creation order)
  unsigned long long key = ~0ull;
  for (int i = fb + threadIdx.x; i < fe; i += 256) {
    float c = cost[i];
    if (any_final) { const float fc = p.fin[tok_state[i]]; c = fc != INFINITY ? __fadd_rn(c, fc) : INFINITY; }
    if (c == INFINITY) continue;
    const unsigned long long k = ((unsigned long long)fkey(c) << 32) | (unsigned int)(0x7fffffff - i);
    if (k < key) key = k;
  }
  key = block_min_key(key);
  if (key == ~0ull) { if (threadIdx.x == 0) { out[0] = 0.f; out[1] = 0.f; out[2] = 0.f; out[3] = 0.f; } return; }
  int tok = 0x7fffffff - (int)(unsigned int)key;
  const float best_final = any_final ? p.fin[tok_state[tok]] : 0.0f;
  int level = F - 1, nh = 0;
  bool ok = true;
  while (tok >= 0 && ok) {
    const int b = bp[tok];
    float gc = 0.0f, ac = 0.0f;
    int ol = 0;
    if (b >= 0) {
      // every link INTO a level-f token is in [flink[f], flink[f+1]); cheapest (graph + ac), first created among equals
      unsigned long long lk = ~0ull;
      for (int l = flink[level] + threadIdx.x; l < flink[level + 1]; l += 256) {
        const DLink k = links[l];
        if (k.src != b || k.dst != tok) continue;
        const unsigned long long kk = ((unsigned long long)fkey(__fadd_rn(k.graph, k.ac)) << 32) | (unsigned int)l;
        if (kk < lk) lk = kk;
      }
      lk = block_min_key(lk);
      if (lk == ~0ull) ok = false;
      else {
        const DLink k = links[(unsigned int)lk];
        const int src_level = b >= ftok[level] ? level : level - 1;
        gc = k.graph; ol = k.ol;
        ac = k.il != 0 ? __fsub_rn(k.ac, coff[src_level]) : k.ac;
        if (b < ftok[level]) level -= 1;
      }
    }
    if (nh >= BP_MAX_PATH) ok = false;
    if (ok && threadIdx.x == 0) { hop[2 * nh] = gc; hop[2 * nh + 1] = ac; hop_ol[nh] = ol; }
    ++nh;
    tok = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tg = 0.0f, ta = 0.0f;
    int nw = 0;
    if (ok) {
      for (int i = nh - 1; i >= 0; --i) {
        tg = __fadd_rn(tg, hop[2 * i]); ta = __fadd_rn(ta, hop[2 * i + 1]);
        if (hop_ol[i] != 0) words[nw++] = hop_ol[i];
      }
      tg = __fadd_rn(tg, best_final);
    }
    out[0] = tg; out[1] = ta; out[2] = ok ? 1.0f : 0.0f; out[3] = (float)nw;
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct HostGraph {
  int start = -1;
  std::vector<float> fin;
  std::vector<long long> off;
  std::vector<DArc> arcs;
  std::vector<unsigned char> has_eps;
};

bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }
bool rd_str(FILE* f, std::string* s) {
  int32_t n;
  if (!rd(f, &n, 4) || n < 0 || n > (1 << 20)) return false;
  s->resize(n);
  return n == 0 || rd(f, &(*s)[0], n);
}
bool skip_symbols(FILE* f) {
  int32_t magic; std::string name; int64_t avail, size;
  if (!rd(f, &magic, 4) || !rd_str(f, &name) || !rd(f, &avail, 8) || !rd(f, &size, 8)) return false;
  for (int64_t i = 0; i < size; ++i) { std::string s; int64_t k; if (!rd_str(f, &s) || !rd(f, &k, 8)) return false; }
  return true;
}
// OpenFST binary "vector" FST with "standard" arcs (what fst::Fst<StdArc>::Read loads for TLG.fst)
int load_fst(const char* path, HostGraph* g) {
  FILE* f = fopen(path, "rb");
  if (!f) return dfail(B2T_ERR_ARG, "cannot open FST %s", path);
  int32_t magic = 0, version = 0, flags = 0; std::string ft, at; uint64_t props; int64_t start, ns, na;
  bool ok = rd(f, &magic, 4) && magic == 2125659606 && rd_str(f, &ft) && rd_str(f, &at) && rd(f, &version, 4) && rd(f, &flags, 4) &&
            rd(f, &props, 8) && rd(f, &start, 8) && rd(f, &ns, 8) && rd(f, &na, 8);
  if (!ok || ft != "vector" || at != "standard") { fclose(f); return dfail(B2T_ERR_UNSUPPORTED, "%s: only OpenFST vector/standard FSTs are supported", path); }
  if (((flags & 1) && !skip_symbols(f)) || ((flags & 2) && !skip_symbols(f))) { fclose(f); return dfail(B2T_ERR_ARG, "%s: bad embedded symbol table", path); }
  g->start = (int)start;
  g->fin.resize(ns); g->off.resize(ns + 1); g->has_eps.assign(ns, 0);
  g->arcs.clear(); g->arcs.reserve(na > 0 ? na : 0);
  for (int64_t s = 0; s < ns; ++s) {
    float fw; int64_t cnt;
    if (!rd(f, &fw, 4) || !rd(f, &cnt, 8)) { fclose(f); return dfail(B2T_ERR_ARG, "%s: truncated", path); }
    g->fin[s] = fw; g->off[s] = (long long)g->arcs.size();
    for (int64_t a = 0; a < cnt; ++a) {
      DArc arc;
      if (!rd(f, &arc, 16)) { fclose(f); return dfail(B2T_ERR_ARG, "%s: truncated", path); }
      if (arc.il == 0) g->has_eps[s] = 1;
      g->arcs.push_back(arc);
    }
  }
  g->off[ns] = (long long)g->arcs.size();
  fclose(f);
  if (g->arcs.size() >= (1ull << 32)) return dfail(B2T_ERR_UNSUPPORTED, "graphs with >= 2^32 arcs are not supported yet");
  return 0;
}

struct HResult { float ac, lm; std::string sentence; };

struct Slot {
  std::vector<float> fed;           // fed frames [n][C]
  int n_fed = 0, num_frames = 0, last_best = 0;
  bool last_blank = false, decoded_any = false, finished = false;
  bool pruned_on_gpu = false;       // the device copy of the lattice has been pruned and compacted (finish or rescore)
  std::vector<float> last_frame;
  std::vector<HResult> results;
};

}  // namespace

struct b2t_decoder {
  HostGraph g;
  std::vector<std::string> words;
  b2t_decode_options opt;
  int max_frames, max_slots, C = 0;
  int tok_cap, link_cap;
  // device
  DArc* d_arcs = nullptr; long long* d_off = nullptr; unsigned char* d_has_eps = nullptr;
  unsigned long long* d_best = nullptr; int* d_tokidx = nullptr;
  int *d_tok_state = nullptr, *d_tok_bp = nullptr; float* d_tok_cost = nullptr; unsigned char* d_dirty = nullptr;
  DLink* d_links = nullptr;
  int *d_ftok = nullptr, *d_flink = nullptr, *d_counters = nullptr, *d_nfed = nullptr, *d_slot_ids = nullptr;
  float *d_coff = nullptr, *d_logp = nullptr;
  float* d_fin = nullptr; unsigned int* d_extra = nullptr; int *d_newidx = nullptr, *d_cftok = nullptr, *d_ccounts = nullptr;   // GPU lattice pruning
  int* d_bp_words = nullptr; float *d_bp_cost = nullptr, *d_bp_hop = nullptr; int* d_bp_ol = nullptr;                                                    // GPU best path
  // strict serial-order mode (decode_strict.cuh), allocated on first use
  bool strict = false;
  int strict_fc = 0, strict_qcap = 0, strict_hcap = 0;
  long long* d_eps_off = nullptr; EArc* d_eps_arcs = nullptr;
  int *d_first_rank = nullptr, *d_eps_bp = nullptr, *d_bucket_min = nullptr, *d_s_cum = nullptr, *d_s_v0 = nullptr, *d_s_v1 = nullptr, *d_s_byrank = nullptr,
      *d_s_queue = nullptr, *d_s_hash = nullptr;
  float* d_s_premin = nullptr; unsigned int *d_s_k0 = nullptr, *d_s_k1 = nullptr;
  std::vector<Slot> slots;
  cudaStream_t stream = nullptr;
  double last_kernel_ms = 0.0;
  std::atomic<float> nbest_margin_hint{1.0f};   // cost margin that sufficed for the last n-best extraction (collect_nbest starts there)
  void *lm_old = nullptr, *lm_new = nullptr;   // LmAcceptor*: the LM the graph was built from / the rescoring LM (DecodeResource lm_fst_path, rescore_lm_fst_path)
};

namespace {

int read_words(const char* path, std::vector<std::string>* w) {
  FILE* f = fopen(path, "r");
  if (!f) return dfail(B2T_ERR_ARG, "cannot open symbol table %s", path);
  char buf[1024]; int id;
  while (fscanf(f, "%1023s %d", buf, &id) == 2) {
    if (id < 0) continue;
    if ((size_t)id >= w->size()) w->resize(id + 1);
    (*w)[id] = buf;
  }
  fclose(f);
  return 0;
}

std::string process_blank(const std::string& s) {   // utils/string.cc:121-146
  std::string r;
  for (char c : s) {
    if (c != ' ') r.push_back(c);
    else if (!r.empty() && r.back() != ' ') r.push_back(' ');
  }
  if (!r.empty() && r.back() == ' ') r.pop_back();
  for (char& c : r) c = (char)tolower((unsigned char)c);
  return r;
}

void push_result(b2t_decoder* d, Slot& s, const std::vector<int>& w, float graph, float acoustic) {   // brain_speech_decoder.cc:113-137
  HResult r;
  r.lm = -graph;
  r.ac = -acoustic / d->opt.acoustic_scale;
  for (int id : w) r.sentence += ' ' + (id >= 0 && id < (int)d->words.size() ? d->words[id] : std::string("<unk>"));
  r.sentence = process_blank(r.sentence);
  s.results.push_back(r);
}

// ---- host view of one slot's token lattice
struct Lattice {
  int F = 0;                               // number of frame_plus_one levels with tokens (frames decoded + 1)
  std::vector<int> ftok, flink;            // offsets per level (size F + 1)
  std::vector<int> state, bp; std::vector<float> cost;
  std::vector<DLink> links;
  std::vector<float> coff;
  std::vector<int> lbeg;                   // links sorted by src: lbeg[tok] .. lbeg[tok+1]
  std::vector<int> frame_of;
};

int fetch_lattice(b2t_decoder* d, int slot, Lattice* L) {
  int counters[4];
  DCK(cudaMemcpyAsync(counters, d->d_counters + slot * 4, sizeof(counters), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaStreamSynchronize(d->stream));
  if (counters[3] == 1) return dfail(B2T_ERR_WORKSPACE, "decoder token pool overflow (capacity %d): raise max_active-derived capacity or lower beam", d->tok_cap);
  if (counters[3] == 2) return dfail(B2T_ERR_WORKSPACE, "decoder link pool overflow (capacity %d)", d->link_cap);
  if (counters[3] == 3) return dfail(B2T_ERR_WORKSPACE, "strict-order decoder scratch overflow (frame capacity %d tokens)", d->strict_fc);
  const int nf = counters[2], nt = counters[0], nl = counters[1];
  L->F = nf + 1;
  L->ftok.resize(L->F + 1); L->flink.resize(L->F + 1);
  DCK(cudaMemcpyAsync(L->ftok.data(), d->d_ftok + (size_t)slot * (d->max_frames + 3), (L->F + 1) * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaMemcpyAsync(L->flink.data(), d->d_flink + (size_t)slot * (d->max_frames + 3), (L->F + 1) * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  L->state.resize(nt); L->bp.resize(nt); L->cost.resize(nt); L->links.resize(nl); L->coff.resize(nf);
  DCK(cudaMemcpyAsync(L->state.data(), d->d_tok_state + (size_t)slot * d->tok_cap, nt * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaMemcpyAsync(L->bp.data(), d->d_tok_bp + (size_t)slot * d->tok_cap, nt * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaMemcpyAsync(L->cost.data(), d->d_tok_cost + (size_t)slot * d->tok_cap, nt * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  if (nl) DCK(cudaMemcpyAsync(L->links.data(), d->d_links + (size_t)slot * d->link_cap, nl * sizeof(DLink), cudaMemcpyDeviceToHost, d->stream));
  if (nf) DCK(cudaMemcpyAsync(L->coff.data(), d->d_coff + (size_t)slot * d->max_frames, nf * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaStreamSynchronize(d->stream));
  // order links by source token (stable counting sort) and index them
  L->lbeg.assign(nt + 1, 0);
  for (const DLink& k : L->links) L->lbeg[k.src + 1]++;
  for (int i = 0; i < nt; ++i) L->lbeg[i + 1] += L->lbeg[i];
  {
    std::vector<int> pos(L->lbeg.begin(), L->lbeg.end() - 1);
    std::vector<DLink> sorted(L->links.size());
    for (const DLink& k : L->links) sorted[pos[k.src]++] = k;
    L->links.swap(sorted);
  }
  L->frame_of.resize(nt);
  for (int f = 0; f < L->F; ++f)
    for (int i = L->ftok[f]; i < L->ftok[f + 1]; ++i) L->frame_of[i] = f;
  return 0;
}

inline float link_ac(const Lattice& L, const DLink& k) {   // acoustic cost with the per-frame offset removed (GetRawLattice :164-171)
  return k.il != 0 ? k.ac - L.coff[L.frame_of[k.src]] : k.ac;
}

// back-pointer best path (lattice-faster-online-decoder.cc:59-177)
bool best_path(const b2t_decoder* d, const Lattice& L, bool use_final, std::vector<int>* words, float* graph, float* acoustic) {
  words->clear();
  const int fb = L.ftok[L.F - 1], fe = L.ftok[L.F];
  bool any_final = false;
  if (use_final)
    for (int i = fb; i < fe; ++i)
      if (d->g.fin[L.state[i]] != INFINITY) any_final = true;
  float best_cost = INFINITY, best_final = 0.0f;
  int best = -1;
  // the reference walks the frame's token list, which is in reverse creation order: the LAST created token among
  // equal costs is seen first and wins; iterate descending to reproduce that
  for (int i = fe - 1; i >= fb; --i) {
    float cost = L.cost[i], fc = 0.0f;
    if (use_final && any_final) {
      fc = d->g.fin[L.state[i]];
      cost = fc != INFINITY ? cost + fc : INFINITY;
    }
    if (cost < best_cost) { best_cost = cost; best = i; best_final = fc; }
  }
  if (best < 0) return false;
  std::vector<std::pair<float, float>> ws;
  std::vector<int> rw;
  int tok = best;
  while (tok >= 0) {
    const int b = L.bp[tok];
    float gc = 0.0f, ac = 0.0f;
    int ol = 0;
    if (b >= 0) {
      float best_link = INFINITY;
      for (int l = L.lbeg[b]; l < L.lbeg[b + 1]; ++l) {
        const DLink& k = L.links[l];
        if (k.dst != tok) continue;
        const float c = k.graph + k.ac;
        if (c < best_link) { best_link = c; gc = k.graph; ac = link_ac(L, k); ol = k.ol; }
      }
      if (best_link == INFINITY) return false;
    }
    ws.push_back({gc, ac});
    if (ol != 0) rw.push_back(ol);
    tok = b;
  }
  float tg = 0.0f, ta = 0.0f;
  for (int i = (int)ws.size() - 1; i >= 0; --i) { tg = tg + ws[i].first; ta = ta + ws[i].second; }
  tg = tg + best_final;
  std::reverse(rw.begin(), rw.end());
  *words = rw; *graph = tg; *acoustic = ta;
  return true;
}

// FinalizeDecoding (:632-647): exact backward pruning with final costs.  Returns per-token extra cost (inf = pruned)
// and a keep flag per link.
void prune_final(const b2t_decoder* d, const Lattice& L, std::vector<float>* extra_out, std::vector<char>* keep_out) {
  const int nt = (int)L.state.size();
  const float lb = d->opt.lattice_beam;
  std::vector<float> extra(nt, 0.0f);
  std::vector<char> keep(L.links.size(), 1);
  const int fb = L.ftok[L.F - 1], fe = L.ftok[L.F];
  bool any_final = false;
  float best_final = INFINITY, best_plain = INFINITY;
  for (int i = fb; i < fe; ++i) {
    const float fc = d->g.fin[L.state[i]];
    if (fc != INFINITY) any_final = true;
    best_plain = std::min(best_plain, L.cost[i]);
    best_final = std::min(best_final, L.cost[i] + fc);
  }
  const float final_best = best_final != INFINITY ? best_final : best_plain;
  auto sweep = [&](int f, bool last) {
    bool changed = true;
    while (changed) {
      changed = false;
      for (int i = L.ftok[f + 1] - 1; i >= L.ftok[f]; --i) {     // the reference walks tokens in reverse creation order
        float te;
        if (last) {
          const float fc = any_final ? d->g.fin[L.state[i]] : 0.0f;
          te = L.cost[i] + fc - final_best;
        } else te = INFINITY;
        for (int l = L.lbeg[i]; l < L.lbeg[i + 1]; ++l) {
          if (!keep[l]) continue;
          const DLink& k = L.links[l];
          float lec = extra[k.dst] + ((L.cost[i] + k.ac + k.graph) - L.cost[k.dst]);
          if (lec > lb) keep[l] = 0;
          else {
            if (lec < 0.0f) lec = 0.0f;
            if (lec < te) te = lec;
          }
        }
        if (last && te > lb) te = INFINITY;
        bool diff;
        if (last) {      // !ApproxEqual(old, new, 1e-5)   (kaldi-math.h)
          const float o = extra[i], df = fabsf(o - te);
          diff = !(o == te || (df != INFINITY && df == df && df <= 1.0e-05f * (fabsf(o) + fabsf(te))));
        } else {         // fabs(new - old) > delta with delta == 0 (NaN from inf - inf compares false)
          diff = fabsf(te - extra[i]) > 0.0f;
        }
        if (diff) changed = true;
        extra[i] = te;
      }
    }
  };
  sweep(L.F - 1, true);
  for (int f = L.F - 2; f >= 0; --f) sweep(f, false);
  *extra_out = extra; *keep_out = keep;
}

// n cheapest distinct word sequences over the pruned lattice: per-token top-K distinct
// word-sequence hypotheses in topological order; exact for K >= n.
struct WordTrie {
  std::vector<std::pair<int, int>> node{{-1, 0}};
  std::unordered_map<unsigned long long, int> idx;
  int extend(int n, int w) {
    const unsigned long long key = ((unsigned long long)(unsigned int)n << 32) | (unsigned int)w;
    auto it = idx.find(key);
    if (it != idx.end()) return it->second;
    node.push_back({n, w});
    return idx[key] = (int)node.size() - 1;
  }
  std::vector<int> words(int n) const {
    std::vector<int> w;
    for (; n > 0; n = node[n].first) w.push_back(node[n].second);
    std::reverse(w.begin(), w.end());
    return w;
  }
};
struct Hyp { int seq; float g, a; };
inline bool hyp_better(const Hyp& x, const Hyp& y) {
  const float fx = x.g + x.a, fy = y.g + y.a;
  if (fx != fy) return fx < fy;
  return x.g < y.g;
}

struct NbOut { std::vector<int> words; float g, a; };
// The K cheapest distinct word sequences of a pruned lattice with the (graph, acoustic) costs of their best alignments.
void collect_nbest(b2t_decoder* d, const Lattice& L, const std::vector<float>& extra, const std::vector<char>& keep, int K, std::vector<NbOut>* out) {
  const int nt = (int)L.state.size();
  K = std::max(1, K);
  // topological order inside each level with respect to epsilon links (Kahn)
  std::vector<int> order; order.reserve(nt);
  std::vector<int> indeg(nt, 0);
  for (size_t l = 0; l < L.links.size(); ++l)
    if (keep[l] && L.links[l].il == 0 && extra[L.links[l].src] != INFINITY && extra[L.links[l].dst] != INFINITY) indeg[L.links[l].dst]++;
  for (int f = 0; f < L.F; ++f) {
    const size_t first = order.size();
    for (int i = L.ftok[f]; i < L.ftok[f + 1]; ++i)
      if (extra[i] != INFINITY && indeg[i] == 0) order.push_back(i);
    for (size_t q = first; q < order.size(); ++q) {
      const int i = order[q];
      for (int l = L.lbeg[i]; l < L.lbeg[i + 1]; ++l)
        if (keep[l] && L.links[l].il == 0 && extra[L.links[l].dst] != INFINITY && --indeg[L.links[l].dst] == 0) order.push_back(L.links[l].dst);
    }
  }
  bool any_final = false;
  for (int i = L.ftok[L.F - 1]; i < L.ftok[L.F]; ++i)
    if (d->g.fin[L.state[i]] != INFINITY) any_final = true;
  auto alive = [&](int l) { return keep[l] && extra[L.links[l].dst] != INFINITY; };
  // exact cost-to-final of every surviving token (un-offset costs), reverse topological order
  std::vector<float> beta(nt, INFINITY);
  for (int i = L.ftok[L.F - 1]; i < L.ftok[L.F]; ++i)
    if (extra[i] != INFINITY) beta[i] = any_final ? d->g.fin[L.state[i]] : 0.0f;
  for (int q = (int)order.size() - 1; q >= 0; --q) {
    const int i = order[q];
    float bi = beta[i];
    for (int l = L.lbeg[i]; l < L.lbeg[i + 1]; ++l) {
      if (!alive(l)) continue;
      const DLink& k = L.links[l];
      const float c = k.graph + link_ac(L, k) + beta[k.dst];
      if (c < bi) bi = c;
    }
    beta[i] = bi;
  }
  if (order.empty() || beta[0] == INFINITY) return;
  // A hypothesis is dropped as soon as even its best completion exceeds `limit` (admissible: beta is exact).  The n best
  // sequences usually lie within a fraction of the lattice beam, so the search runs with a small margin first and widens it
  // until the K-th result is inside the margin (then nothing that was dropped could have ranked among the first K) or the
  // margin is the whole lattice beam.
  const float full = d->opt.lattice_beam;
  auto dedupe_trim = [&](std::vector<Hyp>& v, int cap) {
    // candidates are appended unmerged; one sort keeps the best hypothesis of every word sequence (the earliest inserted
    // among equals) and then the `cap` cheapest
    if (v.size() > 1) {
      std::stable_sort(v.begin(), v.end(), [](const Hyp& x, const Hyp& y) { return x.seq != y.seq ? x.seq < y.seq : hyp_better(x, y); });
      size_t w = 0;
      for (size_t r = 0; r < v.size(); ++r)
        if (r == 0 || v[r].seq != v[r - 1].seq) v[w++] = v[r];
      v.resize(w);
    }
    if (cap > 0 && (int)v.size() > cap) {
      std::partial_sort(v.begin(), v.begin() + cap, v.end(), hyp_better);
      v.resize(cap);
    }
  };
  // (the first margin is the one that sufficed for the previous utterances of this decoder, so that the usual case is one pass)
  const float hint = d->nbest_margin_hint.load(std::memory_order_relaxed);
  int passes = 0;
  for (float margin = std::min(std::max(hint, 1.0f), full);; margin = std::min(margin * 2.0f, full)) {
    ++passes;
    const float limit = beta[0] + margin + 1e-4f;
    WordTrie trie;
    std::vector<std::vector<Hyp>> hyps(nt);
    hyps[0].push_back({0, 0.0f, 0.0f});
    std::vector<Hyp> finals;
    for (int i : order) {
      std::vector<Hyp>& hv = hyps[i];
      if (hv.empty()) continue;
      dedupe_trim(hv, K);
      if (L.frame_of[i] == L.F - 1) {
        const float fc = any_final ? d->g.fin[L.state[i]] : 0.0f;
        if (fc != INFINITY)
          for (const Hyp& h : hv)
            if (h.g + fc + h.a <= limit) finals.push_back(Hyp{h.seq, h.g + fc, h.a});
      }
      for (int l = L.lbeg[i]; l < L.lbeg[i + 1]; ++l) {
        if (!alive(l)) continue;
        const DLink& k = L.links[l];
        const float ac = link_ac(L, k), bn = beta[k.dst];
        if (bn == INFINITY) continue;
        std::vector<Hyp>& dv = hyps[k.dst];
        for (const Hyp& h : hv) {
          const float g = h.g + k.graph, a = h.a + ac;
          if (g + a + bn > limit) continue;
          dv.push_back(Hyp{k.ol != 0 ? trie.extend(h.seq, k.ol) : h.seq, g, a});
        }
        if (dv.size() > (size_t)8 * K + 64) dedupe_trim(dv, 0);      // bound the memory of high in-degree tokens
      }
      std::vector<Hyp>().swap(hv);
    }
    dedupe_trim(finals, 0);
    std::stable_sort(finals.begin(), finals.end(), hyp_better);
    const bool complete = margin >= full || ((int)finals.size() >= K && finals[K - 1].g + finals[K - 1].a <= beta[0] + margin);
    if (!complete) continue;
    for (size_t i = 0; i < finals.size() && (int)i < K; ++i) out->push_back(NbOut{trie.words(finals[i].seq), finals[i].g, finals[i].a});
    d->nbest_margin_hint.store(passes == 1 ? std::max(1.0f, margin * 0.8f) : margin, std::memory_order_relaxed);
    break;
  }
}

void nbest_from_lattice(b2t_decoder* d, Slot& s, const Lattice& L, const std::vector<float>& extra, const std::vector<char>& keep) {
  std::vector<NbOut> nb;
  collect_nbest(d, L, extra, keep, d->opt.nbest, &nb);
  for (const NbOut& e : nb) push_result(d, s, e.words, e.g, e.a);
}

void log_softmax_row(const float* x, int C, float* out) {
  float m = x[0];
  for (int c = 1; c < C; ++c) m = std::max(m, x[c]);
  double s = 0.0;
  for (int c = 0; c < C; ++c) s += std::exp((double)(x[c] - m));
  const float ls = (float)std::log(s);
  for (int c = 0; c < C; ++c) out[c] = (x[c] - m) - ls;
}

// CtcWfstBeamSearch::Search frame selection (ctc_wfst_beam_search.cc:70-111): which frames are fed to the search
int feed_frames(b2t_decoder* d, Slot& s, const float* logp, int T, int C) {
  for (int i = 0; i < T; ++i) {
    const float* row = logp + (size_t)i * C;
    const float blank_score = std::exp(row[0]);
    if (blank_score > d->opt.blank_skip_threshold) {
      s.last_blank = true;
      s.last_frame.assign(row, row + C);
    } else {
      int cur_best = 0;
      for (int c = 1; c < C; ++c)
        if (row[c] > row[cur_best]) cur_best = c;
      if (cur_best != 0 && s.last_blank && cur_best == s.last_best) {
        if (s.n_fed >= d->max_frames) return dfail(B2T_ERR_ARG, "utterance longer than the decoder's max_frames (%d)", d->max_frames);
        s.fed.insert(s.fed.end(), s.last_frame.begin(), s.last_frame.end());
        s.n_fed++;
      }
      s.last_best = cur_best;
      if (s.n_fed >= d->max_frames) return dfail(B2T_ERR_ARG, "utterance longer than the decoder's max_frames (%d)", d->max_frames);
      s.fed.insert(s.fed.end(), row, row + C);
      s.n_fed++;
      s.last_blank = false;
    }
    s.num_frames++;
  }
  return 0;
}

int launch_slots(b2t_decoder* d, const std::vector<int>& ids) {
  if (ids.empty()) return 0;
  std::vector<int> nfed(d->max_slots, 0);
  for (int i = 0; i < d->max_slots; ++i) nfed[i] = d->slots[i].n_fed;
  DCK(cudaMemcpyAsync(d->d_nfed, nfed.data(), d->max_slots * sizeof(int), cudaMemcpyHostToDevice, d->stream));
  DCK(cudaMemcpyAsync(d->d_slot_ids, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice, d->stream));
  DecParams p;
  p.arcs = d->d_arcs; p.off = d->d_off; p.has_eps = d->d_has_eps; p.nstates = (int)d->g.fin.size(); p.start = d->g.start;
  p.beam = d->opt.beam; p.beam_delta = 0.5f; p.acoustic_scale = d->opt.acoustic_scale; p.length_penalty = d->opt.length_penalty;
  p.max_active = d->opt.max_active; p.min_active = d->opt.min_active;
  p.best64 = d->d_best; p.tokidx = d->d_tokidx; p.tok_state = d->d_tok_state; p.tok_cost = d->d_tok_cost; p.tok_bp = d->d_tok_bp;
  p.tok_dirty = d->d_dirty; p.links = d->d_links; p.frame_tok_off = d->d_ftok; p.frame_link_off = d->d_flink; p.cost_offsets = d->d_coff;
  p.counters = d->d_counters; p.logp = d->d_logp; p.n_fed = d->d_nfed; p.slot_ids = d->d_slot_ids;
  p.tok_cap = d->tok_cap; p.link_cap = d->link_cap; p.max_frames = d->max_frames; p.C = d->C;
  p.eps_off = d->d_eps_off; p.eps_arcs = d->d_eps_arcs; p.first_rank = d->d_first_rank; p.eps_bp = d->d_eps_bp; p.bucket_min = d->d_bucket_min;
  p.s_cum = d->d_s_cum; p.s_premin = d->d_s_premin; p.s_k0 = d->d_s_k0; p.s_k1 = d->d_s_k1; p.s_v0 = d->d_s_v0; p.s_v1 = d->d_s_v1;
  p.s_byrank = d->d_s_byrank; p.s_queue = d->d_s_queue; p.s_hash = d->d_s_hash; p.fc = d->strict_fc; p.qcap = d->strict_qcap; p.hcap = d->strict_hcap;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, d->stream);
  if (d->strict) wfst_decode_strict_kernel<<<(int)ids.size(), DEC_THREADS, 0, d->stream>>>(p);
  else wfst_decode_kernel<<<(int)ids.size(), DEC_THREADS, 0, d->stream>>>(p);
  cudaEventRecord(e1, d->stream);
  DCK(cudaGetLastError());
  DCK(cudaStreamSynchronize(d->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  d->last_kernel_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

int upload_fed(b2t_decoder* d, int slot, int from_frame) {
  Slot& s = d->slots[slot];
  if (s.n_fed > from_frame)
    DCK(cudaMemcpyAsync(d->d_logp + ((size_t)slot * d->max_frames + from_frame) * d->C, s.fed.data() + (size_t)from_frame * d->C,
                        (size_t)(s.n_fed - from_frame) * d->C * sizeof(float), cudaMemcpyHostToDevice, d->stream));
  return 0;
}

int reset_slot(b2t_decoder* d, int slot) {
  Slot& s = d->slots[slot];
  s = Slot();
  const int init[4] = {0, 0, -1, 0};
  DCK(cudaMemcpyAsync(d->d_counters + slot * 4, init, sizeof(init), cudaMemcpyHostToDevice, d->stream));
  return 0;
}

bool use_gpu_prune() {   // A/B switch (read at every call so that tests can flip it): host restatement of pruning / best path
  const char* v = getenv("B2T_DECODER_HOST_PRUNE");
  return !(v && atoi(v) != 0);
}

// 1-best of a slot through best_path_kernel; returns 1 when a path was found
int best_path_on_gpu(b2t_decoder* d, int slot, bool use_final, std::vector<int>* words, float* graph, float* acoustic) {
  int counters[4];
  DCK(cudaMemcpyAsync(counters, d->d_counters + slot * 4, sizeof(counters), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaMemcpyAsync(d->d_slot_ids, &slot, sizeof(int), cudaMemcpyHostToDevice, d->stream));
  DCK(cudaStreamSynchronize(d->stream));
  if (counters[3] == 1) return dfail(B2T_ERR_WORKSPACE, "decoder token pool overflow (capacity %d): raise max_active-derived capacity or lower beam", d->tok_cap);
  if (counters[3] == 2) return dfail(B2T_ERR_WORKSPACE, "decoder link pool overflow (capacity %d)", d->link_cap);
  if (counters[3] == 3) return dfail(B2T_ERR_WORKSPACE, "strict-order decoder scratch overflow (frame capacity %d tokens)", d->strict_fc);
  BestPathParams p;
  p.fin = d->d_fin; p.tok_state = d->d_tok_state; p.tok_cost = d->d_tok_cost; p.tok_bp = d->d_tok_bp; p.links = d->d_links;
  p.frame_tok_off = d->d_ftok; p.frame_link_off = d->d_flink; p.cost_offsets = d->d_coff; p.counters = d->d_counters; p.slot_ids = d->d_slot_ids;
  p.out_words = d->d_bp_words; p.out_cost = d->d_bp_cost; p.scratch = d->d_bp_hop; p.scratch_ol = d->d_bp_ol;
  p.tok_cap = d->tok_cap; p.link_cap = d->link_cap; p.max_frames = d->max_frames; p.use_final = use_final ? 1 : 0;
  best_path_kernel<<<1, 256, 0, d->stream>>>(p);
  DCK(cudaGetLastError());
  float oc[4];
  DCK(cudaMemcpyAsync(oc, d->d_bp_cost + slot * 4, sizeof(oc), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaStreamSynchronize(d->stream));
  if (oc[2] == 0.0f) return 0;
  const int nw = (int)oc[3];
  words->resize(nw);
  if (nw) DCK(cudaMemcpyAsync(words->data(), d->d_bp_words + (size_t)slot * BP_MAX_PATH, nw * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaStreamSynchronize(d->stream));
  *graph = oc[0]; *acoustic = oc[1];
  return 1;
}

int partial_result(b2t_decoder* d, int slot) {   // ctc_wfst_beam_search.cc:112-120: 1-best without final costs after every chunk
  Slot& s = d->slots[slot];
  s.results.clear();
  if (s.n_fed == 0) return 0;
  std::vector<int> w; float g, a;
  if (use_gpu_prune()) {                       // (the same A/B switch selects the host restatement of the best path)
    const int rc = best_path_on_gpu(d, slot, false, &w, &g, &a);
    if (rc < 0) return rc;
    if (rc > 0) push_result(d, s, w, g, a);
    return 0;
  }
  Lattice L;
  int rc = fetch_lattice(d, slot, &L);
  if (rc) return rc;
  if (best_path(d, L, false, &w, &g, &a)) push_result(d, s, w, g, a);
  return 0;
}

// Run lattice_prune_kernel for the given (finished) slots; afterwards the slots' token / link pools hold only the survivors.
int prune_slots_on_gpu(b2t_decoder* d, const std::vector<int>& ids) {
  if (ids.empty()) return 0;
  DCK(cudaMemcpyAsync(d->d_slot_ids, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice, d->stream));
  PruneParams p;
  p.fin = d->d_fin; p.lattice_beam = d->opt.lattice_beam;
  p.tok_state = d->d_tok_state; p.tok_cost = d->d_tok_cost; p.links = d->d_links; p.extra = d->d_extra; p.newidx = d->d_newidx;
  p.frame_tok_off = d->d_ftok; p.frame_link_off = d->d_flink; p.counters = d->d_counters; p.c_ftok = d->d_cftok; p.c_counts = d->d_ccounts;
  p.slot_ids = d->d_slot_ids; p.tok_cap = d->tok_cap; p.link_cap = d->link_cap; p.max_frames = d->max_frames;
  lattice_prune_kernel<<<(int)ids.size(), DEC_THREADS, 0, d->stream>>>(p);
  DCK(cudaGetLastError());
  return 0;
}

// links ordered by source token (stable counting sort) + per-token link ranges and frame levels
void index_lattice(Lattice* L) {
  const int nt = (int)L->state.size();
  L->lbeg.assign(nt + 1, 0);
  for (const DLink& k : L->links) L->lbeg[k.src + 1]++;
  for (int i = 0; i < nt; ++i) L->lbeg[i + 1] += L->lbeg[i];
  {
    std::vector<int> pos(L->lbeg.begin(), L->lbeg.end() - 1);
    std::vector<DLink> sorted(L->links.size());
    for (const DLink& k : L->links) sorted[pos[k.src]++] = k;
    L->links.swap(sorted);
  }
  L->frame_of.resize(nt);
  for (int f = 0; f < L->F; ++f)
    for (int i = L->ftok[f]; i < L->ftok[f + 1]; ++i) L->frame_of[i] = f;
}

// Host view of a slot after prune_slots_on_gpu: every token and link is a survivor.  With index_now == false only the copies
// are done here (the caller indexes the lattice on a worker thread).
int fetch_pruned_lattice(b2t_decoder* d, int slot, Lattice* L, bool index_now = true) {
  int counters[4], cc[2];
  DCK(cudaMemcpyAsync(counters, d->d_counters + slot * 4, sizeof(counters), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaMemcpyAsync(cc, d->d_ccounts + slot * 2, sizeof(cc), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaStreamSynchronize(d->stream));
  if (counters[3] == 1) return dfail(B2T_ERR_WORKSPACE, "decoder token pool overflow (capacity %d): raise max_active-derived capacity or lower beam", d->tok_cap);
  if (counters[3] == 2) return dfail(B2T_ERR_WORKSPACE, "decoder link pool overflow (capacity %d)", d->link_cap);
  if (counters[3] == 3) return dfail(B2T_ERR_WORKSPACE, "strict-order decoder scratch overflow (frame capacity %d tokens)", d->strict_fc);
  const int nf = counters[2], nt = cc[0], nl = cc[1];
  L->F = nf + 1;
  L->ftok.resize(L->F + 1); L->flink.assign(L->F + 1, 0);
  L->state.resize(nt); L->bp.assign(nt, -1); L->cost.resize(nt); L->links.resize(nl); L->coff.resize(nf);
  DCK(cudaMemcpyAsync(L->ftok.data(), d->d_cftok + (size_t)slot * (d->max_frames + 3), (L->F + 1) * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
  if (nt) {
    DCK(cudaMemcpyAsync(L->state.data(), d->d_tok_state + (size_t)slot * d->tok_cap, nt * sizeof(int), cudaMemcpyDeviceToHost, d->stream));
    DCK(cudaMemcpyAsync(L->cost.data(), d->d_tok_cost + (size_t)slot * d->tok_cap, nt * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  }
  if (nl) DCK(cudaMemcpyAsync(L->links.data(), d->d_links + (size_t)slot * d->link_cap, nl * sizeof(DLink), cudaMemcpyDeviceToHost, d->stream));
  if (nf) DCK(cudaMemcpyAsync(L->coff.data(), d->d_coff + (size_t)slot * d->max_frames, nf * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  DCK(cudaStreamSynchronize(d->stream));
  if (index_now) index_lattice(L);
  return 0;
}

// FinalizeSearch (ctc_wfst_beam_search.cc:123-160), host part: n-best (or 1-best) of a fetched lattice.  Touches only the slot.
void finish_from_lattice(b2t_decoder* d, int slot, const Lattice& L, double fetch_ms, bool pruned) {
  static const bool timing = getenv("B2T_DECODER_TIMING") != nullptr;      // host-phase timing to stderr (profiling aid)
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  Slot& s = d->slots[slot];
  const double t1 = now();
  double t2 = t1;
  if (d->opt.nbest == 1) {
    std::vector<int> w; float g, a;
    if (best_path(d, L, true, &w, &g, &a)) push_result(d, s, w, g, a);
  } else {
    std::vector<float> extra; std::vector<char> keep;
    if (pruned) { extra.assign(L.state.size(), 0.0f); keep.assign(L.links.size(), 1); }   // lattice_prune_kernel left only survivors
    else prune_final(d, L, &extra, &keep);
    t2 = now();
    nbest_from_lattice(d, s, L, extra, keep);
  }
  if (timing)
    fprintf(stderr, "b2t decoder slot %d: %zu tokens, %zu links; fetch+sort %.2f ms, prune %.2f ms, n-best %.2f ms, %zu results\n", slot, L.state.size(),
            L.links.size(), fetch_ms, t2 - t1, now() - t2, s.results.size());
}

int finish_slot(b2t_decoder* d, int slot) {
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  Slot& s = d->slots[slot];
  if (s.finished) return 0;                       // the lattice was compacted by the first call; its results stand
  s.results.clear();
  s.finished = true;
  if (s.n_fed == 0) return 0;
  const double t0 = now();
  if (d->opt.nbest == 1 && use_gpu_prune()) {     // 1-best with final costs: back-pointer walk on the device
    std::vector<int> w; float g, a;
    const int r = best_path_on_gpu(d, slot, true, &w, &g, &a);
    if (r < 0) return r;
    if (r > 0) push_result(d, s, w, g, a);
    return 0;
  }
  Lattice L;
  const bool gpu_prune = d->opt.nbest != 1 && use_gpu_prune();
  int rc = gpu_prune ? prune_slots_on_gpu(d, std::vector<int>{slot}) : 0;
  if (rc) return rc;
  if (gpu_prune) s.pruned_on_gpu = true;
  rc = gpu_prune ? fetch_pruned_lattice(d, slot, &L) : fetch_lattice(d, slot, &L);
  if (rc) return rc;
  finish_from_lattice(d, slot, L, now() - t0, gpu_prune);
  return 0;
}

// Finish several slots: the lattices are fetched one after the other (one copy stream), the host-side pruning and n-best
// extraction of different utterances run on worker threads.
int finish_slots(b2t_decoder* d, int N) {
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  if (d->opt.nbest == 1 && use_gpu_prune()) {
    for (int n = 0; n < N; ++n) { const int rc = finish_slot(d, n); if (rc) return rc; }
    return 0;
  }
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int n_workers = (int)std::min<unsigned>(hw, 16u);
  std::vector<std::thread> workers;
  std::vector<Lattice> lats(N);
  std::vector<double> fetch_ms(N, 0.0);
  std::mutex mu;
  std::condition_variable cv;
  int fetched = 0, next = 0;
  bool failed = false;
  const bool gpu_prune = d->opt.nbest != 1 && use_gpu_prune();
  auto work = [&] {
    for (;;) {
      int n;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return next < fetched || failed || (fetched == N && next >= N); });
        if (next >= fetched) return;
        n = next++;
      }
      if (d->slots[n].n_fed > 0) {
        if (gpu_prune) index_lattice(&lats[n]);
        finish_from_lattice(d, n, lats[n], fetch_ms[n], gpu_prune);
      }
      Lattice().state.swap(lats[n].state);
      std::vector<DLink>().swap(lats[n].links);
    }
  };
  int rc = 0;
  if (gpu_prune) {
    std::vector<int> ids;
    for (int n = 0; n < N; ++n)
      if (d->slots[n].n_fed > 0) ids.push_back(n);
    if ((rc = prune_slots_on_gpu(d, ids))) return rc;
    for (int n : ids) d->slots[n].pruned_on_gpu = true;
  }
  for (int i = 0; i < std::min(n_workers, N); ++i) workers.emplace_back(work);
  for (int n = 0; n < N && !rc; ++n) {
    Slot& s = d->slots[n];
    s.results.clear();
    s.finished = true;
    const double t0 = now();
    if (s.n_fed > 0) rc = gpu_prune ? fetch_pruned_lattice(d, n, &lats[n], false) : fetch_lattice(d, n, &lats[n]);
    fetch_ms[n] = now() - t0;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (rc) failed = true; else fetched = n + 1;
    }
    cv.notify_all();
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (rc) failed = true;
  }
  cv.notify_all();
  for (auto& w : workers) w.join();
  return rc;
}

int ensure_C(b2t_decoder* d, int C) {
  if (d->C == 0) {
    d->C = C;
    if (cudaMalloc(&d->d_logp, (size_t)d->max_slots * d->max_frames * C * sizeof(float)) != cudaSuccess) return dfail(B2T_ERR_CUDA, "cudaMalloc(logp) failed");
  } else if (d->C != C) return dfail(B2T_ERR_ARG, "class count changed from %d to %d", d->C, C);
  int max_il = 0;
  for (const DArc& a : d->g.arcs) max_il = std::max(max_il, a.il);
  if (max_il > C) return dfail(B2T_ERR_ARG, "graph has input label %d but the posteriors have only %d classes", max_il, C);
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI
namespace {
// ---- lattice LM rescoring, host core (brain_speech_decoder.cc:47-101).  FinalizeSearch leaves the
// determinised word lattice: one path per distinct word sequence W within lattice_beam.  The two LatticeRescore passes
// (LM the graph was built from with scale -1, rescoring LM with scale +1) then reduce to graph'(W) = g_W - c_old(W) + c_new(W),
// acoustic' = a_W, where c(W) is the cheapest path of W through the LM acceptor: back-off arcs are epsilons after
// ReadAndPrepareLmFst (projection on the output labels), the final cost counts, a W the LM rejects disappears.
struct LmAcceptor {
  HostGraph g;
  int load(const char* path) {
    if (load_fst(path, &g)) return B2T_ERR_ARG;
    for (DArc& a : g.arcs) a.il = a.ol;              // PROJECT_OUTPUT
    return 0;
  }
  void closure(std::map<int, float>* st) const {
    std::vector<int> work;
    for (auto& kv : *st) work.push_back(kv.first);
    while (!work.empty()) {
      const int s = work.back(); work.pop_back();
      const float c = (*st)[s];
      for (long long a = g.off[s]; a < g.off[s + 1]; ++a) {
        const DArc& arc = g.arcs[a];
        if (arc.il != 0) continue;
        const float nc = c + arc.w;
        auto it = st->find(arc.next);
        if (it == st->end() || nc < it->second) { (*st)[arc.next] = nc; work.push_back(arc.next); }
      }
    }
  }
  float cost(const int* w, int n) const {
    std::map<int, float> cur;
    cur[g.start] = 0.0f;
    closure(&cur);
    for (int i = 0; i < n; ++i) {
      std::map<int, float> nxt;
      for (auto& kv : cur)
        for (long long a = g.off[kv.first]; a < g.off[kv.first + 1]; ++a) {
          const DArc& arc = g.arcs[a];
          if (arc.il != w[i]) continue;
          const float nc = kv.second + arc.w;
          auto it = nxt.find(arc.next);
          if (it == nxt.end() || nc < it->second) nxt[arc.next] = nc;
        }
      if (nxt.empty()) return INFINITY;
      closure(&nxt);
      cur.swap(nxt);
    }
    float best = INFINITY;
    for (auto& kv : cur)
      if (g.fin[kv.first] != INFINITY) best = std::min(best, kv.second + g.fin[kv.first]);
    return best;
  }
};
}  // namespace

extern "C" {

const char* b2t_decoder_last_error(void) { return g_derr; }

b2t_decoder* b2t_decoder_create(const char* fst_path, const char* words_path, const b2t_decode_options* opt, int max_frames, int max_slots) {
  if (!fst_path || !words_path || !opt || max_frames < 1 || max_slots < 1) { dfail(B2T_ERR_ARG, "bad arguments"); return nullptr; }
  if (fst_path[0] == '\0') { dfail(B2T_ERR_UNSUPPORTED, "empty fst_path selects the LM-free CtcPrefixBeamSearch; use b2t_prefix_beam_search"); return nullptr; }
  if (!(opt->beam > 0.0f && opt->max_active > 1 && opt->lattice_beam > 0.0f && opt->min_active <= opt->max_active)) {
    dfail(B2T_ERR_ARG, "invalid decode options (LatticeFasterDecoderConfig::Check)"); return nullptr;
  }
  b2t_decoder* d = new b2t_decoder();
  if (load_fst(fst_path, &d->g) || read_words(words_path, &d->words)) { delete d; return nullptr; }
  d->opt = *opt; d->max_frames = max_frames; d->max_slots = max_slots;
  // pools: a frame holds at most ~max_active expanded tokens times the out-degree; sized generously and checked
  const long long per_frame = std::min<long long>((long long)std::max(opt->max_active, 1000) * 6, 200000);
  d->tok_cap = (int)std::min<long long>(per_frame * (max_frames + 1), 24000000);
  d->link_cap = (int)std::min<long long>((long long)d->tok_cap * 3, 60000000);
  const size_t ns = d->g.fin.size(), na = d->g.arcs.size();
  bool ok = cudaStreamCreate(&d->stream) == cudaSuccess;
  ok = ok && cudaMalloc(&d->d_arcs, std::max<size_t>(na, 1) * sizeof(DArc)) == cudaSuccess && cudaMalloc(&d->d_off, (ns + 1) * sizeof(long long)) == cudaSuccess &&
       cudaMalloc(&d->d_has_eps, ns) == cudaSuccess && cudaMalloc(&d->d_best, (size_t)max_slots * ns * 8) == cudaSuccess &&
       cudaMalloc(&d->d_tokidx, (size_t)max_slots * ns * 4) == cudaSuccess && cudaMalloc(&d->d_tok_state, (size_t)max_slots * d->tok_cap * 4) == cudaSuccess &&
       cudaMalloc(&d->d_tok_cost, (size_t)max_slots * d->tok_cap * 4) == cudaSuccess && cudaMalloc(&d->d_tok_bp, (size_t)max_slots * d->tok_cap * 4) == cudaSuccess &&
       cudaMalloc(&d->d_dirty, (size_t)max_slots * d->tok_cap) == cudaSuccess && cudaMalloc(&d->d_links, (size_t)max_slots * d->link_cap * sizeof(DLink)) == cudaSuccess &&
       cudaMalloc(&d->d_ftok, (size_t)max_slots * (max_frames + 3) * 4) == cudaSuccess && cudaMalloc(&d->d_flink, (size_t)max_slots * (max_frames + 3) * 4) == cudaSuccess &&
       cudaMalloc(&d->d_coff, (size_t)max_slots * max_frames * 4) == cudaSuccess && cudaMalloc(&d->d_counters, (size_t)max_slots * 16) == cudaSuccess &&
       cudaMalloc(&d->d_nfed, (size_t)max_slots * 4) == cudaSuccess && cudaMalloc(&d->d_slot_ids, (size_t)max_slots * 4) == cudaSuccess &&
       cudaMalloc(&d->d_fin, std::max<size_t>(ns, 1) * 4) == cudaSuccess && cudaMalloc(&d->d_extra, (size_t)max_slots * d->tok_cap * 4) == cudaSuccess &&
       cudaMalloc(&d->d_newidx, (size_t)max_slots * d->tok_cap * 4) == cudaSuccess && cudaMalloc(&d->d_cftok, (size_t)max_slots * (max_frames + 3) * 4) == cudaSuccess &&
       cudaMalloc(&d->d_ccounts, (size_t)max_slots * 8) == cudaSuccess &&
       cudaMalloc(&d->d_bp_words, (size_t)max_slots * BP_MAX_PATH * 4) == cudaSuccess && cudaMalloc(&d->d_bp_cost, (size_t)max_slots * 16) == cudaSuccess &&
       cudaMalloc(&d->d_bp_hop, (size_t)max_slots * BP_MAX_PATH * 8) == cudaSuccess && cudaMalloc(&d->d_bp_ol, (size_t)max_slots * BP_MAX_PATH * 4) == cudaSuccess;
  if (!ok) { dfail(B2T_ERR_CUDA, "decoder allocation failed: %s (states %zu, slots %d, token pool %d)", cudaGetErrorString(cudaGetLastError()), ns, max_slots, d->tok_cap); b2t_decoder_destroy(d); return nullptr; }
  cudaMemcpy(d->d_arcs, d->g.arcs.data(), na * sizeof(DArc), cudaMemcpyHostToDevice);
  cudaMemcpy(d->d_off, d->g.off.data(), (ns + 1) * sizeof(long long), cudaMemcpyHostToDevice);
  cudaMemcpy(d->d_has_eps, d->g.has_eps.data(), ns, cudaMemcpyHostToDevice);
  cudaMemcpy(d->d_fin, d->g.fin.data(), ns * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemset(d->d_best, 0xff, (size_t)max_slots * ns * 8);
  cudaMemset(d->d_tokidx, 0xff, (size_t)max_slots * ns * 4);
  d->slots.resize(max_slots);
  for (int i = 0; i < max_slots; ++i) reset_slot(d, i);
  cudaStreamSynchronize(d->stream);
  if (const char* v = getenv("B2T_DECODER_STRICT"))
    if (atoi(v) != 0 && b2t_decoder_set_strict_order(d, 1)) { b2t_decoder_destroy(d); return nullptr; }
  return d;
}

void b2t_decoder_destroy(b2t_decoder* d) {
  if (d) { delete (LmAcceptor*)d->lm_old; delete (LmAcceptor*)d->lm_new; d->lm_old = d->lm_new = nullptr; }
  if (!d) return;
  void* ptrs[] = {d->d_arcs, d->d_off, d->d_has_eps, d->d_best, d->d_tokidx, d->d_tok_state, d->d_tok_cost, d->d_tok_bp, d->d_dirty, d->d_links,
                  d->d_ftok, d->d_flink, d->d_coff, d->d_counters, d->d_nfed, d->d_slot_ids, d->d_logp, d->d_fin, d->d_extra, d->d_newidx, d->d_cftok, d->d_ccounts, d->d_bp_words, d->d_bp_cost, d->d_bp_hop, d->d_bp_ol};
  for (void* p : ptrs) if (p) cudaFree(p);
  void* sptrs[] = {d->d_eps_off, d->d_eps_arcs, d->d_first_rank, d->d_eps_bp, d->d_bucket_min, d->d_s_cum, d->d_s_premin, d->d_s_k0, d->d_s_k1, d->d_s_v0, d->d_s_v1,
                   d->d_s_byrank, d->d_s_queue, d->d_s_hash};
  for (void* p : sptrs) if (p) cudaFree(p);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

int b2t_decoder_set_options(b2t_decoder* d, const b2t_decode_options* opt) {
  if (!d || !opt) return dfail(B2T_ERR_ARG, "null argument");
  // NOTE: the reference's SetOpt only reaches acoustic_scale / nbest / blank_skip (the Kaldi config copy is never
  // updated).  Here every field takes effect from the next Reset().
  d->opt = *opt;
  return 0;
}

int b2t_decoder_set_strict_order(b2t_decoder* d, int on) {
  if (!d) return dfail(B2T_ERR_ARG, "null decoder");
  for (const Slot& s : d->slots)
    if (s.n_fed > 0 && !s.finished) return dfail(B2T_ERR_STATE, "the search order can only be switched between utterances (Reset() first)");
  if (on && !d->d_first_rank) {
    const size_t ns = d->g.fin.size(), S = (size_t)d->max_slots;
    std::vector<long long> eoff(ns + 1, 0);
    std::vector<EArc> earcs;
    for (size_t st = 0; st < ns; ++st) {
      eoff[st] = (long long)earcs.size();
      if (d->g.has_eps[st])
        for (long long a = d->g.off[st]; a < d->g.off[st + 1]; ++a)
          if (d->g.arcs[a].il == 0) earcs.push_back(EArc{d->g.arcs[a].next, d->g.arcs[a].w, d->g.arcs[a].ol, (int)a});
    }
    eoff[ns] = (long long)earcs.size();
    const long long per_frame = std::min<long long>((long long)std::max(d->opt.max_active, 1000) * 6, 200000);
    d->strict_fc = (int)std::min<long long>(d->tok_cap, per_frame * 4);
    d->strict_qcap = d->strict_fc * 4;
    d->strict_hcap = std::max(2 * d->strict_fc + 16, 1024);
    const size_t fc = (size_t)d->strict_fc;
    bool ok = cudaMalloc(&d->d_eps_off, (ns + 1) * 8) == cudaSuccess && cudaMalloc(&d->d_eps_arcs, std::max<size_t>(earcs.size(), 1) * sizeof(EArc)) == cudaSuccess &&
              cudaMalloc(&d->d_first_rank, S * ns * 4) == cudaSuccess && cudaMalloc(&d->d_eps_bp, S * ns * 4) == cudaSuccess &&
              cudaMalloc(&d->d_bucket_min, S * d->strict_hcap * 4) == cudaSuccess && cudaMalloc(&d->d_s_cum, S * (fc + 1) * 4) == cudaSuccess &&
              cudaMalloc(&d->d_s_premin, S * fc * 4) == cudaSuccess && cudaMalloc(&d->d_s_k0, S * fc * 4) == cudaSuccess && cudaMalloc(&d->d_s_k1, S * fc * 4) == cudaSuccess &&
              cudaMalloc(&d->d_s_v0, S * fc * 4) == cudaSuccess && cudaMalloc(&d->d_s_v1, S * fc * 4) == cudaSuccess && cudaMalloc(&d->d_s_byrank, S * fc * 4) == cudaSuccess &&
              cudaMalloc(&d->d_s_queue, S * (size_t)d->strict_qcap * 4) == cudaSuccess && cudaMalloc(&d->d_s_hash, S * 4) == cudaSuccess;
    if (!ok) return dfail(B2T_ERR_CUDA, "strict-order scratch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    DCK(cudaMemcpy(d->d_eps_off, eoff.data(), (ns + 1) * 8, cudaMemcpyHostToDevice));
    if (!earcs.empty()) DCK(cudaMemcpy(d->d_eps_arcs, earcs.data(), earcs.size() * sizeof(EArc), cudaMemcpyHostToDevice));
    DCK(cudaMemset(d->d_first_rank, 0x7f, S * ns * 4));          // RANK_NONE
    DCK(cudaMemset(d->d_bucket_min, 0x7f, S * d->strict_hcap * 4));
    DCK(cudaMemset(d->d_s_hash, 0, S * 4));
  }
  d->strict = on != 0;
  return 0;
}

int b2t_decoder_reset(b2t_decoder* d, int slot) {
  if (!d || slot < 0 || slot >= d->max_slots) return dfail(B2T_ERR_ARG, "bad slot");
  // the state table is clean after every frame, so only the counters need re-arming
  return reset_slot(d, slot);
}

int b2t_decoder_decode_logprobs(b2t_decoder* d, int slot, const float* logp, int T, int C) {
  if (!d || slot < 0 || slot >= d->max_slots || (!logp && T > 0) || T < 0 || C < 1) return dfail(B2T_ERR_ARG, "bad arguments");
  int rc = ensure_C(d, C);
  if (rc) return rc;
  Slot& s = d->slots[slot];
  if (s.finished) return dfail(B2T_ERR_STATE, "Decode() after FinishDecoding(): call Reset() first");
  if (T == 0) return 0;
  const int before = s.n_fed;
  if ((rc = feed_frames(d, s, logp, T, C))) return rc;
  if ((rc = upload_fed(d, slot, before))) return rc;
  if ((rc = launch_slots(d, std::vector<int>{slot}))) return rc;
  return partial_result(d, slot);
}

int b2t_decoder_decode_logits(b2t_decoder* d, int slot, const float* logits, const float* log_priors, int T, int C, float blank_penalty) {
  if (!logits && T > 0) return dfail(B2T_ERR_ARG, "null logits");
  std::vector<float> lp((size_t)T * C);
  for (int t = 0; t < T; ++t) {                     // lm_decoder.cc:30-36
    log_softmax_row(logits + (size_t)t * C, C, &lp[(size_t)t * C]);
    if (log_priors)
      for (int c = 0; c < C; ++c) lp[(size_t)t * C + c] -= log_priors[(size_t)t * C + c];
    lp[(size_t)t * C] -= blank_penalty;
  }
  return b2t_decoder_decode_logprobs(d, slot, lp.data(), T, C);
}

int b2t_decoder_finish(b2t_decoder* d, int slot) {
  if (!d || slot < 0 || slot >= d->max_slots) return dfail(B2T_ERR_ARG, "bad slot");
  return finish_slot(d, slot);
}

// DecodeResource.lm_fst_path / rescore_lm_fst_path (brain_speech_decoder.h:57-79): the LM the decoding graph was built from and the
// (unpruned) rescoring LM, loaded once.  Either path empty: rescoring stays unavailable, like a null LM pointer in the reference.
int b2t_decoder_set_rescore_lms(b2t_decoder* d, const char* lm_fst_path, const char* rescore_lm_fst_path) {
  if (!d) return dfail(B2T_ERR_ARG, "null decoder");
  delete (LmAcceptor*)d->lm_old; delete (LmAcceptor*)d->lm_new;
  d->lm_old = d->lm_new = nullptr;
  if (!lm_fst_path || !rescore_lm_fst_path || !lm_fst_path[0] || !rescore_lm_fst_path[0]) return 0;
  LmAcceptor* a = new LmAcceptor();
  LmAcceptor* b = new LmAcceptor();
  if (a->load(lm_fst_path) || b->load(rescore_lm_fst_path)) { delete a; delete b; return B2T_ERR_ARG; }
  d->lm_old = a; d->lm_new = b;
  return 0;
}

// Rescore() (brain_speech_decoder.cc:61-101) after FinishDecoding: every distinct word sequence W of the pruned lattice (what
// CtcWfstBeamSearch::Lattice() holds after determinisation) gets graph' = g_W - c_old(W) + c_new(W); the result list keeps its
// first-pass length and is re-filled best first.  The lattice is read back from the device, where it stays until Reset.
int b2t_decoder_rescore(b2t_decoder* d, int slot) {
  if (!d || slot < 0 || slot >= d->max_slots) return dfail(B2T_ERR_ARG, "bad slot");
  if (!d->lm_old || !d->lm_new) return dfail(B2T_ERR_STATE, "Rescore() needs DecodeResource.lm_fst_path and rescore_lm_fst_path (b2t_decoder_set_rescore_lms)");
  Slot& s = d->slots[slot];
  if (!s.finished) return dfail(B2T_ERR_STATE, "Rescore() needs a preceding FinishDecoding()");
  const size_t keep_n = s.results.size();
  if (keep_n == 0 || s.n_fed == 0) return 0;
  Lattice L;
  std::vector<float> extra; std::vector<char> keep;
  int rc;
  if (use_gpu_prune()) {
    if (!s.pruned_on_gpu) {                       // nbest == 1 finished through the back-pointer walk: prune now
      if ((rc = prune_slots_on_gpu(d, std::vector<int>{slot}))) return rc;
      s.pruned_on_gpu = true;
    }
    if ((rc = fetch_pruned_lattice(d, slot, &L))) return rc;
    extra.assign(L.state.size(), 0.0f); keep.assign(L.links.size(), 1);
  } else {
    if ((rc = fetch_lattice(d, slot, &L))) return rc;
    prune_final(d, L, &extra, &keep);
  }
  static const int cap = getenv("B2T_RESCORE_MAX_SEQS") ? atoi(getenv("B2T_RESCORE_MAX_SEQS")) : 20000;   // distinct sequences considered (cheapest first)
  std::vector<NbOut> all;
  collect_nbest(d, L, extra, keep, cap, &all);
  const LmAcceptor* lo = (const LmAcceptor*)d->lm_old;
  const LmAcceptor* ln = (const LmAcceptor*)d->lm_new;
  struct R { size_t idx; float g, a; };
  std::vector<R> scored;
  for (size_t i = 0; i < all.size(); ++i) {
    const float c_old = lo->cost(all[i].words.data(), (int)all[i].words.size()), c_new = ln->cost(all[i].words.data(), (int)all[i].words.size());
    if (c_old == INFINITY || c_new == INFINITY) continue;      // a sequence an LM rejects leaves the lattice in the composition
    scored.push_back({i, -(-all[i].g + c_old) + c_new, all[i].a});
  }
  std::stable_sort(scored.begin(), scored.end(), [](const R& x, const R& y) { return hyp_better(Hyp{0, x.g, x.a}, Hyp{0, y.g, y.a}); });
  s.results.clear();
  for (size_t i = 0; i < scored.size() && i < keep_n; ++i) push_result(d, s, all[scored[i].idx].words, scored[i].g, scored[i].a);
  return 0;
}

// Host core of Rescore(), usable on its own (no GPU involved): n word sequences (ids concatenated in `words`, lengths in `lens`)
// with their first-pass (graph, acoustic) costs are re-scored as graph' = graph - c_old + c_new and ordered by graph' + acoustic
// (then by smaller graph', LatticeWeight Compare); at most `keep` survive.  order_out[i] = index of the i-th best input sequence,
// graph_out[i] its new graph cost.  Returns the number of entries written, or a negative error code.
int b2t_lm_rescore_sequences(const char* lm_fst_path, const char* rescore_lm_fst_path, int n, const int* words, const int* lens,
                             const float* graph, const float* acoustic, int keep, int* order_out, float* graph_out) {
  if (!lm_fst_path || !rescore_lm_fst_path || n < 0 || (n > 0 && (!words || !lens || !graph || !acoustic)) || !order_out || !graph_out)
    return dfail(B2T_ERR_ARG, "bad arguments");
  LmAcceptor lm_old, lm_new;
  if (lm_old.load(lm_fst_path) || lm_new.load(rescore_lm_fst_path)) return B2T_ERR_ARG;
  struct R { int idx; float g, a; };
  std::vector<R> all;
  size_t pos = 0;
  for (int i = 0; i < n; ++i) {
    const float c_old = lm_old.cost(words + pos, lens[i]), c_new = lm_new.cost(words + pos, lens[i]);
    pos += (size_t)lens[i];
    if (c_old == INFINITY || c_new == INFINITY) continue;
    all.push_back({i, -(-graph[i] + c_old) + c_new, acoustic[i]});
  }
  std::stable_sort(all.begin(), all.end(), [](const R& x, const R& y) {
    const float fx = x.g + x.a, fy = y.g + y.a;
    if (fx != fy) return fx < fy;
    return x.g < y.g;
  });
  int m = 0;
  for (; m < (int)all.size() && m < keep; ++m) { order_out[m] = all[m].idx; graph_out[m] = all[m].g; }
  return m;
}

int b2t_decoder_num_results(b2t_decoder* d, int slot) {
  if (!d || slot < 0 || slot >= d->max_slots) return dfail(B2T_ERR_ARG, "bad slot");
  return (int)d->slots[slot].results.size();
}

int b2t_decoder_get_result(b2t_decoder* d, int slot, int i, float* ac_score, float* lm_score, char* sentence, int cap) {
  if (!d || slot < 0 || slot >= d->max_slots) return dfail(B2T_ERR_ARG, "bad slot");
  const Slot& s = d->slots[slot];
  if (i < 0 || i >= (int)s.results.size()) return dfail(B2T_ERR_ARG, "result index out of range");
  if (ac_score) *ac_score = s.results[i].ac;
  if (lm_score) *lm_score = s.results[i].lm;
  if (sentence && cap > 0) snprintf(sentence, cap, "%s", s.results[i].sentence.c_str());
  return (int)s.results[i].sentence.size();
}

// Batched extension: reset + decode + finish N utterances concurrently (one CTA each).  logits: host [N][T][C];
// lens[n] <= T frames are used.  Results are read per slot n afterwards.
int b2t_decoder_decode_batch(b2t_decoder* d, const float* logits, const int* lens, int N, int T, int C, float blank_penalty, int finish) {
  if (!d || !logits || !lens || N < 1 || N > d->max_slots) return dfail(B2T_ERR_ARG, "bad batch arguments (N must be <= max_slots)");
  int rc = ensure_C(d, C);
  if (rc) return rc;
  std::vector<int> ids;
  std::vector<float> lp;
  for (int n = 0; n < N; ++n) {
    if ((rc = reset_slot(d, n))) return rc;
    const int Tn = std::min(std::max(lens[n], 0), T);
    lp.resize((size_t)Tn * C);
    for (int t = 0; t < Tn; ++t) {
      log_softmax_row(logits + ((size_t)n * T + t) * C, C, &lp[(size_t)t * C]);
      lp[(size_t)t * C] -= blank_penalty;
    }
    if ((rc = feed_frames(d, d->slots[n], lp.data(), Tn, C))) return rc;
    if ((rc = upload_fed(d, n, 0))) return rc;
    ids.push_back(n);
  }
  if ((rc = launch_slots(d, ids))) return rc;
  if (finish && (rc = finish_slots(d, N))) return rc;
  return 0;
}

int b2t_decoder_stats(b2t_decoder* d, int slot, int* frames, long long* tokens, long long* links, double* kernel_ms) {
  if (!d || slot < 0 || slot >= d->max_slots) return dfail(B2T_ERR_ARG, "bad slot");
  int c[4];
  if (cudaMemcpy(c, d->d_counters + slot * 4, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return dfail(B2T_ERR_CUDA, "stats copy failed");
  if (frames) *frames = c[2];
  if (tokens) *tokens = c[0];
  if (links) *links = c[1];
  if (kernel_ms) *kernel_ms = d->last_kernel_ms;
  return 0;
}

int b2t_decoder_tokens_per_frame(b2t_decoder* d, int slot, int* out, int cap) {
  if (!d || slot < 0 || slot >= d->max_slots || !out) return dfail(B2T_ERR_ARG, "bad arguments");
  int c[4];
  if (cudaMemcpy(c, d->d_counters + slot * 4, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return dfail(B2T_ERR_CUDA, "copy failed");
  const int nf = std::max(c[2], 0);
  std::vector<int> ft(nf + 2);
  if (cudaMemcpy(ft.data(), d->d_ftok + (size_t)slot * (d->max_frames + 3), (nf + 2) * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return dfail(B2T_ERR_CUDA, "copy failed");
  for (int f = 0; f < nf && f < cap; ++f) out[f] = ft[f + 2] - ft[f + 1];
  return nf;
}

// test hook: states and costs of the tokens of frame_plus_one `fp1`, in pool (= list, in strict mode) order; before finish only
int b2t_decoder_debug_frame_tokens(b2t_decoder* d, int slot, int fp1, int* states, float* costs, int cap) {
  if (!d || slot < 0 || slot >= d->max_slots || !states || !costs) return dfail(B2T_ERR_ARG, "bad arguments");
  if (d->slots[slot].pruned_on_gpu) return dfail(B2T_ERR_STATE, "the pool has been pruned");
  int c[4];
  if (cudaMemcpy(c, d->d_counters + slot * 4, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return dfail(B2T_ERR_CUDA, "copy failed");
  if (fp1 < 0 || fp1 > std::max(c[2], 0)) return dfail(B2T_ERR_ARG, "frame out of range");
  int ft[2];
  if (cudaMemcpy(ft, d->d_ftok + (size_t)slot * (d->max_frames + 3) + fp1, sizeof(ft), cudaMemcpyDeviceToHost) != cudaSuccess) return dfail(B2T_ERR_CUDA, "copy failed");
  const int n = std::min(ft[1] - ft[0], cap);
  if (n > 0 && (cudaMemcpy(states, d->d_tok_state + (size_t)slot * d->tok_cap + ft[0], n * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(costs, d->d_tok_cost + (size_t)slot * d->tok_cap + ft[0], n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess))
    return dfail(B2T_ERR_CUDA, "copy failed");
  return ft[1] - ft[0];
}

}  // extern "C"
