// Strict serial-order mode of the WFST search: reproduces what Kaldi's LatticeFasterDecoder does when it walks its token
// list ONE TOKEN AT A TIME, in parallel.  (Included by decode.cu inside its anonymous namespace.)
//
// What is order dependent in the reference (language_model/runtime/core/kaldi/decoder/lattice-faster-decoder.cc):
//   * ProcessEmitting :785-822 tightens `next_cutoff` while it iterates: the arc (token i, arc a) is accepted iff its cost is
//     below  min(prepass, adaptive_beam + min over all EARLIER (i', a') of their cost)  -- an exclusive prefix-min in the
//     sequence order (token list order, then arc order).  Tokens accepted above the final cutoff stay in the list: they count
//     in the next frame's GetCutoff :651-720 and are expanded when max_active lets them.
//   * the token list order itself: HashList (util/hash-list-inl.h:124-171) chains buckets in the order of their first
//     occupation and appends inside a bucket, with bucket = state % hash_size and hash_size = max(previous, 2 x tokens)
//     (PossiblyResizeHash :216-222).  The list order of frame f+1 therefore is: sort by (insertion rank of the bucket's first
//     token, own insertion rank), where the insertion rank of a token is the sequence position of the first accepted arc
//     that reached its state -- all computable with atomicMin + scans + a stable radix sort.
//   * ProcessNonemitting :839-909 pops a LIFO; the order in which it creates new tokens is inherently serial (and cheap:
//     only states with input-epsilon arcs take part), so one thread replays it literally over a compact epsilon-arc CSR.
// Costs, links and the pruned lattice follow from the accepted set, so per-frame token counts, the 1-best and the n-best set
// equal the oracle's also when max_active binds (tests/test_gpu_decoder.py::test_strict_*).

constexpr int RANK_NONE = 0x7f7f7f7f;   // what cudaMemset(0x7f) leaves behind
constexpr unsigned int EPS_KEY_BASE = 0x80000000u;
constexpr int DEC_WARPS = DEC_THREADS / 32;

// in-place exclusive scans in index order: a[0..n) prefix sum (a[n] = total), b[0..n) prefix min (b may be null); the
// minimum of all of b is returned to every thread
__device__ float block_scan_sum_min(int* a, float* b, int n, int* s_wi, float* s_wf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int carry_i = 0;
  float carry_f = INFINITY;
  for (int base = 0; base < n; base += DEC_THREADS) {
    const int i = base + threadIdx.x;
    const int xi = i < n ? a[i] : 0;
    const float xf = (b && i < n) ? b[i] : INFINITY;
    int si = xi;
    float sf = xf;
    for (int o = 1; o < 32; o <<= 1) {
      const int ti = __shfl_up_sync(0xffffffffu, si, o);
      const float tf = __shfl_up_sync(0xffffffffu, sf, o);
      if (lane >= o) { si += ti; sf = fminf(sf, tf); }
    }
    float pf = __shfl_up_sync(0xffffffffu, sf, 1);
    if (lane == 0) pf = INFINITY;
    if (lane == 31) { s_wi[warp] = si; s_wf[warp] = sf; }
    __syncthreads();
    int wi = 0, tot_i = 0;
    float wf = INFINITY, tot_f = INFINITY;
    for (int w = 0; w < DEC_WARPS; ++w) {
      if (w < warp) { wi += s_wi[w]; wf = fminf(wf, s_wf[w]); }
      tot_i += s_wi[w];
      tot_f = fminf(tot_f, s_wf[w]);
    }
    if (i < n) {
      a[i] = carry_i + wi + (si - xi);
      if (b) b[i] = fminf(carry_f, fminf(wf, pf));
    }
    carry_i += tot_i;
    carry_f = fminf(carry_f, tot_f);
    __syncthreads();
  }
  if (threadIdx.x == 0) a[n] = carry_i;
  __syncthreads();
  return carry_f;
}

// one stable LSD radix pass (8-bit digit at `shift`) over n (key, value) pairs: warp w owns a contiguous slice and walks it in
// order, __match_any_sync ranks equal digits inside each group of 32, so equal keys keep their input order
__device__ void radix_pass(const unsigned int* kin, const int* vin, unsigned int* kout, int* vout, int n, int shift,
                           unsigned int* s_hist, unsigned int* s_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = ((n + DEC_WARPS - 1) / DEC_WARPS + 31) & ~31;
  const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
  for (int i = threadIdx.x; i < DEC_WARPS * 256; i += DEC_THREADS) s_hist[i] = 0;
  __syncthreads();
  for (int i = beg + lane; i < end; i += 32) atomicAdd(&s_hist[warp * 256 + ((kin[i] >> shift) & 255u)], 1u);
  __syncthreads();
  if (threadIdx.x < 256) {
    unsigned int acc = 0;
    for (int w = 0; w < DEC_WARPS; ++w) {
      const unsigned int c = s_hist[w * 256 + threadIdx.x];
      s_hist[w * 256 + threadIdx.x] = acc;
      acc += c;
    }
    s_tot[threadIdx.x] = acc;
  }
  __syncthreads();
  if (warp == 0) {
    unsigned int loc[8], sum = 0;
    for (int j = 0; j < 8; ++j) { loc[j] = s_tot[lane * 8 + j]; sum += loc[j]; }
    unsigned int inc = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    unsigned int ex = inc - sum;
    for (int j = 0; j < 8; ++j) { s_tot[lane * 8 + j] = ex; ex += loc[j]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < DEC_WARPS * 256; i += DEC_THREADS) s_hist[i] += s_tot[i & 255];
  __syncthreads();
  for (int i0 = beg; i0 < end; i0 += 32) {
    const int i = i0 + lane;
    const bool act = i < end;
    const unsigned int amask = __ballot_sync(0xffffffffu, act);
    if (act) {
      const unsigned int k = kin[i];
      const int v = vin[i];
      const unsigned int d = (k >> shift) & 255u;
      const unsigned int peers = __match_any_sync(amask, d);
      const int r = __popc(peers & ((1u << lane) - 1u));
      const unsigned int base = s_hist[warp * 256 + d];
      __syncwarp(amask);
      kout[base + r] = k;
      vout[base + r] = v;
      if (r == 0) s_hist[warp * 256 + d] = base + (unsigned int)__popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
}

struct StrictCtx {
  unsigned long long* best64;
  int *tokidx, *tok_state, *tok_bp, *counters;
  float* tok_cost;
  DLink* links;
  int *first_rank, *eps_bp, *bucket_min, *cum, *byrank, *queue;
  float* premin;
  unsigned int *k0, *k1;
  int *v0, *v1;
  unsigned int *s_hist, *s_tot;
  int* s_wi;
  float* s_wf;
  int* s_i;       // [4] shared scalars
};

// stable sort of (k0, v0)[0..n) by the low `bits` bits of the key; returns 0/1 = which buffer pair holds the result
__device__ int block_sort(const StrictCtx& c, int n, unsigned int max_key) {
  int cur = 0;
  for (int shift = 0; shift < 32 && (max_key >> shift) != 0; shift += 8) {
    if (cur == 0) radix_pass(c.k0, c.v0, c.k1, c.v1, n, shift, c.s_hist, c.s_tot);
    else radix_pass(c.k1, c.v1, c.k0, c.v0, n, shift, c.s_hist, c.s_tot);
    cur ^= 1;
  }
  return cur;
}

// tokens byrank[0..m) (insertion order, dense ranks in first_rank) -> HashList order, written to tok_state[tb ..) / tokidx
__device__ void order_by_bucket(const DecParams& p, const StrictCtx& c, int tb, int m, int hs) {
  for (int q = threadIdx.x; q < m; q += DEC_THREADS) {
    const int st = c.byrank[q];
    c.k0[q] = (unsigned int)c.bucket_min[st % hs];
    c.v0[q] = st;
  }
  __syncthreads();
  const int cur = block_sort(c, m, (unsigned int)max(m, 1));
  const int* v = cur ? c.v1 : c.v0;
  for (int q = threadIdx.x; q < m; q += DEC_THREADS) {
    const int st = v[q];
    c.tok_state[tb + q] = st;
    c.tokidx[st] = tb + q;
  }
  __syncthreads();
}

// The new frame's tokens [tb, counters[0]) were created in arbitrary order with raw insertion ranks in first_rank[state].
// Puts them in list order, replays ProcessNonemitting(cutoff) serially, re-orders, resolves link destinations, emits the
// epsilon links and finishes the frame (costs, back-pointers, table clean-up).  cum_prev/tb_prev describe the source frame.
__device__ void strict_close_frame(const DecParams& p, const StrictCtx& c, int tb, int lb, float cutoff, int hs, unsigned int max_rank,
                                   int tb_prev) {
  int* counters = c.counters;
  const int m1 = counters[0] - tb;
  if (m1 >= p.fc) { if (threadIdx.x == 0) counters[3] = 3; __syncthreads(); return; }
  // ---- insertion order: sort by raw rank, then make ranks dense
  for (int j = threadIdx.x; j < m1; j += DEC_THREADS) {
    const int st = c.tok_state[tb + j];
    const int r = c.first_rank[st];
    c.k0[j] = (unsigned int)r;
    c.v0[j] = st;
    atomicMin(&c.bucket_min[st % hs], r);
  }
  __syncthreads();
  const int cur = block_sort(c, m1, max_rank);
  const unsigned int* kr = cur ? c.k1 : c.k0;
  const int* vr = cur ? c.v1 : c.v0;
  unsigned int* head = cur ? c.k0 : c.k1;
  for (int q = threadIdx.x; q < m1; q += DEC_THREADS) {
    const int st = vr[q];
    c.byrank[q] = st;
    head[q] = (unsigned int)c.bucket_min[st % hs] == kr[q] ? 1u : 0u;
  }
  __syncthreads();
  for (int q = threadIdx.x; q < m1; q += DEC_THREADS) {
    const int st = c.byrank[q];
    c.first_rank[st] = q;
    if (head[q]) c.bucket_min[st % hs] = q;
  }
  __syncthreads();
  order_by_bucket(p, c, tb, m1, hs);
  // ---- ProcessNonemitting: queue = tokens with input-epsilon arcs in list order, popped from the back
  int* tmp = (int*)c.k1;
  for (int q = threadIdx.x; q < m1; q += DEC_THREADS) tmp[q] = p.has_eps[c.tok_state[tb + q]];
  __syncthreads();
  block_scan_sum_min(tmp, nullptr, m1, c.s_wi, c.s_wf);
  for (int q = threadIdx.x; q < m1; q += DEC_THREADS) {
    const int st = c.tok_state[tb + q];
    if (p.has_eps[st]) c.queue[tmp[q]] = st;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int qn = tmp[m1], k = 0, status = 0;
    unsigned int evalc = 0;
    while (qn > 0 && status == 0) {
      const int st = c.queue[--qn];
      const float cur_cost = funkey((unsigned int)(c.best64[st] >> 32));
      if (cur_cost >= cutoff) continue;
      for (long long a = p.eps_off[st]; a < p.eps_off[st + 1]; ++a) {
        const EArc arc = p.eps_arcs[a];
        const float tot = __fadd_rn(cur_cost, arc.w);
        if (!(tot < cutoff)) continue;
        const unsigned long long b = c.best64[arc.next];
        bool changed = false;
        if (b == EMPTY64) {
          const int ns = counters[0];
          if (ns >= p.tok_cap) { status = 1; break; }
          if (m1 + k >= p.fc - 1) { status = 3; break; }
          counters[0] = ns + 1;
          c.tok_state[ns] = arc.next;
          c.tokidx[arc.next] = ns;
          c.byrank[m1 + k] = arc.next;
          c.first_rank[arc.next] = m1 + k;
          int* bm = &c.bucket_min[arc.next % hs];
          if (*bm > m1 + k) *bm = m1 + k;
          ++k;
          changed = true;
        } else if (funkey((unsigned int)(b >> 32)) > tot) {
          changed = true;
        }
        if (changed) {
          c.best64[arc.next] = ((unsigned long long)fkey(tot) << 32) | (EPS_KEY_BASE + (evalc & 0x7fffffffu));
          c.eps_bp[arc.next] = st;
          if (p.has_eps[arc.next]) {
            if (qn >= p.qcap) { status = 3; break; }
            c.queue[qn++] = arc.next;
          }
        }
        ++evalc;
      }
    }
    if (status) counters[3] = status;
    c.s_i[0] = k;
  }
  __syncthreads();
  if (counters[3] != 0) return;
  const int k = c.s_i[0];
  if (k > 0) order_by_bucket(p, c, tb, m1 + k, hs);
  const int te = counters[0];
  // ---- emitting links: destination state -> token index
  for (int l = lb + threadIdx.x; l < counters[1]; l += DEC_THREADS) c.links[l].dst = c.tokidx[c.links[l].dst];
  __syncthreads();
  const int le_emit = counters[1];
  __syncthreads();
  // ---- epsilon links from the converged costs (the last time Kaldi processes a token it sees its final cost)
  for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
    const int st = c.tok_state[i];
    if (!p.has_eps[st]) continue;
    const float cur_cost = funkey((unsigned int)(c.best64[st] >> 32));
    if (cur_cost >= cutoff) continue;
    for (long long a = p.eps_off[st]; a < p.eps_off[st + 1]; ++a) {
      const EArc arc = p.eps_arcs[a];
      const float tot = __fadd_rn(cur_cost, arc.w);
      if (tot < cutoff) {
        const int li = atomicAdd(&counters[1], 1);
        if (li < p.link_cap) c.links[li] = DLink{i, c.tokidx[arc.next], 0, arc.ol, arc.w, 0.0f};
        else counters[3] = 2;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && counters[1] > p.link_cap) counters[1] = p.link_cap;
  __syncthreads();
  // ---- final costs and back-pointers: the earliest arc (in Kaldi's sequence) that reached the final cost
  for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
    const int st = c.tok_state[i];
    const unsigned long long b = c.best64[st];
    c.tok_cost[i] = funkey((unsigned int)(b >> 32));
    if ((unsigned int)b >= EPS_KEY_BASE) c.tok_bp[i] = c.tokidx[c.eps_bp[st]];
  }
  __syncthreads();
  for (int l = lb + threadIdx.x; l < le_emit; l += DEC_THREADS) {
    const DLink kk = c.links[l];
    const unsigned long long b = c.best64[c.tok_state[kk.dst]];
    const unsigned int r = (unsigned int)b;
    if (r < EPS_KEY_BASE && (int)r >= c.cum[kk.src - tb_prev] && (int)r < c.cum[kk.src - tb_prev + 1]) c.tok_bp[kk.dst] = kk.src;
  }
  __syncthreads();
  for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
    const int st = c.tok_state[i];
    c.best64[st] = EMPTY64;
    c.tokidx[st] = -1;
    c.first_rank[st] = RANK_NONE;
    c.bucket_min[st % hs] = RANK_NONE;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(DEC_THREADS, 1)
wfst_decode_strict_kernel(const DecParams p) {
  __shared__ float sred[DEC_WARPS];
  __shared__ unsigned int hist[256];
  __shared__ unsigned int sh[2];
  __shared__ float s_prep;
  __shared__ int s_cand[DEC_WARPS];
  __shared__ int s_wi[DEC_WARPS];
  __shared__ float s_wf[DEC_WARPS];
  __shared__ unsigned int s_hist[DEC_WARPS * 256];
  __shared__ unsigned int s_tot[256];
  __shared__ int s_i[4];
  const int slot = p.slot_ids[blockIdx.x];
  StrictCtx c;
  c.best64 = p.best64 + (size_t)slot * p.nstates;
  c.tokidx = p.tokidx + (size_t)slot * p.nstates;
  c.tok_state = p.tok_state + (size_t)slot * p.tok_cap;
  c.tok_cost = p.tok_cost + (size_t)slot * p.tok_cap;
  c.tok_bp = p.tok_bp + (size_t)slot * p.tok_cap;
  c.links = p.links + (size_t)slot * p.link_cap;
  c.counters = p.counters + slot * 4;
  c.first_rank = p.first_rank + (size_t)slot * p.nstates;
  c.eps_bp = p.eps_bp + (size_t)slot * p.nstates;
  c.bucket_min = p.bucket_min + (size_t)slot * p.hcap;
  c.cum = p.s_cum + (size_t)slot * (p.fc + 1);
  c.premin = p.s_premin + (size_t)slot * p.fc;
  c.k0 = p.s_k0 + (size_t)slot * p.fc; c.k1 = p.s_k1 + (size_t)slot * p.fc;
  c.v0 = p.s_v0 + (size_t)slot * p.fc; c.v1 = p.s_v1 + (size_t)slot * p.fc;
  c.byrank = p.s_byrank + (size_t)slot * p.fc;
  c.queue = p.s_queue + (size_t)slot * p.qcap;
  c.s_hist = s_hist; c.s_tot = s_tot; c.s_wi = s_wi; c.s_wf = s_wf; c.s_i = s_i;
  int* ftok = p.frame_tok_off + (size_t)slot * (p.max_frames + 3);
  int* flink = p.frame_link_off + (size_t)slot * (p.max_frames + 3);
  float* coff = p.cost_offsets + (size_t)slot * p.max_frames;
  int* counters = c.counters;
  int* hashp = p.s_hash + slot;
  const float* logp = p.logp + (size_t)slot * p.max_frames * p.C;
  const int n_fed = p.n_fed[slot];

  if (counters[2] < 0) {
    // ---- InitDecoding (:58-75): start token, ProcessNonemitting(beam).  The HashList starts at 1000 buckets (constructor)
    // and is never shrunk, also not by InitDecoding: its size carries over from the utterances this slot decoded before.
    if (threadIdx.x == 0) {
      counters[0] = 1; counters[1] = 0;
      c.tok_state[0] = p.start; c.tok_bp[0] = -1;
      c.best64[p.start] = ((unsigned long long)fkey(0.0f) << 32);
      c.tokidx[p.start] = 0;
      c.first_rank[p.start] = 0;
      c.cum[0] = 0; c.cum[1] = 0;
      ftok[0] = 0; flink[0] = 0;
      if (*hashp < 1000) *hashp = 1000;
      s_i[1] = *hashp;
    }
    __syncthreads();
    strict_close_frame(p, c, 0, 0, p.beam, s_i[1], 1u, 0);
    if (threadIdx.x == 0) { ftok[1] = counters[0]; flink[1] = counters[1]; counters[2] = 0; }
    __syncthreads();
    if (threadIdx.x == 0) {     // the start token has no predecessor, wherever the ordering put it
      for (int i = 0; i < ftok[1]; ++i) if (c.tok_state[i] == p.start && c.tok_cost[i] == 0.0f) { c.tok_bp[i] = -1; break; }
    }
    __syncthreads();
  }

  for (int frame = counters[2]; frame < n_fed; ++frame) {
    if (counters[3] != 0) break;
    const int tb = ftok[frame], te = ftok[frame + 1], n = te - tb;
    if (n >= p.fc) { if (threadIdx.x == 0) counters[3] = 3; __syncthreads(); break; }
    const float* ll = logp + (size_t)frame * p.C;
    // ---- GetCutoff (:651-720) over every token of the list, those beyond the previous cutoff included
    float v = INFINITY;
    for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) v = fminf(v, c.tok_cost[i]);
    const float best = block_min(v, sred);
    float cur_cutoff, adaptive_beam;
    {
      const float beam_cutoff = __fadd_rn(best, p.beam);
      float max_cut = INFINITY, min_cut = INFINITY;
      if (n > p.max_active) max_cut = select_kth(c.tok_cost + tb, n, p.max_active, hist, sh);
      if (max_cut < beam_cutoff) {
        adaptive_beam = __fadd_rn(__fsub_rn(max_cut, best), p.beam_delta);
        cur_cutoff = max_cut;
      } else {
        if (n > p.min_active) min_cut = p.min_active == 0 ? best : select_kth(c.tok_cost + tb, n, p.min_active, hist, sh);
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
    // ---- PossiblyResizeHash (:216-222) + prepass on the first best token (:756-774)
    if (threadIdx.x == 0) {
      s_prep = INFINITY;
      const int want = (int)((float)n * 2.0f);
      if (want > *hashp) *hashp = want;
      s_i[1] = *hashp;
      if (*hashp > p.hcap) counters[3] = 3;
    }
    __syncthreads();
    const int hs = s_i[1];
    if (counters[3] != 0) break;
    {
      int cand = 0x7fffffff;
      for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS)
        if (c.tok_cost[i] == best) { cand = i; break; }
      for (int o = 16; o; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
      if ((threadIdx.x & 31) == 0) s_cand[threadIdx.x >> 5] = cand;
      __syncthreads();
      if (threadIdx.x == 0) {
        int b = s_cand[0];
        for (int i = 1; i < DEC_WARPS; ++i) b = min(b, s_cand[i]);
        float nc = INFINITY;
        if (b != 0x7fffffff) {
          const int st = c.tok_state[b];
          const float tc = c.tok_cost[b];
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
    // ---- per token: out-degree (sequence ranks) and cheapest emitting arc; exclusive scans in list order
    for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
      const float tc = c.tok_cost[i];
      const int st = c.tok_state[i];
      c.cum[i - tb] = (int)(p.off[st + 1] - p.off[st]);
      float m = INFINITY;
      if (tc <= cur_cutoff) {
        for (long long a = p.off[st]; a < p.off[st + 1]; ++a) {
          const DArc arc = p.arcs[a];
          if (arc.il == 0) continue;
          const float ac = __fsub_rn(cost_offset, __fmul_rn(p.acoustic_scale, ll[arc.il - 1]));
          const float g = st != arc.next ? __fadd_rn(arc.w, p.length_penalty) : arc.w;
          m = fminf(m, __fadd_rn(__fadd_rn(tc, ac), g));
        }
      }
      c.premin[i - tb] = m;
    }
    __syncthreads();
    const float next_best = block_scan_sum_min(c.cum, c.premin, n, s_wi, s_wf);
    if (c.cum[n] < 0 || c.cum[n] >= RANK_NONE) { if (threadIdx.x == 0) counters[3] = 3; __syncthreads(); break; }
    const float next_cutoff = fminf(s_prep, __fadd_rn(next_best, adaptive_beam));
    // ---- the sequential walk, in parallel: each token starts from the cutoff its predecessors left behind
    const int tb_new = counters[0], lb_new = counters[1];
    __syncthreads();
    for (int i = tb + threadIdx.x; i < te; i += DEC_THREADS) {
      const float tc = c.tok_cost[i];
      if (!(tc <= cur_cutoff)) continue;
      const int st = c.tok_state[i];
      float rc = fminf(s_prep, __fadd_rn(c.premin[i - tb], adaptive_beam));
      const int rank0 = c.cum[i - tb];
      const long long a0 = p.off[st];
      for (long long a = a0; a < p.off[st + 1]; ++a) {
        const DArc arc = p.arcs[a];
        if (arc.il == 0) continue;
        const float ac = __fsub_rn(cost_offset, __fmul_rn(p.acoustic_scale, ll[arc.il - 1]));
        const float g = st != arc.next ? __fadd_rn(arc.w, p.length_penalty) : arc.w;
        const float tot = __fadd_rn(__fadd_rn(tc, ac), g);
        if (tot >= rc) continue;
        const float tight = __fadd_rn(tot, adaptive_beam);
        if (tight < rc) rc = tight;
        const int li = atomicAdd(&counters[1], 1);
        if (li >= p.link_cap) { counters[3] = 2; continue; }
        const int rank = rank0 + (int)(a - a0);
        const unsigned long long key = ((unsigned long long)fkey(tot) << 32) | (unsigned int)rank;
        const unsigned long long old = atomicMin(&c.best64[arc.next], key);
        if (old == EMPTY64) {
          const int ns = atomicAdd(&counters[0], 1);
          if (ns < p.tok_cap) c.tok_state[ns] = arc.next;
          else counters[3] = 1;
        }
        atomicMin(&c.first_rank[arc.next], rank);
        c.links[li] = DLink{i, arc.next /* state for now */, arc.il, arc.ol, g, ac};
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) { counters[0] = min(counters[0], p.tok_cap); counters[1] = min(counters[1], p.link_cap); coff[frame] = cost_offset; }
    __syncthreads();
    if (counters[3] != 0) break;
    strict_close_frame(p, c, tb_new, lb_new, next_cutoff, hs, (unsigned int)max(c.cum[n], 1), tb);
    if (threadIdx.x == 0) { ftok[frame + 2] = counters[0]; flink[frame + 2] = counters[1]; counters[2] = frame + 1; }
    __syncthreads();
  }
}
