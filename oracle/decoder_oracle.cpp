// CPU oracle for the n-gram CTC decoders.  TEST INFRASTRUCTURE ONLY -- nothing in the product links
// or loads this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
//
// It restates, serially and in the reference's own order of operations, the algorithms of
//   language_model/runtime/core/decoder/ctc_prefix_beam_search.{h,cc}        (CtcPrefixBeamSearch)
//   language_model/runtime/core/utils/utils.cc:24-30                          (LogAdd)
//   language_model/runtime/core/decoder/ctc_wfst_beam_search.{h,cc}           (frame feeder, blank skipping)
//   language_model/runtime/core/kaldi/decoder/lattice-faster-decoder.cc       (token passing, pruning;
//                                                                              incl. the local length_penalty patch :767-769,797-799)
//   language_model/runtime/core/kaldi/decoder/lattice-faster-online-decoder.cc:59-177 (back-pointer best path)
//   language_model/runtime/core/kaldi/util/hash-list-inl.h                    (iteration order of the token hash)
//   language_model/runtime/core/decoder/brain_speech_decoder.cc:113-137       (UpdateResult)
//   language_model/runtime/core/utils/string.cc:121-146                       (ProcessBlank)
//   language_model/runtime/server/x86/python/lm_decoder.cc:14-49              (DecodeNumpy / DecodeNumpyLogProbs)
//
// PARITY STATUS.  The reference decoder cannot be built offline (it needs OpenFST 1.6.5, libtorch 1.13.1, glog,
// gflags fetched from the network).  The prefix search is pinned by the reference's own golden
// 3x3 vector (ctc_prefix_beam_search_test.cc:18-59).  The WFST search has NO reference test or runnable
// binary: "parity unpinned" -- this file is a line-by-line restatement reviewed against the sources above, and
// the n-best extraction follows the observed contract of SURVEY.md Appendix B (OpenFST's ShortestPath /
// DeterminizeLatticePruned sources are not available): the n cheapest DISTINCT word sequences whose best
// alignment lies within lattice_beam of the best path, each with the (graph, acoustic) cost of that alignment.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

const float kFloatMax = std::numeric_limits<float>::max();
const float kInf = std::numeric_limits<float>::infinity();

// ------------------------------------------------------------------------------------------------
// LogAdd (utils.cc:24-30)
float LogAdd(float x, float y) {
  static float num_min = -std::numeric_limits<float>::max();
  if (x <= num_min) return y;
  if (y <= num_min) return x;
  float xmax = std::max(x, y);
  return std::log(std::exp(x - xmax) + std::exp(y - xmax)) + xmax;
}

// ------------------------------------------------------------------------------------------------
// CTC prefix beam search (ctc_prefix_beam_search.cc:44-136)
struct PrefixScore {
  float s = -kFloatMax, ns = -kFloatMax, v_s = -kFloatMax, v_ns = -kFloatMax, cur_token_prob = -kFloatMax;
  std::vector<int> times_s, times_ns;
  float score() const { return LogAdd(s, ns); }
  float viterbi_score() const { return v_s > v_ns ? v_s : v_ns; }
  const std::vector<int>& times() const { return v_s > v_ns ? times_s : times_ns; }
};
struct PrefixHash {
  size_t operator()(const std::vector<int>& prefix) const {
    size_t h = 0;
    for (int id : prefix) h = id + 31 * h;
    return h;
  }
};
typedef std::unordered_map<std::vector<int>, PrefixScore, PrefixHash> HypMap;

struct PrefixSearch {
  int blank = 0, first_beam = 10, second_beam = 10, abs_t = 0;
  HypMap cur;
  std::vector<std::vector<int>> hyps, times;
  std::vector<float> like, viterbi;
  void Reset() {
    hyps.clear(); like.clear(); cur.clear(); viterbi.clear(); times.clear();
    abs_t = 0;
    PrefixScore p;
    p.s = 0.0f; p.ns = -kFloatMax; p.v_s = 0.0f; p.v_ns = 0.0f;
    cur[std::vector<int>()] = p;
  }
  static bool Cmp(const std::pair<std::vector<int>, PrefixScore>& a, const std::pair<std::vector<int>, PrefixScore>& b) {
    return a.second.score() > b.second.score();
  }
  void Search(const float* logp, int T, int C) {
    for (int t = 0; t < T; ++t, ++abs_t) {
      const float* row = logp + (size_t)t * C;
      HypMap next;
      // 1. first beam: top-k tokens (torch.topk: descending value; ties -> lower index first)
      std::vector<int> idx(C);
      for (int i = 0; i < C; ++i) idx[i] = i;
      const int k = std::min(first_beam, C);
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return row[a] > row[b]; });
      // 2. token passing
      for (int i = 0; i < k; ++i) {
        const int id = idx[i];
        const float prob = row[id];
        for (const auto& it : cur) {
          const std::vector<int>& prefix = it.first;
          const PrefixScore& ps = it.second;
          if (id == blank) {
            PrefixScore& n = next[prefix];
            n.s = LogAdd(n.s, ps.score() + prob);
            n.v_s = ps.viterbi_score() + prob;
            n.times_s = ps.times();
          } else if (!prefix.empty() && id == prefix.back()) {
            PrefixScore& n1 = next[prefix];
            n1.ns = LogAdd(n1.ns, ps.ns + prob);
            if (n1.v_ns < ps.v_ns + prob) {
              n1.v_ns = ps.v_ns + prob;
              if (n1.cur_token_prob < prob) {
                n1.cur_token_prob = prob;
                n1.times_ns = ps.times_ns;
                if (!n1.times_ns.empty()) n1.times_ns.back() = abs_t;
              }
            }
            std::vector<int> np(prefix);
            np.push_back(id);
            PrefixScore& n2 = next[np];
            n2.ns = LogAdd(n2.ns, ps.s + prob);
            if (n2.v_ns < ps.v_s + prob) {
              n2.v_ns = ps.v_s + prob;
              n2.cur_token_prob = prob;
              n2.times_ns = ps.times_s;
              n2.times_ns.push_back(abs_t);
            }
          } else {
            std::vector<int> np(prefix);
            np.push_back(id);
            PrefixScore& n = next[np];
            n.ns = LogAdd(n.ns, ps.score() + prob);
            if (n.v_ns < ps.viterbi_score() + prob) {
              n.v_ns = ps.viterbi_score() + prob;
              n.cur_token_prob = prob;
              n.times_ns = ps.times();
              n.times_ns.push_back(abs_t);
            }
          }
        }
      }
      // 3. second beam
      std::vector<std::pair<std::vector<int>, PrefixScore>> arr(next.begin(), next.end());
      const int sb = std::min((int)arr.size(), second_beam);
      std::nth_element(arr.begin(), arr.begin() + sb, arr.end(), Cmp);
      arr.resize(sb);
      std::sort(arr.begin(), arr.end(), Cmp);
      // 4. new result
      cur.clear(); hyps.clear(); like.clear(); viterbi.clear(); times.clear();
      for (auto& item : arr) {
        cur[item.first] = item.second;
        hyps.push_back(item.first);
        like.push_back(item.second.score());
        viterbi.push_back(item.second.viterbi_score());
        times.push_back(item.second.times());
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// OpenFST binary reader: fst type "vector", arc type "standard" (what Fst<StdArc>::Read accepts for TLG.fst,
// brain_speech_decoder.h:53).  Layout per the OpenFST 1.6 sources as documented in SURVEY.md 2a(7).
struct Arc { int il, ol; float w; int next; };
struct Graph {
  int start = -1;
  std::vector<float> fin;        // +inf = not final
  std::vector<int64_t> off;      // arcs of state s: [off[s], off[s+1])
  std::vector<Arc> arcs;
  std::vector<int> n_ieps;       // number of input-epsilon arcs per state
};
bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }
bool rd_str(FILE* f, std::string* s) {
  int32_t n;
  if (!rd(f, &n, 4) || n < 0 || n > 1 << 20) return false;
  s->resize(n);
  return n == 0 || rd(f, &(*s)[0], n);
}
bool skip_symtab(FILE* f) {   // SymbolTable binary: magic, name, available_key, size, {symbol, key}*
  int32_t magic;
  std::string name;
  int64_t avail, size;
  if (!rd(f, &magic, 4) || !rd_str(f, &name) || !rd(f, &avail, 8) || !rd(f, &size, 8)) return false;
  for (int64_t i = 0; i < size; ++i) {
    std::string sym;
    int64_t key;
    if (!rd_str(f, &sym) || !rd(f, &key, 8)) return false;
  }
  return true;
}
bool ReadFst(const char* path, Graph* g, std::string* err) {
  FILE* f = fopen(path, "rb");
  if (!f) { *err = std::string("cannot open ") + path; return false; }
  int32_t magic, version, flags;
  std::string fsttype, arctype;
  uint64_t props;
  int64_t start, nstates, narcs;
  bool ok = rd(f, &magic, 4) && magic == 2125659606 && rd_str(f, &fsttype) && rd_str(f, &arctype) && rd(f, &version, 4) &&
            rd(f, &flags, 4) && rd(f, &props, 8) && rd(f, &start, 8) && rd(f, &nstates, 8) && rd(f, &narcs, 8);
  if (!ok || fsttype != "vector" || arctype != "standard") { *err = "not an OpenFST vector/standard file"; fclose(f); return false; }
  if ((flags & 1) && !skip_symtab(f)) { *err = "bad input symbol table"; fclose(f); return false; }
  if ((flags & 2) && !skip_symtab(f)) { *err = "bad output symbol table"; fclose(f); return false; }
  g->start = (int)start;
  g->fin.resize(nstates); g->off.assign(nstates + 1, 0); g->n_ieps.assign(nstates, 0);
  g->arcs.clear();
  for (int64_t s = 0; s < nstates; ++s) {
    float fw;
    int64_t na;
    if (!rd(f, &fw, 4) || !rd(f, &na, 8)) { *err = "truncated state"; fclose(f); return false; }
    g->fin[s] = fw;
    g->off[s] = (int64_t)g->arcs.size();
    for (int64_t a = 0; a < na; ++a) {
      int32_t v[2]; float w; int32_t nx;
      if (!rd(f, v, 8) || !rd(f, &w, 4) || !rd(f, &nx, 4)) { *err = "truncated arc"; fclose(f); return false; }
      g->arcs.push_back({v[0], v[1], w, nx});
      if (v[0] == 0) g->n_ieps[s]++;
    }
  }
  g->off[nstates] = (int64_t)g->arcs.size();
  fclose(f);
  return true;
}

// ------------------------------------------------------------------------------------------------
// Token passing (lattice-faster-decoder.cc, BackpointerToken variant)
struct Token;
struct Link {
  Token* next_tok; int il, ol; float graph_cost, acoustic_cost; Link* next;
};
struct Token {
  float tot_cost, extra_cost; Link* links; Token* next; Token* backpointer;
};
struct Elem { int key; Token* val; };

// HashList restated with the same iteration order: buckets in order of first occupation, insertion order inside.
struct HashList {
  size_t hash_size = 0;
  std::vector<std::vector<Elem>> buckets;
  std::vector<size_t> order;
  void SetSize(size_t sz) {
    hash_size = sz;
    if (sz > buckets.size()) buckets.resize(sz);
  }
  size_t Size() const { return hash_size; }
  std::vector<Elem> Clear() {      // returns the list (in list order) and empties the table
    std::vector<Elem> out;
    for (size_t b : order) {
      for (const Elem& e : buckets[b]) out.push_back(e);
      buckets[b].clear();
    }
    order.clear();
    return out;
  }
  std::vector<Elem> GetList() const {
    std::vector<Elem> out;
    for (size_t b : order)
      for (const Elem& e : buckets[b]) out.push_back(e);
    return out;
  }
  Elem* Insert(int key, Token* val) {
    const size_t idx = (size_t)key % hash_size;
    std::vector<Elem>& b = buckets[idx];
    for (Elem& e : b)
      if (e.key == key) return &e;
    if (b.empty()) order.push_back(idx);
    b.push_back({key, val});
    return &b.back();
  }
};

struct Config {
  float beam = 16.0f; int max_active = std::numeric_limits<int>::max(); int min_active = 200; float lattice_beam = 10.0f;
  int prune_interval = 25; float beam_delta = 0.5f, hash_ratio = 2.0f, prune_scale = 0.1f, length_penalty = 0.0f;
};

struct TokenList { Token* toks = nullptr; bool must_prune_forward_links = true, must_prune_tokens = true; };

struct Decoder {
  const Graph* fst = nullptr;
  Config cfg;
  HashList toks;
  std::vector<TokenList> active;
  std::vector<float> cost_offsets, tmp;
  int num_toks = 0;
  bool finalized = false;
  std::unordered_map<Token*, float> final_costs;
  float final_relative_cost = 0, final_best_cost = 0;
  const float* loglikes = nullptr;   // current frame, already scaled

  ~Decoder() { ClearActive(); }
  int NumFramesDecoded() const { return (int)active.size() - 1; }

  void DeleteLinks(Token* t) {
    for (Link* l = t->links; l;) { Link* m = l->next; delete l; l = m; }
    t->links = nullptr;
  }
  void ClearActive() {
    for (auto& tl : active)
      for (Token* t = tl.toks; t;) { DeleteLinks(t); Token* n = t->next; delete t; --num_toks; t = n; }
    active.clear();
  }
  void InitDecoding() {                                    // :58-75
    toks.Clear();
    if (toks.Size() == 0) toks.SetSize(1000);
    cost_offsets.clear();
    ClearActive();
    num_toks = 0; finalized = false; final_costs.clear();
    active.resize(1);
    Token* st = new Token{0.0f, 0.0f, nullptr, nullptr, nullptr};
    active[0].toks = st;
    toks.Insert(fst->start, st);
    num_toks++;
    ProcessNonemitting(cfg.beam);
  }
  Elem* FindOrAddToken(int state, int frame_plus_one, float tot_cost, Token* bp, bool* changed) {   // :250-292
    Token*& head = active[frame_plus_one].toks;
    Elem* e = toks.Insert(state, nullptr);
    if (e->val == nullptr) {
      Token* nt = new Token{tot_cost, 0.0f, nullptr, head, bp};
      head = nt;
      num_toks++;
      e->val = nt;
      if (changed) *changed = true;
    } else {
      Token* t = e->val;
      if (t->tot_cost > tot_cost) {
        t->tot_cost = tot_cost;
        t->backpointer = bp;
        if (changed) *changed = true;
      } else if (changed) {
        *changed = false;
      }
    }
    return e;
  }
  void PossiblyResizeHash(size_t n) {                      // :216-222
    size_t nsz = (size_t)((float)n * cfg.hash_ratio);
    if (nsz > toks.Size()) toks.SetSize(nsz);
  }
  float GetCutoff(const std::vector<Elem>& list, size_t* tok_count, float* adaptive_beam, int* best_idx) {   // :651-720
    float best = kInf;
    size_t count = 0;
    if (cfg.max_active == std::numeric_limits<int>::max() && cfg.min_active == 0) {
      for (size_t i = 0; i < list.size(); ++i, ++count) {
        const float w = list[i].val->tot_cost;
        if (w < best) { best = w; if (best_idx) *best_idx = (int)i; }
      }
      if (tok_count) *tok_count = count;
      if (adaptive_beam) *adaptive_beam = cfg.beam;
      return best + cfg.beam;
    }
    tmp.clear();
    for (size_t i = 0; i < list.size(); ++i, ++count) {
      const float w = list[i].val->tot_cost;
      tmp.push_back(w);
      if (w < best) { best = w; if (best_idx) *best_idx = (int)i; }
    }
    if (tok_count) *tok_count = count;
    float beam_cutoff = best + cfg.beam, min_active_cutoff = kInf, max_active_cutoff = kInf;
    if (tmp.size() > (size_t)cfg.max_active) {
      std::nth_element(tmp.begin(), tmp.begin() + cfg.max_active, tmp.end());
      max_active_cutoff = tmp[cfg.max_active];
    }
    if (max_active_cutoff < beam_cutoff) {
      if (adaptive_beam) *adaptive_beam = max_active_cutoff - best + cfg.beam_delta;
      return max_active_cutoff;
    }
    if (tmp.size() > (size_t)cfg.min_active) {
      if (cfg.min_active == 0) min_active_cutoff = best;
      else {
        std::nth_element(tmp.begin(), tmp.begin() + cfg.min_active,
                         tmp.size() > (size_t)cfg.max_active ? tmp.begin() + cfg.max_active : tmp.end());
        min_active_cutoff = tmp[cfg.min_active];
      }
    }
    if (min_active_cutoff > beam_cutoff) {
      if (adaptive_beam) *adaptive_beam = min_active_cutoff - best + cfg.beam_delta;
      return min_active_cutoff;
    }
    *adaptive_beam = cfg.beam;
    return beam_cutoff;
  }
  float LogLike(int ilabel) const { return loglikes[ilabel - 1]; }   // ctc_wfst_beam_search.cc:27-33 (scale applied by caller)

  float ProcessEmitting() {                                // :723-824
    const int frame = (int)active.size() - 1;
    active.resize(active.size() + 1);
    std::vector<Elem> final_toks = toks.Clear();
    int best_idx = -1;
    float adaptive_beam;
    size_t tok_cnt;
    const float cur_cutoff = GetCutoff(final_toks, &tok_cnt, &adaptive_beam, &best_idx);
    PossiblyResizeHash(tok_cnt);
    float next_cutoff = kInf, cost_offset = 0.0f;
    if (best_idx >= 0) {
      const int state = final_toks[best_idx].key;
      Token* tok = final_toks[best_idx].val;
      cost_offset = -tok->tot_cost;
      for (int64_t a = fst->off[state]; a < fst->off[state + 1]; ++a) {
        const Arc& arc = fst->arcs[a];
        if (arc.il != 0) {
          float new_weight = arc.w + cost_offset - LogLike(arc.il) + tok->tot_cost;
          if (state != arc.next) new_weight += cfg.length_penalty;
          if (new_weight + adaptive_beam < next_cutoff) next_cutoff = new_weight + adaptive_beam;
        }
      }
    }
    cost_offsets.resize(frame + 1, 0.0f);
    cost_offsets[frame] = cost_offset;
    for (const Elem& e : final_toks) {
      const int state = e.key;
      Token* tok = e.val;
      if (tok->tot_cost <= cur_cutoff) {
        for (int64_t a = fst->off[state]; a < fst->off[state + 1]; ++a) {
          const Arc& arc = fst->arcs[a];
          if (arc.il != 0) {
            const float ac_cost = cost_offset - LogLike(arc.il);
            float graph_cost = arc.w;
            if (state != arc.next) graph_cost += cfg.length_penalty;
            const float cur_cost = tok->tot_cost, tot_cost = cur_cost + ac_cost + graph_cost;
            if (tot_cost >= next_cutoff) continue;
            else if (tot_cost + adaptive_beam < next_cutoff) next_cutoff = tot_cost + adaptive_beam;
            Elem* en = FindOrAddToken(arc.next, frame + 1, tot_cost, tok, nullptr);
            tok->links = new Link{en->val, arc.il, arc.ol, graph_cost, ac_cost, tok->links};
          }
        }
      }
    }
    return next_cutoff;
  }
  void ProcessNonemitting(float cutoff) {                  // :839-909
    const int frame = (int)active.size() - 2;
    std::vector<Elem> queue;
    for (const Elem& e : toks.GetList())
      if (fst->n_ieps[e.key] != 0) queue.push_back(e);
    while (!queue.empty()) {
      const Elem e = queue.back();
      queue.pop_back();
      const int state = e.key;
      Token* tok = e.val;
      const float cur_cost = tok->tot_cost;
      if (cur_cost >= cutoff) continue;
      DeleteLinks(tok);
      for (int64_t a = fst->off[state]; a < fst->off[state + 1]; ++a) {
        const Arc& arc = fst->arcs[a];
        if (arc.il == 0) {
          const float graph_cost = arc.w, tot_cost = cur_cost + graph_cost;
          if (tot_cost < cutoff) {
            bool changed;
            Elem* en = FindOrAddToken(arc.next, frame + 1, tot_cost, tok, &changed);
            tok->links = new Link{en->val, 0, arc.ol, graph_cost, 0.0f, tok->links};
            if (changed && fst->n_ieps[arc.next] != 0) queue.push_back(*en);
          }
        }
      }
    }
  }
  void PruneForwardLinks(int f, bool* extra_costs_changed, bool* links_pruned, float delta) {   // :298-373
    *extra_costs_changed = false; *links_pruned = false;
    bool changed = true;
    while (changed) {
      changed = false;
      for (Token* tok = active[f].toks; tok; tok = tok->next) {
        Link *link, *prev = nullptr;
        float tok_extra = kInf;
        for (link = tok->links; link;) {
          Token* nt = link->next_tok;
          float lec = nt->extra_cost + ((tok->tot_cost + link->acoustic_cost + link->graph_cost) - nt->tot_cost);
          if (lec > cfg.lattice_beam) {
            Link* nl = link->next;
            if (prev) prev->next = nl; else tok->links = nl;
            delete link;
            link = nl;
            *links_pruned = true;
          } else {
            if (lec < 0.0f) lec = 0.0f;
            if (lec < tok_extra) tok_extra = lec;
            prev = link;
            link = link->next;
          }
        }
        if (fabsf(tok_extra - tok->extra_cost) > delta) changed = true;
        tok->extra_cost = tok_extra;
      }
      if (changed) *extra_costs_changed = true;
    }
  }
  void ComputeFinalCosts(std::unordered_map<Token*, float>* fc, float* rel, float* best_out) const {   // :544-581
    if (fc) fc->clear();
    float best = kInf, best_final = kInf;
    for (const Elem& e : toks.GetList()) {
      const float final_cost = fst->fin[e.key];
      const float cost = e.val->tot_cost, cwf = cost + final_cost;
      best = std::min(cost, best);
      best_final = std::min(cwf, best_final);
      if (fc && final_cost != kInf) (*fc)[e.val] = final_cost;
    }
    if (rel) *rel = (best == kInf && best_final == kInf) ? kInf : best_final - best;
    if (best_out) *best_out = best_final != kInf ? best_final : best;
  }
  static bool ApproxEqual(float a, float b, float tol) {   // kaldi-math.h
    if (a == b) return true;
    float diff = fabsf(a - b);
    if (diff == kInf || diff != diff) return false;
    return diff <= tol * (fabsf(a) + fabsf(b));
  }
  void PruneForwardLinksFinal() {                          // :379-465
    const int f = (int)active.size() - 1;
    ComputeFinalCosts(&final_costs, &final_relative_cost, &final_best_cost);
    finalized = true;
    toks.Clear();
    bool changed = true;
    const float delta = 1.0e-05f;
    while (changed) {
      changed = false;
      for (Token* tok = active[f].toks; tok; tok = tok->next) {
        Link *link, *prev = nullptr;
        float final_cost;
        if (final_costs.empty()) final_cost = 0.0f;
        else {
          auto it = final_costs.find(tok);
          final_cost = it != final_costs.end() ? it->second : kInf;
        }
        float tok_extra = tok->tot_cost + final_cost - final_best_cost;
        for (link = tok->links; link;) {
          Token* nt = link->next_tok;
          float lec = nt->extra_cost + ((tok->tot_cost + link->acoustic_cost + link->graph_cost) - nt->tot_cost);
          if (lec > cfg.lattice_beam) {
            Link* nl = link->next;
            if (prev) prev->next = nl; else tok->links = nl;
            delete link;
            link = nl;
          } else {
            if (lec < 0.0f) lec = 0.0f;
            if (lec < tok_extra) tok_extra = lec;
            prev = link;
            link = link->next;
          }
        }
        if (tok_extra > cfg.lattice_beam) tok_extra = kInf;
        if (!ApproxEqual(tok->extra_cost, tok_extra, delta)) changed = true;
        tok->extra_cost = tok_extra;
      }
    }
  }
  void PruneTokensForFrame(int f) {                        // :485-506
    Token*& head = active[f].toks;
    Token *tok, *next, *prev = nullptr;
    for (tok = head; tok; tok = next) {
      next = tok->next;
      if (tok->extra_cost == kInf) {
        if (prev) prev->next = tok->next; else head = tok->next;
        DeleteLinks(tok);
        delete tok;
        num_toks--;
      } else prev = tok;
    }
  }
  void PruneActiveTokens(float delta) {                    // :514-541
    const int cur = NumFramesDecoded();
    for (int f = cur - 1; f >= 0; f--) {
      if (active[f].must_prune_forward_links) {
        bool ecc = false, lp = false;
        PruneForwardLinks(f, &ecc, &lp, delta);
        if (ecc && f > 0) active[f - 1].must_prune_forward_links = true;
        if (lp) active[f].must_prune_tokens = true;
        active[f].must_prune_forward_links = false;
      }
      if (f + 1 < cur && active[f + 1].must_prune_tokens) {
        PruneTokensForFrame(f + 1);
        active[f + 1].must_prune_tokens = false;
      }
    }
  }
  void AdvanceOneFrame(const float* scaled_loglikes) {     // :584-626 with max_num_frames == 1
    loglikes = scaled_loglikes;
    if (NumFramesDecoded() % cfg.prune_interval == 0) PruneActiveTokens(cfg.lattice_beam * cfg.prune_scale);
    const float cutoff = ProcessEmitting();
    ProcessNonemitting(cutoff);
  }
  void FinalizeDecoding() {                                // :632-647
    const int fpo = NumFramesDecoded();
    PruneForwardLinksFinal();
    for (int f = fpo - 1; f >= 0; f--) {
      bool b1, b2;
      PruneForwardLinks(f, &b1, &b2, 0.0f);
      PruneTokensForFrame(f + 1);
    }
    PruneTokensForFrame(0);
  }
  // lattice-faster-online-decoder.cc:59-177.  Returns olabels (words) of the best path and its (graph, acoustic) weight.
  bool GetBestPath(bool use_final, std::vector<int>* words, std::vector<int>* ilabels, float* graph, float* acoustic) const {
    words->clear(); ilabels->clear();
    std::unordered_map<Token*, float> local;
    const std::unordered_map<Token*, float>& fc = finalized ? final_costs : local;
    if (!finalized && use_final) ComputeFinalCosts(&local, nullptr, nullptr);
    float best_cost = kInf, best_final = 0.0f;
    Token* best = nullptr;
    for (Token* tok = active.back().toks; tok; tok = tok->next) {
      float cost = tok->tot_cost, final_cost = 0.0f;
      if (use_final && !fc.empty()) {
        auto it = fc.find(tok);
        if (it != fc.end()) { final_cost = it->second; cost += final_cost; }
        else cost = kInf;
      }
      if (cost < best_cost) { best_cost = cost; best = tok; best_final = final_cost; }
    }
    if (!best) return false;
    double g = best_final, a = 0.0;
    // the reference accumulates LatticeWeight along the linear lattice with float Times(); keep float
    float gf = best_final, af = 0.0f;
    (void)g; (void)a;
    std::vector<int> rw, ri;
    std::vector<std::pair<float, float>> ws;
    Token* tok = best;
    int cur_t = NumFramesDecoded() - 1;
    while (tok) {
      int il = 0, ol = 0;
      float gc = 0.0f, ac = 0.0f;
      int step = 0;
      if (tok->backpointer) {
        float best_link = kInf;
        for (Link* l = tok->backpointer->links; l; l = l->next) {
          if (l->next_tok == tok) {
            float graph_cost = l->graph_cost, acoustic_cost = l->acoustic_cost;
            const float cost = graph_cost + acoustic_cost;
            if (cost < best_link) {
              il = l->il; ol = l->ol;
              if (l->il != 0) { acoustic_cost -= cost_offsets[cur_t]; step = -1; }
              else step = 0;
              gc = graph_cost; ac = acoustic_cost;
              best_link = cost;
            }
          }
        }
        if (best_link == kInf) return false;
      }
      ws.push_back({gc, ac});
      if (il != 0) ri.push_back(il);
      if (ol != 0) rw.push_back(ol);
      tok = tok->backpointer;
      cur_t += step;
    }
    // GetLinearSymbolSequence multiplies weights from the start state forward: Times is float addition
    float tg = 0.0f, ta = 0.0f;   // LatticeWeight::One()
    for (int i = (int)ws.size() - 1; i >= 0; --i) { tg = tg + ws[i].first; ta = ta + ws[i].second; }
    tg = tg + gf; ta = ta + af;
    std::reverse(rw.begin(), rw.end());
    std::reverse(ri.begin(), ri.end());
    *words = rw; *ilabels = ri; *graph = tg; *acoustic = ta;
    return true;
  }
};

// ------------------------------------------------------------------------------------------------
// n-best: all distinct word sequences within lattice_beam over the pruned token lattice.
struct NBestEntry { std::vector<int> words; float graph, acoustic; };
struct SeqTrie {
  std::vector<std::pair<int, int>> nodes{{-1, 0}};   // (parent, word); node 0 = empty sequence
  std::map<std::pair<int, int>, int> idx;
  int Extend(int node, int word) {
    auto key = std::make_pair(node, word);
    auto it = idx.find(key);
    if (it != idx.end()) return it->second;
    nodes.push_back(key);
    return idx[key] = (int)nodes.size() - 1;
  }
  std::vector<int> Words(int node) const {
    std::vector<int> w;
    for (; node > 0; node = nodes[node].first) w.push_back(nodes[node].second);
    std::reverse(w.begin(), w.end());
    return w;
  }
};
struct Hyp { float g, a; };
inline bool Better(const Hyp& x, const Hyp& y) {   // LatticeWeight Compare (lattice-weight.h:291-309): smaller total, then smaller graph
  const float fx = x.g + x.a, fy = y.g + y.a;
  if (fx < fy) return true;
  if (fx > fy) return false;
  return x.g < y.g;
}
std::vector<NBestEntry> NBest(const Decoder& d, int nbest) {
  std::vector<NBestEntry> out;
  const int F = (int)d.active.size();
  // topological order: frames ascending; inside a frame, order tokens so that epsilon links go forward
  std::vector<Token*> order;
  std::unordered_map<Token*, int> pos;
  for (int f = 0; f < F; ++f) {
    std::vector<Token*> ft;
    for (Token* t = d.active[f].toks; t; t = t->next) ft.push_back(t);
    std::unordered_map<Token*, int> indeg;
    for (Token* t : ft) indeg[t];
    for (Token* t : ft)
      for (Link* l = t->links; l; l = l->next)
        if (l->il == 0) indeg[l->next_tok]++;
    std::vector<Token*> ready;
    for (Token* t : ft) if (indeg[t] == 0) ready.push_back(t);
    size_t done = 0;
    while (done < ready.size()) {
      Token* t = ready[done++];
      for (Link* l = t->links; l; l = l->next)
        if (l->il == 0 && --indeg[l->next_tok] == 0) ready.push_back(l->next_tok);
    }
    for (Token* t : ready) { pos[t] = (int)order.size(); order.push_back(t); }
  }
  if (order.empty()) return out;
  // backward best cost-to-final (beta) for pruning
  std::unordered_map<Token*, float> beta;
  for (Token* t : order) beta[t] = kInf;
  for (Token* t = d.active[F - 1].toks; t; t = t->next) {
    float fc = 0.0f;
    if (!d.final_costs.empty()) {
      auto it = d.final_costs.find(t);
      fc = it != d.final_costs.end() ? it->second : kInf;
    }
    beta[t] = fc;
  }
  std::vector<int> frame_of(order.size(), 0);
  {
    int k = 0;
    for (int f = 0; f < F; ++f)
      for (Token* t = d.active[f].toks; t; t = t->next) { frame_of[pos[t]] = f; ++k; }
  }
  for (int i = (int)order.size() - 1; i >= 0; --i) {
    Token* t = order[i];
    for (Link* l = t->links; l; l = l->next) {
      float ac = l->acoustic_cost;
      if (l->il != 0) ac -= d.cost_offsets[frame_of[i]];
      const float c = l->graph_cost + ac + beta[l->next_tok];
      if (c < beta[t]) beta[t] = c;
    }
  }
  Token* start_tok = nullptr;                    // the start token is the one without a back-pointer (InitDecoding)
  for (Token* t = d.active[0].toks; t; t = t->next)
    if (!t->backpointer) start_tok = t;
  if (!start_tok) return out;
  const float best_total = beta[start_tok];
  if (best_total == kInf) return out;
  const float limit = best_total + d.cfg.lattice_beam;
  SeqTrie trie;
  std::vector<std::map<int, Hyp>> hyps(order.size());
  hyps[pos[start_tok]][0] = Hyp{0.0f, 0.0f};
  std::map<int, Hyp> finals;
  for (size_t i = 0; i < order.size(); ++i) {
    Token* t = order[i];
    if (hyps[i].empty()) continue;
    if (frame_of[i] == F - 1 && beta[t] != kInf) {
      float fc = 0.0f;
      bool is_final = true;
      if (!d.final_costs.empty()) {
        auto it = d.final_costs.find(t);
        if (it == d.final_costs.end()) is_final = false; else fc = it->second;
      }
      if (is_final)
        for (auto& kv : hyps[i]) {
          Hyp h{kv.second.g + fc, kv.second.a};
          if (h.g + h.a > limit + 1e-4f) continue;
          auto it = finals.find(kv.first);
          if (it == finals.end() || Better(h, it->second)) finals[kv.first] = h;
        }
    }
    for (Link* l = t->links; l; l = l->next) {
      float ac = l->acoustic_cost;
      if (l->il != 0) ac -= d.cost_offsets[frame_of[i]];
      const int j = pos[l->next_tok];
      const float bnext = beta[l->next_tok];
      if (bnext == kInf) continue;
      for (auto& kv : hyps[i]) {
        Hyp h{kv.second.g + l->graph_cost, kv.second.a + ac};
        if (h.g + h.a + bnext > limit + 1e-4f) continue;
        const int node = l->ol != 0 ? trie.Extend(kv.first, l->ol) : kv.first;
        auto it = hyps[j].find(node);
        if (it == hyps[j].end() || Better(h, it->second)) hyps[j][node] = h;
      }
    }
    if (frame_of[i] < F - 1) std::map<int, Hyp>().swap(hyps[i]);
  }
  std::vector<std::pair<Hyp, int>> v;
  for (auto& kv : finals) v.push_back({kv.second, kv.first});
  std::stable_sort(v.begin(), v.end(), [](const std::pair<Hyp, int>& a, const std::pair<Hyp, int>& b) { return Better(a.first, b.first); });
  for (size_t i = 0; i < v.size() && (int)i < nbest; ++i) out.push_back({trie.Words(v[i].second), v[i].first.g, v[i].first.a});
  return out;
}

// ------------------------------------------------------------------------------------------------
// ProcessBlank (string.cc:121-146) for ASCII/UTF-8 input: collapse spaces, trim, lower-case
std::string ProcessBlank(const std::string& s) {
  std::string r;
  for (char c : s) {
    if (c != ' ') r.push_back(c);
    else if (!r.empty() && r.back() != ' ') r.push_back(' ');
  }
  if (!r.empty() && r.back() == ' ') r.pop_back();
  for (char& c : r) c = (char)tolower((unsigned char)c);
  return r;
}

// ------------------------------------------------------------------------------------------------
// BrainSpeechDecoder::Rescore (brain_speech_decoder.cc:47-101).  The lattice it starts from is CtcWfstBeamSearch::Lattice(),
// i.e. the DETERMINISED word lattice of FinalizeSearch (ctc_wfst_beam_search.cc:139-142): one path per distinct word sequence W
// within lattice_beam, weight (g_W, a_W).  LatticeRescore(lat, lm, scale) = scale the graph costs, compose with the LM acceptor,
// determinise, scale back; with scale = -1 for the LM the graph was built from and +1 for the rescoring LM this gives, per W,
//     graph' = g_W - c_old(W) + c_new(W),   acoustic' = a_W,
// where c(W) is the cheapest path of W through the LM FST INCLUDING its final cost.  ReadAndPrepareLmFst
// (kaldi/fstext/kaldi-fst-io.cc) projects the LM on its output labels, so the back-off arcs (#0 : eps on disk) are epsilon
// arcs that composition may take at any time; determinisation keeps the cheapest route.  A W the LM does not accept disappears.
// ShortestPath(lat, result_.size()) then keeps as many entries as the first pass had, ordered by graph' + acoustic'.
// The reference has no test for this path and OpenFST is not available offline: PARITY UNPINNED.
struct LmFst {
  Graph g;                                       // ilabels == olabels after projection on the output side
  bool Read(const char* path, std::string* err) {
    if (!ReadFst(path, &g, err)) return false;
    for (auto& a : g.arcs) a.il = a.ol;          // PROJECT_OUTPUT
    return true;
  }
  void Closure(std::map<int, float>* st) const {  // epsilon (back-off) arcs; back-off chains are acyclic
    std::vector<int> work;
    for (auto& kv : *st) work.push_back(kv.first);
    while (!work.empty()) {
      const int s = work.back(); work.pop_back();
      const float c = (*st)[s];
      for (long long a = g.off[s]; a < g.off[s + 1]; ++a) {
        const Arc& arc = g.arcs[a];
        if (arc.il != 0) continue;
        auto it = st->find(arc.next);
        const float nc = c + arc.w;
        if (it == st->end() || nc < it->second) { (*st)[arc.next] = nc; work.push_back(arc.next); }
      }
    }
  }
  float Cost(const std::vector<int>& words) const {
    std::map<int, float> cur;
    cur[g.start] = 0.0f;
    Closure(&cur);
    for (int w : words) {
      std::map<int, float> nxt;
      for (auto& kv : cur)
        for (long long a = g.off[kv.first]; a < g.off[kv.first + 1]; ++a) {
          const Arc& arc = g.arcs[a];
          if (arc.il != w) continue;
          auto it = nxt.find(arc.next);
          const float nc = kv.second + arc.w;
          if (it == nxt.end() || nc < it->second) nxt[arc.next] = nc;
        }
      if (nxt.empty()) return kInf;
      Closure(&nxt);
      cur.swap(nxt);
    }
    float best = kInf;
    for (auto& kv : cur)
      if (g.fin[kv.first] != kInf) best = std::min(best, kv.second + g.fin[kv.first]);
    return best;
  }
};

struct Result { float ac_score, lm_score; std::string sentence; };

// CtcWfstBeamSearch + BrainSpeechDecoder facade
struct Facade {
  Graph graph;
  std::vector<std::string> words;
  Decoder dec;
  float acoustic_scale = 1.0f, blank_skip = 0.98f;
  int nbest = 10;
  int num_frames = 0, last_best = 0;
  bool last_blank = false;
  std::vector<float> last_frame, scaled;
  std::vector<int> frames_map;
  std::vector<Result> results;
  std::vector<int> tokens_per_frame;

  void Reset() {
    num_frames = 0; frames_map.clear(); last_blank = false; last_best = 0; results.clear(); tokens_per_frame.clear();
    dec.InitDecoding();
  }
  void Feed(const float* row, int C) {
    scaled.resize(C);
    for (int c = 0; c < C; ++c) scaled[c] = acoustic_scale * row[c];
    dec.AdvanceOneFrame(scaled.data());
    tokens_per_frame.push_back((int)dec.toks.GetList().size());
  }
  void Search(const float* logp, int T, int C) {           // ctc_wfst_beam_search.cc:70-121
    if (T == 0) return;
    for (int i = 0; i < T; ++i) {
      const float* row = logp + (size_t)i * C;
      const float blank_score = std::exp(row[0]);
      if (blank_score > blank_skip) {
        last_blank = true;
        last_frame.assign(row, row + C);
      } else {
        int cur_best = 0;
        for (int c = 1; c < C; ++c) if (row[c] > row[cur_best]) cur_best = c;
        if (cur_best != 0 && last_blank && cur_best == last_best) {
          Feed(last_frame.data(), C);
          frames_map.push_back(num_frames - 1);
        }
        last_best = cur_best;
        Feed(row, C);
        frames_map.push_back(num_frames);
        last_blank = false;
      }
      num_frames++;
    }
    results.clear();
    if (!frames_map.empty()) {
      std::vector<int> w, il;
      float g, a;
      if (dec.GetBestPath(false, &w, &il, &g, &a)) Push(w, -g, -a);
    }
  }
  void Push(const std::vector<int>& w, float lm, float ac) {   // brain_speech_decoder.cc:113-137
    Result r;
    r.lm_score = lm;
    r.ac_score = ac / acoustic_scale;
    for (int id : w) r.sentence += ' ' + (id >= 0 && id < (int)words.size() ? words[id] : std::string("<unk>"));
    r.sentence = ProcessBlank(r.sentence);
    results.push_back(r);
  }
  void Finish() {                                          // ctc_wfst_beam_search.cc:123-160
    results.clear();
    if (frames_map.empty()) return;
    dec.FinalizeDecoding();
    if (nbest == 1) {
      std::vector<int> w, il;
      float g, a;
      if (dec.GetBestPath(true, &w, &il, &g, &a)) Push(w, -g, -a);
    } else {
      for (const NBestEntry& e : NBest(dec, nbest)) Push(e.words, -e.graph, -e.acoustic);
    }
  }
  // Rescore(): see the comment above LmFst.  Returns false when an LM is missing.
  bool Rescore(const LmFst* old_lm, const LmFst* new_lm) {
    if (!old_lm || !new_lm) return false;
    const size_t keep = results.size();
    struct R { std::vector<int> words; float g, a; };
    std::vector<R> all;
    for (const NBestEntry& e : NBest(dec, 0x7fffffff)) {           // every distinct word sequence within lattice_beam
      const float c_old = old_lm->Cost(e.words), c_new = new_lm->Cost(e.words);
      if (c_old == kInf || c_new == kInf) continue;
      // graph cost through the two scalings: -( -g + c_old ) then + c_new
      all.push_back({e.words, -(-e.graph + c_old) + c_new, e.acoustic});
    }
    std::stable_sort(all.begin(), all.end(), [](const R& x, const R& y) { return Better(Hyp{x.g, x.a}, Hyp{y.g, y.a}); });
    results.clear();
    for (size_t i = 0; i < all.size() && i < keep; ++i) Push(all[i].words, -all[i].g, -all[i].a);
    return true;
  }
};

bool ReadWords(const char* path, std::vector<std::string>* words) {
  FILE* f = fopen(path, "r");
  if (!f) return false;
  char buf[1024];
  int id;
  while (fscanf(f, "%1023s %d", buf, &id) == 2) {
    if (id < 0) continue;
    if ((size_t)id >= words->size()) words->resize(id + 1);
    (*words)[id] = buf;
  }
  fclose(f);
  return true;
}

void LogSoftmaxRow(const float* x, int C, float* out) {
  float m = x[0];
  for (int c = 1; c < C; ++c) m = std::max(m, x[c]);
  double s = 0.0;
  for (int c = 0; c < C; ++c) s += std::exp((double)(x[c] - m));
  const float ls = (float)std::log(s);
  for (int c = 0; c < C; ++c) out[c] = (x[c] - m) - ls;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C API (ctypes)
extern "C" {

void* orc_create(const char* fst_path, const char* words_path, int max_active, int min_active, float beam, float lattice_beam,
                 float acoustic_scale, float blank_skip, float length_penalty, int nbest, char* err, int err_cap) {
  Facade* f = new Facade();
  std::string e;
  if (!ReadFst(fst_path, &f->graph, &e) || !ReadWords(words_path, &f->words)) {
    if (err) snprintf(err, err_cap, "%s", e.empty() ? "cannot read words file" : e.c_str());
    delete f;
    return nullptr;
  }
  f->dec.fst = &f->graph;
  f->dec.cfg.max_active = max_active; f->dec.cfg.min_active = min_active; f->dec.cfg.beam = beam;
  f->dec.cfg.lattice_beam = lattice_beam; f->dec.cfg.length_penalty = length_penalty;
  f->acoustic_scale = acoustic_scale; f->blank_skip = blank_skip; f->nbest = nbest;
  f->Reset();
  return f;
}
void orc_destroy(void* h) { delete (Facade*)h; }
void orc_reset(void* h) { ((Facade*)h)->Reset(); }
void orc_decode_logprobs(void* h, const float* logp, int T, int C) { ((Facade*)h)->Search(logp, T, C); }
// lm_decoder.cc:14-37: log_softmax, minus log-priors, blank column minus blank_penalty
void orc_decode_logits(void* h, const float* logits, const float* log_priors, int T, int C, float blank_penalty) {
  std::vector<float> lp((size_t)T * C);
  for (int t = 0; t < T; ++t) {
    LogSoftmaxRow(logits + (size_t)t * C, C, &lp[(size_t)t * C]);
    if (log_priors)
      for (int c = 0; c < C; ++c) lp[(size_t)t * C + c] -= log_priors[(size_t)t * C + c];
    lp[(size_t)t * C] -= blank_penalty;
  }
  ((Facade*)h)->Search(lp.data(), T, C);
}
void orc_finish(void* h) { ((Facade*)h)->Finish(); }
// Rescore with the LM the graph was built from (lm_fst_path) and the rescoring LM; returns 0, or -1 on error
int orc_rescore(void* h, const char* lm_fst_path, const char* rescore_lm_fst_path, char* err, int err_cap) {
  LmFst a, b;
  std::string e;
  if (!a.Read(lm_fst_path, &e) || !b.Read(rescore_lm_fst_path, &e)) {
    if (err) snprintf(err, err_cap, "%s", e.c_str());
    return -1;
  }
  return ((Facade*)h)->Rescore(&a, &b) ? 0 : -1;
}
int orc_num_results(void* h) { return (int)((Facade*)h)->results.size(); }
int orc_get_result(void* h, int i, float* ac, float* lm, char* buf, int cap) {
  Facade* f = (Facade*)h;
  if (i < 0 || i >= (int)f->results.size()) return -1;
  *ac = f->results[i].ac_score; *lm = f->results[i].lm_score;
  snprintf(buf, cap, "%s", f->results[i].sentence.c_str());
  return (int)f->results[i].sentence.size();
}
int orc_tokens_per_frame(void* h, int* out, int cap) {
  Facade* f = (Facade*)h;
  const int n = std::min(cap, (int)f->tokens_per_frame.size());
  for (int i = 0; i < n; ++i) out[i] = f->tokens_per_frame[i];
  return (int)f->tokens_per_frame.size();
}
int orc_token_list(void* h, int* states, float* costs, int cap) {      // debugging aid: the current token list, in list order
  Facade* f = (Facade*)h;
  const std::vector<Elem> l = f->dec.toks.GetList();
  for (size_t i = 0; i < l.size() && (int)i < cap; ++i) { states[i] = l[i].key; costs[i] = l[i].val->tot_cost; }
  return (int)l.size();
}
int orc_graph_info(void* h, long long* nstates, long long* narcs) {
  Facade* f = (Facade*)h;
  *nstates = (long long)f->graph.fin.size(); *narcs = (long long)f->graph.arcs.size();
  return f->graph.start;
}

void* orc_prefix_create(int blank, int first_beam, int second_beam) {
  PrefixSearch* p = new PrefixSearch();
  p->blank = blank; p->first_beam = first_beam; p->second_beam = second_beam;
  p->Reset();
  return p;
}
void orc_prefix_destroy(void* h) { delete (PrefixSearch*)h; }
void orc_prefix_reset(void* h) { ((PrefixSearch*)h)->Reset(); }
void orc_prefix_search(void* h, const float* logp, int T, int C) { ((PrefixSearch*)h)->Search(logp, T, C); }
int orc_prefix_num(void* h) { return (int)((PrefixSearch*)h)->hyps.size(); }
int orc_prefix_get(void* h, int i, int* ids, int cap, float* score, float* viterbi, int* times, int tcap) {
  PrefixSearch* p = (PrefixSearch*)h;
  if (i < 0 || i >= (int)p->hyps.size()) return -1;
  const int n = (int)p->hyps[i].size();
  for (int k = 0; k < n && k < cap; ++k) ids[k] = p->hyps[i][k];
  for (int k = 0; k < (int)p->times[i].size() && k < tcap; ++k) times[k] = p->times[i][k];
  *score = p->like[i]; *viterbi = p->viterbi[i];
  return n;
}

}  // extern "C"
