/*
 * ob_terms.hpp -- compiles a `terms` table (K x d levels, src/linalg.cpp:70-75 semantics:
 * Phi[n,k] = prod_{l: t_kl>0} B_l[n, t_kl]) into warp programs for the CUDA kernels.
 *
 * The term set chosen by outermod::selectterms (src/modandbase.cpp:387-440) is
 * downward closed, so the terms form a prefix trie: the parent of a term is the term
 * with its last non-zero dimension zeroed.  Walking that trie
 *   - top-down   gives Phi^T r with ONE multiply per (row, term)   [stream `fwd`]
 *   - bottom-up  gives Phi a   with ONE fused multiply-add per (row, term) (nested
 *     Horner form)                                                 [stream `bwd`]
 * instead of the reference's nnz_k multiplies per (row, term) (linalg.cpp:72-74).
 * Prefixes that are not themselves terms become "pass" nodes (no coefficient/output),
 * so arbitrary term tables are accepted.
 *
 * The trie is cut into segments (sub-trees) that are spread over G warps of a CTA;
 * each warp interprets its own word stream over the same row tile held in shared
 * memory.  The interpreter keeps the running product in registers (`cur`) and only
 * touches its small register stack when the host compiler says so (SAVE / *_STK /
 * *_LOAD words), so leaves -- ~75-80% of all terms -- cost one shared-memory operand
 * and one FMA.
 *
 * "Augmented" programs serve the hyper-gradient operators (prodmmge_/tprodmmge_,
 * linalg.cpp:139-163,364-386): for hyper-parameter h of dimension l the result is the
 * plain product with dimension l's factor replaced by the gradient column
 * G_h[:, t_kl] for EVERY level including 0 (rearranged).  Dimension l is
 * then made the first trie level so its level-0 column is shared.
 */
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace obt {

using u64 = uint64_t;

constexpr int kMaxDepth = 8;      /* register stack slots of the interpreter */
constexpr int kEmitBatch = 16;    /* forward emits are padded to a multiple of this per warp */
constexpr int kPrefetch = 8;      /* END words appended to every stream: the interpreters read up to seven words ahead */

/* forward (Phi^T) opcodes */
/* bit 31 of a word (op >= 8) marks a LEAF: the interpreters test it first (~75% of all words) */
enum : uint32_t { F_END = 0, F_LEAF = 8, F_DESC_CUR = 2, F_DESC_STK = 3, F_ROOT = 4, F_LOADCUR = 5, F_EMITZERO = 6 };
/* backward (Phi a) opcodes */
enum : uint32_t { B_END = 0, B_LEAF = 8, B_CLOSE_FRESH = 2, B_CLOSE_LOAD = 3, B_SAVE = 4, B_ROOT = 5 };
constexpr uint32_t FLAG_EMIT = 1, FLAG_SAVE = 2, FLAG_HAS_A = 1;
inline uint32_t mkword(uint32_t op, uint32_t depth, uint32_t flags, uint32_t col) {
  return (op << 28) | ((depth & 15u) << 24) | ((flags & 15u) << 20) | (col & 0xFFFFu);
}

/* one packed basis column the program reads */
struct ColRef {
  uint32_t dim;
  uint32_t level;  /* column knotptst[dim]+level of basemat, or gest[h]+level of the gradient matrix */
  uint32_t aug;    /* 1: comes from the gradient matrix of the augmented dimension */
};

struct Program {
  u64 K = 0, d = 0;
  int aug_dim = -1;
  int G = 0;
  std::vector<ColRef> cols;                 /* packed column table, index = `col` in the words */
  std::vector<uint32_t> fwd, bwd;           /* concatenated per-warp streams */
  std::vector<uint32_t> fwd_off, bwd_off;   /* G+1 offsets */
  std::vector<uint32_t> slot_base;          /* G+1: first slot of each warp (padded counts) */
  std::vector<uint32_t> slot_real;          /* G: real (unpadded) slots of each warp */
  std::vector<int32_t> slot_term;           /* slot -> term index, -1 for padding */
  /* brute-force form (fallback kernels, getmat): CSR of packed columns per term */
  std::vector<uint32_t> csr_ptr, csr_col;
  std::vector<u64> col_use;                 /* loads of each packed column per stream pass (cols are sorted by it) */
  u64 W = 0, nodes = 0, maxdepth = 0, Lcols = 0;
  int tmem_cap = 0;                         /* >0: words address columns >= cap as 0x8000|(col-cap) (shared memory side) */
  int fwd_stack = 1, bwd_stack = 1;         /* register-stack slots the streams touch (1 + highest index) */
  bool fast_ok = true;                      /* false: deeper than kMaxDepth or duplicate terms */
  u64 nslots() const { return slot_base.empty() ? 0 : slot_base.back(); }
};

namespace detail {
struct Node {
  int parent = -1, col = -1, term = -1, depth = 0;
  std::vector<int> kids;
};
} // namespace detail

/* terms: K x d column-major; G: warps per CTA; aug_dim: -1 or the dimension to augment */
inline uint32_t word_col(const Program& P, uint32_t w) {
  const uint32_t c = w & 0xFFFFu;
  return (P.tmem_cap > 0 && (c & 0x8000u)) ? (uint32_t)P.tmem_cap + (c & 0x7FFFu) : c;
}

inline Program compile(const u64* terms, u64 K, u64 d, int G, int aug_dim = -1, int tmem_cap = 0) {
  using detail::Node;
  Program P;
  P.K = K; P.d = d; P.G = G; P.aug_dim = aug_dim; P.tmem_cap = tmem_cap;
  /* ---- packed column table */
  std::vector<u64> lmax(d, 0);
  for (u64 l = 0; l < d; ++l) for (u64 k = 0; k < K; ++k) lmax[l] = std::max(lmax[l], terms[k + l * K]);
  std::vector<int> coloff(d, -1);
  int augoff = -1;
  if (aug_dim >= 0) { /* augmented dimension first: levels 0..lmax */
    augoff = (int)P.cols.size();
    for (u64 j = 0; j <= lmax[aug_dim]; ++j) P.cols.push_back({(uint32_t)aug_dim, (uint32_t)j, 1u});
  }
  for (u64 l = 0; l < d; ++l) {
    if ((int)l == aug_dim) continue;
    coloff[l] = (int)P.cols.size();
    for (u64 j = 1; j <= lmax[l]; ++j) P.cols.push_back({(uint32_t)l, (uint32_t)j, 0u});
  }
  for (u64 l = 0; l < d; ++l) P.Lcols += lmax[l];
  if (P.cols.size() > 0x7FFF) P.fast_ok = false;
  /* ---- per-term paths, CSR, W */
  P.csr_ptr.assign(K + 1, 0);
  std::vector<std::vector<int>> paths(K);
  for (u64 k = 0; k < K; ++k) {
    std::vector<int>& p = paths[k];
    if (aug_dim >= 0) p.push_back(augoff + (int)terms[k + (u64)aug_dim * K]);
    u64 nnz = 0;
    for (u64 l = 0; l < d; ++l) {
      const u64 t = terms[k + l * K];
      if (t > 0) { ++nnz; if ((int)l != aug_dim) p.push_back(coloff[l] + (int)t - 1); }
    }
    P.W += nnz + 1;
    P.maxdepth = std::max<u64>(P.maxdepth, p.size());
    P.csr_ptr[k + 1] = P.csr_ptr[k] + (uint32_t)p.size();
    for (int c : p) P.csr_col.push_back((uint32_t)c);
  }
  if (P.maxdepth > (u64)kMaxDepth) P.fast_ok = false;
  /* ---- trie */
  std::vector<Node> nd(1);
  std::map<std::pair<int, int>, int> index; /* (parent, col) -> node */
  for (u64 k = 0; k < K; ++k) {
    int cur = 0;
    for (int c : paths[k]) {
      auto key = std::make_pair(cur, c);
      auto it = index.find(key);
      if (it == index.end()) {
        Node n;
        n.parent = cur; n.col = c; n.depth = nd[cur].depth + 1;
        nd.push_back(n);
        const int id = (int)nd.size() - 1;
        nd[cur].kids.push_back(id);
        index[key] = id;
        cur = id;
      } else cur = it->second;
    }
    if (nd[cur].term >= 0) P.fast_ok = false; /* duplicate multi-index */
    else nd[cur].term = (int)k;
  }
  P.nodes = nd.size();
  if (!P.fast_ok) return P;
  const int NN = (int)nd.size();
  for (auto& n : nd) std::sort(n.kids.begin(), n.kids.end(), [&](int a, int b) { return nd[a].col < nd[b].col; });

  /* ---- segmentation: cut[v] starts a segment = subtree(v) minus deeper cuts */
  std::vector<char> cut(NN, 0);
  cut[0] = 1;
  std::vector<int> order; /* pre-order of the whole trie */
  {
    std::vector<int> st{0};
    while (!st.empty()) {
      int v = st.back(); st.pop_back();
      order.push_back(v);
      for (auto it = nd[v].kids.rbegin(); it != nd[v].kids.rend(); ++it) st.push_back(*it);
    }
  }
  std::vector<int> lsize(NN), segof(NN);
  auto recompute = [&]() {
    for (int i = NN - 1; i >= 0; --i) {
      const int v = order[i];
      int s = 1;
      for (int c : nd[v].kids) if (!cut[c]) s += lsize[c];
      lsize[v] = s;
    }
    for (int v : order) segof[v] = cut[v] ? v : segof[nd[v].parent];
  };
  recompute();
  const int target = std::max(24, (NN + 3 * G - 1) / (3 * G));
  for (int iter = 0; iter < 16 * G; ++iter) {
    int big = -1;
    for (int v : order) if (cut[v] && (big < 0 || lsize[v] > lsize[big])) big = v;
    if (lsize[big] <= target) break;
    int best = -1, bestscore = -1;
    const int tot = lsize[big];
    for (int v : order) {
      if (cut[v] || segof[v] != big) continue;
      const int sc = std::min(lsize[v], tot - lsize[v]);
      if (sc > bestscore) { bestscore = sc; best = v; }
    }
    if (best < 0 || bestscore <= 0) break;
    cut[best] = 1;
    recompute();
  }
  std::vector<int> segs;
  for (int v : order) if (cut[v]) segs.push_back(v);
  /* ---- LPT assignment of segments to warps */
  std::vector<int> by = segs;
  auto cost = [&](int s) { return lsize[s] + std::max(0, nd[s].depth - 1); };
  std::stable_sort(by.begin(), by.end(), [&](int a, int b) { return cost(a) > cost(b); });
  std::vector<std::vector<int>> gsegs(G);
  std::vector<long> load(G, 0);
  for (int s : by) {
    int g = 0;
    for (int i = 1; i < G; ++i) if (load[i] < load[g]) g = i;
    gsegs[g].push_back(s);
    load[g] += cost(s);
  }
  for (auto& v : gsegs) { /* the top segment (trie root) must open its warp's forward stream */
    auto it = std::find(v.begin(), v.end(), 0);
    if (it != v.end()) std::rotate(v.begin(), it, it + 1);
  }

  /* ---- emission */
  auto inseg_kids = [&](int v, std::vector<int>& leaves, std::vector<int>& inner) {
    leaves.clear(); inner.clear();
    for (int c : nd[v].kids) {
      if (cut[c]) continue;
      bool haskid = false;
      for (int cc : nd[c].kids) if (!cut[cc]) { haskid = true; break; }
      (haskid ? inner : leaves).push_back(c);
    }
  };
  P.fwd_off.assign(G + 1, 0); P.bwd_off.assign(G + 1, 0);
  P.slot_base.assign(G + 1, 0); P.slot_real.assign(G, 0);
  for (int g = 0; g < G; ++g) {
    std::vector<uint32_t> fw, bw;
    std::vector<int32_t> slots;
    /* forward */
    struct Fwd {
      const std::vector<Node>& nd; std::vector<uint32_t>& fw; std::vector<int32_t>& slots;
      decltype(inseg_kids)& kidsof;
      void node(int k, bool cur_is_parent) {
        std::vector<int> leaves, inner;
        kidsof(k, leaves, inner);
        const int e = nd[k].depth;
        if (leaves.empty() && inner.empty()) {
          if (nd[k].term < 0) return; /* pass node with every child cut away */
          if (!cur_is_parent) fw.push_back(mkword(F_LOADCUR, e - 1, 0, 0));
          fw.push_back(mkword(F_LEAF, e, FLAG_EMIT, nd[k].col));
          slots.push_back(nd[k].term);
          return;
        }
        uint32_t fl = 0;
        if (nd[k].term >= 0) { fl |= FLAG_EMIT; slots.push_back(nd[k].term); }
        if (inner.size() >= 2) fl |= FLAG_SAVE;
        fw.push_back(mkword(cur_is_parent ? F_DESC_CUR : F_DESC_STK, e, fl, nd[k].col));
        for (int c : leaves) { fw.push_back(mkword(F_LEAF, e + 1, FLAG_EMIT, nd[c].col)); slots.push_back(nd[c].term); }
        bool first = true;
        for (int c : inner) { node(c, first); first = false; }
      }
    } F{nd, fw, slots, inseg_kids};
    for (int s : gsegs[g]) {
      if (s == 0) {
        std::vector<int> leaves, inner;
        inseg_kids(0, leaves, inner);
        if (nd[0].term >= 0) { fw.push_back(mkword(F_ROOT, 0, FLAG_EMIT, 0)); slots.push_back(nd[0].term); }
        for (int c : leaves) { fw.push_back(mkword(F_LEAF, 1, FLAG_EMIT, nd[c].col)); slots.push_back(nd[c].term); }
        for (int c : inner) F.node(c, false);
      } else {
        std::vector<int> anc;
        for (int v = nd[s].parent; v > 0; v = nd[v].parent) anc.push_back(v);
        std::reverse(anc.begin(), anc.end());
        bool first = true;
        for (int v : anc) { fw.push_back(mkword(first ? F_DESC_STK : F_DESC_CUR, nd[v].depth, 0, nd[v].col)); first = false; }
        F.node(s, !anc.empty());
      }
    }
    const uint32_t nreal = (uint32_t)slots.size();
    while (slots.size() % kEmitBatch) { fw.push_back(mkword(F_EMITZERO, 0, FLAG_EMIT, 0)); slots.push_back(-1); }
    for (int i = 0; i <= kPrefetch; ++i) fw.push_back(mkword(F_END, 0, 0, 0));
    while (fw.size() % 4) fw.push_back(mkword(F_END, 0, 0, 0)); /* streams start 16-byte aligned: 4-word fetches */
    /* backward: exact reverse node order of the forward stream */
    struct Bwd {
      const std::vector<Node>& nd; std::vector<uint32_t>& bw;
      decltype(inseg_kids)& kidsof;
      int cur_scope = 0;
      bool saved[kMaxDepth + 1] = {false};
      void leaf(int k) {
        const int need = nd[k].depth - 1;
        if (cur_scope != need) { bw.push_back(mkword(B_SAVE, cur_scope, 0, 0)); saved[cur_scope] = true; cur_scope = need; }
        bw.push_back(mkword(B_LEAF, nd[k].depth, FLAG_HAS_A, nd[k].col));
      }
      void close(int k, bool has_a) {
        const int e = nd[k].depth;
        if (saved[e - 1]) { bw.push_back(mkword(B_CLOSE_LOAD, e, has_a ? FLAG_HAS_A : 0, nd[k].col)); saved[e - 1] = false; }
        else bw.push_back(mkword(B_CLOSE_FRESH, e, has_a ? FLAG_HAS_A : 0, nd[k].col));
        cur_scope = e - 1;
      }
      void node(int k) {
        std::vector<int> leaves, inner;
        kidsof(k, leaves, inner);
        if (leaves.empty() && inner.empty()) { if (nd[k].term >= 0) leaf(k); return; }
        for (auto it = inner.rbegin(); it != inner.rend(); ++it) node(*it);
        for (auto it = leaves.rbegin(); it != leaves.rend(); ++it) leaf(*it);
        close(k, nd[k].term >= 0);
      }
    } Bk{nd, bw, inseg_kids};
    for (auto it = gsegs[g].rbegin(); it != gsegs[g].rend(); ++it) {
      const int s = *it;
      if (s == 0) {
        std::vector<int> leaves, inner;
        inseg_kids(0, leaves, inner);
        for (auto i2 = inner.rbegin(); i2 != inner.rend(); ++i2) Bk.node(*i2);
        for (auto i2 = leaves.rbegin(); i2 != leaves.rend(); ++i2) Bk.leaf(*i2);
        if (nd[0].term >= 0) bw.push_back(mkword(B_ROOT, 0, FLAG_HAS_A, 0));
      } else {
        Bk.node(s);
        bool only_pass_leaf = false;
        { std::vector<int> l2, i2; inseg_kids(s, l2, i2); only_pass_leaf = l2.empty() && i2.empty() && nd[s].term < 0; }
        if (!only_pass_leaf)
          for (int v = nd[s].parent; v > 0; v = nd[v].parent) Bk.close(v, false);
      }
    }
    for (int i = 0; i <= kPrefetch; ++i) bw.push_back(mkword(B_END, 0, 0, 0));
    while (bw.size() % 4) bw.push_back(mkword(B_END, 0, 0, 0));
    P.fwd_off[g + 1] = P.fwd_off[g] + (uint32_t)fw.size();
    P.bwd_off[g + 1] = P.bwd_off[g] + (uint32_t)bw.size();
    P.fwd.insert(P.fwd.end(), fw.begin(), fw.end());
    P.bwd.insert(P.bwd.end(), bw.begin(), bw.end());
    P.slot_real[g] = nreal;
    P.slot_base[g + 1] = P.slot_base[g] + (uint32_t)slots.size();
    P.slot_term.insert(P.slot_term.end(), slots.begin(), slots.end());
  }
  { /* order the packed columns by how often the streams load them (hottest first): the v2 kernels
       keep columns [0, lt) in TMEM and the rest in shared memory */
    const size_t nc = P.cols.size();
    std::vector<u64> use(nc, 0);
    for (uint32_t w : P.fwd) { const uint32_t op = w >> 28; if (op >= 8 || op == F_DESC_CUR || op == F_DESC_STK) use[w & 0xFFFFu]++; }
    std::vector<uint32_t> order(nc), rank(nc);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return use[a] > use[b]; });
    for (uint32_t i = 0; i < nc; ++i) rank[order[i]] = i;
    std::vector<ColRef> nc_cols(nc);
    for (uint32_t i = 0; i < nc; ++i) nc_cols[i] = P.cols[order[i]];
    P.cols = nc_cols;
    P.col_use.resize(nc);
    for (uint32_t i = 0; i < nc; ++i) P.col_use[i] = use[order[i]];
    auto remap = [&](std::vector<uint32_t>& ws, bool fwd) {
      for (uint32_t& w : ws) {
        const uint32_t op = w >> 28;
        const bool hascol = fwd ? (op >= 8 || op == F_DESC_CUR || op == F_DESC_STK) : (op >= 8 || op == B_CLOSE_FRESH || op == B_CLOSE_LOAD);
        if (hascol) {
          uint32_t c = rank[w & 0xFFFFu];
          if (tmem_cap > 0 && c >= (uint32_t)tmem_cap) c = 0x8000u | (c - (uint32_t)tmem_cap);
          w = (w & 0xFFFF0000u) | c;
        }
      }
    };
    remap(P.fwd, true); remap(P.bwd, false);
    for (uint32_t& c : P.csr_col) c = rank[c];
  }
  for (uint32_t w : P.fwd) {
    const uint32_t op = w >> 28, e = (w >> 24) & 15, fl = (w >> 20) & 15;
    if (op == F_DESC_STK) P.fwd_stack = std::max(P.fwd_stack, (int)e);          /* reads e-1 */
    if ((op == F_DESC_STK || op == F_DESC_CUR) && (fl & FLAG_SAVE)) P.fwd_stack = std::max(P.fwd_stack, (int)e + 1);
    if (op == F_LOADCUR) P.fwd_stack = std::max(P.fwd_stack, (int)e + 1);
  }
  for (uint32_t w : P.bwd) {
    const uint32_t op = w >> 28, e = (w >> 24) & 15;
    if (op == B_SAVE) P.bwd_stack = std::max(P.bwd_stack, (int)e + 1);
    if (op == B_CLOSE_LOAD) P.bwd_stack = std::max(P.bwd_stack, (int)e);           /* reads e-1 */
  }
  return P;
}

/* CPU interpreters of the two streams -- used by the host unit tests of the compiler
 * (tests/test_terms_program.py through the C ABI) and as executable documentation of
 * what the CUDA interpreters in ob_kernels.cu do.  B: packed columns for ONE row. */
inline double run_bwd_row(const Program& P, const double* Bcols, const double* a) {
  double total = 0;
  for (int g = 0; g < P.G; ++g) {
    double cur = 0, stk[kMaxDepth + 1] = {0};
    int slot = (int)P.slot_base[g] + (int)P.slot_real[g] - 1;
    for (uint32_t i = P.bwd_off[g];; ++i) {
      const uint32_t w = P.bwd[i], op = w >> 28, e = (w >> 24) & 15, fl = (w >> 20) & 15, col = word_col(P, w);
      if (op == B_END) break;
      if (op == B_LEAF) cur += Bcols[col] * a[P.slot_term[slot--]];
      else if (op == B_CLOSE_FRESH || op == B_CLOSE_LOAD) {
        const double val = ((fl & FLAG_HAS_A) ? a[P.slot_term[slot--]] : 0.0) + cur;
        cur = (op == B_CLOSE_LOAD ? stk[e - 1] : 0.0) + Bcols[col] * val;
      } else if (op == B_SAVE) { stk[e] = cur; cur = 0; }
      else if (op == B_ROOT) cur += a[P.slot_term[slot--]];
    }
    total += cur;
  }
  return total;
}
inline void run_fwd_row(const Program& P, const double* Bcols, double b, double* out /* K, accumulated */) {
  for (int g = 0; g < P.G; ++g) {
    double cur = b, stk[kMaxDepth + 1] = {0};
    stk[0] = b;
    int slot = (int)P.slot_base[g];
    for (uint32_t i = P.fwd_off[g];; ++i) {
      const uint32_t w = P.fwd[i], op = w >> 28, e = (w >> 24) & 15, fl = (w >> 20) & 15, col = word_col(P, w);
      if (op == F_END) break;
      double emit = 0;
      if (op == F_LEAF) emit = cur * Bcols[col];
      else if (op == F_DESC_CUR || op == F_DESC_STK) {
        cur = (op == F_DESC_STK ? stk[e - 1] : cur) * Bcols[col];
        if (fl & FLAG_SAVE) stk[e] = cur;
        emit = cur;
      } else if (op == F_ROOT) emit = stk[0];
      else if (op == F_LOADCUR) cur = stk[e];
      if (fl & FLAG_EMIT) { const int t = P.slot_term[slot++]; if (t >= 0) out[t] += emit; }
    }
  }
}

} // namespace obt
