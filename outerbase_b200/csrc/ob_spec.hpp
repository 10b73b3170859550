/*
 * ob_spec.hpp -- generator of the terms-specialised Phi kernels (host side, plain C++).
 *
 * Partial evaluation of the two stream interpreters of ob_terms.hpp (run_bwd_row /
 * run_fwd_row, the CPU statements of what phi_a_kernel / phi_t_kernel do) on ONE compiled
 * terms table: every word of a warp's stream becomes one or two FP64 statements of
 * straight-line CUDA C, the register stack becomes named variables, and the factor
 * columns a stream reads more than once are loaded into registers once per 32R rows.
 * The result is spliced into the hand-written frame ob_spec_scaffold.inc and compiled
 * for sm_100a with NVRTC by ob_kernels.cu.  Semantics and operation order are those of
 * the interpreters, i.e. of prodmm_/tprodmm_ (src/linalg.cpp:57-131, 286-355) up to the
 * trie re-association documented in ob_terms.hpp.
 */
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <map>
#include <string>
#include <vector>

#include "ob_terms.hpp"

namespace obs {

using obt::Program;
using u64 = uint64_t;

/* mirrored by `struct SpecParams` in ob_spec_scaffold.inc */
struct SpecParams {
  const double* const* load_src;
  const int* col_op;
  const double* scale;
  const double* y;
  const double* win;
  double* out;
  double* w;
  double* ssq_partial;
  double* partial;
  const double* a;
  const double* yh;
  const int* slot_term;
  double sd;
  unsigned long long N;
  int ncol, has_ops, sq, mode, ntiles, nstage, nslots;
  unsigned off_tile, tile_doubles, off_vec, off_flags;
  /* Phi^T: cross-CTA (and cross-GPU) reduction in the tail of the same launch (fuse_tail != 0) */
  unsigned* sync_ctr;            /* monotonic arrival counter of the context (grid barrier) */
  const unsigned* ready;         /* Phi^T: rows of the input vector that have arrived so far (host -> device copy in
                                    flight on another stream), or null: the whole vector is there */
  const double* extra;           /* fused tail: element K of the result = fixed-order sum of extra[0..extra_n) */
  void* peer[8];                 /* the peer-memory slot block of every rank (Ctx::P2P), G > 1 */
  unsigned long long timeout_ns;
  unsigned sync_target, seq;
  int G, rank, fuse_tail, extra_n;
  const void* tmaps;             /* Phi a: one 2-D tensor map per run of adjacent basis columns (option tmap), or null:
                                    one bulk copy per column */
};

/* mirrored by `struct MatParams` in ob_spec_scaffold.inc (phi_am_spec) */
struct MatParams {
  const double* aperm;
  double* out;
  unsigned long long ldo;
  int ncols, nblocks;
  unsigned off_phi, off_coef;
};

/* mirrored by `struct TMatParams` in ob_spec_scaffold.inc (phi_tm_spec) */
struct TMatParams {
  const double* A;       /* N x ncols right-hand sides, column-major, leading dimension lda (padded like basemat) */
  double* partial;       /* J x nslots x 64 per-row-group sums */
  unsigned long long lda;
  int ncols, nslots;
  unsigned off_park, off_r; /* shared memory: per-warp Phi blocks | within a stage: offset of the 64 right-hand-side columns */
};

/* mirrored by `struct DotParams` in ob_spec_scaffold.inc (phi_d_spec) */
struct DotParams {
  const double* gmat;    /* basemat_gradhyp: column gest[h] + j = G_h[:, j], leading dimension ld */
  const double* bmat;    /* squared operator: the PLAIN basemat (basematsq_gradhyp = 2 G % B, modandbase.cpp:588-590), else null */
  const double* wdot;    /* row weights w (N) or null = ones */
  double* partial;       /* H x gridDim per-CTA sums */
  unsigned long long ld;
  int H, d;
  int hst[33];           /* hypst: first hyper-parameter of each dimension (d + 1) */
  int gest[65];          /* first basemat_gradhyp column of each hyper-parameter */
  int kst[33];           /* knotptst: first basemat column of each dimension */
  unsigned off_hsm;      /* shared memory: per-warp sums, warps x H doubles */
};

struct SpecOptions {
  /* Phi a (one stream = the whole program): rows per lane, passes per tile (= warps per tile), tiles
   * in work at once (compute warps = qa * tga), register-cached columns */
  /* round 2: 64-row tiles, four in work (qa 2, tga 4) -- 0.272 against 0.282 ms at 1M rows and 44 against 47 us on a
   * 125k-row shard (finer tail); all columns of C3 register-cached now that a compute warp has 232 registers */
  int ra = 1, qa = 2, tga = 4, cache_a = 80;
  /* Phi^T: warps per CTA (= streams per CTA type), rows per lane, passes per tile, cached columns,
   * accumulators per warp (cap) */
  /* acc_cap bounds the LENGTH of a stream as much as its registers: the two compute warps of an SM sub-partition run
   * different streams and their loops must fit the ~6 KB L0 instruction cache together with the producer's
   * (profiles/r02_ifetch.txt, r02_spec_sweeps.txt: at C3 56 -> 5 CTA types 0.298 ms, 72 -> 4 types 0.422 ms,
   * 92 -> 3 types 0.87 ms) */
  int wt = 8, rt = 1, pt = 4, cache_t = 16, acc_cap = 56;
  /* producer warps of both kernels (one warp issues one bulk copy per ~100 cycles, tools/tma_bench.cu) */
  int np = 4;
  /* Phi^T: 1 = form a cluster of the CTA types and multicast the tile (when 2 <= types <= 8, np >= 2).
   * Correct but measured 2-3x slower than L2-served repeats on B200 (profiles/r01_spec_sweeps.txt): off. */
  int mc = 0;
  /* Phi^T: passes of a tile interleaved by the compiler (unroll factor of the pass loop) */
  int ut = 1;
  /* multi-RHS kernel: compute warps (32 rows each) per tile and terms per block */
  int mw = 8, kc = 16;
  /* hyper-gradient sweep (phi_d_spec): warps per tile, tiles in work, register-cached columns, most tile columns
   * whose derivative accumulators still fit the register file */
  int qd = 2, tgd = 3, cache_d = 10, maxcols_d = 96, npd = 2;
  /* warp-specialised register reallocation (setmaxnreg, sm_90+): the register file is physically split over the four
   * SM sub-partitions (16384 each, warp w lives on sub-partition w % 4), so a CTA of 12 warps = 3 per sub-partition is
   * launched with 168 registers per thread; the producer warpgroup then shrinks to nreg_p and the two compute
   * warpgroups grow to nreg_c (2 * 232 + 40 = 504 <= 512 per lane of a sub-partition).  That gives EIGHT compute warps
   * -- two on every sub-partition, the FP64 pipe of each one equally loaded -- with 232 registers each, where the
   * uniform allocation allowed 6 + 2 warps at 255 (compute warps 2:2:1:1 over the sub-partitions: the busiest one
   * carries a third of the SM's work) or 8 + 2 at 168.  0 = off (uniform registers).  Needs compute and producer warp
   * counts that are multiples of 4 (setmaxnreg acts on aligned groups of four warps). */
  int nreg_c = 232, nreg_p = 40;
  /* nanoseconds a producer warp sleeps between two polls of a stage's `empty` barrier (0 = poll at full speed) */
  int psleep = 128;
  /* Phi a stages a tile with one 2-D tensor copy (cp.async.bulk.tensor, SASS UTMALDG) per dimension -- the levels
   * 1..max of a dimension are adjacent columns of the basis matrix -- instead of one bulk copy per column */
  int tmap = 0;
};

struct SpecSource {
  std::string src;
  int types = 0, nacc = 0; /* Phi^T: CTA types, accumulators declared per thread */
  int maxcols_t = 0;       /* Phi^T: most columns any type stages */
  int cluster = 1;         /* Phi^T: CTAs per cluster (= types when multicasting, else 1) */
  int tr_a = 0, tr_t = 0;  /* rows per tile */
  SpecOptions opt;
  bool ok = false;
  std::string why;
};

namespace detail {

struct Emitter {
  std::string s;
  char buf[512];
  template <class... A>
  void f(const char* fmt, A... a) { std::snprintf(buf, sizeof buf, fmt, a...); s += buf; }
  void f(const char* lit) { s += lit; }
};

/* factor access for one stream: cached columns become variables, the rest load at every use */
struct Factors {
  std::vector<int> cached; /* -1 / 1 per column */
  std::vector<char> declared;
  std::vector<int> pos;    /* program column -> column of the staged tile */
  int R, TR;
  Factors(const Program& P, const std::vector<uint32_t>& words, uint32_t lo, uint32_t hi, bool fwd, int cache, int R_, int TR_,
          const std::vector<int>* pos_ = nullptr)
      : R(R_), TR(TR_) {
    if (pos_) pos = *pos_;
    else { pos.resize(P.cols.size()); for (size_t i = 0; i < pos.size(); ++i) pos[i] = (int)i; }
    const size_t nc = P.cols.size();
    std::vector<int> use(nc, 0);
    for (uint32_t i = lo; i < hi; ++i) {
      const uint32_t w = words[i], op = w >> 28;
      const bool hascol = fwd ? (op >= 8 || op == obt::F_DESC_CUR || op == obt::F_DESC_STK)
                              : (op >= 8 || op == obt::B_CLOSE_FRESH || op == obt::B_CLOSE_LOAD);
      if (hascol) use[w & 0xFFFFu]++;
    }
    std::vector<int> order(nc);
    for (size_t i = 0; i < nc; ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return use[a] > use[b]; });
    cached.assign(nc, 0);
    declared.assign(nc, 0);
    for (int i = 0; i < (int)nc && i < cache; ++i) if (use[order[i]] >= 2) cached[order[i]] = 1;
  }
  /* expression for factor (col, r); may first emit the declaration of a cached column */
  std::string get(Emitter& e, int col, int r) {
    char b[96];
    const unsigned off = (unsigned)(pos[col] * TR + 32 * r) * 8u;
    if (!cached[col]) { std::snprintf(b, sizeof b, "ldv(tp + %uu)", off); return b; }
    if (!declared[col]) {
      declared[col] = 1;
      for (int q = 0; q < R; ++q) e.f("const double f%d_%d = lds(tp + %uu);\n", col, q, (unsigned)(pos[col] * TR + 32 * q) * 8u);
    }
    std::snprintf(b, sizeof b, "f%d_%d", col, r);
    return b;
  }
};

/* backward (Horner) stream g -> statements accumulating into o[r] */
inline void emit_bwd(Emitter& e, const Program& P, int g, int R, int TR, int cache, const std::vector<int>* pos = nullptr) {
  using namespace obt;
  Factors F(P, P.bwd, P.bwd_off[g], P.bwd_off[g + 1], false, cache, R, TR, pos);
  int slot = (int)P.slot_base[g] + (int)P.slot_real[g] - 1;
  int nv = 0;
  /* symbolic values: "" = known zero, else a variable stem (stem_r) */
  std::string cur, stk[kMaxDepth + 2];
  auto fresh = [&]() { char b[24]; std::snprintf(b, sizeof b, "x%d", nv++); return std::string(b); };
  /* coefficient of slot k -> variable a<k>.  Slots are consumed downwards: an odd slot fetches
   * the aligned pair (k-1, k) with one broadcast LDS.128, the even partner is used next. */
  const int slot_lo = (int)P.slot_base[g];
  std::vector<char> have(P.nslots() + 2, 0);
  auto coef = [&](int k) {
    if (have[k]) return;
    if ((k & 1) && k - 1 >= slot_lo) {
      e.f("double a%d, a%d; lda2<%d>(as, a%d, a%d);\n", k - 1, k, 8 * (k - 1), k - 1, k);
      have[k - 1] = have[k] = 1;
    } else {
      e.f("const double a%d = lda<%d>(as);\n", k, 8 * k);
      have[k] = 1;
    }
  };
  for (uint32_t i = P.bwd_off[g];; ++i) {
    const uint32_t w = P.bwd[i], op = w >> 28, d = (w >> 24) & 15, fl = (w >> 20) & 15, col = w & 0xFFFFu;
    if (op == B_END) break;
    if (op == B_LEAF) {
      const int k = slot--;
      const std::string x = fresh();
      coef(k);
      for (int r = 0; r < R; ++r) {
        const std::string f = F.get(e, (int)col, r);
        if (cur.empty()) e.f("const double %s_%d = %s * a%d;\n", x.c_str(), r, f.c_str(), k);
        else e.f("const double %s_%d = fma(%s, a%d, %s_%d);\n", x.c_str(), r, f.c_str(), k, cur.c_str(), r);
      }
      cur = x;
    } else if (op == B_CLOSE_FRESH || op == B_CLOSE_LOAD) {
      const bool has_a = fl & FLAG_HAS_A;
      const int k = has_a ? slot-- : -1;
      const std::string base = (op == B_CLOSE_LOAD) ? stk[d - 1] : std::string();
      const std::string x = fresh();
      if (has_a) coef(k);
      for (int r = 0; r < R; ++r) {
        const std::string f = F.get(e, (int)col, r);
        char val[96];
        if (has_a && !cur.empty()) std::snprintf(val, sizeof val, "(a%d + %s_%d)", k, cur.c_str(), r);
        else if (has_a) std::snprintf(val, sizeof val, "a%d", k);
        else if (!cur.empty()) std::snprintf(val, sizeof val, "%s_%d", cur.c_str(), r);
        else std::snprintf(val, sizeof val, "0.0");
        if (base.empty()) e.f("const double %s_%d = %s * %s;\n", x.c_str(), r, f.c_str(), val);
        else e.f("const double %s_%d = fma(%s, %s, %s_%d);\n", x.c_str(), r, f.c_str(), val, base.c_str(), r);
      }
      if (op == B_CLOSE_LOAD) stk[d - 1].clear();
      cur = x;
    } else if (op == B_SAVE) {
      stk[d] = cur;
      cur.clear();
    } else if (op == B_ROOT) {
      const int k = slot--;
      const std::string x = fresh();
      coef(k);
      for (int r = 0; r < R; ++r) {
        if (cur.empty()) e.f("const double %s_%d = a%d;\n", x.c_str(), r, k);
        else e.f("const double %s_%d = %s_%d + a%d;\n", x.c_str(), r, cur.c_str(), r, k);
      }
      cur = x;
    }
  }
  if (!cur.empty()) for (int r = 0; r < R; ++r) e.f("o[%d] = %s_%d;\n", r, cur.c_str(), r);
}

/* forward (top-down) stream g -> statements accumulating into acc<i>; returns the emit count */
inline int emit_fwd(Emitter& e, const Program& P, int g, int R, int TR, int cache, const std::vector<int>& pos, bool park = false) {
  using namespace obt;
  /* park: every emit is Phi[row, term] itself and goes to the warp's shared-memory block (OBS_TM_PARK) instead of a
   * register accumulator -- phi_tm_spec contracts the block with the right-hand sides on the tensor cores */
  Factors F(P, P.fwd, P.fwd_off[g], P.fwd_off[g + 1], true, cache, R, TR, &pos);
  int nv = 0, ia = 0;
  std::string cur = "b", stk[kMaxDepth + 2];
  stk[0] = "b";
  auto name = [&](const std::string& stem, int r) {
    char b[48];
    if (stem == "b") std::snprintf(b, sizeof b, "b[%d]", r);
    else std::snprintf(b, sizeof b, "%s_%d", stem.c_str(), r);
    return std::string(b);
  };
  auto fresh = [&]() { char b[24]; std::snprintf(b, sizeof b, "v%d", nv++); return std::string(b); };
  for (uint32_t i = P.fwd_off[g];; ++i) {
    const uint32_t w = P.fwd[i], op = w >> 28, d = (w >> 24) & 15, fl = (w >> 20) & 15, col = w & 0xFFFFu;
    if (op == F_END) break;
    if (op == F_LEAF) {
      for (int r = 0; r < R; ++r) {
        const std::string f = F.get(e, (int)col, r);
        if (park) e.f("OBS_TM_PARK(%d, %s * %s);\n", ia, name(cur, r).c_str(), f.c_str());
        else e.f("acc%d = fma(%s, %s, acc%d);\n", ia, name(cur, r).c_str(), f.c_str(), ia);
      }
      ++ia;
    } else if (op == F_DESC_CUR || op == F_DESC_STK) {
      const std::string src = (op == F_DESC_STK) ? stk[d - 1] : cur;
      const std::string x = fresh();
      for (int r = 0; r < R; ++r) {
        const std::string f = F.get(e, (int)col, r);
        e.f("const double %s_%d = %s * %s;\n", x.c_str(), r, name(src, r).c_str(), f.c_str());
      }
      cur = x;
      if (fl & FLAG_SAVE) stk[d] = x;
      if (fl & FLAG_EMIT) {
        for (int r = 0; r < R; ++r) {
          if (park) e.f("OBS_TM_PARK(%d, %s_%d);\n", ia, x.c_str(), r);
          else e.f("acc%d += %s_%d;\n", ia, x.c_str(), r);
        }
        ++ia;
      }
    } else if (op == F_ROOT) {
      for (int r = 0; r < R; ++r) {
        if (park) e.f("OBS_TM_PARK(%d, b[%d]);\n", ia, r);
        else e.f("acc%d += b[%d];\n", ia, r);
      }
      ++ia;
    } else if (op == F_LOADCUR) {
      cur = stk[d];
    } /* F_EMITZERO: padding of the interpreter's emit batches, nothing to do */
  }
  return ia;
}

/* forward stream 0 of a G = 1 program -> a loop over blocks of 32 emits: `switch (blk)` selects the statements that
 * park Phi[row, term] of the block's emits in the warp's shared-memory block (OBS_M_EMIT), then ONE shared instance
 * of the tensor-core contraction runs (OBS_M_BLOCK).  The interpreter's running product and register stack are
 * plain variables (cu, s0..s8) that live across the cases.  Returns the emit count. */
inline int emit_mat(Emitter& e, const Program& P, int TR, int KC) {
  using namespace obt;
  int ia = 0;
  bool open = false;
  /* One case = the statements of KC emits.  Its basis-column loads are hoisted to the top of the case (each column
   * once): the loads, the products and the parking stores are volatile asm statements that keep their order, and a
   * load right in front of its use made every emit wait a full shared-memory latency (~40 cycles x KC per block, a
   * quarter of the block's tensor-core time, profiles/r02_traffic.json c5: DMMA pipe 75 % active). */
  Emitter loads, stmts;
  std::map<uint32_t, int> have; /* column -> f index in this case */
  auto begin_case = [&]() { if (!open) { loads.s.clear(); stmts.s.clear(); have.clear(); open = true; } };
  auto end_case = [&]() {
    e.f("case %d: {\n", (ia - 1) / KC);
    e.s += loads.s;
    e.s += stmts.s;
    e.f("} break;\n");
    open = false;
  };
  auto emitted = [&]() { if (++ia % KC == 0) end_case(); };
  auto fac = [&](uint32_t col) {
    auto it = have.find(col);
    if (it == have.end()) {
      it = have.emplace(col, (int)have.size()).first;
      loads.f("const double f%d = ldv(tp + %uu);\n", it->second, (unsigned)(col * TR) * 8u);
    }
    char b[24];
    std::snprintf(b, sizeof b, "f%d", it->second);
    return std::string(b);
  };
  e.f("double cu = 1.0, s0 = 1.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0, s6 = 0.0, s7 = 0.0, s8 = 0.0;\n");
  e.f("_Pragma(\"unroll 1\") for (int blk = 0; blk < OBS_NBLK; ++blk) {\nswitch (blk) {\n");
  for (uint32_t i = P.fwd_off[0];; ++i) {
    const uint32_t w = P.fwd[i], op = w >> 28, d = (w >> 24) & 15, fl = (w >> 20) & 15, col = w & 0xFFFFu;
    if (op == F_END) break;
    if (op == F_EMITZERO) continue;
    begin_case();
    if (op == F_LEAF) {
      stmts.f("OBS_M_EMIT(%d, cu * %s);\n", ia % KC, fac(col).c_str());
      emitted();
    } else if (op == F_DESC_CUR || op == F_DESC_STK) {
      if (op == F_DESC_STK) stmts.f("cu = s%u * %s;\n", d - 1, fac(col).c_str());
      else stmts.f("cu = cu * %s;\n", fac(col).c_str());
      if (fl & FLAG_SAVE) stmts.f("s%u = cu;\n", d);
      if (fl & FLAG_EMIT) { stmts.f("OBS_M_EMIT(%d, cu);\n", ia % KC); emitted(); }
    } else if (op == F_ROOT) {
      stmts.f("OBS_M_EMIT(%d, s0);\n", ia % KC);
      emitted();
    } else if (op == F_LOADCUR) {
      stmts.f("cu = s%u;\n", d);
    }
  }
  const int n = ia;
  if (ia % KC) { begin_case(); while (ia % KC) { stmts.f("OBS_M_EMIT(%d, 0.0);\n", ia % KC); ++ia; } end_case(); }
  if (open) { /* trailing statements without an emit (cannot follow the last term of a well-formed program) */
    e.f("case %d: {\n", ia / KC);
    e.s += loads.s; e.s += stmts.s;
    e.f("} break;\n");
    open = false;
  }
  e.f("default: break;\n}\nOBS_M_BLOCK()\n}\n");
  return n;
}

/* ---- reverse-mode sweep over the trie (phi_d_spec): for one row, yhat/basescale = sum_k a_k prod B and
 *   D_c = d(yhat/basescale) / d B_c   for EVERY tile column c = (dimension, level),
 * from one depth-first walk: descending an edge (parent -> child, column c) forms the prefix product
 * v_child = v_parent * B_c, returning from it brings the child's Horner sum u_child = a_child + sum B u, and the edge
 * contributes  u_parent += B_c u_child  and  D_c += v_parent u_child.  3 FMAs per inner edge, 2 per leaf, against the
 * H separate plain products of domultgesub_ (src/linalg.cpp:139-163) this replaces.  The epilogue contracts D with the
 * stored gradient columns per dimension (linalg.cpp:139-163, 273-276 regrouped):
 *   outge[n,h] = basescale * ( sum_j D_(l,j) G_h[n,j] + G_h[n,0] (u_root - sum_j D_(l,j) B_(l,j)) ),   l = hypmatch[h]. */
struct DNode { int col = -1, term = -1; std::vector<int> kids; };

inline int emit_dot(Emitter& e, Emitter& tab, const Program& P, int TR, int cache) {
  const size_t nc = P.cols.size();
  /* trie of the terms' column paths (csr rows are ascending in dimension) */
  std::vector<DNode> N(1);
  for (u64 k = 0; k < P.K; ++k) {
    int at = 0;
    for (uint32_t i = P.csr_ptr[k]; i < P.csr_ptr[k + 1]; ++i) {
      const int c = (int)P.csr_col[i];
      int nx = -1;
      for (int kid : N[at].kids) if (N[kid].col == c) { nx = kid; break; }
      if (nx < 0) { nx = (int)N.size(); DNode nd; nd.col = c; N.push_back(nd); N[at].kids.push_back(nx); }
      at = nx;
    }
    N[at].term = (int)k;
  }
  std::vector<int> use(nc, 1); /* the epilogue reads every column once */
  for (size_t i = 1; i < N.size(); ++i) use[N[i].col]++;
  std::vector<int> order(nc);
  for (size_t i = 0; i < nc; ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return use[a] > use[b]; });
  std::vector<char> cached(nc, 0);
  for (int i = 0; i < (int)nc && i < cache; ++i) cached[order[i]] = 1;
  auto fac = [&](int c) {
    char b[64];
    if (cached[c]) std::snprintf(b, sizeof b, "f%d", c);
    else std::snprintf(b, sizeof b, "ldv(tp + %uu)", (unsigned)(c * TR) * 8u);
    return std::string(b);
  };
  for (size_t c = 0; c < nc; ++c) if (cached[c]) e.f("const double f%d = lds(tp + %uu);\n", (int)c, (unsigned)(c * TR) * 8u);
  for (size_t c = 0; c < nc; ++c) e.f("double D%d = 0.0;\n", (int)c);
  /* coefficients in consumption order: slot i of the shared-memory copy holds a[dterm[i]] */
  std::vector<int> dterm;
  { /* pre-pass: the order in which the walk below consumes coefficients */
    std::vector<std::pair<int, size_t>> st{{0, 0}};
    if (N[0].term >= 0) dterm.push_back(N[0].term);
    while (!st.empty()) {
      auto& [n, i] = st.back();
      if (i == N[n].kids.size()) { st.pop_back(); continue; }
      const int kid = N[n].kids[i++];
      if (N[kid].term >= 0) dterm.push_back(N[kid].term);
      if (!N[kid].kids.empty()) st.push_back({kid, 0});
    }
  }
  int slot = 0, nv = 0, blk = 1;
  std::vector<int> have(dterm.size() + 2, 0); /* block in which the coefficient's variable was declared */
  auto coef = [&]() { /* next coefficient -> variable a<slot>; even slots fetch the aligned pair with one LDS.128 */
    const int k = slot++;
    if (have[k] != blk) {
      if (!(k & 1) && k + 1 < (int)dterm.size()) { e.f("double a%d, a%d; lda2<%d>(as, a%d, a%d);\n", k, k + 1, 8 * k, k, k + 1); have[k] = have[k + 1] = blk; }
      else { e.f("const double a%d = lda<%d>(as);\n", k, 8 * k); have[k] = blk; }
    }
    char b[24]; std::snprintf(b, sizeof b, "a%d", k); return std::string(b);
  };
  auto fresh = [&](const char* stem) { char b[24]; std::snprintf(b, sizeof b, "%s%d", stem, nv++); return std::string(b); };
  /* returns the expression of the node's Horner sum u ("" = zero); v = prefix product at the node (never the root) */
  std::function<std::string(int, const std::string&)> walk = [&](int n, const std::string& v) -> std::string {
    std::string u;
    if (N[n].term >= 0) u = coef();
    for (int kid : N[n].kids) {
      const int c = N[kid].col;
      std::string uc;
      if (N[kid].kids.empty()) uc = coef(); /* leaf: its sum is its coefficient (a trie leaf is always a term) */
      else {
        const std::string vc = fresh("v");
        e.f("const double %s = %s * %s;\n", vc.c_str(), v.c_str(), fac(c).c_str());
        uc = walk(kid, vc);
      }
      const std::string x = fresh("u");
      if (u.empty()) e.f("const double %s = %s * %s;\n", x.c_str(), fac(c).c_str(), uc.c_str());
      else e.f("const double %s = fma(%s, %s, %s);\n", x.c_str(), fac(c).c_str(), uc.c_str(), u.c_str());
      u = x;
      e.f("D%d = fma(%s, %s, D%d);\n", c, v.c_str(), uc.c_str(), c);
    }
    return u;
  };
  /* The root's children are emitted as separate basic blocks behind an always-true, compiler-opaque branch
   * (OBS_D_BLOCK): NVVM's time on ONE straight-line block of 10^4 statements is superlinear -- 28 s at C4's table
   * against 6.5 s split.  Only ur, the D accumulators and the cached columns live across blocks; small sub-tries share
   * a block.  Coefficient variables are block-local, so a pair loaded in one block is reloaded in the next. */
  if (N[0].term >= 0) e.f("double ur = %s;\n", coef().c_str());
  else e.f("double ur = 0.0;\n");
  size_t mark = 0;
  bool open = false;
  auto lines_since = [&]() { size_t n = 0; for (size_t i = mark; i < e.s.size(); ++i) n += e.s[i] == '\n'; return n; };
  for (int kid : N[0].kids) {
    if (open && lines_since() > 300) { e.f("}\n"); open = false; }
    if (!open) { ++blk; e.f("OBS_D_BLOCK {\n"); mark = e.s.size(); open = true; }
    const int c = N[kid].col;
    std::string uc;
    if (N[kid].kids.empty()) uc = coef();
    else {
      std::string vc = fac(c);
      if (!cached[c]) { const std::string t = fresh("v"); e.f("const double %s = %s;\n", t.c_str(), vc.c_str()); vc = t; }
      uc = walk(kid, vc);
    }
    e.f("ur = fma(%s, %s, ur);\n", fac(c).c_str(), uc.c_str());
    e.f("D%d += %s;\n", c, uc.c_str());
  }
  if (open) e.f("}\n");
  e.f("const double uroot = ur;\n");
  if (slot != (int)dterm.size()) throw std::logic_error("phi_d_spec: coefficient order mismatch");
  /* epilogue: one block per dimension; the hyper-parameters of a dimension are a run-time loop (their number
   * belongs to the covariance functions, not to the terms table) */
  for (u64 l = 0; l < P.d; ++l) {
    std::vector<int> cl;
    for (size_t c = 0; c < nc; ++c) if (P.cols[c].dim == l) cl.push_back((int)c);
    if (cl.empty()) { e.f("OBS_D_DIM_EMPTY(%d)\n", (int)l); continue; }
    e.f("{ double E = 0.0;\n");
    for (int c : cl) e.f("E = fma(D%d, %s, E);\n", c, fac(c).c_str());
    e.f("OBS_D_HYP_BEGIN(%d)\n", (int)l);
    for (int c : cl) e.f("OBS_D_ACC(D%d, %d, %u)\n", c, (int)l, P.cols[c].level);
    e.f("OBS_D_HYP_END(%d)\n}\n", (int)l);
  }
  tab.f("__device__ const int obs_dterm[] = {");
  for (int t : dterm) tab.f("%d,", t);
  tab.f("0};\n");
  for (size_t i = 0; i < dterm.size(); ++i) tab.f("// OBS_SLOT_D %d %d\n", (int)i, dterm[i]);
  return (int)dterm.size();
}

inline void replace_marker(std::string& s, const char* marker, const std::string& with) {
  const size_t at = s.find(marker);
  if (at == std::string::npos) throw std::logic_error(std::string("scaffold marker missing: ") + marker);
  s.replace(at, std::strlen(marker), with);
}

} // namespace detail

inline const char* scaffold_text() {
  static const char* text =
#include "ob_spec_scaffold.inc"
      ;
  return text;
}

/* Phi^T: smallest number of CTA types (streams = types * wt) whose busiest stream keeps its
 * accumulators within the cap.  0: the table is not trie-compilable, or the cap is below what one
 * trie segment needs (ob_terms.hpp cuts segments of >= 24 nodes). */
inline int choose_types(const u64* terms, u64 K, u64 d, const SpecOptions& opt) {
  if (K == 0) return 0;
  int types = (int)std::max<u64>(1, (K + (u64)opt.wt * opt.acc_cap - 1) / ((u64)opt.wt * opt.acc_cap));
  for (; types <= 64; ++types) {
    const Program P = obt::compile(terms, K, d, types * opt.wt);
    if (!P.fast_ok) return 0;
    u64 mx = 0;
    for (uint32_t r : P.slot_real) mx = std::max<u64>(mx, r);
    if ((int)mx <= opt.acc_cap) return types;
  }
  return 0;
}

/* Phi a's tile layout: program columns sorted by (dimension, level); a run = consecutive levels of one dimension */
inline std::vector<int> tile_order(const Program& P) {
  std::vector<int> order(P.cols.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    const obt::ColRef &x = P.cols[a], &y = P.cols[b];
    if (x.aug != y.aug) return x.aug < y.aug;
    return x.dim != y.dim ? x.dim < y.dim : x.level < y.level;
  });
  return order;
}
/* first tile column of every run */
inline std::vector<int> tile_runs(const Program& P, const std::vector<int>& order) {
  std::vector<int> runs;
  for (size_t c = 0; c < order.size(); ++c) {
    const obt::ColRef& x = P.cols[order[c]];
    const bool cont = c > 0 && P.cols[order[c - 1]].aug == x.aug && P.cols[order[c - 1]].dim == x.dim && P.cols[order[c - 1]].level + 1 == x.level;
    if (!cont) runs.push_back((int)c);
  }
  return runs;
}

/* pa: program compiled with G = 1 (or null: no Phi a kernel); pt: G = types * opt.wt (or null) */
inline SpecSource generate(const Program* pa, const Program* pt, int types, const SpecOptions& opt) {
  using namespace detail;
  SpecSource S;
  S.opt = opt;
  S.tr_a = 32 * opt.ra * opt.qa;
  S.tr_t = 32 * opt.rt * opt.pt;
  const bool want_a = pa != nullptr, want_t = pt != nullptr;
  if (!want_a && !want_t) { S.why = "nothing to generate"; return S; }
  const u64 K = want_a ? pa->K : pt->K;
  if (K == 0) { S.why = "no terms"; return S; }
  if (256 % S.tr_a || 256 % S.tr_t || opt.np < 1 || opt.np > 8 || opt.qa * opt.tga + opt.np > 32 || opt.wt + opt.np > 32 || opt.qa < 1 || opt.tga < 1 || opt.tga > 8 || opt.wt > 31) {
    S.why = "inconsistent tile options";
    return S;
  }
  if ((want_a && (!pa->fast_ok || pa->G != 1 || pa->tmem_cap)) || (want_t && (!pt->fast_ok || pt->G != types * opt.wt || pt->tmem_cap))) {
    S.why = "programs do not match the options";
    return S;
  }
  /* register reallocation needs warp counts that are multiples of 4 and register counts that are multiples of 8:
   * geometries that do not qualify run with the uniform allocation */
  int nreg_c = opt.nreg_c;
  if (nreg_c && ((opt.qa * opt.tga) % 4 || opt.wt % 4 || opt.np % 4 || nreg_c % 8 || opt.nreg_p % 8 || opt.nreg_p < 24 || nreg_c > 256)) nreg_c = 0;
  Emitter hdr, tab, ca, ct, decl, red;
  hdr.f("#define OBS_PARAM_BYTES %d\n#define OBS_K %llu\n#define OBS_NP %d\n#define OBS_NREG_C %d\n#define OBS_NREG_P %d\n#define OBS_PSLEEP %d\n", (int)sizeof(SpecParams),
        (unsigned long long)K, opt.np, nreg_c, opt.nreg_p, opt.psleep);
  if (want_a) {
    const Program& P = *pa;
    hdr.f("#define OBS_HAVE_A 1\n#define OBS_RA %d\n#define OBS_QA %d\n#define OBS_TGA %d\n#define OBS_NCOLS_A %d\n", opt.ra, opt.qa,
          opt.tga, (int)P.cols.size());
    /* tile columns in (dimension, level) order: the levels of one dimension are adjacent columns of the basis matrix,
     * so a run of them is ONE 2-D tensor copy (tile_runs) */
    std::vector<int> order = tile_order(P);
    if (!opt.tmap) for (size_t c = 0; c < order.size(); ++c) order[c] = (int)c; /* bulk copies: the program's own (hottest first) order */
    std::vector<int> pos(P.cols.size());
    for (size_t c = 0; c < order.size(); ++c) pos[order[c]] = (int)c;
    tab.f("__device__ const unsigned short obs_cols_a[] = {");
    for (size_t c = 0; c < order.size(); ++c) tab.f("%d,", order[c]);
    tab.f("0};\n");
    const std::vector<int> runs = tile_runs(P, order);
    hdr.f("#define OBS_TMAP_A %d\n#define OBS_NRUNS_A %d\n", (opt.tmap && runs.size() <= 31) ? 1 : 0, (int)runs.size());
    tab.f("__device__ const unsigned short obs_run_col_a[] = {");
    for (int r : runs) tab.f("%d,", r);
    tab.f("%d};\n", (int)order.size());
    /* machine-readable layout (comments): tile column -> (dimension, level), coefficient slot -> term */
    for (size_t c = 0; c < order.size(); ++c) tab.f("// OBS_LAYOUT_A %d %u %u\n", (int)c, P.cols[order[c]].dim, P.cols[order[c]].level);
    for (u64 i = 0; i < P.nslots(); ++i) tab.f("// OBS_SLOT_A %d %d\n", (int)i, (int)P.slot_term[i]);
    ca.f("/*BEGIN_BODY_A*/\n");
    emit_bwd(ca, P, 0, opt.ra, S.tr_a, opt.cache_a, &pos);
    ca.f("/*END_BODY_A*/\n");
  }
  if (want_t) {
    const Program& P = *pt;
    u64 mx = 0;
    for (uint32_t r : P.slot_real) mx = std::max<u64>(mx, r);
    S.nacc = (int)mx;
    S.types = types;
    const int G = types * opt.wt;
    hdr.f("#define OBS_HAVE_T 1\n#define OBS_WT %d\n#define OBS_RT %d\n#define OBS_PT %d\n#define OBS_TYPES %d\n", opt.wt, opt.rt,
          opt.pt, types);
    /* per-type column lists */
    std::vector<std::vector<int>> tcols(types);
    for (int g = 0; g < G; ++g) {
      std::vector<char> used(P.cols.size(), 0);
      for (uint32_t i = P.fwd_off[g]; i < P.fwd_off[g + 1]; ++i) {
        const uint32_t w = P.fwd[i], op = w >> 28;
        if (op >= 8 || op == obt::F_DESC_CUR || op == obt::F_DESC_STK) used[w & 0xFFFFu] = 1;
      }
      auto& tc = tcols[g / opt.wt];
      for (size_t c = 0; c < used.size(); ++c) if (used[c] && std::find(tc.begin(), tc.end(), (int)c) == tc.end()) tc.push_back((int)c);
    }
    S.cluster = (opt.mc && types >= 2 && types <= 8 && opt.np >= 2) ? types : 1;
    size_t maxcols = 1;
    for (auto& tc : tcols) maxcols = std::max(maxcols, tc.size());
    if (S.cluster > 1) maxcols = P.cols.size(); /* multicast: every CTA receives every column */
    hdr.f("#define OBS_CL %d\n#define OBS_UT %d\n", S.cluster, std::max(1, opt.ut));
    tab.f("__device__ const unsigned short obs_cols_all[] = {");
    for (size_t c = 0; c < P.cols.size(); ++c) tab.f("%d,", (int)c);
    tab.f("0};\n");
    hdr.f("#define OBS_MAXCOLS_T %d\n", (int)maxcols);
    tab.f("__device__ const unsigned short obs_cols_t[] = {");
    std::vector<int> coff{0};
    for (auto& tc : tcols) { std::sort(tc.begin(), tc.end()); for (int c : tc) tab.f("%d,", c); coff.push_back(coff.back() + (int)tc.size()); }
    tab.f("0};\n__device__ const unsigned short obs_coff_t[] = {");
    for (int v : coff) tab.f("%d,", v);
    tab.f("};\n__device__ const unsigned obs_slot_base[] = {");
    for (int g = 0; g < G; ++g) tab.f("%u,", P.slot_base[g]);
    tab.f("};\n__device__ const unsigned short obs_slot_real[] = {");
    for (int g = 0; g < G; ++g) tab.f("%u,", P.slot_real[g]);
    tab.f("};\n");
    S.maxcols_t = (int)maxcols;
    for (int t = 0; t < types; ++t) /* tile column of every type -> (dimension, level) */
      for (size_t i = 0; i < (S.cluster > 1 ? P.cols.size() : tcols[t].size()); ++i) {
        const int c = S.cluster > 1 ? (int)i : tcols[t][i];
        tab.f("// OBS_LAYOUT_T %d %d %u %u\n", t, (int)i, P.cols[c].dim, P.cols[c].level);
      }
    for (int g = 0; g < G; ++g) {
      const auto& tc = tcols[g / opt.wt]; /* sorted above */
      std::vector<int> pos(P.cols.size(), -1);
      for (size_t i = 0; i < tc.size(); ++i) pos[tc[i]] = (int)i;
      if (S.cluster > 1) for (size_t c = 0; c < pos.size(); ++c) pos[c] = (int)c;
      tab.f("// OBS_STREAM_T %d type %d terms", g, g / opt.wt);
      for (uint32_t i = 0; i < P.slot_real[g]; ++i) tab.f(" %d", (int)P.slot_term[P.slot_base[g] + i]);
      tab.f("\n");
      ct.f("case %d: {\nOBS_T_LOOP_BEGIN\n/*BEGIN_CASE_T %d*/\n", g, g);
      const int n = emit_fwd(ct, P, g, opt.rt, S.tr_t, opt.cache_t, pos);
      if (n != (int)P.slot_real[g]) { S.why = "emit count mismatch"; return S; }
      ct.f("/*END_CASE_T*/\nOBS_T_LOOP_END\n} break;\n");
    }
    for (int i = 0; i < S.nacc; ++i) decl.f("  double acc%d = 0.0;\n", i);
    /* one butterfly per accumulator, once per launch; lane 0 stores */
    for (int i = 0; i < S.nacc; ++i) {
      red.f("    { double v = acc%d;", i);
      for (int m = 16; m > 0; m >>= 1) red.f(" v += shfl_xor_d(v, %d);", m);
      red.f(" if (lane == 0 && %d < nacc) dst[%d] = v; }\n", i, i);
    }
  }
  std::string body = scaffold_text();
  replace_marker(body, "//@@TABLES@@", tab.s);
  if (want_a) replace_marker(body, "//@@BODY_A@@", ca.s);
  if (want_t) {
    replace_marker(body, "//@@ACC_DECL@@", decl.s);
    replace_marker(body, "//@@CASES_T@@", ct.s);
    replace_marker(body, "//@@ACC_REDUCE@@", red.s);
  }
  S.src = hdr.s + body;
  S.ok = true;
  return S;
}

/* the multi right-hand-side module (phi_am_spec only): pa = program compiled with G = 1 */
inline SpecSource generate_mat(const Program& pa, const SpecOptions& opt) {
  using namespace detail;
  SpecSource S;
  S.opt = opt;
  const int MW = opt.mw, KC = opt.kc;
  S.tr_a = 32 * MW;
  if (!pa.fast_ok || pa.G != 1 || pa.tmem_cap || pa.K == 0) { S.why = "program does not match"; return S; }
  if ((MW != 4 && MW != 8) || (KC != 16 && KC != 32)) { S.why = "unsupported multi-RHS geometry"; return S; }
  Emitter hdr, tab, body;
  hdr.f("#define OBS_PARAM_BYTES %d\n#define OBS_K %llu\n#define OBS_NP 1\n#define OBS_HAVE_M 1\n#define OBS_NCOLS_A %d\n", (int)sizeof(SpecParams),
        (unsigned long long)pa.K, (int)pa.cols.size());
  tab.f("__device__ const unsigned short obs_cols_a[] = {");
  for (size_t c = 0; c < pa.cols.size(); ++c) tab.f("%d,", (int)c);
  tab.f("0};\n");
  /* machine-readable layout (comments): tile column -> (dimension, level), emit slot -> term */
  for (size_t c = 0; c < pa.cols.size(); ++c) tab.f("// OBS_LAYOUT_M %d %u %u\n", (int)c, pa.cols[c].dim, pa.cols[c].level);
  for (u64 i = 0; i < pa.nslots(); ++i) tab.f("// OBS_SLOT_M %d %d\n", (int)i, (int)pa.slot_term[i]);
  body.f("/*BEGIN_BODY_M*/\n");
  const int n = emit_mat(body, pa, S.tr_a, KC);
  body.f("/*END_BODY_M*/\n");
  if (n != (int)pa.slot_real[0]) { S.why = "emit count mismatch"; return S; }
  S.nacc = (n + KC - 1) / KC; /* coefficient blocks per pass */
  hdr.f("#define OBS_NBLK %d\n#define OBS_KC %d\n#define OBS_MW %d\n", S.nacc, KC, MW);
  std::string src = scaffold_text();
  replace_marker(src, "//@@TABLES@@", tab.s);
  replace_marker(src, "//@@BODY_M@@", body.s);
  S.src = hdr.s + src;
  S.ok = true;
  return S;
}

/* Phi^T . A on the tensor cores (phi_tm_spec only): pt = program compiled with G = types * 8 streams of at most
 * kTmTerms terms each (choose_types with that cap) */
constexpr int kTmTerms = 32, kTmWarps = 8;
inline SpecSource generate_tmat(const Program& pt, int types, const SpecOptions& opt) {
  using namespace detail;
  SpecSource S;
  S.opt = opt;
  S.types = types;
  S.tr_t = 64;
  const int G = types * kTmWarps;
  if (!pt.fast_ok || pt.G != G || pt.tmem_cap || pt.K == 0) { S.why = "program does not match"; return S; }
  Emitter hdr, tab, ct;
  hdr.f("#define OBS_PARAM_BYTES %d\n#define OBS_TM_BYTES %d\n#define OBS_K %llu\n#define OBS_NP 4\n#define OBS_HAVE_TM 1\n#define OBS_TYPES %d\n#define OBS_PSLEEP %d\n", (int)sizeof(SpecParams),
        (int)sizeof(TMatParams), (unsigned long long)pt.K, types, opt.psleep);
  hdr.f("#define OBS_NREG_C %d\n#define OBS_NREG_P %d\n", opt.nreg_c % 8 == 0 && opt.nreg_c >= 200 ? opt.nreg_c : 232, 40);
  std::vector<std::vector<int>> tcols(types);
  for (int g = 0; g < G; ++g) {
    if ((int)pt.slot_real[g] > kTmTerms) { S.why = "stream longer than the Phi block"; return S; }
    std::vector<char> used(pt.cols.size(), 0);
    for (uint32_t i = pt.fwd_off[g]; i < pt.fwd_off[g + 1]; ++i) {
      const uint32_t w = pt.fwd[i], op = w >> 28;
      if (op >= 8 || op == obt::F_DESC_CUR || op == obt::F_DESC_STK) used[w & 0xFFFFu] = 1;
    }
    auto& tc = tcols[g / kTmWarps];
    for (size_t c = 0; c < used.size(); ++c) if (used[c] && std::find(tc.begin(), tc.end(), (int)c) == tc.end()) tc.push_back((int)c);
  }
  size_t maxcols = 1;
  for (auto& tc : tcols) { std::sort(tc.begin(), tc.end()); maxcols = std::max(maxcols, tc.size()); }
  S.maxcols_t = (int)maxcols;
  hdr.f("#define OBS_MAXCOLS_T %d\n", (int)maxcols);
  tab.f("__device__ const unsigned short obs_cols_t[] = {");
  std::vector<int> coff{0};
  for (auto& tc : tcols) { for (int c : tc) tab.f("%d,", c); coff.push_back(coff.back() + (int)tc.size()); }
  tab.f("0};\n__device__ const unsigned short obs_coff_t[] = {");
  for (int v : coff) tab.f("%d,", v);
  tab.f("};\n__device__ const unsigned obs_slot_base[] = {");
  for (int g = 0; g < G; ++g) tab.f("%u,", pt.slot_base[g]);
  tab.f("};\n__device__ const unsigned short obs_slot_real[] = {");
  for (int g = 0; g < G; ++g) tab.f("%u,", pt.slot_real[g]);
  tab.f("};\n");
  for (int g = 0; g < G; ++g) {
    const auto& tc = tcols[g / kTmWarps];
    std::vector<int> pos(pt.cols.size(), -1);
    for (size_t i = 0; i < tc.size(); ++i) pos[tc[i]] = (int)i;
    ct.f("case %d: {\n/*BEGIN_CASE_TM %d*/\n", g, g);
    const int n = emit_fwd(ct, pt, g, 1, S.tr_t, 0, pos, /*park=*/true);
    if (n != (int)pt.slot_real[g]) { S.why = "emit count mismatch"; return S; }
    ct.f("/*END_CASE_TM*/\n} break;\n");
    /* machine-readable (comments): the terms of stream g in parking order */
    tab.f("// OBS_STREAM_TM %d type %d terms", g, g / kTmWarps);
    for (uint32_t i = 0; i < pt.slot_real[g]; ++i) tab.f(" %d", (int)pt.slot_term[pt.slot_base[g] + i]);
    tab.f("\n");
  }
  for (int t = 0; t < types; ++t) /* tile column of a type -> (dimension, level) */
    for (size_t i = 0; i < tcols[t].size(); ++i) tab.f("// OBS_LAYOUT_TM %d %d %u %u\n", t, (int)i, pt.cols[tcols[t][i]].dim, pt.cols[tcols[t][i]].level);
  std::string src = scaffold_text();
  replace_marker(src, "//@@TABLES@@", tab.s);
  replace_marker(src, "//@@CASES_TM@@", ct.s);
  S.src = hdr.s + src;
  S.ok = true;
  return S;
}

/* the hyper-gradient module (phi_d_spec only): pa = program compiled with G = 1 (its column table and CSR) */
inline SpecSource generate_dot(const Program& pa, const SpecOptions& opt) {
  using namespace detail;
  SpecSource S;
  S.opt = opt;
  S.tr_a = 32 * opt.qd;
  if (!pa.fast_ok || pa.G != 1 || pa.tmem_cap || pa.K == 0 || pa.aug_dim >= 0) { S.why = "program does not match"; return S; }
  if (256 % S.tr_a || opt.qd < 1 || opt.tgd < 1 || opt.tgd > 8 || opt.npd < 1 || opt.qd * opt.tgd + opt.npd > 8) { S.why = "inconsistent tile options"; return S; }
  if ((int)pa.cols.size() > opt.maxcols_d || pa.d > 32) { S.why = "too many basis columns for the register-resident derivative accumulators"; return S; }
  Emitter hdr, tab, body;
  hdr.f("#define OBS_PARAM_BYTES %d\n#define OBS_DOT_BYTES %d\n#define OBS_K %llu\n#define OBS_NP %d\n#define OBS_HAVE_D 1\n#define OBS_NCOLS_A %d\n#define OBS_PSLEEP %d\n", (int)sizeof(SpecParams),
        (int)sizeof(DotParams), (unsigned long long)pa.K, opt.npd, (int)pa.cols.size(), opt.psleep);
  hdr.f("#define OBS_QD %d\n#define OBS_TGD %d\n", opt.qd, opt.tgd);
  tab.f("__device__ const unsigned short obs_cols_a[] = {");
  for (size_t c = 0; c < pa.cols.size(); ++c) tab.f("%d,", (int)c);
  tab.f("0};\n");
  for (size_t c = 0; c < pa.cols.size(); ++c) tab.f("// OBS_LAYOUT_D %d %u %u\n", (int)c, pa.cols[c].dim, pa.cols[c].level);
  body.f("/*BEGIN_BODY_D*/\n");
  S.nacc = emit_dot(body, tab, pa, S.tr_a, opt.cache_d);
  body.f("/*END_BODY_D*/\n");
  if (S.nacc != (int)pa.K) { S.why = "coefficient count mismatch"; return S; }
  std::string src = scaffold_text();
  replace_marker(src, "//@@TABLES@@", tab.s);
  replace_marker(src, "//@@BODY_D@@", body.s);
  S.src = hdr.s + src;
  S.ok = true;
  return S;
}

} // namespace obs
