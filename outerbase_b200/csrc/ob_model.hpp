/*
 * ob_model.hpp -- host side of the product: the `outermod` state that stays on the CPU.
 *
 * Index tables, per-dimension knot eigenbasis (O(d m^3), m <= ~70) and greedy term
 * selection.  Reference: class outermod, src/modandbase.h:9-54 and
 * src/modandbase.cpp:67-440; setcovfs/setknot src/interfaceR.cpp:53-149;
 * covariance hyper-priors and bounds src/covfuncs.cpp:35-111,166-195,254-283.
 * Everything N-sized lives on the GPU (ob_device.cuh); nothing here touches rows.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

namespace obh {

using u64 = uint64_t;
using i64 = int64_t;

enum CovKind : int { COV_MAT25 = 0, COV_MAT25POW = 1, COV_MAT25ANG = 2 };

struct CovSpec { /* constructor constants of covf_mat25{,pow,ang}, covfuncs.cpp:87-111,166-195,254-283 */
  CovKind kind;
  int numhyp;
  double lb[2], ub[2], h0[2], hvar[2];
  double lowbnd, uppbnd;
};

inline CovSpec cov_spec(const std::string& name) {
  if (name == "mat25") return {COV_MAT25, 1, {-2.25, 0}, {1.5, 0}, {0, 0}, {0.1, 1}, 0.0, 1.0};
  if (name == "mat25pow") return {COV_MAT25POW, 2, {-2.25, -1.25}, {1.5, 1.25}, {0, 0}, {0.1, 0.01}, 0.0, 1.0};
  if (name == "mat25ang") return {COV_MAT25ANG, 2, {-2.25, -2.25}, {1.5, 1.5}, {0, 0}, {0.1, 0.1}, 0.0, 6.283185};
  throw std::range_error("need to choose one of the existing cov functions");
}

/* The 1-D Matern-5/2 family on the host (knot-by-knot, m x m).  The same formulas
 * run on the device for (x, knots); see cov_eval in ob_kernels.cu. */
struct CovPoint { double t, s, c, lg; }; /* transformed coordinate(s) of one input */

inline CovPoint cov_transform(CovKind k, const double* hyp, double x) {
  CovPoint p{0, 0, 0, 0};
  if (k == COV_MAT25) p.t = x / std::exp(2. * hyp[0]);
  else if (k == COV_MAT25POW) {
    const double powv = std::exp(0.25 * hyp[1]);
    p.t = std::pow(x, powv) / std::exp(2. * hyp[0] + 0.25 * hyp[1]);
    p.lg = std::log(x) * p.t;
  } else {
    p.s = std::sin(x) / std::exp(2. * hyp[0]);
    p.c = std::cos(x) / std::exp(2. * hyp[1]);
  }
  return p;
}

/* value and hyper-gradients of the covariance between two transformed points */
inline void cov_pair(CovKind k, const double* hyp, const CovPoint& a, const CovPoint& b, double& v, double* g) {
  if (k == COV_MAT25ANG) {
    const double hs = a.s - b.s, hc = a.c - b.c;
    const double t = std::sqrt(hs * hs + hc * hc);
    const double e = std::exp(-t);
    v = (1 + t + (t * t) / 3) * e;
    if (g) { const double w = e * (t + 1); g[0] = ((2. / 3) * (hs * hs)) * w; g[1] = ((2. / 3) * (hc * hc)) * w; }
    return;
  }
  double h = a.t - b.t;
  const double ah = std::fabs(h), e = std::exp(-ah);
  v = (1 + ah + (ah * ah) / 3) * e;
  if (!g) return;
  const double h2 = (h * (1 + ah)) * e;
  if (k == COV_MAT25) { g[0] = (2. / 3) * (h * h2); return; }
  const double powv = std::exp(0.25 * hyp[1]);
  double s1 = a.lg - b.lg;
  s1 *= (-(0.25 * powv / 3)) * h2;
  h *= h2;
  g[1] = s1 + (0.25 / 3) * h;
  g[0] = (2. / 3) * h;
}

/* small dense column-major matrix */
struct Mat {
  u64 nr = 0, nc = 0;
  std::vector<double> a;
  Mat() {}
  Mat(u64 r, u64 c) : nr(r), nc(c), a(r * c, 0.0) {}
  double& operator()(u64 i, u64 j) { return a[i + j * nr]; }
  double operator()(u64 i, u64 j) const { return a[i + j * nr]; }
};

inline Mat mul(const Mat& A, const Mat& B) {
  Mat C(A.nr, B.nc);
  for (u64 j = 0; j < B.nc; ++j)
    for (u64 i = 0; i < A.nr; ++i) {
      double s = 0;
      for (u64 p = 0; p < A.nc; ++p) s += A(i, p) * B(p, j);
      C(i, j) = s;
    }
  return C;
}

/* two-accumulator sum (Armadillo accumulate order; keeps basisvar sums and hence
 * selectterms' scores identical to what the reference's `sum(...elem(...))` yields) */
inline double sum2(const double* x, u64 n) {
  double s1 = 0, s2 = 0;
  u64 j;
  for (j = 1; j < n; j += 2) { s1 += x[j - 1]; s2 += x[j]; }
  if ((j - 1) < n) s1 += x[j - 1];
  return s1 + s2;
}

/* symmetric eigenproblem by cyclic Jacobi rotations; eigenvalues DESCENDING */
inline void sym_eig_desc(const Mat& S, std::vector<double>& w, Mat& V) {
  const u64 m = S.nr;
  Mat A = S;
  V = Mat(m, m);
  for (u64 i = 0; i < m; ++i) V(i, i) = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0, dg = 0;
    for (u64 j = 0; j < m; ++j)
      for (u64 i = 0; i < m; ++i) (i == j ? dg : off) += A(i, j) * A(i, j);
    if (off <= 1e-60 * dg || off == 0.0) break;
    for (u64 p = 0; p + 1 < m; ++p)
      for (u64 q = p + 1; q < m; ++q) {
        const double apq = A(p, q);
        if (apq == 0.0 || std::fabs(apq) < 1e-300) continue;
        const double theta = (A(q, q) - A(p, p)) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (u64 k = 0; k < m; ++k) { const double x = A(k, p), y = A(k, q); A(k, p) = c * x - s * y; A(k, q) = s * x + c * y; }
        for (u64 k = 0; k < m; ++k) { const double x = A(p, k), y = A(q, k); A(p, k) = c * x - s * y; A(q, k) = s * x + c * y; }
        A(p, q) = 0.0; A(q, p) = 0.0;
        for (u64 k = 0; k < m; ++k) { const double x = V(k, p), y = V(k, q); V(k, p) = c * x - s * y; V(k, q) = s * x + c * y; }
      }
  }
  std::vector<u64> idx(m);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](u64 x, u64 y) { return A(x, x) < A(y, y); });
  std::reverse(idx.begin(), idx.end()); /* ascending (eig_sym) then reversed, modandbase.cpp:236-238 */
  w.resize(m);
  Mat Vs(m, m);
  for (u64 j = 0; j < m; ++j) { w[j] = A(idx[j], idx[j]); for (u64 i = 0; i < m; ++i) Vs(i, j) = V(i, idx[j]); }
  V = Vs;
}

struct OuterMod {
  u64 d = 0;
  std::vector<CovSpec> cov;
  std::vector<double> hyp;
  std::vector<u64> knotptst, hypmatch, hypst, gest, knotptstge;
  std::vector<double> knotpt, basisvar, logbasisvar_gradhyp;
  std::vector<i64> maxlevel;
  Mat rotmat, rotmat_gradhyp;
  bool covs_set = false, knots_set = false;
  u64 select_seed = 0;
  u64 version = 0; /* bumped by every build(); device handles compare it */

  u64 nhyp() const { return hyp.size(); }
  u64 nknot() const { return knotpt.size(); }
  u64 nge() const { return knotptstge.empty() ? 0 : knotptstge[d]; }
  u64 mdim(u64 l) const { return knotptst[l + 1] - knotptst[l]; }

  void layout_hyp() { /* hypst / hypmatch, modandbase.cpp:130-148 */
    hypst.assign(d + 1, 0);
    for (u64 l = 0; l < d; ++l) hypst[l + 1] = hypst[l] + (u64)cov[l].numhyp;
    hypmatch.assign(hypst[d], 0);
    for (u64 l = 0; l < d; ++l) for (u64 h = hypst[l]; h < hypst[l + 1]; ++h) hypmatch[h] = l;
  }
  void layout_ge() { /* gest / knotptstge, modandbase.cpp:183-197 == interfaceR.cpp:133-146 */
    knotptstge.assign(d + 1, 0);
    gest.assign(hypst[d] + 1, 0);
    u64 cur = 0;
    for (u64 l = 0; l < d; ++l) {
      knotptstge[l] = cur;
      for (u64 h = hypst[l]; h < hypst[l + 1]; ++h) { gest[h] = cur; cur += mdim(l); }
    }
    knotptstge[d] = cur;
    gest[hypst[d]] = cur;
  }

  void set_covfs(const std::vector<std::string>& names) { /* interfaceR.cpp:53-73 + hyp_init :128 */
    std::vector<CovSpec> c;
    for (auto& n : names) c.push_back(cov_spec(n));
    cov = c;
    d = cov.size();
    layout_hyp();
    hyp.assign(hypst[d], 0.0);
    for (u64 l = 0; l < d; ++l) for (int h = 0; h < cov[l].numhyp; ++h) hyp[hypst[l] + h] = cov[l].h0[h];
    covs_set = true;
    knots_set = false;
  }

  void set_knot(const double* knots, const u64* lens) { /* interfaceR.cpp:94-149 */
    if (!covs_set) throw std::range_error("Need to set cov. funcs before setting knots.");
    const double* p = knots;
    for (u64 l = 0; l < d; ++l) {
      if (lens[l] < 2) throw std::range_error("need at least two knots per dimension");
      double mn = p[0], mx = p[0];
      for (u64 i = 1; i < lens[l]; ++i) { mn = std::min(mn, p[i]); mx = std::max(mx, p[i]); }
      if (mn < cov[l].lowbnd || mx > cov[l].uppbnd)
        throw std::range_error(std::to_string(l + 1) + "knot point needs to be between " +
                               std::to_string(cov[l].lowbnd) + " and " + std::to_string(cov[l].uppbnd));
      p += lens[l];
    }
    knotptst.assign(d + 1, 0);
    for (u64 l = 0; l < d; ++l) knotptst[l + 1] = knotptst[l] + lens[l];
    knotpt.assign(knots, knots + knotptst[d]);
    knots_set = true;
    layout_ge();
    build();
  }

  void hyp_set(const double* h, u64 n) { /* modandbase.cpp:161-202 */
    layout_hyp();
    if (n != hypst[d]) throw std::range_error("wrongsized vector");
    hyp.assign(h, h + n);
    if (knots_set) { layout_ge(); build(); }
  }

  void build() { /* modandbase.cpp:210-276 */
    u64 mmax = 0;
    for (u64 l = 0; l < d; ++l) mmax = std::max(mmax, mdim(l));
    rotmat = Mat(mmax, nknot());
    rotmat_gradhyp = Mat(mmax, nge());
    basisvar.assign(nknot(), 0.0);
    logbasisvar_gradhyp.assign(nge(), 0.0);
    maxlevel.assign(d, 0);
    for (u64 l = 0; l < d; ++l) {
      const u64 m = mdim(l), H = (u64)cov[l].numhyp;
      const double* hl = hyp.data() + hypst[l];
      std::vector<CovPoint> pts(m);
      for (u64 i = 0; i < m; ++i) pts[i] = cov_transform(cov[l].kind, hl, knotpt[knotptst[l] + i]);
      Mat R(m, m);
      std::vector<Mat> dR(H, Mat(m, m));
      for (u64 j = 0; j < m; ++j)
        for (u64 i = 0; i < m; ++i) {
          double v, g[2];
          cov_pair(cov[l].kind, hl, pts[i], pts[j], v, g);
          R(i, j) = v;
          for (u64 h = 0; h < H; ++h) dR[h](i, j) = g[h];
        }
      std::vector<double> sr;
      Mat U;
      sym_eig_desc(R, sr, U);
      const u64 half = m / 2; /* sign convention :241-242 */
      for (u64 j = 0; j < m; ++j) {
        const double s = U(half, j) + 2.71828 * U(half + 1 < m ? half + 1 : half, j);
        const double sg = s > 0 ? 1.0 : (s < 0 ? -1.0 : 0.0);
        for (u64 i = 0; i < m; ++i) U(i, j) *= sg;
      }
      const double minsv = 0.00000000001 * (sum2(sr.data(), m) / double(m)); /* :245 */
      i64 ml = (i64)m - 1;
      for (u64 j = 0; j + 1 < m; ++j) if (-(sr[j + 1] - sr[j]) < minsv) { ml = (i64)j; break; }
      maxlevel[l] = ml;
      { /* :249  sr += linspace(minsv/1000, m*minsv/1000, m) */
        const double st = minsv / 1000, en = double(m) * minsv / 1000, delta = (en - st) / double(m - 1);
        for (u64 j = 0; j + 1 < m; ++j) sr[j] = sr[j] + (st + double(j) * delta);
        sr[m - 1] = sr[m - 1] + en;
      }
      const double sq = std::sqrt(double(m));
      for (u64 j = 0; j < m; ++j) {
        const double den = sr[j] / sq;
        for (u64 i = 0; i < m; ++i) rotmat(i, knotptst[l] + j) = U(i, j) / den;
        basisvar[knotptst[l] + j] = std::log(sr[j] / double(m));
      }
      Mat Ut(m, m);
      for (u64 j = 0; j < m; ++j) for (u64 i = 0; i < m; ++i) Ut(i, j) = U(j, i);
      for (u64 h = 0; h < H; ++h) { /* :258-274 */
        Mat UtdRU = mul(mul(Ut, dR[h]), U);
        Mat Wm(m, m);
        for (u64 j = 0; j < m; ++j) {
          logbasisvar_gradhyp[knotptstge[l] + h * m + j] = UtdRU(j, j) / sr[j];
          for (u64 i = 0; i < m; ++i) Wm(i, j) = UtdRU(i, j) * (1 / (((i == j) ? 0.0 : sr[j]) - sr[i]));
        }
        Mat Ah = mul(U, Wm);
        for (u64 j = 0; j < m; ++j) {
          const double den = sr[j] / sq;
          for (u64 i = 0; i < m; ++i) rotmat_gradhyp(i, knotptstge[l] + h * m + j) = Ah(i, j) / den;
        }
      }
    }
    ++version;
  }

  void check_terms(const u64* terms, u64 K) const {
    for (u64 l = 0; l < d; ++l)
      for (u64 k = 0; k < K; ++k)
        if (terms[k + l * K] >= mdim(l)) throw std::range_error("terms level exceeds the number of knots");
  }

  void getvar(const u64* terms, u64 K, double* out) const { /* modandbase.cpp:350-356 */
    check_terms(terms, K);
    std::vector<double> t(d);
    for (u64 k = 0; k < K; ++k) {
      for (u64 l = 0; l < d; ++l) t[l] = basisvar[knotptst[l] + terms[k + l * K]];
      out[k] = std::exp(sum2(t.data(), d));
    }
  }

  void getlvar_gradhyp(const u64* terms, u64 K, double* out /* K x H */) const { /* :364-379 */
    check_terms(terms, K);
    std::fill(out, out + K * nhyp(), 0.0);
    for (u64 l = 0; l < d; ++l)
      for (u64 h = hypst[l]; h < hypst[l + 1]; ++h)
        for (u64 k = 0; k < K; ++k) out[k + h * K] += logbasisvar_gradhyp[gest[h] + terms[k + l * K]];
  }

  /* best-first growth of a downward-closed multi-index set, modandbase.cpp:387-440 */
  void selectterms(u64 numele, u64* out /* numele x d col-major */) const {
    if (!knots_set) throw std::range_error("Need to set covfs and knots before selecting terms.");
    std::vector<std::vector<i64>> chosen, cand;
    std::vector<double> cscore;
    auto score = [&](const std::vector<i64>& t) {
      std::vector<double> s(d);
      for (u64 l = 0; l < d; ++l) s[l] = basisvar[knotptst[l] + (u64)t[l]];
      return sum2(s.data(), d);
    };
    cand.push_back(std::vector<i64>(d, 0));
    cscore.push_back(score(cand[0]));
    u64 rng = select_seed;
    auto next = [&]() {
      u64 z = (rng += 0x9e3779b97f4a7c15ULL);
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
      return z ^ (z >> 31);
    };
    for (u64 it = 0; it < numele; ++it) {
      if (cand.empty()) throw std::range_error("selectterms: no candidate terms left");
      const double cut = -0.1 + *std::max_element(cscore.begin(), cscore.end());
      std::vector<u64> big;
      for (u64 i = 0; i < cscore.size(); ++i) if (cscore[i] > cut) big.push_back(i);
      const u64 pick = (select_seed == 0) ? big[0] : big[next() % big.size()];
      chosen.push_back(cand[pick]);
      const u64 last = cand.size() - 1;
      if (last > pick) { cand[pick] = cand[last]; cscore[pick] = cscore[last]; }
      cand.pop_back(); cscore.pop_back();
      const std::vector<i64>& nw = chosen.back();
      const u64 nd = chosen.size();
      std::vector<char> swapped(nd); /* rows equal to nw with one unit moved between two dims (:420-422) */
      for (u64 i = 0; i < nd; ++i) {
        i64 s = 0, sa = 0;
        for (u64 l = 0; l < d; ++l) { const i64 df = nw[l] - chosen[i][l]; s += df; sa += df < 0 ? -df : df; }
        swapped[i] = (s == 0) && (sa == 2);
      }
      i64 nnz = 0;
      for (u64 l = 0; l < d; ++l) nnz += nw[l] > 0;
      for (u64 l = 0; l < d; ++l) {
        if (nw[l] >= maxlevel[l]) continue;
        i64 have = 1;
        for (u64 i = 0; i < nd; ++i) have += swapped[i] && (nw[l] - chosen[i][l] == -1);
        const i64 need = nnz + (nw[l] < 1);
        if (need == have) {
          std::vector<i64> c = nw;
          c[l] += 1;
          cand.push_back(c);
          cscore.push_back(score(c));
        }
      }
    }
    for (u64 k = 0; k < numele; ++k) for (u64 l = 0; l < d; ++l) out[k + l * numele] = (u64)chosen[k][l];
  }

  /* hyper-prior: covf::lpdf / lpdf_gradhyp (covfuncs.cpp:35-70) summed over dims (modandbase.cpp:89-119) */
  double hyplpdf(const double* hp, u64 n) const {
    if (n != hyp.size()) return -std::numeric_limits<double>::infinity();
    double out = 0;
    for (u64 l = 0; l < d; ++l) {
      double o = 0;
      std::vector<double> t(cov[l].numhyp);
      for (int h = 0; h < cov[l].numhyp; ++h) {
        const double v = hp[hypst[l] + h];
        if (cov[l].ub[h] < v || cov[l].lb[h] > v) return -std::numeric_limits<double>::infinity();
        o += 5 * std::log(cov[l].ub[h] - v);
        o += 5 * std::log(v - cov[l].lb[h]);
        const double e = v - cov[l].h0[h];
        t[h] = e * e / cov[l].hvar[h];
      }
      o -= 0.5 * sum2(t.data(), t.size());
      out += o;
    }
    return out;
  }
  void hyplpdf_grad(const double* hp, u64 n, double* out) const {
    std::fill(out, out + hyp.size(), 0.0);
    if (n != hyp.size()) return;
    for (u64 l = 0; l < d; ++l) {
      bool ok = true;
      std::vector<double> g(cov[l].numhyp, 0.0);
      for (int h = 0; h < cov[l].numhyp && ok; ++h) {
        const double v = hp[hypst[l] + h];
        if (cov[l].ub[h] < v || cov[l].lb[h] > v) { ok = false; break; }
        g[h] -= 5 / (cov[l].ub[h] - v);
        g[h] += 5 / (v - cov[l].lb[h]);
      }
      if (!ok) { /* covfuncs.cpp:63-64 returns the partially filled vector */
        for (int h = 0; h < cov[l].numhyp; ++h) out[hypst[l] + h] = g[h];
        continue;
      }
      for (int h = 0; h < cov[l].numhyp; ++h) out[hypst[l] + h] = g[h] - (hp[hypst[l] + h] - cov[l].h0[h]) / cov[l].hvar[h];
    }
  }
};

/* outerbase::setloopvals_ (modandbase.cpp:504-513): R-visible, ignored by the GPU path */
inline void loopvals(u64 n_row, u64 nthreads, u64& chunksize, u64& loopsize, bool& vertpl) {
  if (nthreads < 1) nthreads = 1;
  const u64 maxchunk = 1 + (2048 / nthreads), minchunk = 32;
  chunksize = std::max(minchunk, std::min(maxchunk, n_row / (4 * nthreads) + 1));
  loopsize = (n_row + chunksize - 1) / chunksize;
  vertpl = loopsize > 20;
}

} // namespace obh
