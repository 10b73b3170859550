/*
 * ob_oracle.cpp -- CPU ORACLE (test infrastructure only; see ob_oracle.hpp header).
 * Restates, with the reference's loop structure and floating-point order:
 *   src/covfuncs.cpp, src/modandbase.cpp, src/linalg.cpp, src/fit.cpp,
 *   src/lpdfs/loglik_gauss.cpp, src/lpdfs/logpr_gauss.cpp, src/interfaceR.cpp:53-149.
 */
#include "ob_oracle.hpp"

#include <algorithm>
#include <numeric>
#ifdef _OPENMP
#include <omp.h>
#else
static inline int omp_get_num_procs() { return 1; }
static inline int omp_get_thread_num() { return 0; }
static inline int omp_get_num_threads() { return 1; }
#endif

namespace orc {

/* ------------------------------------------------------------------ helpers */

double accu2(const double* x, u64 n) {
  double acc1 = 0, acc2 = 0;
  u64 j;
  for (j = 1; j < n; j += 2) { acc1 += x[j - 1]; acc2 += x[j]; }
  if ((j - 1) < n) acc1 += x[j - 1];
  return acc1 + acc2;
}

double dot2(const double* x, const double* y, u64 n) {
  double acc1 = 0, acc2 = 0;
  u64 j;
  for (j = 1; j < n; j += 2) { acc1 += x[j - 1] * y[j - 1]; acc2 += x[j] * y[j]; }
  if ((j - 1) < n) acc1 += x[j - 1] * y[j - 1];
  return acc1 + acc2;
}

/* What Armadillo hands to BLAS is restated in the order of netlib's reference BLAS -- R's bundled libRblas, the
 * implementation `PKG_LIBS = $(BLAS_LIBS)` (src/Makevars:13) links by default: ddot and dgemv('T') accumulate LEFT TO
 * RIGHT in one accumulator.  Call sites: `x.t() * M` (gemv: linalg.cpp:574, logpr_gauss.cpp:103, loglik_gauss.cpp:127,
 * loglik_gda.cpp:139-150) always; `dot(x, y)` (linalg.cpp:382-384) above 32 elements -- op_dot::direct_dot runs
 * Armadillo's own two-accumulator loop up to 32.  oracle/_ref (the unmodified reference sources on the Armadillo
 * shim, ARMA_SHIM_BLAS=1) makes the same choice; tests/test_oracle_ref.py holds the two bitwise equal. */
double dot_seq(const double* x, const double* y, u64 n) {
  double s = 0;
  for (u64 i = 0; i < n; ++i) s = s + x[i] * y[i];
  return s;
}
double dot_arma(const double* x, const double* y, u64 n) { return n <= 32 ? dot2(x, y, n) : dot_seq(x, y, n); }

static bool all_finite(const vec& v) {
  for (double x : v) if (!std::isfinite(x)) return false;
  return true;
}

/* Armadillo op_mean::direct_mean / op_var::direct_var (call site loglik_gauss.cpp:48) */
static double arma_mean(const double* x, u64 n) { return accu2(x, n) / double(n); }
static double arma_var(const double* x, u64 n) {
  if (n < 2) return 0.0;
  const double acc1 = arma_mean(x, n);
  double acc2 = 0, acc3 = 0;
  u64 i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) {
    const double ti = acc1 - x[i], tj = acc1 - x[j];
    acc2 += ti * ti + tj * tj;
    acc3 += ti + tj;
  }
  if (i < n) { const double ti = acc1 - x[i]; acc2 += ti * ti; acc3 += ti; }
  return (acc2 - acc3 * acc3 / double(n)) / double(n - 1);
}

/* Cyclic Jacobi eigen-decomposition of a symmetric m x m matrix (stands in for
 * LAPACK dsyev behind eig_sym, modandbase.cpp:236).  Returns eigenvalues in
 * ASCENDING order (eig_sym's convention) with matching eigenvector columns. */
static void eig_sym_jacobi(vec& w, mat& V, const mat& Ain) {
  const u64 m = Ain.nr;
  mat A = Ain;
  V.set_size(m, m);
  V.zeros();
  for (u64 i = 0; i < m; ++i) V(i, i) = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0, dg = 0;
    for (u64 j = 0; j < m; ++j)
      for (u64 i = 0; i < m; ++i) (i == j ? dg : off) += A(i, j) * A(i, j);
    if (off <= 1e-60 * dg || off == 0.0) break;
    for (u64 p = 0; p + 1 < m; ++p) {
      for (u64 q = p + 1; q < m; ++q) {
        const double apq = A(p, q);
        if (apq == 0.0) continue;
        const double app = A(p, p), aqq = A(q, q);
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (u64 k = 0; k < m; ++k) { /* A <- A J */
          const double akp = A(k, p), akq = A(k, q);
          A(k, p) = c * akp - s * akq;
          A(k, q) = s * akp + c * akq;
        }
        for (u64 k = 0; k < m; ++k) { /* A <- J^T A */
          const double apk = A(p, k), aqk = A(q, k);
          A(p, k) = c * apk - s * aqk;
          A(q, k) = s * apk + c * aqk;
        }
        A(p, q) = 0.0; A(q, p) = 0.0;
        for (u64 k = 0; k < m; ++k) {
          const double vkp = V(k, p), vkq = V(k, q);
          V(k, p) = c * vkp - s * vkq;
          V(k, q) = s * vkp + c * vkq;
        }
      }
    }
  }
  std::vector<u64> idx(m);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](u64 a, u64 b) { return A(a, a) < A(b, b); });
  w.resize(m);
  mat Vs(m, m);
  for (u64 j = 0; j < m; ++j) {
    w[j] = A(idx[j], idx[j]);
    for (u64 i = 0; i < m; ++i) Vs(i, j) = V(i, idx[j]);
  }
  V = Vs;
}

/* C = A * B, k-ascending (stands in for BLAS dgemm behind Armadillo `*`) */
static void matmul(mat& C, const mat& A, const mat& B) {
  C.set_size(A.nr, B.nc);
  for (u64 j = 0; j < B.nc; ++j)
    for (u64 i = 0; i < A.nr; ++i) {
      double s = 0;
      for (u64 p = 0; p < A.nc; ++p) s += A(i, p) * B(p, j);
      C(i, j) = s;
    }
}

/* ------------------------------------------------------------------ covf */

double covf::lpdf(const vec& hypp) const { /* covfuncs.cpp:35-50 */
  double out = 0;
  if (hyp.size() != hypp.size()) return -std::numeric_limits<double>::infinity();
  for (size_t l = 0; l < hypp.size(); ++l) {
    if (hypub[l] < hypp[l]) return -std::numeric_limits<double>::infinity();
    if (hyplb[l] > hypp[l]) return -std::numeric_limits<double>::infinity();
    out += 5 * std::log(hypub[l] - hypp[l]);
    out += 5 * std::log(hypp[l] - hyplb[l]);
  }
  vec t(hypp.size());
  for (size_t l = 0; l < hypp.size(); ++l) { const double e = hypp[l] - hyp0[l]; t[l] = e * e / hypvar[l]; }
  out -= 0.5 * accu2(t.data(), t.size());
  return out;
}

vec covf::lpdf_gradhyp(const vec& hypp) const { /* covfuncs.cpp:53-70 */
  vec out(hyp.size(), 0.0);
  if (hyp.size() != hypp.size()) return out;
  for (size_t l = 0; l < hypp.size(); ++l) {
    if (hypub[l] < hypp[l]) return out;
    if (hyplb[l] > hypp[l]) return out;
    out[l] -= 5 / (hypub[l] - hypp[l]);
    out[l] += 5 / (hypp[l] - hyplb[l]);
  }
  for (size_t l = 0; l < hypp.size(); ++l) out[l] -= (hypp[l] - hyp0[l]) / hypvar[l];
  return out;
}

bool covf::inputcheck(const double* x, u64 n) const { /* covfuncs.h:23-27 */
  double mn = x[0], mx = x[0];
  for (u64 i = 1; i < n; ++i) { mn = std::min(mn, x[i]); mx = std::max(mx, x[i]); }
  if (mn < lowbnd) return false;
  if (mx > uppbnd) return false;
  return true;
}

namespace {

struct covf_mat25 : covf { /* covfuncs.cpp:87-150 */
  double a = 2.;
  covf_mat25() {
    numhyp = 1; hypnames = {"scale"};
    hyp = {0}; hyplb = {-2.25}; hypub = {1.5}; hyp0 = {0}; hypvar = {0.1};
    lowbnd = 0; uppbnd = 1;
  }
  void cov(mat& h, const double* x1, u64 n1, const double* x2, u64 n2) const override { /* :113-126 */
    const double expLS = std::exp(a * hyp[0]);
    vec x1t(n1), x2t(n2);
    for (u64 i = 0; i < n1; ++i) x1t[i] = x1[i] / expLS;
    for (u64 j = 0; j < n2; ++j) x2t[j] = x2[j] / expLS;
    h.set_size(n1, n2);
    for (u64 j = 0; j < n2; ++j)
      for (u64 i = 0; i < n1; ++i) {
        const double t = std::fabs(x1t[i] - x2t[j]);
        h(i, j) = (1 + t + (t * t) / 3) * std::exp(-t);
      }
  }
  void cov_gradhyp(std::vector<mat>& g, const double* x1, u64 n1, const double* x2, u64 n2) const override { /* :134-150 */
    const double expLS = std::exp(a * hyp[0]);
    vec x1t(n1), x2t(n2);
    for (u64 i = 0; i < n1; ++i) x1t[i] = x1[i] / expLS;
    for (u64 j = 0; j < n2; ++j) x2t[j] = x2[j] / expLS;
    g.assign(1, mat(n1, n2));
    for (u64 j = 0; j < n2; ++j)
      for (u64 i = 0; i < n1; ++i) {
        const double h = x1t[i] - x2t[j];
        const double h2 = (h * (1 + std::fabs(h))) * std::exp(-std::fabs(h));
        g[0](i, j) = (a / 3) * (h * h2);
      }
  }
};

struct covf_mat25pow : covf { /* covfuncs.cpp:166-243 */
  double a = 2., b = 0.25;
  covf_mat25pow() {
    numhyp = 2; hypnames = {"scale", "power"};
    hyp = {0, 0}; hyplb = {-2.25, -1.25}; hypub = {1.5, 1.25}; hyp0 = {0, 0}; hypvar = {0.1, 0.01};
    lowbnd = 0; uppbnd = 1;
  }
  void cov(mat& h, const double* x1, u64 n1, const double* x2, u64 n2) const override { /* :197-212 */
    const double powv = std::exp(b * hyp[1]);
    const double expLS = std::exp(a * hyp[0] + b * hyp[1]);
    vec x1t(n1), x2t(n2);
    for (u64 i = 0; i < n1; ++i) x1t[i] = std::pow(x1[i], powv) / expLS;
    for (u64 j = 0; j < n2; ++j) x2t[j] = std::pow(x2[j], powv) / expLS;
    h.set_size(n1, n2);
    for (u64 j = 0; j < n2; ++j)
      for (u64 i = 0; i < n1; ++i) {
        const double t = std::fabs(x1t[i] - x2t[j]);
        h(i, j) = (1 + t + (t * t) / 3) * std::exp(-t);
      }
  }
  void cov_gradhyp(std::vector<mat>& g, const double* x1, u64 n1, const double* x2, u64 n2) const override { /* :220-243 */
    const double powv = std::exp(b * hyp[1]);
    const double expLS = std::exp(a * hyp[0] + b * hyp[1]);
    vec x1t(n1), x2t(n2), l1(n1), l2(n2);
    for (u64 i = 0; i < n1; ++i) { x1t[i] = std::pow(x1[i], powv) / expLS; l1[i] = std::log(x1[i]) * x1t[i]; }
    for (u64 j = 0; j < n2; ++j) { x2t[j] = std::pow(x2[j], powv) / expLS; l2[j] = std::log(x2[j]) * x2t[j]; }
    g.assign(2, mat(n1, n2));
    const double c1 = -(b * powv / 3);
    for (u64 j = 0; j < n2; ++j)
      for (u64 i = 0; i < n1; ++i) {
        double h = x1t[i] - x2t[j];
        const double h2 = (h * (1 + std::fabs(h))) * std::exp(-std::fabs(h));
        double s1 = l1[i] - l2[j];
        s1 *= c1 * h2;
        h *= h2;
        s1 += (b / 3) * h;
        g[1](i, j) = s1;
        g[0](i, j) = (a / 3) * h;
      }
  }
};

struct covf_mat25ang : covf { /* covfuncs.cpp:254-347 */
  double a = 2.;
  covf_mat25ang() {
    numhyp = 2; hypnames = {"sin.sc", "cos.sc"};
    hyp = {0, 0}; hyplb = {-2.25, -2.25}; hypub = {1.5, 1.5}; hyp0 = {0, 0}; hypvar = {0.1, 0.1};
    lowbnd = 0; uppbnd = 6.283185;
  }
  void prep(vec& s1, vec& c1, vec& s2, vec& c2, const double* x1, u64 n1, const double* x2, u64 n2) const {
    const double expLSs = std::exp(a * hyp[0]), expLSc = std::exp(a * hyp[1]);
    s1.resize(n1); c1.resize(n1); s2.resize(n2); c2.resize(n2);
    for (u64 i = 0; i < n1; ++i) { s1[i] = std::sin(x1[i]) / expLSs; c1[i] = std::cos(x1[i]) / expLSc; }
    for (u64 j = 0; j < n2; ++j) { s2[j] = std::sin(x2[j]) / expLSs; c2[j] = std::cos(x2[j]) / expLSc; }
  }
  void cov(mat& h, const double* x1, u64 n1, const double* x2, u64 n2) const override { /* :285-310 */
    vec s1, c1, s2, c2;
    prep(s1, c1, s2, c2, x1, n1, x2, n2);
    h.set_size(n1, n2);
    for (u64 j = 0; j < n2; ++j)
      for (u64 i = 0; i < n1; ++i) {
        const double hs = s1[i] - s2[j], hc = c1[i] - c2[j];
        const double t = std::sqrt(hs * hs + hc * hc);
        h(i, j) = (1 + t + (t * t) / 3) * std::exp(-t);
      }
  }
  void cov_gradhyp(std::vector<mat>& g, const double* x1, u64 n1, const double* x2, u64 n2) const override { /* :318-347 */
    vec s1, c1, s2, c2;
    prep(s1, c1, s2, c2, x1, n1, x2, n2);
    g.assign(2, mat(n1, n2));
    for (u64 j = 0; j < n2; ++j)
      for (u64 i = 0; i < n1; ++i) {
        const double hs = s1[i] - s2[j], hc = c1[i] - c2[j];
        const double t = std::sqrt(hs * hs + hc * hc);
        const double e = std::exp(-t) * (t + 1);
        g[0](i, j) = ((a / 3) * (hs * hs)) * e;
        g[1](i, j) = ((a / 3) * (hc * hc)) * e;
      }
  }
};

} // namespace

std::unique_ptr<covf> make_covf(const std::string& name) {
  if (name == "mat25") return std::unique_ptr<covf>(new covf_mat25());
  if (name == "mat25pow") return std::unique_ptr<covf>(new covf_mat25pow());
  if (name == "mat25ang") return std::unique_ptr<covf>(new covf_mat25ang());
  throw std::range_error("need to choose one of the existing cov functions");
}

/* ------------------------------------------------------------------ outermod */

void outermod::set_covfs(const std::vector<std::string>& names) { /* interfaceR.cpp:53-73 */
  d = names.size();
  covflist.clear();
  for (u64 k = 0; k < d; ++k) covflist.push_back(make_covf(names[k]));
  hyp_init();
  setcovfs = true;
  setknots = false;
}

void outermod::set_knot(const std::vector<vec>& L) { /* interfaceR.cpp:94-149 */
  if (!setcovfs) throw std::range_error("Need to set cov. funcs before setting knots.");
  if (L.size() != d) throw std::range_error("dim needs to match" + std::to_string(d) + ".");
  for (u64 l = 0; l < d; ++l) {
    if (L[l].size() < 2) throw std::range_error("need at least two knots per dimension");
    if (!covflist[l]->inputcheck(L[l].data(), L[l].size()))
      throw std::range_error(std::to_string(l + 1) + "knot point needs to be between " +
                             std::to_string(covflist[l]->lowbnd) + " and " + std::to_string(covflist[l]->uppbnd));
  }
  knotptst.assign(d + 1, 0);
  u64 currst = 0;
  for (u64 l = 0; l < d; ++l) { knotptst[l] = currst; currst += L[l].size(); }
  knotptst[d] = currst;
  knotpt.assign(currst, 0.0);
  for (u64 l = 0; l < d; ++l) std::copy(L[l].begin(), L[l].end(), knotpt.begin() + knotptst[l]);
  setknots = true;

  knotptstge.assign(d + 1, 0);
  gest.assign(hypst[d] + 1, 0);
  currst = 0;
  u64 currstalt = 0;
  for (u64 l = 0; l < d; ++l) {
    knotptstge[l] = currst;
    for (u64 k = 0; k < hypst[l + 1] - hypst[l]; ++k) {
      hypmatch[currstalt] = l;
      gest[currstalt] = currst;
      currst += knotptst[l + 1] - knotptst[l];
      currstalt += 1;
    }
  }
  knotptstge[d] = currst;
  gest[hypst[d]] = currst;
  build();
}

void outermod::hyp_init() { /* modandbase.cpp:128-153 */
  hypst.assign(d + 1, 0);
  u64 currst = 0;
  for (u64 l = 0; l < d; ++l) { hypst[l] = currst; currst += covflist[l]->numhyp; }
  hypst[d] = currst;
  hyp.assign(currst, 0.0);
  for (u64 l = 0; l < d; ++l)
    for (u64 k = 0; k < covflist[l]->numhyp; ++k) hyp[hypst[l] + k] = covflist[l]->hyp0[k];
  hypmatch.assign(hypst[d], 0);
  u64 currstalt = 0;
  for (u64 l = 0; l < d; ++l)
    for (u64 k = 0; k < hypst[l + 1] - hypst[l]; ++k) { hypmatch[currstalt] = l; currstalt++; }
  hyp_set(hyp);
}

void outermod::hyp_set(const vec& hyp_) { /* modandbase.cpp:161-202 */
  hypst.assign(d + 1, 0);
  u64 currst = 0;
  for (u64 l = 0; l < d; ++l) { hypst[l] = currst; currst += covflist[l]->numhyp; }
  hypst[d] = currst;
  if (hyp_.size() != currst) throw std::range_error("wrongsized vector");
  hyp = hyp_;
  for (u64 l = 0; l < d; ++l)
    for (u64 k = 0; k < covflist[l]->numhyp; ++k) covflist[l]->hyp[k] = hyp[hypst[l] + k];
  if (setknots) {
    knotptstge.assign(d + 1, 0);
    hypmatch.assign(hypst[d], 0);
    gest.assign(hypst[d] + 1, 0);
    currst = 0;
    u64 currstalt = 0;
    for (u64 l = 0; l < d; ++l) {
      knotptstge[l] = currst;
      for (u64 k = 0; k < hypst[l + 1] - hypst[l]; ++k) {
        hypmatch[currstalt] = l;
        gest[currstalt] = currst;
        currst += knotptst[l + 1] - knotptst[l];
        currstalt += 1;
      }
    }
    knotptstge[d] = currst;
    gest[hypst[d]] = currst;
    build();
  }
}

void outermod::setsizes_() { /* modandbase.cpp:67-81 */
  u64 mmax = 0;
  for (u64 l = 0; l < d; ++l) mmax = std::max(mmax, knotptst[l + 1] - knotptst[l]);
  rotmat.set_size(mmax, knotpt.size()); rotmat.zeros();
  rotmat_gradhyp.set_size(mmax, knotptstge[d]); rotmat_gradhyp.zeros();
  basisvar.assign(knotpt.size(), 0.0);
  logbasisvar_gradhyp.assign(knotptstge[d], 0.0);
  maxlevel.assign(d, 0);
}

void outermod::build() { /* modandbase.cpp:210-276 */
  setsizes_();
  for (u64 k = 0; k < d; ++k) {
    const double* xsh = knotpt.data() + knotptst[k];
    const u64 lenh = knotptst[k + 1] - knotptst[k];
    mat R;
    covflist[k]->cov(R, xsh, lenh, xsh, lenh);
    vec sra;
    mat Ua;
    eig_sym_jacobi(sra, Ua, R);
    vec sr(lenh);
    mat U(lenh, lenh);
    for (u64 j = 0; j < lenh; ++j) { /* reverse / fliplr :237-238 */
      sr[j] = sra[lenh - 1 - j];
      for (u64 i = 0; i < lenh; ++i) U(i, j) = Ua(i, lenh - 1 - j);
    }
    const u64 halfw = lenh / 2; /* sign ambiguity :241-242 */
    for (u64 j = 0; j < lenh; ++j) {
      const double s = U(halfw, j) + 2.71828 * U(halfw + 1 < lenh ? halfw + 1 : halfw, j);
      const double sg = (s > 0) ? 1.0 : ((s < 0) ? -1.0 : 0.0);
      for (u64 i = 0; i < lenh; ++i) U(i, j) *= sg;
    }
    const double minsv = 0.00000000001 * arma_mean(sr.data(), lenh); /* :245 */
    i64 ml = (i64)lenh - 1;
    for (u64 j = 0; j + 1 < lenh; ++j)
      if (-(sr[j + 1] - sr[j]) < minsv) { ml = (i64)j; break; }
    maxlevel[k] = ml;
    { /* sr = sr + linspace(minsv/1000, lenh*minsv/1000, lenh) :249 */
      const double st = minsv / 1000, en = double(lenh) * minsv / 1000;
      const double delta = (en - st) / double(lenh - 1);
      for (u64 j = 0; j + 1 < lenh; ++j) sr[j] = sr[j] + (st + double(j) * delta);
      sr[lenh - 1] = sr[lenh - 1] + en;
    }
    const double sq = std::sqrt(double(lenh));
    for (u64 j = 0; j < lenh; ++j) {
      const double den = sr[j] / sq;
      for (u64 i = 0; i < lenh; ++i) rotmat(i, knotptst[k] + j) = U(i, j) / den;
      basisvar[knotptst[k] + j] = std::log(sr[j] / double(lenh));
    }
    /* gradient matrices :258-274 */
    std::vector<mat> Rge;
    covflist[k]->cov_gradhyp(Rge, xsh, lenh, xsh, lenh);
    mat Fm(lenh, lenh);
    for (u64 j = 0; j < lenh; ++j)
      for (u64 i = 0; i < lenh; ++i) Fm(i, j) = 1 / (((i == j) ? 0.0 : sr[j]) - sr[i]);
    mat Ut(lenh, lenh);
    for (u64 j = 0; j < lenh; ++j)
      for (u64 i = 0; i < lenh; ++i) Ut(i, j) = U(j, i);
    for (u64 l = 0; l < hypst[k + 1] - hypst[k]; ++l) {
      mat T1, UtdRV, Ah;
      matmul(T1, Ut, Rge[l]);
      matmul(UtdRV, T1, U);
      for (u64 j = 0; j < lenh; ++j) logbasisvar_gradhyp[knotptstge[k] + l * lenh + j] = UtdRV(j, j) / sr[j];
      mat W(lenh, lenh);
      for (u64 j = 0; j < lenh; ++j)
        for (u64 i = 0; i < lenh; ++i) W(i, j) = UtdRV(i, j) * Fm(i, j);
      matmul(Ah, U, W);
      for (u64 j = 0; j < lenh; ++j) {
        const double den = sr[j] / sq;
        for (u64 i = 0; i < lenh; ++i) rotmat_gradhyp(i, knotptstge[k] + l * lenh + j) = Ah(i, j) / den;
      }
    }
  }
}

/* R = cov(x, knots_k) * rotmat_k, k-ascending contraction; then cols 1.. /= col 0 */
void outermod::buildob(mat& R, const double* xcol, u64 n, u64 k) const { /* :285-298 */
  const u64 lenh = knotptst[k + 1] - knotptst[k];
  mat C;
  covflist[k]->cov(C, xcol, n, knotpt.data() + knotptst[k], lenh);
  R.set_size(n, lenh);
  for (u64 j = 0; j < lenh; ++j)
    for (u64 i = 0; i < n; ++i) {
      double s = 0;
      for (u64 p = 0; p < lenh; ++p) s += C(i, p) * rotmat(p, knotptst[k] + j);
      R(i, j) = s;
    }
  for (u64 j = 1; j < lenh; ++j)
    for (u64 i = 0; i < n; ++i) R(i, j) /= R(i, 0);
}

void outermod::buildob(mat& R, std::vector<mat>& Rt, const double* xcol, u64 n, u64 k) const { /* :306-327 */
  const u64 lenh = knotptst[k + 1] - knotptst[k];
  mat C;
  std::vector<mat> Cg;
  covflist[k]->cov(C, xcol, n, knotpt.data() + knotptst[k], lenh);
  covflist[k]->cov_gradhyp(Cg, xcol, n, knotpt.data() + knotptst[k], lenh);
  Rt.assign(hypst[k + 1] - hypst[k], mat(n, lenh));
  for (u64 l = hypst[k]; l < hypst[k + 1]; ++l) {
    mat& S = Rt[l - hypst[k]];
    const mat& G = Cg[l - hypst[k]];
    for (u64 j = 0; j < lenh; ++j)
      for (u64 i = 0; i < n; ++i) {
        double s1 = 0, s2 = 0;
        for (u64 p = 0; p < lenh; ++p) s1 += G(i, p) * rotmat(p, knotptst[k] + j);
        for (u64 p = 0; p < lenh; ++p) s2 += C(i, p) * rotmat_gradhyp(p, gest[l] + j);
        S(i, j) = s1 + s2;
      }
  }
  R.set_size(n, lenh);
  for (u64 j = 0; j < lenh; ++j)
    for (u64 i = 0; i < n; ++i) {
      double s = 0;
      for (u64 p = 0; p < lenh; ++p) s += C(i, p) * rotmat(p, knotptst[k] + j);
      R(i, j) = s;
    }
  for (u64 l = 0; l < Rt.size(); ++l)
    for (u64 j = 0; j < lenh; ++j)
      for (u64 i = 0; i < n; ++i) Rt[l](i, j) /= R(i, 0);
  for (u64 j = 1; j < lenh; ++j)
    for (u64 i = 0; i < n; ++i) R(i, j) /= R(i, 0);
}

vec outermod::getvar(const umat& terms) const { /* :350-356 */
  vec out(terms.nr), t(d);
  for (u64 k = 0; k < terms.nr; ++k) {
    for (u64 l = 0; l < d; ++l) t[l] = basisvar[knotptst[l] + terms(k, l)];
    out[k] = std::exp(accu2(t.data(), d));
  }
  return out;
}

mat outermod::getlvar_gradhyp(const umat& terms) const { /* :364-379 */
  mat out(terms.nr, hypmatch.size());
  out.zeros();
  for (u64 k = 0; k < d; ++k)
    for (u64 l = hypst[k]; l < hypst[k + 1]; ++l)
      for (u64 r = 0; r < terms.nr; ++r) out(r, l) += logbasisvar_gradhyp[gest[l] + terms(r, k)];
  return out;
}

static inline u64 splitmix64(u64& s) {
  u64 z = (s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

umat outermod::selectterms(unsigned numele) const { /* :387-440 */
  std::vector<std::vector<i64>> terms, pterms; /* growable instead of 10*numele rows (:394) */
  vec ptv;
  terms.reserve(numele);
  auto score = [&](const std::vector<i64>& t) {
    vec s(d);
    for (u64 l = 0; l < d; ++l) s[l] = basisvar[knotptst[l] + (u64)t[l]];
    return accu2(s.data(), d);
  };
  pterms.push_back(std::vector<i64>(d, 0));
  ptv.push_back(score(pterms[0]));
  u64 rng = select_seed;
  for (unsigned k = 0; k < numele; ++k) {
    if (pterms.empty()) throw std::range_error("selectterms: no candidate terms left");
    const double mval = -0.1 + *std::max_element(ptv.begin(), ptv.end());
    std::vector<u64> islarge;
    for (u64 i = 0; i < ptv.size(); ++i) if (ptv[i] > mval) islarge.push_back(i);
    u64 kstar = islarge[0];
    if (select_seed != 0) kstar = islarge[splitmix64(rng) % islarge.size()];
    terms.push_back(pterms[kstar]);
    const u64 np = pterms.size() - 1; /* swap-remove :413-417 */
    if (np > kstar) { pterms[kstar] = pterms[np]; ptv[kstar] = ptv[np]; }
    pterms.pop_back(); ptv.pop_back();
    const u64 nd = terms.size();
    const std::vector<i64>& nw = terms[nd - 1];
    /* gt(i): new - terms[i] has sum 0 and abs-sum 2 (:420-422) */
    std::vector<char> gt(nd);
    for (u64 i = 0; i < nd; ++i) {
      i64 s = 0, sa = 0;
      for (u64 l = 0; l < d; ++l) { const i64 df = nw[l] - terms[i][l]; s += df; sa += (df < 0 ? -df : df); }
      gt[i] = (s == 0) && (sa == 2);
    }
    i64 nnz = 0;
    for (u64 l = 0; l < d; ++l) nnz += (nw[l] > 0);
    for (u64 l = 0; l < d; ++l) {
      if (nw[l] < maxlevel[l]) {
        i64 cnt = 0;
        for (u64 i = 0; i < nd; ++i) cnt += (gt[i] && (nw[l] - terms[i][l] == -1));
        const i64 h3 = nnz + (nw[l] < 1);
        const i64 h4 = 1 + cnt;
        if (h3 == h4) {
          std::vector<i64> c = nw;
          c[l] += 1;
          pterms.push_back(c);
          ptv.push_back(score(c));
        }
      }
    }
  }
  umat out(numele, d);
  for (u64 k = 0; k < numele; ++k)
    for (u64 l = 0; l < d; ++l) out(k, l) = (u64)terms[k][l];
  return out;
}

double outermod::hyplpdf(const vec& hypp) const { /* :89-99 */
  double out = 0;
  if (hyp.size() != hypp.size()) return -std::numeric_limits<double>::infinity();
  for (u64 l = 0; l < d; ++l)
    out += covflist[l]->lpdf(vec(hypp.begin() + hypst[l], hypp.begin() + hypst[l + 1]));
  return out;
}

vec outermod::hyplpdf_grad(const vec& hypp) const { /* :107-119 */
  vec out(hyp.size(), 0.0);
  if (hyp.size() == hypp.size())
    for (u64 l = 0; l < d; ++l) {
      vec g = covflist[l]->lpdf_gradhyp(vec(hypp.begin() + hypst[l], hypp.begin() + hypst[l + 1]));
      std::copy(g.begin(), g.end(), out.begin() + hypst[l]);
    }
  return out;
}

/* ------------------------------------------------------------------ linalg
 * A row block of basemat is copied to a contiguous temporary exactly as
 * `basemat.rows(startind,endind)` does in the reference (linalg.cpp:119,325). */

static void copy_rows(mat& dst, const mat& src, u64 s, u64 e) {
  const u64 n = e - s + 1;
  dst.set_size(n, src.nc);
  for (u64 j = 0; j < src.nc; ++j) std::memcpy(dst.col(j), src.col(j) + s, n * sizeof(double));
}

static inline void colmul(double* t, const double* c, u64 n) { for (u64 i = 0; i < n; ++i) t[i] *= c[i]; }

/* domult_ body for one k (linalg.cpp:70-75) */
static inline void term_product(double* temp, double init, const umat& terms, u64 k,
                                const std::vector<u64>& knotptst, const mat& bm, u64 n, i64 skip = -1) {
  for (u64 i = 0; i < n; ++i) temp[i] = init;
  for (u64 l = 0; l < terms.nc; ++l)
    if (terms(k, l) > 0 && (i64)l != skip) colmul(temp, bm.col(knotptst[l] + terms(k, l)), n);
}
static inline void term_product_from(double* temp, const double* b, const umat& terms, u64 k,
                                     const std::vector<u64>& knotptst, const mat& bm, u64 n, i64 skip = -1) {
  std::memcpy(temp, b, n * sizeof(double));
  for (u64 l = 0; l < terms.nc; ++l)
    if (terms(k, l) > 0 && (i64)l != skip) colmul(temp, bm.col(knotptst[l] + terms(k, l)), n);
}

/* Wide-case helper: omp-for over k with thread-local accumulators summed in
 * thread order (the reference's `omp critical` order is nondeterministic). */
template <class Body, class Merge>
static void wide_over_terms(u64 K, int nthreads, Body body, Merge merge) {
#pragma omp parallel num_threads(nthreads)
  {
    const int T = omp_get_num_threads(), tid = omp_get_thread_num();
    const u64 per = (K + T - 1) / T, k0 = std::min(K, per * tid), k1 = std::min(K, k0 + per);
    body(tid, T, k0, k1);
#pragma omp barrier
#pragma omp single
    merge(T);
  }
}

void prodmm_(vec& out, const umat& terms, const vec& a, const mat& basemat, const vec& basescale,
             const std::vector<u64>& knotptst, const loopvals& lv) { /* linalg.cpp:102-131 */
  const u64 N = basemat.nr, K = terms.nr;
  out.assign(N, 0.0);
  if (lv.vertpl) {
#pragma omp parallel num_threads(lv.nthreads)
    {
      mat bm;
      vec out_, temp_;
#pragma omp for schedule(static)
      for (u64 lcv = 0; lcv < lv.loopsize; ++lcv) {
        const u64 s = lcv * lv.chunksize, e = std::min((lcv + 1) * lv.chunksize - 1, N - 1), n = e - s + 1;
        copy_rows(bm, basemat, s, e);
        out_.assign(n, 0.0);
        temp_.resize(n);
        for (u64 k = 0; k < K; ++k) { /* domult_ :70-75 */
          term_product(temp_.data(), a[k], terms, k, knotptst, bm, n);
          for (u64 i = 0; i < n; ++i) out_[i] += temp_[i];
        }
        std::memcpy(out.data() + s, out_.data(), n * sizeof(double));
      }
    }
  } else { /* domult_ :77-92 */
    std::vector<vec> loc;
    wide_over_terms(K, lv.nthreads,
      [&](int tid, int T, u64 k0, u64 k1) {
#pragma omp critical
        if (loc.empty()) loc.assign(T, vec());
#pragma omp barrier
        vec& o = loc[tid];
        o.assign(N, 0.0);
        vec temp(N);
        for (u64 k = k0; k < k1; ++k) {
          term_product(temp.data(), a[k], terms, k, knotptst, basemat, N);
          for (u64 i = 0; i < N; ++i) o[i] += temp[i];
        }
      },
      [&](int T) { for (int t = 0; t < T; ++t) for (u64 i = 0; i < N; ++i) out[i] += loc[t][i]; });
  }
  for (u64 i = 0; i < N; ++i) out[i] *= basescale[i];
}

void tprodmm_(vec& out, const umat& terms, const vec& a, const mat& basemat, const vec& basescale,
              const std::vector<u64>& knotptst, const loopvals& lv) { /* linalg.cpp:303-355 */
  const u64 N = basemat.nr, K = terms.nr;
  out.assign(K, 0.0);
  vec b(N);
  for (u64 i = 0; i < N; ++i) b[i] = basescale[i] * a[i];
  std::vector<vec> loc;
  if (lv.vertpl) {
#pragma omp parallel num_threads(lv.nthreads)
    {
      const int T = omp_get_num_threads(), tid = omp_get_thread_num();
#pragma omp critical
      if (loc.empty()) loc.assign(T, vec());
#pragma omp barrier
      vec& out_ = loc[tid];
      out_.assign(K, 0.0);
      mat bm;
      vec temp_;
#pragma omp for schedule(static)
      for (u64 lcv = 0; lcv < lv.loopsize; ++lcv) {
        const u64 s = lcv * lv.chunksize, e = std::min((lcv + 1) * lv.chunksize - 1, N - 1), n = e - s + 1;
        copy_rows(bm, basemat, s, e);
        temp_.resize(n);
        for (u64 k = 0; k < K; ++k) { /* dotmultsub_ :286-294 */
          term_product_from(temp_.data(), b.data() + s, terms, k, knotptst, bm, n);
          out_[k] += accu2(temp_.data(), n);
        }
      }
#pragma omp barrier
#pragma omp single
      for (int t = 0; t < T; ++t) for (u64 k = 0; k < K; ++k) out[k] += loc[t][k];
    }
  } else {
    wide_over_terms(K, lv.nthreads,
      [&](int, int, u64 k0, u64 k1) {
        vec temp(N);
        for (u64 k = k0; k < k1; ++k) {
          term_product_from(temp.data(), b.data(), terms, k, knotptst, basemat, N);
          out[k] += accu2(temp.data(), N); /* disjoint k per thread */
        }
      },
      [&](int) {});
  }
}

/* domultgesub_ for a row block (linalg.cpp:139-163) */
static void domultge_block(double* out, mat& outge, const vec& a, const umat& terms,
                           const std::vector<u64>& knotptst, const mat& bm, const mat& bmge,
                           const std::vector<u64>& gest, const std::vector<u64>& hypmatch,
                           u64 n, u64 k0, u64 k1) {
  vec temp(n), tempalt(n);
  const u64 H = gest.size() - 1;
  for (u64 k = k0; k < k1; ++k) {
    term_product(temp.data(), a[k], terms, k, knotptst, bm, n);
    for (u64 i = 0; i < n; ++i) out[i] += temp[i];
    for (u64 l = 0; l < H; ++l) {
      if (terms(k, hypmatch[l]) > 0) {
        term_product(tempalt.data(), a[k], terms, k, knotptst, bm, n, (i64)hypmatch[l]);
        const double* g1 = bmge.col(gest[l] + terms(k, hypmatch[l]));
        const double* g0 = bmge.col(gest[l]);
        double* og = outge.col(l);
        for (u64 i = 0; i < n; ++i) og[i] += tempalt[i] * g1[i] - temp[i] * g0[i];
      }
    }
  }
}

void prodmmge_(vec& out, mat& outge, const umat& terms, const vec& a, const mat& basemat,
               const vec& basescale, const std::vector<u64>& knotptst, const mat& basematge,
               const std::vector<u64>& gest, const std::vector<u64>& hypmatch, const loopvals& lv) { /* :225-277 */
  const u64 N = basemat.nr, K = terms.nr, H = gest.size() - 1;
  out.assign(N, 0.0);
  outge.set_size(N, H);
  outge.zeros();
  if (lv.vertpl) {
#pragma omp parallel num_threads(lv.nthreads)
    {
      mat bm, bmge, outge_;
      vec out_;
#pragma omp for schedule(static)
      for (u64 lcv = 0; lcv < lv.loopsize; ++lcv) {
        const u64 s = lcv * lv.chunksize, e = std::min((lcv + 1) * lv.chunksize - 1, N - 1), n = e - s + 1;
        copy_rows(bm, basemat, s, e);
        copy_rows(bmge, basematge, s, e);
        out_.assign(n, 0.0);
        outge_.set_size(n, H);
        outge_.zeros();
        domultge_block(out_.data(), outge_, a, terms, knotptst, bm, bmge, gest, hypmatch, n, 0, K);
        std::memcpy(out.data() + s, out_.data(), n * sizeof(double));
        for (u64 l = 0; l < H; ++l) std::memcpy(outge.col(l) + s, outge_.col(l), n * sizeof(double));
      }
    }
  } else {
    std::vector<vec> loco;
    std::vector<mat> locg;
    wide_over_terms(K, lv.nthreads,
      [&](int tid, int T, u64 k0, u64 k1) {
#pragma omp critical
        if (loco.empty()) { loco.assign(T, vec()); locg.assign(T, mat()); }
#pragma omp barrier
        loco[tid].assign(N, 0.0);
        locg[tid].set_size(N, H);
        locg[tid].zeros();
        domultge_block(loco[tid].data(), locg[tid], a, terms, knotptst, basemat, basematge, gest, hypmatch, N, k0, k1);
      },
      [&](int T) {
        for (int t = 0; t < T; ++t) {
          for (u64 i = 0; i < N; ++i) out[i] += loco[t][i];
          for (u64 i = 0; i < N * H; ++i) outge.a[i] += locg[t].a[i];
        }
      });
  }
  for (u64 l = 0; l < H; ++l) { /* :273-274 */
    const double* g0 = basematge.col(gest[l]);
    double* og = outge.col(l);
    for (u64 i = 0; i < N; ++i) og[i] += g0[i] * out[i];
  }
  for (u64 i = 0; i < N; ++i) out[i] *= basescale[i];
  for (u64 l = 0; l < H; ++l) { double* og = outge.col(l); for (u64 i = 0; i < N; ++i) og[i] *= basescale[i]; }
}

/* dotmultgesub_ for a row block (linalg.cpp:364-386) */
static void dotmultge_block(double* out, mat& outge, const double* b, const umat& terms,
                            const std::vector<u64>& knotptst, const mat& bm, const mat& bmge,
                            const std::vector<u64>& gest, const std::vector<u64>& hypmatch,
                            u64 n, u64 k0, u64 k1) {
  vec temp(n), tempalt(n);
  const u64 H = gest.size() - 1;
  for (u64 k = k0; k < k1; ++k) {
    term_product_from(temp.data(), b, terms, k, knotptst, bm, n);
    out[k] += accu2(temp.data(), n);
    for (u64 l = 0; l < H; ++l) {
      if (terms(k, hypmatch[l]) > 0) {
        term_product_from(tempalt.data(), b, terms, k, knotptst, bm, n, (i64)hypmatch[l]);
        outge(k, l) += dot_arma(tempalt.data(), bmge.col(gest[l] + terms(k, hypmatch[l])), n);
      } else {
        outge(k, l) += dot_arma(temp.data(), bmge.col(gest[l]), n);
      }
    }
  }
}

void tprodmmge_(vec& out, mat& outge, const umat& terms, const vec& a, const mat& basemat,
                const vec& basescale, const std::vector<u64>& knotptst, const mat& basematge,
                const std::vector<u64>& gest, const std::vector<u64>& hypmatch, const loopvals& lv) { /* :394-471 */
  const u64 N = basemat.nr, K = terms.nr, H = gest.size() - 1;
  vec b(N);
  for (u64 i = 0; i < N; ++i) b[i] = basescale[i] * a[i];
  out.assign(K, 0.0);
  outge.set_size(K, H);
  outge.zeros();
  if (lv.vertpl) {
    std::vector<vec> loco;
    std::vector<mat> locg;
#pragma omp parallel num_threads(lv.nthreads)
    {
      const int T = omp_get_num_threads(), tid = omp_get_thread_num();
#pragma omp critical
      if (loco.empty()) { loco.assign(T, vec()); locg.assign(T, mat()); }
#pragma omp barrier
      loco[tid].assign(K, 0.0);
      locg[tid].set_size(K, H);
      locg[tid].zeros();
      mat bm, bmge;
#pragma omp for schedule(static)
      for (u64 lcv = 0; lcv < lv.loopsize; ++lcv) {
        const u64 s = lcv * lv.chunksize, e = std::min((lcv + 1) * lv.chunksize - 1, N - 1), n = e - s + 1;
        copy_rows(bm, basemat, s, e);
        copy_rows(bmge, basematge, s, e);
        dotmultge_block(loco[tid].data(), locg[tid], b.data() + s, terms, knotptst, bm, bmge, gest, hypmatch, n, 0, K);
      }
#pragma omp barrier
#pragma omp single
      for (int t = 0; t < T; ++t) {
        for (u64 k = 0; k < K; ++k) out[k] += loco[t][k];
        for (u64 i = 0; i < K * H; ++i) outge.a[i] += locg[t].a[i];
      }
    }
  } else {
    wide_over_terms(K, lv.nthreads,
      [&](int, int, u64 k0, u64 k1) {
        dotmultge_block(out.data(), outge, b.data(), terms, knotptst, basemat, basematge, gest, hypmatch, N, k0, k1);
      },
      [&](int) {});
  }
}

void prodmm_mat_(mat& out, const umat& terms, const mat& a, const mat& basemat, const vec& basescale,
                 const std::vector<u64>& knotptst, const loopvals& lv) { /* linalg.cpp:527-557, :481-519 */
  const u64 N = basemat.nr, K = terms.nr, C = a.nc;
  out.set_size(N, C);
  out.zeros();
  if (lv.vertpl) {
#pragma omp parallel num_threads(lv.nthreads)
    {
      mat bm, out_;
      vec temp_;
#pragma omp for schedule(static)
      for (u64 lcv = 0; lcv < lv.loopsize; ++lcv) {
        const u64 s = lcv * lv.chunksize, e = std::min((lcv + 1) * lv.chunksize - 1, N - 1), n = e - s + 1;
        copy_rows(bm, basemat, s, e);
        out_.set_size(n, C);
        out_.zeros();
        temp_.resize(n);
        for (u64 k = 0; k < K; ++k) { /* out += temp * a.row(k) :496 */
          term_product(temp_.data(), 1.0, terms, k, knotptst, bm, n);
          for (u64 c = 0; c < C; ++c) { const double ak = a(k, c); double* oc = out_.col(c); for (u64 i = 0; i < n; ++i) oc[i] += temp_[i] * ak; }
        }
        for (u64 c = 0; c < C; ++c) std::memcpy(out.col(c) + s, out_.col(c), n * sizeof(double));
      }
    }
  } else {
    std::vector<mat> loc;
    wide_over_terms(K, lv.nthreads,
      [&](int tid, int T, u64 k0, u64 k1) {
#pragma omp critical
        if (loc.empty()) loc.assign(T, mat());
#pragma omp barrier
        loc[tid].set_size(N, C);
        loc[tid].zeros();
        vec temp(N);
        for (u64 k = k0; k < k1; ++k) {
          term_product(temp.data(), 1.0, terms, k, knotptst, basemat, N);
          for (u64 c = 0; c < C; ++c) { const double ak = a(k, c); double* oc = loc[tid].col(c); for (u64 i = 0; i < N; ++i) oc[i] += temp[i] * ak; }
        }
      },
      [&](int T) { for (int t = 0; t < T; ++t) for (u64 i = 0; i < N * C; ++i) out.a[i] += loc[t].a[i]; });
  }
  for (u64 c = 0; c < C; ++c) { double* oc = out.col(c); for (u64 i = 0; i < N; ++i) oc[i] *= basescale[i]; }
}

void tprodmm_mat_(mat& out, const umat& terms, const mat& a, const mat& basemat, const vec& basescale,
                  const std::vector<u64>& knotptst, const loopvals& lv) { /* linalg.cpp:583-637, :567-575 */
  const u64 N = basemat.nr, K = terms.nr, C = a.nc;
  out.set_size(K, C);
  out.zeros();
  mat b(N, C);
  for (u64 c = 0; c < C; ++c) for (u64 i = 0; i < N; ++i) b(i, c) = a(i, c) * basescale[i];
  if (lv.vertpl) {
    std::vector<mat> loc;
#pragma omp parallel num_threads(lv.nthreads)
    {
      const int T = omp_get_num_threads(), tid = omp_get_thread_num();
#pragma omp critical
      if (loc.empty()) loc.assign(T, mat());
#pragma omp barrier
      loc[tid].set_size(K, C);
      loc[tid].zeros();
      mat bm;
      vec temp_;
#pragma omp for schedule(static)
      for (u64 lcv = 0; lcv < lv.loopsize; ++lcv) {
        const u64 s = lcv * lv.chunksize, e = std::min((lcv + 1) * lv.chunksize - 1, N - 1), n = e - s + 1;
        copy_rows(bm, basemat, s, e);
        temp_.resize(n);
        for (u64 k = 0; k < K; ++k) { /* out.row(k) += temp.t() * b_ :574 */
          term_product(temp_.data(), 1.0, terms, k, knotptst, bm, n);
          for (u64 c = 0; c < C; ++c) loc[tid](k, c) += dot_seq(temp_.data(), b.col(c) + s, n);
        }
      }
#pragma omp barrier
#pragma omp single
      for (int t = 0; t < T; ++t) for (u64 i = 0; i < K * C; ++i) out.a[i] += loc[t].a[i];
    }
  } else {
    wide_over_terms(K, lv.nthreads,
      [&](int, int, u64 k0, u64 k1) {
        vec temp(N);
        for (u64 k = k0; k < k1; ++k) {
          term_product(temp.data(), 1.0, terms, k, knotptst, basemat, N);
          for (u64 c = 0; c < C; ++c) out(k, c) += dot_seq(temp.data(), b.col(c), N);
        }
      },
      [&](int) {});
  }
}

void getm_(mat& out, const umat& terms, const mat& basemat, const vec& basescale,
           const std::vector<u64>& knotptst, const loopvals& lv) { /* linalg.cpp:685-715, :647-677 */
  const u64 N = basemat.nr, K = terms.nr;
  out.set_size(N, K);
#pragma omp parallel for num_threads(lv.nthreads) schedule(static)
  for (u64 k = 0; k < K; ++k) {
    term_product(out.col(k), 1.0, terms, k, knotptst, basemat, N);
    double* oc = out.col(k);
    for (u64 i = 0; i < N; ++i) oc[i] *= basescale[i];
  }
}

/* ------------------------------------------------------------------ outerbase */

outerbase::outerbase(const outermod& om_, const mat& xp_, bool dograd_) : om(om_), xp(xp_) { /* :459-483 */
  dograd = dograd_;
  n_row = xp.nr;
  nthreads = omp_get_num_procs();
  build();
}

void outerbase::setvals_() { /* :492-501 */
  d = om.d; hypmatch = om.hypmatch; hypst = om.hypst; gest = om.gest; knotptst = om.knotptst;
  n_hyp = om.hypmatch.size();
}

void outerbase::setloopvals_() { /* :504-513 */
  const u64 maxchunk = 1 + (2048 / nthreads);
  const u64 minchunk = 32;
  chunksize = std::max(minchunk, std::min(maxchunk, n_row / (4 * nthreads) + 1));
  loopsize = (n_row + chunksize - 1) / chunksize;
  vertpl = loopsize > 20;
}

void outerbase::setsizes_() { /* :521-539 */
  basemat.set_size(xp.nr, om.knotpt.size());
  basemat_gradhyp.set_size(xp.nr, om.knotptstge[d]);
  basematsq.set_size(xp.nr, om.knotpt.size());
  basematsq_gradhyp.set_size(xp.nr, om.knotptstge[d]);
  basescalemat.set_size(xp.nr, d);
  basescale.assign(xp.nr, 1.0);
  basescalesq.assign(xp.nr, 0.0);
}

void outerbase::build() { /* :547-626 -- row-chunked; the short branch's race (:600-607) is not reproduced */
  setvals_();
  setsizes_();
  setloopvals_();
  const u64 N = n_row;
#pragma omp parallel num_threads((int)nthreads)
  {
    mat R;
    std::vector<mat> Rt;
#pragma omp for schedule(static)
    for (u64 j = 0; j < loopsize; ++j) {
      const u64 s = j * chunksize, e = std::min((j + 1) * chunksize - 1, N - 1), n = e - s + 1;
      for (u64 k = 0; k < d; ++k) {
        const u64 lenh = knotptst[k + 1] - knotptst[k];
        if (dograd) om.buildob(R, Rt, xp.col(k) + s, n, k);
        else om.buildob(R, xp.col(k) + s, n, k);
        for (u64 i = 0; i < n; ++i) {
          basescalemat(s + i, k) = R(i, 0);
          basescale[s + i] *= R(i, 0);
          R(i, 0) = 1.0;
        }
        for (u64 c = 0; c < lenh; ++c)
          for (u64 i = 0; i < n; ++i) {
            basemat(s + i, knotptst[k] + c) = R(i, c);
            basematsq(s + i, knotptst[k] + c) = R(i, c) * R(i, c);
          }
        if (dograd)
          for (u64 l = hypst[k]; l < hypst[k + 1]; ++l) {
            const mat& S = Rt[l - hypst[k]];
            for (u64 c = 0; c < lenh; ++c)
              for (u64 i = 0; i < n; ++i) {
                basemat_gradhyp(s + i, gest[l] + c) = S(i, c);
                basematsq_gradhyp(s + i, gest[l] + c) = 2 * (S(i, c) * R(i, c));
              }
          }
      }
      for (u64 i = 0; i < n; ++i) basescalesq[s + i] = basescale[s + i] * basescale[s + i];
    }
  }
}

mat outerbase::getbase(u64 dim) const { /* :634-639 */
  const u64 k = dim - 1, lenh = knotptst[k + 1] - knotptst[k];
  mat out(n_row, lenh);
  for (u64 c = 0; c < lenh; ++c)
    for (u64 i = 0; i < n_row; ++i) out(i, c) = basemat(i, knotptst[k] + c) * basescalemat(i, k);
  return out;
}
mat outerbase::getmat(const umat& terms) const { mat o; getm_(o, terms, basemat, basescale, knotptst, lv()); return o; }
/* getmge_ / dogetmge_, linalg.cpp:724-822, unchunked branch (the chunked one cannot work, see ob_oracle.hpp) */
void getmge_(std::vector<mat>& outge, const umat& terms, const mat& basemat, const vec& basescale, const std::vector<u64>& knotptst,
             const mat& basematge, const std::vector<u64>& gest, const std::vector<u64>& hypmatch) {
  const u64 N = basemat.nr, K = terms.nr, H = gest.size() - 1;
  outge.assign(H, mat(N, K));
  vec tempalt(N);
  for (u64 k = 0; k < K; ++k)
    for (u64 l = 0; l < H; ++l) {
      std::fill(tempalt.begin(), tempalt.end(), 1.0);
      for (u64 m = 0; m < terms.nc; ++m)
        if (terms(k, m) > 0 && m != hypmatch[l]) {
          const double* c = basemat.col(knotptst[m] + terms(k, m));
          for (u64 i = 0; i < N; ++i) tempalt[i] *= c[i];
        }
      const double* g = basematge.col(gest[l] + terms(k, hypmatch[l]));
      double* o = outge[l].col(k);
      for (u64 i = 0; i < N; ++i) o[i] = tempalt[i] * g[i];
    }
  for (u64 l = 0; l < H; ++l)
    for (u64 k = 0; k < K; ++k) {
      double* o = outge[l].col(k);
      for (u64 i = 0; i < N; ++i) o[i] *= basescale[i];
    }
}
std::vector<mat> outerbase::getmat_gradhyp(const umat& terms) const { /* :663-669 */
  std::vector<mat> outge;
  getmge_(outge, terms, basemat, basescale, knotptst, basemat_gradhyp, gest, hypmatch);
  return outge;
}
void outerbase::mm(vec& out, const umat& terms, const vec& a) const { prodmm_(out, terms, a, basemat, basescale, knotptst, lv()); }
void outerbase::tmm(vec& out, const umat& terms, const vec& a) const { tprodmm_(out, terms, a, basemat, basescale, knotptst, lv()); }
void outerbase::mm_gradhyp(vec& out, mat& outge, const umat& terms, const vec& a) const {
  prodmmge_(out, outge, terms, a, basemat, basescale, knotptst, basemat_gradhyp, gest, hypmatch, lv());
}
void outerbase::tmm_gradhyp(vec& out, mat& outge, const umat& terms, const vec& a) const {
  tprodmmge_(out, outge, terms, a, basemat, basescale, knotptst, basemat_gradhyp, gest, hypmatch, lv());
}
vec outerbase::sqmm(const umat& terms, const vec& a) const { vec o; prodmm_(o, terms, a, basematsq, basescalesq, knotptst, lv()); return o; }
mat outerbase::sqmm_gradhyp(const umat& terms, const vec& a) const {
  vec o; mat g;
  prodmmge_(o, g, terms, a, basematsq, basescalesq, knotptst, basematsq_gradhyp, gest, hypmatch, lv());
  return g;
}
vec outerbase::sqtmm(const umat& terms, const vec& a) const { vec o; tprodmm_(o, terms, a, basematsq, basescalesq, knotptst, lv()); return o; }
mat outerbase::sqtmmm(const umat& terms, const mat& a) const { mat o; tprodmm_mat_(o, terms, a, basematsq, basescalesq, knotptst, lv()); return o; }
mat outerbase::sqtmm_gradhyp(const umat& terms, const vec& a) const {
  vec o; mat g;
  tprodmmge_(o, g, terms, a, basematsq, basescalesq, knotptst, basematsq_gradhyp, gest, hypmatch, lv());
  return g;
}
vec outerbase::sqcolsums(const umat& terms) const { vec ro(n_row, 1.0); return sqtmm(terms, ro); }
vec outerbase::residvar(const umat& terms) const { /* :889-896, assumes a correlation function */
  vec out = sqmm(terms, om.getvar(terms));
  for (double& v : out) v = 1 - v;
  return out;
}
mat outerbase::residvar_gradhyp(const umat& terms) const { /* :904-922 */
  const vec varc = om.getvar(terms);
  mat outge = sqmm_gradhyp(terms, varc);
  for (double& v : outge.a) v = -v;
  mat l2 = om.getlvar_gradhyp(terms);
  for (u64 h = 0; h < l2.nc; ++h)
    for (u64 k = 0; k < l2.nr; ++k) l2(k, h) *= varc[k];
  mat l3;
  prodmm_mat_(l3, terms, l2, basematsq, basescalesq, knotptst, lv());
  for (u64 i = 0; i < outge.a.size(); ++i) outge.a[i] -= l3.a[i];
  return outge;
}
mat outerbase::sqcolsums_gradhyp(const umat& terms) const { vec ro(n_row, 1.0); return sqtmm_gradhyp(terms, ro); }

/* ------------------------------------------------------------------ lpdf */

double lpdf::paralpdf(const vec& parap) const { /* fit.cpp:133-142 */
  double out = 0;
  if (npara != parap.size()) return -std::numeric_limits<double>::infinity();
  vec t(parap.size());
  for (size_t l = 0; l < parap.size(); ++l) { const double e = parap[l] - para0[l]; t[l] = e * e / paravar[l]; }
  out -= 0.5 * accu2(t.data(), t.size());
  return out;
}

vec lpdf::paralpdf_grad(const vec& parap) const { /* fit.cpp:146-157 */
  vec out(para.size(), 0.0);
  if (npara != parap.size()) return out;
  for (size_t l = 0; l < parap.size(); ++l) out[l] -= (parap[l] - para0[l]) / paravar[l];
  return out;
}

/* ---- small dense algebra in the order of the BLAS / LAPACK stand-ins of oracle/arma_shim (netlib reference BLAS:
 * every output element is one left-to-right sum; solve / inv: elimination with partial pivoting) ---- */
static mat dense_mul(const mat& A, const mat& B) { /* dgemm('N','N') / dgemv('N'): axpy sweeps, exact zeros of B skipped */
  mat C(A.nr, B.nc);
  for (u64 j = 0; j < B.nc; ++j)
    for (u64 l = 0; l < A.nc; ++l) {
      const double t = B(l, j);
      if (t == 0.0) continue;
      double* c = C.col(j);
      const double* a = A.col(l);
      for (u64 i = 0; i < A.nr; ++i) c[i] += t * a[i];
    }
  return C;
}
static vec dense_mul(const mat& A, const vec& b) {
  vec c(A.nr, 0.0);
  for (u64 l = 0; l < A.nc; ++l) {
    if (b[l] == 0.0) continue;
    const double* a = A.col(l);
    for (u64 i = 0; i < A.nr; ++i) c[i] += b[l] * a[i];
  }
  return c;
}
static mat dense_tmul(const mat& A, const mat& B) { /* A^T B, dgemm('T','N'): one sequential dot per element */
  mat C(A.nc, B.nc);
  for (u64 j = 0; j < B.nc; ++j)
    for (u64 i = 0; i < A.nc; ++i) C(i, j) = dot_seq(A.col(i), B.col(j), A.nr);
  return C;
}
static mat dense_solve(const mat& A, const mat& B) {
  if (A.nr != A.nc || A.nr != B.nr) throw std::invalid_argument("solve(): incompatible dimensions");
  const u64 n = A.nr, m = B.nc;
  mat L = A, X = B;
  for (u64 k = 0; k < n; ++k) {
    u64 p = k;
    for (u64 i = k + 1; i < n; ++i) if (std::abs(L(i, k)) > std::abs(L(p, k))) p = i;
    if (L(p, k) == 0.0) throw std::runtime_error("solve(): solution not found");
    if (p != k) { for (u64 j = 0; j < n; ++j) std::swap(L(k, j), L(p, j)); for (u64 j = 0; j < m; ++j) std::swap(X(k, j), X(p, j)); }
    for (u64 i = k + 1; i < n; ++i) {
      const double f = L(i, k) / L(k, k);
      if (f == 0.0) continue;
      for (u64 j = k; j < n; ++j) L(i, j) -= f * L(k, j);
      for (u64 j = 0; j < m; ++j) X(i, j) -= f * X(k, j);
    }
  }
  for (u64 j = 0; j < m; ++j)
    for (u64 ii = n; ii-- > 0;) {
      double sacc = X(ii, j);
      for (u64 c = ii + 1; c < n; ++c) sacc -= L(ii, c) * X(c, j);
      X(ii, j) = sacc / L(ii, ii);
    }
  return X;
}
static mat dense_inv(const mat& A) {
  mat I(A.nr, A.nr);
  for (u64 i = 0; i < A.nr; ++i) I(i, i) = 1.0;
  return dense_solve(A, I);
}

void lpdf::optnewton() { /* fit.cpp:98-131: one Newton step on the full Hessian */
  fullhess = true;
  compute_val = true; compute_grad = true; compute_gradhyp = false; compute_gradpara = false;
  if (coeff.size() != nterms) coeff.assign(nterms, 0.0);
  update(vec(coeff));
  mat h = hess();
  vec r = grad;
  if (!all_finite(h.a) && !all_finite(r)) { val = -std::numeric_limits<double>::infinity(); return; }
  mat rm(r.size(), 1);
  rm.a = r;
  const mat step = dense_solve(h, rm);
  for (u64 i = 0; i < coeff.size(); ++i) coeff[i] += step.a[i];
  compute_gradhyp = true; compute_gradpara = true;
  update(vec(coeff));
  compute_gradhyp = false; compute_gradpara = false;
}

void lpdf::optcg(double tol, unsigned maxepch) { /* fit.cpp:37-96 */
  fullhess = false;
  compute_val = true; compute_grad = true; compute_gradhyp = false; compute_gradpara = false;
  if (coeff.size() != nterms) coeff.assign(nterms, 0.0);
  update(vec(coeff));
  vec m = diaghess();
  cg_iters = 0;
  if (!all_finite(m) && !all_finite(grad)) { val = -std::numeric_limits<double>::infinity(); return; }
  const u64 K = nterms;
  vec rm(K), t(K);
  for (u64 i = 0; i < K; ++i) rm[i] = grad[i] / m[i];
  double num = 0;
  vec p = rm;
  vec q = hessmult(p);
  double beta = 0, denom = 1, alpha = 0, num2 = 0, valo = val, valdiff = 10;
  unsigned k;
  for (k = 0; k < maxepch; k++) {
    for (u64 i = 0; i < K; ++i) t[i] = grad[i] * rm[i];
    num = accu2(t.data(), K);
    if (num < tol && valdiff < tol) break;
    for (u64 i = 0; i < K; ++i) t[i] = q[i] * p[i];
    denom = accu2(t.data(), K);
    alpha = num / denom;
    for (u64 i = 0; i < K; ++i) coeff[i] += alpha * p[i];
    valo = val;
    update(vec(coeff));
    valdiff = val - valo;
    for (u64 i = 0; i < K; ++i) rm[i] = grad[i] / m[i];
    for (u64 i = 0; i < K; ++i) t[i] = (alpha * q[i]) * rm[i];
    num2 = -accu2(t.data(), K);
    beta = num2 / num;
    for (u64 i = 0; i < K; ++i) p[i] = rm[i] + beta * p[i];
    q = hessmult(p);
  }
  cg_iters = k;
  compute_gradhyp = true; compute_gradpara = true;
  update(vec(coeff));
  compute_gradhyp = false; compute_gradpara = false;
}

/* ---- logpr_gauss ---- */
logpr_gauss::logpr_gauss(const outermod& om_, const umat& terms_) : om(om_) { /* :41-58 */
  npara = 1;
  terms = terms_;
  para0 = {6}; paravar = {4};
  nterms = terms.nr;
  para = para0;
  sca = std::exp(para[0]);
  updateom();
}
void logpr_gauss::updateom() { /* :65-68 */
  coeffsd = om.getvar(terms);
  for (double& v : coeffsd) v = std::sqrt(v);
  coefflvarge = om.getlvar_gradhyp(terms);
}
void logpr_gauss::updatepara(const vec& p) { para = p; sca = std::exp(para[0]); } /* :75-78 */
void logpr_gauss::updateterms(const umat& t) { terms = t; nterms = terms.nr; updateom(); } /* :85-90 */
void logpr_gauss::update(const vec& coeff_) { /* :98-106 */
  coeff = coeff_;
  const u64 K = coeff.size(), H = coefflvarge.nc;
  stdresid.resize(K);
  for (u64 i = 0; i < K; ++i) stdresid[i] = coeff[i] / (coeffsd[i] * sca);
  vec t(K), t2(K);
  if (compute_val) {
    for (u64 i = 0; i < K; ++i) { t[i] = stdresid[i] * stdresid[i]; t2[i] = std::log(coeffsd[i] * sca); }
    val = -0.5 * accu2(t.data(), K) - accu2(t2.data(), K);
  }
  if (compute_gradhyp) {
    gradhyp.assign(H, 0.0);
    for (u64 i = 0; i < K; ++i) t[i] = stdresid[i] * stdresid[i] - 1;
    for (u64 h = 0; h < H; ++h) {
      for (u64 i = 0; i < K; ++i) t2[i] = 0.5 * coefflvarge(i, h);
      gradhyp[h] = dot_seq(t2.data(), t.data(), K);
    }
  }
  if (compute_gradpara) {
    for (u64 i = 0; i < K; ++i) t[i] = stdresid[i] * stdresid[i];
    gradpara = {accu2(t.data(), K) - double(coeffsd.size())};
  }
  if (compute_grad) {
    grad.resize(K);
    for (u64 i = 0; i < K; ++i) grad[i] = -1. * stdresid[i] / (coeffsd[i] * sca);
  }
}
vec logpr_gauss::hessmult(const vec& g) { /* :113-115 */
  vec o(g.size());
  for (u64 i = 0; i < g.size(); ++i) { const double s = coeffsd[i] * sca; o[i] = g[i] / (s * s); }
  return o;
}
vec logpr_gauss::diaghess() { /* :122-124 */
  vec o(coeffsd.size());
  for (u64 i = 0; i < o.size(); ++i) { const double s = coeffsd[i] * sca; o[i] = 1. / (s * s); }
  return o;
}
mat logpr_gauss::diaghessgradhyp() { /* :131-135 */
  mat o = coefflvarge;
  for (u64 h = 0; h < o.nc; ++h)
    for (u64 i = 0; i < o.nr; ++i) { const double s = coeffsd[i] * sca; o(i, h) = -(o(i, h) / (s * s)); }
  return o;
}
mat logpr_gauss::diaghessgradpara() { /* :143-145 */
  mat o(coeffsd.size(), 1);
  for (u64 i = 0; i < o.nr; ++i) { const double s = coeffsd[i] * sca; o(i, 0) = -2. / (s * s); }
  return o;
}

mat logpr_gauss::hess() { /* :153-158 */
  mat h(nterms, nterms);
  for (u64 i = 0; i < nterms; ++i) { const double sd = coeffsd[i] * sca; h(i, i) = 1. / (sd * sd); }
  return h;
}
std::vector<mat> logpr_gauss::hessgradhyp() { /* :165-173 */
  std::vector<mat> o(coefflvarge.nc, mat(nterms, nterms));
  for (u64 l = 0; l < coefflvarge.nc; ++l)
    for (u64 i = 0; i < nterms; ++i) { const double sd = coeffsd[i] * sca; o[l](i, i) = -(coefflvarge(i, l) / (sd * sd)); }
  return o;
}
std::vector<mat> logpr_gauss::hessgradpara() { /* :181-186 */
  std::vector<mat> o(1, mat(nterms, nterms));
  for (u64 i = 0; i < nterms; ++i) { const double sd = coeffsd[i] * sca; o[0](i, i) = -2. / (sd * sd); }
  return o;
}

/* ---- loglik_std ---- */
loglik_std::loglik_std(const outermod& om_, const umat& terms_, const vec& y_, const mat& x_)
    : om(om_), ob(om_, x_, true), y(y_), x(x_) { /* :41-59 */
  terms = terms_;
  npara = 1;
  basismat = ob.getmat(terms);
  basismat_gradhyp = ob.getmat_gradhyp(terms);
  para0 = {std::log(0.01 * arma_var(y.data(), y.size()))};
  paravar = {1};
  para = para0;
  nterms = terms.nr;
}
void loglik_std::updateom() { /* :67-71 */
  ob.build();
  basismat = ob.getmat(terms);
  basismat_gradhyp = ob.getmat_gradhyp(terms);
}
void loglik_std::updatepara(const vec& p) { para = p; } /* :78-80 */
void loglik_std::updateterms(const umat& t) { /* :87-92 */
  terms = t;
  nterms = terms.nr;
  basismat = ob.getmat(terms);
  basismat_gradhyp = ob.getmat_gradhyp(terms);
}
void loglik_std::update(const vec& coeff_) { /* :100-123 */
  coeff = coeff_;
  const u64 N = y.size(), H = basismat_gradhyp.size();
  yhat = dense_mul(basismat, coeff);
  mat yhatge;
  if (compute_gradhyp) {
    yhatge.set_size(N, H);
    for (u64 l = 0; l < H; ++l) {
      const vec c = dense_mul(basismat_gradhyp[l], coeff);
      std::copy(c.begin(), c.end(), yhatge.col(l));
    }
  }
  const double e = std::exp(-para[0]);
  vec residtemp(N), residtemp2(N);
  for (u64 i = 0; i < N; ++i) { residtemp[i] = e * (yhat[i] - y[i]); residtemp2[i] = residtemp[i] * residtemp[i]; }
  if (compute_val) val = -0.5 * accu2(residtemp2.data(), N) - double(N) * para[0];
  if (compute_grad) {
    for (u64 i = 0; i < N; ++i) residtemp[i] = (-e) * residtemp[i];
    ob.tmm(grad, terms, residtemp);
    if (compute_gradhyp) {
      gradhyp.assign(H, 0.0);
      for (u64 h = 0; h < H; ++h) gradhyp[h] = dot_seq(residtemp.data(), yhatge.col(h), N);
    }
    if (compute_gradpara) gradpara = {accu2(residtemp2.data(), N) - double(N)};
  }
}
vec loglik_std::hessmult(const vec& g) { /* :130-133 */
  const vec t = dense_mul(basismat, g);
  vec lh(basismat.nc);
  const double c = std::exp(-2 * para[0]);
  for (u64 k = 0; k < lh.size(); ++k) lh[k] = c * dot_seq(basismat.col(k), t.data(), basismat.nr);
  return lh;
}
static vec colsums_of_squares(const mat& B) { /* sum(square(B), 0): one two-accumulator sum per column */
  vec lh(B.nc), t(B.nr);
  for (u64 k = 0; k < B.nc; ++k) {
    const double* c = B.col(k);
    for (u64 i = 0; i < B.nr; ++i) t[i] = c[i] * c[i];
    lh[k] = accu2(t.data(), B.nr);
  }
  return lh;
}
vec loglik_std::diaghess() { /* :140-143 */
  vec lh = colsums_of_squares(basismat);
  const double c = std::exp(-2 * para[0]);
  for (double& v : lh) v = c * v;
  return lh;
}
mat loglik_std::diaghessgradhyp() { /* :150-155 */
  const u64 N = basismat.nr, K = basismat.nc, H = basismat_gradhyp.size();
  mat lh(K, H);
  vec t(N);
  const double c = std::exp(-2 * para[0]);
  for (u64 h = 0; h < H; ++h)
    for (u64 k = 0; k < K; ++k) {
      const double* b = basismat.col(k);
      const double* g = basismat_gradhyp[h].col(k);
      for (u64 i = 0; i < N; ++i) t[i] = g[i] * (2 * b[i]);
      lh(k, h) = c * accu2(t.data(), N);
    }
  return lh;
}
mat loglik_std::diaghessgradpara() { /* :162-165 */
  const vec lh = colsums_of_squares(basismat);
  mat o(lh.size(), 1);
  const double c = -2 * std::exp(-2 * para[0]);
  for (u64 i = 0; i < lh.size(); ++i) o(i, 0) = c * lh[i];
  return o;
}
mat loglik_std::hess() { /* :173-176 */
  mat lh = dense_tmul(basismat, basismat);
  const double c = std::exp(-2 * para[0]);
  for (double& v : lh.a) v = c * v;
  return lh;
}
std::vector<mat> loglik_std::hessgradhyp() { /* :183-195 */
  const u64 K = basismat.nc;
  std::vector<mat> o;
  const double c = std::exp(-2 * para[0]);
  for (const mat& G : basismat_gradhyp) {
    mat S = dense_tmul(basismat, G);
    for (double& v : S.a) v = c * v;
    mat R(K, K);
    for (u64 j = 0; j < K; ++j)
      for (u64 i = 0; i < K; ++i) R(i, j) = S(i, j) + S(j, i);
    o.push_back(std::move(R));
  }
  return o;
}
std::vector<mat> loglik_std::hessgradpara() { /* :202-206 */
  mat lh = dense_tmul(basismat, basismat);
  const double c = -2 * std::exp(-2 * para[0]);
  for (double& v : lh.a) v = c * v;
  return {lh};
}

/* ---- loglik_gauss ---- */
loglik_gauss::loglik_gauss(const outermod& om_, const umat& terms_, const vec& y_, const mat& x_)
    : om(om_), ob(om_, x_, true), y(y_), x(x_) { /* :41-59 */
  terms = terms_;
  npara = 1;
  para0 = {std::log(0.01 * arma_var(y.data(), y.size()))};
  paravar = {1};
  para = para0;
  obssd.assign(y.size(), std::exp(para[0]));
  obsvar.assign(y.size(), std::exp(2. * para[0]));
  nterms = terms.nr;
  lobsvar.resize(y.size());
  for (u64 i = 0; i < y.size(); ++i) lobsvar[i] = std::log(obsvar[i]);
}
void loglik_gauss::updateom() { ob.build(); }          /* :67-69 */
void loglik_gauss::setnthreads(int k) { ob.nthreads = k; } /* :76-78 */
void loglik_gauss::updatepara(const vec& p) {          /* :86-91 */
  para = p;
  obssd.assign(y.size(), std::exp(para[0]));
  obsvar.assign(y.size(), std::exp(2. * para[0]));
  for (u64 i = 0; i < y.size(); ++i) lobsvar[i] = std::log(obsvar[i]);
}
void loglik_gauss::updateterms(const umat& t) { terms = t; nterms = terms.nr; } /* :98-101 */
void loglik_gauss::update(const vec& coeff_) { /* :110-130 */
  coeff = coeff_;
  const u64 N = y.size();
  if (compute_gradhyp) ob.mm_gradhyp(yhat, yhatge, terms, coeff);
  else ob.mm(yhat, terms, coeff);
  residtemp.resize(N); residtemp2.resize(N);
  for (u64 i = 0; i < N; ++i) { residtemp[i] = (yhat[i] - y[i]) / obssd[i]; residtemp2[i] = residtemp[i] * residtemp[i]; }
  if (compute_val) {
    vec t(N);
    for (u64 i = 0; i < N; ++i) t[i] = std::log(obssd[i]);
    val = -0.5 * accu2(residtemp2.data(), N) - accu2(t.data(), N);
  }
  if (compute_grad) {
    for (u64 i = 0; i < N; ++i) residtemp[i] = -1. * (residtemp[i] / obssd[i]);
    ob.tmm(grad, terms, residtemp);
    if (compute_gradhyp) {
      gradhyp.assign(ob.n_hyp, 0.0);
      for (u64 h = 0; h < ob.n_hyp; ++h) gradhyp[h] = dot_seq(residtemp.data(), yhatge.col(h), N);
    }
    if (compute_gradpara) gradpara = {accu2(residtemp2.data(), N) - double(N)};
  }
}
vec loglik_gauss::hessmult(const vec& g) { /* :137-145 */
  ob.mm(yhattemp, terms, g);
  for (u64 i = 0; i < yhattemp.size(); ++i) { yhattemp[i] /= obssd[i]; yhattemp[i] /= obssd[i]; }
  ob.tmm(gradtemp, terms, yhattemp);
  return gradtemp;
}
vec loglik_gauss::diaghess() { /* :154-157 */
  vec lh = ob.sqcolsums(terms);
  const double c = std::exp(-2 * para[0]);
  for (double& v : lh) v = c * v;
  return lh;
}
mat loglik_gauss::diaghessgradhyp() { /* :165-168 */
  mat h = ob.sqcolsums_gradhyp(terms);
  const double c = std::exp(-2 * para[0]);
  for (double& v : h.a) v = c * v;
  return h;
}
mat loglik_gauss::diaghessgradpara() { /* :176-179 */
  vec lh = ob.sqcolsums(terms);
  mat o(lh.size(), 1);
  const double c = -2 * std::exp(-2 * para[0]);
  for (u64 i = 0; i < lh.size(); ++i) o(i, 0) = c * lh[i];
  return o;
}

/* ---- loglik_gda ---- */
loglik_gda::loglik_gda(const outermod& om_, const umat& terms_, const vec& y_, const mat& x_)
    : om(om_), ob(om_, x_, true), y(y_), x(x_) { /* :47-68 */
  terms = terms_;
  npara = 2;
  para0 = {0.5 * std::log(0.01 * arma_var(y.data(), y.size())), 0.0};
  paravar = {4, 4};
  para = para0;
  buildstd();
  nterms = terms.nr;
}
void loglik_gda::setnthreads(int k) { ob.nthreads = k; }                                /* :75-77 */
void loglik_gda::updateom() { ob.build(); if (doda) redostd = true; }                    /* :84-87 */
void loglik_gda::updatepara(const vec& p) { para = p; redostd = true; }                   /* :94-97 */
void loglik_gda::updateterms(const umat& t) { terms = t; nterms = terms.nr; if (doda) redostd = true; } /* :104-108 */
void loglik_gda::buildstd() { /* :217-236 */
  if (redostd) {
    const u64 N = y.size();
    vec obsvar(N, std::exp(2 * para[0]));
    const vec rterms = ob.residvar(terms);
    if (doda) for (u64 i = 0; i < N; ++i) obsvar[i] += std::exp(2 * para[1]) * rterms[i];
    obssd.resize(N);
    for (u64 i = 0; i < N; ++i) obssd[i] = std::sqrt(obsvar[i]);
    if (doda) {
      obssd_gradhyp = ob.residvar_gradhyp(terms);
      for (u64 h = 0; h < obssd_gradhyp.nc; ++h)
        for (u64 i = 0; i < N; ++i) obssd_gradhyp(i, h) *= (std::exp(2 * para[1]) * 0.5) / obssd[i];
    }
    obssd_gradpara = mat(N, 2);
    for (u64 i = 0; i < N; ++i) {
      obssd_gradpara(i, 0) = std::exp(2 * para[0]) / obssd[i];
      obssd_gradpara(i, 1) = doda ? std::exp(2 * para[1]) * rterms[i] / obssd[i] : 0.0;
    }
  }
  redostd = false;
}
void loglik_gda::update(const vec& coeff_) { /* :116-149 */
  coeff = coeff_;
  const u64 N = y.size();
  if (compute_gradhyp) ob.mm_gradhyp(yhat, yhatge, terms, coeff);
  else ob.mm(yhat, terms, coeff);
  buildstd();
  residtemp.resize(N); residtemp2.resize(N);
  for (u64 i = 0; i < N; ++i) { residtemp[i] = (yhat[i] - y[i]) / obssd[i]; residtemp2[i] = residtemp[i] * residtemp[i]; }
  if (compute_val) {
    vec t(N);
    for (u64 i = 0; i < N; ++i) t[i] = std::log(obssd[i]);
    val = -0.5 * accu2(residtemp2.data(), N) - accu2(t.data(), N);
  }
  if (compute_grad) {
    vec inv(N);
    for (u64 i = 0; i < N; ++i) { residtemp[i] = -1. * (residtemp[i] / obssd[i]); residtemp2[i] /= obssd[i]; inv[i] = 1 / obssd[i]; }
    ob.tmm(grad, terms, residtemp);
    if (compute_gradhyp) {
      gradhyp.assign(ob.n_hyp, 0.0);
      for (u64 h = 0; h < ob.n_hyp; ++h) {
        gradhyp[h] = dot_seq(residtemp.data(), yhatge.col(h), N);
        if (doda) {
          gradhyp[h] += dot_seq(residtemp2.data(), obssd_gradhyp.col(h), N);
          gradhyp[h] -= dot_seq(inv.data(), obssd_gradhyp.col(h), N);
        }
      }
    }
    if (compute_gradpara) {
      gradpara.assign(2, 0.0);
      for (u64 c = 0; c < 2; ++c) {
        gradpara[c] = dot_seq(residtemp2.data(), obssd_gradpara.col(c), N);
        gradpara[c] -= dot_seq(inv.data(), obssd_gradpara.col(c), N);
      }
    }
  }
}
vec loglik_gda::hessmult(const vec& g) { /* :156-164 */
  ob.mm(yhattemp, terms, g);
  for (u64 i = 0; i < yhattemp.size(); ++i) { yhattemp[i] /= obssd[i]; yhattemp[i] /= obssd[i]; }
  ob.tmm(gradtemp, terms, yhattemp);
  return gradtemp;
}
vec loglik_gda::diaghess() { /* :172-175 */
  vec t(obssd.size());
  for (u64 i = 0; i < t.size(); ++i) t[i] = 1 / (obssd[i] * obssd[i]);
  return ob.sqtmm(terms, t);
}
mat loglik_gda::diaghessgradhyp() { /* :182-196 */
  const u64 N = obssd.size();
  vec temp(N);
  for (u64 i = 0; i < N; ++i) temp[i] = 1 / (obssd[i] * obssd[i]);
  mat lh = ob.sqtmm_gradhyp(terms, temp);
  for (u64 i = 0; i < N; ++i) temp[i] *= -2 / obssd[i];
  if (doda) {
    mat t2 = obssd_gradhyp;
    for (u64 h = 0; h < t2.nc; ++h)
      for (u64 i = 0; i < N; ++i) t2(i, h) *= temp[i];
    const mat add = ob.sqtmmm(terms, t2);
    for (u64 i = 0; i < lh.a.size(); ++i) lh.a[i] += add.a[i];
  }
  return lh;
}
mat loglik_gda::diaghessgradpara() { /* :203-211 */
  const u64 N = obssd.size();
  mat t2 = obssd_gradpara;
  for (u64 c = 0; c < t2.nc; ++c)
    for (u64 i = 0; i < N; ++i) t2(i, c) *= (1 / (obssd[i] * obssd[i])) * (-2 / obssd[i]);
  return ob.sqtmmm(terms, t2);
}

/* ---- pred_gda ---- */
pred_gda::pred_gda(const loglik_gda& loglik) : om(loglik.om), para(loglik.para), terms(loglik.terms) { /* :249-266 */
  nthreads = (int)loglik.ob.nthreads;
  doda = loglik.doda;
  coeff = loglik.coeff;
  if (coeff.size() != terms.nr) coeff.assign(terms.nr, 0.0);
  ob.reset(new outerbase(om, loglik.x, false));
  ob->nthreads = nthreads;
  if (!loglik.didnotothess) {
    coeffvar.resize(loglik.totdiaghess.size());
    for (u64 i = 0; i < coeffvar.size(); ++i) coeffvar[i] = 1 / loglik.totdiaghess[i];
  } else coeffvar.assign(coeff.size(), 0.0);
}
void pred_gda::update(const mat& x_) { ob.reset(new outerbase(om, x_, false)); ob->nthreads = nthreads; } /* :268-272 */
vec pred_gda::mean() const { vec o; ob->mm(o, terms, coeff); return o; }                                   /* :273-275 */
vec pred_gda::var() const { /* :276-282 */
  vec out = ob->sqmm(terms, coeffvar);
  const vec rv = doda ? ob->residvar(terms) : vec();
  for (u64 i = 0; i < out.size(); ++i) { out[i] += std::exp(2 * para[0]); if (doda) out[i] += std::exp(2 * para[1]) * rv[i]; }
  return out;
}

/* ---- lpdfvec ---- */
lpdfvec::lpdfvec(lpdf& a, lpdf& b) { /* fit.cpp:174-200 */
  parasrt.assign(2, 0); paraend.assign(2, 0);
  terms = a.terms;
  nterms = a.nterms;
  lpdflist.push_back(&a);
  parasrt[0] = 0; paraend[0] = a.npara - 1;
  lpdflist.push_back(&b);
  parasrt[1] = paraend[0] + 1; paraend[1] = parasrt[1] + b.npara - 1;
  para.assign(1 + paraend[1], 0.0);
  std::copy(a.para.begin(), a.para.end(), para.begin() + parasrt[0]);
  std::copy(b.para.begin(), b.para.end(), para.begin() + parasrt[1]);
  npara = 0; /* the reference never sets lpdfvec::npara; paralpdf is overridden */
  npara = (unsigned)para.size();
}
void lpdfvec::setnthreads(int k) { for (lpdf* l : lpdflist) l->setnthreads(k); }
void lpdfvec::updateom() { for (lpdf* l : lpdflist) l->updateom(); redohess = true; } /* :207-210 */
void lpdfvec::updatepara(const vec& p) { /* :219-228 */
  for (u64 c = 0; c < lpdflist.size(); ++c) {
    std::copy(p.begin() + parasrt[c], p.begin() + paraend[c] + 1, para.begin() + parasrt[c]);
    lpdflist[c]->updatepara(vec(p.begin() + parasrt[c], p.begin() + paraend[c] + 1));
  }
  redohess = true;
}
void lpdfvec::updateterms(const umat& t) { /* :237-244 */
  terms = t;
  for (lpdf* l : lpdflist) { l->updateterms(t); nterms = l->nterms; }
  redohess = true;
}
void lpdfvec::buildhess() { /* :252-301 */
  if (redohess) {
    diaghessv = diaghess_();
    settotdiaghess(diaghessv);
    if (domargadj) {
      diaghessgradhypv = diaghessgradhyp_();
      diaghessgradparav = diaghessgradpara_();
      const u64 K = diaghessv.size();
      vec t(K);
      for (u64 i = 0; i < K; ++i) t[i] = std::log(diaghessv[i]);
      val_margadj = -0.5 * accu2(t.data(), K);
      gradhyp_margadj.assign(diaghessgradhypv.nc, 0.0);
      for (u64 h = 0; h < diaghessgradhypv.nc; ++h) {
        for (u64 i = 0; i < K; ++i) t[i] = diaghessgradhypv(i, h) / diaghessv[i];
        gradhyp_margadj[h] = -0.5 * accu2(t.data(), K);
      }
      gradpara_margadj.assign(diaghessgradparav.nc, 0.0);
      for (u64 h = 0; h < diaghessgradparav.nc; ++h) {
        for (u64 i = 0; i < K; ++i) t[i] = diaghessgradparav(i, h) / diaghessv[i];
        gradpara_margadj[h] = -0.5 * accu2(t.data(), K);
      }
    }
  }
  if (redohess && fullhess) { /* :269-299 */
    hessv = hess_();
    settothess(hessv);
    if (domargadj) {
      hessgradhypv = hessgradhyp_();
      hessgradparav = hessgradpara_();
      const u64 K = hessv.nr;
      vec heigval;
      mat heigvec;
      eig_sym_jacobi(heigval, heigvec, hessv);
      mat hessi(K, K);
      for (u64 j = 0; j < K; ++j)
        for (u64 i = 0; i < K; ++i) hessi(i, j) = heigvec(j, i) / heigval[i];
      hessi = dense_mul(heigvec, hessi);
      vec t(K);
      for (u64 i = 0; i < K; ++i) t[i] = std::log(heigval[i]);
      val_margadj = -0.5 * accu2(t.data(), K);
      vec tm(K * K);
      gradhyp_margadj.assign(hessgradhypv.size(), 0.0);
      for (u64 l = 0; l < hessgradhypv.size(); ++l) {
        for (u64 i = 0; i < K * K; ++i) tm[i] = hessgradhypv[l].a[i] * hessi.a[i];
        gradhyp_margadj[l] = -0.5 * accu2(tm.data(), K * K);
      }
      gradpara_margadj.assign(hessgradparav.size(), 0.0);
      for (u64 l = 0; l < hessgradparav.size(); ++l) {
        for (u64 i = 0; i < K * K; ++i) tm[i] = hessgradparav[l].a[i] * hessi.a[i];
        gradpara_margadj[l] = -0.5 * accu2(tm.data(), K * K);
      }
    }
  }
  redohess = false;
}
void lpdfvec::settothess(const mat& h) { /* :609-612 */
  tothess = h;
  for (lpdf* l : lpdflist) l->settothess(h);
}
mat lpdfvec::hess_() { /* :503-512 */
  mat out;
  u64 cnt = 0;
  for (lpdf* l : lpdflist) {
    mat h = l->hess();
    if (cnt == 0) out = h;
    else {
      if (h.a.size() != out.a.size()) throw std::invalid_argument("addition: incompatible matrix dimensions");
      for (u64 i = 0; i < out.a.size(); ++i) out.a[i] += h.a[i];
    }
    cnt++;
  }
  return out;
}
std::vector<mat> lpdfvec::hessgradhyp_() { /* :520-531 */
  std::vector<mat> out;
  u64 cnt = 0;
  for (lpdf* l : lpdflist) {
    std::vector<mat> h = l->hessgradhyp();
    if (cnt == 0) out = h;
    else {
      if (h.size() != out.size()) throw std::invalid_argument("addition: incompatible cube dimensions");
      for (u64 s = 0; s < out.size(); ++s)
        for (u64 i = 0; i < out[s].a.size(); ++i) out[s].a[i] += h[s].a[i];
    }
    cnt++;
  }
  return out;
}
std::vector<mat> lpdfvec::hessgradpara_() { /* :539-549 */
  const u64 K = lpdflist[0]->nterms;
  std::vector<mat> out(para.size(), mat(K, K));
  for (u64 c = 0; c < lpdflist.size(); ++c) {
    std::vector<mat> h = lpdflist[c]->hessgradpara();
    if (h.size() != paraend[c] + 1 - parasrt[c]) throw std::invalid_argument("copy into subcube: incompatible dimensions");
    for (u64 j = 0; j < h.size(); ++j) out[parasrt[c] + j] = h[j];
  }
  return out;
}
void lpdfvec::update(const vec& coeff_) { /* :323-361 */
  coeff = coeff_;
  for (lpdf* l : lpdflist) {
    l->compute_val = compute_val; l->compute_grad = compute_grad;
    l->compute_gradhyp = compute_gradhyp; l->compute_gradpara = compute_gradpara;
  }
  for (lpdf* l : lpdflist) l->update(coeff);
  if (compute_val) val = 0;
  if (compute_grad) grad.assign(lpdflist[0]->grad.size(), 0.0);
  if (compute_gradhyp) gradhyp.assign(lpdflist[0]->gradhyp.size(), 0.0);
  if (compute_gradpara) gradpara.assign(para.size(), 0.0);
  buildhess();
  u64 cnt = 0;
  for (lpdf* l : lpdflist) {
    if (compute_val) val += l->val;
    if (compute_grad) for (u64 i = 0; i < grad.size(); ++i) grad[i] += l->grad[i];
    if (compute_gradhyp) for (u64 i = 0; i < gradhyp.size(); ++i) gradhyp[i] += l->gradhyp[i];
    if (compute_gradpara)
      for (u64 i = parasrt[cnt]; i <= paraend[cnt]; ++i) gradpara[i] += l->gradpara[i - parasrt[cnt]];
    cnt++;
  }
  if (domargadj) margadj();
}
void lpdfvec::margadj() { /* :371-380 */
  if (compute_val) val += val_margadj;
  if (compute_gradhyp) for (u64 i = 0; i < gradhyp.size(); ++i) gradhyp[i] += gradhyp_margadj[i];
  if (compute_gradpara) for (u64 i = 0; i < gradpara.size(); ++i) gradpara[i] += gradpara_margadj[i];
}
vec lpdfvec::hessmult(const vec& g) { /* :389-398 */
  vec out;
  u64 cnt = 0;
  for (lpdf* l : lpdflist) {
    vec h = l->hessmult(g);
    if (cnt == 0) out = h;
    else for (u64 i = 0; i < out.size(); ++i) out[i] += h[i];
    cnt++;
  }
  return out;
}
void lpdfvec::settotdiaghess(const vec& dh) { /* :604-607 */
  totdiaghess = dh;
  for (lpdf* l : lpdflist) l->settotdiaghess(dh);
}
double lpdfvec::paralpdf(const vec& parap) const { /* :468-476 */
  double out = 0;
  if (parap.size() != para.size()) return -std::numeric_limits<double>::infinity();
  for (u64 c = 0; c < lpdflist.size(); ++c)
    out += lpdflist[c]->paralpdf(vec(parap.begin() + parasrt[c], parap.begin() + paraend[c] + 1));
  return out;
}
vec lpdfvec::paralpdf_grad(const vec& parap) const { /* :485-494 */
  vec out(parap.size(), 0.0);
  if (parap.size() != para.size()) return out;
  for (u64 c = 0; c < lpdflist.size(); ++c) {
    vec g = lpdflist[c]->paralpdf_grad(vec(parap.begin() + parasrt[c], parap.begin() + paraend[c] + 1));
    std::copy(g.begin(), g.end(), out.begin() + parasrt[c]);
  }
  return out;
}
vec lpdfvec::diaghess_() { /* :557-566 */
  vec out;
  u64 cnt = 0;
  for (lpdf* l : lpdflist) {
    vec h = l->diaghess();
    if (cnt == 0) out = h;
    else for (u64 i = 0; i < out.size(); ++i) out[i] += h[i];
    cnt++;
  }
  return out;
}
mat lpdfvec::diaghessgradhyp_() { /* :574-584 */
  mat out;
  u64 cnt = 0;
  for (lpdf* l : lpdflist) {
    mat h = l->diaghessgradhyp();
    if (cnt == 0) out = h;
    else for (u64 i = 0; i < out.a.size(); ++i) out.a[i] += h.a[i];
    cnt++;
  }
  return out;
}
mat lpdfvec::diaghessgradpara_() { /* :592-602 */
  mat out(lpdflist[0]->nterms, para.size());
  out.zeros();
  for (u64 c = 0; c < lpdflist.size(); ++c) {
    mat h = lpdflist[c]->diaghessgradpara();
    for (u64 j = 0; j < h.nc; ++j)
      for (u64 i = 0; i < h.nr; ++i) out(i, parasrt[c] + j) = h(i, j);
  }
  return out;
}

/* ---- pred_gauss ---- */
pred_gauss::pred_gauss(const loglik_gauss& loglik) : om(loglik.om), para(loglik.para), terms(loglik.terms) { /* :196-212 */
  ob.reset(new outerbase(om, loglik.x, false));
  nthreads = (int)loglik.ob.nthreads;
  ob->nthreads = nthreads;
  coeff = loglik.coeff;
  if (!loglik.didnotothess) {
    coeffvar.resize(loglik.totdiaghess.size());
    for (u64 i = 0; i < coeffvar.size(); ++i) coeffvar[i] = 1 / loglik.totdiaghess[i];
  } else coeffvar.assign(coeff.size(), 0.0);
}
void pred_gauss::update(const mat& x_) { ob.reset(new outerbase(om, x_, false)); ob->nthreads = nthreads; } /* :214-218 */
vec pred_gauss::mean() const { vec o; ob->mm(o, terms, coeff); return o; } /* :220-222 */
vec pred_gauss::var() const { /* :223-227 */
  vec o = ob->sqmm(terms, coeffvar);
  const double c = std::exp(2 * para[0]);
  for (double& v : o) v += c;
  return o;
}

/* ---- predr_std ---- */
predr_std::predr_std(const loglik_std& loglik) : om(loglik.om), para(loglik.para), terms(loglik.terms) { /* :219-238 */
  ob.reset(new outerbase(om, loglik.x, false));
  nthreads = (int)loglik.ob.nthreads;
  ob->nthreads = nthreads;
  coeff = loglik.coeff;
  basismat = ob->getmat(terms);
  const u64 K = coeff.size();
  if (!loglik.didnotothess) {
    if (loglik.didfulltothess) coeffcov = dense_inv(loglik.tothess);
    else {
      coeffcov = mat(K, K);
      /* as written in the reference (:229): the diagonal of the total Hessian itself, not its inverse */
      for (u64 i = 0; i < K; ++i) coeffcov(i, i) = loglik.totdiaghess[i];
    }
  } else coeffcov = mat(K, K);
}
void predr_std::update(const mat& x_) { /* :240-245 */
  ob.reset(new outerbase(om, x_, false));
  ob->nthreads = nthreads;
  basismat = ob->getmat(terms);
}
vec predr_std::mean() const { return dense_mul(basismat, coeff); } /* :247-249 */
vec predr_std::var() const { /* :251-257 */
  mat adj = dense_mul(basismat, coeffcov);
  for (u64 i = 0; i < adj.a.size(); ++i) adj.a[i] *= basismat.a[i];
  vec out(adj.nr, 0.0);
  if (adj.nc) std::copy(adj.col(0), adj.col(0) + adj.nr, out.begin());
  for (u64 k = 1; k < adj.nc; ++k) { const double* c = adj.col(k); for (u64 i = 0; i < adj.nr; ++i) out[i] += c[i]; }
  const double e2 = std::exp(2 * para[0]);
  for (double& v : out) v += e2;
  return out;
}

} // namespace orc

/* The eigen-solver above, exported for oracle/ref_capi.cpp: the reference build (unmodified sources + Armadillo shim)
 * routes eig_sym to the SAME routine, so that both sides share one eigenbasis (SURVEY 7 "hard parts", 8c). */
extern "C" void orc_eig_sym_jacobi(uint64_t n, const double* A, double* w, double* V) {
  orc::mat Ain(n, n), Vm;
  std::copy(A, A + n * n, Ain.a.begin());
  orc::vec wv;
  orc::eig_sym_jacobi(wv, Vm, Ain);
  std::copy(wv.begin(), wv.end(), w);
  std::copy(Vm.a.begin(), Vm.a.end(), V);
}
