/*
 * ob_oracle.hpp -- CPU ORACLE for the outerbase hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain C++17 + OpenMP restatement (no Armadillo, no Rcpp) of the reference's
 * algorithm, following its loop structure and floating-point order.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  The product (outerbase_b200/) never does.
 *
 * PARITY PINNING: **pinned against the reference itself** (round 2).  `make -C oracle ref` compiles the UNMODIFIED
 * reference sources (/root/reference/src/{linalg,covfuncs,modandbase,fit}.cpp + lpdfs/) against oracle/arma_shim (a
 * header-only Armadillo subset) into oracle/_ref/libob_ref.so; tests/test_oracle_ref.py holds this oracle BITWISE equal
 * to it with one OpenMP thread -- terms, index tables, eigenbasis, basis matrices, every linalg.h kernel on both vertpl
 * branches, loglik_gauss / logpr_gauss / lpdfvec / loglik_gda, optcg iterates, a whole BFGS_lpdf run -- and to rounding
 * (<= 1e-13) with several threads, where the reference itself adds thread-local sums in arrival order.  The reference
 * ships no golden vectors of its own; committed fixtures (tests/golden/) freeze these outputs.
 * Still third-party, and SHARED by both sides of that comparison rather than pinned: eig_sym (LAPACK upstream; cyclic
 * Jacobi here), the BLAS behind Armadillo (netlib reference order here: dot_seq / k-ascending gemm) and R's RNG in
 * shuffle (injectable tie-break).  tests/test_oracle_ref.py::test_blas_flavour_sensitivity shows what the BLAS choice
 * moves (nothing in the linalg.h kernels on shared inputs; the high levels of the basis build).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference tree, src/...).
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

using u64 = uint64_t;
using i64 = int64_t;
using vec = std::vector<double>;

/* column-major dense matrix (Armadillo `mat`) */
struct mat {
  u64 nr = 0, nc = 0;
  std::vector<double> a;
  mat() {}
  mat(u64 r, u64 c) : nr(r), nc(c), a(r * c, 0.0) {}
  void set_size(u64 r, u64 c) { nr = r; nc = c; a.resize(r * c); }
  void zeros() { std::fill(a.begin(), a.end(), 0.0); }
  double& operator()(u64 i, u64 j) { return a[i + j * nr]; }
  double operator()(u64 i, u64 j) const { return a[i + j * nr]; }
  double* col(u64 j) { return a.data() + j * nr; }
  const double* col(u64 j) const { return a.data() + j * nr; }
};

/* column-major unsigned table (Armadillo `umat`) */
struct umat {
  u64 nr = 0, nc = 0;
  std::vector<u64> a;
  umat() {}
  umat(u64 r, u64 c) : nr(r), nc(c), a(r * c, 0) {}
  u64& operator()(u64 i, u64 j) { return a[i + j * nr]; }
  u64 operator()(u64 i, u64 j) const { return a[i + j * nr]; }
};

/* Armadillo arrayops::accumulate / op_dot::direct_dot_arma: two accumulators,
 * even/odd interleave, acc1 + acc2 at the end (call sites src/linalg.cpp:293,374,382). */
double accu2(const double* x, u64 n);
double dot2(const double* x, const double* y, u64 n);
/* netlib ddot / dgemv('T') order (left to right), and Armadillo's dot() = dot2 up to 32 elements, BLAS above */
double dot_seq(const double* x, const double* y, u64 n);
double dot_arma(const double* x, const double* y, u64 n);

/* ---- covariance functions: src/covfuncs.h:4-69, src/covfuncs.cpp:35-347 ---- */
struct covf {
  vec hyp, hypub, hyplb, hyp0, hypvar;
  double lowbnd = 0, uppbnd = 1;
  unsigned numhyp = 0;
  std::vector<std::string> hypnames;
  virtual ~covf() {}
  double lpdf(const vec& hypp) const;          /* covfuncs.cpp:35-50 */
  vec lpdf_gradhyp(const vec& hypp) const;     /* covfuncs.cpp:53-70 */
  bool inputcheck(const double* x, u64 n) const; /* covfuncs.h:23-27 */
  virtual void cov(mat& out, const double* x1, u64 n1, const double* x2, u64 n2) const = 0;
  /* out: numhyp slices, each n1 x n2 */
  virtual void cov_gradhyp(std::vector<mat>& out, const double* x1, u64 n1,
                           const double* x2, u64 n2) const = 0;
};
std::unique_ptr<covf> make_covf(const std::string& name); /* interfaceR.cpp:58-68 */

/* ---- outermod: src/modandbase.h:9-54, src/modandbase.cpp:67-440 ---- */
struct outermod {
  u64 d = 0;
  vec basisvar;
  bool setcovfs = false;
  std::vector<std::unique_ptr<covf>> covflist;
  vec hyp;
  std::vector<u64> knotptst, hypmatch, hypst, gest, knotptstge;
  vec knotpt;
  bool setknots = false;
  std::vector<i64> maxlevel;
  mat rotmat, rotmat_gradhyp;
  vec logbasisvar_gradhyp;
  u64 select_seed = 0; /* 0 = deterministic lowest-index tie-break */

  void set_covfs(const std::vector<std::string>& names); /* interfaceR.cpp:53-73 */
  void set_knot(const std::vector<vec>& L);              /* interfaceR.cpp:94-149 */
  void hyp_init();                                       /* modandbase.cpp:128-153 */
  void hyp_set(const vec& hyp_);                         /* modandbase.cpp:161-202 */
  void build();                                          /* modandbase.cpp:210-276 */
  void setsizes_();                                      /* modandbase.cpp:67-81 */
  /* xp: n x d column-major with leading dimension ldx, rows [0,n) */
  void buildob(mat& R, const double* xcol, u64 n, u64 k) const;                       /* :285-298 */
  void buildob(mat& R, std::vector<mat>& Rt, const double* xcol, u64 n, u64 k) const; /* :306-327 */
  vec getvar(const umat& terms) const;            /* :350-356 */
  mat getlvar_gradhyp(const umat& terms) const;   /* :364-379 */
  umat selectterms(unsigned numele) const;        /* :387-440 */
  double hyplpdf(const vec& hypp) const;          /* :89-99 */
  vec hyplpdf_grad(const vec& hypp) const;        /* :107-119 */
};

/* ---- linalg: src/linalg.h:9-58, src/linalg.cpp:57-715 ---- */
struct loopvals { bool vertpl; u64 chunksize, loopsize; int nthreads; };
void prodmm_(vec& out, const umat& terms, const vec& a, const mat& basemat, const vec& basescale,
             const std::vector<u64>& knotptst, const loopvals& lv);
void tprodmm_(vec& out, const umat& terms, const vec& a, const mat& basemat, const vec& basescale,
              const std::vector<u64>& knotptst, const loopvals& lv);
void prodmmge_(vec& out, mat& outge, const umat& terms, const vec& a, const mat& basemat,
               const vec& basescale, const std::vector<u64>& knotptst, const mat& basematge,
               const std::vector<u64>& gest, const std::vector<u64>& hypmatch, const loopvals& lv);
void tprodmmge_(vec& out, mat& outge, const umat& terms, const vec& a, const mat& basemat,
                const vec& basescale, const std::vector<u64>& knotptst, const mat& basematge,
                const std::vector<u64>& gest, const std::vector<u64>& hypmatch, const loopvals& lv);
void prodmm_mat_(mat& out, const umat& terms, const mat& a, const mat& basemat, const vec& basescale,
                 const std::vector<u64>& knotptst, const loopvals& lv);
void tprodmm_mat_(mat& out, const umat& terms, const mat& a, const mat& basemat, const vec& basescale,
                  const std::vector<u64>& knotptst, const loopvals& lv);
void getm_(mat& out, const umat& terms, const mat& basemat, const vec& basescale,
           const std::vector<u64>& knotptst, const loopvals& lv);
void getmge_(std::vector<mat>& outge, const umat& terms, const mat& basemat, const vec& basescale, const std::vector<u64>& knotptst,
             const mat& basematge, const std::vector<u64>& gest, const std::vector<u64>& hypmatch);

/* ---- outerbase: src/modandbase.h:57-125, src/modandbase.cpp:459-922 ---- */
struct outerbase {
  const outermod& om;
  mat xp;
  mat basemat;
  u64 d = 0, n_row = 0, n_hyp = 0;
  bool dograd = true;
  std::vector<u64> hypst;
  u64 loopsize = 0, chunksize = 128, nthreads = 1;
  bool vertpl = false;
  std::vector<u64> knotptst, gest, hypmatch;
  vec basescale, basescalesq;
  mat basescalemat, basemat_gradhyp, basematsq, basematsq_gradhyp;

  outerbase(const outermod& om_, const mat& xp_, bool dograd_); /* :459-483 */
  void setvals_();      /* :492-501 */
  void setloopvals_();  /* :504-513 */
  void setsizes_();     /* :521-539 */
  void build();         /* :547-626 */
  loopvals lv() const { return {vertpl, chunksize, loopsize, (int)nthreads}; }
  mat getbase(u64 dim) const;                      /* :634-639 */
  mat getmat(const umat& terms) const;             /* :649-654 */
  /* :663-669 -> getmge_ / dogetmge_, linalg.cpp:724-822.  One N x K slice per hyper-parameter.  The reference's
   * row-chunked branch (vertpl, linalg.cpp:788-810) cannot work (dogetmge_ resizes the whole cube to one chunk and the
   * chunk result is never copied back); the formula of the unchunked branch is used for every shape here. */
  std::vector<mat> getmat_gradhyp(const umat& terms) const;
  void mm(vec& out, const umat& terms, const vec& a) const;   /* :677-680 */
  void tmm(vec& out, const umat& terms, const vec& a) const;  /* :700-703 */
  void mm_gradhyp(vec& out, mat& outge, const umat& terms, const vec& a) const;  /* :725-731 */
  void tmm_gradhyp(vec& out, mat& outge, const umat& terms, const vec& a) const; /* :755-761 */
  vec sqmm(const umat& terms, const vec& a) const;           /* :784-790 */
  mat sqmm_gradhyp(const umat& terms, const vec& a) const;   /* :798-808 */
  vec sqtmm(const umat& terms, const vec& a) const;          /* :816-822 */
  mat sqtmmm(const umat& terms, const mat& a) const;         /* :831-837 */
  mat sqtmm_gradhyp(const umat& terms, const vec& a) const;  /* :845-855 */
  vec sqcolsums(const umat& terms) const;                    /* :863-867 */
  mat sqcolsums_gradhyp(const umat& terms) const;            /* :875-879 */
  vec residvar(const umat& terms) const;                     /* :889-896 */
  mat residvar_gradhyp(const umat& terms) const;             /* :904-922 */
};

/* ---- lpdf family: src/fit.h:23-148,236-268 ; src/fit.cpp ; src/lpdfs/ ---- */
struct lpdf {
  double val = 0;
  vec grad, gradhyp, gradpara, para;
  umat terms;
  vec coeff, totdiaghess;
  mat tothess;
  bool didfulltothess = false, didnotothess = true, fullhess = false;
  bool compute_val = true, compute_grad = true, compute_gradhyp = false, compute_gradpara = false;
  unsigned npara = 0, nterms = 0;
  vec para0, paravar;
  u64 cg_iters = 0;
  virtual ~lpdf() {}
  virtual void setnthreads(int) {}
  virtual double paralpdf(const vec& parap) const;    /* fit.cpp:133-142 */
  virtual vec paralpdf_grad(const vec& parap) const;  /* fit.cpp:146-157 */
  virtual void optcg(double tol, unsigned maxepch);   /* fit.cpp:37-96 */
  virtual void optnewton();                           /* fit.cpp:98-131 */
  virtual void updateom() {}
  virtual void updatepara(const vec&) {}
  virtual void updateterms(const umat&) {}
  virtual void update(const vec&) {}
  virtual vec hessmult(const vec&) { return {}; }
  virtual vec diaghess() { return {}; }
  virtual mat diaghessgradhyp() { return {}; }
  virtual mat diaghessgradpara() { return {}; }
  virtual void settotdiaghess(const vec& dh) { totdiaghess = dh; didfulltothess = false; didnotothess = false; }
  virtual void settothess(const mat& h) { tothess = h; didfulltothess = true; didnotothess = false; } /* fit.h:81-85 */
  virtual mat hess() { return {}; }
  virtual std::vector<mat> hessgradhyp() { return {}; }  /* cube: one K x K slice per hyper-parameter */
  virtual std::vector<mat> hessgradpara() { return {}; } /* one slice per parameter */
  virtual u64 nhyp() const { return 0; }
  virtual u64 nrow() const { return 0; }
};

struct logpr_gauss : lpdf { /* src/lpdfs/logpr_gauss.cpp:41-145 */
  const outermod& om;
  vec coeffsd;
  mat coefflvarge;
  vec stdresid;
  double sca = 1;
  logpr_gauss(const outermod& om_, const umat& terms_);
  void updateom() override;
  void updatepara(const vec&) override;
  void updateterms(const umat&) override;
  void update(const vec& coeff_) override;
  vec hessmult(const vec& g) override;
  vec diaghess() override;
  mat diaghessgradhyp() override;
  mat diaghessgradpara() override;
  mat hess() override;                      /* :153-158 */
  std::vector<mat> hessgradhyp() override;  /* :165-173 */
  std::vector<mat> hessgradpara() override; /* :181-186 */
  u64 nhyp() const override { return om.hypmatch.size(); }
};

struct loglik_std : lpdf { /* src/lpdfs/loglik_std.cpp:41-203, src/fit.h:182-211: the same model as loglik_gauss on the
                             EXPLICIT basis matrix (N x K) and its hyper-gradient cube, with the full Hessian */
  const outermod& om;
  outerbase ob;
  mat basismat;
  std::vector<mat> basismat_gradhyp;
  vec y, yhat;
  mat x;
  loglik_std(const outermod& om_, const umat& terms_, const vec& y_, const mat& x_);
  void setnthreads(int) override {} /* not overridden in the reference: ob keeps its thread count */
  void updateom() override;
  void updatepara(const vec&) override;
  void updateterms(const umat&) override;
  void update(const vec& coeff_) override;
  vec hessmult(const vec& g) override;
  vec diaghess() override;
  mat diaghessgradhyp() override;
  mat diaghessgradpara() override;
  mat hess() override;
  std::vector<mat> hessgradhyp() override;
  std::vector<mat> hessgradpara() override;
  u64 nhyp() const override { return ob.n_hyp; }
  u64 nrow() const override { return ob.n_row; }
};

struct loglik_gauss : lpdf { /* src/lpdfs/loglik_gauss.cpp:41-179 */
  const outermod& om;
  outerbase ob;
  vec y;
  mat x;
  vec yhat;
  vec obsvar, lobsvar, obssd;
  mat yhatge;
  vec gradtemp, yhattemp, residtemp, residtemp2;
  loglik_gauss(const outermod& om_, const umat& terms_, const vec& y_, const mat& x_);
  void setnthreads(int k) override;
  void updateom() override;
  void updatepara(const vec&) override;
  void updateterms(const umat&) override;
  void update(const vec& coeff_) override;
  vec hessmult(const vec& g) override;
  vec diaghess() override;
  mat diaghessgradhyp() override;
  mat diaghessgradpara() override;
  u64 nhyp() const override { return ob.n_hyp; }
  u64 nrow() const override { return ob.n_row; }
};

struct loglik_gda : lpdf { /* src/lpdfs/loglik_gda.cpp:47-239, src/fit.h:272-310: Gaussian likelihood whose noise
                             variance carries the variance of the terms left out of the basis (obfit stage 1) */
  const outermod& om;
  outerbase ob;
  vec y;
  mat x;
  bool doda = true, redostd = true;
  vec yhat, obssd;
  mat yhatge, obssd_gradhyp, obssd_gradpara;
  vec gradtemp, yhattemp, residtemp, residtemp2;
  loglik_gda(const outermod& om_, const umat& terms_, const vec& y_, const mat& x_);
  void setnthreads(int k) override;
  void updateom() override;
  void updatepara(const vec&) override;
  void updateterms(const umat&) override;
  void update(const vec& coeff_) override;
  vec hessmult(const vec& g) override;
  vec diaghess() override;
  mat diaghessgradhyp() override;
  mat diaghessgradpara() override;
  void buildstd();
  u64 nhyp() const override { return ob.n_hyp; }
  u64 nrow() const override { return ob.n_row; }
};

struct lpdfvec : lpdf { /* src/fit.h:93-148 ; src/fit.cpp:174-267,310-428,557-607 */
  double val_margadj = 0;
  vec gradhyp_margadj, gradpara_margadj;
  bool domargadj = true;
  vec diaghessv;
  mat diaghessgradhypv, diaghessgradparav;
  mat hessv;
  std::vector<mat> hessgradhypv, hessgradparav;
  bool redohess = true;
  std::vector<lpdf*> lpdflist;
  std::vector<u64> parasrt, paraend;
  lpdfvec(lpdf& a, lpdf& b);
  void setnthreads(int k) override;
  void updateom() override;
  void updatepara(const vec&) override;
  void updateterms(const umat&) override;
  void update(const vec&) override;
  vec hessmult(const vec&) override;
  vec diaghess() override { return diaghessv; }
  mat diaghessgradhyp() override { return diaghessgradhypv; }
  mat diaghessgradpara() override { return diaghessgradparav; }
  void settotdiaghess(const vec& dh) override;
  void settothess(const mat& h) override;                           /* :609-612 */
  mat hess() override { return hessv; }                              /* :436-438 */
  std::vector<mat> hessgradhyp() override { return hessgradhypv; }   /* :446-448 */
  std::vector<mat> hessgradpara() override { return hessgradparav; } /* :456-458 */
  mat hess_();                       /* :503-512 */
  std::vector<mat> hessgradhyp_();   /* :520-531 */
  std::vector<mat> hessgradpara_();  /* :539-549 */
  double paralpdf(const vec& parap) const override;
  vec paralpdf_grad(const vec& parap) const override;
  void buildhess();
  void margadj();
  vec diaghess_();
  mat diaghessgradhyp_();
  mat diaghessgradpara_();
  u64 nhyp() const override { return lpdflist[0]->nhyp(); }
  u64 nrow() const override { return std::max(lpdflist[0]->nrow(), lpdflist[1]->nrow()); }
};

struct pred_gauss { /* src/lpdfs/loglik_gauss.cpp:196-227 */
  const outermod& om;
  vec para;
  umat terms;
  int nthreads = 0;
  vec coeff, coeffvar;
  std::unique_ptr<outerbase> ob;
  pred_gauss(const loglik_gauss& loglik);
  void update(const mat& x_);
  vec mean() const;
  vec var() const;
};

struct predr_std { /* src/lpdfs/loglik_std.cpp:219-255 */
  const outermod& om;
  vec para;
  umat terms;
  mat basismat;
  int nthreads = 0;
  vec coeff;
  mat coeffcov;
  std::unique_ptr<outerbase> ob;
  predr_std(const loglik_std& loglik);
  void update(const mat& x_);
  vec mean() const;
  vec var() const;
};

struct pred_gda { /* src/lpdfs/loglik_gda.cpp:249-283 */
  const outermod& om;
  vec para;
  umat terms;
  int nthreads = 0;
  bool doda = true;
  vec coeff, coeffvar;
  std::unique_ptr<outerbase> ob;
  pred_gda(const loglik_gda& loglik);
  void update(const mat& x_);
  vec mean() const;
  vec var() const;
};

} // namespace orc
