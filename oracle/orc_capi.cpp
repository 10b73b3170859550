/*
 * orc_capi.cpp -- C ABI of the CPU ORACLE (test infrastructure only).
 * Same shape as include/outerbase_b200.h with every `ob_` replaced by `orc_`
 * (the header oracle/_gen/orc_api.h is generated from it by the Makefile), so
 * tests drive the oracle and the CUDA product through one harness.
 * "Device" pointers of the *_dev entry points are plain host pointers here.
 */
#include "_gen/orc_api.h"
#include "ob_oracle.hpp"

#include <cstdio>
#include <map>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH                                                                      \
  }                                                                                    \
  catch (const std::range_error& e) { g_err = e.what(); return ORC_ERR_INVALID; }      \
  catch (const std::invalid_argument& e) { g_err = e.what(); return ORC_ERR_INVALID; } \
  catch (const std::exception& e) { g_err = e.what(); return ORC_ERR_STATE; }          \
  return ORC_OK;

struct orc_ctx { int nthreads; uint64_t launches = 0; };
struct orc_outermod { outermod om; };
struct orc_outerbase {
  orc_ctx* ctx;
  std::unique_ptr<outerbase> ob;
  umat terms;
};
struct orc_lpdf {
  std::unique_ptr<lpdf> p;
  int kind; /* 0 loglik_gauss, 1 logpr_gauss, 2 lpdfvec */
};
struct orc_predictor { std::unique_ptr<pred_gauss> p; std::unique_ptr<pred_gda> pg; std::unique_ptr<predr_std> ps; uint64_t d; };

static umat to_umat(const uint64_t* t, uint64_t K, uint64_t d) {
  umat m(K, d);
  std::copy(t, t + K * d, m.a.begin());
  return m;
}
static mat to_mat(const double* x, uint64_t r, uint64_t c) {
  mat m(r, c);
  std::copy(x, x + r * c, m.a.begin());
  return m;
}

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }
int orc_version(void) { return 1; }

int orc_ctx_create(int device, orc_ctx** out) {
  (void)device;
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_num_procs();
#endif
  *out = new orc_ctx{nt};
  return ORC_OK;
}
int orc_ctx_destroy(orc_ctx* c) { delete c; return ORC_OK; }
int orc_ctx_synchronize(orc_ctx*) { return ORC_OK; }
int orc_ctx_stream(orc_ctx*, void** s) { *s = nullptr; return ORC_OK; }
int orc_comm_get_unique_id(void* id) { std::memset(id, 0, 128); return ORC_OK; }
int orc_ctx_comm_init(orc_ctx*, int, int, const void*) { g_err = "oracle is single-rank"; return ORC_ERR_STATE; }
int orc_ctx_comm_info(orc_ctx*, int* n, int* r) { *n = 1; *r = 0; return ORC_OK; }
int orc_ctx_allreduce_dev(orc_ctx*, double*, uint64_t) { return ORC_OK; } /* one rank: the sum over ranks is the input */
int orc_ctx_fp64_peak(orc_ctx*, double* t) { *t = 0; return ORC_OK; }
int orc_ctx_launch_count(orc_ctx*, uint64_t* c) { *c = 0; return ORC_OK; }
/* kernel specialisation is a property of the CUDA product; the oracle accepts and ignores the requests */
int orc_ctx_set_option(orc_ctx*, const char*, double) { return ORC_OK; }
int orc_outerbase_specialize(orc_outerbase*, const uint64_t*, uint64_t, double* s) { if (s) *s = 0; return ORC_OK; }
int orc_outerbase_spec_state(orc_outerbase*, const uint64_t*, uint64_t, int* st) { if (st) *st = 0; return ORC_OK; }
int orc_spec_source(const uint64_t*, uint64_t, uint64_t, const int*, char*, uint64_t*, uint64_t*) { g_err = "the oracle has no kernels"; return ORC_ERR_STATE; }
int orc_spec_source_dot(const uint64_t*, uint64_t, uint64_t, char*, uint64_t*, uint64_t*) { g_err = "the oracle has no kernels"; return ORC_ERR_STATE; }
int orc_spec_source_mat(const uint64_t*, uint64_t, uint64_t, char*, uint64_t*, uint64_t*) { g_err = "the oracle has no kernels"; return ORC_ERR_STATE; }
int orc_spec_source_tmat(const uint64_t*, uint64_t, uint64_t, char*, uint64_t*, uint64_t*) { g_err = "the oracle has no kernels"; return ORC_ERR_STATE; }
int orc_spec_compile_check(const char*, uint64_t*, double*) { g_err = "the oracle has no kernels"; return ORC_ERR_STATE; }

int orc_covf_numhyp(const char* name, uint64_t* n) { ORC_TRY *n = make_covf(name)->numhyp; ORC_CATCH }
int orc_covf_cov(orc_ctx*, const char* name, const double* hyp, const double* x1, uint64_t n1,
                 const double* x2, uint64_t n2, double* out) {
  ORC_TRY
  auto c = make_covf(name);
  for (unsigned i = 0; i < c->numhyp; ++i) c->hyp[i] = hyp[i];
  mat h;
  c->cov(h, x1, n1, x2, n2);
  std::copy(h.a.begin(), h.a.end(), out);
  ORC_CATCH
}
int orc_covf_cov_gradhyp(orc_ctx*, const char* name, const double* hyp, const double* x1, uint64_t n1,
                         const double* x2, uint64_t n2, double* out) {
  ORC_TRY
  auto c = make_covf(name);
  for (unsigned i = 0; i < c->numhyp; ++i) c->hyp[i] = hyp[i];
  std::vector<mat> g;
  c->cov_gradhyp(g, x1, n1, x2, n2);
  for (size_t s = 0; s < g.size(); ++s) std::copy(g[s].a.begin(), g[s].a.end(), out + s * n1 * n2);
  ORC_CATCH
}

/* ---- outermod ---- */
int orc_outermod_create(orc_outermod** out) { *out = new orc_outermod(); return ORC_OK; }
int orc_outermod_destroy(orc_outermod* om) { delete om; return ORC_OK; }
int orc_outermod_setcovfs(orc_outermod* om, uint64_t d, const char* const* names) {
  ORC_TRY
  std::vector<std::string> v;
  for (uint64_t i = 0; i < d; ++i) v.push_back(names[i]);
  om->om.set_covfs(v);
  ORC_CATCH
}
int orc_outermod_setknot(orc_outermod* om, const double* knots, const uint64_t* lens) {
  ORC_TRY
  std::vector<vec> L;
  const double* p = knots;
  for (uint64_t l = 0; l < om->om.d; ++l) { L.push_back(vec(p, p + lens[l])); p += lens[l]; }
  om->om.set_knot(L);
  ORC_CATCH
}
int orc_outermod_updatehyp(orc_outermod* om, const double* hyp, uint64_t n) { ORC_TRY om->om.hyp_set(vec(hyp, hyp + n)); ORC_CATCH }
int orc_outermod_gethyp(orc_outermod* om, double* hyp) { std::copy(om->om.hyp.begin(), om->om.hyp.end(), hyp); return ORC_OK; }
int orc_outermod_sizes(orc_outermod* om, uint64_t* d, uint64_t* nhyp, uint64_t* nknot, uint64_t* nge) {
  *d = om->om.d; *nhyp = om->om.hyp.size(); *nknot = om->om.knotpt.size();
  *nge = om->om.knotptstge.empty() ? 0 : om->om.knotptstge[om->om.d];
  return ORC_OK;
}
int orc_outermod_set_select_seed(orc_outermod* om, uint64_t seed) { om->om.select_seed = seed; return ORC_OK; }
int orc_outermod_selectterms(orc_outermod* om, uint64_t numele, uint64_t* terms) {
  ORC_TRY
  umat t = om->om.selectterms((unsigned)numele);
  std::copy(t.a.begin(), t.a.end(), terms);
  ORC_CATCH
}
int orc_outermod_getvar(orc_outermod* om, const uint64_t* terms, uint64_t K, double* out) {
  ORC_TRY
  vec v = om->om.getvar(to_umat(terms, K, om->om.d));
  std::copy(v.begin(), v.end(), out);
  ORC_CATCH
}
int orc_outermod_getlvar_gradhyp(orc_outermod* om, const uint64_t* terms, uint64_t K, double* out) {
  ORC_TRY
  mat v = om->om.getlvar_gradhyp(to_umat(terms, K, om->om.d));
  std::copy(v.a.begin(), v.a.end(), out);
  ORC_CATCH
}
int orc_outermod_hyplpdf(orc_outermod* om, const double* hyp, uint64_t n, double* out) { ORC_TRY *out = om->om.hyplpdf(vec(hyp, hyp + n)); ORC_CATCH }
int orc_outermod_hyplpdf_grad(orc_outermod* om, const double* hyp, uint64_t n, double* out) {
  ORC_TRY
  vec g = om->om.hyplpdf_grad(vec(hyp, hyp + n));
  std::copy(g.begin(), g.end(), out);
  ORC_CATCH
}
int orc_outermod_get_index(orc_outermod* om, const char* which, int64_t* out, uint64_t* n) {
  ORC_TRY
  const std::string w = which;
  const outermod& m = om->om;
  std::vector<int64_t> v;
  auto cp = [&](const std::vector<u64>& s) { v.assign(s.begin(), s.end()); };
  if (w == "knotptst") cp(m.knotptst);
  else if (w == "hypst") cp(m.hypst);
  else if (w == "hypmatch") cp(m.hypmatch);
  else if (w == "gest") cp(m.gest);
  else if (w == "knotptstge") cp(m.knotptstge);
  else if (w == "maxlevel") v.assign(m.maxlevel.begin(), m.maxlevel.end());
  else throw std::invalid_argument("unknown index table " + w);
  *n = v.size();
  if (out) std::copy(v.begin(), v.end(), out);
  ORC_CATCH
}
int orc_outermod_get_real(orc_outermod* om, const char* which, double* out, uint64_t* nrow, uint64_t* ncol) {
  ORC_TRY
  const std::string w = which;
  const outermod& m = om->om;
  const std::vector<double>* src = nullptr;
  if (w == "basisvar") { src = &m.basisvar; *nrow = src->size(); *ncol = 1; }
  else if (w == "knotpt") { src = &m.knotpt; *nrow = src->size(); *ncol = 1; }
  else if (w == "logbasisvar_gradhyp") { src = &m.logbasisvar_gradhyp; *nrow = src->size(); *ncol = 1; }
  else if (w == "rotmat") { src = &m.rotmat.a; *nrow = m.rotmat.nr; *ncol = m.rotmat.nc; }
  else if (w == "rotmat_gradhyp") { src = &m.rotmat_gradhyp.a; *nrow = m.rotmat_gradhyp.nr; *ncol = m.rotmat_gradhyp.nc; }
  else throw std::invalid_argument("unknown real table " + w);
  if (out) std::copy(src->begin(), src->end(), out);
  ORC_CATCH
}

/* ---- outerbase ---- */
int orc_outerbase_create(orc_ctx* ctx, orc_outermod* om, const double* x, uint64_t N, int dograd, orc_outerbase** out) {
  ORC_TRY
  if (!om->om.setknots) throw std::range_error("Need to set covfs and knots before building.");
  auto* h = new orc_outerbase();
  h->ctx = ctx;
  h->ob.reset(new outerbase(om->om, to_mat(x, N, om->om.d), dograd != 0));
  *out = h;
  ORC_CATCH
}
int orc_outerbase_destroy(orc_outerbase* ob) { delete ob; return ORC_OK; }
int orc_outerbase_build(orc_outerbase* ob) { ORC_TRY ob->ob->build(); ORC_CATCH }
int orc_outerbase_set_nthreads(orc_outerbase* ob, int n) { ob->ob->nthreads = n; return ORC_OK; }
int orc_outerbase_loopvals(orc_outerbase* ob, uint64_t* nthreads, uint64_t* chunksize, uint64_t* loopsize, int* vertpl) {
  *nthreads = ob->ob->nthreads; *chunksize = ob->ob->chunksize; *loopsize = ob->ob->loopsize; *vertpl = ob->ob->vertpl;
  return ORC_OK;
}
int orc_outerbase_get_real(orc_outerbase* ob, const char* which, double* out, uint64_t* nrow, uint64_t* ncol) {
  ORC_TRY
  const std::string w = which;
  const outerbase& b = *ob->ob;
  const std::vector<double>* src = nullptr;
  if (w == "basemat") { src = &b.basemat.a; *nrow = b.basemat.nr; *ncol = b.basemat.nc; }
  else if (w == "basemat_gradhyp") { src = &b.basemat_gradhyp.a; *nrow = b.basemat_gradhyp.nr; *ncol = b.basemat_gradhyp.nc; }
  else if (w == "basescale") { src = &b.basescale; *nrow = b.basescale.size(); *ncol = 1; }
  else if (w == "basescalemat") { src = &b.basescalemat.a; *nrow = b.basescalemat.nr; *ncol = b.basescalemat.nc; }
  else throw std::invalid_argument("unknown matrix " + w);
  if (out) std::copy(src->begin(), src->end(), out);
  ORC_CATCH
}
int orc_outerbase_getbase(orc_outerbase* ob, uint64_t dim, double* out) {
  ORC_TRY
  if (dim < 1 || dim > ob->ob->d) throw std::range_error("dim out of range");
  mat m = ob->ob->getbase(dim);
  std::copy(m.a.begin(), m.a.end(), out);
  ORC_CATCH
}
int orc_outerbase_getmat_gradhyp(orc_outerbase* ob, const uint64_t* terms, uint64_t K, double* out) {
  ORC_TRY
  std::vector<mat> g = ob->ob->getmat_gradhyp(to_umat(terms, K, ob->ob->d));
  uint64_t off = 0;
  for (const mat& m : g) { std::copy(m.a.begin(), m.a.end(), out + off); off += m.a.size(); }
  ORC_CATCH
}
int orc_outerbase_getmat(orc_outerbase* ob, const uint64_t* terms, uint64_t K, double* out) {
  ORC_TRY
  mat m = ob->ob->getmat(to_umat(terms, K, ob->ob->d));
  std::copy(m.a.begin(), m.a.end(), out);
  ORC_CATCH
}
static void mm_impl(const outerbase& ob, int sq, const umat& t, const double* a, double* out) {
  vec av(a, a + t.nr), o;
  if (sq) o = ob.sqmm(t, av); else ob.mm(o, t, av);
  std::copy(o.begin(), o.end(), out);
}
static void tmm_impl(const outerbase& ob, int sq, const umat& t, const double* a, double* out) {
  vec av(a, a + ob.n_row), o;
  if (sq) o = ob.sqtmm(t, av); else ob.tmm(o, t, av);
  std::copy(o.begin(), o.end(), out);
}
int orc_outerbase_mm(orc_outerbase* ob, int sq, const uint64_t* terms, uint64_t K, const double* a, double* out) {
  ORC_TRY mm_impl(*ob->ob, sq, to_umat(terms, K, ob->ob->d), a, out); ORC_CATCH
}
int orc_outerbase_tmm(orc_outerbase* ob, int sq, const uint64_t* terms, uint64_t K, const double* a, double* out) {
  ORC_TRY tmm_impl(*ob->ob, sq, to_umat(terms, K, ob->ob->d), a, out); ORC_CATCH
}
int orc_outerbase_mm_gradhyp(orc_outerbase* ob, int sq, const uint64_t* terms, uint64_t K, const double* a,
                             double* out, double* outge) {
  ORC_TRY
  const outerbase& b = *ob->ob;
  umat t = to_umat(terms, K, b.d);
  vec av(a, a + K), o;
  mat g;
  if (sq) prodmmge_(o, g, t, av, b.basematsq, b.basescalesq, b.knotptst, b.basematsq_gradhyp, b.gest, b.hypmatch, b.lv());
  else b.mm_gradhyp(o, g, t, av);
  if (out) std::copy(o.begin(), o.end(), out);
  std::copy(g.a.begin(), g.a.end(), outge);
  ORC_CATCH
}
int orc_outerbase_tmm_gradhyp(orc_outerbase* ob, int sq, const uint64_t* terms, uint64_t K, const double* a,
                              double* out, double* outge) {
  ORC_TRY
  const outerbase& b = *ob->ob;
  umat t = to_umat(terms, K, b.d);
  vec av(a, a + b.n_row), o;
  mat g;
  if (sq) tprodmmge_(o, g, t, av, b.basematsq, b.basescalesq, b.knotptst, b.basematsq_gradhyp, b.gest, b.hypmatch, b.lv());
  else b.tmm_gradhyp(o, g, t, av);
  if (out) std::copy(o.begin(), o.end(), out);
  std::copy(g.a.begin(), g.a.end(), outge);
  ORC_CATCH
}
int orc_outerbase_mm_mat(orc_outerbase* ob, int sq, const uint64_t* terms, uint64_t K, const double* A, uint64_t C, double* out) {
  ORC_TRY
  const outerbase& b = *ob->ob;
  mat o;
  prodmm_mat_(o, to_umat(terms, K, b.d), to_mat(A, K, C), sq ? b.basematsq : b.basemat, sq ? b.basescalesq : b.basescale, b.knotptst, b.lv());
  std::copy(o.a.begin(), o.a.end(), out);
  ORC_CATCH
}
int orc_outerbase_tmm_mat(orc_outerbase* ob, int sq, const uint64_t* terms, uint64_t K, const double* A, uint64_t C, double* out) {
  ORC_TRY
  const outerbase& b = *ob->ob;
  mat o;
  tprodmm_mat_(o, to_umat(terms, K, b.d), to_mat(A, b.n_row, C), sq ? b.basematsq : b.basemat, sq ? b.basescalesq : b.basescale, b.knotptst, b.lv());
  std::copy(o.a.begin(), o.a.end(), out);
  ORC_CATCH
}
int orc_outerbase_residvar(orc_outerbase* ob, const uint64_t* terms, uint64_t K, double* out) {
  ORC_TRY vec o = ob->ob->residvar(to_umat(terms, K, ob->ob->d)); std::copy(o.begin(), o.end(), out); ORC_CATCH
}
int orc_outerbase_residvar_gradhyp(orc_outerbase* ob, const uint64_t* terms, uint64_t K, double* out) {
  ORC_TRY mat o = ob->ob->residvar_gradhyp(to_umat(terms, K, ob->ob->d)); std::copy(o.a.begin(), o.a.end(), out); ORC_CATCH
}
int orc_outerbase_set_terms(orc_outerbase* ob, const uint64_t* terms, uint64_t K) { ORC_TRY ob->terms = to_umat(terms, K, ob->ob->d); ORC_CATCH }
int orc_outerbase_mm_dev(orc_outerbase* ob, int sq, const double* a, double* out) { ORC_TRY mm_impl(*ob->ob, sq, ob->terms, a, out); ORC_CATCH }
int orc_outerbase_tmm_dev(orc_outerbase* ob, int sq, const double* a, double* out) { ORC_TRY tmm_impl(*ob->ob, sq, ob->terms, a, out); ORC_CATCH }
int orc_outerbase_mm_mat_dev(orc_outerbase* ob, int sq, const double* A, uint64_t C, double* out) {
  return orc_outerbase_mm_mat(ob, sq, ob->terms.a.data(), ob->terms.nr, A, C, out);
}
int orc_outerbase_tmm_mat_dev(orc_outerbase* ob, int sq, const double* A, uint64_t C, double* out) {
  return orc_outerbase_tmm_mat(ob, sq, ob->terms.a.data(), ob->terms.nr, A, C, out);
}
int orc_outerbase_terms_stats(orc_outerbase* ob, uint64_t* W, uint64_t* Lcols, uint64_t* nodes, uint64_t* maxdepth) {
  ORC_TRY
  const umat& t = ob->terms;
  uint64_t w = 0, lc = 0, md = 0;
  for (uint64_t l = 0; l < t.nc; ++l) { uint64_t mx = 0; for (uint64_t k = 0; k < t.nr; ++k) mx = std::max(mx, t(k, l)); lc += mx; }
  for (uint64_t k = 0; k < t.nr; ++k) { uint64_t nz = 0; for (uint64_t l = 0; l < t.nc; ++l) nz += t(k, l) > 0; w += nz + 1; md = std::max(md, nz); }
  *W = w; *Lcols = lc; *nodes = t.nr; *maxdepth = md;
  ORC_CATCH
}

/* ---- stateless linalg.h seam ---- */
static std::vector<u64> kp(const uint64_t* p, uint64_t n) { return std::vector<u64>(p, p + n); }
static loopvals default_lv(uint64_t N) {
  int T = 1;
#ifdef _OPENMP
  T = omp_get_num_procs();
#endif
  const u64 chunk = std::max<u64>(32, std::min<u64>(1 + 2048 / T, N / (4 * T) + 1));
  const u64 loops = (N + chunk - 1) / chunk;
  return {loops > 20, chunk, loops, T};
}
int orc_prodmm_vec(orc_ctx*, double* out, const uint64_t* terms, uint64_t K, uint64_t d, const double* a,
                   const double* basemat, uint64_t N, uint64_t M, const double* basescale, const uint64_t* knotptst) {
  ORC_TRY
  vec o;
  prodmm_(o, to_umat(terms, K, d), vec(a, a + K), to_mat(basemat, N, M), vec(basescale, basescale + N), kp(knotptst, d + 1), default_lv(N));
  std::copy(o.begin(), o.end(), out);
  ORC_CATCH
}
int orc_tprodmm_vec(orc_ctx*, double* out, const uint64_t* terms, uint64_t K, uint64_t d, const double* a,
                    const double* basemat, uint64_t N, uint64_t M, const double* basescale, const uint64_t* knotptst) {
  ORC_TRY
  vec o;
  tprodmm_(o, to_umat(terms, K, d), vec(a, a + N), to_mat(basemat, N, M), vec(basescale, basescale + N), kp(knotptst, d + 1), default_lv(N));
  std::copy(o.begin(), o.end(), out);
  ORC_CATCH
}
int orc_prodmm_mat(orc_ctx*, double* out, const uint64_t* terms, uint64_t K, uint64_t d, const double* A, uint64_t C,
                   const double* basemat, uint64_t N, uint64_t M, const double* basescale, const uint64_t* knotptst) {
  ORC_TRY
  mat o;
  prodmm_mat_(o, to_umat(terms, K, d), to_mat(A, K, C), to_mat(basemat, N, M), vec(basescale, basescale + N), kp(knotptst, d + 1), default_lv(N));
  std::copy(o.a.begin(), o.a.end(), out);
  ORC_CATCH
}
int orc_tprodmm_mat(orc_ctx*, double* out, const uint64_t* terms, uint64_t K, uint64_t d, const double* A, uint64_t C,
                    const double* basemat, uint64_t N, uint64_t M, const double* basescale, const uint64_t* knotptst) {
  ORC_TRY
  mat o;
  tprodmm_mat_(o, to_umat(terms, K, d), to_mat(A, N, C), to_mat(basemat, N, M), vec(basescale, basescale + N), kp(knotptst, d + 1), default_lv(N));
  std::copy(o.a.begin(), o.a.end(), out);
  ORC_CATCH
}
int orc_prodmmge(orc_ctx*, double* out, double* outge, const uint64_t* terms, uint64_t K, uint64_t d, const double* a,
                 const double* basemat, uint64_t N, uint64_t M, const double* basescale, const uint64_t* knotptst,
                 const double* basematge, uint64_t Mge, const uint64_t* gest, const uint64_t* hypmatch, uint64_t H) {
  ORC_TRY
  vec o; mat g;
  prodmmge_(o, g, to_umat(terms, K, d), vec(a, a + K), to_mat(basemat, N, M), vec(basescale, basescale + N), kp(knotptst, d + 1),
            to_mat(basematge, N, Mge), kp(gest, H + 1), kp(hypmatch, H), default_lv(N));
  std::copy(o.begin(), o.end(), out);
  std::copy(g.a.begin(), g.a.end(), outge);
  ORC_CATCH
}
int orc_tprodmmge(orc_ctx*, double* out, double* outge, const uint64_t* terms, uint64_t K, uint64_t d, const double* a,
                  const double* basemat, uint64_t N, uint64_t M, const double* basescale, const uint64_t* knotptst,
                  const double* basematge, uint64_t Mge, const uint64_t* gest, const uint64_t* hypmatch, uint64_t H) {
  ORC_TRY
  vec o; mat g;
  tprodmmge_(o, g, to_umat(terms, K, d), vec(a, a + N), to_mat(basemat, N, M), vec(basescale, basescale + N), kp(knotptst, d + 1),
             to_mat(basematge, N, Mge), kp(gest, H + 1), kp(hypmatch, H), default_lv(N));
  std::copy(o.begin(), o.end(), out);
  std::copy(g.a.begin(), g.a.end(), outge);
  ORC_CATCH
}
int orc_getm(orc_ctx*, double* out, const uint64_t* terms, uint64_t K, uint64_t d, const double* basemat, uint64_t N,
             uint64_t M, const double* basescale, const uint64_t* knotptst) {
  ORC_TRY
  mat o;
  getm_(o, to_umat(terms, K, d), to_mat(basemat, N, M), vec(basescale, basescale + N), kp(knotptst, d + 1), default_lv(N));
  std::copy(o.a.begin(), o.a.end(), out);
  ORC_CATCH
}
int orc_getmge(orc_ctx*, double* outge, const uint64_t* terms, uint64_t K, uint64_t d, const double* basemat, uint64_t N, uint64_t M,
               const double* basescale, const uint64_t* knotptst, const double* basematge, uint64_t Mge, const uint64_t* gest,
               const uint64_t* hypmatch, uint64_t H) {
  ORC_TRY
  std::vector<mat> g;
  getmge_(g, to_umat(terms, K, d), to_mat(basemat, N, M), vec(basescale, basescale + N), kp(knotptst, d + 1), to_mat(basematge, N, Mge),
          kp(gest, H + 1), kp(hypmatch, H));
  for (uint64_t h = 0; h < g.size(); ++h) std::copy(g[h].a.begin(), g[h].a.end(), outge + h * N * K);
  ORC_CATCH
}

/* ---- lpdf family ---- */
int orc_loglik_gauss_create(orc_ctx*, orc_outermod* om, const uint64_t* terms, uint64_t K, const double* y,
                            const double* x, uint64_t N, orc_lpdf** out) {
  ORC_TRY
  auto* h = new orc_lpdf();
  h->kind = 0;
  h->p.reset(new loglik_gauss(om->om, to_umat(terms, K, om->om.d), vec(y, y + N), to_mat(x, N, om->om.d)));
  *out = h;
  ORC_CATCH
}
int orc_loglik_std_create(orc_ctx*, orc_outermod* om, const uint64_t* terms, uint64_t K, const double* y,
                          const double* x, uint64_t N, orc_lpdf** out) {
  ORC_TRY
  auto* h = new orc_lpdf();
  h->kind = 3;
  h->p.reset(new loglik_std(om->om, to_umat(terms, K, om->om.d), vec(y, y + N), to_mat(x, N, om->om.d)));
  *out = h;
  ORC_CATCH
}
static void put_cube(const std::vector<mat>& c, double* out, uint64_t* n) {
  *n = 0;
  for (const mat& m : c) { std::copy(m.a.begin(), m.a.end(), out + *n); *n += m.a.size(); }
}
int orc_lpdf_optnewton(orc_lpdf* l) { ORC_TRY l->p->optnewton(); ORC_CATCH }
int orc_lpdf_hess(orc_lpdf* l, double* out, uint64_t* n) { ORC_TRY mat h = l->p->hess(); std::copy(h.a.begin(), h.a.end(), out); *n = h.a.size(); ORC_CATCH }
int orc_lpdf_hessgradhyp(orc_lpdf* l, double* out, uint64_t* n) { ORC_TRY put_cube(l->p->hessgradhyp(), out, n); ORC_CATCH }
int orc_lpdf_hessgradpara(orc_lpdf* l, double* out, uint64_t* n) { ORC_TRY put_cube(l->p->hessgradpara(), out, n); ORC_CATCH }
int orc_loglik_gda_create(orc_ctx*, orc_outermod* om, const uint64_t* terms, uint64_t K, const double* y,
                          const double* x, uint64_t N, orc_lpdf** out) {
  ORC_TRY
  auto* h = new orc_lpdf();
  h->kind = 3;
  h->p.reset(new loglik_gda(om->om, to_umat(terms, K, om->om.d), vec(y, y + N), to_mat(x, N, om->om.d)));
  *out = h;
  ORC_CATCH
}
int orc_logpr_gauss_create(orc_ctx*, orc_outermod* om, const uint64_t* terms, uint64_t K, orc_lpdf** out) {
  ORC_TRY
  auto* h = new orc_lpdf();
  h->kind = 1;
  h->p.reset(new logpr_gauss(om->om, to_umat(terms, K, om->om.d)));
  *out = h;
  ORC_CATCH
}
int orc_lpdfvec_create(orc_lpdf* a, orc_lpdf* b, orc_lpdf** out) {
  ORC_TRY
  auto* h = new orc_lpdf();
  h->kind = 2;
  h->p.reset(new lpdfvec(*a->p, *b->p));
  *out = h;
  ORC_CATCH
}
int orc_lpdf_destroy(orc_lpdf* l) { delete l; return ORC_OK; }
int orc_lpdf_setnthreads(orc_lpdf* l, int k) { l->p->setnthreads(k); return ORC_OK; }
int orc_lpdf_update(orc_lpdf* l, const double* coeff, uint64_t K) { ORC_TRY l->p->update(vec(coeff, coeff + K)); ORC_CATCH }
int orc_lpdf_updateom(orc_lpdf* l) { ORC_TRY l->p->updateom(); ORC_CATCH }
int orc_lpdf_updatepara(orc_lpdf* l, const double* para, uint64_t n) { ORC_TRY l->p->updatepara(vec(para, para + n)); ORC_CATCH }
int orc_lpdf_updateterms(orc_lpdf* l, const uint64_t* terms, uint64_t K) {
  ORC_TRY l->p->updateterms(to_umat(terms, K, l->p->terms.nc)); ORC_CATCH
}
int orc_lpdf_optcg(orc_lpdf* l, double tol, uint64_t maxepch) { ORC_TRY l->p->optcg(tol, (unsigned)maxepch); ORC_CATCH }
int orc_lpdf_hessmult(orc_lpdf* l, const double* g, double* out) {
  ORC_TRY
  vec o = l->p->hessmult(vec(g, g + l->p->nterms));
  std::copy(o.begin(), o.end(), out);
  ORC_CATCH
}
int orc_lpdf_diaghess(orc_lpdf* l, double* out) { ORC_TRY vec o = l->p->diaghess(); std::copy(o.begin(), o.end(), out); ORC_CATCH }
int orc_lpdf_diaghessgradhyp(orc_lpdf* l, double* out) { ORC_TRY mat o = l->p->diaghessgradhyp(); std::copy(o.a.begin(), o.a.end(), out); ORC_CATCH }
int orc_lpdf_diaghessgradpara(orc_lpdf* l, double* out) { ORC_TRY mat o = l->p->diaghessgradpara(); std::copy(o.a.begin(), o.a.end(), out); ORC_CATCH }
int orc_lpdf_paralpdf(orc_lpdf* l, const double* para, uint64_t n, double* out) { ORC_TRY *out = l->p->paralpdf(vec(para, para + n)); ORC_CATCH }
int orc_lpdf_paralpdf_grad(orc_lpdf* l, const double* para, uint64_t n, double* out) {
  ORC_TRY vec g = l->p->paralpdf_grad(vec(para, para + n)); std::copy(g.begin(), g.end(), out); ORC_CATCH
}
int orc_lpdf_set_flag(orc_lpdf* l, const char* which, int value) {
  ORC_TRY
  const std::string w = which;
  if (w == "compute_val") l->p->compute_val = value;
  else if (w == "compute_grad") l->p->compute_grad = value;
  else if (w == "compute_gradhyp") l->p->compute_gradhyp = value;
  else if (w == "compute_gradpara") l->p->compute_gradpara = value;
  else if (w == "fullhess") l->p->fullhess = value;
  else if (w == "domarg") {
    auto* v = dynamic_cast<lpdfvec*>(l->p.get());
    if (!v) throw std::invalid_argument("domarg is a field of lpdfvec");
    v->domargadj = value;
  } else if (w == "dodiag") { /* R field name of loglik_gda::doda, interfaceR.cpp:748 */
    auto* v = dynamic_cast<loglik_gda*>(l->p.get());
    if (!v) throw std::invalid_argument("dodiag is a field of loglik_gda");
    v->doda = value; v->redostd = true;
  } else throw std::invalid_argument("unknown flag " + w);
  ORC_CATCH
}
int orc_lpdf_sizes(orc_lpdf* l, uint64_t* nterms, uint64_t* npara, uint64_t* nhyp, uint64_t* nrow) {
  *nterms = l->p->nterms; *npara = l->p->para.size(); *nhyp = l->p->nhyp(); *nrow = l->p->nrow();
  return ORC_OK;
}
int orc_lpdf_get(orc_lpdf* l, const char* which, double* out, uint64_t* n) {
  ORC_TRY
  const std::string w = which;
  vec v;
  if (w == "val") v = {l->p->val};
  else if (w == "grad") v = l->p->grad;
  else if (w == "gradhyp") v = l->p->gradhyp;
  else if (w == "gradpara") v = l->p->gradpara;
  else if (w == "coeff") v = l->p->coeff;
  else if (w == "para") v = l->p->para;
  else if (w == "totdiaghess") v = l->p->totdiaghess;
  else if (w == "tothess") v = l->p->tothess.a;
  else if (w == "cg_iters") v = {double(l->p->cg_iters)};
  else if (w == "yhat") {
    if (auto* g = dynamic_cast<loglik_gauss*>(l->p.get())) v = g->yhat;
    else if (auto* g2 = dynamic_cast<loglik_gda*>(l->p.get())) v = g2->yhat;
    else if (auto* g3 = dynamic_cast<loglik_std*>(l->p.get())) v = g3->yhat;
    else throw std::invalid_argument("yhat is a field of loglik_gauss / loglik_gda / loglik_std");
  } else if (w == "coeffsd") {
    auto* g = dynamic_cast<logpr_gauss*>(l->p.get());
    if (!g) throw std::invalid_argument("coeffsd is a field of logpr_gauss");
    v = g->coeffsd;
  } else throw std::invalid_argument("unknown field " + w);
  *n = v.size();
  if (out) std::copy(v.begin(), v.end(), out);
  ORC_CATCH
}
int orc_lpdf_set_coeff(orc_lpdf* l, const double* coeff, uint64_t K) { ORC_TRY l->p->coeff.assign(coeff, coeff + K); ORC_CATCH }

int orc_predictor_create(orc_lpdf* loglik, orc_predictor** out) {
  ORC_TRY
  auto* h = new orc_predictor();
  if (auto* g = dynamic_cast<loglik_gauss*>(loglik->p.get())) { h->p.reset(new pred_gauss(*g)); h->d = g->om.d; }
  else if (auto* g2 = dynamic_cast<loglik_gda*>(loglik->p.get())) { h->pg.reset(new pred_gda(*g2)); h->d = g2->om.d; }
  else if (auto* g3 = dynamic_cast<loglik_std*>(loglik->p.get())) { h->ps.reset(new predr_std(*g3)); h->d = g3->om.d; }
  else { delete h; throw std::invalid_argument("cannot produce a predictor from this obj."); }
  *out = h;
  ORC_CATCH
}
int orc_predictor_destroy(orc_predictor* p) { delete p; return ORC_OK; }
int orc_predictor_update(orc_predictor* p, const double* x, uint64_t N) {
  ORC_TRY if (p->p) p->p->update(to_mat(x, N, p->d)); else if (p->pg) p->pg->update(to_mat(x, N, p->d)); else p->ps->update(to_mat(x, N, p->d)); ORC_CATCH
}
int orc_predictor_mean(orc_predictor* p, double* out) { ORC_TRY vec o = p->p ? p->p->mean() : p->pg ? p->pg->mean() : p->ps->mean(); std::copy(o.begin(), o.end(), out); ORC_CATCH }
int orc_predictor_var(orc_predictor* p, double* out) { ORC_TRY vec o = p->p ? p->p->var() : p->pg ? p->pg->var() : p->ps->var(); std::copy(o.begin(), o.end(), out); ORC_CATCH }

} // extern "C"
