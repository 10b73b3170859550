/*
 * integration/linalg_shim.cpp -- the reference-side binding: a REPLACEMENT for the reference's src/linalg.cpp that
 * keeps every declaration of src/linalg.h:9-58 and forwards each function to the C ABI of libouterbase_b200.so
 * (include/outerbase_b200.h, "stateless linalg.h seam").  A maintainer drops this file in place of src/linalg.cpp and
 * adds the two Makevars lines of INTEGRATION.md; nothing else of the package changes -- outerbase, loglik_gauss,
 * lpdfvec::optcg ... (src/modandbase.cpp, src/fit.cpp) call these functions exactly as before.
 *
 * Armadillo objects are column-major fp64 / 64-bit uword, which is the ABI's layout: arguments pass through as raw
 * pointers, outputs are sized as the reference sizes them (src/linalg.cpp:107, 230-233, 308, 399-402, 532, 588, 690).
 * vertpl / chunksize / loopsize / num_threads (the OpenMP tuning of src/modandbase.cpp:504-513) are accepted and
 * ignored.  Errors: a non-zero status becomes a C++ exception, i.e. an R error through Rcpp, as before.
 *
 * In THIS repository the file is compiled (make -C oracle refgpu) against the reference's unmodified
 * modandbase.cpp / covfuncs.cpp / fit.cpp and the Armadillo-subset shim into oracle/_ref/libob_refgpu.so, and
 * tests/test_integration_shim.py runs the reference's own classes on the GPU through it -- the closest thing to
 * re-attaching the Rcpp module that an image without R allows.
 */
#include "customconfig.h"
#include <RcppArmadillo.h>

#include <stdexcept>
#include <string>

#include "outerbase_b200.h"

using namespace arma;

#include "linalg.h"

namespace {

ob_ctx* ctx() { /* one context per process (per R session), created at first use */
  static ob_ctx* c = nullptr;
  if (!c && ob_ctx_create(0, &c) != OB_OK) throw std::runtime_error(std::string("outerbase_b200: ") + ob_last_error());
  return c;
}
void chk(int rc) {
  if (rc == OB_OK) return;
  if (rc == OB_ERR_INVALID) throw std::range_error(ob_last_error()); /* the reference's own exception classes */
  throw std::runtime_error(ob_last_error());
}
const uint64_t* u64p(const umat& m) { return reinterpret_cast<const uint64_t*>(m.memptr()); }

} // namespace

/* src/linalg.cpp:102-131 */
void prodmm_(vec& out, const umat& terms, const vec& a, const mat& basemat, const vec& basescale, const uvec& knotptst,
             bool, uword, uword, int) {
  if (out.n_elem != basemat.n_rows) out.set_size(basemat.n_rows);
  chk(ob_prodmm_vec(ctx(), out.memptr(), u64p(terms), terms.n_rows, terms.n_cols, a.memptr(), basemat.memptr(), basemat.n_rows,
                    basemat.n_cols, basescale.memptr(), u64p(knotptst)));
}
/* src/linalg.cpp:527-557 */
void prodmm_(mat& out, const umat& terms, const mat& a, const mat& basemat, const vec& basescale, const uvec& knotptst,
             bool, uword, uword, int) {
  if (out.n_rows != basemat.n_rows || out.n_cols != a.n_cols) out.set_size(basemat.n_rows, a.n_cols);
  chk(ob_prodmm_mat(ctx(), out.memptr(), u64p(terms), terms.n_rows, terms.n_cols, a.memptr(), a.n_cols, basemat.memptr(),
                    basemat.n_rows, basemat.n_cols, basescale.memptr(), u64p(knotptst)));
}
/* src/linalg.cpp:303-355 */
void tprodmm_(vec& out, const umat& terms, const vec& a, const mat& basemat, const vec& basescale, const uvec& knotptst,
              bool, uword, uword, int) {
  if (out.n_elem != terms.n_rows) out.set_size(terms.n_rows);
  chk(ob_tprodmm_vec(ctx(), out.memptr(), u64p(terms), terms.n_rows, terms.n_cols, a.memptr(), basemat.memptr(), basemat.n_rows,
                     basemat.n_cols, basescale.memptr(), u64p(knotptst)));
}
/* src/linalg.cpp:583-637 */
void tprodmm_(mat& out, const umat& terms, const mat& a, const mat& basemat, const vec& basescale, const uvec& knotptst,
              bool, uword, uword, int) {
  if (out.n_rows != terms.n_rows || out.n_cols != a.n_cols) out.set_size(terms.n_rows, a.n_cols);
  chk(ob_tprodmm_mat(ctx(), out.memptr(), u64p(terms), terms.n_rows, terms.n_cols, a.memptr(), a.n_cols, basemat.memptr(),
                     basemat.n_rows, basemat.n_cols, basescale.memptr(), u64p(knotptst)));
}
/* src/linalg.cpp:225-277 */
void prodmmge_(vec& out, mat& outge, const umat& terms, const vec& a, const mat& basemat, const vec& basescale,
               const uvec& knotptst, const mat& basematge, const uvec& gest, const uvec& hypmatch, bool, uword, uword, int) {
  if (out.n_elem != basemat.n_rows) out.set_size(basemat.n_rows);
  if (outge.n_rows != basemat.n_rows || outge.n_cols != gest.n_elem - 1) outge.set_size(basemat.n_rows, gest.n_elem - 1);
  chk(ob_prodmmge(ctx(), out.memptr(), outge.memptr(), u64p(terms), terms.n_rows, terms.n_cols, a.memptr(), basemat.memptr(),
                  basemat.n_rows, basemat.n_cols, basescale.memptr(), u64p(knotptst), basematge.memptr(), basematge.n_cols, u64p(gest),
                  u64p(hypmatch), hypmatch.n_elem));
}
/* src/linalg.cpp:394-471 */
void tprodmmge_(vec& out, mat& outge, const umat& terms, const vec& a, const mat& basemat, const vec& basescale,
                const uvec& knotptst, const mat& basematge, const uvec& gest, const uvec& hypmatch, bool, uword, uword, int) {
  if (out.n_elem != terms.n_rows) out.set_size(terms.n_rows);
  if (outge.n_rows != terms.n_rows || outge.n_cols != gest.n_elem - 1) outge.set_size(terms.n_rows, gest.n_elem - 1);
  chk(ob_tprodmmge(ctx(), out.memptr(), outge.memptr(), u64p(terms), terms.n_rows, terms.n_cols, a.memptr(), basemat.memptr(),
                   basemat.n_rows, basemat.n_cols, basescale.memptr(), u64p(knotptst), basematge.memptr(), basematge.n_cols, u64p(gest),
                   u64p(hypmatch), hypmatch.n_elem));
}
/* src/linalg.cpp:685-715 */
void getm_(mat& out, const umat& terms, const mat& basemat, const vec& basescale, const uvec& knotptst, bool, uword, uword, int) {
  if (out.n_rows != basemat.n_rows || out.n_cols != terms.n_rows) out.set_size(basemat.n_rows, terms.n_rows);
  chk(ob_getm(ctx(), out.memptr(), u64p(terms), terms.n_rows, terms.n_cols, basemat.memptr(), basemat.n_rows, basemat.n_cols,
              basescale.memptr(), u64p(knotptst)));
}
/* src/linalg.cpp:778-822 -- explicit d(Phi)/d(hyp) cube, used by loglik_std only.  The ABI follows the unchunked branch for
 * every shape (the row-chunk branch, :788-810, cannot work upstream: dogetmge_ resizes the whole cube to one chunk and
 * the chunk is never copied back) */
void getmge_(cube& outge, const umat& terms, const mat& basemat, const vec& basescale, const uvec& knotptst, const mat& basematge,
             const uvec& gest, const uvec& hypmatch, bool, uword, uword, int) {
  const uword H = gest.n_elem - 1;
  if (outge.n_rows != basemat.n_rows || outge.n_cols != terms.n_rows || outge.n_slices != H) outge.set_size(basemat.n_rows, terms.n_rows, H);
  chk(ob_getmge(ctx(), outge.memptr(), u64p(terms), terms.n_rows, terms.n_cols, basemat.memptr(), basemat.n_rows, basemat.n_cols,
                basescale.memptr(), u64p(knotptst), basematge.memptr(), basematge.n_cols, u64p(gest), u64p(hypmatch), H));
}
