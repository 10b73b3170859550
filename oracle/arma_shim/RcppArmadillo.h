/*
 * RcppArmadillo.h -- header-only SUBSET of Armadillo, sufficient to compile the UNMODIFIED reference sources
 *   /root/reference/src/{linalg,covfuncs,modandbase,fit}.cpp (+ src/lpdfs/ *.cpp, which fit.cpp #includes)
 * in this image, where neither R, Rcpp nor Armadillo exist.  TEST INFRASTRUCTURE ONLY: it is used by the
 * recipe `make -C oracle ref` to build oracle/_ref/libob_ref.so, which pins the CPU oracle (tests/test_oracle_ref.py)
 * and serves as bench.py's `--impl reference` arm.  Nothing under outerbase_b200/ includes or links it.
 *
 * It is NOT Armadillo: original code, eager evaluation (every operator returns a matrix), written from the library's
 * documented behaviour.  What matters for parity is the floating-point ORDER of the reductions, restated from
 * Armadillo's published algorithms (the dependency is unpinned upstream: DESCRIPTION:16 `LinkingTo: RcppArmadillo`):
 *   - element-wise operators: one rounding per operator per element, like Armadillo's expression templates
 *     (build with -ffp-contract=off; x86-64 baseline has no FMA anyway);
 *   - accu / sum / mean of contiguous data: TWO accumulators over even / odd elements, acc1 + acc2 (arrayops::accumulate);
 *   - var: op_var::direct_var (mean by accumulate, two-accumulator squared deviations with the acc3 correction);
 *   - dot: two accumulators for n <= 32 (op_dot::direct_dot_arma), BLAS ddot above;
 *   - matrix products (gemv / gemm): BLAS.
 * The BLAS Armadillo would call is R's, equally unpinned.  ARMA_SHIM_BLAS selects the restated flavour:
 *     1 (default)  netlib reference BLAS, R's bundled libRblas: ddot / dgemv('T') / dgemm accumulate k-ascending in ONE
 *                  accumulator; dgemv('N') and dgemm('N','N') are axpy sweeps (also k-ascending per output element);
 *     0            Armadillo built with ARMA_DONT_USE_BLAS: every inner product is the two-accumulator direct_dot_arma.
 *   - eig_sym (LAPACK dsyevd upstream) and solve / inv (dgesv / dgetri) have no restatable order: eig_sym is routed to
 *     `arma_shim_eig_sym`, a hook the including program defines (oracle/ref_capi.cpp forwards it to the oracle's cyclic
 *     Jacobi so both sides share ONE eigenbasis, SURVEY 7 "hard parts"); solve / inv are partial-pivot elimination.
 *   - shuffle (R's RNG upstream, modandbase.cpp:408) is routed to `arma_shim_shuffle` (default: identity, i.e. the
 *     lowest-index tie-break the oracle and the product use by default).
 * Bounds / size checks that Armadillo performs with ARMA_NO_DEBUG unset (src/customconfig.h:3) throw std::logic_error.
 */
#ifndef OB_ARMA_SHIM_H
#define OB_ARMA_SHIM_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <limits>
#include <new>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#ifndef ARMA_SHIM_BLAS
#define ARMA_SHIM_BLAS 1
#endif

namespace Rcpp {}

namespace arma {

typedef std::uint64_t uword;
typedef std::int64_t sword;

struct datum {
  static constexpr double inf = std::numeric_limits<double>::infinity();
  static constexpr double nan = std::numeric_limits<double>::quiet_NaN();
  static constexpr double pi = 3.14159265358979323846;
};

struct arma_tag {};
template <class T> struct Mat;
template <class T> struct Col;
template <class T> struct Row;
template <class T> struct Cube;
template <class T, int VS> struct subview;
template <class T> struct diagview;
template <class P, int DIM> struct each_proxy;

template <class X> using is_arma = std::is_base_of<arma_tag, typename std::decay<X>::type>;
template <class X> using enable_arma = typename std::enable_if<is_arma<X>::value, int>::type;
template <class X> using enable_scalar = typename std::enable_if<std::is_arithmetic<X>::value, int>::type;

template <class T, int VS> struct result_of { typedef Mat<T> type; };
template <class T> struct result_of<T, 1> { typedef Col<T> type; };
template <class T> struct result_of<T, 2> { typedef Row<T> type; };

template <class A, class B> struct promote { typedef typename std::conditional<std::is_floating_point<A>::value || std::is_floating_point<B>::value, double,
    typename std::conditional<std::is_same<A, B>::value, A, sword>::type>::type type; };

inline void shim_check(bool ok, const char* what) { if (!ok) throw std::logic_error(what); }

/* ------------------------------------------------------------------ reductions in Armadillo's order */
template <class T> inline T accumulate2(const T* src, uword n) { /* arrayops::accumulate */
  T acc1 = T(0), acc2 = T(0);
  uword j;
  for (j = 1; j < n; j += 2) { acc1 += (*src); src++; acc2 += (*src); src++; }
  if ((j - 1) < n) acc1 += (*src);
  return acc1 + acc2;
}
template <class T> inline T dot_arma(uword n, const T* A, const T* B) { /* op_dot::direct_dot_arma */
  T val1 = T(0), val2 = T(0);
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { val1 += A[i] * B[i]; val2 += A[j] * B[j]; }
  if (i < n) val1 += A[i] * B[i];
  return val1 + val2;
}
template <class T> inline T dot_blas(uword n, const T* A, const T* B) { /* netlib ddot, unit stride: left-to-right sum */
#if ARMA_SHIM_BLAS
  T s = T(0);
  for (uword i = 0; i < n; ++i) s = s + A[i] * B[i];
  return s;
#else
  return dot_arma(n, A, B);
#endif
}
template <class T> inline T dot_direct(uword n, const T* A, const T* B) { return n <= 32u ? dot_arma(n, A, B) : dot_blas(n, A, B); }

/* ------------------------------------------------------------------ Mat */
template <class T> struct Mat : arma_tag {
  typedef T elem_type;
  static constexpr int vec_state = 0;
  uword n_rows = 0, n_cols = 0, n_elem = 0;
  T* mem = nullptr;
  bool owns = true;

  Mat() {}
  Mat(uword r, uword c) { init(r, c); std::fill(mem, mem + n_elem, T(0)); }
  Mat(const Mat& o) { init(o.n_rows, o.n_cols); if (n_elem) std::memcpy(mem, o.mem, n_elem * sizeof(T)); }
  Mat(Mat&& o) noexcept { steal(o); }
  /* non-owning alias (Armadillo's auxiliary-memory constructor) */
  Mat(T* aux, uword r, uword c, bool /*copy_aux_mem*/, bool /*strict*/) : n_rows(r), n_cols(c), n_elem(r * c), mem(aux), owns(false) {}
  Mat(const T* src, uword r, uword c) { init(r, c); if (n_elem) std::memcpy(mem, src, n_elem * sizeof(T)); }
  template <class X, int VS> Mat(const subview<X, VS>& s);
  template <class X> Mat(const diagview<X>& s);
  Mat(const Cube<T>& c);
  template <class U, typename std::enable_if<!std::is_same<U, T>::value, int>::type = 0> explicit Mat(const Mat<U>& o) {
    init(o.n_rows, o.n_cols);
    for (uword i = 0; i < n_elem; ++i) mem[i] = T(o.mem[i]);
  }
  ~Mat() { if (owns && mem) ::operator delete(mem); }

  void init(uword r, uword c) {
    n_rows = r; n_cols = c; n_elem = r * c; owns = true;
    mem = n_elem ? static_cast<T*>(::operator new(n_elem * sizeof(T))) : nullptr;
  }
  void steal(Mat& o) {
    n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem; mem = o.mem; owns = o.owns;
    o.mem = nullptr; o.n_rows = o.n_cols = o.n_elem = 0; o.owns = true;
  }
  Mat& operator=(const Mat& o) {
    if (this == &o) return *this;
    set_size(o.n_rows, o.n_cols);
    if (n_elem) std::memmove(mem, o.mem, n_elem * sizeof(T));
    return *this;
  }
  Mat& operator=(Mat&& o) {
    if (this == &o) return *this;
    /* same shape: write into the existing memory, as Armadillo's expression evaluation does -- the reference relies on
     * it: every thread of outerbase::build's parallel region executes `basescalesq = square(basescale)`
     * (modandbase.cpp:623), a benign race only as long as nobody reallocates */
    if (n_rows == o.n_rows && n_cols == o.n_cols) { if (n_elem) std::memcpy(mem, o.mem, n_elem * sizeof(T)); return *this; }
    if (owns && o.owns) { if (mem) ::operator delete(mem); steal(o); return *this; }
    return (*this = static_cast<const Mat&>(o));
  }
  Mat& operator=(T val) { set_size(1, 1); mem[0] = val; return *this; } /* Mat::operator=(eT): 1x1 */
  template <class X, int VS> Mat& operator=(const subview<X, VS>& s) { Mat t(s); return (*this = std::move(t)); }
  template <class X> Mat& operator=(const diagview<X>& s) { Mat t(s); return (*this = std::move(t)); }

  void set_size(uword r, uword c) {
    if (r == n_rows && c == n_cols) return;
    if (!owns) { shim_check(r * c == n_elem, "set_size: cannot resize a matrix that aliases foreign memory"); n_rows = r; n_cols = c; return; }
    if (r * c == n_elem) { n_rows = r; n_cols = c; return; }
    if (mem) ::operator delete(mem);
    init(r, c);
  }
  void set_size(uword n) { vec_state_size(n); }
  virtual void vec_state_size(uword n) { set_size(n, 1); }
  template <class X> void copy_size(const X& o) { set_size(o.n_rows, o.n_cols); }
  void resize(uword r, uword c) { /* keeps the overlapping block */
    if (r == n_rows && c == n_cols) return;
    Mat t(r, c);
    for (uword j = 0; j < std::min(c, n_cols); ++j) for (uword i = 0; i < std::min(r, n_rows); ++i) t.at(i, j) = at(i, j);
    *this = std::move(t);
  }
  void resize(uword n) { if (n_cols == 1 || n_elem == 0 || vec_is_col()) resize(n, 1); else resize(1, n); }
  virtual bool vec_is_col() const { return true; }
  Mat& zeros() { std::fill(mem, mem + n_elem, T(0)); return *this; }
  Mat& zeros(uword n) { set_size(n); return zeros(); }
  Mat& zeros(uword r, uword c) { set_size(r, c); return zeros(); }
  Mat& ones() { std::fill(mem, mem + n_elem, T(1)); return *this; }
  Mat& ones(uword n) { set_size(n); return ones(); }
  Mat& fill(T v) { std::fill(mem, mem + n_elem, v); return *this; }
  T* memptr() { return mem; }
  const T* memptr() const { return mem; }
  T* colptr(uword c) { return mem + c * n_rows; }
  const T* colptr(uword c) const { return mem + c * n_rows; }
  bool is_empty() const { return n_elem == 0; }
  bool is_finite() const { for (uword i = 0; i < n_elem; ++i) if (!std::isfinite(double(mem[i]))) return false; return true; }

  T& at(uword i, uword j) { return mem[i + j * n_rows]; }
  const T& at(uword i, uword j) const { return mem[i + j * n_rows]; }
  T& at(uword i) { return mem[i]; }
  const T& at(uword i) const { return mem[i]; }
  T& operator()(uword i, uword j) { shim_check(i < n_rows && j < n_cols, "Mat::operator(): index out of bounds"); return mem[i + j * n_rows]; }
  const T& operator()(uword i, uword j) const { shim_check(i < n_rows && j < n_cols, "Mat::operator(): index out of bounds"); return mem[i + j * n_rows]; }
  T& operator()(uword i) { shim_check(i < n_elem, "Mat::operator(): index out of bounds"); return mem[i]; }
  const T& operator()(uword i) const { shim_check(i < n_elem, "Mat::operator(): index out of bounds"); return mem[i]; }
  T& operator[](uword i) { return mem[i]; }
  const T& operator[](uword i) const { return mem[i]; }

  /* views */
  subview<T, 1> col(uword c) { return subview<T, 1>(*this, 0, c, n_rows, 1); }
  const subview<T, 1> col(uword c) const { return subview<T, 1>(const_cast<Mat&>(*this), 0, c, n_rows, 1); }
  subview<T, 2> row(uword r) { return subview<T, 2>(*this, r, 0, 1, n_cols); }
  const subview<T, 2> row(uword r) const { return subview<T, 2>(const_cast<Mat&>(*this), r, 0, 1, n_cols); }
  subview<T, 0> cols(uword a, uword b) { return subview<T, 0>(*this, 0, a, n_rows, b + 1 - a); }
  const subview<T, 0> cols(uword a, uword b) const { return subview<T, 0>(const_cast<Mat&>(*this), 0, a, n_rows, b + 1 - a); }
  subview<T, 0> rows(uword a, uword b) { return subview<T, 0>(*this, a, 0, b + 1 - a, n_cols); }
  const subview<T, 0> rows(uword a, uword b) const { return subview<T, 0>(const_cast<Mat&>(*this), a, 0, b + 1 - a, n_cols); }
  subview<T, 0> head_rows(uword n) { return subview<T, 0>(*this, 0, 0, n, n_cols); }
  const subview<T, 0> head_rows(uword n) const { return subview<T, 0>(const_cast<Mat&>(*this), 0, 0, n, n_cols); }
  subview<T, 0> submat(uword r1, uword c1, uword r2, uword c2) { return subview<T, 0>(*this, r1, c1, r2 + 1 - r1, c2 + 1 - c1); }
  const subview<T, 0> submat(uword r1, uword c1, uword r2, uword c2) const { return subview<T, 0>(const_cast<Mat&>(*this), r1, c1, r2 + 1 - r1, c2 + 1 - c1); }
  Col<T> unsafe_col(uword c);
  const Col<T> unsafe_col(uword c) const;
  diagview<T> diag() { return diagview<T>(*this); }
  const diagview<T> diag() const { return diagview<T>(const_cast<Mat&>(*this)); }
  each_proxy<Mat<T>, 0> each_col() { return each_proxy<Mat<T>, 0>(*this); }
  each_proxy<Mat<T>, 1> each_row() { return each_proxy<Mat<T>, 1>(*this); }
  Mat t() const { Mat o(n_cols, n_rows); for (uword j = 0; j < n_cols; ++j) for (uword i = 0; i < n_rows; ++i) o.mem[j + i * n_cols] = mem[i + j * n_rows]; return o; }
  Row<T> as_row() const;
  Col<T> as_col() const;
  template <class I> Col<T> elem(const I& idx) const;

  T min() const { shim_check(n_elem > 0, "min(): object has no elements"); return *std::min_element(mem, mem + n_elem); }
  T max() const { shim_check(n_elem > 0, "max(): object has no elements"); return *std::max_element(mem, mem + n_elem); }
  uword index_max() const { shim_check(n_elem > 0, "index_max(): object has no elements"); return uword(std::max_element(mem, mem + n_elem) - mem); }
  uword index_min() const { shim_check(n_elem > 0, "index_min(): object has no elements"); return uword(std::min_element(mem, mem + n_elem) - mem); }

  /* compound assignment with any matrix-like operand or a scalar */
  template <class X, enable_arma<X> = 0> Mat& operator+=(const X& x);
  template <class X, enable_arma<X> = 0> Mat& operator-=(const X& x);
  template <class X, enable_arma<X> = 0> Mat& operator%=(const X& x);
  template <class X, enable_arma<X> = 0> Mat& operator/=(const X& x);
  template <class X, enable_arma<X> = 0> Mat& operator*=(const X& x);
  template <class S, enable_scalar<S> = 0> Mat& operator+=(S s) { for (uword i = 0; i < n_elem; ++i) mem[i] += T(s); return *this; }
  template <class S, enable_scalar<S> = 0> Mat& operator-=(S s) { for (uword i = 0; i < n_elem; ++i) mem[i] -= T(s); return *this; }
  template <class S, enable_scalar<S> = 0> Mat& operator*=(S s) { for (uword i = 0; i < n_elem; ++i) mem[i] *= T(s); return *this; }
  template <class S, enable_scalar<S> = 0> Mat& operator/=(S s) { for (uword i = 0; i < n_elem; ++i) mem[i] /= T(s); return *this; }
};

template <class T> struct Col : Mat<T> {
  static constexpr int vec_state = 1;
  Col() { this->n_cols = 1; }
  explicit Col(uword n) : Mat<T>(n, 1) {}
  Col(const Col& o) : Mat<T>(o) {}
  Col(Col&& o) noexcept : Mat<T>(std::move(o)) {}
  Col(const Mat<T>& o) : Mat<T>(o) { as_vec(); }
  Col(Mat<T>&& o) : Mat<T>(std::move(o)) { as_vec(); }
  Col(T* aux, uword n, bool c, bool s) : Mat<T>(aux, n, 1, c, s) {}
  Col(const T* src, uword n) : Mat<T>(src, n, 1) {}
  Col(std::initializer_list<T> l) : Mat<T>(l.size(), 1) { std::copy(l.begin(), l.end(), this->mem); }
  template <class X, int VS> Col(const subview<X, VS>& s) : Mat<T>(s) { as_vec(); }
  template <class X> Col(const diagview<X>& s) : Mat<T>(s) {}
  void as_vec() {
    if (this->n_cols != 1) { shim_check(this->n_rows == 1 || this->n_elem == 0, "Col: incompatible matrix dimensions"); this->n_rows = this->n_elem; this->n_cols = 1; }
  }
  Col& operator=(const Col& o) { Mat<T>::operator=(o); return *this; }
  Col& operator=(Col&& o) { Mat<T>::operator=(std::move(o)); return *this; }
  Col& operator=(const Mat<T>& o) { Mat<T>::operator=(o); as_vec(); return *this; }
  Col& operator=(Mat<T>&& o) { Mat<T>::operator=(std::move(o)); as_vec(); return *this; }
  Col& operator=(T v) { Mat<T>::operator=(v); return *this; }
  template <class X, int VS> Col& operator=(const subview<X, VS>& s) { Mat<T>::operator=(s); as_vec(); return *this; }
  template <class X> Col& operator=(const diagview<X>& s) { Mat<T>::operator=(s); return *this; }
  void vec_state_size(uword n) override { Mat<T>::set_size(n, 1); }
  using Mat<T>::rows;
  subview<T, 1> rows(uword a, uword b) { return subview<T, 1>(*this, a, 0, b + 1 - a, 1); }
  const subview<T, 1> rows(uword a, uword b) const { return subview<T, 1>(const_cast<Col&>(*this), a, 0, b + 1 - a, 1); }
  subview<T, 1> subvec(uword a, uword b) { shim_check(a <= b + 1 && b < this->n_elem, "Col::subvec(): indices out of bounds"); return subview<T, 1>(*this, a, 0, b + 1 - a, 1); }
  const subview<T, 1> subvec(uword a, uword b) const { shim_check(a <= b + 1 && b < this->n_elem, "Col::subvec(): indices out of bounds"); return subview<T, 1>(const_cast<Col&>(*this), a, 0, b + 1 - a, 1); }
  subview<T, 1> head(uword n) { return subview<T, 1>(*this, 0, 0, n, 1); }
  const subview<T, 1> head(uword n) const { return subview<T, 1>(const_cast<Col&>(*this), 0, 0, n, 1); }
  subview<T, 1> row(uword r) { return subview<T, 1>(*this, r, 0, 1, 1); }
  T* begin() { return this->mem; }
  T* end() { return this->mem + this->n_elem; }
  const T* begin() const { return this->mem; }
  const T* end() const { return this->mem + this->n_elem; }
};

template <class T> struct Row : Mat<T> {
  static constexpr int vec_state = 2;
  Row() { this->n_rows = 1; }
  explicit Row(uword n) : Mat<T>(1, n) {}
  Row(const Row& o) : Mat<T>(o) {}
  Row(Row&& o) noexcept : Mat<T>(std::move(o)) {}
  Row(const Mat<T>& o) : Mat<T>(o) { as_vec(); }
  Row(Mat<T>&& o) : Mat<T>(std::move(o)) { as_vec(); }
  template <class X, int VS> Row(const subview<X, VS>& s) : Mat<T>(s) { as_vec(); }
  void as_vec() {
    if (this->n_rows != 1) { shim_check(this->n_cols == 1 || this->n_elem == 0, "Row: incompatible matrix dimensions"); this->n_cols = this->n_elem; this->n_rows = 1; }
  }
  Row& operator=(const Row& o) { Mat<T>::operator=(o); return *this; }
  Row& operator=(Row&& o) { Mat<T>::operator=(std::move(o)); return *this; }
  Row& operator=(const Mat<T>& o) { Mat<T>::operator=(o); as_vec(); return *this; }
  Row& operator=(Mat<T>&& o) { Mat<T>::operator=(std::move(o)); as_vec(); return *this; }
  void vec_state_size(uword n) override { Mat<T>::set_size(1, n); }
  bool vec_is_col() const override { return false; }
};

typedef Mat<double> mat; typedef Col<double> vec; typedef Col<double> colvec; typedef Row<double> rowvec; typedef Cube<double> cube;
typedef Mat<uword> umat; typedef Col<uword> uvec; typedef Row<uword> urowvec;
typedef Mat<sword> imat; typedef Col<sword> ivec; typedef Row<sword> irowvec;

template <class T> Col<T> Mat<T>::unsafe_col(uword c) { shim_check(c < n_cols, "Mat::unsafe_col(): index out of bounds"); return Col<T>(colptr(c), n_rows, false, true); }
template <class T> const Col<T> Mat<T>::unsafe_col(uword c) const { shim_check(c < n_cols, "Mat::unsafe_col(): index out of bounds"); return Col<T>(const_cast<T*>(colptr(c)), n_rows, false, true); }
template <class T> Row<T> Mat<T>::as_row() const { Row<T> o(n_elem); if (n_elem) std::memcpy(o.mem, mem, n_elem * sizeof(T)); return o; }
template <class T> Col<T> Mat<T>::as_col() const { Col<T> o(n_elem); if (n_elem) std::memcpy(o.mem, mem, n_elem * sizeof(T)); return o; }

/* ------------------------------------------------------------------ views */
template <class T, int VS> struct subview : arma_tag {
  typedef T elem_type;
  static constexpr int vec_state = VS;
  Mat<T>& m;
  uword r0, c0, n_rows, n_cols, n_elem;
  subview(Mat<T>& m_, uword r0_, uword c0_, uword nr, uword nc) : m(m_), r0(r0_), c0(c0_), n_rows(nr), n_cols(nc), n_elem(nr * nc) {
    shim_check(r0 + nr <= m.n_rows && c0 + nc <= m.n_cols, "subview: indices out of bounds or incorrectly used");
  }
  subview(const subview&) = default;
  T& at(uword i, uword j) { return m.mem[(r0 + i) + (c0 + j) * m.n_rows]; }
  const T& at(uword i, uword j) const { return m.mem[(r0 + i) + (c0 + j) * m.n_rows]; }
  T& operator()(uword i, uword j) { shim_check(i < n_rows && j < n_cols, "subview::operator(): index out of bounds"); return at(i, j); }
  T& operator()(uword i) { return VS == 2 ? at(0, i) : at(i % n_rows, i / n_rows); }
  const T& operator()(uword i) const { return VS == 2 ? at(0, i) : at(i % n_rows, i / n_rows); }
  T& operator[](uword i) { return (*this)(i); }
  const T& operator[](uword i) const { return (*this)(i); }
  typename result_of<T, VS>::type eval() const {
    typename result_of<T, VS>::type o;
    static_cast<Mat<T>&>(o).set_size(n_rows, n_cols);
    for (uword j = 0; j < n_cols; ++j) if (n_rows) std::memcpy(o.mem + j * n_rows, &at(0, j), n_rows * sizeof(T));
    return o;
  }
  template <class F> void apply(const Mat<T>& x, F f) {
    shim_check(x.n_rows == n_rows && x.n_cols == n_cols, "subview: incompatible matrix dimensions");
    for (uword j = 0; j < n_cols; ++j) for (uword i = 0; i < n_rows; ++i) f(at(i, j), x.mem[i + j * n_rows]);
  }
  /* the right-hand side is materialised first, so views of the same parent may overlap */
  template <class X, enable_arma<X> = 0> subview& operator=(const X& x);
  subview& operator=(const subview& x);
  template <class X, enable_arma<X> = 0> subview& operator+=(const X& x);
  template <class X, enable_arma<X> = 0> subview& operator-=(const X& x);
  template <class X, enable_arma<X> = 0> subview& operator%=(const X& x);
  template <class X, enable_arma<X> = 0> subview& operator/=(const X& x);
  subview& zeros() { for (uword j = 0; j < n_cols; ++j) for (uword i = 0; i < n_rows; ++i) at(i, j) = T(0); return *this; }
  subview& ones() { for (uword j = 0; j < n_cols; ++j) for (uword i = 0; i < n_rows; ++i) at(i, j) = T(1); return *this; }
  subview& fill(T v) { for (uword j = 0; j < n_cols; ++j) for (uword i = 0; i < n_rows; ++i) at(i, j) = v; return *this; }
  subview<T, VS> rows(uword a, uword b) { return subview<T, VS>(m, r0 + a, c0, b + 1 - a, n_cols); }
  const subview<T, VS> rows(uword a, uword b) const { return subview<T, VS>(m, r0 + a, c0, b + 1 - a, n_cols); }
  subview<T, VS> head(uword n) { return VS == 2 ? subview<T, VS>(m, r0, c0, 1, n) : subview<T, VS>(m, r0, c0, n, n_cols); }
  const subview<T, VS> head(uword n) const { return VS == 2 ? subview<T, VS>(m, r0, c0, 1, n) : subview<T, VS>(m, r0, c0, n, n_cols); }
  subview<T, 1> col(uword c) { return subview<T, 1>(m, r0, c0 + c, n_rows, 1); }
  each_proxy<subview<T, VS>, 0> each_col() { return each_proxy<subview<T, VS>, 0>(*this); }
  each_proxy<subview<T, VS>, 1> each_row() { return each_proxy<subview<T, VS>, 1>(*this); }
  Mat<T> t() const { return static_cast<const Mat<T>&>(eval()).t(); }
  Row<T> as_row() const { return static_cast<const Mat<T>&>(eval()).as_row(); }
  Col<T> as_col() const { return static_cast<const Mat<T>&>(eval()).as_col(); }
  T min() const { return eval().min(); }
  T max() const { return eval().max(); }
};

template <class T> struct diagview : arma_tag {
  typedef T elem_type;
  static constexpr int vec_state = 1;
  Mat<T>& m;
  uword n_rows, n_cols = 1, n_elem;
  explicit diagview(Mat<T>& m_) : m(m_), n_rows(std::min(m_.n_rows, m_.n_cols)), n_elem(std::min(m_.n_rows, m_.n_cols)) {}
  T& at(uword i) { return m.mem[i + i * m.n_rows]; }
  const T& at(uword i) const { return m.mem[i + i * m.n_rows]; }
  Col<T> eval() const { Col<T> o(n_elem); for (uword i = 0; i < n_elem; ++i) o.mem[i] = at(i); return o; }
  diagview& zeros() { for (uword i = 0; i < n_elem; ++i) at(i) = T(0); return *this; }
  template <class X, enable_arma<X> = 0> diagview& operator=(const X& x);
};

/* as_mat: a plain matrix for any operand (identity for matrices, a copy for views) */
template <class T> inline const Mat<T>& as_mat(const Mat<T>& x) { return x; }
template <class T, int VS> inline typename result_of<T, VS>::type as_mat(const subview<T, VS>& s) { return s.eval(); }
template <class T> inline Col<T> as_mat(const diagview<T>& s) { return s.eval(); }

template <class T> template <class X, int VS> Mat<T>::Mat(const subview<X, VS>& s) {
  static_assert(std::is_same<X, T>::value, "subview element type");
  init(s.n_rows, s.n_cols);
  for (uword j = 0; j < n_cols; ++j) if (n_rows) std::memcpy(mem + j * n_rows, &s.at(0, j), n_rows * sizeof(T));
}
template <class T> template <class X> Mat<T>::Mat(const diagview<X>& s) { init(s.n_elem, 1); for (uword i = 0; i < n_elem; ++i) mem[i] = s.at(i); }

#define SHIM_MAT_COMPOUND(OP, EXPR)                                                                               \
  template <class T> template <class X, enable_arma<X>> Mat<T>& Mat<T>::operator OP(const X& x) {                 \
    const auto& b = as_mat(x);                                                                                    \
    shim_check(b.n_rows == n_rows && b.n_cols == n_cols, "element-wise operation: incompatible matrix dimensions"); \
    for (uword i = 0; i < n_elem; ++i) { T& l = mem[i]; const T r = T(b.mem[i]); EXPR; }                          \
    return *this;                                                                                                 \
  }
SHIM_MAT_COMPOUND(+=, l += r)
SHIM_MAT_COMPOUND(-=, l -= r)
SHIM_MAT_COMPOUND(%=, l *= r)
SHIM_MAT_COMPOUND(/=, l /= r)
#undef SHIM_MAT_COMPOUND

template <class T, int VS> template <class X, enable_arma<X>> subview<T, VS>& subview<T, VS>::operator=(const X& x) {
  const auto b = as_mat(x); /* a copy when x is a view (possibly of the same parent) */
  Mat<T> c(b.n_rows, b.n_cols);
  for (uword i = 0; i < c.n_elem; ++i) c.mem[i] = T(b.mem[i]);
  if (c.n_rows != n_rows && c.n_elem == n_elem && (c.n_rows == 1 || c.n_cols == 1) && (n_rows == 1 || n_cols == 1)) { c.n_rows = n_rows; c.n_cols = n_cols; }
  apply(c, [](T& l, const T& r) { l = r; });
  return *this;
}
template <class T, int VS> subview<T, VS>& subview<T, VS>::operator=(const subview& x) { return this->template operator=<subview<T, VS>>(x); }
#define SHIM_SV_COMPOUND(OP, EXPR)                                                                       \
  template <class T, int VS> template <class X, enable_arma<X>> subview<T, VS>& subview<T, VS>::operator OP(const X& x) { \
    const Mat<T> c(as_mat(x));                                                                           \
    apply(c, [](T& l, const T& r) { EXPR; });                                                            \
    return *this;                                                                                        \
  }
SHIM_SV_COMPOUND(+=, l += r)
SHIM_SV_COMPOUND(-=, l -= r)
SHIM_SV_COMPOUND(%=, l *= r)
SHIM_SV_COMPOUND(/=, l /= r)
#undef SHIM_SV_COMPOUND
template <class T> template <class X, enable_arma<X>> diagview<T>& diagview<T>::operator=(const X& x) {
  const auto& b = as_mat(x);
  shim_check(b.n_elem == n_elem, "diagview: incompatible dimensions");
  for (uword i = 0; i < n_elem; ++i) at(i) = b.mem[i];
  return *this;
}

/* each_col() / each_row(): the vector is broadcast over the columns / rows of the parent */
template <class P, int DIM> struct each_proxy {
  typedef typename P::elem_type T;
  P& p;
  explicit each_proxy(P& p_) : p(p_) {}
  template <class X, class F> void run(const X& x, F f) {
    const Mat<T> v(as_mat(x));
    if (DIM == 0) shim_check(v.n_elem == p.n_rows && (v.n_cols == 1), "each_col(): incompatible size");
    else shim_check(v.n_elem == p.n_cols && (v.n_rows == 1), "each_row(): incompatible size");
    for (uword j = 0; j < p.n_cols; ++j) for (uword i = 0; i < p.n_rows; ++i) f(p.at(i, j), v.mem[DIM == 0 ? i : j]);
  }
  template <class X> void operator=(const X& x) { run(x, [](T& l, const T& r) { l = r; }); }
  template <class X> void operator+=(const X& x) { run(x, [](T& l, const T& r) { l += r; }); }
  template <class X> void operator-=(const X& x) { run(x, [](T& l, const T& r) { l -= r; }); }
  template <class X> void operator%=(const X& x) { run(x, [](T& l, const T& r) { l *= r; }); }
  template <class X> void operator/=(const X& x) { run(x, [](T& l, const T& r) { l /= r; }); }
};
/* X - M.each_row(): out(i,j) = X(j) - M(i,j)   (operator-(Base, subview_each1), used by selectterms) */
template <class X, class P, enable_arma<X> = 0> Mat<typename P::elem_type> operator-(const X& x, const each_proxy<P, 1>& e) {
  typedef typename P::elem_type T;
  const Mat<T> v(as_mat(x));
  shim_check(v.n_elem == e.p.n_cols, "each_row(): incompatible size");
  Mat<T> o(e.p.n_rows, e.p.n_cols);
  for (uword j = 0; j < o.n_cols; ++j) for (uword i = 0; i < o.n_rows; ++i) o.at(i, j) = v.mem[j] - e.p.at(i, j);
  return o;
}

/* ------------------------------------------------------------------ element-wise operators (eager) */
template <class A, class B> struct bin_result {
  typedef typename promote<typename A::elem_type, typename B::elem_type>::type T;
  static constexpr int VS = A::vec_state == 1 || B::vec_state == 1 ? 1 : (A::vec_state == 2 || B::vec_state == 2 ? 2 : 0);
  typedef typename result_of<T, VS>::type type;
};
#define SHIM_BINARY(OP, EXPR)                                                                                        \
  template <class A, class B, enable_arma<A> = 0, enable_arma<B> = 0> typename bin_result<A, B>::type operator OP(const A& a_, const B& b_) { \
    typedef typename bin_result<A, B>::T T;                                                                          \
    const auto& a = as_mat(a_); const auto& b = as_mat(b_);                                                          \
    shim_check(a.n_rows == b.n_rows && a.n_cols == b.n_cols, "element-wise operation: incompatible matrix dimensions"); \
    typename bin_result<A, B>::type o;                                                                               \
    static_cast<Mat<T>&>(o).set_size(a.n_rows, a.n_cols);                                                            \
    for (uword i = 0; i < o.n_elem; ++i) { const T l = T(a.mem[i]), r = T(b.mem[i]); o.mem[i] = EXPR; }              \
    return o;                                                                                                        \
  }                                                                                                                  \
  template <class A, class S, enable_arma<A> = 0, enable_scalar<S> = 0> typename result_of<typename promote<typename A::elem_type, S>::type, A::vec_state>::type operator OP(const A& a_, S s) { \
    typedef typename promote<typename A::elem_type, S>::type T;                                                      \
    const auto& a = as_mat(a_);                                                                                      \
    typename result_of<T, A::vec_state>::type o;                                                                     \
    static_cast<Mat<T>&>(o).set_size(a.n_rows, a.n_cols);                                                            \
    const T r = T(s);                                                                                                \
    for (uword i = 0; i < o.n_elem; ++i) { const T l = T(a.mem[i]); o.mem[i] = EXPR; }                               \
    return o;                                                                                                        \
  }                                                                                                                  \
  template <class S, class B, enable_scalar<S> = 0, enable_arma<B> = 0> typename result_of<typename promote<typename B::elem_type, S>::type, B::vec_state>::type operator OP(S s, const B& b_) { \
    typedef typename promote<typename B::elem_type, S>::type T;                                                      \
    const auto& b = as_mat(b_);                                                                                      \
    typename result_of<T, B::vec_state>::type o;                                                                     \
    static_cast<Mat<T>&>(o).set_size(b.n_rows, b.n_cols);                                                            \
    const T l = T(s);                                                                                                \
    for (uword i = 0; i < o.n_elem; ++i) { const T r = T(b.mem[i]); o.mem[i] = EXPR; }                               \
    return o;                                                                                                        \
  }
SHIM_BINARY(+, l + r)
SHIM_BINARY(-, l - r)
SHIM_BINARY(/, l / r)
#undef SHIM_BINARY
/* % : element-wise product (matrix operands only) */
template <class A, class B, enable_arma<A> = 0, enable_arma<B> = 0> typename bin_result<A, B>::type operator%(const A& a_, const B& b_) {
  typedef typename bin_result<A, B>::T T;
  const auto& a = as_mat(a_); const auto& b = as_mat(b_);
  shim_check(a.n_rows == b.n_rows && a.n_cols == b.n_cols, "element-wise multiplication: incompatible matrix dimensions");
  typename bin_result<A, B>::type o;
  static_cast<Mat<T>&>(o).set_size(a.n_rows, a.n_cols);
  for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = T(a.mem[i]) * T(b.mem[i]);
  return o;
}
/* scalar * X, X * scalar */
template <class A, class S, enable_arma<A> = 0, enable_scalar<S> = 0> typename result_of<typename promote<typename A::elem_type, S>::type, A::vec_state>::type operator*(const A& a_, S s) {
  typedef typename promote<typename A::elem_type, S>::type T;
  const auto& a = as_mat(a_);
  typename result_of<T, A::vec_state>::type o;
  static_cast<Mat<T>&>(o).set_size(a.n_rows, a.n_cols);
  for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = T(a.mem[i]) * T(s);
  return o;
}
template <class S, class A, enable_scalar<S> = 0, enable_arma<A> = 0> typename result_of<typename promote<typename A::elem_type, S>::type, A::vec_state>::type operator*(S s, const A& a_) { return a_ * s; }
template <class A, enable_arma<A> = 0> typename result_of<typename A::elem_type, A::vec_state>::type operator-(const A& a_) {
  typedef typename A::elem_type T;
  const auto& a = as_mat(a_);
  typename result_of<T, A::vec_state>::type o;
  static_cast<Mat<T>&>(o).set_size(a.n_rows, a.n_cols);
  for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = -a.mem[i];
  return o;
}
/* relational operators against a scalar -> uword matrices of the same shape */
#define SHIM_REL(OP)                                                                                                  \
  template <class A, class S, enable_arma<A> = 0, enable_scalar<S> = 0> typename result_of<uword, A::vec_state>::type operator OP(const A& a_, S s) { \
    const auto& a = as_mat(a_);                                                                                       \
    typename result_of<uword, A::vec_state>::type o;                                                                  \
    static_cast<Mat<uword>&>(o).set_size(a.n_rows, a.n_cols);                                                         \
    typedef typename promote<typename A::elem_type, S>::type T;                                                       \
    for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = (T(a.mem[i]) OP T(s)) ? 1u : 0u;                                  \
    return o;                                                                                                         \
  }
SHIM_REL(>)
SHIM_REL(<)
SHIM_REL(==)
SHIM_REL(>=)
SHIM_REL(<=)
#undef SHIM_REL

/* element-wise functions */
#define SHIM_UNARY(NAME, EXPR)                                                                            \
  template <class A, enable_arma<A> = 0> typename result_of<typename A::elem_type, A::vec_state>::type NAME(const A& a_) { \
    typedef typename A::elem_type T;                                                                      \
    const auto& a = as_mat(a_);                                                                           \
    typename result_of<T, A::vec_state>::type o;                                                          \
    static_cast<Mat<T>&>(o).set_size(a.n_rows, a.n_cols);                                                 \
    for (uword i = 0; i < o.n_elem; ++i) { const T v = a.mem[i]; o.mem[i] = EXPR; }                       \
    return o;                                                                                             \
  }
SHIM_UNARY(exp, std::exp(v))
SHIM_UNARY(log, std::log(v))
SHIM_UNARY(sqrt, std::sqrt(v))
SHIM_UNARY(sin, std::sin(v))
SHIM_UNARY(cos, std::cos(v))
SHIM_UNARY(square, v * v)
SHIM_UNARY(abs, (v < T(0) ? T(-v) : v))
SHIM_UNARY(sign, (v > T(0) ? T(1) : (v < T(0) ? T(-1) : T(0))))
#undef SHIM_UNARY
using std::exp; using std::log; using std::sqrt; using std::sin; using std::cos; using std::abs; using std::pow;
template <class A, class S, enable_arma<A> = 0, enable_scalar<S> = 0> typename result_of<typename A::elem_type, A::vec_state>::type pow(const A& a_, S p) {
  typedef typename A::elem_type T;
  const auto& a = as_mat(a_);
  typename result_of<T, A::vec_state>::type o;
  static_cast<Mat<T>&>(o).set_size(a.n_rows, a.n_cols);
  for (uword i = 0; i < o.n_elem; ++i) o.mem[i] = std::pow(a.mem[i], T(p));
  return o;
}

/* ------------------------------------------------------------------ reductions */
template <class A, enable_arma<A> = 0> typename A::elem_type accu(const A& a_) { const auto& a = as_mat(a_); return accumulate2(a.mem, a.n_elem); }
/* sum(vector) -> scalar; sum(matrix) -> row of column sums; sum(X, dim) -> matrix */
template <class A, typename std::enable_if<is_arma<A>::value && A::vec_state != 0, int>::type = 0> typename A::elem_type sum(const A& a_) { return accu(a_); }
template <class A, enable_arma<A> = 0> Mat<typename A::elem_type> sum(const A& a_, uword dim) {
  typedef typename A::elem_type T;
  const auto& a = as_mat(a_);
  if (dim == 0) { /* op_sum::apply_noalias_unwrap: per column arrayops::accumulate */
    Mat<T> o(1, a.n_cols);
    for (uword j = 0; j < a.n_cols; ++j) o.mem[j] = accumulate2(a.colptr(j), a.n_rows);
    return o;
  }
  Mat<T> o(a.n_rows, 1); /* dim 1: out = col 0; out += col j  (arrayops::inplace_plus) */
  if (a.n_cols) for (uword i = 0; i < a.n_rows; ++i) o.mem[i] = a.mem[i];
  for (uword j = 1; j < a.n_cols; ++j) for (uword i = 0; i < a.n_rows; ++i) o.mem[i] += a.at(i, j);
  return o;
}
template <class A, typename std::enable_if<is_arma<A>::value && A::vec_state == 0, int>::type = 0> Row<typename A::elem_type> sum(const A& a_) { return Row<typename A::elem_type>(sum(a_, 0)); }
template <class A, enable_arma<A> = 0> double mean(const A& a_) { /* op_mean::direct_mean */
  const auto& a = as_mat(a_);
  const double r = double(accumulate2(a.mem, a.n_elem)) / double(a.n_elem);
  if (std::isfinite(r)) return r;
  double rm = 0; /* direct_mean_robust */
  for (uword i = 0; i < a.n_elem; ++i) rm = rm + (double(a.mem[i]) - rm) / double(i + 1);
  return rm;
}
template <class A, enable_arma<A> = 0> double var(const A& a_) { /* op_var::direct_var, norm_type 0 */
  const auto& a = as_mat(a_);
  const uword n = a.n_elem;
  if (n < 2) return 0.0;
  const double acc1 = mean(a);
  double acc2 = 0, acc3 = 0;
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { const double ti = acc1 - a.mem[i], tj = acc1 - a.mem[j]; acc2 += ti * ti + tj * tj; acc3 += ti + tj; }
  if (i < n) { const double ti = acc1 - a.mem[i]; acc2 += ti * ti; acc3 += ti; }
  const double v = (acc2 - acc3 * acc3 / double(n)) / double(n - 1);
  return v; /* (the robust fallback only triggers for non-finite values) */
}
template <class A, enable_arma<A> = 0> typename A::elem_type max(const A& a_) { return as_mat(a_).max(); }
template <class A, enable_arma<A> = 0> typename A::elem_type min(const A& a_) { return as_mat(a_).min(); }
template <class A, class B, enable_arma<A> = 0, enable_arma<B> = 0> double dot(const A& a_, const B& b_) {
  const auto& a = as_mat(a_); const auto& b = as_mat(b_);
  shim_check(a.n_elem == b.n_elem, "dot(): objects must have the same number of elements");
  return dot_direct(a.n_elem, a.mem, b.mem);
}

/* ------------------------------------------------------------------ matrix product */
template <class T> Mat<T> matmul(const Mat<T>& A, const Mat<T>& B) {
  shim_check(A.n_cols == B.n_rows, "matrix multiplication: incompatible matrix dimensions");
  Mat<T> C(A.n_rows, B.n_cols);
  const uword M = A.n_rows, N = B.n_cols, Kd = A.n_cols;
#if ARMA_SHIM_BLAS
  /* netlib dgemv('N') / dgemm('N','N'): C(:,j) += B(l,j) * A(:,l), l ascending (skipping exact zeros of B, as the
   * reference BLAS does: IF (B(L,J).NE.ZERO)); every C(i,j) is a left-to-right sum over l */
  for (uword j = 0; j < N; ++j)
    for (uword l = 0; l < Kd; ++l) {
      const T t = B.mem[l + j * Kd];
      if (t == T(0)) continue;
      T* c = C.mem + j * M;
      const T* a = A.mem + l * M;
      for (uword i = 0; i < M; ++i) c[i] += t * a[i];
    }
#else
  /* gemm_emul_large / gemv_emul: row of A against column of B with the two-accumulator dot */
  std::vector<T> tmp(Kd);
  for (uword i = 0; i < M; ++i) {
    for (uword l = 0; l < Kd; ++l) tmp[l] = A.mem[i + l * M];
    for (uword j = 0; j < N; ++j) C.mem[i + j * M] = dot_arma(Kd, tmp.data(), B.mem + j * Kd);
  }
#endif
  return C;
}
/* A^T * B without forming A^T: dgemv('T') / dgemm('T','N') -- one left-to-right dot per output */
template <class T> Mat<T> matmul_tn(const Mat<T>& A, const Mat<T>& B) {
  shim_check(A.n_rows == B.n_rows, "matrix multiplication: incompatible matrix dimensions");
  Mat<T> C(A.n_cols, B.n_cols);
  for (uword j = 0; j < B.n_cols; ++j)
    for (uword i = 0; i < A.n_cols; ++i) C.mem[i + j * A.n_cols] = dot_blas(A.n_rows, A.mem + i * A.n_rows, B.mem + j * B.n_rows);
  return C;
}
/* lazy transpose marker so that X.t() * Y keeps BLAS's transposed-operand order */
template <class A, class B, enable_arma<A> = 0, enable_arma<B> = 0>
typename result_of<typename promote<typename A::elem_type, typename B::elem_type>::type, (B::vec_state == 1 ? 1 : (A::vec_state == 2 ? 2 : 0))>::type
operator*(const A& a_, const B& b_) {
  typedef typename promote<typename A::elem_type, typename B::elem_type>::type T;
  static_assert(std::is_same<T, typename A::elem_type>::value && std::is_same<T, typename B::elem_type>::value, "matrix product of one element type");
  const auto& a = as_mat(a_); const auto& b = as_mat(b_);
  typename result_of<T, (B::vec_state == 1 ? 1 : (A::vec_state == 2 ? 2 : 0))>::type o;
  if (a.n_rows == 1 && a.n_cols == b.n_rows && b.n_cols != 1) {
    /* row vector times matrix: Armadillo calls gemv<true>(B, a) -- one dot per column of B */
    Mat<T> at(a.mem ? const_cast<T*>(a.mem) : nullptr, a.n_cols, 1, false, true);
    static_cast<Mat<T>&>(o) = matmul_tn(b, at).t();
  } else static_cast<Mat<T>&>(o) = matmul(a, b);
  return o;
}
template <class T> template <class X, enable_arma<X>> Mat<T>& Mat<T>::operator*=(const X& x) { Mat<T> r = matmul(*this, Mat<T>(as_mat(x))); return (*this = std::move(r)); }

/* ------------------------------------------------------------------ generators / rearrangements */
inline vec linspace(double start, double end, uword N) { /* Armadillo: start + i*delta, last element = end */
  vec o(N);
  if (N == 1) { o.mem[0] = end; return o; }
  const double delta = (end - start) / double(N - 1);
  for (uword i = 0; i + 1 < N; ++i) o.mem[i] = start + double(i) * delta;
  if (N) o.mem[N - 1] = end;
  return o;
}
template <class A, enable_arma<A> = 0> typename result_of<typename A::elem_type, A::vec_state>::type reverse(const A& a_) {
  const auto& a = as_mat(a_);
  typename result_of<typename A::elem_type, A::vec_state>::type o(a);
  std::reverse(o.mem, o.mem + o.n_elem);
  return o;
}
template <class T> Mat<T> fliplr(const Mat<T>& a) {
  Mat<T> o(a.n_rows, a.n_cols);
  for (uword j = 0; j < a.n_cols; ++j) if (a.n_rows) std::memcpy(o.colptr(j), a.colptr(a.n_cols - 1 - j), a.n_rows * sizeof(T));
  return o;
}
template <class A, enable_arma<A> = 0> typename result_of<typename A::elem_type, A::vec_state>::type diff(const A& a_) {
  typedef typename A::elem_type T;
  const auto& a = as_mat(a_);
  typename result_of<T, A::vec_state>::type o;
  const uword n = a.n_elem ? a.n_elem - 1 : 0;
  static_cast<Mat<T>&>(o).set_size(A::vec_state == 2 ? 1 : n, A::vec_state == 2 ? n : 1);
  for (uword i = 0; i < n; ++i) o.mem[i] = a.mem[i + 1] - a.mem[i];
  return o;
}
template <class T> Col<T> diagvec(const Mat<T>& a) { return diagview<T>(const_cast<Mat<T>&>(a)).eval(); }
template <class A, enable_arma<A> = 0> uvec find(const A& a_, uword k = 0) {
  const auto& a = as_mat(a_);
  std::vector<uword> idx;
  for (uword i = 0; i < a.n_elem && (k == 0 || idx.size() < k); ++i) if (a.mem[i] != 0) idx.push_back(i);
  uvec o(idx.size());
  std::copy(idx.begin(), idx.end(), o.mem);
  return o;
}
template <class T> template <class I> Col<T> Mat<T>::elem(const I& idx_) const {
  const auto& idx = as_mat(idx_);
  Col<T> o(idx.n_elem);
  for (uword i = 0; i < idx.n_elem; ++i) { const uword k = uword(idx.mem[i]); shim_check(k < n_elem, "Mat::elem(): index out of bounds"); o.mem[i] = mem[k]; }
  return o;
}
template <class O> struct conv_to {
  template <class A, enable_arma<A> = 0> static O from(const A& a_) {
    const auto& a = as_mat(a_);
    O o;
    if (O::vec_state == 0) static_cast<Mat<typename O::elem_type>&>(o).set_size(a.n_rows, a.n_cols);
    else o.set_size(a.n_elem);
    for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = typename O::elem_type(a.mem[i]);
    return o;
  }
};

/* third-party substitutions (see the header comment) */
inline std::function<void(uword*, uword)>& arma_shim_shuffle() { static std::function<void(uword*, uword)> f; return f; }
template <class T> Col<T> shuffle(const Col<T>& a) {
  Col<T> o(a);
  if (arma_shim_shuffle()) {
    std::vector<uword> perm(a.n_elem);
    for (uword i = 0; i < a.n_elem; ++i) perm[i] = i;
    arma_shim_shuffle()(perm.data(), a.n_elem);
    for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[perm[i]];
  }
  return o;
}
/* symmetric eigenproblem: ascending eigenvalues, eigenvectors in columns -- provided by the including program */
}  // namespace arma
extern "C" void arma_shim_eig_sym(std::uint64_t n, const double* A /* n x n col-major */, double* w /* n */, double* V /* n x n */);
namespace arma {
inline bool eig_sym(vec& w, mat& V, const mat& A) {
  shim_check(A.n_rows == A.n_cols, "eig_sym(): given matrix must be square sized");
  w.set_size(A.n_rows);
  V.set_size(A.n_rows, A.n_rows);
  arma_shim_eig_sym(A.n_rows, A.mem, w.mem, V.mem);
  return true;
}
/* solve / inv: Gaussian elimination with partial pivoting (dgesv upstream) */
inline mat solve(const mat& A, const mat& B) {
  shim_check(A.n_rows == A.n_cols && A.n_rows == B.n_rows, "solve(): incompatible dimensions");
  const uword n = A.n_rows, m = B.n_cols;
  mat L(A), X(B);
  for (uword k = 0; k < n; ++k) {
    uword p = k;
    for (uword i = k + 1; i < n; ++i) if (std::abs(L.at(i, k)) > std::abs(L.at(p, k))) p = i;
    if (L.at(p, k) == 0.0) throw std::runtime_error("solve(): solution not found");
    if (p != k) { for (uword j = 0; j < n; ++j) std::swap(L.at(k, j), L.at(p, j)); for (uword j = 0; j < m; ++j) std::swap(X.at(k, j), X.at(p, j)); }
    for (uword i = k + 1; i < n; ++i) {
      const double f = L.at(i, k) / L.at(k, k);
      if (f == 0.0) continue;
      for (uword j = k; j < n; ++j) L.at(i, j) -= f * L.at(k, j);
      for (uword j = 0; j < m; ++j) X.at(i, j) -= f * X.at(k, j);
    }
  }
  for (uword j = 0; j < m; ++j)
    for (uword ii = n; ii-- > 0;) {
      double s = X.at(ii, j);
      for (uword c = ii + 1; c < n; ++c) s -= L.at(ii, c) * X.at(c, j);
      X.at(ii, j) = s / L.at(ii, ii);
    }
  return X;
}
inline vec solve(const mat& A, const vec& b) { return vec(solve(A, static_cast<const mat&>(b))); }
inline mat inv(const mat& A) { mat I(A.n_rows, A.n_rows); for (uword i = 0; i < A.n_rows; ++i) I.at(i, i) = 1.0; return solve(A, I); }

/* ------------------------------------------------------------------ Cube */
template <class T> struct cube_each_slice;
template <class T> struct Cube : arma_tag {
  typedef T elem_type;
  uword n_rows = 0, n_cols = 0, n_slices = 0, n_elem = 0;
  std::vector<T> store;
  mutable std::vector<Mat<T>> views; /* slice(l): a matrix aliasing the cube's memory, like Armadillo's mat_ptrs */
  Cube() {}
  Cube(uword r, uword c, uword s) { set_size(r, c, s); }
  Cube(const Cube& o) : n_rows(o.n_rows), n_cols(o.n_cols), n_slices(o.n_slices), n_elem(o.n_elem), store(o.store) { rebuild(); }
  Cube(Cube&& o) noexcept : n_rows(o.n_rows), n_cols(o.n_cols), n_slices(o.n_slices), n_elem(o.n_elem), store(std::move(o.store)) { o.n_rows = o.n_cols = o.n_slices = o.n_elem = 0; o.views.clear(); rebuild(); }
  Cube& operator=(const Cube& o) { if (this != &o) { n_rows = o.n_rows; n_cols = o.n_cols; n_slices = o.n_slices; n_elem = o.n_elem; store = o.store; rebuild(); } return *this; }
  Cube& operator=(Cube&& o) { if (this != &o) { n_rows = o.n_rows; n_cols = o.n_cols; n_slices = o.n_slices; n_elem = o.n_elem; store = std::move(o.store); o.n_rows = o.n_cols = o.n_slices = o.n_elem = 0; o.views.clear(); rebuild(); } return *this; }
  void rebuild() {
    views.clear();
    views.reserve(n_slices);
    for (uword s = 0; s < n_slices; ++s) views.emplace_back(store.data() + s * n_rows * n_cols, n_rows, n_cols, false, true);
  }
  void set_size(uword r, uword c, uword s) {
    if (r == n_rows && c == n_cols && s == n_slices) return;
    n_rows = r; n_cols = c; n_slices = s; n_elem = r * c * s;
    store.assign(n_elem, T(0));
    rebuild();
  }
  Cube& zeros() { std::fill(store.begin(), store.end(), T(0)); return *this; }
  Mat<T>& slice(uword s) { shim_check(s < n_slices, "Cube::slice(): index out of bounds"); return views[s]; }
  const Mat<T>& slice(uword s) const { shim_check(s < n_slices, "Cube::slice(): index out of bounds"); return views[s]; }
  T* memptr() { return store.data(); }
  const T* memptr() const { return store.data(); }
  Cube& operator+=(const Cube& o) { shim_check(o.n_elem == n_elem, "Cube: incompatible dimensions"); for (uword i = 0; i < n_elem; ++i) store[i] += o.store[i]; return *this; }
  struct slices_view {
    Cube& c; uword a, b;
    void operator=(const Cube& o) {
      shim_check(o.n_rows == c.n_rows && o.n_cols == c.n_cols && o.n_slices == b + 1 - a, "Cube::slices(): incompatible dimensions");
      std::copy(o.store.begin(), o.store.end(), c.store.begin() + a * c.n_rows * c.n_cols);
    }
  };
  slices_view slices(uword a, uword b) { shim_check(a <= b + 1 && b < n_slices, "Cube::slices(): indices out of bounds"); return slices_view{*this, a, b}; }
  struct rows_view { /* subview_cube = matrix: Armadillo accepts it at compile time and checks sizes at run time */
    Cube& c; uword a, b;
    void operator=(const Mat<T>& m) {
      shim_check(c.n_slices == 1 && m.n_rows == b + 1 - a && m.n_cols == c.n_cols, "copy into subcube: incompatible dimensions");
      for (uword j = 0; j < m.n_cols; ++j) for (uword i = 0; i < m.n_rows; ++i) c.views[0].at(a + i, j) = m.at(i, j);
    }
  };
  rows_view rows(uword a, uword b) { return rows_view{*this, a, b}; }
  cube_each_slice<T> each_slice();
};
template <class T> struct cube_each_slice {
  Cube<T>& c;
  template <class X> void operator%=(const X& x) { const Mat<T> m(as_mat(x)); for (uword s = 0; s < c.n_slices; ++s) c.slice(s) %= m; }
};
template <class T> cube_each_slice<T> Cube<T>::each_slice() { return cube_each_slice<T>{*this}; }
template <class T, class S, enable_scalar<S> = 0> Cube<T> operator*(S s, const Cube<T>& c) { Cube<T> o(c); for (uword i = 0; i < o.n_elem; ++i) o.store[i] = T(s) * o.store[i]; return o; }
template <class T> Cube<T> sum(const Cube<T>& c, uword dim) {
  shim_check(dim == 0, "sum(cube, dim): only dim 0 is implemented");
  Cube<T> o(1, c.n_cols, c.n_slices);
  for (uword s = 0; s < c.n_slices; ++s) for (uword j = 0; j < c.n_cols; ++j) o.store[j + s * c.n_cols] = accumulate2(c.slice(s).colptr(j), c.n_rows);
  return o;
}
template <class T> Mat<T>::Mat(const Cube<T>& c) { /* Mat(const BaseCube&): a cube with one singleton dimension */
  if (c.n_slices == 1) { init(c.n_rows, c.n_cols); }
  else if (c.n_rows == 1) { init(c.n_cols, c.n_slices); }
  else if (c.n_cols == 1) { init(c.n_rows, c.n_slices); }
  else throw std::logic_error("Mat(cube): cube is not interpretable as a matrix");
  if (n_elem) std::memcpy(mem, c.store.data(), n_elem * sizeof(T));
}

}  // namespace arma
#endif
