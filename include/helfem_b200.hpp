// Header-only C++ shims over the C ABI of libhelfemqc_b200 (include/helfem_b200.h) with the member
// names and argument meaning of the reference's classes, for callers written against HelFEM:
//
//   helfem::atomic::basis::TwoDBasisT<double>      src/atomic/TwoDBasis.h:181,223-227
//   helfem::diatomic::basis::TwoDBasis             src/diatomic/basis.h:265,300-302,344
//   helfem::sadatom::basis::TwoDBasis              src/sadatom/basis.h:72,110-114
//   helfem::{atomic,diatomic,sadatom}::dftgrid::DFTGrid::eval_Fxc   src/atomic/dftgrid.h:156,159 etc.
//
// Every class is a template on the dense matrix type `Mat`, which only needs
//   Mat(rows, cols),  double* data(),  rows(),  cols(),  column-major storage
// -- Eigen::MatrixXd (= helfem::Matrix) and arma::mat both qualify, so a HelFEM driver switches by
// changing one typedef.  Errors are thrown as the reference throws them: std::logic_error for misuse
// (integrals not computed, wrong matrix size), std::runtime_error for everything else.
// No computation happens in this header; there is no CPU fallback behind it.
#ifndef HELFEM_B200_HPP
#define HELFEM_B200_HPP

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "helfem_b200.h"

namespace helfem_b200 {

inline void check(long long rc) {
  if (rc >= 0) return;
  const std::string msg = hfq_last_error();
  if (rc == HFQ_ERR_INVALID || rc == HFQ_ERR_STATE) throw std::logic_error(msg);
  throw std::runtime_error(msg);
}

namespace detail {

// owns one hfq_tables (host caches) and, lazily, one hfq_ctx (the same basis bound to a GPU)
class Handle {
 public:
  Handle() = default;
  Handle(const Handle &) = delete;
  Handle &operator=(const Handle &) = delete;
  Handle(Handle &&o) noexcept : tab_(o.tab_), ctx_(o.ctx_), device_(o.device_) { o.tab_ = nullptr, o.ctx_ = nullptr; }
  Handle &operator=(Handle &&o) noexcept {
    if (this != &o) {
      reset();
      tab_ = o.tab_, ctx_ = o.ctx_, device_ = o.device_;
      o.tab_ = nullptr, o.ctx_ = nullptr;
    }
    return *this;
  }
  ~Handle() { reset(); }
  void reset() {
    if (ctx_) hfq_destroy(ctx_);
    if (tab_) hfq_tables_destroy(tab_);
    ctx_ = nullptr, tab_ = nullptr;
  }
  void adopt(hfq_tables *t) {
    reset();
    tab_ = t;
  }
  void set_device(int d) { device_ = d; }
  bool has_tables() const { return tab_ != nullptr; }
  const hfq_tables *tables() const {
    if (!tab_) throw std::logic_error("Primitive teis have not been computed!\n");   // TwoDBasis.cpp:775
    return tab_;
  }
  hfq_ctx *ctx() {
    if (!ctx_) check(hfq_create(&ctx_, tables(), device_));
    return ctx_;
  }
  hfq_tables_info info() const {
    hfq_tables_info i;
    check(hfq_tables_get_info(tables(), &i));
    return i;
  }

 private:
  hfq_tables *tab_ = nullptr;
  hfq_ctx *ctx_ = nullptr;
  int device_ = 0;
};

template <class Mat>
void expect_square(const Mat &P, long long n) {
  if ((long long)P.rows() != n || (long long)P.cols() != n)
    throw std::logic_error("Matrix does not have expected size! Got " + std::to_string(P.rows()) + " x " +
                           std::to_string(P.cols()) + ", expected " + std::to_string(n) + " x " + std::to_string(n) +
                           "!\n");   // src/diatomic/basis.cpp:2093-2097
}

template <class Mat>
Mat coulomb(Handle &h, const Mat &P) {
  const long long n = h.info().Nbf;
  expect_square(P, n);
  Mat J(P.rows(), P.cols());
  check(hfq_coulomb(h.ctx(), P.data(), n, J.data(), n));
  return J;
}

template <class Mat>
Mat exchange(Handle &h, const Mat &P) {
  const long long n = h.info().Nbf;
  expect_square(P, n);
  Mat K(P.rows(), P.cols());
  check(hfq_exchange(h.ctx(), P.data(), n, K.data(), n));
  return K;
}

}  // namespace detail

// One Fock-build step of the drivers' fock_builder lambdas in a single call (src/diatomic/main.cpp:413-426):
// J = coulomb(P), K = exchange(kscale * P); one upload of P, J travels back while K is built.
template <class Mat>
void coulomb_exchange(detail::Handle &h, const Mat &P, double kscale, Mat &J, Mat &K) {
  const long long n = h.info().Nbf;
  detail::expect_square(P, n);
  J = Mat(P.rows(), P.cols());
  K = Mat(P.rows(), P.cols());
  check(hfq_coulomb_exchange(h.ctx(), P.data(), n, kscale, J.data(), n, K.data(), n));
}

namespace atomic {

// helfem::atomic::basis::TwoDBasisT<double>
template <class Mat>
class TwoDBasis {
 public:
  // arguments as parsed by src/atomic/main.cpp (primbas = 4 LIP, nnodes per element)
  TwoDBasis(int Z, int lmax, int mmax, int nelem, int nnodes = 15, double Rmax = 40.0, int igrid = 4, double zexp = 2.0,
            int nquad = 0, int device = 0)
      : Z_(Z), lmax_(lmax), mmax_(mmax), nelem_(nelem), nnodes_(nnodes), igrid_(igrid), nquad_(nquad), Rmax_(Rmax), zexp_(zexp) {
    bare_.set_device(device);
    rs_.set_device(device);
  }
  void compute_tei(bool /*exchange*/ = true) {
    hfq_tables *t = nullptr;
    check(hfq_tables_atomic(&t, Z_, lmax_, mmax_, nelem_, nnodes_, Rmax_, igrid_, zexp_, nquad_));
    bare_.adopt(t);
  }
  void compute_yukawa(double lambda) {
    hfq_tables *t = nullptr;
    check(hfq_tables_atomic_yukawa(&t, Z_, lmax_, mmax_, nelem_, nnodes_, Rmax_, igrid_, zexp_, nquad_, lambda));
    rs_.adopt(t);
  }
  void compute_erfc(double mu) {
    hfq_tables *t = nullptr;
    check(hfq_tables_atomic_erfc(&t, Z_, lmax_, mmax_, nelem_, nnodes_, Rmax_, igrid_, zexp_, nquad_, mu));
    rs_.adopt(t);
  }
  size_t Nbf() const { return (size_t)bare_.info().Nbf; }
  Mat coulomb(const Mat &P) { return detail::coulomb(bare_, P); }
  Mat exchange(const Mat &P) { return detail::exchange(bare_, P); }          // returns -K, like the reference
  Mat rs_exchange(const Mat &P) { return detail::exchange(rs_, P); }
  void coulomb_exchange(const Mat &P, double kscale, Mat &J, Mat &K) { helfem_b200::coulomb_exchange(bare_, P, kscale, J, K); }
  // overlap / kinetic / nuclear attraction (setup helpers)
  void one_electron(Mat &S, Mat &T, Mat &V) const {
    const long long n = bare_.info().Nbf;
    S = Mat(n, n), T = Mat(n, n), V = Mat(n, n);
    check(hfq_tables_one_electron(bare_.tables(), S.data(), T.data(), V.data()));
  }
  detail::Handle &handle() { return bare_; }

 private:
  int Z_, lmax_, mmax_, nelem_, nnodes_, igrid_, nquad_;
  double Rmax_, zexp_;
  detail::Handle bare_, rs_;
};

}  // namespace atomic

namespace diatomic {

// helfem::diatomic::basis::TwoDBasis
template <class Mat>
class TwoDBasis {
 public:
  // lmmax[|m|] = largest l of the m shell (the --lmax list of src/diatomic/main.cpp)
  TwoDBasis(int Z1, int Z2, double Rbond, std::vector<int> lmmax, int nelem, int nnodes = 15, double Rmax = 40.0,
            int igrid = 4, double zexp = 1.0, int nquad = 0, int device = 0)
      : Z1_(Z1), Z2_(Z2), nelem_(nelem), nnodes_(nnodes), igrid_(igrid), nquad_(nquad), Rbond_(Rbond), Rmax_(Rmax),
        zexp_(zexp), lmmax_(std::move(lmmax)) {
    h_.set_device(device);
  }
  void compute_tei(bool /*exchange*/ = true) {
    hfq_tables *t = nullptr;
    check(hfq_tables_diatomic(&t, Z1_, Z2_, Rbond_, lmmax_.data(), (int)lmmax_.size(), nelem_, nnodes_, Rmax_, igrid_, zexp_,
                              nquad_));
    h_.adopt(t);
    if (absm_) check(hfq_set_absm_symmetric(h_.ctx(), 1));
  }
  void set_absm_symmetric(bool s) {   // src/diatomic/basis.cpp:913-915
    absm_ = s;
    if (h_.has_tables()) check(hfq_set_absm_symmetric(h_.ctx(), s ? 1 : 0));
  }
  size_t Nbf() const { return (size_t)h_.info().Nbf; }
  Mat coulomb(const Mat &P) { return detail::coulomb(h_, P); }
  Mat exchange(const Mat &P) { return detail::exchange(h_, P); }
  void coulomb_exchange(const Mat &P, double kscale, Mat &J, Mat &K) { helfem_b200::coulomb_exchange(h_, P, kscale, J, K); }
  void one_electron(Mat &S, Mat &T, Mat &V) const {
    const long long n = h_.info().Nbf;
    S = Mat(n, n), T = Mat(n, n), V = Mat(n, n);
    check(hfq_tables_one_electron(h_.tables(), S.data(), T.data(), V.data()));
  }
  detail::Handle &handle() { return h_; }

 private:
  int Z1_, Z2_, nelem_, nnodes_, igrid_, nquad_;
  double Rbond_, Rmax_, zexp_;
  std::vector<int> lmmax_;
  bool absm_ = false;
  detail::Handle h_;
};

}  // namespace diatomic

// DFTGrid::eval_Fxc of the atomic / diatomic drivers for the functionals built into the library and evaluated on
// the device, selected by their libxc ids like the reference's call: x_func = 1 Slater, 101 PBE, 202 TPSS exchange;
// c_func = 7 VWN5, 130 PBE, 231 TPSS correlation (correlation: restricted densities); <= 0 none (the HF drivers still
// call it to integrate Nel).  For any other functional use hfq_grid_density / hfq_grid_fxc around the caller's own
// libxc calls.  The x_pars / c_pars vectors of the reference's signature are dropped:
// none of the built-in functionals has external parameters.
template <class Mat>
class DFTGrid {
 public:
  // atomic basis: DFTGrid(&basis, ldft, mdft) (src/atomic/dftgrid.h:139); diatomic: mang <= 1 selects the
  // pure-m grid of src/diatomic/dftgrid_purem.h, mang >= 2 the 3D grid of src/diatomic/dftgrid.h
  DFTGrid(detail::Handle &basis, int lang, int mang) : h_(basis) { check(hfq_grid_attach(h_.ctx(), lang, mang)); }
  void eval_Fxc(int x_func, int c_func, const Mat &P, Mat &H, double &Exc, double &Nel, double &Ekin, double thr) {
    const long long n = h_.info().Nbf;
    detail::expect_square(P, n);
    H = Mat(n, n);
    check(hfq_eval_fxc(h_.ctx(), x_func, c_func, P.data(), n, nullptr, n, H.data(), n, nullptr, n, &Exc, &Nel, &Ekin, 1, thr));
  }
  void eval_Fxc(int x_func, int c_func, const Mat &Pa, const Mat &Pb, Mat &Ha, Mat &Hb, double &Exc, double &Nel,
                double &Ekin, bool beta, double thr) {
    const long long n = h_.info().Nbf;
    detail::expect_square(Pa, n);
    detail::expect_square(Pb, n);
    Ha = Mat(n, n);
    Hb = Mat(n, n);
    check(hfq_eval_fxc(h_.ctx(), x_func, c_func, Pa.data(), n, Pb.data(), n, Ha.data(), n, Hb.data(), n, &Exc, &Nel, &Ekin,
                       beta ? 1 : 0, thr));
  }

 private:
  detail::Handle &h_;
};

}  // namespace helfem_b200
#endif  // HELFEM_B200_HPP
