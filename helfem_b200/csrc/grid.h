// DFT quadrature grid of the atomic basis (r x theta x phi): host-side tables and the
// device engine behind eval_Fxc.  The reference materialises (functions x points) tables per
// element and runs ~10 GEMMs of size nbf_el^2 x npts on them every build
// (src/atomic/dftgrid.cpp:470-576, :51-242, :304-465).  Every basis function is separable,
// chi_(a,r)(ia,ir) = conj(Y_a(ia)) R_r(ir), so here all contractions factor into a radial
// and an angular stage on small pair tables -- O(100x) fewer flops and no per-build table
// construction.
#pragma once
#include <cuda_runtime.h>

#include <complex>
#include <cstdint>
#include <memory>
#include <vector>

#include "tables.h"

namespace hfq {

struct GridTables {
  int lang = 0, mang = 0, nang = 0, nrad = 0, Nel = 0, Nang = 0, NI = 0;
  std::vector<double> cth, phi, wang;              // [nang]
  std::vector<double> r, wrad;                     // [Nel*nrad]
  // radial tables per element, [Nel][NI][nrad] (zero rows for missing functions)
  std::vector<double> F, D, L1, F2;
  // angular tables [Nang][nang]
  std::vector<std::complex<double>> Y, Th;
};

GridTables build_atomic_grid(const BasisTables &t, int lang, int mang);

enum GridFlags { GRID_GRAD = 1, GRID_TAU = 2, GRID_LAPL = 4 };

class GridEngine {
 public:
  GridEngine(const BasisTables &t, const GridTables &g, int device, cudaStream_t stream);
  ~GridEngine();
  int64_t npoints() const;
  // densities in libxc layout on the host (any output pointer may be NULL); Pb == NULL: restricted
  void density(const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb, int flags, double *rho, double *sigma,
               double *tau, double *lapl, double *weights, double *Nel, double *Ekin);
  // assembly from functional output (host arrays, libxc layout); uses the density kept on the device
  void fxc(int flags, bool beta, const double *exc, const double *vrho, const double *vsigma, const double *vtau,
           const double *vlapl, double *Ha, int64_t ldHa, double *Hb, int64_t ldHb, double *Exc);

 private:
  struct Impl;
  std::unique_ptr<Impl> p_;
};

}  // namespace hfq
