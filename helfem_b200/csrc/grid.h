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

// Point p = (element e, angular point ia, radial point ir), p = (e*nang + ia)*nrad + ir.
// Basis function (a, r) at p:  value conj(Y_a(ia)) F_r(ir); derivatives along the three
// orthogonal directions conj(Y_a) D_r / s0, conj(Th_a) F_r / s1, conj(i m_a Y_a) F_r / s2;
// Laplacian lfac(p) * conj(Y_a) [ L1_r - l_a(l_a+1) F2_r - m_a^2 F3_r ].
struct GridTables {
  int lang = 0, mang = 0, nang = 0, nrad = 0, Nel = 0, Nang = 0, NI = 0;
  bool pure_m = false;                             // only same-m pairs couple (phi integrated analytically)
  bool same_l_only = false;                        // only a == b couples (spherically averaged atom: per-l cube)
  bool clamp_theta_kin = false;                    // tau: max(theta-direction term, 0) (src/sadatom/dftgrid.cpp:106)
  std::vector<double> cth, phi, wang;              // [nang]
  std::vector<double> r, wrad;                     // [Nel*nrad]
  std::vector<double> wtot, scale[3], lfac;        // [N] per point
  // radial tables per element, [Nel][NI][nrad] (zero rows for missing functions)
  std::vector<double> F, D, L1, F2, F3;
  // angular tables [Nang][nang]
  std::vector<std::complex<double>> Y, Th;
};

// atomic 3D grid (src/atomic/dftgrid.cpp), r x theta x phi
GridTables build_atomic_grid(const BasisTables &t, int lang, int mang);
// spherically averaged atom (src/sadatom/dftgrid.cpp:45-125, :256-328, :464-486): radial points only, weight
// 4 pi w_r r^2, one "angular function" per l with Y = 1 and a theta-direction table i sqrt(l(l+1)), which
// reproduces the l(l+1) rho_l / r^2 term of tau and its Fock contribution while adding nothing to grad rho;
// the l(l+1) parts of the Laplacian cancel identically, as they do in the reference's formula
GridTables build_sadatom_grid(const BasisTables &t);
// diatomic grids: mang <= 1 -> pure-m 2D grid (src/diatomic/dftgrid_purem.cpp, mu x nu, phi analytic),
// mang >= 2 -> general 3D grid (src/diatomic/dftgrid.cpp)
GridTables build_diatomic_grid(const BasisTables &t, int lang, int mang);

enum GridFlags { GRID_GRAD = 1, GRID_TAU = 2, GRID_LAPL = 4 };

class GridEngine {
 public:
  GridEngine(const BasisTables &t, const GridTables &g, int device, cudaStream_t stream);
  ~GridEngine();
  int64_t npoints() const;
  bool polarized() const;     // the last density() call was unrestricted
  int density_flags() const;  // GridFlags of the last density() call
  // densities in libxc layout on the host (any output pointer may be NULL); Pb == NULL: restricted
  void density(const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb, int flags, double *rho, double *sigma,
               double *tau, double *lapl, double *weights, double *Nel, double *Ekin);
  // the same in two steps: density_launch queues the whole chain on the grid stream and returns at once (so that
  // the J/K build can be issued next to it), density_collect waits and hands the results out
  void density_launch(const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb, int flags);
  void density_collect(double *rho, double *sigma, double *tau, double *lapl, double *weights, double *Nel, double *Ekin);
  // built-in functionals on the device (x_func = 1 Slater exchange, <= 0 none) + assembly; H may be device pointers
  void fxc_builtin(int x_func, int c_func, double thr, bool beta, double *Ha, int64_t ldHa, double *Hb, int64_t ldHb,
                   double *Exc);
  static bool builtin_supported(int x_func, int c_func);
  static bool builtin_needs_gradient(int x_func, int c_func);
  static bool builtin_needs_tau(int x_func, int c_func);
  static int builtin_density_flags(int x_func, int c_func);   // GridFlags the density call must compute
  cudaStream_t stream() const;
  // assembly from functional output (host arrays, libxc layout); uses the density kept on the device
  void fxc(int flags, bool beta, const double *exc, const double *vrho, const double *vsigma, const double *vtau,
           const double *vlapl, double *Ha, int64_t ldHa, double *Hb, int64_t ldHb, double *Exc);

 private:
  void assemble(int flags, bool beta, bool exc, bool gga, bool vtau, bool vlapl, double *Ha, int64_t ldHa, double *Hb,
                int64_t ldHb, double *Exc);
  struct Impl;
  std::unique_ptr<Impl> p_;
};

}  // namespace hfq
