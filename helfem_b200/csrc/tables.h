// Host-side description of one basis for the J/K engine: the angular function
// list, the radial element structure and the integral caches (cross-element
// "disjoint" factors and the in-element low-rank factor) per multipole channel.
//
// The same structure describes the three reference basis types:
//   atomic   (src/atomic/TwoDBasis.h:83-128):  nch = 1, channel index = L,
//            small = r^L, big = r^(-L-1), factor = pivoted Cholesky, sigma = +1,
//            prefactor 4 pi/(2L+1), couplings = Gaunt coefficients.
//   diatomic (src/diatomic/basis.h:151-188):   nch = 2 (cos^2-modified Gaunt
//            channel "0" and plain Gaunt channel "2"), channel index =
//            (L,|M|), small = P0/P2, big = Q0/Q2, factor = sign-aware Cholesky
//            of the 2-channel kernel, prefactor (-1)^M 4 pi Rh^5 (L-|M|)!/(L+|M|)!.
// It can be filled either by this library's own setup (atomic_setup.cpp,
// diatomic_setup.cpp) or by a caller that already owns the reference's caches
// (hfq_create_from_tables in the C ABI).
#pragma once
#include <cstdint>
#include <vector>

namespace hfq {

enum class BasisKind : int { Atomic = 0, Diatomic = 1, Sadatom = 2 };

struct ChannelBlock {        // one (multipole channel, element)
  int n = 0;                 // functions in the element (Ni)
  int rank = 0;              // columns of the in-element factor
  std::vector<double> small; // nch blocks of n*n, column-major
  std::vector<double> big;   // nch blocks of n*n (may be empty on element 0)
  std::vector<double> B;     // (nch*n*n) x rank, column-major
  std::vector<double> sigma; // rank entries, +-1
};

struct BasisTables {
  BasisKind kind = BasisKind::Atomic;
  int nch = 1;
  int Nrad = 0, Nel = 0;
  std::vector<int> efirst, en;          // element -> first radial function, count
  std::vector<int> lval, mval;          // angular functions, reference order
  bool drop_first_m_nonzero = false;    // diatomic: m != 0 shells lose radial fn 0
  // multipole channels; lmM[i] = -1 means "any M" (atomic: channel = L)
  std::vector<int> lmL, lmM;
  std::vector<double> pref;             // |prefactor| per channel
  bool sign_by_M = false;               // multiply prefactor by (-1)^M
  int Lext = 0;                         // coupling range |lj-li|-Lext .. lj+li+Lext
  std::vector<ChannelBlock> blocks;     // [ilm*Nel + iel]
  // Optional dense pair tensors (kernels that do not factorise across elements: erfc attenuation,
  // src/atomic/TwoDBasis.h rs_ktei).  pair[(ilm*Nel + ei)*Nel + ej] is the exchange-ordered matrix
  // A[(rj*Nj + rk)][(ri*Nj + rl)] (row-major, Ni*Nj square), rj, ri in element ei, rk, rl in ej:
  // K_(ei,ej)(rj, rk) += sum A R(ri, rl).  When present, exchange() uses them for EVERY element pair
  // and coulomb() is unavailable (the reference has no range-separated Coulomb build either).
  std::vector<std::vector<double>> pair;
  bool pairwise() const { return !pair.empty(); }
  // basis parameters kept for bookkeeping / grid construction
  double Rhalf = 0.0;
  // > 1: a BATCH of independent spherically averaged atoms on one radial basis (the two-electron caches do not
  // depend on Z): angular function a = atom * (lmax + 1) + l.  Such tables serve the batched operators of the SAP
  // workload only (radial Coulomb of many densities, radial DFT grid with one "angular point" per atom).
  int batch = 1;
  int Z1 = 0, Z2 = 0;
  int nnodes = 0, nquad = 0;
  std::vector<double> bval;

  int Nang() const { return (int)lval.size(); }
  int Ndummy() const { return Nang() * Nrad; }
  int Nbf() const;                       // after boundary removal
  std::vector<int64_t> pure_idx() const; // dense index -> dummy index
  int channel(int L, int Mabs) const;    // -1 if absent
};

// atomic: Z, lmax, mmax, nelem, nnodes (LIP, primbas=4), Rmax, grid type, zexp, nquad (0 -> 5*nnodes)
BasisTables build_atomic_tables(int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid, double zexp,
                                int nquad);
// Yukawa-screened caches of the same basis (TwoDBasis::compute_yukawa, src/atomic/TwoDBasis.cpp:737-758):
// exchange() on these tables is the reference's rs_exchange() for a Yukawa range separation
// (:1001-1131): i_L / k_L weighted cross-element factors, Yukawa in-element kernel, prefactor 4 pi lambda
BasisTables build_atomic_yukawa_tables(int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid,
                                       double zexp, int nquad, double lambda);
// erfc-attenuated exchange caches (TwoDBasis::compute_erfc, src/atomic/TwoDBasis.cpp:762-771,
// CoulombExchangeFE.h:275-297, RadialBasis.cpp:742-810, quadrature.cpp:201-249, erfc_expn.cpp):
// dense pair tensors of the Green's function Phi_L(mu r, mu r'), prefactor 4 pi mu/(2L+1) (:1083).
// exchange() on these tables is the reference's rs_exchange() for erfc range separation.
BasisTables build_atomic_erfc_tables(int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid,
                                     double zexp, int nquad, double mu);
// Phi_L(Xi, xi) of the erfc expansion (exported for tests)
double erfc_phi(int n, double Xi, double xi);
// spherically averaged atom (src/sadatom/basis.{h,cpp}): one angular function per l (m summed
// out), same radial caches as the atomic basis; exchange couples density block l_in to output
// block l_out through the m-averaged squared Gaunt coefficient (src/sadatom/basis.cpp:209-312)
BasisTables build_sadatom_tables(int Z, int lmax, int nelem, int nnodes, double Rmax, int igrid, double zexp, int nquad);
// a batch of nbatch spherically averaged atoms sharing one radial basis (see BasisTables::batch)
BasisTables build_sadatom_batch_tables(int lmax, int nbatch, int nelem, int nnodes, double Rmax, int igrid, double zexp,
                                       int nquad);
// range-separated caches of the spherically averaged atom (src/sadatom/basis.cpp:154-184): rs = 1 Yukawa
// (param = lambda), rs = 2 erfc (param = mu); exchange() on them is sadatom rs_exchange (:314-420)
BasisTables build_sadatom_rs_tables(int Z, int lmax, int nelem, int nnodes, double Rmax, int igrid, double zexp, int nquad,
                                    int rs, double param);
// diatomic: lmax_per_m[|m|], other arguments as src/diatomic/main.cpp:60-118
BasisTables build_diatomic_tables(int Z1, int Z2, double Rbond, const std::vector<int> &lmax_per_m, int nelem,
                                  int nnodes, double Rmax, int igrid, double zexp, int nquad, int device = -1);
// device >= 0: the in-element kernels and their factorisations are computed on that GPU (tei_device.cu)

// Radial effective-potential ("SAP") table of the spherically averaged atom after an SCF
// (effective_potential_table, src/sadatom/main.cpp:55-107, with coulomb_screening / xc_screening /
// electron_density* / kinetic_energy_density of src/sadatom/basis.cpp).  Pl_a / Pl_b: nl per-l radial density
// matrices (Nrad x Nrad, column-major, concatenated); Pl_b == nullptr: restricted, Pl_a is the total density.
// x_func: 1 = LDA exchange (the SAP functional), <= 0 = no xc screening.  Returns (Nel*nquad + 1) rows x 9
// columns, column-major: r, rho, grad rho, lapl rho, tau, v_coul, v_xc, quadrature weight, Z_eff.
// Host-side post-processing (once per atom), like the reference's.
std::vector<double> sap_table(const BasisTables &t, const double *Pl_a, const double *Pl_b, int nl, int x_func);

// One-electron matrices (overlap, kinetic, nuclear attraction), dense Nbf x Nbf
// column-major -- setup helpers so that callers/tests can run an SCF around the
// Fock-build path (src/atomic/TwoDBasis.cpp:320-375, src/diatomic/basis.cpp:1032-1166).
void one_electron_matrices(const BasisTables &t, std::vector<double> &S, std::vector<double> &T, std::vector<double> &V);
// diatomic basis: straight into caller-owned Nbf x Nbf matrices (no intermediate copies)
void diatomic_one_electron_into(const BasisTables &t, double *S, double *T, double *V);

}  // namespace hfq
