/* libhelfemqc_b200 -- C ABI of the B200-native SCF Fock-build path (J, K).
 *
 * Drop-in boundary for the reference's (susilehtola/HelFEM) hot path.  The
 * reference has no C ABI of its own; each entry point below names the C++
 * member it replaces.  All matrices are dense FP64, column-major, caller
 * owned, with explicit leading dimensions.  Every function returns 0 on
 * success and a negative code on error; the message of the last error of the
 * calling thread is available from hfq_last_error().
 *
 * Two object kinds:
 *   hfq_tables  host-side basis description + integral caches
 *               (what TwoDBasis' constructor and compute_tei() produce)
 *   hfq_ctx     one basis bound to one GPU: device-resident caches and the
 *               CUDA kernels behind coulomb() / exchange()
 * There is no CPU fallback: hfq_create fails if no CUDA device is usable.
 */
#ifndef HELFEM_B200_H
#define HELFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hfq_tables hfq_tables;
typedef struct hfq_ctx hfq_ctx;

#define HFQ_OK 0
#define HFQ_ERR_INVALID (-1)  /* bad argument / size mismatch (reference: std::logic_error)   */
#define HFQ_ERR_STATE (-2)    /* called before the integrals exist ("Primitive teis have not
                                 been computed!", src/atomic/TwoDBasis.cpp:775)              */
#define HFQ_ERR_CUDA (-3)     /* CUDA runtime failure                                         */
#define HFQ_ERR_INTERNAL (-4)

/* flags of the DFT-grid calls: which densities are produced / which potentials are supplied */
#define HFQ_GRAD 1
#define HFQ_TAU 2
#define HFQ_LAPL 4

const char *hfq_last_error(void);
/* OpenMP threads used by the host-side setup (compute_tei, coupling tables) and the host zero-fill / copies of the
 * host-pointer calls.  Launchers such as torchrun export OMP_NUM_THREADS=1; a rank calls this with its share of
 * the cores (the reference's setup loops are OpenMP too: src/diatomic/basis.cpp:1410-1454). */
int hfq_set_host_threads(int n);

/* ---- basis construction + compute_tei() ------------------------------------------------- */

/* Atomic basis: TwoDBasisT<double> ctor + compute_tei()
 * (src/atomic/TwoDBasis.cpp:66-94, :708-735; flags src/atomic/main.cpp:53-129).
 * LIP primitive basis (primbas=4), point nucleus.  nquad = 0 -> 5*nnodes. */
int hfq_tables_atomic(hfq_tables **out, int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid,
                      double zexp, int nquad);

/* Range-separated exchange, Yukawa kernel: TwoDBasisT::compute_yukawa(lambda) (src/atomic/TwoDBasis.cpp:737-758).
 * hfq_exchange on a context created from these tables is TwoDBasisT::rs_exchange(P) (:1001-1131, Yukawa branch).
 */
int hfq_tables_atomic_yukawa(hfq_tables **out, int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid,
                             double zexp, int nquad, double lambda);

/* Range-separated exchange, erfc kernel: TwoDBasisT::compute_erfc(mu) (src/atomic/TwoDBasis.cpp:762-771,
 * CoulombExchangeFE.h:275-297): dense pair tensors rs_ktei for every element pair, prefactor 4 pi mu/(2L+1).
 * hfq_exchange on a context created from these tables is TwoDBasisT::rs_exchange(P) (:1001-1131, erfc branch);
 * hfq_coulomb on it fails (the reference has no range-separated Coulomb build). */
int hfq_tables_atomic_erfc(hfq_tables **out, int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid,
                           double zexp, int nquad, double mu);
/* Hand over the reference's own rs_ktei cache (src/atomic/TwoDBasis.h): for flat index (L*Nel + iel)*Nel + jel a
 * column-major (Ni*Nj) x (Ni*Nj) matrix ktei(kk*Ni + jj, ll*Ni + ii) (utils::exchange_tei), concatenated in flat
 * order; count = total number of doubles.  pref[L] of the tables must already be 4 pi mu/(2L+1). */
int hfq_tables_set_pair_tensors(hfq_tables *t, const double *ktei, int64_t count);
/* copy one pair tensor out in the same layout; returns its number of doubles (0 if the tables have none) */
int64_t hfq_tables_get_pair_tensor(const hfq_tables *t, int L, int iel, int jel, double *out, int64_t cap);
/* Phi_L(Xi, xi) of the erfc expansion (erfc_expn::Phi, libhelfem/src/erfc_expn.cpp:225-235) */
double hfq_erfc_phi(int L, double Xi, double xi);

/* Spherically averaged atom: sadatom::basis::TwoDBasis ctor + compute_tei()
 * (src/sadatom/basis.cpp:50-184).  One angular function per l = 0..lmax; matrices passed to
 * hfq_exchange are the block-diagonal dense form of the reference's per-l Cube
 * (block l at rows/cols [l*Nrad, (l+1)*Nrad)); exchange = src/sadatom/basis.cpp:209-312.
 * The Coulomb matrix of the spherical density (src/sadatom/basis.cpp:186-207) is
 * 4*pi * hfq_coulomb of an lmax = 0 atomic context over the same radial basis. */
int hfq_tables_sadatom(hfq_tables **out, int Z, int lmax, int nelem, int nnodes, double Rmax, int igrid, double zexp,
                       int nquad);

/* A BATCH of nbatch spherically averaged atoms on one radial basis -- the gen_sap_table workload (one SCF per
 * element; src/sadatom/scf.cpp:50-350 evaluates several densities concurrently, :313-350).  The radial two-electron
 * caches do not depend on Z, so all atoms share them: angular function a = atom * (lmax + 1) + l.  A context created
 * from these tables serves two batched, device-resident operators: hfq_coulomb_radial_batch and the radial DFT grid
 * (hfq_grid_attach, then hfq_grid_density / hfq_grid_fxc / hfq_eval_fxc / fxc with x_func = 1): the atom index is
 * laid out along the grid's "angular point" axis, point p = (element, atom, radial node), and matrices are
 * BLOCK-COMPACT: the Nrad x Nrad block of function a at offset a * Nrad * (Nrad + 1) doubles, column-major, passed
 * with leading dimension Nrad (device pointers only).  hfq_coulomb / hfq_exchange are not available on them. */
int hfq_tables_sadatom_batch(hfq_tables **out, int lmax, int nbatch, int nelem, int nnodes, double Rmax, int igrid,
                             double zexp, int nquad);

/* Range-separated caches of the spherically averaged atom: sadatom TwoDBasis::compute_yukawa(lambda) (rs = 1)
 * or compute_erfc(mu) (rs = 2), src/sadatom/basis.cpp:154-184; hfq_exchange on a context created from these
 * tables is sadatom rs_exchange(cube) (:314-420). */
int hfq_tables_sadatom_rs(hfq_tables **out, int Z, int lmax, int nelem, int nnodes, double Rmax, int igrid, double zexp,
                          int nquad, int rs, double param);

/* Diatomic basis: diatomic::basis::TwoDBasis ctor + compute_tei()
 * (src/diatomic/basis.cpp:525-647, :1382-1547; flags src/diatomic/main.cpp:60-118).
 * lmax_per_m[|m|], |m| = 0..nm-1, is the --lmax list. */
int hfq_tables_diatomic(hfq_tables **out, int Z1, int Z2, double Rbond, const int *lmax_per_m, int nm, int nelem,
                        int nnodes, double Rmax, int igrid, double zexp, int nquad);
/* The same with the in-element two-electron kernels computed ON THE GPU `device` (compute_tei is the
 * dominant non-SCF cost at large lmax): nested Gauss-Chebyshev quadratures as FP64 tensor-core GEMMs and one CTA per
 * channel for the sign-aware pivoted Cholesky (csrc/tei_device.cu; src/diatomic/quadrature.cpp:188-257,
 * src/diatomic/basis.cpp:1382-1537).  The tables come back to the host like those of hfq_tables_diatomic; the
 * factors agree with the host path to round-off of the quadrature sums (reconstructed kernels to ~1e-14 of their
 * largest element).  HFQ_ERR_CUDA without that GPU: there is no silent fallback. */
int hfq_tables_diatomic_device(hfq_tables **out, int Z1, int Z2, double Rbond, const int *lmax_per_m, int nm, int nelem,
                               int nnodes, double Rmax, int igrid, double zexp, int nquad, int device);

/* Caller-supplied caches: lets an existing HelFEM build hand over the caches its own
 * compute_tei() produced (disjoint_L/disjoint_m1L/prim_chol, src/atomic/TwoDBasis.h:83-128;
 * disjoint_P0/P2/Q0/Q2, cd_B, cd_sigma, src/diatomic/basis.h:151-188).
 * Block (ilm, iel), flat index ilm*Nel+iel:
 *   small/big: nch matrices n x n column-major, concatenated (big may be all zero on element 0)
 *   B: (nch*n*n) x rank column-major; sigma: rank entries (+-1)
 * The per-block arrays are concatenated in flat-index order. */
typedef struct hfq_tables_desc {
  int kind;               /* 0 atomic, 1 diatomic, 2 sadatom */
  int nch;                /* 2 diatomic, 1 otherwise */
  int Nrad, Nel, Nang, nlm;
  const int *efirst;      /* [Nel] first radial function of element */
  const int *en;          /* [Nel] functions in element */
  const int *lval;        /* [Nang] */
  const int *mval;        /* [Nang] */
  const int *lmL;         /* [nlm] multipole L of channel */
  const int *lmM;         /* [nlm] |M| of channel, -1 = any (atomic) */
  const double *pref;     /* [nlm] |prefactor| */
  const int *rank;        /* [nlm*Nel] */
  const double *small_;   /* concatenated */
  const double *big_;     /* concatenated */
  const double *B;        /* concatenated */
  const double *sigma;    /* concatenated */
  double Rhalf;
} hfq_tables_desc;
int hfq_tables_from_arrays(hfq_tables **out, const hfq_tables_desc *desc);

typedef struct hfq_tables_info {
  int kind, nch, Nrad, Nel, Nang, nlm, Nbf, Ndummy;
} hfq_tables_info;
int hfq_tables_get_info(const hfq_tables *t, hfq_tables_info *info);
/* what: 0 lval, 1 mval, 2 efirst, 3 en, 4 lmL, 5 lmM, 6 rank[nlm*Nel] */
int hfq_tables_get_ints(const hfq_tables *t, int what, int *out, int64_t cap);
/* what: 0 pref[nlm], 1 bval[Nel+1] */
int hfq_tables_get_doubles(const hfq_tables *t, int what, double *out, int64_t cap);
/* copy one cache block out (any pointer may be NULL) */
int hfq_tables_get_block(const hfq_tables *t, int ilm, int iel, double *small_, double *big_, double *B,
                         double *sigma);
/* overlap / kinetic / nuclear attraction, Nbf x Nbf column-major (setup helpers,
 * src/atomic/TwoDBasis.cpp:320-375, src/diatomic/basis.cpp:1032-1166) */
int hfq_tables_one_electron(const hfq_tables *t, double *S, double *T, double *V);
/* Radial effective-potential (SAP) table of a spherically averaged atom after an SCF:
 * effective_potential_table (src/sadatom/main.cpp:55-107) = radii, electron_density{,_gradient,_laplacian},
 * kinetic_energy_density, coulomb_screening, xc_screening, quadrature_weights of src/sadatom/basis.cpp.
 * Pl_a / Pl_b: nl per-l radial density matrices (Nrad x Nrad column-major, concatenated); Pl_b == NULL:
 * restricted (Pl_a holds the total per-l density).  x_func = 1: LDA exchange (the SAP functional,
 * src/general/sap.h:40-43), <= 0: no xc screening.  out: (Nel*nquad + 1) x 9 column-major -- r, rho, grad rho,
 * lapl rho, tau, v_coul, v_xc, weight, Z_eff = Z - (v_coul + v_xc).  Returns the number of rows; with
 * out == NULL the number of doubles needed.  Host-side post-processing, no GPU involved (as in the reference:
 * once per atom, after the SCF). */
int64_t hfq_sap_table(const hfq_tables *t, const double *Pl_a, const double *Pl_b, int nl, int x_func, double *out,
                      int64_t cap);
void hfq_tables_destroy(hfq_tables *t);

/* ---- the Fock-build path ------------------------------------------------------------------- */

/* Bind a basis to CUDA device `device`, upload the caches. */
int hfq_create(hfq_ctx **out, const hfq_tables *t, int device);
void hfq_destroy(hfq_ctx *ctx);
int hfq_nbf(const hfq_ctx *ctx);

/* diatomic::basis::TwoDBasis::set_absm_symmetric (src/diatomic/basis.cpp:913-915) */
int hfq_set_absm_symmetric(hfq_ctx *ctx, int flag);

/* J = coulomb(P): atomic src/atomic/TwoDBasis.cpp:773-877, diatomic src/diatomic/basis.cpp:1627-1816.
 * Host buffers; copies are part of the call. */
int hfq_coulomb(hfq_ctx *ctx, const double *P, int64_t ldP, double *J, int64_t ldJ);
/* K = exchange(P) with the reference's sign (returns -K, added to the Fock matrix):
 * atomic src/atomic/TwoDBasis.cpp:879-999, diatomic src/diatomic/basis.cpp:1818-2089. */
int hfq_exchange(hfq_ctx *ctx, const double *P, int64_t ldP, double *K, int64_t ldK);

/* Same with device-resident matrices on the context's GPU; `stream` is a cudaStream_t
 * (NULL = the context's own stream).  The call returns after the work is complete.
 * shard/nshards (contexts WITHOUT a communicator, see hfq_comm_init): the exchange is split by output block
 * -- unit = (output sector pair, radial element pair), dealt longest-first to the least loaded shard -- and each
 * shard writes only the contributions of its own units (everything else zero), so the sum of the nshards matrices
 * is the full K; J is split over multipoles the same way.  With a communicator the arguments are ignored: rank
 * and size come from it and the results are complete on every rank. */
int hfq_coulomb_device(hfq_ctx *ctx, const double *dP, int64_t ldP, double *dJ, int64_t ldJ, void *stream);
int hfq_exchange_device(hfq_ctx *ctx, const double *dP, int64_t ldP, double *dK, int64_t ldK, int shard,
                        int nshards, void *stream);

/* Radial Coulomb matrices of nb spherical densities in one launch: J_b = fac * coulomb(P_b) with the sadatom
 * convention coulomb(P) = 4 pi J_0(P) (src/sadatom/basis.cpp:186-207; the reference's gensap builder calls it with
 * Prad / 4 pi, src/sadatom/scf.cpp:199).  dP, dJ: nb contiguous Nrad x Nrad device matrices.  Any context over an
 * atomic / sadatom radial basis.  stream: a CUDA stream of the caller (the launch is asynchronous on it), or NULL = the
 * context's own stream, in which case the call returns when the result is complete.  Note that NULL is also the
 * handle of the legacy default stream (PyTorch's default): such callers get the synchronous form and must make sure
 * themselves that dP is complete before the call (all *_device entry points read their inputs on the stream they run on). */
int hfq_coulomb_radial_batch(hfq_ctx *ctx, const double *dP, double *dJ, int nb, double fac, void *stream);

/* Batched symmetric eigensolver for small matrices on the current device (cyclic Jacobi, one CTA per matrix, n <= 118):
 * dA = nb row-major n x n symmetric matrices, overwritten by the eigenvectors (V[i][j] = component i of vector j);
 * dW = nb x n eigenvalues, NOT sorted.  Caller-side helper of the batched atomic SCFs (the reference leaves the
 * eigenproblems to OpenOrbitalOptimizer / Eigen). */
int hfq_syev_batch(double *dA, double *dW, int n, int64_t nb, void *stream);

/* One Fock-build step as the reference's fock_builder issues it (src/diatomic/main.cpp:413-426:
 * J = coulomb(P); K = exchange(P/2) back to back): J = coulomb(P) and K = exchange(kscale * P) from a
 * single upload and a single packed copy of P; the host version copies J back while K is being built.
 * With nshards > 1 (device version) J and K are partial sums to be all-reduced. */
int hfq_coulomb_exchange(hfq_ctx *ctx, const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K,
                         int64_t ldK);
int hfq_coulomb_exchange_device(hfq_ctx *ctx, const double *dP, int64_t ldP, double kscale, double *dJ, int64_t ldJ,
                                double *dK, int64_t ldK, int shard, int nshards, void *stream);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink (the reference's OpenMP axis over output
 * blocks, src/diatomic/basis.cpp:1866-1875, becomes the shard axis) -----------------------------------------------
 * hfq_comm_unique_id: ncclGetUniqueId on rank 0 (128 bytes); the caller distributes it (MPI_Bcast,
 * torch.distributed broadcast, a file ...).  hfq_comm_init: ncclCommInitRank for the context's device -- collective
 * over all ranks.  From then on hfq_exchange_device / hfq_coulomb_exchange_device build only the units this rank
 * owns, complete the compact result with ONE in-place ncclAllGather issued by the library on the build stream
 * (N2 lmax=30: 55 MB in total, each rank contributes 1/nranks), and unpack the complete K on every rank; J is built
 * in full on every rank (0.7 ms; cheaper than a second collective).  NCCL is bound at run time (dlopen of
 * libnccl.so.2, or $HFQ_NCCL_LIB); single-GPU users never load it.  nranks == 1 removes the communicator. */
int hfq_comm_unique_id(void *id128);
int hfq_comm_init(hfq_ctx *ctx, const void *id128, int rank, int nranks);
int hfq_comm_size(const hfq_ctx *ctx);
/* The ownership rule itself (host only, testable without a GPU): owner[i] in [0, nranks) for n work units of the
 * given costs, longest first onto the least loaded rank, ties by index. */
int hfq_shard_assign(const double *cost, int n, int nranks, int *owner);

/* ---- DFT quadrature grid: DFTGrid::eval_Fxc of the atomic basis --------------------------------
 * (src/atomic/dftgrid.h:156,159; worker src/atomic/dftgrid.cpp:51-242, :304-465, :470-576).
 * The functional evaluation itself stays on libxc's definitions: the GPU produces the densities in
 * libxc's layout, the caller (or hfq_eval_fxc for the built-in Slater exchange) evaluates the
 * functional, the GPU assembles the matrix.  Point p = (element, angular point, radial point);
 * spin components are interleaved per point exactly as libxc expects them
 * (src/general/dftgrid_common.cpp:60-73): rho[p*ns + s], sigma[p*3 + {aa,ab,bb}] (1 component if
 * restricted), tau, lapl like rho.  flags: 1 gradient, 2 tau, 4 Laplacian.
 * The matrix and per-point array arguments (P, H, rho .. weights, exc .. vlapl) may be host OR device
 * pointers (copies use cudaMemcpyDefault): with P and H resident on the device a call moves no matrix
 * over PCIe.  Scalars (Nel, Ekin, Exc) are host pointers. */
/* atomic basis: DFTGrid(&basis, ldft, mdft) (src/atomic/dftgrid.h:139).  Diatomic basis: mang <= 1 ->
 * PureMDFTGrid(&basis, ldft) (src/diatomic/dftgrid_purem.h, phi analytic, used at --symmetry >= 1);
 * mang >= 2 -> the general 3D DFTGrid(&basis, ldft, mdft) of src/diatomic/dftgrid.h (no Laplacian).
 * Sadatom basis (hfq_tables_sadatom): the radial-only grid of src/sadatom/dftgrid.h:124,127 (lang, mang ignored);
 * matrices are the per-l cube laid out block-diagonally, (lmax+1)*Nrad square, and only the diagonal blocks
 * of H are meaningful (src/sadatom/dftgrid.cpp:45-125, :256-328, :464-486, :505-653). */
int hfq_grid_attach(hfq_ctx *ctx, int lang, int mang);
int64_t hfq_grid_npoints(const hfq_ctx *ctx);
/* DFTGridWorker::update_density + compute_Nel/compute_Ekin over all elements.  Pb == NULL: restricted.
 * Output pointers may be NULL.  weights[p] = w_ang w_rad r^2. */
int hfq_grid_density(hfq_ctx *ctx, const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb, int flags,
                     double *rho, double *sigma, double *tau, double *lapl, double *weights, double *Nel, double *Ekin);
/* DFTGridWorker::eval_Fxc + eval_Exc for the densities of the last hfq_grid_density call.
 * vsigma/vtau/vlapl/exc may be NULL; Hb is written only if the density was polarised and beta != 0. */
int hfq_grid_fxc(hfq_ctx *ctx, int flags, int beta, const double *exc, const double *vrho, const double *vsigma,
                 const double *vtau, const double *vlapl, double *Ha, int64_t ldHa, double *Hb, int64_t ldHb, double *Exc);
/* DFTGrid::eval_Fxc(x_func, .., c_func, .., P[a,b] -> H[a,b], Exc, Nel, Ekin, beta, thr) for the
 * functionals built into this library and evaluated on the device (csrc/xc_builtin.cuh), libxc ids:
 * x_func = 1 Slater, 101 PBE, 202 TPSS exchange; c_func = 7 VWN5, 130 PBE, 231 TPSS correlation (correlation:
 * restricted densities only; polarised exchange through the spin-scaling relation); <= 0: none -- the HF drivers still
 * call eval_Fxc to integrate Nel.  These are the functionals of the reference's recorded LDA / PBE / TPSS energies
 * (tests/refs/ci.json).  Ekin is returned for the meta-GGA (tau is evaluated), 0 otherwise, as in the reference.  Other ids return HFQ_ERR_INVALID: use hfq_grid_density + libxc + hfq_grid_fxc. */
int hfq_eval_fxc(hfq_ctx *ctx, int x_func, int c_func, const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb,
                 double *Ha, int64_t ldHa, double *Hb, int64_t ldHb, double *Exc, double *Nel, double *Ekin, int beta,
                 double thr);

/* One complete restricted Fock build as the reference's fock_builder issues it (src/diatomic/main.cpp:385-426,
 * src/atomic/main.cpp:385-446): XC = eval_Fxc(x_func, c_func, P), J = coulomb(P), K = exchange(kscale * P), everything
 * device-resident.  The grid's density chain is queued first on its own stream and runs next to the J/K kernels.
 * x_func <= 0 and c_func <= 0 (Hartree-Fock: the reference still calls eval_Fxc, which then only integrates Nel):
 * dHxc may be NULL; if given it is zero-filled (the reference's XC matrix of an HF build).  Built-in functionals
 * (see hfq_eval_fxc) are evaluated on the device.  Needs hfq_grid_attach.  Exc, Nel: host scalars. */
int hfq_fock_build_device(hfq_ctx *ctx, const double *dP, int64_t ldP, double kscale, double *dJ, int64_t ldJ, double *dK,
                          int64_t ldK, int x_func, int c_func, double *dHxc, int64_t ldH, double *Exc, double *Nel,
                          double thr, void *stream);
/* The same with HOST matrices, what a CPU-side SCF driver calls: J, K as hfq_coulomb_exchange (one upload of P, result
 * copies overlapped with the build), the grid pass on the device copy of P.
 * Multi-GPU (context with a communicator, one process per GPU): P, J, K must be memory that EVERY rank can address
 * (one shared segment, ideally page-locked by each rank); every rank uploads its 1/nranks column slice of P over its own
 * PCIe link, one in-place ncclAllGather over NVLink completes the density on every GPU, the build is sharded, and
 * every rank copies back / zero-fills its column slice of J and K.  The matrices are complete when every rank has
 * returned: the caller places a barrier after the call.  Hxc (if given) is written by every rank in full. */
int hfq_fock_build(hfq_ctx *ctx, const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K, int64_t ldK,
                   int x_func, int c_func, double *Hxc, int64_t ldH, double *Exc, double *Nel, double thr);

/* Non-zero structure of the last hfq_exchange* result, for compact collectives / copies:
 * bf_sector[Nbf] = sector id of every basis function; pairs = (row sector, column sector) of the
 * blocks that were written (everything else in K is exactly zero).  Returns the number of pairs. */
int hfq_exchange_output_pattern(const hfq_ctx *ctx, int *bf_sector, int64_t cap_bf, int *pairs, int64_t cap_pairs);
/* the same for the last hfq_coulomb* result */
int hfq_coulomb_output_pattern(const hfq_ctx *ctx, int *bf_sector, int64_t cap_bf, int *pairs, int64_t cap_pairs);

/* Timings / work counters of the last call on this context:
 * out[0..5] = ms {pack, fold, in-element GEMM, cross-element, unpack, total},
 * out[6..8] = executed flops {fold, in-element GEMM, cross-element}, out[9] = kernel launches,
 * out[10] = device bytes held by the context, out[11..13] = algorithmic (unpadded) flops
 * {fold, in-element GEMM, cross-element}, out[14..16] = launches of those three kernels,
 * out[17..18] = bytes copied host->device / device->host by the last host-pointer call (the
 * result copies skip the rows outside the non-zero blocks; those are zero-filled on the host),
 * out[19] = number of hfq_coulomb_exchange calls on this context that ran on a predicted sparse
 * upload of P (pinned host buffers only; verified bit-for-bit against the full upload). */
int hfq_last_timings(const hfq_ctx *ctx, double *out, int n);

#ifdef __cplusplus
}
#endif
#endif /* HELFEM_B200_H */
