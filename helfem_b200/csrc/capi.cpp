// C ABI of libhelfemqc_b200 (see include/helfem_b200.h).
#include "../../include/helfem_b200.h"

#include <omp.h>

#include <cmath>
#include <cstring>
#include <map>
#include <vector>
#include <exception>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>

#include "comm.h"
#include "engine.h"
#include "grid.h"
#include "tables.h"

namespace hfq {
void syev_batch(double *dA, double *dW, int n, int64_t nb, cudaStream_t st);   // solver.cu
}

struct hfq_tables {
  hfq::BasisTables t;
};

struct hfq_ctx {
  std::unique_ptr<hfq::Engine> eng;
  // every DFT grid attached to this basis, keyed by (lang, mang): the reference's diatomic driver holds a 3D
  // DFTGrid and a PureMDFTGrid on one basis (src/diatomic/main.cpp:329-330); `grid` is the selected one
  std::map<std::pair<int, int>, std::unique_ptr<hfq::GridEngine>> grids;
  hfq::GridEngine *grid = nullptr;
  std::mutex mu;  // the reference's gensap driver calls the build from several threads (src/sadatom/scf.cpp:329-334)
};

namespace {
thread_local std::string g_err;

template <typename F>
int guarded(F &&f) {
  try {
    return f();
  } catch (const std::logic_error &e) {
    g_err = e.what();
    return HFQ_ERR_INVALID;
  } catch (const std::runtime_error &e) {
    g_err = e.what();
    return (g_err.find("CUDA") != std::string::npos) ? HFQ_ERR_CUDA : HFQ_ERR_INTERNAL;
  } catch (const std::exception &e) {
    g_err = e.what();
    return HFQ_ERR_INTERNAL;
  } catch (...) {
    g_err = "unknown error";
    return HFQ_ERR_INTERNAL;
  }
}

int fail(int code, const char *msg) {
  g_err = msg;
  return code;
}
}  // namespace

extern "C" {

const char *hfq_last_error(void) { return g_err.c_str(); }

int hfq_tables_atomic(hfq_tables **out, int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid,
                      double zexp, int nquad) {
  if (!out || lmax < 0 || mmax < 0 || mmax > lmax || nelem < 1 || nnodes < 2 || nnodes > 16 || !(Rmax > 0.0))
    return fail(HFQ_ERR_INVALID, "hfq_tables_atomic: invalid argument");
  return guarded([&] {
    auto *h = new hfq_tables;
    h->t = hfq::build_atomic_tables(Z, lmax, mmax, nelem, nnodes, Rmax, igrid, zexp, nquad);
    *out = h;
    return HFQ_OK;
  });
}

int hfq_tables_atomic_yukawa(hfq_tables **out, int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid,
                             double zexp, int nquad, double lambda) {
  if (!out || lmax < 0 || mmax < 0 || mmax > lmax || nelem < 1 || nnodes < 2 || nnodes > 16 || !(Rmax > 0.0) || !(lambda > 0.0))
    return fail(HFQ_ERR_INVALID, "hfq_tables_atomic_yukawa: invalid argument");
  return guarded([&] {
    auto *h = new hfq_tables;
    h->t = hfq::build_atomic_yukawa_tables(Z, lmax, mmax, nelem, nnodes, Rmax, igrid, zexp, nquad, lambda);
    *out = h;
    return HFQ_OK;
  });
}

int hfq_tables_atomic_erfc(hfq_tables **out, int Z, int lmax, int mmax, int nelem, int nnodes, double Rmax, int igrid,
                           double zexp, int nquad, double mu) {
  if (!out || lmax < 0 || mmax < 0 || mmax > lmax || nelem < 1 || nnodes < 2 || nnodes > 16 || !(Rmax > 0.0) || !(mu > 0.0))
    return fail(HFQ_ERR_INVALID, "hfq_tables_atomic_erfc: invalid argument");
  return guarded([&] {
    auto *h = new hfq_tables;
    h->t = hfq::build_atomic_erfc_tables(Z, lmax, mmax, nelem, nnodes, Rmax, igrid, zexp, nquad, mu);
    *out = h;
    return HFQ_OK;
  });
}

int hfq_tables_set_pair_tensors(hfq_tables *h, const double *ktei, int64_t count) {
  if (!h || !ktei) return fail(HFQ_ERR_INVALID, "hfq_tables_set_pair_tensors: null argument");
  hfq::BasisTables &t = h->t;
  if (t.nch != 1) return fail(HFQ_ERR_INVALID, "hfq_tables_set_pair_tensors: one-channel (atomic) tables only");
  const int nlm = (int)t.lmL.size(), Nel = t.Nel;
  int64_t need = 0;
  for (int ei = 0; ei < Nel; ei++)
    for (int ej = 0; ej < Nel; ej++) need += (int64_t)t.en[ei] * t.en[ej] * t.en[ei] * t.en[ej];
  need *= nlm;
  if (count != need) return fail(HFQ_ERR_INVALID, "hfq_tables_set_pair_tensors: unexpected size");
  return guarded([&] {
    t.pair.assign((size_t)nlm * Nel * Nel, {});
    const double *src = ktei;
    for (int L = 0; L < nlm; L++)
      for (int ei = 0; ei < Nel; ei++)
        for (int ej = 0; ej < Nel; ej++) {
          const int Ni = t.en[ei], Nj = t.en[ej];
          const size_t M = (size_t)Ni * Nj;
          std::vector<double> &A = t.pair[((size_t)L * Nel + ei) * Nel + ej];
          A.resize(M * M);
          // reference: ktei(rk*Ni + rj, rl*Ni + ri), column-major  ->  A[(rj*Nj + rk)][(ri*Nj + rl)]
          for (int rj = 0; rj < Ni; rj++)
            for (int rk = 0; rk < Nj; rk++)
              for (int ri = 0; ri < Ni; ri++)
                for (int rl = 0; rl < Nj; rl++)
                  A[((size_t)rj * Nj + rk) * M + (size_t)ri * Nj + rl] = src[((size_t)rk * Ni + rj) + ((size_t)rl * Ni + ri) * M];
          src += M * M;
        }
    return HFQ_OK;
  });
}

int64_t hfq_tables_get_pair_tensor(const hfq_tables *h, int L, int iel, int jel, double *out, int64_t cap) {
  if (!h) return fail(HFQ_ERR_INVALID, "hfq_tables_get_pair_tensor: null argument");
  const hfq::BasisTables &t = h->t;
  if (!t.pairwise()) return 0;
  const int nlm = (int)t.lmL.size(), Nel = t.Nel;
  if (L < 0 || L >= nlm || iel < 0 || iel >= Nel || jel < 0 || jel >= Nel)
    return fail(HFQ_ERR_INVALID, "hfq_tables_get_pair_tensor: index out of range");
  const int Ni = t.en[iel], Nj = t.en[jel];
  const size_t M = (size_t)Ni * Nj;
  if (!out) return (int64_t)(M * M);
  if (cap < (int64_t)(M * M)) return fail(HFQ_ERR_INVALID, "hfq_tables_get_pair_tensor: buffer too small");
  const std::vector<double> &A = t.pair[((size_t)L * Nel + iel) * Nel + jel];
  for (int rj = 0; rj < Ni; rj++)
    for (int rk = 0; rk < Nj; rk++)
      for (int ri = 0; ri < Ni; ri++)
        for (int rl = 0; rl < Nj; rl++)
          out[((size_t)rk * Ni + rj) + ((size_t)rl * Ni + ri) * M] = A[((size_t)rj * Nj + rk) * M + (size_t)ri * Nj + rl];
  return (int64_t)(M * M);
}

double hfq_erfc_phi(int L, double Xi, double xi) {
  try {
    return hfq::erfc_phi(L, Xi, xi);
  } catch (const std::exception &e) {
    fail(HFQ_ERR_INTERNAL, e.what());
    return std::nan("");
  }
}

int hfq_tables_sadatom(hfq_tables **out, int Z, int lmax, int nelem, int nnodes, double Rmax, int igrid, double zexp,
                       int nquad) {
  if (!out || lmax < 0 || nelem < 1 || nnodes < 2 || nnodes > 16 || !(Rmax > 0.0))
    return fail(HFQ_ERR_INVALID, "hfq_tables_sadatom: invalid argument");
  return guarded([&] {
    auto *h = new hfq_tables;
    h->t = hfq::build_sadatom_tables(Z, lmax, nelem, nnodes, Rmax, igrid, zexp, nquad);
    *out = h;
    return HFQ_OK;
  });
}

int hfq_tables_sadatom_batch(hfq_tables **out, int lmax, int nbatch, int nelem, int nnodes, double Rmax, int igrid,
                             double zexp, int nquad) {
  if (!out || lmax < 0 || nbatch < 1 || nelem < 1 || nnodes < 2 || nnodes > 16 || !(Rmax > 0.0))
    return fail(HFQ_ERR_INVALID, "hfq_tables_sadatom_batch: invalid argument");
  return guarded([&] {
    auto *h = new hfq_tables;
    h->t = hfq::build_sadatom_batch_tables(lmax, nbatch, nelem, nnodes, Rmax, igrid, zexp, nquad);
    *out = h;
    return HFQ_OK;
  });
}

int hfq_tables_sadatom_rs(hfq_tables **out, int Z, int lmax, int nelem, int nnodes, double Rmax, int igrid, double zexp,
                          int nquad, int rs, double param) {
  if (!out || lmax < 0 || nelem < 1 || nnodes < 2 || nnodes > 16 || !(Rmax > 0.0) || (rs != 1 && rs != 2) || !(param > 0.0))
    return fail(HFQ_ERR_INVALID, "hfq_tables_sadatom_rs: invalid argument");
  return guarded([&] {
    auto *h = new hfq_tables;
    h->t = hfq::build_sadatom_rs_tables(Z, lmax, nelem, nnodes, Rmax, igrid, zexp, nquad, rs, param);
    *out = h;
    return HFQ_OK;
  });
}

int hfq_tables_diatomic(hfq_tables **out, int Z1, int Z2, double Rbond, const int *lmax_per_m, int nm, int nelem,
                        int nnodes, double Rmax, int igrid, double zexp, int nquad) {
  if (!out || !lmax_per_m || nm < 1 || nelem < 1 || nnodes < 2 || nnodes > 16 || !(Rbond > 0.0) || !(Rmax > 0.5 * Rbond))
    return fail(HFQ_ERR_INVALID, "hfq_tables_diatomic: invalid argument");
  for (int m = 0; m < nm; m++)
    if (lmax_per_m[m] < m) return fail(HFQ_ERR_INVALID, "hfq_tables_diatomic: lmax(|m|) < |m|");
  return guarded([&] {
    auto *h = new hfq_tables;
    h->t = hfq::build_diatomic_tables(Z1, Z2, Rbond, std::vector<int>(lmax_per_m, lmax_per_m + nm), nelem, nnodes,
                                      Rmax, igrid, zexp, nquad);
    *out = h;
    return HFQ_OK;
  });
}

int hfq_tables_diatomic_device(hfq_tables **out, int Z1, int Z2, double Rbond, const int *lmax_per_m, int nm, int nelem,
                               int nnodes, double Rmax, int igrid, double zexp, int nquad, int device) {
  if (!out || !lmax_per_m || nm < 1 || nelem < 1 || nnodes < 2 || nnodes > 16 || !(Rbond > 0.0) || !(Rmax > 0.5 * Rbond) ||
      device < 0)
    return fail(HFQ_ERR_INVALID, "hfq_tables_diatomic_device: invalid argument");
  for (int m = 0; m < nm; m++)
    if (lmax_per_m[m] < m) return fail(HFQ_ERR_INVALID, "hfq_tables_diatomic_device: lmax(|m|) < |m|");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev)
    return fail(HFQ_ERR_CUDA, "hfq_tables_diatomic_device: no such CUDA device");
  return guarded([&] {
    auto *h = new hfq_tables;
    h->t = hfq::build_diatomic_tables(Z1, Z2, Rbond, std::vector<int>(lmax_per_m, lmax_per_m + nm), nelem, nnodes,
                                      Rmax, igrid, zexp, nquad, device);
    *out = h;
    return HFQ_OK;
  });
}

int hfq_tables_from_arrays(hfq_tables **out, const hfq_tables_desc *d) {
  if (!out || !d || !d->efirst || !d->en || !d->lval || !d->mval || !d->lmL || !d->lmM || !d->pref || !d->rank ||
      !d->small_ || !d->big_ || !d->B || !d->sigma)
    return fail(HFQ_ERR_INVALID, "hfq_tables_from_arrays: null argument");
  if ((d->kind != 0 && d->kind != 1 && d->kind != 2) || d->nch != (d->kind == 1 ? 2 : 1) || d->Nrad < 1 || d->Nel < 1 ||
      d->Nang < 1 || d->nlm < 1)
    return fail(HFQ_ERR_INVALID, "hfq_tables_from_arrays: inconsistent sizes");
  return guarded([&] {
    auto *h = new hfq_tables;
    hfq::BasisTables &t = h->t;
    t.kind = d->kind == 0 ? hfq::BasisKind::Atomic : d->kind == 1 ? hfq::BasisKind::Diatomic : hfq::BasisKind::Sadatom;
    t.nch = d->nch;
    t.Nrad = d->Nrad;
    t.Nel = d->Nel;
    t.efirst.assign(d->efirst, d->efirst + d->Nel);
    t.en.assign(d->en, d->en + d->Nel);
    t.lval.assign(d->lval, d->lval + d->Nang);
    t.mval.assign(d->mval, d->mval + d->Nang);
    t.lmL.assign(d->lmL, d->lmL + d->nlm);
    t.lmM.assign(d->lmM, d->lmM + d->nlm);
    t.pref.assign(d->pref, d->pref + d->nlm);
    t.drop_first_m_nonzero = d->kind == 1;
    t.sign_by_M = d->kind == 1;
    t.Lext = d->kind == 1 ? 2 : 0;
    t.Rhalf = d->Rhalf;
    for (int e = 0; e < t.Nel; e++)
      if (t.efirst[e] < 0 || t.en[e] < 1 || t.efirst[e] + t.en[e] > t.Nrad) {
        delete h;
        throw std::logic_error("hfq_tables_from_arrays: element outside the radial basis");
      }
    t.blocks.resize((size_t)d->nlm * d->Nel);
    size_t os = 0, ob = 0, og = 0;
    for (int ilm = 0; ilm < d->nlm; ilm++)
      for (int e = 0; e < d->Nel; e++) {
        hfq::ChannelBlock &b = t.blocks[(size_t)ilm * d->Nel + e];
        b.n = t.en[e];
        b.rank = d->rank[(size_t)ilm * d->Nel + e];
        const size_t nn = (size_t)t.nch * b.n * b.n;
        b.small.assign(d->small_ + os, d->small_ + os + nn);
        b.big.assign(d->big_ + os, d->big_ + os + nn);
        os += nn;
        b.B.assign(d->B + ob, d->B + ob + nn * b.rank);
        ob += nn * b.rank;
        b.sigma.assign(d->sigma + og, d->sigma + og + b.rank);
        og += b.rank;
      }
    *out = h;
    return HFQ_OK;
  });
}

int64_t hfq_sap_table(const hfq_tables *h, const double *Pl_a, const double *Pl_b, int nl, int x_func, double *out,
                      int64_t cap) {
  if (!h) return fail(HFQ_ERR_INVALID, "hfq_sap_table: null argument");
  const int64_t need = ((int64_t)h->t.Nel * h->t.nquad + 1) * 9;
  if (!out) return need;
  if (cap < need) return fail(HFQ_ERR_INVALID, "hfq_sap_table: buffer too small");
  return guarded([&] {
    const std::vector<double> v = hfq::sap_table(h->t, Pl_a, Pl_b, nl, x_func);
    std::memcpy(out, v.data(), v.size() * sizeof(double));
    return (int)(v.size() / 9);
  });
}

int hfq_tables_get_info(const hfq_tables *h, hfq_tables_info *info) {
  if (!h || !info) return fail(HFQ_ERR_INVALID, "hfq_tables_get_info: null argument");
  const hfq::BasisTables &t = h->t;
  info->kind = (int)t.kind;
  info->nch = t.nch;
  info->Nrad = t.Nrad;
  info->Nel = t.Nel;
  info->Nang = t.Nang();
  info->nlm = (int)t.lmL.size();
  info->Nbf = t.Nbf();
  info->Ndummy = t.Ndummy();
  return HFQ_OK;
}

int hfq_tables_get_ints(const hfq_tables *h, int what, int *out, int64_t cap) {
  if (!h || !out) return fail(HFQ_ERR_INVALID, "hfq_tables_get_ints: null argument");
  const hfq::BasisTables &t = h->t;
  std::vector<int> ranks;
  const std::vector<int> *src = nullptr;
  switch (what) {
    case 0: src = &t.lval; break;
    case 1: src = &t.mval; break;
    case 2: src = &t.efirst; break;
    case 3: src = &t.en; break;
    case 4: src = &t.lmL; break;
    case 5: src = &t.lmM; break;
    case 6:
      for (const auto &b : t.blocks) ranks.push_back(b.rank);
      src = &ranks;
      break;
    default: return fail(HFQ_ERR_INVALID, "hfq_tables_get_ints: unknown selector");
  }
  if ((int64_t)src->size() > cap) return fail(HFQ_ERR_INVALID, "hfq_tables_get_ints: buffer too small");
  std::memcpy(out, src->data(), src->size() * sizeof(int));
  return (int)src->size();
}

int hfq_tables_get_doubles(const hfq_tables *h, int what, double *out, int64_t cap) {
  if (!h || !out) return fail(HFQ_ERR_INVALID, "hfq_tables_get_doubles: null argument");
  const std::vector<double> *src = what == 0 ? &h->t.pref : what == 1 ? &h->t.bval : nullptr;
  if (!src) return fail(HFQ_ERR_INVALID, "hfq_tables_get_doubles: unknown selector");
  if ((int64_t)src->size() > cap) return fail(HFQ_ERR_INVALID, "hfq_tables_get_doubles: buffer too small");
  std::memcpy(out, src->data(), src->size() * sizeof(double));
  return (int)src->size();
}

int hfq_tables_get_block(const hfq_tables *h, int ilm, int iel, double *small_, double *big_, double *B,
                         double *sigma) {
  if (!h) return fail(HFQ_ERR_INVALID, "hfq_tables_get_block: null argument");
  const hfq::BasisTables &t = h->t;
  if (ilm < 0 || ilm >= (int)t.lmL.size() || iel < 0 || iel >= t.Nel)
    return fail(HFQ_ERR_INVALID, "hfq_tables_get_block: index out of range");
  const hfq::ChannelBlock &b = t.blocks[(size_t)ilm * t.Nel + iel];
  const size_t nn = (size_t)t.nch * b.n * b.n;
  if (small_) std::memcpy(small_, b.small.data(), nn * sizeof(double));
  if (big_) {
    if (b.big.size() == nn)
      std::memcpy(big_, b.big.data(), nn * sizeof(double));
    else
      std::memset(big_, 0, nn * sizeof(double));
  }
  if (B) std::memcpy(B, b.B.data(), b.B.size() * sizeof(double));
  if (sigma) std::memcpy(sigma, b.sigma.data(), b.sigma.size() * sizeof(double));
  return HFQ_OK;
}

int hfq_tables_one_electron(const hfq_tables *h, double *S, double *T, double *V) {
  if (!h || !S || !T || !V) return fail(HFQ_ERR_INVALID, "hfq_tables_one_electron: null argument");
  if (h->t.bval.empty()) return fail(HFQ_ERR_STATE, "hfq_tables_one_electron: tables were not built by this library");
  return guarded([&] {
    if (h->t.kind == hfq::BasisKind::Diatomic) {
      hfq::diatomic_one_electron_into(h->t, S, T, V);
      return HFQ_OK;
    }
    std::vector<double> s, t, v;
    hfq::one_electron_matrices(h->t, s, t, v);
    std::memcpy(S, s.data(), s.size() * sizeof(double));
    std::memcpy(T, t.data(), t.size() * sizeof(double));
    std::memcpy(V, v.data(), v.size() * sizeof(double));
    return HFQ_OK;
  });
}

void hfq_tables_destroy(hfq_tables *h) { delete h; }

int hfq_create(hfq_ctx **out, const hfq_tables *h, int device) {
  if (!out || !h) return fail(HFQ_ERR_INVALID, "hfq_create: null argument");
  if (h->t.blocks.empty())
    return fail(HFQ_ERR_STATE, "Primitive teis have not been computed!");
  return guarded([&] {
    auto c = std::make_unique<hfq_ctx>();
    c->eng = std::make_unique<hfq::Engine>(h->t, device);
    *out = c.release();
    return HFQ_OK;
  });
}

void hfq_destroy(hfq_ctx *ctx) { delete ctx; }

int hfq_nbf(const hfq_ctx *ctx) { return ctx ? ctx->eng->Nbf() : fail(HFQ_ERR_INVALID, "hfq_nbf: null context"); }

int hfq_set_absm_symmetric(hfq_ctx *ctx, int flag) {
  if (!ctx) return fail(HFQ_ERR_INVALID, "hfq_set_absm_symmetric: null context");
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->eng->set_absm_symmetric(flag != 0);
  return HFQ_OK;
}

static int check_mat(const hfq_ctx *ctx, const void *a, int64_t lda, const void *b, int64_t ldb) {
  if (!ctx || !a || !b) return fail(HFQ_ERR_INVALID, "null argument");
  const int64_t n = ctx->eng->Nbf();
  if (lda < n || ldb < n) {
    g_err = "Matrix does not have expected size! Leading dimension smaller than Nbf = " + std::to_string(n);
    return HFQ_ERR_INVALID;
  }
  return HFQ_OK;
}

int hfq_coulomb(hfq_ctx *ctx, const double *P, int64_t ldP, double *J, int64_t ldJ) {
  if (int rc = check_mat(ctx, P, ldP, J, ldJ)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->eng->coulomb(P, ldP, J, ldJ);
    return HFQ_OK;
  });
}

int hfq_exchange(hfq_ctx *ctx, const double *P, int64_t ldP, double *K, int64_t ldK) {
  if (int rc = check_mat(ctx, P, ldP, K, ldK)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->eng->exchange(P, ldP, K, ldK);
    return HFQ_OK;
  });
}

int hfq_coulomb_device(hfq_ctx *ctx, const double *dP, int64_t ldP, double *dJ, int64_t ldJ, void *stream) {
  if (int rc = check_mat(ctx, dP, ldP, dJ, ldJ)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->eng->coulomb_dev(dP, ldP, dJ, ldJ, 0, 1, stream ? (cudaStream_t)stream : ctx->eng->stream());
    return HFQ_OK;
  });
}

int hfq_exchange_device(hfq_ctx *ctx, const double *dP, int64_t ldP, double *dK, int64_t ldK, int shard,
                        int nshards, void *stream) {
  if (int rc = check_mat(ctx, dP, ldP, dK, ldK)) return rc;
  if (nshards < 1 || shard < 0 || shard >= nshards) return fail(HFQ_ERR_INVALID, "hfq_exchange_device: bad shard");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->eng->exchange_dev(dP, ldP, dK, ldK, shard, nshards, stream ? (cudaStream_t)stream : ctx->eng->stream());
    return HFQ_OK;
  });
}

int hfq_coulomb_exchange(hfq_ctx *ctx, const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K,
                         int64_t ldK) {
  if (int rc = check_mat(ctx, P, ldP, J, ldJ)) return rc;
  if (int rc = check_mat(ctx, P, ldP, K, ldK)) return rc;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->eng->coulomb_exchange(P, ldP, kscale, J, ldJ, K, ldK);
    return HFQ_OK;
  });
}

int hfq_coulomb_exchange_device(hfq_ctx *ctx, const double *dP, int64_t ldP, double kscale, double *dJ, int64_t ldJ,
                                double *dK, int64_t ldK, int shard, int nshards, void *stream) {
  if (int rc = check_mat(ctx, dP, ldP, dJ, ldJ)) return rc;
  if (int rc = check_mat(ctx, dP, ldP, dK, ldK)) return rc;
  if (nshards < 1 || shard < 0 || shard >= nshards) return fail(HFQ_ERR_INVALID, "hfq_coulomb_exchange_device: bad shard");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->eng->jk_dev(dP, ldP, kscale, dJ, ldJ, dK, ldK, shard, nshards, stream ? (cudaStream_t)stream : ctx->eng->stream());
    return HFQ_OK;
  });
}

static int output_pattern(const hfq_ctx *ctx, bool coulomb, int *bf_sector, int64_t cap_bf, int *pairs, int64_t cap_pairs);

int hfq_exchange_output_pattern(const hfq_ctx *ctx, int *bf_sector, int64_t cap_bf, int *pairs, int64_t cap_pairs) {
  return output_pattern(ctx, false, bf_sector, cap_bf, pairs, cap_pairs);
}

int hfq_coulomb_output_pattern(const hfq_ctx *ctx, int *bf_sector, int64_t cap_bf, int *pairs, int64_t cap_pairs) {
  return output_pattern(ctx, true, bf_sector, cap_bf, pairs, cap_pairs);
}

static int output_pattern(const hfq_ctx *ctx, bool coulomb, int *bf_sector, int64_t cap_bf, int *pairs, int64_t cap_pairs) {
  if (!ctx || !bf_sector || !pairs) return fail(HFQ_ERR_INVALID, "hfq_*_output_pattern: null argument");
  std::vector<int> bs, pr;
  ctx->eng->output_pattern(bs, pr, coulomb);
  if ((int64_t)bs.size() > cap_bf || (int64_t)pr.size() > cap_pairs)
    return fail(HFQ_ERR_INVALID, "hfq_exchange_output_pattern: buffer too small");
  std::memcpy(bf_sector, bs.data(), bs.size() * sizeof(int));
  std::memcpy(pairs, pr.data(), pr.size() * sizeof(int));
  return (int)(pr.size() / 2);
}

int hfq_grid_attach(hfq_ctx *ctx, int lang, int mang) {
  if (!ctx || lang < 1) return fail(HFQ_ERR_INVALID, "hfq_grid_attach: invalid argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    const hfq::BasisTables &bt = ctx->eng->tables();
    if (bt.kind == hfq::BasisKind::Sadatom) lang = mang = 1;   // radial-only grid: one per basis
    auto it = ctx->grids.find({lang, mang});
    if (it != ctx->grids.end()) {   // already attached: select it (densities of its last call are kept)
      ctx->grid = it->second.get();
      return HFQ_OK;
    }
    // atomic: 3D grid; diatomic: mang <= 1 selects the pure-m grid the reference uses at --symmetry >= 1,
    // mang >= 2 the general 3D grid of --symmetry=0
    const hfq::GridTables g = bt.kind == hfq::BasisKind::Atomic    ? hfq::build_atomic_grid(bt, lang, mang)
                              : bt.kind == hfq::BasisKind::Sadatom ? hfq::build_sadatom_grid(bt)
                                                                   : hfq::build_diatomic_grid(bt, lang, mang);
    auto ge = std::make_unique<hfq::GridEngine>(ctx->eng->tables(), g, ctx->eng->device(), ctx->eng->stream());
    ctx->grid = ge.get();
    ctx->grids[{lang, mang}] = std::move(ge);
    return HFQ_OK;
  });
}

int64_t hfq_grid_npoints(const hfq_ctx *ctx) {
  if (!ctx || !ctx->grid) return fail(HFQ_ERR_STATE, "hfq_grid_npoints: no grid attached");
  return ctx->grid->npoints();
}

int hfq_grid_density(hfq_ctx *ctx, const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb, int flags,
                     double *rho, double *sigma, double *tau, double *lapl, double *weights, double *Nel, double *Ekin) {
  if (!ctx || !ctx->grid) return fail(HFQ_ERR_STATE, "hfq_grid_density: no grid attached");
  if (!Pa) return fail(HFQ_ERR_INVALID, "Error - density matrix is empty!");   // src/atomic/dftgrid.cpp:53-55
  // batch tables: block-compact matrices, leading dimension Nrad (include/helfem_b200.h, hfq_tables_sadatom_batch)
  const int64_t n = ctx->eng->tables().batch > 1 ? ctx->eng->tables().Nrad : ctx->eng->Nbf();
  if (ldPa < n || (Pb && ldPb < n)) return fail(HFQ_ERR_INVALID, "hfq_grid_density: leading dimension smaller than Nbf");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->grid->density(Pa, ldPa, Pb, ldPb, flags, rho, sigma, tau, lapl, weights, Nel, Ekin);
    return HFQ_OK;
  });
}

int hfq_grid_fxc(hfq_ctx *ctx, int flags, int beta, const double *exc, const double *vrho, const double *vsigma,
                 const double *vtau, const double *vlapl, double *Ha, int64_t ldHa, double *Hb, int64_t ldHb, double *Exc) {
  if (!ctx || !ctx->grid) return fail(HFQ_ERR_STATE, "hfq_grid_fxc: no grid attached");
  if (!vrho || !Ha) return fail(HFQ_ERR_INVALID, "hfq_grid_fxc: null argument");
  {
    const int64_t n = ctx->eng->tables().batch > 1 ? ctx->eng->tables().Nrad : ctx->eng->Nbf();
    if (ldHa < n) return fail(HFQ_ERR_INVALID, "hfq_grid_fxc: ldHa smaller than Nbf");
    if (ctx->grid->polarized() && beta && (!Hb || ldHb < n))
      return fail(HFQ_ERR_INVALID, "hfq_grid_fxc: polarised density with beta set needs Hb with ldHb >= Nbf");
    if ((vtau && !(ctx->grid->density_flags() & HFQ_TAU)) || (vlapl && !(ctx->grid->density_flags() & HFQ_LAPL)))
      return fail(HFQ_ERR_INVALID, "hfq_grid_fxc: vtau / vlapl given but tau / the Laplacian was not computed by hfq_grid_density");
  }
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->grid->fxc(flags, beta != 0, exc, vrho, vsigma, vtau, vlapl, Ha, ldHa, Hb, ldHb, Exc);
    return HFQ_OK;
  });
}

int hfq_eval_fxc(hfq_ctx *ctx, int x_func, int c_func, const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb,
                 double *Ha, int64_t ldHa, double *Hb, int64_t ldHb, double *Exc, double *Nel, double *Ekin, int beta,
                 double thr) {
  if (!ctx || !ctx->grid) return fail(HFQ_ERR_STATE, "hfq_eval_fxc: no grid attached");
  if (!Pa || !Ha) return fail(HFQ_ERR_INVALID, "Error - density matrix is empty!");
  {
    const int64_t n = ctx->eng->Nbf();
    if (ldPa < n || ldHa < n || (Pb && ldPb < n) || (Pb && beta && (!Hb || ldHb < n)))
      return fail(HFQ_ERR_INVALID, "hfq_eval_fxc: leading dimension smaller than Nbf, or Hb missing for a polarised build");
  }
  if (!hfq::GridEngine::builtin_supported(x_func, c_func) || (Pb && c_func > 0))
    return fail(HFQ_ERR_INVALID,
                "hfq_eval_fxc: built in are the exchange functionals 1 (Slater), 101 (PBE), 202 (TPSS) and, for restricted "
                "densities, the correlation functionals 7 (VWN5), 130 (PBE), 231 (TPSS); evaluate other functionals with libxc "
                "between hfq_grid_density and hfq_grid_fxc");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    // densities, functional and assembly stay on the device; P and H may be host or device matrices
    ctx->grid->density_launch(Pa, ldPa, Pb, ldPb, hfq::GridEngine::builtin_density_flags(x_func, c_func));
    ctx->grid->density_collect(nullptr, nullptr, nullptr, nullptr, nullptr, Nel, Ekin);
    // tau is evaluated for meta-GGAs only (src/general/dftgrid_common.cpp compute_Ekin)
    if (Ekin && !hfq::GridEngine::builtin_needs_tau(x_func, c_func)) *Ekin = 0.0;
    ctx->grid->fxc_builtin(x_func, c_func, thr, beta != 0, Ha, ldHa, Hb, ldHb, Exc);
    return HFQ_OK;
  });
}

int hfq_fock_build(hfq_ctx *ctx, const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K, int64_t ldK,
                   int x_func, int c_func, double *Hxc, int64_t ldH, double *Exc, double *Nel, double thr) {
  if (!ctx || !P || !J || !K) return fail(HFQ_ERR_INVALID, "hfq_fock_build: null argument");
  if (!ctx->grid) return fail(HFQ_ERR_STATE, "hfq_fock_build: no grid attached");
  const int64_t n = ctx->eng->Nbf();
  if (ldP < n || ldJ < n || ldK < n || (Hxc && ldH < n)) return fail(HFQ_ERR_INVALID, "hfq_fock_build: leading dimension smaller than Nbf");
  if (!hfq::GridEngine::builtin_supported(x_func, c_func))
    return fail(HFQ_ERR_INVALID, "hfq_fock_build: built-in functionals are exchange 1, 101, 202 and correlation 7, 130, 231 (libxc ids)");
  if ((x_func > 0 || c_func > 0) && !Hxc) return fail(HFQ_ERR_INVALID, "hfq_fock_build: Hxc needed for a density functional");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    // J, K through the host-pointer path (one upload of P, copies overlapped with the build); with a communicator
    // the matrices are shared by all ranks and every rank moves its column slice.  The grid then works on the
    // device copy of P that this call left behind: nothing crosses PCIe a second time.
    if (ctx->eng->comm_size() > 1)
      ctx->eng->jk_spmd_host(P, ldP, kscale, J, ldJ, K, ldK);
    else
      ctx->eng->coulomb_exchange(P, ldP, kscale, J, ldJ, K, ldK);
    const hfq::EngineTimings tm = ctx->eng->timings();
    double ekin = 0.0;
    ctx->grid->density_launch(ctx->eng->device_density(), n, nullptr, 0,
                              hfq::GridEngine::builtin_density_flags(x_func, c_func));
    ctx->grid->density_collect(nullptr, nullptr, nullptr, nullptr, nullptr, Nel, &ekin);
    ctx->grid->fxc_builtin(x_func, c_func, thr, true, Hxc, ldH, nullptr, 0, Exc);
    (void)tm;
    return HFQ_OK;
  });
}

int hfq_fock_build_device(hfq_ctx *ctx, const double *dP, int64_t ldP, double kscale, double *dJ, int64_t ldJ, double *dK,
                          int64_t ldK, int x_func, int c_func, double *dHxc, int64_t ldH, double *Exc, double *Nel,
                          double thr, void *stream) {
  if (!ctx || !dP || !dJ || !dK) return fail(HFQ_ERR_INVALID, "hfq_fock_build_device: null argument");
  if (!ctx->grid) return fail(HFQ_ERR_STATE, "hfq_fock_build_device: no grid attached");
  const int64_t n = ctx->eng->Nbf();
  if (ldP < n || ldJ < n || ldK < n || (dHxc && ldH < n))
    return fail(HFQ_ERR_INVALID, "hfq_fock_build_device: leading dimension smaller than Nbf");
  if (!hfq::GridEngine::builtin_supported(x_func, c_func))
    return fail(HFQ_ERR_INVALID, "hfq_fock_build_device: built-in functionals are exchange 1, 101, 202 and correlation 7, 130, 231 (libxc ids)");
  if ((x_func > 0 || c_func > 0) && !dHxc) return fail(HFQ_ERR_INVALID, "hfq_fock_build_device: Hxc needed for a density functional");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->eng->stream();
    // the grid density chain (small GEMMs, bandwidth-bound) is queued first on the grid's stream and runs next to
    // the tensor-pipe-bound J/K kernels
    ctx->eng->fence_stream(st, ctx->grid->stream());   // P may still be written by earlier work on st
    ctx->grid->density_launch(dP, ldP, nullptr, 0, hfq::GridEngine::builtin_density_flags(x_func, c_func));
    ctx->eng->jk_dev(dP, ldP, kscale, dJ, ldJ, dK, ldK, 0, 1, st);
    double ekin = 0.0;
    ctx->grid->density_collect(nullptr, nullptr, nullptr, nullptr, nullptr, Nel, &ekin);
    ctx->grid->fxc_builtin(x_func, c_func, thr, true, dHxc, ldH, nullptr, 0, Exc);
    return HFQ_OK;
  });
}

int hfq_coulomb_radial_batch(hfq_ctx *ctx, const double *dP, double *dJ, int nb, double fac, void *stream) {
  if (!ctx || !dP || !dJ || nb < 0) return fail(HFQ_ERR_INVALID, "hfq_coulomb_radial_batch: invalid argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    const int64_t N = ctx->eng->tables().Nrad;
    ctx->eng->coulomb_radial_batch(dP, dJ, nb, N * N, fac, stream ? (cudaStream_t)stream : ctx->eng->stream());
    // on the context's own stream (stream == NULL, which is also the handle of the legacy default stream that e.g.
    // PyTorch uses) the caller has no way to order later work after the launch: complete it here
    if (!stream && cudaStreamSynchronize(ctx->eng->stream()) != cudaSuccess) throw std::runtime_error("hfq_coulomb_radial_batch: stream synchronisation failed");
    return HFQ_OK;
  });
}

int hfq_syev_batch(double *dA, double *dW, int n, int64_t nb, void *stream) {
  if (!dA || !dW || n < 1 || nb < 0) return fail(HFQ_ERR_INVALID, "hfq_syev_batch: invalid argument");
  return guarded([&] {
    hfq::syev_batch(dA, dW, n, nb, (cudaStream_t)stream);
    return HFQ_OK;
  });
}

int hfq_set_host_threads(int n) {
  if (n < 1) return fail(HFQ_ERR_INVALID, "hfq_set_host_threads: n < 1");
  omp_set_num_threads(n);
  hfq::set_host_threads(n);
  return HFQ_OK;
}

int hfq_comm_unique_id(void *id128) {
  if (!id128) return fail(HFQ_ERR_INVALID, "hfq_comm_unique_id: null argument");
  return guarded([&] {
    hfq::Comm::unique_id(id128);
    return HFQ_OK;
  });
}

int hfq_comm_init(hfq_ctx *ctx, const void *id128, int rank, int nranks) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(HFQ_ERR_INVALID, "hfq_comm_init: invalid argument");
  std::lock_guard<std::mutex> lk(ctx->mu);
  return guarded([&] {
    ctx->eng->set_comm(id128, rank, nranks);
    return HFQ_OK;
  });
}

int hfq_comm_size(const hfq_ctx *ctx) { return ctx ? ctx->eng->comm_size() : 0; }

int hfq_shard_assign(const double *cost, int n, int nranks, int *owner) {
  if (!cost || !owner || n < 0 || nranks < 1) return fail(HFQ_ERR_INVALID, "hfq_shard_assign: invalid argument");
  std::vector<int> o;
  hfq::assign_units(std::vector<double>(cost, cost + n), nranks, o);
  std::memcpy(owner, o.data(), (size_t)n * sizeof(int));
  return HFQ_OK;
}

int hfq_last_timings(const hfq_ctx *ctx, double *out, int n) {
  if (!ctx || !out) return fail(HFQ_ERR_INVALID, "hfq_last_timings: null argument");
  const hfq::EngineTimings &t = ctx->eng->timings();
  const double v[20] = {t.pack, t.fold, t.tgemm, t.offdiag, t.unpack, t.total, t.flops_fold, t.flops_tgemm,
                        t.flops_offdiag, (double)t.launches, (double)ctx->eng->device_bytes(), t.alg_fold,
                        t.alg_tgemm, t.alg_offdiag, (double)t.launches_fold, (double)t.launches_tgemm,
                        (double)t.launches_offdiag, t.h2d_bytes, t.d2h_bytes, (double)ctx->eng->speculative_hits()};
  for (int i = 0; i < n && i < 20; i++) out[i] = v[i];
  return HFQ_OK;
}

}  // extern "C"
