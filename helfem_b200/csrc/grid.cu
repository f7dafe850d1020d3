// Device engine of the atomic DFT grid (see grid.h).  All contractions are small FP64 GEMMs on
// pair tables, run through the multi-entry DMMA GEMM kernel of kernels.cuh:
//   density  : Q[(a,b)][(type,ir)] = Pe[(a,b)][(r,c)] . RR[(r,c)][(type,ir)]
//              D_j[ia][ir]          = YY_j^T[ia][(a,b)] . Q[(a,b)][(type_j, ir)]
//   assembly : T_j[ia][(r,c)]       = c_j[ia][ir] . RR_j^T[ir][(r,c)]
//              H_e[(a,b)][(r,c)]   += YY_j[(a,b)][ia] . T_j[ia][(r,c)]
#include "grid.h"

#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

#include "kernels.cuh"
#include "xc_builtin.cuh"

namespace hfq {

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t err__ = (call);                                                                   \
    if (err__ != cudaSuccess)                                                                     \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " + \
                               __FILE__ + ":" + std::to_string(__LINE__));                        \
  } while (0)

namespace {

template <typename T>
struct Buf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t c) {
    if (c <= n) return;
    if (p) cudaFree(p);
    CK(cudaMalloc(&p, c * sizeof(T)));
    n = c;
  }
  void upload(const std::vector<T> &h) {
    alloc(h.size());
    if (!h.empty()) CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  ~Buf() {
    if (p) cudaFree(p);
  }
};

constexpr int NCOMBO = 8;   // 00, 10, 20, 30, 11, 22, 33, L0
constexpr int NRR = 6;      // FF, DF, DD, LF, 2F, 3F
constexpr int NYY = 7;      // YY, ThY, PhY, ThTh, PhPh, YlY, Ym2Y

struct GridDev {
  int Nel, Nang, NA2, NI, NN, nang, nrad, Nrad;   // NA2 = number of coupled angular pairs
  int64_t npe;   // points per element
  const int *efirst, *en;
  const int *pair_a, *pair_b;   // [NA2]
  const int *pair_of;           // [Nang*Nang] pair index or -1
  const int *ang_off, *ang_skip;   // dense index map (boundary functions of dropped shells removed)
  const double *w;       // [N] total quadrature weight
  const double *sc;      // [3][N] scale factors of the three orthogonal directions
  const double *lfac;    // [N] prefactor of the Laplacian
  int clamp1;            // tau: theta-direction term clamped at zero (spherically averaged atom)
  int skip_uncoupled;    // unpack: leave the blocks of uncoupled angular pairs untouched
};

// Pe[e][(a,b)][(r,c)] = P[(a, f+r), (b, f+c)]     grid (Nel, NA2)
__global__ void k_grid_pack(GridDev g, const double *__restrict__ P, int64_t ld, double *__restrict__ Pe) {
  const int e = blockIdx.x, ab = blockIdx.y, a = g.pair_a[ab], b = g.pair_b[ab];
  const int f = g.efirst[e], n = g.en[e], sa = g.ang_skip[a], sb = g.ang_skip[b];
  double *dst = Pe + ((int64_t)e * g.NA2 + ab) * g.NN;
  for (int idx = threadIdx.x; idx < g.NN; idx += blockDim.x) {
    const int r = idx / g.NI, c = idx % g.NI;
    const bool ok = r < n && c < n && f + r >= sa && f + c >= sb;
    dst[idx] = ok ? P[(int64_t)g.ang_off[a] + f + r - sa + ((int64_t)g.ang_off[b] + f + c - sb) * ld] : 0.0;
  }
}

// point quantities from the contracted tables D[j][p], p = (e*nang + ia)*nrad + ir
// out: rho[p], grho[3][p], tau[p], lapl[p]; sums[0] += w rho, sums[1] += w tau
__global__ void k_grid_points(GridDev g, const double *__restrict__ D, int flags, double *__restrict__ rho,
                              double *__restrict__ grho, double *__restrict__ tau, double *__restrict__ lapl,
                              double *__restrict__ sums) {
  const int64_t N = (int64_t)g.Nel * g.npe;
  double sn = 0.0, sk = 0.0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
    const double s0 = g.sc[p], st = g.sc[N + p], sp = g.sc[2 * N + p], w = g.w[p];
    const double d0 = D[p];
    rho[p] = d0;
    sn += w * d0;
    if (flags & GRID_GRAD) {
      grho[p] = 2.0 * D[N + p] / s0;
      grho[N + p] = 2.0 * D[2 * N + p] / st;
      grho[2 * N + p] = 2.0 * D[3 * N + p] / sp;
    }
    if (flags & (GRID_TAU | GRID_LAPL)) {
      const double kth = D[5 * N + p] / (st * st);
      const double krest = D[4 * N + p] / (s0 * s0) + D[6 * N + p] / (sp * sp);
      const double kin = krest + kth;
      // sadatom: tau = (radial term + max(l(l+1) rho_l / r^2, 0)) / 2 (src/sadatom/dftgrid.cpp:100-107); the
      // l(l+1) parts of the Laplacian cancel identically in the spherical average and are left out on both
      // sides (kinetic part here, -l(l+1) f/r^2 in the angular pair table), as in the reference (:110-123)
      const double kt = g.clamp1 ? krest + fmax(kth, 0.0) : kin;
      tau[p] = 0.5 * kt;
      sk += w * 0.5 * kt;
      if (flags & GRID_LAPL) lapl[p] = 2.0 * ((g.clamp1 ? krest : kin) + g.lfac[p] * D[7 * N + p]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    sn += __shfl_down_sync(0xffffffffu, sn, o);
    sk += __shfl_down_sync(0xffffffffu, sk, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums, sn);
    atomicAdd(sums + 1, sk);
  }
}

// assembly weights c_j[p] for one spin channel
//   vr: v_rho, vs_same/vs_ab: v_sigma(ss), v_sigma(ab) (vs_ab == nullptr: restricted, factor 2 on vs_same only),
//   g_same/g_other: gradients [3][N]
__global__ void k_grid_weights(GridDev g, int flags, const double *__restrict__ vr, const double *__restrict__ vs_same,
                               const double *__restrict__ vs_ab, const double *__restrict__ vt,
                               const double *__restrict__ vl, const double *__restrict__ g_same,
                               const double *__restrict__ g_other, int use_vtl, double *__restrict__ C) {
  const int64_t N = (int64_t)g.Nel * g.npe;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
    const double s0 = g.sc[p], st = g.sc[N + p], sp = g.sc[2 * N + p], w = g.w[p];
    C[p] = w * vr[p];
    if (flags & GRID_GRAD) {
      const double sc[3] = {s0, st, sp};
      for (int c = 0; c < 3; c++) {
        double gv = 2.0 * vs_same[p] * g_same[c * N + p];
        if (vs_ab) gv += vs_ab[p] * g_other[c * N + p];
        C[(1 + c) * N + p] = w * gv / sc[c];
      }
    }
    if (use_vtl) {
      double vtl = 0.0;
      if (vt) vtl += 0.5 * vt[p];
      if (vl) vtl += 2.0 * vl[p];
      vtl *= w;
      C[4 * N + p] = vtl / (s0 * s0);
      // sadatom: the l(l+1) matrix takes the tau potential only (src/sadatom/dftgrid.cpp:311-315)
      C[5 * N + p] = (g.clamp1 ? (vt ? 0.5 * vt[p] * w : 0.0) : vtl) / (st * st);
      C[6 * N + p] = vtl / (sp * sp);
    }
    if (vl) C[7 * N + p] = w * vl[p] * g.lfac[p];
  }
}

// H[(a,f+r),(b,f+c)] = sum over elements containing both radial functions of
//   Hs[e][(a,b)][(r,c)] + Hx[e][(a,b)][(r,c)] + Hx[e][(b,a)][(c,r)]
// one CTA per COUPLED angular pair (grid NA2); the blocks of uncoupled pairs are exactly zero and are cleared by the
// caller at memset speed (pure-m grid: 10 071 of 130 321 blocks are coupled for N2)
__global__ void k_grid_unpack(GridDev g, const double *__restrict__ Hs, const double *__restrict__ Hx,
                              double *__restrict__ H, int64_t ld) {
  const int a = g.pair_a[blockIdx.x], b = g.pair_b[blockIdx.x];
  const int sa = g.ang_skip[a], sb = g.ang_skip[b];
  const int pab = blockIdx.x, pba = g.pair_of[b * g.Nang + a];
  for (int idx = threadIdx.x; idx < g.Nrad * g.Nrad; idx += blockDim.x) {
    const int R = idx % g.Nrad, Cc = idx / g.Nrad;
    if (R < sa || Cc < sb) continue;
    double s = 0.0;
    if (pba >= 0)
      for (int e = 0; e < g.Nel; e++) {
        const int r = R - g.efirst[e], c = Cc - g.efirst[e];
        if (r < 0 || c < 0 || r >= g.en[e] || c >= g.en[e]) continue;
        const int64_t eb = (int64_t)e * g.NA2;
        s += Hs[(eb + pab) * g.NN + r * g.NI + c] + Hx[(eb + pab) * g.NN + r * g.NI + c] +
             Hx[(eb + pba) * g.NN + c * g.NI + r];
      }
    H[(int64_t)g.ang_off[a] + R - sa + ((int64_t)g.ang_off[b] + Cc - sb) * ld] = s;
  }
}

// out[i] = sum_s part[s * n + i]  (K-split partial outputs of the density GEMM; combos that were not computed are summed
// too: they are never read)
__global__ void k_sum_partials(const double *__restrict__ part, int nparts, int64_t n, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int s = 0; s < nparts; s++) v += part[(size_t)s * n + i];
    out[i] = v;
  }
}

// interleave spin components: out[p*nc + c] = in[c][p]
__global__ void k_interleave(const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c,
                             int nc, int64_t N, double *__restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
    out[p * nc] = a[p];
    if (nc > 1) out[p * nc + 1] = b[p];
    if (nc > 2) out[p * nc + 2] = c[p];
  }
}
// sigma components from gradients
__global__ void k_sigma(const double *__restrict__ ga, const double *__restrict__ gb, int64_t N, double *__restrict__ saa,
                        double *__restrict__ sab, double *__restrict__ sbb) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
    const double a0 = ga[p], a1 = ga[N + p], a2 = ga[2 * N + p];
    saa[p] = a0 * a0 + a1 * a1 + a2 * a2;
    if (gb) {
      const double b0 = gb[p], b1 = gb[N + p], b2 = gb[2 * N + p];
      sab[p] = a0 * b0 + a1 * b1 + a2 * b2;
      sbb[p] = b0 * b0 + b1 * b1 + b2 * b2;
    }
  }
}
// de-interleave: out[c][p] = in[p*nc + c]
__global__ void k_deinterleave(const double *__restrict__ in, int nc, int c, int64_t N, double *__restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x)
    out[p] = in[p * nc + c];
}
// Exc = sum w exc (rho_a + rho_b)
__global__ void k_exc(const double *__restrict__ w, const double *__restrict__ exc, const double *__restrict__ ra,
                      const double *__restrict__ rb, int64_t N, double *__restrict__ sum) {
  double s = 0.0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x)
    s += w[p] * exc[p] * (ra[p] + (rb ? rb[p] : 0.0));
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(sum, s);
}

// Built-in functionals (xc_builtin.cuh: libxc ids 1, 7, 101, 130, 202, 231; id <= 0: none) on the device, from the
// densities, gradients and kinetic energy densities of the last density call: exc per particle, v_rho, v_sigma and
// v_tau in libxc's conventions, summed over the exchange and the correlation functional like
// DFTGridWorkerBase::compute_xc accumulates them; zero below the density threshold (xc_func_set_dens_threshold,
// src/general/dftgrid_common.cpp:131).  Polarised densities: exchange only, through the spin-scaling relation
// E_x[na, nb] = (E_x[2 na] + E_x[2 nb]) / 2 (v_sigma(ab) = 0).
// ga / gb: gradient components [3][N]; ta / tb: tau (null for LDAs / GGAs); the vs* / vt* outputs may be null likewise.
__global__ void k_xc_builtin(int64_t N, int x_func, int c_func, const double *__restrict__ ra, const double *__restrict__ rb,
                             const double *__restrict__ ga, const double *__restrict__ gb, const double *__restrict__ ta,
                             const double *__restrict__ tb, double thr, double *__restrict__ exc, double *__restrict__ va,
                             double *__restrict__ vb, double *__restrict__ vsa, double *__restrict__ vsab,
                             double *__restrict__ vsb, double *__restrict__ vta, double *__restrict__ vtb) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
    auto sig = [&](const double *g) {
      if (!g) return 0.0;
      const double a0 = g[p], a1 = g[N + p], a2 = g[2 * N + p];
      return a0 * a0 + a1 * a1 + a2 * a2;
    };
    if (!rb) {
      const double n = ra[p];
      double e = 0.0, vr = 0.0, vs = 0.0, vt = 0.0;
      if (!(n < thr)) {
        const double sg = sig(ga), tt = ta ? ta[p] : 1.0;
        for (int k = 0; k < 2; k++) {
          const int id = k ? c_func : x_func;
          if (id <= 0) continue;
          const xc::D2 d = xc::energy(id, n, sg, tt);
          e += d.v;
          vr += d.v + n * d.n;
          vs += n * d.s;
          vt += n * d.t;
        }
      }
      exc[p] = e;
      va[p] = vr;
      if (vsa) vsa[p] = vs;
      if (vta) vta[p] = vt;
    } else {
      const double na = ra[p], nb = rb[p], n = na + nb;
      double e = 0.0, vra = 0.0, vrb = 0.0, vsaa = 0.0, vsbb = 0.0, vtaa = 0.0, vtbb = 0.0;
      if (!(n < thr) && x_func > 0) {
        if (na > 0.0) {
          const xc::D2 d = xc::energy(x_func, 2.0 * na, 4.0 * sig(ga), ta ? 2.0 * ta[p] : 1.0);
          e += na * d.v;
          vra = d.v + 2.0 * na * d.n;
          vsaa = 4.0 * na * d.s;
          vtaa = 2.0 * na * d.t;
        }
        if (nb > 0.0) {
          const xc::D2 d = xc::energy(x_func, 2.0 * nb, 4.0 * sig(gb), tb ? 2.0 * tb[p] : 1.0);
          e += nb * d.v;
          vrb = d.v + 2.0 * nb * d.n;
          vsbb = 4.0 * nb * d.s;
          vtbb = 2.0 * nb * d.t;
        }
        e /= n;
      }
      exc[p] = e;
      va[p] = vra;
      vb[p] = vrb;
      if (vsa) {
        vsa[p] = vsaa;
        vsab[p] = 0.0;
        vsb[p] = vsbb;
      }
      if (vta) {
        vta[p] = vtaa;
        vtb[p] = vtbb;
      }
    }
  }
}

}  // namespace

struct GridEngine::Impl {
  int device = 0;
  cudaStream_t st = nullptr;
  GridDev gd{};
  int nbf = 0;
  int64_t N = 0;
  Buf<int> d_efirst, d_en, d_bo_q, d_bo_t, d_pa, d_pb, d_pof, d_aoff, d_askip;
  Buf<double> d_w, d_sc, d_lfac, d_RR, d_RRT, d_YY, d_YYT;
  Buf<double> d_P, d_Pe, d_Q, d_D, d_Dpart, d_dens, d_C, d_T, d_Hs, d_Hx, d_H, d_sums, d_io, d_v;
  int ksplit = 1;
  Buf<dev::GemmItem> d_items;
  Buf<dev::GemmEntry> d_entries;
  bool polarized = false, pure_m = false;
  int dens_flags = 0;
  // GEMM descriptors of the density stages depend only on (spin, flags): built once, kept on the device, so a
  // density evaluation is a chain of launches without a host synchronisation
  struct GemmPlan {
    Buf<dev::GemmItem> items;
    Buf<dev::GemmEntry> entries;
    int n = 0, maxM = 0, maxN = 0;
  };
  std::map<int, std::unique_ptr<GemmPlan>> plans;
  template <typename Build>
  void gemm_cached(int key, int maxM, int maxN, Build build) {
    auto it = plans.find(key);
    if (it == plans.end()) {
      std::vector<dev::GemmItem> items;
      std::vector<dev::GemmEntry> entries;
      build(items, entries);
      auto pl = std::make_unique<GemmPlan>();
      pl->items.upload(items);
      pl->entries.upload(entries);
      pl->n = (int)items.size();
      pl->maxM = maxM;
      pl->maxN = maxN;
      it = plans.emplace(key, std::move(pl)).first;
    }
    GemmPlan &pl = *it->second;
    if (!pl.n) return;
    const dim3 grid((pl.maxN + 63) / 64, (pl.maxM + 63) / 64, (unsigned)pl.n);
    dev::k_gemm<64, 64, 2, 2, false><<<grid, 128, 0, st>>>(pl.items.p, pl.entries.p);
    CK(cudaGetLastError());
  }
  // dens layout per spin s: rho [N], grho [3N], tau [N], lapl [N]  -> 6N
  double *dens(int s, int which) { return d_dens.p + ((size_t)s * 6 + which) * N; }

  void gemm(const std::vector<dev::GemmItem> &items, const std::vector<dev::GemmEntry> &entries, int maxM, int maxN) {
    if (items.empty()) return;
    d_items.alloc(items.size());
    d_entries.alloc(entries.size());
    CK(cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(items[0]), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_entries.p, entries.data(), entries.size() * sizeof(entries[0]), cudaMemcpyHostToDevice, st));
    const dim3 grid((maxN + 63) / 64, (maxM + 63) / 64, (unsigned)items.size());
    dev::k_gemm<64, 64, 2, 2, false><<<grid, 128, 0, st>>>(d_items.p, d_entries.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));   // descriptor buffers are reused by the next stage
  }
};

GridEngine::GridEngine(const BasisTables &t, const GridTables &g, int device, cudaStream_t stream) : p_(new Impl) {
  Impl &s = *p_;
  s.device = device;
  s.st = stream;
  CK(cudaSetDevice(device));
  const int NA = g.Nang, NI = g.NI, NN = NI * NI, nang = g.nang, nrad = g.nrad, Nel = g.Nel;
  s.nbf = t.Nbf();
  s.pure_m = g.pure_m;
  s.N = (int64_t)Nel * nang * nrad;
  s.d_efirst.upload(t.efirst);
  s.d_en.upload(t.en);
  // coupled angular pairs: all of them (3D grid) or same-m only (phi integrated analytically)
  std::vector<int> pa, pb, pof((size_t)NA * NA, -1), aoff(NA), askip(NA);
  for (int a = 0; a < NA; a++)
    for (int b = 0; b < NA; b++)
      if ((!g.pure_m || t.mval[a] == t.mval[b]) && (!g.same_l_only || a == b)) {
        pof[(size_t)a * NA + b] = (int)pa.size();
        pa.push_back(a);
        pb.push_back(b);
      }
  const int NA2 = (int)pa.size();
  {
    int off = 0;
    for (int a = 0; a < NA; a++) {
      askip[a] = (t.drop_first_m_nonzero && t.mval[a] != 0) ? 1 : 0;
      aoff[a] = off;
      off += t.Nrad - askip[a];
    }
  }
  s.d_pa.upload(pa);
  s.d_pb.upload(pb);
  s.d_pof.upload(pof);
  s.d_aoff.upload(aoff);
  s.d_askip.upload(askip);
  s.d_w.upload(g.wtot);
  {
    std::vector<double> sc;
    for (int c = 0; c < 3; c++) sc.insert(sc.end(), g.scale[c].begin(), g.scale[c].end());
    s.d_sc.upload(sc);
  }
  s.d_lfac.upload(g.lfac);
  // radial pair tables: RR[e][(r,c)][type*nrad + ir] and RRT[e][type][ir][(r,c)]
  {
    std::vector<double> RR((size_t)Nel * NN * NRR * nrad, 0.0), RRT((size_t)Nel * NRR * nrad * NN, 0.0);
    const std::vector<double> *rowt[NRR] = {&g.F, &g.D, &g.D, &g.L1, &g.F2, &g.F3};
    const std::vector<double> *colt[NRR] = {&g.F, &g.F, &g.D, &g.F, &g.F, &g.F};
    for (int e = 0; e < Nel; e++)
      for (int ty = 0; ty < NRR; ty++)
        for (int r = 0; r < NI; r++)
          for (int c = 0; c < NI; c++)
            for (int ir = 0; ir < nrad; ir++) {
              const double v = (*rowt[ty])[((size_t)e * NI + r) * nrad + ir] * (*colt[ty])[((size_t)e * NI + c) * nrad + ir];
              RR[((size_t)e * NN + r * NI + c) * NRR * nrad + ty * nrad + ir] = v;
              RRT[(((size_t)e * NRR + ty) * nrad + ir) * NN + r * NI + c] = v;
            }
    s.d_RR.upload(RR);
    s.d_RRT.upload(RRT);
  }
  // angular pair tables: YY[type][(a,b)][ia], YYT[type][ia][(a,b)]
  {
    std::vector<double> YY((size_t)NYY * NA2 * nang), YYT((size_t)NYY * nang * NA2);
    for (int ab = 0; ab < NA2; ab++) {
      const int a = pa[ab], b = pb[ab];
      for (int ia = 0; ia < nang; ia++) {
        const std::complex<double> ya = g.Y[(size_t)a * nang + ia], yb = g.Y[(size_t)b * nang + ia];
        const std::complex<double> ta = g.Th[(size_t)a * nang + ia], tb = g.Th[(size_t)b * nang + ia];
        const double ma = t.mval[a], mb = t.mval[b], la = t.lval[a];
        const std::complex<double> cyy = std::conj(ya) * yb;
        const double v[NYY] = {cyy.real(), (std::conj(ta) * yb).real(), ma * cyy.imag(), (std::conj(ta) * tb).real(),
                               ma * mb * cyy.real(), g.same_l_only ? 0.0 : -la * (la + 1.0) * cyy.real(),
                               -ma * ma * cyy.real()};
        for (int ty = 0; ty < NYY; ty++) {
          YY[((size_t)ty * NA2 + ab) * nang + ia] = v[ty];
          YYT[((size_t)ty * nang + ia) * NA2 + ab] = v[ty];
        }
      }
    }
    s.d_YY.upload(YY);
    s.d_YYT.upload(YYT);
  }
  {
    std::vector<int> boq(std::max(NA2, NN)), bot(std::max(nang, nrad));
    for (size_t k = 0; k < boq.size(); k++) boq[k] = (int)(k * NRR * nrad);
    for (size_t k = 0; k < bot.size(); k++) bot[k] = (int)(k * NN);
    s.d_bo_q.upload(boq);
    s.d_bo_t.upload(bot);
  }
  s.gd = GridDev{Nel, NA, NA2, NI, NN, nang, nrad, t.Nrad, (int64_t)nang * nrad, s.d_efirst.p, s.d_en.p,
                 s.d_pa.p, s.d_pb.p, s.d_pof.p, s.d_aoff.p, s.d_askip.p, s.d_w.p, s.d_sc.p, s.d_lfac.p,
                 g.clamp_theta_kin ? 1 : 0, t.batch > 1 ? 1 : 0};
  // staging copies of dense host matrices (d_P, d_H) are allocated on first use: device-resident callers never need them
  s.d_Pe.alloc((size_t)2 * Nel * NA2 * NN);
  s.d_Q.alloc((size_t)2 * Nel * NA2 * NRR * nrad);
  s.d_D.alloc((size_t)NCOMBO * s.N);
  // stage 2 of the density contracts over all coupled angular pairs (K = NA2: 10 071 for the N2 pure-m grid) into only
  // nang x nrad outputs per element: K is cut into ksplit chunks with their own partial outputs, so that the launch
  // has a few hundred CTAs instead of a few dozen, and k_sum_partials adds them up
  s.ksplit = std::max(1, std::min(16, NA2 / 512));
  if (s.ksplit > 1) s.d_Dpart.alloc((size_t)s.ksplit * NCOMBO * s.N);
  s.d_dens.alloc((size_t)2 * 6 * s.N);
  s.d_C.alloc((size_t)NCOMBO * s.N);
  s.d_T.alloc((size_t)(NCOMBO + 2) * Nel * nang * NN);
  s.d_Hs.alloc((size_t)Nel * NA2 * NN);
  s.d_Hx.alloc((size_t)Nel * NA2 * NN);
  s.d_sums.alloc(4);
  s.d_io.alloc((size_t)3 * s.N);
  s.d_v.alloc((size_t)12 * s.N);
}

GridEngine::~GridEngine() {}
int64_t GridEngine::npoints() const { return p_->N; }
cudaStream_t GridEngine::stream() const { return p_->st; }
bool GridEngine::polarized() const { return p_->polarized; }
int GridEngine::density_flags() const { return p_->dens_flags; }

static bool is_device_pointer(const void *p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice;
}

// Asynchronous part of a density evaluation: everything is queued on the grid stream, nothing is waited for.
void GridEngine::density_launch(const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb, int flags) {
  Impl &s = *p_;
  CK(cudaSetDevice(s.device));
  const GridDev &g = s.gd;
  const int nspin = Pb ? 2 : 1;
  const size_t n = (size_t)s.nbf;
  const int64_t N = s.N;
  s.polarized = Pb != nullptr;
  s.dens_flags = flags;
  CK(cudaMemsetAsync(s.d_sums.p, 0, 4 * sizeof(double), s.st));
  // combos: (yy type, rr type) per D_j; the Laplacian combo has two entries
  static const int yyt[NCOMBO + 2] = {0, 0, 1, 2, 0, 3, 4, 0, 5, 6}, rrt[NCOMBO + 2] = {0, 1, 0, 0, 2, 0, 0, 3, 4, 5};
  for (int sp = 0; sp < nspin; sp++) {
    const double *P = sp ? Pb : Pa;
    int64_t ld = sp ? ldPb : ldPa;
    if (!is_device_pointer(P)) {   // host matrix: stage it; a device matrix is packed where it lies
      s.d_P.alloc((size_t)2 * n * n);
      double *dP = s.d_P.p + (size_t)sp * n * n;
      CK(cudaMemcpy2DAsync(dP, n * sizeof(double), P, ld * sizeof(double), n * sizeof(double), n, cudaMemcpyDefault, s.st));
      P = dP;
      ld = (int64_t)n;
    }
    double *Pe = s.d_Pe.p + (size_t)sp * g.Nel * g.NA2 * g.NN;
    k_grid_pack<<<dim3(g.Nel, g.NA2), 128, 0, s.st>>>(g, P, ld, Pe);
    CK(cudaGetLastError());
    // stage 1: Q[e] = Pe[e] . RR[e]
    double *Q = s.d_Q.p + (size_t)sp * g.Nel * g.NA2 * NRR * g.nrad;
    s.gemm_cached(100 + sp, g.NA2, NRR * g.nrad, [&](std::vector<dev::GemmItem> &items, std::vector<dev::GemmEntry> &entries) {
      for (int e = 0; e < g.Nel; e++) {
        dev::GemmItem it{};
        it.C = Q + (size_t)e * g.NA2 * NRR * g.nrad;
        it.browoff = s.d_bo_q.p;
        it.M = g.NA2;
        it.N = NRR * g.nrad;
        it.K = g.NN;
        it.ent0 = (int)entries.size();
        entries.push_back(dev::GemmEntry{Pe + (size_t)e * g.NA2 * g.NN, s.d_RR.p + (size_t)e * g.NN * NRR * g.nrad, g.NN});
        it.ent1 = (int)entries.size();
        it.accumulate = 0;
        it.ldc = NRR * g.nrad;
        it.alpha = 1.0;
        items.push_back(it);
      }
    });
    // stage 2: D_j[e][ia][ir] = YYT_j . Q[e][:, type_j]
    s.gemm_cached(200 + sp * 16 + (flags & 7), g.nang, g.nrad, [&](std::vector<dev::GemmItem> &items, std::vector<dev::GemmEntry> &entries) {
      for (int j = 0; j < NCOMBO; j++) {
        const bool need = j == 0 || ((flags & GRID_GRAD) && j >= 1 && j <= 3) ||
                          ((flags & (GRID_TAU | GRID_LAPL)) && j >= 4 && j <= 6) || ((flags & GRID_LAPL) && j == 7);
        if (!need) continue;
        for (int e = 0; e < g.Nel; e++)
          for (int ks = 0; ks < s.ksplit; ks++) {
            // K chunk [k0, k1) of the coupled pairs: A columns k0.., B rows k0.. (row offsets of d_bo_q are linear in k)
            const int k0 = (int)((int64_t)g.NA2 * ks / s.ksplit), k1 = (int)((int64_t)g.NA2 * (ks + 1) / s.ksplit);
            dev::GemmItem it{};
            it.C = (s.ksplit > 1 ? s.d_Dpart.p + (size_t)ks * NCOMBO * N : s.d_D.p) + (size_t)j * N + (size_t)e * g.npe;
            it.browoff = s.d_bo_q.p;
            it.M = g.nang;
            it.N = g.nrad;
            it.K = k1 - k0;
            it.ent0 = (int)entries.size();
            const double *Qe = Q + (size_t)e * g.NA2 * NRR * g.nrad + (size_t)k0 * NRR * g.nrad;
            entries.push_back(dev::GemmEntry{s.d_YYT.p + (size_t)yyt[j] * g.nang * g.NA2 + k0, Qe + rrt[j] * g.nrad, g.NA2});
            if (j == 7)
              for (int x = 8; x <= 9; x++)
                entries.push_back(dev::GemmEntry{s.d_YYT.p + (size_t)yyt[x] * g.nang * g.NA2 + k0, Qe + rrt[x] * g.nrad, g.NA2});
            it.ent1 = (int)entries.size();
            it.accumulate = 0;
            it.ldc = g.nrad;
            it.alpha = 1.0;
            items.push_back(it);
          }
      }
    });
    if (s.ksplit > 1) {
      k_sum_partials<<<592, 256, 0, s.st>>>(s.d_Dpart.p, s.ksplit, (int64_t)NCOMBO * N, s.d_D.p);
      CK(cudaGetLastError());
    }
    k_grid_points<<<592, 256, 0, s.st>>>(g, s.d_D.p, flags, s.dens(sp, 0), s.dens(sp, 1), s.dens(sp, 4), s.dens(sp, 5),
                                         s.d_sums.p);
    CK(cudaGetLastError());
  }
}

// Waits for the density chain; outputs in libxc layout (host or device pointers, any may be NULL)
void GridEngine::density_collect(double *rho, double *sigma, double *tau, double *lapl, double *weights, double *Nel,
                                 double *Ekin) {
  Impl &s = *p_;
  CK(cudaSetDevice(s.device));
  const int64_t N = s.N;
  const int nspin = s.polarized ? 2 : 1;
  const int flags = s.dens_flags;
  auto out = [&](double *host, const double *a, const double *b, const double *c, int nc) {
    if (!host) return;
    k_interleave<<<592, 256, 0, s.st>>>(a, b, c, nc, N, s.d_io.p);
    CK(cudaMemcpyAsync(host, s.d_io.p, (size_t)nc * N * sizeof(double), cudaMemcpyDefault, s.st));
    CK(cudaStreamSynchronize(s.st));
  };
  out(rho, s.dens(0, 0), s.dens(1, 0), nullptr, nspin);
  if ((flags & GRID_GRAD) && sigma) {
    double *saa = s.d_v.p, *sab = saa + N, *sbb = sab + N;
    k_sigma<<<592, 256, 0, s.st>>>(s.dens(0, 1), nspin == 2 ? s.dens(1, 1) : nullptr, N, saa, sab, sbb);
    out(sigma, saa, sab, sbb, nspin == 2 ? 3 : 1);
  }
  if ((flags & (GRID_TAU | GRID_LAPL)) && tau) out(tau, s.dens(0, 4), s.dens(1, 4), nullptr, nspin);
  if ((flags & GRID_LAPL) && lapl) out(lapl, s.dens(0, 5), s.dens(1, 5), nullptr, nspin);
  if (weights) CK(cudaMemcpyAsync(weights, s.d_w.p, (size_t)N * sizeof(double), cudaMemcpyDefault, s.st));
  double sums[4];
  CK(cudaMemcpyAsync(sums, s.d_sums.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, s.st));
  CK(cudaStreamSynchronize(s.st));
  if (Nel) *Nel = sums[0];
  if (Ekin) *Ekin = sums[1];
}

void GridEngine::density(const double *Pa, int64_t ldPa, const double *Pb, int64_t ldPb, int flags, double *rho,
                         double *sigma, double *tau, double *lapl, double *weights, double *Nel, double *Ekin) {
  density_launch(Pa, ldPa, Pb, ldPb, flags);
  density_collect(rho, sigma, tau, lapl, weights, Nel, Ekin);
}

// Built-in functionals evaluated on the device from the densities of the last density call, then the assembly:
// x_func / c_func = libxc ids 1 (Slater), 101 (PBE), 202 (TPSS) exchange / 7 (VWN5), 130 (PBE), 231 (TPSS) correlation;
// <= 0: none (H = 0; the HF drivers call eval_Fxc only to integrate Nel).  GGAs need the gradient, meta-GGAs also tau
// from the density call (builtin_density_flags).
bool GridEngine::builtin_needs_gradient(int x_func, int c_func) {
  return (x_func > 0 && xc::is_gga(x_func)) || (c_func > 0 && xc::is_gga(c_func));
}
bool GridEngine::builtin_needs_tau(int x_func, int c_func) {
  return (x_func > 0 && xc::is_mgga(x_func)) || (c_func > 0 && xc::is_mgga(c_func));
}
int GridEngine::builtin_density_flags(int x_func, int c_func) {
  return (builtin_needs_gradient(x_func, c_func) ? GRID_GRAD : 0) | (builtin_needs_tau(x_func, c_func) ? GRID_TAU : 0);
}
bool GridEngine::builtin_supported(int x_func, int c_func) {
  return (x_func <= 0 || (xc::known(x_func) && xc::is_exchange(x_func))) &&
         (c_func <= 0 || (xc::known(c_func) && !xc::is_exchange(c_func)));
}

void GridEngine::fxc_builtin(int x_func, int c_func, double thr, bool beta, double *Ha, int64_t ldHa, double *Hb,
                             int64_t ldHb, double *Exc) {
  Impl &s = *p_;
  CK(cudaSetDevice(s.device));
  const int64_t N = s.N;
  const size_t n = (size_t)s.nbf;
  const int nspin = s.polarized ? 2 : 1;
  if (x_func <= 0 && c_func <= 0) {
    auto zero = [&](double *H, int64_t ld) {
      if (!H) return;
      if (is_device_pointer(H)) {
        CK(cudaMemset2DAsync(H, (size_t)ld * sizeof(double), 0, n * sizeof(double), n, s.st));
      } else {
        for (size_t c = 0; c < n; c++) std::memset(H + c * ld, 0, n * sizeof(double));
      }
    };
    zero(Ha, ldHa);
    if (nspin == 2) zero(Hb, ldHb);
    CK(cudaStreamSynchronize(s.st));
    if (Exc) *Exc = 0.0;
    return;
  }
  if (!builtin_supported(x_func, c_func))
    throw std::logic_error("built-in functionals: exchange 1 (Slater), 101 (PBE), 202 (TPSS); correlation 7 (VWN5), 130 (PBE), 231 (TPSS)");
  if (nspin == 2 && c_func > 0)
    throw std::logic_error("built-in correlation functionals are spin-unpolarised only: use hfq_grid_density + libxc + hfq_grid_fxc");
  const bool gga = builtin_needs_gradient(x_func, c_func), mgga = builtin_needs_tau(x_func, c_func);
  if (gga && !(s.dens_flags & GRID_GRAD)) throw std::logic_error("built-in GGA: the gradient was not computed by the density call");
  if (mgga && !(s.dens_flags & GRID_TAU)) throw std::logic_error("built-in meta-GGA: tau was not computed by the density call");
  double *v = s.d_v.p;
  k_xc_builtin<<<592, 256, 0, s.st>>>(N, x_func, c_func, s.dens(0, 0), nspin == 2 ? s.dens(1, 0) : nullptr,
                                      gga ? s.dens(0, 1) : nullptr, (gga && nspin == 2) ? s.dens(1, 1) : nullptr,
                                      mgga ? s.dens(0, 4) : nullptr, (mgga && nspin == 2) ? s.dens(1, 4) : nullptr, thr,
                                      v + 9 * N, v, v + N, gga ? v + 2 * N : nullptr, gga ? v + 3 * N : nullptr,
                                      gga ? v + 4 * N : nullptr, mgga ? v + 5 * N : nullptr, mgga ? v + 6 * N : nullptr);
  CK(cudaGetLastError());
  assemble((gga ? GRID_GRAD : 0) | (mgga ? GRID_TAU : 0), beta, true, gga, mgga, false, Ha, ldHa, Hb, ldHb, Exc);
}

void GridEngine::fxc(int flags, bool beta, const double *exc, const double *vrho, const double *vsigma, const double *vtau,
                     const double *vlapl, double *Ha, int64_t ldHa, double *Hb, int64_t ldHb, double *Exc) {
  Impl &s = *p_;
  CK(cudaSetDevice(s.device));
  const GridDev &g = s.gd;
  const int64_t N = s.N;
  const int nspin = s.polarized ? 2 : 1;
  const size_t n = (size_t)s.nbf;
  const bool gga = (flags & GRID_GRAD) && vsigma;
  if (gga && !(s.dens_flags & GRID_GRAD)) throw std::logic_error("hfq_grid_fxc: gradient was not computed by hfq_grid_density");
  // device copies of the functional output, de-interleaved: v[0..1] vrho, v[2..4] vsigma, v[5..6] vtau, v[7..8] vlapl, v[9] exc
  auto put = [&](const double *host, int nc, int slot0) {
    if (!host) return;
    CK(cudaMemcpyAsync(s.d_io.p, host, (size_t)nc * N * sizeof(double), cudaMemcpyDefault, s.st));
    for (int c = 0; c < nc; c++) k_deinterleave<<<592, 256, 0, s.st>>>(s.d_io.p, nc, c, N, s.d_v.p + (size_t)(slot0 + c) * N);
    CK(cudaStreamSynchronize(s.st));
  };
  put(vrho, nspin, 0);
  if (gga) put(vsigma, nspin == 2 ? 3 : 1, 2);
  if (vtau) put(vtau, nspin, 5);
  if (vlapl) put(vlapl, nspin, 7);
  if (exc) put(exc, 1, 9);
  assemble(flags, beta, exc != nullptr, gga, vtau != nullptr, vlapl != nullptr, Ha, ldHa, Hb, ldHb, Exc);
}

// Assembly from the de-interleaved functional output held in d_v (v[0..1] vrho, v[2..4] vsigma, v[5..6] vtau,
// v[7..8] vlapl, v[9] exc)
void GridEngine::assemble(int flags, bool beta, bool exc, bool gga, bool vtau, bool vlapl, double *Ha, int64_t ldHa,
                          double *Hb, int64_t ldHb, double *Exc) {
  Impl &s = *p_;
  const GridDev &g = s.gd;
  const int64_t N = s.N;
  const int nspin = s.polarized ? 2 : 1;
  const size_t n = (size_t)s.nbf;
  (void)flags;
  if (exc) {
    CK(cudaMemsetAsync(s.d_sums.p + 2, 0, sizeof(double), s.st));
    k_exc<<<592, 256, 0, s.st>>>(s.d_w.p, s.d_v.p + 9 * N, s.dens(0, 0), nspin == 2 ? s.dens(1, 0) : nullptr, N, s.d_sums.p + 2);
  }
  static const int yyt[NCOMBO + 2] = {0, 0, 1, 2, 0, 3, 4, 0, 5, 6}, rrt[NCOMBO + 2] = {0, 1, 0, 0, 2, 0, 0, 3, 4, 5};
  for (int sp = 0; sp < nspin; sp++) {
    if (sp == 1 && !beta) continue;
    double *v = s.d_v.p;
    const double *vr = v + (size_t)sp * N;
    const double *vs_same = gga ? v + (size_t)(2 + (nspin == 2 ? 2 * sp : 0)) * N : nullptr;
    const double *vs_ab = (gga && nspin == 2) ? v + 3 * N : nullptr;
    const double *vt = vtau ? v + (size_t)(5 + sp) * N : nullptr;
    const double *vl = vlapl ? v + (size_t)(7 + sp) * N : nullptr;
    // reference quirk: the unrestricted branch of the ATOMIC worker applies the tau/laplacian kinetic
    // term only when tau is present (src/atomic/dftgrid.cpp:426); the pure-m worker tests both (:521)
    const int use_vtl = (nspin == 2 && !s.pure_m) ? (vt != nullptr) : (vt != nullptr || vl != nullptr);
    k_grid_weights<<<592, 256, 0, s.st>>>(g, gga ? GRID_GRAD : 0, vr, vs_same, vs_ab, vt, vl, s.dens(sp, 1),
                                          s.dens(1 - sp, 1), use_vtl, s.d_C.p);
    CK(cudaGetLastError());
    // stage 1: T_j[e][ia][(r,c)] = C_j[e][ia][ir] . RRT[e][type_j][ir][(r,c)]
    std::vector<int> combos = {0};
    if (gga) { combos.push_back(1); combos.push_back(2); combos.push_back(3); }
    if (use_vtl) { combos.push_back(4); combos.push_back(5); combos.push_back(6); }
    if (vl) { combos.push_back(7); combos.push_back(8); combos.push_back(9); }
    const size_t tsz = (size_t)g.Nel * g.nang * g.NN;
    // the descriptors depend only on which terms are present (not on the spin: C and T are re-used per spin)
    const int key = (gga ? 1 : 0) | (use_vtl ? 2 : 0) | (vl ? 4 : 0);
    s.gemm_cached(300 + key, g.nang, g.NN, [&](std::vector<dev::GemmItem> &items, std::vector<dev::GemmEntry> &entries) {
      for (int j : combos)
        for (int e = 0; e < g.Nel; e++) {
          dev::GemmItem it{};
          it.C = s.d_T.p + (size_t)j * tsz + (size_t)e * g.nang * g.NN;
          it.browoff = s.d_bo_t.p;
          it.M = g.nang;
          it.N = g.NN;
          it.K = g.nrad;
          it.ent0 = (int)entries.size();
          const int cj = j >= 8 ? 7 : j;
          entries.push_back(dev::GemmEntry{s.d_C.p + (size_t)cj * N + (size_t)e * g.npe,
                                           s.d_RRT.p + ((size_t)e * NRR + rrt[j]) * g.nrad * g.NN, g.nrad});
          it.ent1 = (int)entries.size();
          it.accumulate = 0;
          it.ldc = g.NN;
          it.alpha = 1.0;
          items.push_back(it);
        }
    });
    // stage 2: Hs / Hx [e][(a,b)][(r,c)] = sum_j YY_j . T_j[e]
    s.gemm_cached(400 + key, g.NA2, g.NN, [&](std::vector<dev::GemmItem> &items, std::vector<dev::GemmEntry> &entries) {
      for (int grp = 0; grp < 2; grp++)
        for (int e = 0; e < g.Nel; e++) {
          dev::GemmItem it{};
          it.C = (grp ? s.d_Hx.p : s.d_Hs.p) + (size_t)e * g.NA2 * g.NN;
          it.browoff = s.d_bo_t.p;
          it.M = g.NA2;
          it.N = g.NN;
          it.K = g.nang;
          it.ent0 = (int)entries.size();
          for (int j : combos) {
            const bool sym = (j == 0 || (j >= 4 && j <= 6));
            if (sym != (grp == 0)) continue;
            entries.push_back(dev::GemmEntry{s.d_YY.p + (size_t)yyt[j] * g.NA2 * g.nang,
                                             s.d_T.p + (size_t)j * tsz + (size_t)e * g.nang * g.NN, g.nang});
          }
          it.ent1 = (int)entries.size();
          it.accumulate = 0;
          it.ldc = g.NN;
          it.alpha = 1.0;
          items.push_back(it);   // an item without entries writes zeros
        }
    });
    double *H = sp ? Hb : Ha;
    const int64_t ld = sp ? ldHb : ldHa;
    const bool sparse = g.NA2 < g.Nang * g.Nang && !g.skip_uncoupled;   // uncoupled blocks exist and belong to the matrix
    if (is_device_pointer(H)) {   // device matrix: written in place, with the caller's leading dimension
      if (sparse) CK(cudaMemset2DAsync(H, (size_t)ld * sizeof(double), 0, n * sizeof(double), n, s.st));
      k_grid_unpack<<<g.NA2, 256, 0, s.st>>>(g, s.d_Hs.p, s.d_Hx.p, H, ld);
      CK(cudaGetLastError());
    } else {
      s.d_H.alloc(n * n);
      if (sparse) CK(cudaMemsetAsync(s.d_H.p, 0, n * n * sizeof(double), s.st));
      k_grid_unpack<<<g.NA2, 256, 0, s.st>>>(g, s.d_Hs.p, s.d_Hx.p, s.d_H.p, (int64_t)n);
      CK(cudaGetLastError());
      CK(cudaMemcpy2DAsync(H, ld * sizeof(double), s.d_H.p, n * sizeof(double), n * sizeof(double), n, cudaMemcpyDefault, s.st));
    }
    CK(cudaStreamSynchronize(s.st));
  }
  if (exc && Exc) {
    CK(cudaMemcpyAsync(Exc, s.d_sums.p + 2, sizeof(double), cudaMemcpyDeviceToHost, s.st));
    CK(cudaStreamSynchronize(s.st));
  }
}

}  // namespace hfq
