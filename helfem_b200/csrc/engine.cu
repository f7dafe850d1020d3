// Host orchestration of the J/K engine.  See engine.h / kernels.cuh.
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <stdexcept>
#include <thread>

#include <omp.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "comm.h"
#include "kernels.cuh"
#include "special.h"

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
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  size_t *counter = nullptr;
  void alloc(size_t count, size_t *ctr) {
    release();
    n = count;
    counter = ctr;
    if (count) {
      CK(cudaMalloc(&p, count * sizeof(T)));
      if (ctr) *ctr += count * sizeof(T);
    }
  }
  void upload(const std::vector<T> &h, size_t *ctr) {
    alloc(h.size(), ctr);
    if (!h.empty()) CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  void release() {
    if (p) {
      cudaFree(p);
      if (counter) *counter -= n * sizeof(T);
    }
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ntasks == 0: configure the kernel (opt-in shared memory size) on the current device, launch nothing
void launch_fold(int NT, int nch, bool par, const dev::BasisDev &bd, const dev::FoldTask *tasks, int ntasks,
                 const int *pixlist, int maxpix, int64_t totpix, const double *G, const double *Ppix, double *R,
                 cudaStream_t st);

}  // namespace

struct Engine::Impl {
  BasisTables t;   // host copy (without the big factor arrays after upload)
  // sector structure
  int ns = 0, NP = 0, NB = 0, NT = 0, NL = 0, nab = 1, Npix = 0;
  bool parity = false;              // sector positions ordered by l-parity class (halves the fold)
  std::vector<int> sec_cls;
  std::vector<int> sec_span;   // 1 + largest occupied position of a sector (= sec_n unless parity-ordered)
  std::vector<int> sec_m, sec_n, sec_ang, ang_sec, ang_pos, ang_off, ang_skip;
  std::vector<int> sec_lmin, sec_lmax;
  int mmin = 0, mmax = 0, nM = 0;   // M = m_a - m_b range: -(mmax-mmin) .. (mmax-mmin)
  // channels
  std::vector<int64_t> blk_off, B_off, sig_off;
  std::vector<int> ranks;
  std::vector<char> G_nonzero;      // [sp*NL + L]
  // element-pair accumulator layout
  std::vector<int64_t> ep_off;
  int64_t op_stride = 0;
  std::vector<int64_t> tperm_off;   // [nlm * Nel] offset of the tiled in-element kernel A_(ilm,e)
  std::vector<int64_t> tperm_tri_off;   // the same for the rows rj <= rk only (symmetric densities)
  int prio_hi = 0;            // highest stream priority of the device
  bool radial_only = false;   // batch tables: see the constructor
  DevBuf<int> d_bchan;        // coulomb_radial_batch: channel / prefactor per batch entry
  DevBuf<double> d_bfac;
  bool p_symmetric = false;   // set by pack_density: P == P^T to 1e-14 of its largest element
  // device
  DevBuf<int> d_ang_off, d_ang_skip, d_sec_n, d_sec_ang, d_efirst, d_en, d_ang_sec, d_ang_pos, d_rad_e0, d_rad_e1;
  int tg_maxM = 0;
  double tri_pix_frac = 1.0;   // share of the pixels (ri, rl) with el(ri) <= el(rl)
  DevBuf<double> d_G, d_small, d_big, d_B, d_sigma, d_tperm, d_tperm_tri, d_zrow;
  DevBuf<int64_t> d_blk_off, d_B_off, d_sig_off, d_ep_off;
  DevBuf<int> d_rank, d_chan_of, d_browoff_T, d_browoff_P, d_browoff_G, d_splist, d_sp_active;
  DevBuf<double> d_jfac, d_norms, d_Ppix, d_R, d_Kacc, d_Paux, d_JauxT, d_Jsec, d_P, d_O, d_O2;
  DevBuf<dev::FoldTask> d_tasks;
  DevBuf<dev::GemmItem> d_gitems;
  DevBuf<dev::GemmEntry> d_gentries;
  DevBuf<dev::OffItem> d_oitems;
  DevBuf<dev::OffEntry> d_oentries;
  std::vector<int> browoff_T_first;  // per element: first index into d_browoff_T
  std::vector<int> browoff_P_first;  // pair-tensor tables: per element pair, first index into d_browoff_P
  std::vector<int64_t> tpair_off, tpair_tri_off;   // [(ilm*Nel + ei)*Nel + ej] offsets of the tiled pair tensors (-1: absent)
  dev::BasisDev bd{};
  size_t r_slots = 0;                // capacity of the R buffer in task slots
  // density packed by pack_density(): shared by coulomb and exchange in a fused build
  std::vector<double> norms_host;
  unsigned long long *flags_host = nullptr;   // page-locked: per sector pair, bit 0 = not exactly zero, bit 1 = norm not
                                              // below 10 eps; then the bits of max |P - P^T| and of max |P|
  bool want_norms_host = false;      // host-pointer calls: the per-block norms are read back too (upload prediction)
  DevBuf<unsigned long long> d_spflags;
  cudaStream_t aux_stream = nullptr;   // clears the output matrix while the exchange kernels run
  cudaEvent_t ev_start = nullptr, ev_kzero = nullptr, ev_fence = nullptr;
  DevBuf<double> d_Kc;               // compact exchange result: [rank segment][unit][rows][NB]
  std::vector<int> packed_splist;
  bool packed_valid = false;
  double kscale = 1.0;               // exchange(kscale * P) = kscale * exchange(P)
  struct JPlan {   // cached Coulomb descriptors (see coulomb_dev)
    bool valid = false;
    std::vector<int> splist, sp_active;
    std::vector<char> M_active;
    int shard = 0, nshards = 1, nfold = 0, nunfold = 0, nblocks = 0;
    DevBuf<dev::GemmItem> d_items, d_uitems;
    DevBuf<dev::GemmEntry> d_entries, d_uentries;
    DevBuf<int> d_blocks;
  } jplan;
  cudaStream_t copy_stream = nullptr, j_stream = nullptr;
  double j_dense_frac = 0.5;   // share of the columns of J that a host-pointer build returns densely over PCIe
  cudaEvent_t ev_packed = nullptr, ev_jdone = nullptr;
  std::vector<int> pred_r0, pred_r1;   // predicted non-zero row range per column of P (from the last verified call)
  cudaEvent_t ev_j = nullptr, ev_jcopied = nullptr;
  cudaEvent_t ev[8];
};

Engine::Engine(const BasisTables &tin, int device) : p_(new Impl), device_(device) {
  CK(cudaSetDevice(device));
  // The engine's own stream carries the host-pointer calls and, in the fused device build, the DFT-grid chain next
  // to the caller's stream: high priority, so that its small kernels are scheduled ahead of the thousands of queued
  // CTAs of the exchange kernels instead of behind them (the same for the Coulomb and memset side streams).
  int prio_lo = 0, prio_hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CK(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, prio_hi));
  Impl &s = *p_;
  s.prio_hi = prio_hi;
  for (auto &e : s.ev) CK(cudaEventCreate(&e));
  const bool timing = getenv("HFQ_SETUP_TIMING") != nullptr;
  const auto t_ctor = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!timing) return;
    cudaDeviceSynchronize();
    fprintf(stderr, "[hfq create] %-40s at %.3f s\n", what,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t_ctor).count());
  };
  s.t = tin;
  lap("tables copied");
  const BasisTables &t = s.t;
  nbf_ = t.Nbf();
  if (t.Nel > 64) throw std::runtime_error("Engine: more than 64 radial elements not supported");
  for (int e = 0; e < t.Nel; e++)
    if (t.en[e] > 16) throw std::runtime_error("Engine: more than 16 functions per element not supported");
  s.nab = t.nch * t.nch;
  s.Npix = t.Nrad * t.Nrad;
  // ---- flattened cache arrays (shared by the full engine and the radial-only mode of batched atoms)
  const int nlm = (int)t.lmL.size();
  auto upload_caches = [&]() {
    std::vector<double> hsmall, hbig, hB, hsig;
    s.blk_off.assign((size_t)nlm * t.Nel, 0);
    s.B_off.assign((size_t)nlm * t.Nel, 0);
    s.sig_off.assign((size_t)nlm * t.Nel, 0);
    s.ranks.assign((size_t)nlm * t.Nel, 0);
    for (int ilm = 0; ilm < nlm; ilm++)
      for (int e = 0; e < t.Nel; e++) {
        const ChannelBlock &b = t.blocks[(size_t)ilm * t.Nel + e];
        const size_t nn = (size_t)t.nch * b.n * b.n;
        if (b.rank > 128) throw std::runtime_error("Engine: in-element factor rank > 128");
        s.blk_off[(size_t)ilm * t.Nel + e] = (int64_t)hsmall.size();
        hsmall.insert(hsmall.end(), b.small.begin(), b.small.end());
        if (b.big.size() == nn)
          hbig.insert(hbig.end(), b.big.begin(), b.big.end());
        else
          hbig.insert(hbig.end(), nn, 0.0);
        s.B_off[(size_t)ilm * t.Nel + e] = (int64_t)hB.size();
        hB.insert(hB.end(), b.B.begin(), b.B.end());
        s.sig_off[(size_t)ilm * t.Nel + e] = (int64_t)hsig.size();
        hsig.insert(hsig.end(), b.sigma.begin(), b.sigma.end());
        s.ranks[(size_t)ilm * t.Nel + e] = b.rank;
      }
    s.d_small.upload(hsmall, &dev_bytes_);
    s.d_big.upload(hbig, &dev_bytes_);
    s.d_B.upload(hB, &dev_bytes_);
    s.d_sigma.upload(hsig, &dev_bytes_);
    s.d_blk_off.upload(s.blk_off, &dev_bytes_);
    s.d_B_off.upload(s.B_off, &dev_bytes_);
    s.d_sig_off.upload(s.sig_off, &dev_bytes_);
    s.d_rank.upload(s.ranks, &dev_bytes_);
    };
  if (t.batch > 1) {
    // Batch of spherically averaged atoms (BasisTables::batch): only the radial Coulomb operator of the L = 0
    // channel is served (coulomb_radial_batch); no sectors, coupling tables or exchange kernels exist.
    if (t.kind != BasisKind::Sadatom || t.nch != 1) throw std::logic_error("Engine: batch tables must be sadatom tables");
    upload_caches();
    for (auto &b : s.t.blocks) {
      std::vector<double>().swap(b.B);
      std::vector<double>().swap(b.small);
      std::vector<double>().swap(b.big);
    }
    s.d_efirst.upload(t.efirst, &dev_bytes_);
    s.d_en.upload(t.en, &dev_bytes_);
    s.NL = 1;
    s.bd = dev::BasisDev{t.Nang(), t.Nrad, s.Npix, 0, 0, 0, 1, 1, t.Nel, 1, nullptr, nullptr, nullptr, nullptr,
                         s.d_efirst.p, s.d_en.p, nullptr, nullptr};
    s.radial_only = true;
    return;
  }
  // ---- sectors: angular functions grouped by m and, for large expansions, by l-parity.
  // A coupling coefficient vanishes unless l_j + l_i + L is even, so a (m, parity) sector pair
  // couples through one L-parity only, and densities of homonuclear molecules (no even-odd l
  // blocks) are screened at sector granularity exactly as the reference screens them per block.
  const int na = t.Nang();
  bool split_parity = false;
  {
    std::map<int, int> cnt;
    for (int a = 0; a < na; a++) cnt[t.mval[a]]++;
    for (auto &kv : cnt) split_parity |= kv.second > 8;
    if (const char *env = getenv("HFQ_SECTOR_SPLIT")) split_parity = atoi(env) != 0;
  }
  std::map<std::pair<int, int>, std::vector<int>> bym;
  if (t.kind == BasisKind::Sadatom) {
    // spherically averaged atom: every l is its own sector and only l-diagonal blocks exist
    split_parity = false;
    for (int a = 0; a < na; a++) bym[{0, t.lval[a]}].push_back(a);
  } else {
    for (int a = 0; a < na; a++) bym[{t.mval[a], split_parity ? (t.lval[a] & 1) : 0}].push_back(a);
  }
  s.ns = (int)bym.size();
  int nmaxsec = 0;
  for (auto &kv : bym) {
    std::sort(kv.second.begin(), kv.second.end(), [&](int x, int y) { return t.lval[x] < t.lval[y]; });
    nmaxsec = std::max(nmaxsec, (int)kv.second.size());
  }
  const int supported_nt[] = {1, 2, 3, 4, 6, 8};
  s.NT = 0;
  for (int nt : supported_nt)
    if (nt * 8 >= nmaxsec) {
      s.NT = nt;
      break;
    }
  if (!s.NT) throw std::runtime_error("Engine: more than 64 angular functions per m sector not supported");
  s.NP = s.NT * 8;
  s.NB = s.NP * s.NP;
  // l-parity ordering: even-l functions in positions [0, NP/2), odd-l in [NP/2, NP), if both
  // classes of every sector fit into half of the padded size
  s.parity = (s.NT % 2 == 0) && !split_parity;
  if (s.parity)
    for (auto &kv : bym) {
      int cnt[2] = {0, 0};
      for (int a : kv.second) cnt[t.lval[a] & 1]++;
      if (std::max(cnt[0], cnt[1]) > s.NP / 2) s.parity = false;
    }
  s.ang_sec.assign(na, 0);
  s.ang_pos.assign(na, 0);
  s.sec_ang.assign((size_t)s.ns * s.NP, -1);
  {
    int si = 0;
    for (auto &kv : bym) {
      s.sec_m.push_back(kv.first.first);
      s.sec_cls.push_back(kv.first.second);
      s.sec_n.push_back((int)kv.second.size());
      int lmn = 1 << 30, lmx = 0, cnt[2] = {0, 0}, span = 0;
      for (size_t k = 0; k < kv.second.size(); k++) {
        const int a = kv.second[k], cls = t.lval[a] & 1;
        const int pos = s.parity ? cls * (s.NP / 2) + cnt[cls]++ : (int)k;
        s.sec_ang[(size_t)si * s.NP + pos] = a;
        s.ang_sec[a] = si;
        s.ang_pos[a] = pos;
        span = std::max(span, pos + 1);
        lmn = std::min(lmn, t.lval[a]);
        lmx = std::max(lmx, t.lval[a]);
      }
      s.sec_span.push_back(span);
      s.sec_lmin.push_back(lmn);
      s.sec_lmax.push_back(lmx);
      si++;
    }
  }
  lap("sectors");
  s.mmin = *std::min_element(s.sec_m.begin(), s.sec_m.end());
  s.mmax = *std::max_element(s.sec_m.begin(), s.sec_m.end());
  s.nM = 2 * (s.mmax - s.mmin) + 1;
  s.ang_off.assign(na, 0);
  s.ang_skip.assign(na, 0);
  {
    int off = 0;
    for (int a = 0; a < na; a++) {
      s.ang_skip[a] = (t.drop_first_m_nonzero && t.mval[a] != 0) ? 1 : 0;
      s.ang_off[a] = off;
      off += t.Nrad - s.ang_skip[a];
    }
  }
  // ---- multipole range and coupling tables
  int lmaxall = 0;
  for (int l : t.lval) lmaxall = std::max(lmaxall, l);
  s.NL = 0;
  for (int L : t.lmL) s.NL = std::max(s.NL, L + 1);
  {
    const GauntTable gt(std::max(s.NL + 2, lmaxall + 2));
    const size_t gsz = (size_t)s.ns * s.ns * s.NL * t.nch * s.NB;
    std::vector<double> G(gsz, 0.0);
    s.G_nonzero.assign((size_t)s.ns * s.ns * s.NL, 0);
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int sa = 0; sa < s.ns; sa++)
      for (int sb = 0; sb < s.ns; sb++) {
        const int sp = sa * s.ns + sb, ma = s.sec_m[sa], mb = s.sec_m[sb], M = ma - mb, aM = std::abs(M);
        std::vector<char> any(s.NL, 0);
        // diatomic: the cos^2-modified coefficient (GauntTable::mod_coeff, src/general/gaunt.cpp:236-272) is a
        // combination of plain coefficients of neighbouring orders,
        //   mod(L) = c0 w0(L) q(L) + c2 sum_{Lp = max(L-2, 0, |M|)}^{L+2} w2(L, Lp) q(Lp),  q(Lp) = coeff(la, ma, lb, mb, Lp),
        // whose weights depend on (L, M) only: they are tabulated per sector pair and q is evaluated once per order
        // instead of up to six times (same operations in the same order as mod_coeff: identical values)
        const double c0 = 2.0 / 3.0 * std::sqrt(std::acos(-1.0)), c2 = 4.0 / 15.0 * std::sqrt(5.0 * std::acos(-1.0));
        std::vector<double> w0, w2, q;
        if (t.kind == BasisKind::Diatomic) {
          w0.assign(s.NL, 0.0);
          w2.assign((size_t)s.NL * 5, 0.0);
          q.assign(s.NL + 3, 0.0);
          for (int L = aM; L < s.NL; L++) {
            w0[L] = gt.coeff(L, M, 0, 0, L);
            for (int Lp = std::max(std::max(L - 2, 0), aM); Lp <= L + 2; Lp++) w2[(size_t)L * 5 + (Lp - (L - 2))] = gt.coeff(Lp, M, 2, 0, L);
          }
        }
        for (int ia = 0; ia < s.NP; ia++)
          for (int ib = 0; ib < s.NP; ib++) {
            const int anga = s.sec_ang[(size_t)sa * s.NP + ia], angb = s.sec_ang[(size_t)sb * s.NP + ib];
            if (anga < 0 || angb < 0) continue;
            const int la = t.lval[anga], lb = t.lval[angb];
            const int Llo = std::max(std::abs(la - lb) - t.Lext, aM), Lhi = std::min(la + lb + t.Lext, s.NL - 1);
            if (t.kind == BasisKind::Diatomic)
              for (int Lp = std::max(Llo - 2, 0); Lp <= Lhi + 2; Lp++) q[Lp] = gt.coeff(la, ma, lb, mb, Lp);
            for (int L = Llo; L <= Lhi; L++) {
              if (t.channel(L, aM) < 0) continue;
              double *dst = &G[(((size_t)sp * s.NL + L) * t.nch) * s.NB + (size_t)ia * s.NP + ib];
              if (t.kind == BasisKind::Atomic) {
                dst[0] = gt.coeff(la, ma, L, M, lb);
                any[L] |= dst[0] != 0.0;
              } else if (t.kind == BasisKind::Sadatom) {
                // sqrt of the m-averaged squared coupling: both fold factors carry the same entry, so
                // their product is sum_{m,m'} G(la,m,L,m-m',lb)^2 / (2 la + 1)  (src/sadatom/basis.cpp:248-259)
                double tot = 0.0;
                for (int mo = -la; mo <= la; mo++)
                  for (int mi = -lb; mi <= lb; mi++) {
                    const double c = gt.coeff(la, mo, L, mo - mi, lb);
                    tot += c * c;
                  }
                dst[0] = std::sqrt(tot / (2 * la + 1));
                any[L] |= dst[0] != 0.0;
              } else {
                const double cpl0 = w0[L] * q[L];
                double cpl2 = 0.0;
                for (int Lp = std::max(std::max(L - 2, 0), aM); Lp <= L + 2; Lp++) cpl2 += w2[(size_t)L * 5 + (Lp - (L - 2))] * q[Lp];
                dst[0] = c0 * cpl0 + c2 * cpl2;
                dst[s.NB] = -gt.coeff(la, ma, L, M, lb);
                any[L] |= dst[0] != 0.0 || dst[s.NB] != 0.0;
              }
            }
          }
        for (int L = 0; L < s.NL; L++) s.G_nonzero[(size_t)sp * s.NL + L] = any[L];
      }
    s.d_G.upload(G, &dev_bytes_);
  }
  upload_caches();
  // free the big host copies
  for (auto &b : s.t.blocks) {
    std::vector<double>().swap(b.B);
    std::vector<double>().swap(b.small);
    std::vector<double>().swap(b.big);
  }
  // ---- small index arrays
  s.d_ang_off.upload(s.ang_off, &dev_bytes_);
  s.d_ang_skip.upload(s.ang_skip, &dev_bytes_);
  s.d_sec_n.upload(s.sec_n, &dev_bytes_);
  s.d_sec_ang.upload(s.sec_ang, &dev_bytes_);
  s.d_efirst.upload(t.efirst, &dev_bytes_);
  s.d_en.upload(t.en, &dev_bytes_);
  s.d_ang_sec.upload(s.ang_sec, &dev_bytes_);
  s.d_ang_pos.upload(s.ang_pos, &dev_bytes_);
  {
    std::vector<int> e0(t.Nrad, t.Nel), e1(t.Nrad, -1);
    for (int e = 0; e < t.Nel; e++)
      for (int r = t.efirst[e]; r < t.efirst[e] + t.en[e]; r++) {
        e0[r] = std::min(e0[r], e);
        e1[r] = std::max(e1[r], e);
      }
    s.d_rad_e0.upload(e0, &dev_bytes_);
    s.d_rad_e1.upload(e1, &dev_bytes_);
    int64_t keep = 0;
    for (int r = 0; r < t.Nrad; r++)
      for (int c = 0; c < t.Nrad; c++) keep += e0[r] <= e1[c];
    s.tri_pix_frac = (double)keep / ((double)t.Nrad * t.Nrad);
  }
  s.bd = dev::BasisDev{na, t.Nrad, s.Npix, s.NP, s.NB, s.ns, t.nch, s.nab, t.Nel, s.NL,
                       s.d_ang_off.p, s.d_ang_skip.p, s.d_sec_n.p, s.d_sec_ang.p, s.d_efirst.p, s.d_en.p,
                       s.d_rad_e0.p, s.d_rad_e1.p};
  // ---- element-pair accumulator layout
  s.ep_off.assign((size_t)t.Nel * t.Nel, 0);
  {
    int64_t off = 0;
    for (int ei = 0; ei < t.Nel; ei++)
      for (int ej = 0; ej < t.Nel; ej++) {
        s.ep_off[(size_t)ei * t.Nel + ej] = off;
        off += (int64_t)t.en[ei] * t.en[ej] * s.NB;
      }
    s.op_stride = off;
  }
  s.d_ep_off.upload(s.ep_off, &dev_bytes_);
  lap("coupling tables + cache upload");
  // ---- dense exchange-ordered in-element kernels, one per (multipole channel, element), stored as
  //      the pre-swizzled k-chunk tiles k_tgemm_ws bulk-copies (kernels.cuh: tperm_index)
  if (t.pairwise()) {
    // dense pair tensors (erfc): tile + swizzle on the host, every element pair; the symmetric-density
    // variant keeps the rows rj <= rk of in-element blocks and the pairs ei < ej
    if (t.nch != 1) throw std::runtime_error("Engine: pair tensors need a one-channel basis");
    for (int tri = 0; tri < 2; tri++) {
      std::vector<int64_t> &offs = tri ? s.tpair_tri_off : s.tpair_off;
      DevBuf<double> &buf = tri ? s.d_tperm_tri : s.d_tperm;
      const int bk = tri ? dev::TP_BK_TRI : dev::TP_BK;
      offs.assign((size_t)nlm * t.Nel * t.Nel, -1);
      int64_t off = 0;
      for (int ilm = 0; ilm < nlm; ilm++)
        for (int ei = 0; ei < t.Nel; ei++)
          for (int ej = 0; ej < t.Nel; ej++) {
            if (tri && ei > ej) continue;
            const int Ni = t.en[ei], Nj = t.en[ej];
            const int rows = (tri && ei == ej) ? Ni * (Ni + 1) / 2 : Ni * Nj;
            offs[((size_t)ilm * t.Nel + ei) * t.Nel + ej] = off;
            off += dev::tperm_doubles(rows, Ni * Nj, bk);
          }
      std::vector<double> h((size_t)off, 0.0);
      for (int ilm = 0; ilm < nlm; ilm++)
        for (int ei = 0; ei < t.Nel; ei++)
          for (int ej = 0; ej < t.Nel; ej++) {
            const int64_t o = offs[((size_t)ilm * t.Nel + ei) * t.Nel + ej];
            if (o < 0) continue;
            const int Ni = t.en[ei], Nj = t.en[ej], K = Ni * Nj;
            const bool half = tri && ei == ej;
            const int rows = half ? Ni * (Ni + 1) / 2 : K;
            const std::vector<double> &A = t.pair[((size_t)ilm * t.Nel + ei) * t.Nel + ej];
            for (int rj = 0; rj < Ni; rj++)
              for (int rk = 0; rk < Nj; rk++) {
                if (half && rj > rk) continue;
                const int row = half ? rk * (rk + 1) / 2 + rj : rj * Nj + rk;
                for (int col = 0; col < K; col++)
                  h[o + dev::tperm_index(row, col, rows, bk)] = A[((size_t)rj * Nj + rk) * K + col];
              }
          }
      buf.upload(h, &dev_bytes_);
    }
  } else
  for (int tri = 0; tri < 2; tri++) {
    std::vector<int64_t> &offs = tri ? s.tperm_tri_off : s.tperm_off;
    DevBuf<double> &buf = tri ? s.d_tperm_tri : s.d_tperm;
    const int bk = tri ? dev::TP_BK_TRI : dev::TP_BK;
    offs.assign((size_t)nlm * t.Nel, 0);
    int64_t off = 0;
    for (int ilm = 0; ilm < nlm; ilm++)
      for (int e = 0; e < t.Nel; e++) {
        const int n = t.en[e], rows = tri ? n * (n + 1) / 2 : n * n;
        offs[(size_t)ilm * t.Nel + e] = off;
        off += dev::tperm_doubles(rows, s.nab * n * n, bk);
      }
    buf.alloc((size_t)off, &dev_bytes_);
    CK(cudaMemsetAsync(buf.p, 0, (size_t)off * sizeof(double), stream_));
    for (int ilm = 0; ilm < nlm; ilm++)
      for (int e = 0; e < t.Nel; e++) {
        const int n = t.en[e], rows = tri ? n * (n + 1) / 2 : n * n;
        dev::k_build_tperm<<<rows, 256, 0, stream_>>>(
            s.d_B.p + s.B_off[(size_t)ilm * t.Nel + e], s.d_sigma.p + s.sig_off[(size_t)ilm * t.Nel + e], n,
            s.ranks[(size_t)ilm * t.Nel + e], t.nch, t.kind == BasisKind::Atomic ? 1 : 0, tri, bk,
            buf.p + offs[(size_t)ilm * t.Nel + e]);
      }
    CK(cudaGetLastError());
  }
  lap("dense in-element kernels (k_build_tperm)");
  // ---- B-row offset tables
  {
    // T-GEMM: k' = ab*n*n + ri*n + rl -> (ab*Npix + pix(ri,rl)) * NB
    std::vector<int> bo;
    for (int e = 0; e < t.Nel; e++) {
      s.browoff_T_first.push_back((int)bo.size());
      const int n = t.en[e], f = t.efirst[e];
      for (int ab = 0; ab < s.nab; ab++)
        for (int ri = 0; ri < n; ri++)
          for (int rl = 0; rl < n; rl++) {
            const int64_t v = ((int64_t)ab * s.Npix + (int64_t)(f + ri) * t.Nrad + f + rl) * s.NB;
            if (v > 0x7fffffffLL) throw std::runtime_error("Engine: R row offset overflows int32");
            bo.push_back((int)v);
          }
    }
    s.d_browoff_T.upload(bo, &dev_bytes_);
    if (t.pairwise()) {
      // pair tensors: k' = ri*Nj + rl -> pix(ri in ei, rl in ej) * NB
      std::vector<int> bp;
      for (int ei = 0; ei < t.Nel; ei++)
        for (int ej = 0; ej < t.Nel; ej++) {
          s.browoff_P_first.push_back((int)bp.size());
          for (int ri = 0; ri < t.en[ei]; ri++)
            for (int rl = 0; rl < t.en[ej]; rl++) {
              const int64_t v = ((int64_t)(t.efirst[ei] + ri) * t.Nrad + t.efirst[ej] + rl) * s.NB;
              if (v > 0x7fffffffLL) throw std::runtime_error("Engine: R row offset overflows int32");
              bp.push_back((int)v);
            }
        }
      s.d_browoff_P.upload(bp, &dev_bytes_);
    }
    // J unfold: q -> q * NB
    std::vector<int> bg((size_t)s.NL * t.nch);
    for (size_t q = 0; q < bg.size(); q++) bg[q] = (int)(q * s.NB);
    s.d_browoff_G.upload(bg, &dev_bytes_);
  }
  // ---- J radial lookup
  {
    std::vector<int> chan_of((size_t)s.NL * s.nM, -1);
    std::vector<double> jfac((size_t)s.NL * s.nM, 0.0);
    for (int L = 0; L < s.NL; L++)
      for (int Mi = 0; Mi < s.nM; Mi++) {
        const int M = Mi - (s.mmax - s.mmin);
        if (std::abs(M) > L) continue;
        const int ilm = t.channel(L, std::abs(M));
        if (ilm < 0) continue;
        chan_of[(size_t)L * s.nM + Mi] = ilm;
        jfac[(size_t)L * s.nM + Mi] = t.pref[ilm] * ((t.sign_by_M && (M & 1)) ? -1.0 : 1.0);
      }
    s.d_chan_of.upload(chan_of, &dev_bytes_);
    s.d_jfac.upload(jfac, &dev_bytes_);
  }
  // ---- work buffers that do not depend on the density
  s.d_norms.alloc((size_t)na * na, &dev_bytes_);
  s.d_Ppix.alloc((size_t)s.ns * s.ns * s.Npix * s.NB, &dev_bytes_);
  s.d_splist.alloc((size_t)s.ns * s.ns, &dev_bytes_);
  s.d_sp_active.alloc((size_t)s.ns * s.ns, &dev_bytes_);
  CK(cudaStreamSynchronize(stream_));
  // opt-in shared memory sizes
  CK(cudaFuncSetAttribute(dev::k_offdiag_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CK(cudaFuncSetAttribute(dev::k_offdiag_mma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));

  {
    s.tg_maxM = 0;
    for (int e = 0; e < t.Nel; e++) s.tg_maxM = std::max(s.tg_maxM, t.en[e] * t.en[e]);
    if (s.tg_maxM > 256) throw std::runtime_error("elements with more than 16 radial functions are not supported");
    CK(cudaFuncSetAttribute(dev::k_tgemm_ws<dev::TP_BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(dev::k_tgemm_ws<dev::TP_BK_TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  s.d_zrow.upload(std::vector<double>(64, 0.0), &dev_bytes_);
  // opt-in shared memory of the fold kernel this basis uses (a per-device attribute: set by every engine)
  launch_fold(s.NT, t.nch, s.parity, s.bd, nullptr, 0, nullptr, 0, 0, nullptr, nullptr, nullptr, stream_);
  lap("done");
}


namespace {

// maxpix = longest pixel list of a task, totpix = sum over the tasks: the pixel chunk per CTA is sized so that the
// launch has a few thousand CTAs and every CTA amortises its coupling-table load over >= 8 pixels per warp
template <int NT, int NCH, bool PAR>
void launch_fold_t(const dev::BasisDev &bd, const dev::FoldTask *tasks, int ntasks, const int *pixlist, int maxpix,
                   int64_t totpix, const double *G, const double *Ppix, double *R, cudaStream_t st) {
  constexpr int NP = NT * 8, LD = NP + 4;
  if constexpr (NT <= 2 && !PAR) {
    // register-resident two-stage fold, one pixel per warp
    constexpr int LDJ = (NP % 16 == 0) ? NP + 8 : NP + 16;
    const size_t smem = (size_t)(NCH * NP * LD + NCH * NP * LDJ + 8 * 2 * NP * LD) * sizeof(double);
    if (ntasks == 0) {   // configuration call from the Engine constructor: the attribute is per device
      CK(cudaFuncSetAttribute(dev::k_fold_reg<NT, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      return;
    }
    int ppc = std::max(64, (int)((totpix + 148 * 32 - 1) / (148 * 32)));
    ppc = round_up(std::min(std::min(ppc, 256), std::max(maxpix, 8)), 8);
    const dim3 grid((maxpix + ppc - 1) / ppc, ntasks);
    dev::k_fold_reg<NT, NCH><<<grid, 256, smem, st>>>(bd, tasks, pixlist, G, Ppix, R, ppc);
    CK(cudaGetLastError());
    return;
  }
  if (ntasks == 0) {
    CK(cudaFuncSetAttribute(dev::k_fold<NT, NCH, PAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    return;
  }
  const size_t gbytes = (size_t)2 * NCH * NP * LD * sizeof(double);
  const size_t slot = (size_t)(2 + NCH) * NP * LD * sizeof(double);   // 2 P buffers + NCH Y tiles per pixel slot
  // pixel slots per group: keep two CTAs per SM (<= ~110 KB each), enough items for 8 warps
  int PB = (int)std::max<size_t>(1, std::min<size_t>((110 * 1024 - gbytes) / slot, 64));
  PB = std::min(PB, std::max(1, 32 / (NT * NCH)) * 2);
  PB = std::max(PB, 1);
  const size_t smem = gbytes + PB * slot;
  // pixels per CTA: enough CTAs to fill the GPU, but amortise the table load
  int ppc = std::max(PB, (int)((totpix + 148 * 8 - 1) / (148 * 8)));
  ppc = std::min(ppc, std::max(PB * 8, 64));
  ppc = round_up(std::min(ppc, std::max(maxpix, 1)), PB);
  const dim3 grid((maxpix + ppc - 1) / ppc, ntasks);
  dev::k_fold<NT, NCH, PAR><<<grid, 256, smem, st>>>(bd, tasks, pixlist, G, Ppix, R, ppc, PB);
  CK(cudaGetLastError());
}

void launch_fold(int NT, int nch, bool par, const dev::BasisDev &bd, const dev::FoldTask *tasks, int ntasks,
                 const int *pixlist, int maxpix, int64_t totpix, const double *G, const double *Ppix, double *R,
                 cudaStream_t st) {
#define HFQ_FOLD_CASE(nt, parity)                                                                        \
  if (NT == nt && par == parity) {                                                                       \
    if (nch == 1)                                                                                        \
      launch_fold_t<nt, 1, parity>(bd, tasks, ntasks, pixlist, maxpix, totpix, G, Ppix, R, st);          \
    else                                                                                                 \
      launch_fold_t<nt, 2, parity>(bd, tasks, ntasks, pixlist, maxpix, totpix, G, Ppix, R, st);          \
    return;                                                                                              \
  }
  HFQ_FOLD_CASE(1, false)
  HFQ_FOLD_CASE(2, false)
  HFQ_FOLD_CASE(3, false)
  HFQ_FOLD_CASE(4, false)
  HFQ_FOLD_CASE(6, false)
  HFQ_FOLD_CASE(8, false)
  HFQ_FOLD_CASE(2, true)
  HFQ_FOLD_CASE(4, true)
  HFQ_FOLD_CASE(6, true)
  HFQ_FOLD_CASE(8, true)
  throw std::runtime_error("fold: unsupported sector size");
#undef HFQ_FOLD_CASE
}

template <bool KC>
void launch_gemm(const dev::GemmItem *items, const dev::GemmEntry *entries, int nitems, int maxM, int maxN,
                 cudaStream_t st) {
  if (nitems == 0) return;
  constexpr int BM = 64, BN = 64;
  const dim3 grid((maxN + BN - 1) / BN, (maxM + BM - 1) / BM, nitems);
  dev::k_gemm<BM, BN, 2, 2, KC><<<grid, 128, 0, st>>>(items, entries);
  CK(cudaGetLastError());
}

}  // namespace

// ---------------------------------------------------------------------------
// exchange
// ---------------------------------------------------------------------------
// Block norms of P (the reference's screening quantity) and the sector-packed copy of every
// sector pair that carries density.  Shared by coulomb and exchange inside a fused build.
// Only the per-sector-pair flags (a few KB) come back to the host; the per-block norms follow only when a
// host-pointer call needs them (row ranges of the next speculative upload).
void Engine::pack_density(const double *dP, int64_t ldP, cudaStream_t st) {
  Impl &s = *p_;
  const BasisTables &t = s.t;
  const int na = t.Nang(), ns = s.ns;
  const size_t nn = (size_t)na * na;
  const size_t nflag = (size_t)ns * ns;
  if (s.norms_host.size() != nn) {
    // page-locked once: the per-call read-back is then a plain DMA instead of a staged pageable copy
    if (!s.norms_host.empty()) cudaHostUnregister(s.norms_host.data());
    s.norms_host.assign(nn, 0.0);
    if (cudaHostRegister(s.norms_host.data(), s.norms_host.size() * sizeof(double), cudaHostRegisterDefault) != cudaSuccess)
      cudaGetLastError();   // stays pageable: slower, still correct
    if (!s.flags_host) CK(cudaMallocHost(&s.flags_host, (nflag + 2) * sizeof(unsigned long long)));
    s.d_spflags.alloc(nflag + 2, &dev_bytes_);
  }
  CK(cudaMemsetAsync(s.d_spflags.p, 0, (nflag + 2) * sizeof(unsigned long long), st));
  // multi-GPU: every rank scans 1/nranks of the columns, one all-reduce(max) of the flags completes them (the
  // per-block norms themselves are only needed by the host-pointer calls: scanned in full there)
  const bool split = comm_ && !s.want_norms_host;
  const int c0 = split ? comm_->rank() : 0, cs = split ? comm_->size() : 1;
  if (c0 < na) {
    dev::k_block_norms<<<dim3((na - c0 + cs - 1) / cs, (na + 3) / 4), 128, 0, st>>>(s.bd, dP, ldP, s.d_ang_sec.p, s.d_norms.p,
                                                                                   s.d_spflags.p, c0, cs);
    CK(cudaGetLastError());
  }
  if (split) comm_->all_reduce_max_u64(s.d_spflags.p, nflag + 2, st);
  CK(cudaMemcpyAsync(s.flags_host, s.d_spflags.p, (nflag + 2) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  if (s.want_norms_host)
    CK(cudaMemcpyAsync(s.norms_host.data(), s.d_norms.p, nn * sizeof(double), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  {
    // symmetric density (every SCF density is): the exchange builds half of each diagonal output pair
    double asym, amax;
    std::memcpy(&asym, &s.flags_host[nflag], sizeof(double));
    std::memcpy(&amax, &s.flags_host[nflag + 1], sizeof(double));
    static const bool allow = !(getenv("HFQ_NO_SYMMETRY") && atoi(getenv("HFQ_NO_SYMMETRY")));
    s.p_symmetric = allow && asym <= 1e-14 * amax;
  }
  s.packed_splist.clear();
  for (int sp = 0; sp < ns * ns; sp++)
    if (s.flags_host[sp] & 1) s.packed_splist.push_back(sp);
  if (!s.packed_splist.empty()) {
    CK(cudaMemcpyAsync(s.d_splist.p, s.packed_splist.data(), s.packed_splist.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    for (int sp : s.packed_splist)
      CK(cudaMemsetAsync(s.d_Ppix.p + (size_t)sp * s.Npix * s.NB, 0, (size_t)s.Npix * s.NB * sizeof(double), st));
    dev::k_pack<<<dim3(t.Nrad, (unsigned)s.packed_splist.size()), 256, 0, st>>>(s.bd, dP, ldP, s.d_splist.p, s.d_Ppix.p);
    CK(cudaGetLastError());
  }
}

// Fused Fock-build step: J = coulomb(P), K = exchange(kscale * P) from ONE packed copy of P.
void Engine::jk_dev(const double *dP, int64_t ldP, double kscale, double *dJ, int64_t ldJ, double *dK, int64_t ldK,
                    int shard, int nshards, cudaStream_t st) {
  Impl &s = *p_;
  CK(cudaSetDevice(device_));
  static const bool trace = getenv("HFQ_TRACE") && atoi(getenv("HFQ_TRACE"));
  const auto t0 = std::chrono::steady_clock::now();
  auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  pack_density(dP, ldP, st);
  const double t_pack = ms_since();
  s.packed_valid = true;
  if (!s.j_stream) {
    CK(cudaStreamCreateWithPriority(&s.j_stream, cudaStreamNonBlocking, s.prio_hi));
    CK(cudaEventCreateWithFlags(&s.ev_packed, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_jdone, cudaEventDisableTiming));
  }
  try {
    // the Coulomb chain (~1 ms of small kernels) runs on its own stream and fills the tails of the
    // exchange kernels instead of preceding them.  With a communicator every rank builds the complete J
    // (replicated: cheaper than a second collective); without one the legacy shard arguments split it over L.
    CK(cudaEventRecord(s.ev_packed, st));
    CK(cudaStreamWaitEvent(s.j_stream, s.ev_packed, 0));
    if (comm_)
      coulomb_run(dP, ldP, dJ, ldJ, 0, 1, s.j_stream, true);
    else
      coulomb_run(dP, ldP, dJ, ldJ, shard, nshards, s.j_stream, true);
    CK(cudaEventRecord(s.ev_jdone, s.j_stream));
    EngineTimings tj;
    tj.launches = 6;
    const double t_j = ms_since();
    s.kscale = kscale;
    exchange_dev(dP, ldP, dK, ldK, shard, nshards, st);
    CK(cudaStreamWaitEvent(st, s.ev_jdone, 0));   // work queued on st by the caller sees J complete
    CK(cudaStreamSynchronize(s.j_stream));
    if (trace)
      fprintf(stderr, "[hfq] jk_dev: pack %.2f  J %.2f (gpu %.2f)  K %.2f (gpu %.2f: fold %.2f gemm %.2f off %.2f unpack %.2f) ms\n",
              t_pack, t_j - t_pack, tj.total, ms_since() - t_j, tm_.total, tm_.fold, tm_.tgemm, tm_.offdiag, tm_.unpack);
    tm_.launches += tj.launches;
    tm_.total += tj.total;
  } catch (...) {
    s.packed_valid = false;
    s.kscale = 1.0;
    cudaStreamSynchronize(s.j_stream);
    throw;
  }
  s.packed_valid = false;
  s.kscale = 1.0;
}

// A plan is everything about an exchange call that depends only on WHICH sector pairs of the
// density are non-zero (plus sharding and the +-m flag): the task list, its split into batches
// and the device-resident kernel descriptors.  SCF iterations reuse it.
//
// Sharding (owner computes).  The unit of ownership is (output sector pair, element pair): its block of the
// accumulator is written by exactly one rank, which folds only the pixels of that element pair, runs the
// in-element GEMM (ei == ej) or the cross-element kernel (ei != ej) over all tasks of the output pair, and
// reduces its K-split partials.  Units are dealt to the ranks longest-first onto the least loaded rank (costs =
// executed flops of the three kernels); the unit blocks of a rank are contiguous in the compact buffer Kc, every
// rank's segment has the same length, so ONE in-place ncclAllGather completes the exchange matrix on all
// ranks; the unpack (dense K, boundary removal, mirrors) then runs everywhere.
struct ExchangeBatch {
  DevBuf<dev::FoldTask> tasks;
  DevBuf<dev::GemmItem> gitems;     // in-element items: first the ngitems_tri items of the symmetric-density layout
  DevBuf<dev::GemmEntry> gentries;  // (A tiles of TP_BK_TRI columns), then the others (TP_BK) -- one launch each
  DevBuf<dev::OffItem> oitems;
  DevBuf<dev::OffEntry> oentries;
  int ntasks = 0, ngitems = 0, ngitems_tri = 0, noitems = 0;
  int maxM_tri = 8, maxM_full = 8;
  int maxpix = 0;        // longest pixel list of a task
  int64_t totpix = 0;    // pixels folded by the batch
};
struct ExchangePlan {
  std::string key;
  std::vector<std::unique_ptr<ExchangeBatch>> batches;
  std::vector<int> splist, op_src, op_tri, op_ldk;
  DevBuf<int> d_op_src, d_op_tri, d_op_ldk, d_blocks;   // device copies + angular blocks (j | k << 16) that can be non-zero
  DevBuf<int64_t> d_unit_off;                 // [(active op * Nel + ei) * Nel + ej] offset in Kc or -1
  DevBuf<dev::ReduceDesc> d_reduce;
  DevBuf<int> d_pixlist;                      // pixel lists of the fold tasks (one per distinct set of owned element pairs)
  int nreduce = 0;
  int64_t seg = 0;                            // doubles per rank segment of Kc
  int64_t part_stride = 0;                    // doubles per K-split partial (partials 1 .. S-1 of the own units)
  int nblocks = 0;
  int nactive = 0, S = 1, maxM = 8;
  double fl_fold = 0, fl_tg = 0, fl_off = 0, al_fold = 0, al_tg = 0, al_off = 0;
  const double *R_base = nullptr, *K_base = nullptr, *Kc_base = nullptr;   // buffers the descriptors point into
};

struct Engine::PlanCache {
  std::vector<std::unique_ptr<ExchangePlan>> plans;
};

int Engine::comm_size() const { return comm_ ? comm_->size() : 1; }

void Engine::set_comm(const void *id128, int rank, int nranks) {
  CK(cudaSetDevice(device_));
  comm_.reset();
  if (nranks > 1) comm_ = std::make_unique<Comm>(id128, rank, nranks, device_);
  if (plans_) plans_->plans.clear();
  p_->jplan.valid = false;
}

// Longest-processing-time assignment of n units to nranks; deterministic (ties by index), identical on every rank.
void assign_units(const std::vector<double> &cost, int nranks, std::vector<int> &owner) {
  const int n = (int)cost.size();
  owner.assign(n, 0);
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
  std::vector<double> load(nranks, 0.0);
  for (int u : order) {
    int best = 0;
    for (int r = 1; r < nranks; r++)
      if (load[r] < load[best]) best = r;
    owner[u] = best;
    load[best] += cost[u];
  }
}

// Radial Coulomb operator of MANY densities in one launch (batched atoms of the SAP workload):
// J_b = fac * J_0(P_b), J_0 the L = 0 multipole of assemble_J_FE_one_multipole_cached_chol
// (libhelfemqc/include/CoulombExchangeFE.h:432-482; sadatom coulomb = 4 pi J_0, src/sadatom/basis.cpp:186-207).
// dP, dJ: nb matrices of Nrad x Nrad, contiguous, stride `stride` doubles (>= Nrad^2); symmetric densities.
void Engine::coulomb_radial_batch(const double *dP, double *dJ, int nb, int64_t stride, double fac, cudaStream_t st) {
  Impl &s = *p_;
  const BasisTables &t = s.t;
  if (t.kind == BasisKind::Diatomic || t.nch != 1 || t.pairwise()) throw std::logic_error("coulomb_radial_batch: atomic radial caches required");
  if (stride != (int64_t)s.Npix) throw std::logic_error("coulomb_radial_batch: matrices must be contiguous (stride = Nrad^2)");
  if (nb < 1) return;
  CK(cudaSetDevice(device_));
  const int ilm = t.channel(0, 0);
  if (ilm < 0) throw std::logic_error("coulomb_radial_batch: no L = 0 channel");
  if ((int)s.d_bchan.n < nb) {
    s.d_bchan.upload(std::vector<int>(nb, ilm), &dev_bytes_);
    s.d_bfac.upload(std::vector<double>(nb, 0.0), &dev_bytes_);
  }
  std::vector<double> hf(nb, fac * t.pref[ilm]);
  CK(cudaMemcpyAsync(s.d_bfac.p, hf.data(), nb * sizeof(double), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));   // hf goes out of scope
  dev::BasisDev b1 = s.bd;
  b1.NL = 1;   // one multipole: Paux[batch][pix], JauxT[batch][pix]
  dev::JRadDev jr{s.d_bchan.p, s.d_bfac.p, s.d_blk_off.p, s.d_B_off.p, s.d_sig_off.p, s.d_rank.p,
                  s.d_small.p, s.d_big.p, s.d_B.p, s.d_sigma.p, nb};
  dev::k_jradial<<<dim3(1, nb), 256, 0, st>>>(b1, jr, 0, dP, dJ);
  CK(cudaGetLastError());
}

void Engine::exchange_dev(const double *dP, int64_t ldP, double *dK, int64_t ldK, int shard, int nshards,
                          cudaStream_t st) {
  Impl &s = *p_;
  const BasisTables &t = s.t;
  if (s.radial_only) throw std::logic_error("exchange is not available on batch tables");
  CK(cudaSetDevice(device_));
  const int na = t.Nang(), ns = s.ns, Nel = t.Nel;
  if (comm_) {   // a communicator overrides the legacy shard arguments
    shard = comm_->rank();
    nshards = comm_->size();
  }
  tm_ = EngineTimings();
  if (!plans_) plans_.reset(new PlanCache);
  CK(cudaEventRecord(s.ev[0], st));
  // everything outside the computed sector pairs is exactly zero: K is cleared at memset speed on a second stream
  // while the (tensor-pipe-bound) exchange kernels run; the unpack then writes the blocks that can be non-zero
  if (!s.aux_stream) {
    CK(cudaStreamCreateWithPriority(&s.aux_stream, cudaStreamNonBlocking, s.prio_hi));
    CK(cudaEventCreate(&s.ev_start));
    CK(cudaEventCreate(&s.ev_kzero));
  }
  CK(cudaEventRecord(s.ev_start, st));   // earlier work on st that still reads K must finish first
  CK(cudaStreamWaitEvent(s.aux_stream, s.ev_start, 0));
  CK(cudaMemset2DAsync(dK, (size_t)ldK * sizeof(double), 0, (size_t)nbf_ * sizeof(double), (size_t)nbf_, s.aux_stream));
  CK(cudaEventRecord(s.ev_kzero, s.aux_stream));
  // 1. which sector pairs of P carry density (reference: block norm >= 10 eps)
  if (!s.packed_valid) pack_density(dP, ldP, st);
  std::string key((size_t)ns * ns + 3, '0');
  for (int sp = 0; sp < ns * ns; sp++)
    if (s.flags_host[sp] & 2) key[sp] = '1';
  key[(size_t)ns * ns] = (char)((absm_symmetric_ ? 'S' : 'N') + (s.p_symmetric ? 1 : 0));
  key[(size_t)ns * ns + 1] = (char)('0' + shard);
  key[(size_t)ns * ns + 2] = (char)('0' + nshards);
  ExchangePlan *plan = nullptr;
  for (auto &pl : plans_->plans)
    if (pl->key == key && pl->R_base == s.d_R.p && pl->K_base == s.d_Kacc.p && pl->Kc_base == s.d_Kc.p) plan = pl.get();
  if (!plan) {
    // ---------------- build the plan ----------------
    CK(cudaStreamSynchronize(st));   // buffers may be re-allocated below
    auto np = std::make_unique<ExchangePlan>();
    np->key = key;
    std::vector<char> sp_nz((size_t)ns * ns, 0);
    for (int sp = 0; sp < ns * ns; sp++) {
      sp_nz[sp] = key[sp] == '1';
      if (sp_nz[sp]) np->splist.push_back(sp);
    }
    struct OpWork {
      int op;
      bool tri = false;   // diagonal output pair of a symmetric density: half storage
      std::vector<dev::FoldTask> tasks;
      std::vector<int> ilm;
      std::vector<double> alg_fold;  // unpadded flops of the fold per task (all pixels)
      std::vector<int> own_pairs;       // ei * Nel + ej of the element pairs built by this rank
      int pix0 = 0, npix = 0;           // its pixel list (pixels of those element pairs)
    };
    std::vector<OpWork> work;   // every active output pair (identical on all ranks)
    np->op_src.assign((size_t)ns * ns, -1);
    // Task list: output pair (sj,sk) <- density pair (si,sl) with mj-mi == mk-ml, every coupled L.
    for (int sj = 0; sj < ns; sj++)
      for (int sk = 0; sk < ns; sk++) {
        const int mj = s.sec_m[sj], mk = s.sec_m[sk];
        if (absm_symmetric_ && (mj < 0 || mk < 0)) continue;
        if (t.kind == BasisKind::Sadatom && sj != sk) continue;   // l-diagonal outputs only
        OpWork w;
        w.op = sj * ns + sk;
        w.tri = s.p_symmetric && sj == sk;
        for (int si = 0; si < ns; si++)
          for (int sl = 0; sl < ns; sl++) {
            if (!sp_nz[(size_t)si * ns + sl]) continue;
            const int mi = s.sec_m[si], ml = s.sec_m[sl];
            if (mj - mi != mk - ml) continue;
            const int M = mj - mi;
            for (int L = std::abs(M); L < s.NL; L++) {
              const int ilm = t.channel(L, std::abs(M));
              if (ilm < 0) continue;
              if (!s.G_nonzero[((size_t)sj * ns + si) * s.NL + L] || !s.G_nonzero[((size_t)sk * ns + sl) * s.NL + L])
                continue;
              dev::FoldTask ft;
              ft.spj = sj * ns + si;
              ft.spk = sk * ns + sl;
              ft.spp = si * ns + sl;
              ft.L = L;
              ft.rslot = 0;
              ft.ldk = round_up(s.sec_span[sk], 2);
              ft.pix0 = ft.npix = 0;
              ft.fac = t.pref[ilm] * ((t.sign_by_M && (M & 1)) ? -1.0 : 1.0);
              w.tasks.push_back(ft);
              w.ilm.push_back(ilm);
              const double nj = s.sec_n[sj], nk = s.sec_n[sk], ni = s.sec_n[si], nl = s.sec_n[sl];
              w.alg_fold.push_back(2.0 * (ni * nl * nk * t.nch + nj * ni * nk * s.nab));
            }
          }
        if (!w.tasks.empty()) work.push_back(std::move(w));
      }
    const size_t slot_doubles = (size_t)s.nab * s.Npix * s.NB;
    const int nactive = (int)work.size();
    np->nactive = nactive;
    for (int a = 0; a < nactive; a++) np->op_src[work[a].op] = a;
    np->op_tri.assign((size_t)ns * ns, 0);
    for (int a = 0; a < nactive; a++) np->op_tri[a] = work[a].tri ? 1 : 0;
    np->op_ldk.assign((size_t)ns * ns, s.NP);
    for (int a = 0; a < nactive; a++) np->op_ldk[a] = round_up(s.sec_span[work[a].op % ns], 2);
    // ---- units = (active output pair, element pair); ownership
    struct Unit {
      int a, ei, ej, rows, ncol;
      double cost;
      int owner = 0;
      int64_t off = -1;   // offset in Kc
    };
    std::vector<Unit> units;
    const bool by_pair = nshards > 1 && Nel <= 8;   // else whole output pairs are dealt (regmask has 64 bits)
    for (int a = 0; a < nactive; a++) {
      const OpWork &w = work[a];
      const int ncol = s.sec_span[w.op / ns] * round_up(s.sec_span[w.op % ns], 2);
      const double ntask = (double)w.tasks.size();
      for (int ei = 0; ei < Nel; ei++)
        for (int ej = 0; ej < Nel; ej++) {
          if (w.tri && ei > ej) continue;
          const int Ni = t.en[ei], Nj = t.en[ej];
          Unit u;
          u.a = a;
          u.ei = ei;
          u.ej = ej;
          u.rows = (w.tri && ei == ej) ? Ni * (Ni + 1) / 2 : Ni * Nj;
          u.ncol = ncol;
          const double fold = 2.0 * Ni * Nj * (double)s.NP * s.NP * s.NP * (t.nch + s.nab) * (s.parity ? 0.5 : 1.0);
          double body;
          if (ei == ej || t.pairwise())
            body = 2.0 * u.rows * (double)ncol * s.nab * Ni * Nj;
          else   // the cross-element kernel runs at about half the flop rate of the GEMM
            body = 2.0 * (2.0 * t.nch * ((double)Ni * Nj * t.nch * Nj + (double)Ni * Nj * Ni) * ncol);
          u.cost = ntask * (fold + body);
          units.push_back(u);
        }
    }
    {
      std::vector<int> owner;
      if (by_pair) {
        std::vector<double> cost;
        for (auto &u : units) cost.push_back(u.cost);
        assign_units(cost, nshards, owner);
      } else {
        std::vector<double> cost(nactive, 0.0);
        for (auto &u : units) cost[u.a] += u.cost;
        std::vector<int> oo;
        assign_units(cost, nshards, oo);
        for (auto &u : units) owner.push_back(oo[u.a]);
      }
      for (size_t i = 0; i < units.size(); i++) units[i].owner = owner[i];
    }
    // compact buffer Kc: [rank segment][unit][rows][NB], all segments of equal length
    {
      std::vector<int64_t> fill(nshards, 0);
      for (auto &u : units) {
        u.off = fill[u.owner];
        fill[u.owner] += (int64_t)u.rows * s.NB;
      }
      np->seg = *std::max_element(fill.begin(), fill.end());
      np->seg = (np->seg + 63) / 64 * 64;
      for (auto &u : units) u.off += (int64_t)u.owner * np->seg;
      const size_t need = (size_t)np->seg * nshards;
      if (s.d_Kc.n < need) {
        s.d_Kc.alloc(need, &dev_bytes_);
        CK(cudaMemsetAsync(s.d_Kc.p, 0, need * sizeof(double), st));   // padding travels through the collective: keep it finite
      }
      std::vector<int64_t> uoff((size_t)std::max(nactive, 1) * Nel * Nel, -1);
      for (auto &u : units)
        if (comm_ || u.owner == shard) uoff[((size_t)u.a * Nel + u.ei) * Nel + u.ej] = u.off;
      np->d_unit_off.upload(uoff, &dev_bytes_);
    }
    // own units per output pair
    int own_gemm_ctas = 0, own_off_ctas = 0;
    for (auto &u : units)
      if (u.owner == shard) {
        OpWork &w = work[u.a];
        w.own_pairs.push_back(u.ei * Nel + u.ej);
        if (u.ei == u.ej || t.pairwise())
          own_gemm_ctas += (u.ncol + 63) / 64;
        else
          own_off_ctas += (u.ncol + 15) / 16;
      }
    {
      // pixel lists: the pixels (ri, rl) of the owned element pairs of an output pair (elements overlap in their
      // boundary functions: each pixel once), one list per distinct set
      std::map<std::vector<int>, std::pair<int, int>> lists;
      std::vector<int> pixlist;
      for (auto &w : work) {
        if (w.own_pairs.empty()) continue;
        auto it = lists.find(w.own_pairs);
        if (it == lists.end()) {
          std::vector<char> need((size_t)s.Npix, 0);
          for (int pr : w.own_pairs) {
            const int ei = pr / Nel, ej = pr % Nel;
            for (int ri = t.efirst[ei]; ri < t.efirst[ei] + t.en[ei]; ri++)
              for (int rl = t.efirst[ej]; rl < t.efirst[ej] + t.en[ej]; rl++) need[(size_t)ri * t.Nrad + rl] = 1;
          }
          const int p0 = (int)pixlist.size();
          for (int pix = 0; pix < s.Npix; pix++)
            if (need[pix]) pixlist.push_back(pix);
          it = lists.emplace(w.own_pairs, std::make_pair(p0, (int)pixlist.size() - p0)).first;
        }
        w.pix0 = it->second.first;
        w.npix = it->second.second;
      }
      np->d_pixlist.upload(pixlist, &dev_bytes_);
    }
    size_t total_tasks = 0;
    for (auto &w : work)
      if (!w.own_pairs.empty()) total_tasks += w.tasks.size();
    // K-split of the in-element GEMM: each (unit, 64-column tile) is cut into S chunks of its task list, each
    // chunk accumulating into its own partial buffer (partial 0 = Kc itself; the others are summed into it by
    // k_reduce_partials), so that the CTAs of a launch fill whole waves of the 148 SMs.
    // The CTAs of a launch differ in length (task counts, partial column tiles), so among splits that fill the waves
    // about equally well the finer one balances better: a small bonus per chunk breaks the ties towards it.
    int S = 1;
    if (own_gemm_ctas > 0) {
      double best = 0.0;
      for (int c = 1; c <= 8; c++) {
        const double u = (double)own_gemm_ctas * c, eff = u / (std::ceil(u / 148.0) * 148.0) + 0.004 * c;
        if (eff > best) {
          best = eff;
          S = c;
        }
      }
    }
    np->S = S;
    // chunks of the cross-element items: enough CTAs for ~8 waves of the 2 x 148 resident ones, at most S (the
    // partial buffers are shared with the in-element items)
    int S_off = 1;
    if (own_off_ctas > 0) S_off = std::max(1, std::min(S, (8 * 296 + own_off_ctas - 1) / own_off_ctas));
    // partial buffers 1 .. S-1: same layout as this rank's segment of Kc
    np->part_stride = np->seg;
    if (S > 1 && s.d_Kacc.n < (size_t)(S - 1) * np->seg) s.d_Kacc.alloc((size_t)(S - 1) * np->seg, &dev_bytes_);
    if (S > 1) {
      std::vector<dev::ReduceDesc> rd;
      for (auto &u : units)
        if (u.owner == shard) {
          const int64_t n = (int64_t)u.rows * s.NB;
          const int nparts = ((u.ei == u.ej || t.pairwise()) ? S : S_off) - 1;
          if (nparts < 1) continue;
          for (int64_t o = 0; o < n; o += 16384)
            rd.push_back(dev::ReduceDesc{u.off + o, u.off - (int64_t)shard * np->seg + o, (int)std::min<int64_t>(16384, n - o), nparts});
        }
      np->nreduce = (int)rd.size();
      np->d_reduce.upload(rd, &dev_bytes_);
    }
    if (absm_symmetric_) {
      // K(-mj,-mk) block = K(mj,mk) block (sectors +-m of one parity class hold the same l list)
      std::map<std::pair<int, int>, int> sec_of_m;
      for (int i = 0; i < ns; i++) sec_of_m[{s.sec_m[i], s.sec_cls[i]}] = i;
      for (int sj = 0; sj < ns; sj++)
        for (int sk = 0; sk < ns; sk++) {
          const int mj = s.sec_m[sj], mk = s.sec_m[sk];
          if (mj >= 0 || mk >= 0) continue;
          auto pj = sec_of_m.find({-mj, s.sec_cls[sj]}), pk = sec_of_m.find({-mk, s.sec_cls[sk]});
          if (pj == sec_of_m.end() || pk == sec_of_m.end()) continue;
          np->op_src[(size_t)sj * ns + sk] = np->op_src[(size_t)pj->second * ns + pk->second];
        }
    }
    {
      if (na > 0xffff) throw std::runtime_error("Engine: more than 65535 angular functions");
      std::vector<int> blocks;
      for (int aj = 0; aj < na; aj++)
        for (int ak = 0; ak < na; ak++)
          if (np->op_src[(size_t)s.ang_sec[aj] * ns + s.ang_sec[ak]] >= 0) blocks.push_back(aj | (ak << 16));
      np->nblocks = (int)blocks.size();
      np->d_blocks.upload(blocks, &dev_bytes_);
      np->d_op_src.upload(np->op_src, &dev_bytes_);
      np->d_op_tri.upload(np->op_tri, &dev_bytes_);
      np->d_op_ldk.upload(np->op_ldk, &dev_bytes_);
    }
    {
      size_t free_b = 0, total_b = 0;
      CK(cudaMemGetInfo(&free_b, &total_b));
      const size_t have = s.d_R.n * sizeof(double);
      size_t budget = std::min<size_t>((size_t)48 << 30, (size_t)((free_b + have) * 0.6));
      if (const char *env = getenv("HFQ_R_BUDGET_MB"))   // tests: force the multi-batch path
        budget = std::min<size_t>(budget, (size_t)std::max(1, atoi(env)) << 20);
      size_t want_slots = std::min<size_t>(total_tasks, std::max<size_t>(1, budget / (slot_doubles * sizeof(double))));
      if (want_slots > s.r_slots) {
        s.d_R.alloc(want_slots * slot_doubles, &dev_bytes_);
        s.r_slots = want_slots;
      }
    }
    if (total_tasks && s.r_slots == 0) throw std::runtime_error("Engine: no memory for the exchange work buffer");
    np->R_base = s.d_R.p;
    np->K_base = s.d_Kacc.p;
    np->Kc_base = s.d_Kc.p;
    std::vector<std::vector<int64_t>> own_off(work.size());   // parallel to OpWork::own_pairs (both follow the unit order)
    for (auto &u : units)
      if (u.owner == shard) own_off[u.a].push_back(u.off);
    // Batches: every batch takes an equal share of tasks from every output pair this rank works on, so each
    // launch works on all of them at once (grid size independent of the batch count).
    std::vector<size_t> done(work.size(), 0);
    std::map<std::pair<size_t, int>, char> started;   // (work index * 4096 + pair, partial) -> written before
    size_t remaining = total_tasks;
    while (remaining > 0) {
      std::vector<dev::FoldTask> tasks;
      std::vector<dev::GemmItem> gitems;
      std::vector<dev::GemmEntry> gentries;
      std::vector<dev::OffItem> oitems;
      std::vector<dev::OffEntry> oentries;
      int batch_maxpix = 0, batch_maxM_tri = 8, batch_maxM_full = 8;
      int64_t batch_totpix = 0;
      size_t open = 0;
      for (size_t wi = 0; wi < work.size(); wi++) open += !work[wi].own_pairs.empty() && done[wi] < work[wi].tasks.size();
      const size_t share = std::max<size_t>(1, s.r_slots / std::max<size_t>(open, 1));
      for (size_t wi = 0; wi < work.size() && tasks.size() < s.r_slots; wi++) {
        OpWork &w = work[wi];
        if (w.own_pairs.empty()) continue;
        const size_t ti = done[wi];
        size_t take = std::min(w.tasks.size() - ti, std::min(share, s.r_slots - tasks.size()));
        if (remaining <= s.r_slots) take = w.tasks.size() - ti;   // everything fits: finish
        take = std::min(take, s.r_slots - tasks.size());
        if (take == 0) continue;
        const size_t t0 = tasks.size();
        for (size_t k = 0; k < take; k++) {
          dev::FoldTask ft = w.tasks[ti + k];
          ft.rslot = (int)(t0 + k);
          ft.pix0 = w.pix0;
          ft.npix = w.npix;
          tasks.push_back(ft);
        }
        const double own_pix = w.npix;   // pixels this rank folds per task
        batch_maxpix = std::max(batch_maxpix, w.npix);
        batch_totpix += (int64_t)take * w.npix;
        for (size_t pi = 0; pi < w.own_pairs.size(); pi++) {
          const int ei = w.own_pairs[pi] / Nel, ej = w.own_pairs[pi] % Nel;
          const int Ni = t.en[ei], Nj = t.en[ej];
          const int64_t koff = own_off[wi][pi];
          // dense columns (pos_j, pos_k) = pos_j * ldk + pos_k, ldk = even number of positions of the column sector
          const int ncol = s.sec_span[w.op / ns] * round_up(s.sec_span[w.op % ns], 2);
          if (ei == ej || t.pairwise()) {
            // in-element item (tensor-core GEMM against the dense exchange-ordered kernel; for pair-tensor
            // tables (erfc) every element pair), S chunks; chunk 0 always gets work
            const bool half = w.tri && ei == ej;
            for (int c = 0; c < S; c++) {
              const size_t k0 = (take * c + S - 1) / S, k1 = (take * (c + 1) + S - 1) / S;
              if (k1 == k0) continue;
              double *C = c == 0 ? s.d_Kc.p + koff
                                 : s.d_Kacc.p + (size_t)(c - 1) * np->part_stride + (koff - (int64_t)shard * np->seg);
              char &st_flag = started[{wi * 4096 + (size_t)w.own_pairs[pi], c}];
              dev::GemmItem gi{};
              gi.C = C;
              gi.browoff = t.pairwise() ? s.d_browoff_P.p + s.browoff_P_first[(size_t)ei * Nel + ej]
                                        : s.d_browoff_T.p + s.browoff_T_first[ei];
              gi.M = half ? Ni * (Ni + 1) / 2 : Ni * Nj;
              gi.ldb = w.tri ? 1 : 0;   // (unused by the kernel) layout of the item's A tiles: 1 = TP_BK_TRI
              (w.tri ? batch_maxM_tri : batch_maxM_full) = std::max(w.tri ? batch_maxM_tri : batch_maxM_full, gi.M);
              gi.N = round_up(ncol, 8);
              gi.K = s.nab * Ni * Nj;
              gi.ent0 = (int)gentries.size();
              for (size_t k = k0; k < k1; k++) {
                const int ilm = w.ilm[ti + k];
                dev::GemmEntry ge;
                if (t.pairwise()) {
                  const size_t pidx = ((size_t)ilm * Nel + ei) * Nel + ej;
                  ge.A = w.tri ? s.d_tperm_tri.p + s.tpair_tri_off[pidx] : s.d_tperm.p + s.tpair_off[pidx];
                } else {
                  ge.A = w.tri ? s.d_tperm_tri.p + s.tperm_tri_off[(size_t)ilm * Nel + ei]
                               : s.d_tperm.p + s.tperm_off[(size_t)ilm * Nel + ei];
                }
                ge.lda = 0;
                ge.B = s.d_R.p + (t0 + k) * slot_doubles;
                gentries.push_back(ge);
              }
              gi.ent1 = (int)gentries.size();
              gi.accumulate = st_flag ? 1 : 0;
              st_flag = 1;
              gi.ldc = s.NB;
              gi.alpha = 1.0;
              gitems.push_back(gi);
              np->fl_tg += 2.0 * gi.M * gi.N * (double)gi.K * (k1 - k0);
              np->al_tg += 2.0 * gi.M * (double)s.sec_n[w.op / ns] * s.sec_n[w.op % ns] * (double)gi.K * (k1 - k0);
            }
          } else {
            // cross-element items (rank-1 factors): the task list is cut into S_off chunks with their own partial
            // accumulators (chunk 0 = Kc itself), like the in-element items, so that a rank with few units still
            // fills the GPU (8 GPUs: ~10 units x 13 column tiles per rank against 296 resident CTAs)
            for (int c = 0; c < S_off; c++) {
              const size_t k0 = (take * c + S_off - 1) / S_off, k1 = (take * (c + 1) + S_off - 1) / S_off;
              if (k1 == k0) continue;
              char &st_flag = started[{wi * 4096 + (size_t)w.own_pairs[pi], c}];
              dev::OffItem oi{};
              oi.C = c == 0 ? s.d_Kc.p + koff
                            : s.d_Kacc.p + (size_t)(c - 1) * np->part_stride + (koff - (int64_t)shard * np->seg);
              oi.ei = ei;
              oi.ej = ej;
              oi.ent0 = (int)oentries.size();
              for (size_t k = k0; k < k1; k++) oentries.push_back(dev::OffEntry{(int)(t0 + k), w.ilm[ti + k]});
              oi.ent1 = (int)oentries.size();
              oi.accumulate = st_flag ? 1 : 0;
              st_flag = 1;
              oi.ncol = ncol;
              oitems.push_back(oi);
            }
            const double per = 2.0 * t.nch * take * ((double)Ni * Nj * t.nch * Nj + (double)Ni * Nj * Ni);
            np->fl_off += per * round_up(ncol, 16);
            np->al_off += per * s.sec_n[w.op / ns] * s.sec_n[w.op % ns];
          }
        }
        np->fl_fold += 2.0 * (double)take * own_pix * (double)s.NP * s.NP * s.NP * (t.nch + s.nab) * (s.parity ? 0.5 : 1.0);
        for (size_t k = 0; k < take; k++) np->al_fold += w.alg_fold[ti + k] * own_pix;
        done[wi] += take;
        remaining -= take;
      }
      if (tasks.empty()) break;
      // longest items first: CTAs are dispatched in block-index order, so the tail of the launch is made of
      // the cheapest items (accumulating items of one unit never share a launch, order is free)
      std::stable_sort(gitems.begin(), gitems.end(), [](const dev::GemmItem &a, const dev::GemmItem &b) {
        if (a.ldb != b.ldb) return a.ldb > b.ldb;   // the two A layouts run as two launches
        return (double)a.M * a.N * a.K * (a.ent1 - a.ent0) > (double)b.M * b.N * b.K * (b.ent1 - b.ent0);
      });
      int ntri = 0;
      for (auto &gi : gitems) ntri += gi.ldb ? 1 : 0;
      auto bt = std::make_unique<ExchangeBatch>();
      bt->tasks.upload(tasks, &dev_bytes_);
      bt->gitems.upload(gitems, &dev_bytes_);
      bt->gentries.upload(gentries, &dev_bytes_);
      bt->oitems.upload(oitems, &dev_bytes_);
      bt->oentries.upload(oentries, &dev_bytes_);
      bt->ntasks = (int)tasks.size();
      bt->ngitems = (int)gitems.size();
      bt->ngitems_tri = ntri;
      bt->maxM_tri = batch_maxM_tri;
      bt->maxM_full = batch_maxM_full;
      bt->noitems = (int)oitems.size();
      bt->maxpix = batch_maxpix;
      bt->totpix = batch_totpix;
      np->batches.push_back(std::move(bt));
    }
    if (plans_->plans.size() >= 4) plans_->plans.erase(plans_->plans.begin());
    plans_->plans.push_back(std::move(np));
    plan = plans_->plans.back().get();
  }
  // which dense blocks of K can be non-zero (for compact collectives and host copies): the pattern of the
  // COMPLETE matrix, identical on every rank
  last_active_ops_.assign(plan->op_src.begin(), plan->op_src.end());
  if (plan_hook_) plan_hook_();
  // ---------------- run the plan ----------------
  const int S = plan->S;
  if (S > 1) CK(cudaMemsetAsync(s.d_Kacc.p, 0, (size_t)(S - 1) * plan->part_stride * sizeof(double), st));
  CK(cudaEventRecord(s.ev[1], st));
  float ms_fold = 0, ms_tg = 0, ms_off = 0;
  for (auto &btp : plan->batches) {
    ExchangeBatch &bt = *btp;
    CK(cudaEventRecord(s.ev[2], st));
    launch_fold(s.NT, t.nch, s.parity, s.bd, bt.tasks.p, bt.ntasks, plan->d_pixlist.p, bt.maxpix, bt.totpix, s.d_G.p,
                s.d_Ppix.p, s.d_R.p, st);
    CK(cudaEventRecord(s.ev[3], st));
    if (bt.ngitems) {
      // in-element exchange: one CTA tile covers all Ni^2 rows (R rows are read once)
      // stages of the shared-memory ring: as many as fit in 227 KB for the largest A tile of the launch, at most 4
      auto stages_for = [](int maxM, int bk) {
        int stages = 4;
        while (stages > 2 && dev::tgemm_ws_smem(maxM, stages, bk) > 227 * 1024) stages--;
        return stages;
      };
      if (bt.ngitems_tri) {
        const int ns_ = stages_for(bt.maxM_tri, dev::TP_BK_TRI);
        dev::k_tgemm_ws<dev::TP_BK_TRI><<<dim3(s.NB / 64, (unsigned)bt.ngitems_tri), 288,
                                         dev::tgemm_ws_smem(bt.maxM_tri, ns_, dev::TP_BK_TRI), st>>>(
            bt.gitems.p, bt.gentries.p, s.d_zrow.p, ns_, bt.maxM_tri);
        CK(cudaGetLastError());
      }
      if (bt.ngitems > bt.ngitems_tri) {
        const int ns_ = stages_for(bt.maxM_full, dev::TP_BK);
        dev::k_tgemm_ws<dev::TP_BK><<<dim3(s.NB / 64, (unsigned)(bt.ngitems - bt.ngitems_tri)), 288,
                                     dev::tgemm_ws_smem(bt.maxM_full, ns_, dev::TP_BK), st>>>(
            bt.gitems.p + bt.ngitems_tri, bt.gentries.p, s.d_zrow.p, ns_, bt.maxM_full);
        CK(cudaGetLastError());
      }
    }
    CK(cudaEventRecord(s.ev[4], st));
    if (bt.noitems) {
      const dim3 grid(s.NB / 16, (unsigned)bt.noitems);
      const size_t smem = (size_t)(t.nch * 16 * (16 * 20 + 4) + t.nch * 16 * 20 + 16 * (t.nch * 16 + 4)) * sizeof(double);
      if (t.nch == 1)
        dev::k_offdiag_mma<1><<<grid, 256, smem, st>>>(s.bd, bt.oitems.p, bt.oentries.p, s.d_R.p, s.d_small.p,
                                                    s.d_big.p, s.d_blk_off.p);
      else
        dev::k_offdiag_mma<2><<<grid, 256, smem, st>>>(s.bd, bt.oitems.p, bt.oentries.p, s.d_R.p, s.d_small.p,
                                                    s.d_big.p, s.d_blk_off.p);
      CK(cudaGetLastError());
    }
    CK(cudaEventRecord(s.ev[5], st));
    if (plan->batches.size() > 1) CK(cudaStreamSynchronize(st));   // the timing events are re-used by the next batch
    const int ngl = (bt.ngitems_tri ? 1 : 0) + (bt.ngitems > bt.ngitems_tri ? 1 : 0);
    tm_.launches += 1 + ngl + (bt.noitems ? 1 : 0);
    tm_.launches_fold++;
    tm_.launches_tgemm += ngl;
    tm_.launches_offdiag += bt.noitems ? 1 : 0;
    if (plan->batches.size() > 1) {
      float ms;
      CK(cudaEventElapsedTime(&ms, s.ev[2], s.ev[3]));
      ms_fold += ms;
      CK(cudaEventElapsedTime(&ms, s.ev[3], s.ev[4]));
      ms_tg += ms;
      CK(cudaEventElapsedTime(&ms, s.ev[4], s.ev[5]));
      ms_off += ms;
    }
  }
  // 6. reduce the K-split partials of the own units, complete Kc over the ranks, unpack
  CK(cudaEventRecord(s.ev[6], st));
  if (S > 1 && plan->nreduce) {
    dev::k_reduce_partials<<<plan->nreduce, 256, 0, st>>>(plan->d_reduce.p, s.d_Kc.p, s.d_Kacc.p, plan->part_stride);
    CK(cudaGetLastError());
    tm_.launches++;
  }
  if (comm_ && plan->seg > 0) comm_->all_gather_inplace(s.d_Kc.p, (size_t)plan->seg, st);
  CK(cudaStreamWaitEvent(st, s.ev_kzero, 0));
  if (plan->nblocks) {
    dev::UnpackDev u{plan->d_op_src.p, plan->d_op_tri.p, plan->d_op_ldk.p, plan->d_blocks.p, plan->d_unit_off.p, s.d_ang_sec.p, s.d_ang_pos.p,
                     s.kscale};
    dev::k_unpack_K<<<plan->nblocks, 256, 0, st>>>(s.bd, u, s.d_Kc.p, dK, ldK);
    CK(cudaGetLastError());
  }
  CK(cudaEventRecord(s.ev[7], st));
  CK(cudaStreamSynchronize(st));
  if (plan->batches.size() == 1) {
    CK(cudaEventElapsedTime(&ms_fold, s.ev[2], s.ev[3]));
    CK(cudaEventElapsedTime(&ms_tg, s.ev[3], s.ev[4]));
    CK(cudaEventElapsedTime(&ms_off, s.ev[4], s.ev[5]));
  }
  CK(cudaEventElapsedTime(&tm_.pack, s.ev[0], s.ev[1]));
  CK(cudaEventElapsedTime(&tm_.unpack, s.ev[6], s.ev[7]));
  CK(cudaEventElapsedTime(&tm_.total, s.ev[0], s.ev[7]));
  {
    static const bool trace = getenv("HFQ_TRACE") && atoi(getenv("HFQ_TRACE"));
    if (trace) {
      float kz = 0, t6 = 0;
      cudaEventElapsedTime(&kz, s.ev_start, s.ev_kzero);
      cudaEventElapsedTime(&t6, s.ev[0], s.ev[6]);
      fprintf(stderr, "[hfq] exchange_dev: K cleared on the side stream %.2f ms after the start, kernels done at %.2f ms, reduce+gather+unpack %.2f ms\n",
              kz, t6, tm_.unpack);
    }
  }
  tm_.fold = ms_fold;
  tm_.tgemm = ms_tg;
  tm_.offdiag = ms_off;
  tm_.flops_fold = plan->fl_fold;
  tm_.flops_tgemm = plan->fl_tg;
  tm_.flops_offdiag = plan->fl_off;
  tm_.alg_fold = plan->al_fold;
  tm_.alg_tgemm = plan->al_tg;
  tm_.alg_offdiag = plan->al_off;
  tm_.launches += 2;
}

const BasisTables &Engine::tables() const { return p_->t; }

void Engine::fence_stream(cudaStream_t on, cudaStream_t waiter) {
  if (on == waiter) return;
  Impl &s = *p_;
  if (!s.ev_fence) CK(cudaEventCreateWithFlags(&s.ev_fence, cudaEventDisableTiming));
  CK(cudaEventRecord(s.ev_fence, on));
  CK(cudaStreamWaitEvent(waiter, s.ev_fence, 0));
}

Engine::~Engine() {
  plans_.reset();
  if (p_) {
    for (auto &e : p_->ev) cudaEventDestroy(e);
    for (cudaEvent_t e : {p_->ev_packed, p_->ev_jdone, p_->ev_j, p_->ev_jcopied, p_->ev_start, p_->ev_kzero, p_->ev_fence})
      if (e) cudaEventDestroy(e);
    for (cudaStream_t st : {p_->copy_stream, p_->j_stream, p_->aux_stream})
      if (st) cudaStreamDestroy(st);
    if (!p_->norms_host.empty()) cudaHostUnregister(p_->norms_host.data());
    if (p_->flags_host) cudaFreeHost(p_->flags_host);
  }
  comm_.reset();
  if (stream_) cudaStreamDestroy(stream_);
}

// sector of every dense basis function and the list of output sector pairs written by the last
// exchange call (everything else in K is exactly zero)
void Engine::output_pattern(std::vector<int> &bf_sector, std::vector<int> &pairs, bool coulomb) const {
  const Impl &s = *p_;
  bf_sector.clear();
  for (int a = 0; a < s.t.Nang(); a++)
    for (int r = s.ang_skip[a]; r < s.t.Nrad; r++) bf_sector.push_back(s.ang_sec[a]);
  pairs.clear();
  if (coulomb) {
    // J block (ang i, ang j) is read from Jsec[sp = (sector j, sector i)]: rows belong to sector i
    for (size_t sp = 0; sp < last_active_j_.size(); sp++)
      if (last_active_j_[sp]) {
        pairs.push_back((int)(sp % s.ns));
        pairs.push_back((int)(sp / s.ns));
      }
    return;
  }
  for (size_t op = 0; op < last_active_ops_.size(); op++)
    if (last_active_ops_[op] >= 0) {
      pairs.push_back((int)(op / s.ns));
      pairs.push_back((int)(op % s.ns));
    }
}

// ---------------------------------------------------------------------------
// coulomb
// ---------------------------------------------------------------------------
void Engine::coulomb_dev(const double *dP, int64_t ldP, double *dJ, int64_t ldJ, int shard, int nshards,
                         cudaStream_t st) {
  coulomb_run(dP, ldP, dJ, ldJ, shard, nshards, st, false);
}

// async: no timing events (they are shared with the exchange path) and no final synchronisation -- used by
// jk_dev to run the Coulomb chain on a second stream next to the exchange kernels.
void Engine::coulomb_run(const double *dP, int64_t ldP, double *dJ, int64_t ldJ, int shard, int nshards,
                         cudaStream_t st, bool async) {
  Impl &s = *p_;
  const BasisTables &t = s.t;
  if (t.pairwise())
    throw std::logic_error("coulomb is not available on range-separated (erfc pair-tensor) tables\n");
  if (s.radial_only) throw std::logic_error("coulomb is not available on batch tables (use hfq_coulomb_radial_batch)");
  CK(cudaSetDevice(device_));
  const int na = t.Nang(), ns = s.ns, nq = s.NL * t.nch;
  if (!async) tm_ = EngineTimings();
  if (!async) CK(cudaEventRecord(s.ev[0], st));
  if (!s.packed_valid) pack_density(dP, ldP, st);
  const std::vector<int> &splist = s.packed_splist;
  // sharding: this rank handles the multipoles L in [L0, L1) (partial J, summed by the caller's all-reduce)
  const int L0 = (int)((int64_t)s.NL * shard / nshards), L1 = (int)((int64_t)s.NL * (shard + 1) / nshards);
  const int q0 = L0 * t.nch, nqs = (L1 - L0) * t.nch;
  if (!async) CK(cudaEventRecord(s.ev[1], st));
  if (s.d_Paux.n == 0) {
    s.d_Paux.alloc((size_t)s.nM * nq * s.Npix, &dev_bytes_);
    s.d_JauxT.alloc((size_t)s.nM * nq * s.Npix, &dev_bytes_);
    s.d_Jsec.alloc((size_t)ns * ns * s.Npix * s.NB, &dev_bytes_);
  }
  // The descriptors depend only on which sector pairs carry density (and on the shard): cached like
  // the exchange plan, so an SCF iteration issues no host->device descriptor traffic and no mid-call sync.
  Impl::JPlan &jp = s.jplan;
  if (!jp.valid || jp.splist != splist || jp.shard != shard || jp.nshards != nshards) {
    CK(cudaStreamSynchronize(st));   // a previous call may still read the descriptor buffers
    jp.valid = true;
    jp.splist = splist;
    jp.shard = shard;
    jp.nshards = nshards;
    // fold: Paux[Mi][q][pix] = sum over sector pairs with m_a - m_b = M of G[sp][q][:] . Ppix[sp][pix][:]
    std::vector<dev::GemmItem> items, uitems;
    std::vector<dev::GemmEntry> entries, uentries;
    jp.M_active.assign(s.nM, 0);
    for (int Mi = 0; Mi < s.nM; Mi++) {
      const int M = Mi - (s.mmax - s.mmin);
      dev::GemmItem gi{};
      gi.C = s.d_Paux.p + ((size_t)Mi * nq + q0) * s.Npix;
      gi.M = nqs;
      gi.N = s.Npix;
      gi.K = s.NB;
      gi.ent0 = (int)entries.size();
      for (int sp : splist) {
        if (s.sec_m[sp / ns] - s.sec_m[sp % ns] != M) continue;
        dev::GemmEntry ge;
        ge.A = s.d_G.p + ((size_t)sp * nq + q0) * s.NB;
        ge.lda = s.NB;
        ge.B = s.d_Ppix.p + (size_t)sp * s.Npix * s.NB;
        entries.push_back(ge);
      }
      gi.ent1 = (int)entries.size();
      if (gi.ent1 == gi.ent0) continue;
      jp.M_active[Mi] = 1;   // pattern of the COMPLETE J: identical on every shard
      if (nqs == 0) continue;
      gi.accumulate = 0;
      gi.ldb = s.NB;
      gi.ldc = s.Npix;
      gi.alpha = 1.0;
      items.push_back(gi);
    }
    // unfold: Jsec[sp=(sj,si)][pix][j*NP+i] = sum_q JauxT[Mi][pix][q] G[sp][q][j*NP+i]
    jp.sp_active.assign((size_t)ns * ns, 0);
    for (int sp = 0; sp < ns * ns; sp++) {
      const int M = s.sec_m[sp / ns] - s.sec_m[sp % ns], Mi = M + (s.mmax - s.mmin);
      if (!jp.M_active[Mi]) continue;
      jp.sp_active[sp] = 1;
      dev::GemmItem gi{};
      gi.C = s.d_Jsec.p + (size_t)sp * s.Npix * s.NB;
      gi.browoff = s.d_browoff_G.p;
      gi.M = s.Npix;
      gi.N = s.NB;
      gi.K = nqs;
      gi.ent0 = (int)uentries.size();
      dev::GemmEntry ge;
      ge.A = s.d_JauxT.p + (size_t)Mi * s.Npix * nq + q0;
      ge.lda = nq;
      ge.B = s.d_G.p + ((size_t)sp * nq + q0) * s.NB;
      uentries.push_back(ge);
      gi.ent1 = (int)uentries.size();
      gi.accumulate = 0;
      gi.ldc = s.NB;
      gi.alpha = 1.0;
      uitems.push_back(gi);
    }
    // dense blocks (ang i, ang j) that read an active sector pair sp = (sector j, sector i)
    if (na > 0xffff) throw std::runtime_error("Engine: more than 65535 angular functions");
    std::vector<int> blocks;
    for (int ai = 0; ai < na; ai++)
      for (int aj = 0; aj < na; aj++)
        if (jp.sp_active[(size_t)s.ang_sec[aj] * ns + s.ang_sec[ai]]) blocks.push_back(ai | (aj << 16));
    jp.nfold = (int)items.size();
    jp.nunfold = (int)uitems.size();
    jp.nblocks = (int)blocks.size();
    jp.d_items.upload(items, &dev_bytes_);
    jp.d_entries.upload(entries, &dev_bytes_);
    jp.d_uitems.upload(uitems, &dev_bytes_);
    jp.d_uentries.upload(uentries, &dev_bytes_);
    jp.d_blocks.upload(blocks, &dev_bytes_);
  }
  last_active_j_.assign(jp.sp_active.begin(), jp.sp_active.end());
  // inactive M channels must read as zero in the radial step
  for (int Mi = 0; Mi < s.nM; Mi++)
    if (!jp.M_active[Mi]) CK(cudaMemsetAsync(s.d_Paux.p + (size_t)Mi * nq * s.Npix, 0, (size_t)nq * s.Npix * sizeof(double), st));
  launch_gemm<true>(jp.d_items.p, jp.d_entries.p, jp.nfold, nqs, s.Npix, st);
  if (!async) CK(cudaEventRecord(s.ev[2], st));
  if (nshards > 1)   // multipoles of other shards must read as zero in the unfold
    CK(cudaMemsetAsync(s.d_JauxT.p, 0, (size_t)s.nM * nq * s.Npix * sizeof(double), st));
  // radial step
  dev::JRadDev jr{s.d_chan_of.p, s.d_jfac.p, s.d_blk_off.p, s.d_B_off.p, s.d_sig_off.p, s.d_rank.p,
                  s.d_small.p, s.d_big.p, s.d_B.p, s.d_sigma.p, s.nM};
  if (L1 > L0) dev::k_jradial<<<dim3(L1 - L0, s.nM), 256, 0, st>>>(s.bd, jr, L0, s.d_Paux.p, s.d_JauxT.p);
  CK(cudaGetLastError());
  if (!async) CK(cudaEventRecord(s.ev[3], st));
  launch_gemm<false>(jp.d_uitems.p, jp.d_uentries.p, jp.nunfold, s.Npix, s.NB, st);
  if (!async) CK(cudaEventRecord(s.ev[4], st));
  // clear J at memset speed, then unpack only the angular blocks of active sector pairs
  CK(cudaMemset2DAsync(dJ, (size_t)ldJ * sizeof(double), 0, (size_t)nbf_ * sizeof(double), (size_t)nbf_, st));
  if (jp.nblocks) {
    dev::k_unpack_J<<<jp.nblocks, 256, 0, st>>>(s.bd, s.d_ang_sec.p, s.d_ang_pos.p, jp.d_blocks.p, s.d_Jsec.p, dJ, ldJ);
    CK(cudaGetLastError());
  }
  if (!async) CK(cudaEventRecord(s.ev[5], st));
  if (async) return;
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&tm_.pack, s.ev[0], s.ev[1]));
  CK(cudaEventElapsedTime(&tm_.fold, s.ev[1], s.ev[2]));
  CK(cudaEventElapsedTime(&tm_.offdiag, s.ev[2], s.ev[3]));
  CK(cudaEventElapsedTime(&tm_.tgemm, s.ev[3], s.ev[4]));
  CK(cudaEventElapsedTime(&tm_.unpack, s.ev[4], s.ev[5]));
  CK(cudaEventElapsedTime(&tm_.total, s.ev[0], s.ev[5]));
  tm_.launches = 6;
}

// ---------------------------------------------------------------------------
// host-pointer wrappers
// ---------------------------------------------------------------------------
Engine::HostRanges Engine::host_ranges(bool coulomb) const {
  std::vector<int> bfsec, pairs;
  output_pattern(bfsec, pairs, coulomb);
  const int n = nbf_, ns = p_->ns;
  std::vector<int> smin(ns, n), smax(ns, -1);
  for (int i = 0; i < n; i++) {
    smin[bfsec[i]] = std::min(smin[bfsec[i]], i);
    smax[bfsec[i]] = std::max(smax[bfsec[i]], i);
  }
  // bounding row range of the active row sectors of every column sector
  std::vector<int> cmin(ns, n), cmax(ns, 0);
  for (size_t k = 0; k + 1 < pairs.size(); k += 2) {
    const int sr = pairs[k], sc = pairs[k + 1];
    if (smax[sr] < 0) continue;
    cmin[sc] = std::min(cmin[sc], smin[sr]);
    cmax[sc] = std::max(cmax[sc], smax[sr] + 1);
  }
  HostRanges hr;
  hr.r0.resize(n);
  hr.r1.resize(n);
  for (int c = 0; c < n; c++) {
    const int sc = bfsec[c];
    hr.r0[c] = cmin[sc] < cmax[sc] ? cmin[sc] : 0;
    hr.r1[c] = cmin[sc] < cmax[sc] ? cmax[sc] : 0;
  }
  return hr;
}

double Engine::copy_ranges_async(double *H, int64_t ldH, const double *D, const HostRanges &hr, cudaStream_t st, int cb,
                                 int ce) const {
  const int n = nbf_;
  double bytes = 0.0;
  if ((int)hr.r0.size() != n) return 0.0;
  if (ce < 0) ce = n;
  for (int c0 = cb; c0 < ce;) {
    int c1 = c0 + 1;
    while (c1 < ce && hr.r0[c1] == hr.r0[c0] && hr.r1[c1] == hr.r1[c0]) c1++;
    const int r0 = hr.r0[c0], r1 = hr.r1[c0];
    bytes += (double)(r1 - r0) * (c1 - c0) * sizeof(double);
    if (r1 > r0)
      CK(cudaMemcpy2DAsync(H + (int64_t)c0 * ldH + r0, ldH * sizeof(double), D + (int64_t)c0 * n + r0, (size_t)n * sizeof(double),
                           (size_t)(r1 - r0) * sizeof(double), (size_t)(c1 - c0), cudaMemcpyDeviceToHost, st));
    c0 = c1;
  }
  return bytes;
}

namespace {
std::atomic<int> g_host_threads{0};
}
void set_host_threads(int n) { g_host_threads.store(n); }
// threads of the host-side helpers (zero-fill, verification scan): all of this process' share of the cores, minus two for
// the threads that feed the GPU when the share is large enough (a rank of an 8-GPU job may own only 2-4 cores)
static int helper_threads() {
  const int ht = host_threads();
  return std::max(1, ht > 4 ? ht - 2 : ht);
}
int host_threads() {
  const int n = g_host_threads.load();
  return n > 0 ? n : omp_get_max_threads();
}

// true if any element of the n x n column-major matrix H outside the row ranges [r0[c], r1[c]) has a non-zero bit
// pattern (-0.0 and NaN count as non-zero: the caller then falls back to the complete upload)
bool Engine::nonzero_outside(const double *H, int64_t ldH, int n, const std::vector<int> &r0, const std::vector<int> &r1) {
  const int nthr = helper_threads();
  int found = 0;
#pragma omp parallel for schedule(static) num_threads(nthr) reduction(| : found)
  for (int c = 0; c < n; c++) {
    const unsigned long long *col = reinterpret_cast<const unsigned long long *>(H + (int64_t)c * ldH);
    unsigned long long acc = 0;
    for (int i = 0; i < r0[c]; i++) acc |= col[i];
    for (int i = r1[c]; i < n; i++) acc |= col[i];
    found |= acc != 0 ? 1 : 0;
  }
  return found != 0;
}

namespace {
// Zero-fill with non-temporal stores: a column piece (~100 KB) is below the size from which memset switches to
// streaming stores, so memset would read every cache line before overwriting it (read-for-ownership) and double the
// memory traffic of what is the longest part of a host-pointer call.
inline void stream_zero(double *p, size_t count) {
#if defined(__SSE2__)
  size_t i = 0;
  while (i < count && (reinterpret_cast<uintptr_t>(p + i) & 15)) p[i++] = 0.0;
  const __m128d z = _mm_setzero_pd();
  for (; i + 2 <= count; i += 2) _mm_stream_pd(p + i, z);
  for (; i < count; i++) p[i] = 0.0;
#else
  std::memset(p, 0, count * sizeof(double));
#endif
}
}  // namespace

void Engine::zero_outside(double *H, int64_t ldH, int n, const HostRanges &hr, int cb, int ce) {
  const bool none = (int)hr.r0.size() != n;   // no pattern: everything is zero
  // leave cores to the thread that feeds the GPU.  The count set through hfq_set_host_threads is kept in a process
  // global: this runs in a helper std::thread, whose OpenMP ICVs start from OMP_NUM_THREADS again (torchrun exports
  // 1), not from the omp_set_num_threads of the calling thread
  const int nthr = helper_threads();
  if (ce < 0) ce = n;
#pragma omp parallel num_threads(nthr)
  {
#pragma omp for schedule(static)
    for (int c = cb; c < ce; c++) {
      double *col = H + (int64_t)c * ldH;
      if (none) {
        stream_zero(col, (size_t)n);
        continue;
      }
      if (hr.r0[c] > 0) stream_zero(col, (size_t)hr.r0[c]);
      if (hr.r1[c] < n) stream_zero(col + hr.r1[c], (size_t)(n - hr.r1[c]));
    }
#if defined(__SSE2__)
    _mm_sfence();   // every thread orders its own streaming stores before it leaves the region
#endif
  }
}

namespace {
// joins on scope exit so that an exception cannot leave a worker writing into the caller's buffer
struct Joiner {
  std::thread t;
  ~Joiner() {
    if (t.joinable()) t.join();
  }
};
}  // namespace

void Engine::coulomb(const double *P, int64_t ldP, double *J, int64_t ldJ) {
  Impl &s = *p_;
  CK(cudaSetDevice(device_));
  const size_t n = (size_t)nbf_;
  if (s.d_P.n < n * n) s.d_P.alloc(n * n, &dev_bytes_);
  if (s.d_O.n < n * n) s.d_O.alloc(n * n, &dev_bytes_);
  CK(cudaMemcpy2DAsync(s.d_P.p, n * sizeof(double), P, ldP * sizeof(double), n * sizeof(double), n,
                       cudaMemcpyHostToDevice, stream_));
  coulomb_dev(s.d_P.p, (int64_t)n, s.d_O.p, (int64_t)n, 0, 1, stream_);
  const HostRanges hr = host_ranges(true);
  tm_.h2d_bytes = (double)n * n * sizeof(double);
  tm_.d2h_bytes = copy_ranges_async(J, ldJ, s.d_O.p, hr, stream_);
  zero_outside(J, ldJ, nbf_, hr);
  CK(cudaStreamSynchronize(stream_));
}

void Engine::exchange(const double *P, int64_t ldP, double *K, int64_t ldK) {
  Impl &s = *p_;
  CK(cudaSetDevice(device_));
  const size_t n = (size_t)nbf_;
  if (s.d_P.n < n * n) s.d_P.alloc(n * n, &dev_bytes_);
  if (s.d_O.n < n * n) s.d_O.alloc(n * n, &dev_bytes_);
  CK(cudaMemcpy2DAsync(s.d_P.p, n * sizeof(double), P, ldP * sizeof(double), n * sizeof(double), n,
                       cudaMemcpyHostToDevice, stream_));
  HostRanges hr;
  Joiner zero;
  plan_hook_ = [&]() {
    hr = host_ranges(false);
    zero.t = std::thread([&]() { zero_outside(K, ldK, nbf_, hr); });
  };
  try {
    exchange_dev(s.d_P.p, (int64_t)n, s.d_O.p, (int64_t)n, 0, 1, stream_);
  } catch (...) {
    plan_hook_ = nullptr;
    throw;
  }
  plan_hook_ = nullptr;
  tm_.h2d_bytes = (double)n * n * sizeof(double);
  tm_.d2h_bytes = copy_ranges_async(K, ldK, s.d_O.p, hr, stream_);
  CK(cudaStreamSynchronize(stream_));
}

// Bounding non-zero row range of every dense column of the density that was packed last (from the
// per-block max |P| of k_block_norms).
Engine::HostRanges Engine::density_ranges() const {
  const Impl &s = *p_;
  const int na = s.t.Nang(), n = nbf_;
  const size_t nn = (size_t)na * na;
  HostRanges hr;
  hr.r0.assign(n, 0);
  hr.r1.assign(n, 0);
  for (int c = 0; c < na; c++) {
    int lo = n, hi = 0;
    for (int a = 0; a < na; a++)
      if (!(s.norms_host[(size_t)a * na + c] == 0.0)) {   // block (a, c): rows of a, columns of c
        lo = std::min(lo, s.ang_off[a]);
        hi = std::max(hi, s.ang_off[a] + s.t.Nrad - s.ang_skip[a]);
      }
    if (hi <= lo) lo = hi = 0;
    for (int k = 0; k < s.t.Nrad - s.ang_skip[c]; k++) {
      hr.r0[s.ang_off[c] + k] = lo;
      hr.r1[s.ang_off[c] + k] = hi;
    }
  }
  return hr;
}

// Fused host entry point: the non-zero row ranges of J are copied back on a second stream while the
// exchange kernels run, then those of K; host threads zero-fill the rest of both result matrices
// in the meantime.
//
// Upload of P.  An SCF calls this with the same block structure of P every iteration, and the
// 1.76 GB dense upload (N2) is longer than the whole build.  So, for pinned host buffers, the call
// SPECULATES: it uploads only the row ranges that were non-zero in the previous call into a zeroed
// device matrix and starts computing, while the complete matrix is uploaded on a third stream; a
// compare kernel then checks that the assembled and the complete matrix are bit-identical.  If
// they are not (the structure changed), the build is repeated from the complete upload.  The
// result never depends on the prediction; all of P still crosses PCIe, but behind the compute.
void Engine::coulomb_exchange(const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K,
                              int64_t ldK) {
  Impl &s = *p_;
  CK(cudaSetDevice(device_));
  static const bool allow = !(getenv("HFQ_NO_SPECULATION") && atoi(getenv("HFQ_NO_SPECULATION")));
  bool spec = false;
  if (allow && (int)s.pred_r0.size() == nbf_) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, P) == cudaSuccess)
      spec = attr.type == cudaMemoryTypeHost;   // page-locked: the background upload is truly asynchronous
    else
      cudaGetLastError();
  }
  if (spec && fused_host(P, ldP, kscale, J, ldJ, K, ldK, true)) {
    spec_hits_++;
    return;
  }
  fused_host(P, ldP, kscale, J, ldJ, K, ldK, false);
}

bool Engine::fused_host(const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K, int64_t ldK,
                        bool spec) {
  Impl &s = *p_;
  const size_t n = (size_t)nbf_;
  static const bool trace = getenv("HFQ_TRACE") && atoi(getenv("HFQ_TRACE"));
  const auto tstart = std::chrono::steady_clock::now();
  auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tstart).count(); };
  double t_pack = 0, t_j = 0, t_k = 0, t_sync = 0, t_zero = 0, t_zero_end = 0;
  if (s.d_P.n < n * n) s.d_P.alloc(n * n, &dev_bytes_);
  if (s.d_O.n < n * n) s.d_O.alloc(n * n, &dev_bytes_);
  if (s.d_O2.n < n * n) s.d_O2.alloc(n * n, &dev_bytes_);
  if (!s.copy_stream) {
    CK(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&s.ev_j, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&s.ev_jcopied, cudaEventDisableTiming));
  }
  double h2d = spec ? 0.0 : (double)n * n * sizeof(double);
  std::atomic<int> host_mismatch{0};
  Joiner verify;   // joined before the function returns (declared before the scopes that may throw)
  if (spec) {
    CK(cudaMemsetAsync(s.d_P.p, 0, n * n * sizeof(double), stream_));
    for (size_t c0 = 0; c0 < n;) {
      size_t c1 = c0 + 1;
      while (c1 < n && s.pred_r0[c1] == s.pred_r0[c0] && s.pred_r1[c1] == s.pred_r1[c0]) c1++;
      const int r0 = s.pred_r0[c0], r1 = s.pred_r1[c0];
      if (r1 > r0) {
        CK(cudaMemcpy2DAsync(s.d_P.p + c0 * n + r0, n * sizeof(double), P + (int64_t)c0 * ldP + r0, ldP * sizeof(double),
                             (size_t)(r1 - r0) * sizeof(double), c1 - c0, cudaMemcpyHostToDevice, stream_));
        h2d += (double)(r1 - r0) * (c1 - c0) * sizeof(double);
      }
      c0 = c1;
    }
    // The prediction is verified on the HOST while the GPU computes: everything of P outside the uploaded row ranges
    // must be exactly zero (bit pattern 0).  Only then is the device matrix -- zeros plus the uploaded ranges -- the
    // caller's matrix; the scan reads P once at host-memory speed instead of sending all of it over PCIe.
    verify.t = std::thread([&, this]() { host_mismatch.store(nonzero_outside(P, ldP, nbf_, s.pred_r0, s.pred_r1) ? 1 : 0); });
  } else {
    CK(cudaMemcpy2DAsync(s.d_P.p, n * sizeof(double), P, ldP * sizeof(double), n * sizeof(double), n,
                         cudaMemcpyHostToDevice, stream_));
  }
  s.want_norms_host = true;   // density_ranges() below predicts the next upload from the per-block norms
  try {
    pack_density(s.d_P.p, (int64_t)n, stream_);
  } catch (...) {
    s.want_norms_host = false;
    throw;
  }
  s.want_norms_host = false;
  t_pack = ms_since();
  s.packed_valid = true;
  HostRanges hrj, hrk;
  double jbytes = 0.0;
  int mismatch = 0;
  {
    Joiner zero;
    try {
      coulomb_dev(s.d_P.p, (int64_t)n, s.d_O2.p, (int64_t)n, 0, 1, stream_);
      const EngineTimings tj = tm_;
      t_j = ms_since();
      // J is complete (coulomb_dev synchronises): copy it out while K is being built
      // Zero-filling the two dense host results (3.5 GB for N2) at host-memory speed is the longest part of the
      // call (55-65 ms measured), while the device-to-host copy engine has nothing to do during the K build.  With a
      // page-locked J the complete dense J -- zeros included, the device matrix is fully defined -- therefore
      // travels over PCIe while K is being built, and the host threads only zero-fill K.
      bool j_dense = false;
      {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, J) == cudaSuccess)
          j_dense = attr.type == cudaMemoryTypeHost;
        else
          cudaGetLastError();
      }
      hrj = host_ranges(true);
      // columns [0, cj) of J travel densely, the others as row ranges + host zero-fill; the split follows the
      // measured finish times of the two sides (PCIe rate and host-memory speed differ from host to host)
      const int cj = j_dense ? std::min((int)n, std::max(0, (int)std::lround(s.j_dense_frac * (double)n))) : 0;
      if (cj > 0) {
        CK(cudaMemcpy2DAsync(J, (size_t)ldJ * sizeof(double), s.d_O2.p, n * sizeof(double), n * sizeof(double), (size_t)cj,
                             cudaMemcpyDeviceToHost, s.copy_stream));
        jbytes = (double)n * cj * sizeof(double);
      }
      if (cj < (int)n) jbytes += copy_ranges_async(J, ldJ, s.d_O2.p, hrj, s.copy_stream, cj, (int)n);
      s.kscale = kscale;
      plan_hook_ = [&, cj]() {
        hrk = host_ranges(false);
        zero.t = std::thread([&, cj]() {
          const double z0 = ms_since();
          if (cj < nbf_) zero_outside(J, ldJ, nbf_, hrj, cj, nbf_);
          zero_outside(K, ldK, nbf_, hrk);
          t_zero = ms_since() - z0;
          t_zero_end = ms_since();
        });
      };
      exchange_dev(s.d_P.p, (int64_t)n, s.d_O.p, (int64_t)n, 0, 1, stream_);
      t_k = ms_since();
      tm_.launches += tj.launches;
      tm_.total += tj.total;
    } catch (...) {
      plan_hook_ = nullptr;
      s.packed_valid = false;
      s.kscale = 1.0;
      cudaStreamSynchronize(s.copy_stream);   // the J copy-back must not write into the caller's buffer after the error return
      throw;
    }
    plan_hook_ = nullptr;
    s.packed_valid = false;
    s.kscale = 1.0;
    tm_.h2d_bytes = h2d;
    tm_.d2h_bytes = jbytes + copy_ranges_async(K, ldK, s.d_O.p, hrk, stream_);
    CK(cudaStreamSynchronize(stream_));
    CK(cudaStreamSynchronize(s.copy_stream));
    t_sync = ms_since();
  }   // zero-fill thread joined here
  if (verify.t.joinable()) verify.t.join();
  mismatch = host_mismatch.load();
  // steer the dense share of J: the side that finished later gives work to the other
  if (t_zero_end > 0.0) {
    if (t_sync > t_zero_end + 2.0)
      s.j_dense_frac = std::max(0.0, s.j_dense_frac - 0.1);
    else if (t_zero_end > t_sync + 2.0)
      s.j_dense_frac = std::min(1.0, s.j_dense_frac + 0.1);
  }
  if (trace)
    fprintf(stderr, "[hfq] fused_host spec=%d: pack done %.1f  J done %.1f  K done %.1f  copies done %.1f  joined %.1f ms (zero-fill %.1f, dense share of J next %.1f)\n",
            (int)spec, t_pack, t_j, t_k, t_sync, ms_since(), t_zero, s.j_dense_frac);
  if (mismatch) {
    s.pred_r0.clear();
    s.pred_r1.clear();
    return false;
  }
  const HostRanges pr = density_ranges();
  s.pred_r0 = pr.r0;
  s.pred_r1 = pr.r1;
  return true;
}

// Multi-GPU build with HOST matrices shared by all ranks (one process per GPU; P, J, K in memory every rank can
// address, e.g. a POSIX shared-memory segment, ideally page-locked by each rank).  Every rank moves 1/nranks of the
// bytes over ITS OWN PCIe link: it uploads its column slice of P, one in-place ncclAllGather over NVLink completes
// the density on every GPU, the sharded build runs (jk_dev with the communicator), and every rank copies its column
// slice of the non-zero row ranges of J and K back and zero-fills the rest of its slice.  The caller synchronises
// the ranks afterwards (the matrices are complete when every rank has returned).
void Engine::jk_spmd_host(const double *P, int64_t ldP, double kscale, double *J, int64_t ldJ, double *K, int64_t ldK) {
  Impl &s = *p_;
  if (!comm_) throw std::logic_error("hfq_coulomb_exchange_spmd needs a communicator (hfq_comm_init)");
  CK(cudaSetDevice(device_));
  const size_t n = (size_t)nbf_;
  const int N = comm_->size(), r = comm_->rank();
  const size_t chunk = (n + N - 1) / N;                 // columns per rank
  const int cb = (int)std::min(n, r * chunk), ce = (int)std::min(n, (r + 1) * chunk);
  if (s.d_P.n < chunk * N * n) s.d_P.alloc(chunk * N * n, &dev_bytes_);
  if (s.d_O.n < n * n) s.d_O.alloc(n * n, &dev_bytes_);
  if (s.d_O2.n < n * n) s.d_O2.alloc(n * n, &dev_bytes_);
  if (ce > cb)
    CK(cudaMemcpy2DAsync(s.d_P.p + (size_t)cb * n, n * sizeof(double), P + (int64_t)cb * ldP, ldP * sizeof(double),
                         n * sizeof(double), (size_t)(ce - cb), cudaMemcpyHostToDevice, stream_));
  comm_->all_gather_inplace(s.d_P.p, chunk * n, stream_);
  HostRanges hrj, hrk;
  {
    Joiner zero;   // host threads zero-fill this rank's column slice of J and K while the GPUs compute
    plan_hook_ = [&]() {
      hrj = host_ranges(true);
      hrk = host_ranges(false);
      zero.t = std::thread([&, this]() {
        zero_outside(J, ldJ, nbf_, hrj, cb, ce);
        zero_outside(K, ldK, nbf_, hrk, cb, ce);
      });
    };
    try {
      jk_dev(s.d_P.p, (int64_t)n, kscale, s.d_O2.p, (int64_t)n, s.d_O.p, (int64_t)n, 0, 1, stream_);
    } catch (...) {
      plan_hook_ = nullptr;
      throw;
    }
    plan_hook_ = nullptr;
    tm_.h2d_bytes = (double)(ce - cb) * n * sizeof(double);
    tm_.d2h_bytes = copy_ranges_async(J, ldJ, s.d_O2.p, hrj, stream_, cb, ce) + copy_ranges_async(K, ldK, s.d_O.p, hrk, stream_, cb, ce);
    CK(cudaStreamSynchronize(stream_));
  }   // zero-fill thread joined here
}

const double *Engine::device_density() const { return p_->d_P.p; }

}  // namespace hfq
