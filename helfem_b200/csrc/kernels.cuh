// CUDA kernels of the J/K engine (sm_100a).  All arithmetic is FP64; the
// GEMM-shaped work runs on the FP64 tensor pipe (mma.sync m8n8k4 -> SASS
// DMMA.8x8x4) with shared-memory staged operands.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace hfq {
namespace dev {

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// Dense <-> sector-packed layout
//   dense:  P[(ang a, radial r), (ang b, radial c)], column-major, boundary
//           functions of dropped shells removed (index maps ang_off/ang_skip)
//   packed: Ppix[sp][pix = r*Nrad + c][ia][ib], sp = sector pair (sa, sb),
//           ia/ib = position of the angular function inside its m-sector
// ---------------------------------------------------------------------------

struct BasisDev {
  int Nang, Nrad, Npix, NP, NB, ns, nch, nab, Nel, NL;
  const int *ang_off;    // [Nang] first dense index of angular function
  const int *ang_skip;   // [Nang] 1 if radial function 0 is removed
  const int *sec_n;      // [ns]
  const int *sec_ang;    // [ns*NP] angular function index or -1
  const int *efirst;     // [Nel]
  const int *en;         // [Nel]
  const int *rad_e0;     // [Nrad] first element containing the radial function
  const int *rad_e1;     // [Nrad] last element containing it (elements overlap by one function)
};

// Block norms of P (the reference's screening quantity, src/diatomic/basis.cpp:1855-1864) and the symmetry test,
// one pass over the matrix.  One WARP per unordered pair of angular functions (a <= c): it reads block (a, c) and
// block (c, a) once each, both coalesced (the mirrored tile goes through a warp-private shared-memory tile), and
// produces out[a*Nang+c], out[c*Nang+a] (sums of squares) and, in flags[] (zeroed by the caller), the sector-pair
// flags the plan needs: flags[sp] |= 1 if the block is not exactly zero (it must be packed), |= 2 if its norm is not
// below 10 eps (the reference's skip test, negated so that NaN blocks are kept); flags[ns^2] = bits of
// max |P(i,j) - P(j,i)|, flags[ns^2 + 1] = bits of max |P(i,j)| (non-negative doubles order like their bit
// patterns).  Every entry is combined with OR / MAX, which coincide on {0, 1, 3}: a multi-GPU build lets each rank
// scan the columns c = c0, c0 + cstride, .. and completes the flags with one all-reduce(max).
// grid (number of columns of this rank, ceil(Nang / 4)), 128 threads.
static __global__ void __launch_bounds__(128)
k_block_norms(BasisDev b, const double *__restrict__ P, int64_t ld, const int *__restrict__ ang_sec,
              double *__restrict__ out, unsigned long long *__restrict__ flags, int c0, int cstride) {
  __shared__ double tiles[4][32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + blockIdx.x * cstride, a = blockIdx.y * 4 + warp;
  if (a > c || a >= b.Nang) return;
  double(*tm)[33] = tiles[warp];
  const int sa = b.ang_skip[a], sc = b.ang_skip[c];
  const int na = b.Nrad - sa, nc = b.Nrad - sc;
  const double *base = P + b.ang_off[a] + (int64_t)b.ang_off[c] * ld;   // block (a, c): na x nc
  const double *mirr = P + b.ang_off[c] + (int64_t)b.ang_off[a] * ld;   // block (c, a): nc x na
  double s1 = 0.0, s2 = 0.0, d = 0.0, m = 0.0;
  for (int j0 = 0; j0 < nc; j0 += 32)
    for (int i0 = 0; i0 < na; i0 += 32) {
      __syncwarp();
      // mirror tile: element (j0 + lane, i0 + k) of block (c, a) -> tm[k][lane].  Partial tiles (42 = 32 + 10 radial
      // functions) loop over their own columns only; tile entries outside are never read below
      const int k1 = min(32, na - i0), k2 = min(32, nc - j0);
#pragma unroll 8
      for (int k = 0; k < k1; k++) {
        const int jj = j0 + lane, ii = i0 + k;
        const double v = (jj < nc) ? mirr[jj + (int64_t)ii * ld] : 0.0;
        tm[k][lane] = v;
        if (a != c) {
          s2 += v * v;
          m = fmax(m, fabs(v));
        }
      }
      __syncwarp();
#pragma unroll 8
      for (int k = 0; k < k2; k++) {   // element (i0 + lane, j0 + k) of block (a, c)
        const int ii = i0 + lane, jj = j0 + k;
        if (ii < na) {
          const double v = base[ii + (int64_t)jj * ld];
          s1 += v * v;
          d = fmax(d, fabs(v - tm[lane][k]));
          m = fmax(m, fabs(v));
        }
      }
    }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    d = fmax(d, __shfl_down_sync(0xffffffffu, d, o));
    m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
  }
  if (lane == 0) {
    const double thr2 = 100.0 * 2.220446049250313e-16 * 2.220446049250313e-16;
    out[a * b.Nang + c] = s1;
    unsigned long long f = (!(s1 == 0.0) ? 1ull : 0ull) | (!(s1 < thr2) ? 3ull : 0ull);
    if (f) atomicOr(flags + ang_sec[a] * b.ns + ang_sec[c], f);
    if (a != c) {
      out[c * b.Nang + a] = s2;
      f = (!(s2 == 0.0) ? 1ull : 0ull) | (!(s2 < thr2) ? 3ull : 0ull);
      if (f) atomicOr(flags + ang_sec[c] * b.ns + ang_sec[a], f);
    }
    unsigned long long *mx = flags + b.ns * b.ns;
    if (!(d == 0.0)) atomicMax(mx, (unsigned long long)__double_as_longlong(d != d ? 1e300 : d));   // NaN: not symmetric
    atomicMax(mx + 1, (unsigned long long)__double_as_longlong(m));
  }
}

// grid (Nrad [column radial c], nactive sector pairs); splist[y] = sp index
static __global__ void k_pack(BasisDev b, const double *__restrict__ P, int64_t ld, const int *__restrict__ splist,
                       double *__restrict__ Ppix) {
  const int c = blockIdx.x, sp = splist[blockIdx.y];
  const int sa = sp / b.ns, sb = sp % b.ns;
  double *dst = Ppix + (int64_t)sp * b.Npix * b.NB;
  for (int ib = 0; ib < b.NP; ib++) {   // sector positions may have gaps (parity-class ordering)
    const int angb = b.sec_ang[sb * b.NP + ib];
    if (angb < 0 || c < b.ang_skip[angb]) continue;
    const int64_t col = b.ang_off[angb] + c - b.ang_skip[angb];
    for (int idx = threadIdx.x; idx < b.NP * b.Nrad; idx += blockDim.x) {
      const int ia = idx / b.Nrad, r = idx % b.Nrad;
      const int anga = b.sec_ang[sa * b.NP + ia];
      if (anga < 0 || r < b.ang_skip[anga]) continue;
      const double v = P[b.ang_off[anga] + r - b.ang_skip[anga] + col * ld];
      dst[((int64_t)r * b.Nrad + c) * b.NB + ia * b.NP + ib] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// Exchange fold: R_ab[pix] = fac * Gj_a * P[pix] * Gk_b^T  per (task, pixel)
// ---------------------------------------------------------------------------

struct FoldTask {
  int spj, spk, spp;   // sector pairs of the two coupling tables and of P
  int L;               // multipole order (index into the coupling tables)
  int rslot;           // slot in the R buffer
  int ldk;             // row stride of the folded block R[j][k] inside its NP*NP slot: the (even) number of positions
                       // of the output column sector, so that the consumers see span_j * ldk dense columns
  int pix0, npix;      // the pixels (ri, rl) to fold: pixlist[pix0 .. pix0 + npix).  Not all of them in general: a
                       // symmetric-density diagonal output pair needs el(ri) <= el(rl) only (mirrored by the unpack),
                       // and under owner-computes sharding a rank folds the element pairs it builds
  double fac;          // prefactor incl. (-1)^M
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, int src_bytes) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// NT = NP/8.  Shared memory: Gj[NCH][NP][LD] Gk[NCH][NP][LD], P[2][PB][NP][LD] (cp.async double
// buffer), Yt[PB][NCH][NP][LD]; LD = NP+4 (conflict-free DMMA fragment reads).
// PAR: the angular functions of a sector are ordered by l-parity class (first half / second
// half of the NP positions).  A coupling coefficient is non-zero only if l_j + l_i + L is even,
// so every 8-row tile couples to ONE parity class of the contracted index: both contractions
// run over NP/2 instead of NP.
template <int NT, int NCH, bool PAR>
__global__ void __launch_bounds__(256)
k_fold(BasisDev b, const FoldTask *__restrict__ tasks, const int *__restrict__ pixlist, const double *__restrict__ G,
       const double *__restrict__ Ppix, double *__restrict__ R, int pix_per_cta, int PB) {
  constexpr int NP = NT * 8, LD = NP + 4, NAB = NCH * NCH, NPH = NP / 2, NTH = NT / 2;
  extern __shared__ double sm[];
  const FoldTask t = tasks[blockIdx.y];
  double *sGj = sm, *sGk = sGj + NCH * NP * LD, *sP = sGk + NCH * NP * LD;
  double *sY = sP + 2 * PB * NP * LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int64_t gstride = (int64_t)NP * NP;
  const int pix0 = blockIdx.x * pix_per_cta;   // positions in the task's pixel list
  const int pix1 = min(pix0 + pix_per_cta, t.npix);
  if (pix0 >= pix1) return;
  const int *plist = pixlist + t.pix0;
  const double *Psrc = Ppix + (int64_t)t.spp * b.Npix * gstride;
  double *Rdst = R + (int64_t)t.rslot * NAB * b.Npix * gstride;
  auto prefetch = [&](int pg, int buf) {
    const int npx = min(PB, pix1 - pg);
    if (npx <= 0) return;
    for (int idx = tid; idx < npx * NP * (NP / 2); idx += blockDim.x) {
      const int s = idx / (NP * (NP / 2)), rem = idx % (NP * (NP / 2)), r = rem / (NP / 2), c2 = rem % (NP / 2);
      cp_async16(sP + ((buf * PB + s) * NP + r) * LD + 2 * c2, Psrc + (int64_t)plist[pg + s] * gstride + r * NP + 2 * c2, 16);
    }
  };
  prefetch(pix0, 0);
  cp_async_commit();
  {
    const double *gj = G + ((int64_t)t.spj * b.NL + t.L) * NCH * gstride;
    const double *gk = G + ((int64_t)t.spk * b.NL + t.L) * NCH * gstride;
    for (int idx = tid; idx < NCH * NP * NP; idx += blockDim.x) {
      const int ch = idx / (NP * NP), rem = idx % (NP * NP), r = rem / NP, c = rem % NP;
      sGj[(ch * NP + r) * LD + c] = gj[idx];
      sGk[(ch * NP + r) * LD + c] = gk[idx];
    }
  }
  const int lr = lane >> 2, lc = lane & 3;
  const int Lpar = t.L & 1;
  int buf = 0;
  for (int pg = pix0; pg < pix1; pg += PB, buf ^= 1) {
    const int npx = min(PB, pix1 - pg);
    cp_async_wait<0>();
    __syncthreads();  // P(pg) landed; previous group's stage 2 no longer reads sY; G is loaded
    prefetch(pg + PB, buf ^ 1);
    cp_async_commit();
    // stage 1: Yt_b[k][i] = sum_l Gk_b[k][l] P[i][l]   items: (slot, b, row tile of k)
    for (int item = warp; item < npx * NCH * NT; item += nwarp) {
      const int s = item / (NCH * NT), rem = item % (NCH * NT), bb = rem / NT, rt = rem % NT;
      const int k0 = PAR ? (((rt / NTH) ^ Lpar) * NPH) : 0;
      const double *A = sGk + (bb * NP + rt * 8 + lr) * LD + lc + k0;
      const double *B = sP + ((buf * PB + s) * NP + lr) * LD + lc + k0;
      double c[NT][2];
#pragma unroll
      for (int n = 0; n < NT; n++) c[n][0] = c[n][1] = 0.0;
#pragma unroll 2
      for (int kk = 0; kk < (PAR ? NPH : NP); kk += 4) {
        const double a = A[kk];
#pragma unroll
        for (int n = 0; n < NT; n++) dmma(c[n][0], c[n][1], a, B[n * 8 * LD + kk]);
      }
      double *Y = sY + ((s * NCH + bb) * NP + rt * 8 + lr) * LD + 2 * lc;
#pragma unroll
      for (int n = 0; n < NT; n++) {
        double2 v;
        v.x = c[n][0];
        v.y = c[n][1];
        *reinterpret_cast<double2 *>(Y + n * 8) = v;
      }
    }
    __syncthreads();
    // stage 2: R_ab[j][k] = fac sum_i Gj_a[j][i] Yt_b[k][i]   items: (slot, a, b, row tile of j)
    for (int item = warp; item < npx * NAB * NT; item += nwarp) {
      const int s = item / (NAB * NT), rem = item % (NAB * NT), ab = rem / NT, rt = rem % NT;
      const int aa = ab / NCH, bb = ab % NCH;
      const int k0 = PAR ? (((rt / NTH) ^ Lpar) * NPH) : 0;
      const double *A = sGj + (aa * NP + rt * 8 + lr) * LD + lc + k0;
      const double *B = sY + ((s * NCH + bb) * NP + lr) * LD + lc + k0;
      double c[NT][2];
#pragma unroll
      for (int n = 0; n < NT; n++) c[n][0] = c[n][1] = 0.0;
#pragma unroll 2
      for (int kk = 0; kk < (PAR ? NPH : NP); kk += 4) {
        const double a = A[kk];
#pragma unroll
        for (int n = 0; n < NT; n++) dmma(c[n][0], c[n][1], a, B[n * 8 * LD + kk]);
      }
      double *out = Rdst + ((int64_t)ab * b.Npix + plist[pg + s]) * gstride + (rt * 8 + lr) * t.ldk + 2 * lc;
#pragma unroll
      for (int n = 0; n < NT; n++) {
        if (n * 8 + 2 * lc >= t.ldk) continue;   // column past the sector (ldk is even: both halves or none)
        double2 v;
        v.x = t.fac * c[n][0];
        v.y = t.fac * c[n][1];
        *reinterpret_cast<double2 *>(out + n * 8) = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Exchange fold for sectors of <= 16 functions (NT <= 2, no parity ordering): one warp = one pixel,
// both contractions back to back in registers.
//   stage 1: Yt_b[k][i] = sum_l Gk_b[k][l] P[i][l]        (accumulator fragments c1)
//   stage 2: R_ab[j][k] = sum_i (fac Gj_a[j][i]) Yt_b[k][i]
// A DMMA accumulator fragment holds Yt[row lr][cols 2lc, 2lc+1] and a B fragment wants
// B[k = lc][n = lr]: the same row, so c1 IS the stage-2 B operand if the contraction index i is
// visited in the order (8n + 2lc + f) instead of (4q + lc) -- only the A operand (Gj, read from shared
// memory as one 16-byte load per two k-steps) has to follow that order.  No shared-memory round
// trip for Yt, no CTA barrier in the pixel loop: every warp runs its own cp.async double buffer of
// P tiles.  Shared memory: Gk[NCH][NP][NP+4], Gj[NCH][NP][LDJ], P[8 warps][2][NP][NP+4].
// ---------------------------------------------------------------------------
template <int NT, int NCH>
__global__ void __launch_bounds__(256)
k_fold_reg(BasisDev b, const FoldTask *__restrict__ tasks, const int *__restrict__ pixlist, const double *__restrict__ G,
           const double *__restrict__ Ppix, double *__restrict__ R, int pix_per_cta) {
  constexpr int NP = NT * 8, LDP = NP + 4, LDK = NP + 4, LDJ = (NP % 16 == 0) ? NP + 8 : NP + 16, NAB = NCH * NCH;
  extern __shared__ double sm[];
  double *sGk = sm, *sGj = sGk + NCH * NP * LDK, *sP = sGj + NCH * NP * LDJ;
  const FoldTask t = tasks[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  constexpr int64_t gstride = (int64_t)NP * NP;
  const int pix0 = blockIdx.x * pix_per_cta;   // positions in the task's pixel list
  const int pix1 = min(pix0 + pix_per_cta, t.npix);
  if (pix0 >= pix1) return;
  const int *plist = pixlist + t.pix0;
  const double *Psrc = Ppix + (int64_t)t.spp * b.Npix * gstride;
  double *Rdst = R + (int64_t)t.rslot * NAB * b.Npix * gstride;
  double *myP = sP + warp * 2 * NP * LDP;
  auto prefetch = [&](int pos, int buf) {
    if (pos < pix1) {
      const int64_t pix = plist[pos];
#pragma unroll
      for (int idx = lane; idx < NP * (NP / 2); idx += 32) {
        const int r = idx / (NP / 2), c2 = idx % (NP / 2);
        cp_async16(myP + (buf * NP + r) * LDP + 2 * c2, Psrc + pix * gstride + r * NP + 2 * c2, 16);
      }
    }
  };
  prefetch(pix0 + warp, 0);
  cp_async_commit();
  {
    const double *gj = G + ((int64_t)t.spj * b.NL + t.L) * NCH * gstride;
    const double *gk = G + ((int64_t)t.spk * b.NL + t.L) * NCH * gstride;
    for (int idx = tid; idx < NCH * NP * NP; idx += 256) {
      const int ch = idx / (NP * NP), rem = idx % (NP * NP), r = rem / NP, c = rem % NP;
      sGj[(ch * NP + r) * LDJ + c] = t.fac * gj[idx];   // prefactor folded into the table (no DMUL per output)
      sGk[(ch * NP + r) * LDK + c] = gk[idx];
    }
  }
  __syncthreads();
  // store pattern of a lane, fixed for the task: fragment (rtj, rtk) goes to row (rtj*8 + lr) * ldk, column rtk*8 + 2*lc
  const int st_off = lr * t.ldk + 2 * lc, st_row = 8 * t.ldk;
  bool st_ok[NT];
#pragma unroll
  for (int rtk = 0; rtk < NT; rtk++) st_ok[rtk] = rtk * 8 + 2 * lc < t.ldk;
  int buf = 0;
  for (int pos = pix0 + warp; pos < pix1; pos += 8, buf ^= 1) {
    __syncwarp();   // every lane is done reading the buffer the next prefetch overwrites
    prefetch(pos + 8, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const int pix = plist[pos];
    const double *P = myP + buf * NP * LDP + lr * LDP + lc;
    // ---- stage 1
    double c1[NCH][NT][NT][2];
#pragma unroll
    for (int q = 0; q < NCH * NT * NT * 2; q++) (&c1[0][0][0][0])[q] = 0.0;
#pragma unroll
    for (int kk = 0; kk < NP; kk += 4) {
      double bp[NT];
#pragma unroll
      for (int n = 0; n < NT; n++) bp[n] = P[n * 8 * LDP + kk];
#pragma unroll
      for (int bb = 0; bb < NCH; bb++)
#pragma unroll
        for (int rt = 0; rt < NT; rt++) {
          const double a = sGk[(bb * NP + rt * 8 + lr) * LDK + lc + kk];
#pragma unroll
          for (int n = 0; n < NT; n++) dmma(c1[bb][rt][n][0], c1[bb][rt][n][1], a, bp[n]);
        }
    }
    // ---- stage 2
#pragma unroll
    for (int aa = 0; aa < NCH; aa++) {
      double2 ga[NT][NT];   // Gj_a[j = rtj*8 + lr][i = 8*ni + 2*lc + {0,1}]
#pragma unroll
      for (int rtj = 0; rtj < NT; rtj++)
#pragma unroll
        for (int ni = 0; ni < NT; ni++)
          ga[rtj][ni] = *reinterpret_cast<const double2 *>(sGj + (aa * NP + rtj * 8 + lr) * LDJ + 8 * ni + 2 * lc);
#pragma unroll
      for (int bb = 0; bb < NCH; bb++) {
        double c2[NT][NT][2];
#pragma unroll
        for (int q = 0; q < NT * NT * 2; q++) (&c2[0][0][0])[q] = 0.0;
#pragma unroll
        for (int ni = 0; ni < NT; ni++)
#pragma unroll
          for (int rtj = 0; rtj < NT; rtj++)
#pragma unroll
            for (int rtk = 0; rtk < NT; rtk++) {
              dmma(c2[rtj][rtk][0], c2[rtj][rtk][1], ga[rtj][ni].x, c1[bb][rtk][ni][0]);
              dmma(c2[rtj][rtk][0], c2[rtj][rtk][1], ga[rtj][ni].y, c1[bb][rtk][ni][1]);
            }
        double *out = Rdst + ((int64_t)(aa * NCH + bb) * b.Npix + pix) * gstride + st_off;
#pragma unroll
        for (int rtj = 0; rtj < NT; rtj++)
#pragma unroll
          for (int rtk = 0; rtk < NT; rtk++)
            if (st_ok[rtk])   // columns past the sector are not stored (ldk is even)
              *reinterpret_cast<double2 *>(out + rtj * st_row + rtk * 8) =
                  make_double2(c2[rtj][rtk][0], c2[rtj][rtk][1]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Multi-entry FP64 tensor-core GEMM:  C[M x N] (+)= alpha * sum_e A_e[M x K] B_e[K x N]
//   A_e: row-major, row stride lda_e, K contiguous.
//   B_e: KCONTIG = false: row k of B_e lives at B_e + browoff[k], N contiguous
//        KCONTIG = true : B_e[n][k] row-major with stride ldb (K contiguous)
//   C  : row-major with stride ldc
// ---------------------------------------------------------------------------

struct GemmEntry {
  const double *A;
  const double *B;
  int64_t lda;
};

struct GemmItem {
  double *C;
  const int *browoff;  // KCONTIG=false only
  int M, N, K;
  int ent0, ent1;
  int accumulate;
  int64_t ldb;         // KCONTIG=true only
  int64_t ldc;
  double alpha;
};

template <int BM, int BN, int WGM, int WGN, bool KCONTIG>
__global__ void __launch_bounds__(WGM *WGN * 32)
k_gemm(const GemmItem *__restrict__ items, const GemmEntry *__restrict__ entries) {
  constexpr int BK = 16, NTHR = WGM * WGN * 32;
  constexpr int LDA_S = BK + 4;
  constexpr int LDB_S = KCONTIG ? (BK + 4) : (BN + 4);
  constexpr int WM = BM / WGM, WN = BN / WGN, TM = WM / 8, TN = WN / 8;
  constexpr int A_PER = BM * BK / NTHR, B_PER = BN * BK / NTHR;
  static_assert(BM * BK % NTHR == 0 && BN * BK % NTHR == 0, "tile/threads mismatch");
  __shared__ double As[BM * LDA_S];
  __shared__ double Bs[KCONTIG ? BN * LDB_S : BK * LDB_S];

  const GemmItem it = items[blockIdx.z];
  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  if (bm >= it.M || bn >= it.N) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / WGN) * WM, wn0 = (warp % WGN) * WN;
  const int lr = lane >> 2, lc = lane & 3;

  double acc[TM][TN][2];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  double ra[A_PER], rb[B_PER];
  const int nkc = (it.K + BK - 1) / BK;
  const int nsteps = (it.ent1 - it.ent0) * nkc;

  auto fetch = [&](int step) {
    const GemmEntry e = entries[it.ent0 + step / nkc];
    const int kc = (step % nkc) * BK;
#pragma unroll
    for (int u = 0; u < A_PER; u++) {
      const int idx = tid + u * NTHR, m = idx / BK, k = idx % BK;
      ra[u] = (bm + m < it.M && kc + k < it.K) ? __ldg(e.A + (int64_t)(bm + m) * e.lda + kc + k) : 0.0;
    }
    if (KCONTIG) {
#pragma unroll
      for (int u = 0; u < B_PER; u++) {
        const int idx = tid + u * NTHR, n = idx / BK, k = idx % BK;
        rb[u] = (bn + n < it.N && kc + k < it.K) ? __ldg(e.B + (int64_t)(bn + n) * it.ldb + kc + k) : 0.0;
      }
    } else {
#pragma unroll
      for (int u = 0; u < B_PER; u++) {
        const int idx = tid + u * NTHR, k = idx / BN, n = idx % BN;
        rb[u] = (bn + n < it.N && kc + k < it.K) ? __ldg(e.B + it.browoff[kc + k] + bn + n) : 0.0;
      }
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < A_PER; u++) {
      const int idx = tid + u * NTHR, m = idx / BK, k = idx % BK;
      As[m * LDA_S + k] = ra[u];
    }
    if (KCONTIG) {
#pragma unroll
      for (int u = 0; u < B_PER; u++) {
        const int idx = tid + u * NTHR, n = idx / BK, k = idx % BK;
        Bs[n * LDB_S + k] = rb[u];
      }
    } else {
#pragma unroll
      for (int u = 0; u < B_PER; u++) {
        const int idx = tid + u * NTHR, k = idx / BN, n = idx % BN;
        Bs[k * LDB_S + n] = rb[u];
      }
    }
  };

  if (nsteps > 0) fetch(0);
  for (int step = 0; step < nsteps; step++) {
    __syncthreads();
    stash();
    __syncthreads();
    if (step + 1 < nsteps) fetch(step + 1);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[TM], bf[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) af[i] = As[(wm0 + i * 8 + lr) * LDA_S + kk + lc];
#pragma unroll
      for (int j = 0; j < TN; j++)
        bf[j] = KCONTIG ? Bs[(wn0 + j * 8 + lr) * LDB_S + kk + lc] : Bs[(kk + lc) * LDB_S + wn0 + j * 8 + lr];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
#pragma unroll
  for (int i = 0; i < TM; i++) {
    const int m = bm + wm0 + i * 8 + lr;
    if (m >= it.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const int n = bn + wn0 + j * 8 + 2 * lc;
      double *c = it.C + (int64_t)m * it.ldc + n;
      if (n < it.N) c[0] = it.alpha * acc[i][j][0] + (it.accumulate ? c[0] : 0.0);
      if (n + 1 < it.N) c[1] = it.alpha * acc[i][j][1] + (it.accumulate ? c[1] : 0.0);
    }
  }
}

// ---------------------------------------------------------------------------
// Storage of one dense exchange-ordered in-element kernel A[(rj,rk)][K] (K = nab*n*n):
// k-chunks of TP_BK columns, each chunk a dense [n*n rows][TP_BK] tile that is the exact
// shared-memory image the GEMM wants (one bulk copy per tile).  Within a row the 4-double groups
// are XOR-swizzled with (row & 7) so that the DMMA A-fragment reads (8 rows x 4 consecutive k)
// hit every bank exactly twice (= the 2-wavefront minimum for 256 bytes).  Columns >= K are zero.
// ---------------------------------------------------------------------------
constexpr int TP_BK = 32;       // general layout (all M = Ni * Nj rows), swizzled
#ifndef HFQ_TP_BK_TRI
#define HFQ_TP_BK_TRI 36
#endif
constexpr int TP_BK_TRI = HFQ_TP_BK_TRI;   // symmetric-density layout (rows rj <= rk): K = nab * 15^2 = 900 = 25 chunks, no K padding;
                                // the row stride of 36 doubles = 4 banks already spreads the 8 fragment rows over all
                                // 32 banks, so this layout is not swizzled
__host__ __device__ inline int64_t tperm_doubles(int nn, int K, int bk = TP_BK) { return (int64_t)((K + bk - 1) / bk) * nn * bk; }
__host__ __device__ inline int64_t tperm_index(int row, int col, int nn, int bk = TP_BK) {
  const int c = col % bk;
  return (int64_t)(col / bk) * nn * bk + (int64_t)row * bk + (bk == TP_BK ? (c ^ (4 * (row & 7))) : c);
}

// ---------------------------------------------------------------------------
// Setup: dense exchange-ordered in-element kernel from the low-rank factor
//   T[(rj*n + rk)][ab][(ri*n + rl)] = s_ab sum_p sigma_p B[a*nn + pair1, p] B[b*nn + rk + rl*n, p]
//   pair1 = out_fast ? rj + ri*n : ri + rj*n
// grid (rows); stored at tperm_index(row, ab*nn + ri*n + rl, rows)
// ---------------------------------------------------------------------------
static __global__ void k_build_tperm(const double *__restrict__ B, const double *__restrict__ sigma, int n, int rank, int nch,
                              int out_fast, int tri, int bk, double *__restrict__ dst) {
  // tri: rows are the pairs rj <= rk only, row t = rk (rk + 1) / 2 + rj (symmetric densities)
  const int row = blockIdx.x, nn = n * n;
  int rj, rk;
  if (tri) {
    rk = 0;
    while ((rk + 1) * (rk + 2) / 2 <= row) rk++;
    rj = row - rk * (rk + 1) / 2;
  } else {
    rj = row / n;
    rk = row % n;
  }
  const int nrows = tri ? n * (n + 1) / 2 : nn;
  for (int col = threadIdx.x; col < nch * nch * nn; col += blockDim.x) {
    const int ab = col / nn, il = col % nn, a = ab / nch, bb = ab % nch, ri = il / n, rl = il % n;
    const int p1 = out_fast ? (rj + ri * n) : (ri + rj * n);
    const double *b1 = B + a * nn + p1, *b2 = B + bb * nn + rk + rl * n;
    const int64_t ldB = (int64_t)nch * nn;
    double s = 0.0;
    for (int p = 0; p < rank; p++) s += sigma[p] * b1[p * ldB] * b2[p * ldB];
    const double sgn = (nch == 2 && a != bb) ? -1.0 : 1.0;
    dst[tperm_index(row, col, nrows, bk)] = sgn * s;
  }
}

// ---------------------------------------------------------------------------
// Exchange, cross-element pairs:  K_(ei,ej)(rj,rk) += sum_ab sum_(ri,rl) I_a(rj,ri) R_ab(ri,rl) J_b(rk,rl)
// Work descriptors of k_offdiag_mma: item = output pair x element pair, entries = (R slot, channel).
// ---------------------------------------------------------------------------
struct OffItem {
  double *C;        // [Ni*Nj][NB]
  int ei, ej;
  int ent0, ent1;   // entries: (R slot, channel)
  int accumulate;
  int ncol;         // blk columns >= ncol are padding (multiple of 16)
};
struct OffEntry {
  int rslot, ilm;
};

// ---------------------------------------------------------------------------
// Exchange, cross-element pairs on the FP64 tensor pipe.
//   K_(ei,ej)(rj,rk)[blk] += sum_a sum_ri I_a(rj,ri) U_a(ri,rk)[blk]
//   U_a(ri,rk)[blk]        = sum_b sum_rl J_b(rk,rl) R_ab(ri,rl)[blk]
// Element sizes are padded 15 -> 16 with zero rows/columns of I and J.
// One CTA (8 warps) = (item, tile of 16 blk columns).  Stage 1 streams the R rows
// straight from global memory into DMMA B fragments (every R element is used by
// exactly one CTA), stage 2 keeps the 16x16x16 output tile of two rk values per
// warp in registers across the whole entry loop.
// Shared memory: U[NCH][16 rk][16 ri x LDU (+4)] + I[NCH][16][LDS] + J[NCH][16][LDJ]
// ---------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256, 2)
k_offdiag_mma(BasisDev b, const OffItem *__restrict__ items, const OffEntry *__restrict__ entries,
              const double *__restrict__ R, const double *__restrict__ dsmall, const double *__restrict__ dbig,
              const int64_t *__restrict__ blk_off) {
  constexpr int BT = 16, LDU = BT + 4, LDK = 16 * LDU + 4, LDI = 16 + 4, LDJ = NCH * 16 + 4, NAB = NCH * NCH;
  constexpr int NCHUNK = 2 * NAB;   // chunks (ri slot, a, b) of one warp per entry
  extern __shared__ double sm[];
  double *sU = sm;                          // [NCH][16 rk][LDK: 16 ri x LDU]
  double *sI = sU + NCH * 16 * LDK;         // [NCH][16 rj][LDI]
  double *sJ = sI + NCH * 16 * LDI;         // J[rk][b*16 + rl]
  const OffItem it = items[blockIdx.y];
  const int Ni = b.en[it.ei], Nj = b.en[it.ej], fi = b.efirst[it.ei], fj = b.efirst[it.ej];
  const int blk0 = blockIdx.x * BT;
  if (blk0 >= it.ncol) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int64_t gstride = (int64_t)b.NB;
  for (int idx = tid; idx < NCH * 16 * LDK; idx += 256) sU[idx] = 0.0;
  for (int idx = tid; idx < NCH * 16 * LDI; idx += 256) sI[idx] = 0.0;
  for (int idx = tid; idx < 16 * LDJ; idx += 256) sJ[idx] = 0.0;
  double kacc[2][2][2][2];   // [rk slot][rj tile][blk tile][frag]
#pragma unroll
  for (int q = 0; q < 16; q++) (&kacc[0][0][0][0])[q] = 0.0;

  // Stage-1 operand stream.  Chunk q = (slot, a, bb) of entry e is the 16 (rl) x 16 (blk) block
  // R_ab(ri = warp + 8*slot, :)[blk0..+16).  A lane holds, for k-step ks, the two columns
  // (2*lr, 2*lr+1) of row rl = 4*ks + lc as one 16-byte load: DMMA column tile nt therefore maps to
  // the blk columns 2*n + nt.  The next chunk is fetched into registers while the current one
  // feeds the tensor pipe (nothing else hides the HBM latency of this stream; an additional
  // prefetch.global.L2 two chunks ahead was measured and did not help: 5.57 vs 5.31 ms; neither did a variant with
  // one 16-warp CTA per SM and a 4-deep per-warp cp.async ring in shared memory, 96 KB in flight per SM:
  // 6.38 vs 5.59 ms, profiles/r02i -- the CTA-wide barriers between the two stages cost more than the latency).
  auto fetch = [&](int e, int q, double2 (&pf)[4]) {
    const int slot = q / NAB, ab = q % NAB, ri = warp + 8 * slot;
    const bool ok = e < it.ent1 && ri < Ni;
    const double *row = R + ((int64_t)(ok ? entries[e].rslot : 0) * NAB * b.Npix + (int64_t)ab * b.Npix +
                             (int64_t)(fi + (ok ? ri : 0)) * b.Nrad + fj) * gstride + blk0 + 2 * lr;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      const int rl = ks * 4 + lc;
      pf[ks] = (ok && rl < Nj) ? __ldg(reinterpret_cast<const double2 *>(row + (int64_t)rl * gstride)) : make_double2(0.0, 0.0);
    }
  };
  double2 pf[4];
  fetch(it.ent0, 0, pf);

  for (int e = it.ent0; e < it.ent1; e++) {
    const OffEntry en = entries[e];
    const double *srcI = (it.ei > it.ej ? dbig : dsmall) + blk_off[en.ilm * b.Nel + it.ei];
    const double *srcJ = (it.ei > it.ej ? dsmall : dbig) + blk_off[en.ilm * b.Nel + it.ej];
    __syncthreads();   // previous stage 2 finished with sI/sU
    for (int idx = tid; idx < NCH * Ni * Ni; idx += 256) {
      const int ch = idx / (Ni * Ni), rem = idx % (Ni * Ni), ri = rem / Ni, rj = rem % Ni;   // I(rj,ri) column-major
      sI[(ch * 16 + rj) * LDI + ri] = srcI[idx];
    }
    for (int idx = tid; idx < NCH * Nj * Nj; idx += 256) {
      const int ch = idx / (Nj * Nj), rem = idx % (Nj * Nj), rl = rem / Nj, rk = rem % Nj;   // J(rk,rl) column-major
      sJ[rk * LDJ + ch * 16 + rl] = srcJ[idx];
    }
    __syncthreads();
    // ---- stage 1: warp -> ri = warp, warp + 8
#pragma unroll
    for (int slot = 0; slot < 2; slot++) {
      const int ri = warp + 8 * slot;
#pragma unroll
      for (int a = 0; a < NCH; a++) {
        double c[2][2][2];
#pragma unroll
        for (int q = 0; q < 8; q++) (&c[0][0][0])[q] = 0.0;
#pragma unroll
        for (int bb = 0; bb < NCH; bb++) {
          double2 cur[4];
#pragma unroll
          for (int ks = 0; ks < 4; ks++) cur[ks] = pf[ks];
          const int q = (slot * NCH + a) * NCH + bb + 1;
          if (q < NCHUNK)
            fetch(e, q, pf);
          else
            fetch(e + 1, 0, pf);
          if (ri < Ni) {
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {
              const int rl = ks * 4 + lc;
              const double a0 = sJ[lr * LDJ + bb * 16 + rl], a1 = sJ[(8 + lr) * LDJ + bb * 16 + rl];
              dmma(c[0][0][0], c[0][0][1], a0, cur[ks].x);
              dmma(c[0][1][0], c[0][1][1], a0, cur[ks].y);
              dmma(c[1][0][0], c[1][0][1], a1, cur[ks].x);
              dmma(c[1][1][0], c[1][1][1], a1, cur[ks].y);
            }
          }
        }
        if (ri < Ni) {
          // fragment (row lr, tile columns 2lc, 2lc+1 of tiles 0/1) = blk columns 4lc .. 4lc+3
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            double *u = sU + (a * 16 + mt * 8 + lr) * LDK + ri * LDU + 4 * lc;
            *reinterpret_cast<double2 *>(u) = make_double2(c[mt][0][0], c[mt][1][0]);
            *reinterpret_cast<double2 *>(u + 2) = make_double2(c[mt][0][1], c[mt][1][1]);
          }
        }
      }
    }
    __syncthreads();
    // ---- stage 2: warp -> rk = warp, warp + 8
#pragma unroll
    for (int slot = 0; slot < 2; slot++) {
      const int rk = warp + slot * 8;
      if (rk >= Nj) continue;
#pragma unroll
      for (int a = 0; a < NCH; a++) {
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const int ri = ks * 4 + lc;
          const double a0 = sI[(a * 16 + lr) * LDI + ri], a1 = sI[(a * 16 + 8 + lr) * LDI + ri];
          const double *u = sU + (a * 16 + rk) * LDK + ri * LDU + lr;
          const double bf0 = u[0], bf1 = u[8];
          dmma(kacc[slot][0][0][0], kacc[slot][0][0][1], a0, bf0);
          dmma(kacc[slot][0][1][0], kacc[slot][0][1][1], a0, bf1);
          dmma(kacc[slot][1][0][0], kacc[slot][1][0][1], a1, bf0);
          dmma(kacc[slot][1][1][0], kacc[slot][1][1][1], a1, bf1);
        }
      }
    }
  }
#pragma unroll
  for (int slot = 0; slot < 2; slot++) {
    const int rk = warp + slot * 8;
    if (rk >= Nj) continue;
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      const int rj = mt * 8 + lr;
      if (rj >= Ni) continue;
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        double *c = it.C + (int64_t)(rj * Nj + rk) * gstride + blk0 + nt * 8 + 2 * lc;
        double2 v;
        v.x = kacc[slot][mt][nt][0];
        v.y = kacc[slot][mt][nt][1];
        if (it.accumulate) {
          const double2 o = *reinterpret_cast<const double2 *>(c);
          v.x += o.x;
          v.y += o.y;
        }
        *reinterpret_cast<double2 *>(c) = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// mbarrier / bulk-copy helpers (sm_90+ PTX; SASS UBLKCP + SYNCS)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------
// In-element exchange GEMM, warp-specialised:
//   C[M x 64-column tile] (+)= sum_entries A_e[M x K] R_e[K x N],  M = Ni^2 <= 256, K = nab*Ni^2
// A_e is stored as pre-swizzled [k-chunk][M][TP_BK] tiles (tperm_index), so the producer warp moves
// a whole A stage with ONE cp.async.bulk (<= 64 KB) plus one 512-byte bulk copy per R row, all
// counted on the stage's "full" mbarrier; the 8 DMMA warps wait on it, run their k-steps and
// release the stage through an "empty" mbarrier.  No CTA-wide barrier in the main loop, so the
// FP64 tensor pipe is fed by whichever warps hold data while others wait.  Row tiles past M read
// stale shared memory; they only feed accumulator rows >= M, which are never stored.  A columns
// >= K are zero in the tiles and R rows past K come from a zero row.  it.N (a multiple of 8) may be
// smaller than the padded column count: column tiles past it are skipped.  NS = stages that fit
// (3 for the 15-node elements).
// ---------------------------------------------------------------------------
// The consumer loop of one warp, specialised at compile time on its tile counts (the dispatch happens once per CTA,
// outside the stage loop, so that every variant keeps its accumulators in registers).
//   BAL:  the warp owns 4 column tiles: NF full row tiles + REM left-over units
//   !BAL: the warp owns w.ncj < 4 column tiles and computes NCJ = 2 or 4 of them (columns past w.ncj are padding and
//         are not stored; REM unused): NF row tiles rg, rg + 4, ..; NF = 0 also serves idle warps
struct TgemmWarp {
  const double *As, *Bs;     // stage 0 of the ring
  uint64_t *full, *empty;
  int NS, A_STAGE, B_STAGE, nsteps;
  int lr, lc, cg, rg, lane;
  int nrt_tot, ncj;
  int rs;                    // row-tile stride of the !BAL split: 4 (row groups of one column group) or 8 (all warps)
};

template <int BK, int NF, int REM, bool BAL, int NCJ = 4>
__device__ __forceinline__ void tgemm_ws_consume(const TgemmWarp &w, const GemmItem &it, int bn) {
  constexpr int LDB_S = 68, NFF = NF > 0 ? NF : 1;
  constexpr bool SWZ = BK == TP_BK;   // A fragment of k-step ks sits at row*BK + lc + 4*ks, ks XOR lr in the swizzled layout
  double acc[NFF][4][2], ex[4][2];
#pragma unroll
  for (int i = 0; i < NFF; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
  for (int j = 0; j < 4; j++) ex[j][0] = ex[j][1] = 0.0;
  int bcol[4];
#pragma unroll
  for (int j = 0; j < 4; j++) bcol[j] = BAL ? ((j + w.rg) & 3) * 8 : j * 8;
  const int xrow = (w.nrt_tot & ~3) * 8 + w.lr;   // first left-over row tile
  int stage = 0, phase = 0;
  for (int step = 0; step < w.nsteps; step++) {
    mbar_wait(w.full + stage, phase);
    if (NF > 0 || REM > 0) {
      const double *as = w.As + (size_t)stage * w.A_STAGE + (w.rg * 8 + w.lr) * BK + w.lc;
      const double *bs = w.Bs + stage * w.B_STAGE + w.lc * LDB_S + w.cg * 32 + w.lr;
      if (BAL) {
        const double *ax = w.As + (size_t)stage * w.A_STAGE + xrow * BK + w.lc;
        double bf[2][4], af[2][NFF], ae[2][REM > 0 ? REM : 1];
#pragma unroll
        for (int j = 0; j < 4; j++) bf[0][j] = bs[bcol[j]];
#pragma unroll
        for (int i = 0; i < NF; i++) af[0][i] = as[i * 32 * BK + (SWZ ? 4 * w.lr : 0)];
#pragma unroll
        for (int e = 0; e < REM; e++) ae[0][e] = ax[e * 8 * BK + (SWZ ? 4 * w.lr : 0)];
#pragma unroll
        for (int ks = 0; ks < BK / 4; ks++) {
          const int cur = ks & 1, nxt = cur ^ 1;
          if (ks + 1 < BK / 4) {
            const int ko = SWZ ? 4 * ((ks + 1) ^ w.lr) : 4 * (ks + 1);
#pragma unroll
            for (int j = 0; j < 4; j++) bf[nxt][j] = bs[(ks + 1) * 4 * LDB_S + bcol[j]];
#pragma unroll
            for (int i = 0; i < NF; i++) af[nxt][i] = as[i * 32 * BK + ko];
#pragma unroll
            for (int e = 0; e < REM; e++) ae[nxt][e] = ax[e * 8 * BK + ko];
          }
#pragma unroll
          for (int i = 0; i < NF; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
#pragma unroll
          for (int e = 0; e < REM; e++) dmma(ex[e][0], ex[e][1], ae[cur][e], bf[cur][e]);
        }
      } else {
        double bf[2][NCJ], af[2][NFF];
        const int rstep = w.rs * 8 * BK;
#pragma unroll
        for (int j = 0; j < NCJ; j++) bf[0][j] = bs[j * 8];
#pragma unroll
        for (int i = 0; i < NF; i++) af[0][i] = as[i * rstep + (SWZ ? 4 * w.lr : 0)];
#pragma unroll
        for (int ks = 0; ks < BK / 4; ks++) {
          const int cur = ks & 1, nxt = cur ^ 1;
          if (ks + 1 < BK / 4) {
            const int ko = SWZ ? 4 * ((ks + 1) ^ w.lr) : 4 * (ks + 1);
#pragma unroll
            for (int j = 0; j < NCJ; j++) bf[nxt][j] = bs[(ks + 1) * 4 * LDB_S + j * 8];
#pragma unroll
            for (int i = 0; i < NF; i++) af[nxt][i] = as[i * rstep + ko];
          }
#pragma unroll
          for (int i = 0; i < NF; i++)
#pragma unroll
            for (int j = 0; j < NCJ; j++) dmma(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
        }
      }
    }
    __syncwarp();
    if (w.lane == 0) mbar_arrive(w.empty + stage);
    if (++stage == w.NS) {
      stage = 0;
      phase ^= 1;
    }
  }
  auto store = [&](int m, int col, double v0, double v1) {
    double *c = it.C + (int64_t)m * it.ldc + bn + w.cg * 32 + col + 2 * w.lc;
    double2 v;
    v.x = it.alpha * v0;
    v.y = it.alpha * v1;
    if (it.accumulate) {
      const double2 o = *reinterpret_cast<const double2 *>(c);
      v.x += o.x;
      v.y += o.y;
    }
    *reinterpret_cast<double2 *>(c) = v;
  };
#pragma unroll
  for (int i = 0; i < NF; i++) {
    const int m = (w.rg + (BAL ? 4 : w.rs) * i) * 8 + w.lr;
    if (m >= it.M) continue;
#pragma unroll
    for (int j = 0; j < NCJ; j++)
      if (BAL || j < w.ncj) store(m, bcol[j], acc[i][j][0], acc[i][j][1]);
  }
#pragma unroll
  for (int e = 0; e < REM; e++) {
    const int m = xrow - w.lr + e * 8 + w.lr;
    if (m < it.M) store(m, bcol[e], ex[e][0], ex[e][1]);
  }
}

template <int BK, int REM, bool BAL, int NCJ = 4>
__device__ __forceinline__ void tgemm_ws_dispatch_nf(const TgemmWarp &w, const GemmItem &it, int bn, int nf) {
  switch (nf) {
    case 8: tgemm_ws_consume<BK, 8, REM, BAL, NCJ>(w, it, bn); break;
    case 7: tgemm_ws_consume<BK, 7, REM, BAL, NCJ>(w, it, bn); break;
    case 6: tgemm_ws_consume<BK, 6, REM, BAL, NCJ>(w, it, bn); break;
    case 5: tgemm_ws_consume<BK, 5, REM, BAL, NCJ>(w, it, bn); break;
    case 4: tgemm_ws_consume<BK, 4, REM, BAL, NCJ>(w, it, bn); break;
    case 3: tgemm_ws_consume<BK, 3, REM, BAL, NCJ>(w, it, bn); break;
    case 2: tgemm_ws_consume<BK, 2, REM, BAL, NCJ>(w, it, bn); break;
    case 1: tgemm_ws_consume<BK, 1, REM, BAL, NCJ>(w, it, bn); break;
    default: tgemm_ws_consume<BK, 0, REM, BAL, NCJ>(w, it, bn); break;
  }
}

__host__ __device__ inline size_t tgemm_ws_smem(int M, int ns, int bk = TP_BK) {
  // ns stages of (A tile [M][TP_BK] + R tile [TP_BK][68]); 8 slack rows so that the last row tile (rows up to
  // round_up(M, 8), the furthest any consumer reads) of the last stage stays inside the allocation; 2*ns mbarriers
  return ((size_t)ns * ((size_t)M * bk + bk * 68) + (size_t)8 * bk) * sizeof(double) + 16 * ns;
}

template <int BK>
__global__ void __launch_bounds__(288, 1)
k_tgemm_ws(const GemmItem *__restrict__ items, const GemmEntry *__restrict__ entries, const double *__restrict__ zrow,
           int NS, int maxM) {
  constexpr int BN = 64, NW = 8;
  constexpr int LDB_S = BN + 4, B_STAGE = BK * LDB_S;
  extern __shared__ double sm[];
  const GemmItem it = items[blockIdx.y];
  const int A_STAGE = maxM * BK;   // stage stride (largest M of the launch); rows >= it.M hold stale data
  double *Bs = sm, *As = sm + NS * B_STAGE;
  uint64_t *full = reinterpret_cast<uint64_t *>(As + (size_t)NS * A_STAGE + 8 * BK), *empty = full + NS;
  const int bn = blockIdx.x * BN;
  if (bn >= it.N) return;   // column tile holds padding only
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nkc = (it.K + BK - 1) / BK;
  const int nsteps = (it.ent1 - it.ent0) * nkc;

  if (tid == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (warp == NW) {
    // ---- producer ----
    const uint32_t abytes = (uint32_t)(it.M * BK * 8);
    int stage = 0, phase = 1, ent = it.ent0, kci = 0;
    GemmEntry e = entries[ent];
    for (int step = 0; step < nsteps; step++) {
      mbar_wait(empty + stage, phase);
      const int kc = kci * BK;
      if (lane == 0) {
        mbar_expect_tx(full + stage, abytes + BK * BN * 8);
        bulk_g2s(As + (size_t)stage * A_STAGE, e.A + (int64_t)kci * it.M * BK, abytes, full + stage);
      }
      __syncwarp();
      for (int r = lane; r < BK; r += 32) {   // one 512-byte bulk copy per R row of the stage
        const double *src = (kc + r < it.K) ? e.B + it.browoff[kc + r] + bn : zrow;
        bulk_g2s(Bs + stage * B_STAGE + r * LDB_S, src, BN * 8, full + stage);
      }
      if (++kci == nkc) {
        kci = 0;
        if (++ent < it.ent1) e = entries[ent];
      }
      if (++stage == NS) {
        stage = 0;
        phase ^= 1;
      }
    }
    return;
  }

  // ---- consumers ----
  // Work split: warp w owns the column tiles 4*(w/4) .. +3 (8 columns each) and the row group rg = w & 3.  No
  // DMMA is issued for row tiles past M or for column tiles past N.  Warps with all 4 column tiles (the usual
  // case) take the balanced split (tgemm_ws_consume<.., BAL = true>): the row tiles are dealt in full rounds and
  // the (row tiles mod 4) left-over tiles are cut by column tile, so that every warp issues the same number of
  // DMMAs per k-step (M = 120: 15 each instead of 16/16/16/12).  A warp with 2 column tiles (last tile of a sector
  // pair whose column count is not a multiple of 64) takes every 4th row tile; one with none only keeps the ring going.
  TgemmWarp w;
  w.As = As;
  w.Bs = Bs;
  w.full = full;
  w.empty = empty;
  w.NS = NS;
  w.A_STAGE = A_STAGE;
  w.B_STAGE = B_STAGE;
  w.nsteps = nsteps;
  w.lane = lane;
  w.lr = lane >> 2;
  w.lc = lane & 3;
  w.cg = warp >> 2;
  w.rg = warp & 3;
  w.nrt_tot = (it.M + 7) >> 3;
  w.ncj = min(4, ((it.N - bn) >> 3) - 4 * w.cg);   // column tiles of this warp (<= 0: idle)
  w.rs = 4;
  const int ctiles = min(8, (it.N - bn) >> 3);     // column tiles of this CTA
  if (ctiles <= 4) {
    // the second column group would idle: all 8 warps share the first one, every 8th row tile each
    w.cg = 0;
    w.rg = warp;
    w.rs = 8;
    w.ncj = ctiles;
    const int nf = (w.nrt_tot - w.rg + 7) >> 3;
    if (ctiles > 2)
      tgemm_ws_dispatch_nf<BK, 0, false, 4>(w, it, bn, nf);
    else
      tgemm_ws_dispatch_nf<BK, 0, false, 2>(w, it, bn, nf);
  } else if (w.ncj == 4) {
    const int nf = w.nrt_tot >> 2;
    switch (w.nrt_tot & 3) {
      case 3: tgemm_ws_dispatch_nf<BK, 3, true>(w, it, bn, nf); break;
      case 2: tgemm_ws_dispatch_nf<BK, 2, true>(w, it, bn, nf); break;
      case 1: tgemm_ws_dispatch_nf<BK, 1, true>(w, it, bn, nf); break;
      default: tgemm_ws_dispatch_nf<BK, 0, true>(w, it, bn, nf); break;
    }
  } else if (w.ncj > 2) {
    tgemm_ws_dispatch_nf<BK, 0, false, 4>(w, it, bn, (w.nrt_tot - w.rg + 3) >> 2);
  } else if (w.ncj > 0) {
    tgemm_ws_dispatch_nf<BK, 0, false, 2>(w, it, bn, (w.nrt_tot - w.rg + 3) >> 2);
  } else {
    tgemm_ws_consume<BK, 0, 0, false>(w, it, bn);
  }
}

// ---------------------------------------------------------------------------
// Exchange unpack: dense K = - sum over element-pair blocks, boundary removal,
// optional +-m mirror (mirror_op maps an output sector pair to the pair that
// was actually computed).
// grid (dense column tiles), one thread per dense element
// ---------------------------------------------------------------------------
struct UnpackDev {
  const int *op_src;        // [ns*ns] active output pair whose units hold this block (or -1: zero)
  const int *op_tri;        // [active op] 1: symmetric storage (see below)
  const int *op_ldk;        // [active op] column index of block (pos_j, pos_k) = pos_j * ldk + pos_k (FoldTask::ldk)
  const int *blocks;        // angular blocks (j | k << 16) whose sector pair was computed
  const int64_t *unit_off;  // [(active op * Nel + ei) * Nel + ej] offset of the unit's [rows][NB] block in Kc, -1: absent
                            // (built by another rank and not gathered: the caller sums the partial matrices)
  const int *ang_sec;       // [Nang] sector of angular function
  const int *ang_pos;       // [Nang] position inside the sector
  double scale;             // exchange(scale * P) = scale * exchange(P)
};

// Sum of the K-split partial accumulators of the units this rank built into the compact unit buffer:
// Kc[dst + i] += sum_c part[src + c * cstride + i].  One descriptor per (unit, 16K-element slab).
struct ReduceDesc {
  int64_t dst, src;   // offsets in doubles into Kc / the partial buffer
  int n;              // doubles
  int nparts;         // partial buffers this unit wrote (in-element units: S - 1, cross-element units: S_off - 1)
};
static __global__ void k_reduce_partials(const ReduceDesc *__restrict__ descs, double *__restrict__ Kc,
                                         const double *__restrict__ part, int64_t cstride) {
  const ReduceDesc d = descs[blockIdx.x];
  for (int i = threadIdx.x; i < d.n; i += blockDim.x) {
    double s = Kc[d.dst + i];
    for (int c = 0; c < d.nparts; c++) s += part[d.src + (int64_t)c * cstride + i];
    Kc[d.dst + i] = s;
  }
}

// Symmetric storage (diagonal output pairs of a symmetric density): K(j rj, k rk) = K(k rk, j rj), so
// only element pairs ei <= ej were computed, and of an in-element block only the rows rj <= rk
// (row t = rk (rk+1)/2 + rj); the other half is read from the mirrored entry, column (pos_k, pos_j).
// One CTA per angular block that can be non-zero (u.blocks[] = angj | angk << 16); the caller zero-fills K first.
static __global__ void k_unpack_K(BasisDev b, UnpackDev u, const double *__restrict__ Kc, double *__restrict__ K, int64_t ld) {
  const int angj = u.blocks[blockIdx.x] & 0xffff, angk = u.blocks[blockIdx.x] >> 16;
  const int sj = b.ang_skip[angj], sk = b.ang_skip[angk];
  const int nj = b.Nrad - sj, nk = b.Nrad - sk;
  double *dst = K + b.ang_off[angj] + (int64_t)b.ang_off[angk] * ld;
  const int op = u.ang_sec[angj] * b.ns + u.ang_sec[angk];
  const int src = u.op_src[op];
  const int ldk = src >= 0 ? u.op_ldk[src] : b.NP;
  const int blk = u.ang_pos[angj] * ldk + u.ang_pos[angk];
  const int blkT = u.ang_pos[angk] * ldk + u.ang_pos[angj];
  const bool tri = src >= 0 && u.op_tri[src];
  const int64_t *uo = u.unit_off + (int64_t)(src < 0 ? 0 : src) * b.Nel * b.Nel;
  for (int idx = threadIdx.x; idx < nj * nk; idx += blockDim.x) {
    const int r = idx % nj + sj, c = idx / nj + sk;
    double s = 0.0;
    if (src >= 0) {
      for (int ei = b.rad_e0[r]; ei <= b.rad_e1[r]; ei++) {
        const int ri = r - b.efirst[ei];
        for (int ej = b.rad_e0[c]; ej <= b.rad_e1[c]; ej++) {
          const int rk = c - b.efirst[ej];
          int64_t base, off;
          if (!tri || ei < ej) {
            base = uo[ei * b.Nel + ej];
            off = (int64_t)(ri * b.en[ej] + rk) * b.NB + blk;
          } else if (ei > ej) {
            base = uo[ej * b.Nel + ei];
            off = (int64_t)(rk * b.en[ei] + ri) * b.NB + blkT;
          } else if (ri <= rk) {
            base = uo[ei * b.Nel + ei];
            off = (int64_t)(rk * (rk + 1) / 2 + ri) * b.NB + blk;
          } else {
            base = uo[ei * b.Nel + ei];
            off = (int64_t)(ri * (ri + 1) / 2 + rk) * b.NB + blkT;
          }
          if (base >= 0) s += Kc[base + off];
        }
      }
    }
    dst[(r - sj) + (int64_t)(c - sk) * ld] = -u.scale * s;
  }
}

// ---------------------------------------------------------------------------
// Coulomb: radial step per (L, M) channel.
//   in : Paux[Midx][q = L*nch + ch][pix]      (fold output)
//   out: JauxT[Midx][pix][q]
// One CTA per (L, Midx).
// ---------------------------------------------------------------------------
struct JRadDev {
  const int *chan_of;        // [NL * nM] channel index (ilm) or -1
  const double *fac;         // [NL * nM] prefactor incl. sign
  const int64_t *blk_off;    // [nlm*Nel] offsets into dsmall/dbig
  const int64_t *B_off;      // [nlm*Nel] offsets into dB
  const int64_t *sig_off;    // [nlm*Nel] offsets into dsigma
  const int *rank;           // [nlm*Nel]
  const double *dsmall, *dbig, *dB, *dsigma;
  int nM;
};

static __global__ void __launch_bounds__(256)
k_jradial(BasisDev b, JRadDev jr, int L0, const double *__restrict__ Paux, double *__restrict__ JauxT) {
  const int L = L0 + blockIdx.x, Mi = blockIdx.y, tid = threadIdx.x;
  const int nq = b.NL * b.nch;
  const int ilm = jr.chan_of[L * jr.nM + Mi];
  double *out = JauxT + (int64_t)Mi * b.Npix * nq;
  for (int ch = 0; ch < b.nch; ch++)
    for (int pix = tid; pix < b.Npix; pix += blockDim.x) out[(int64_t)pix * nq + L * b.nch + ch] = 0.0;
  if (ilm < 0) return;
  const double fac = jr.fac[L * jr.nM + Mi];
  const double *pin = Paux + ((int64_t)Mi * nq + L * b.nch) * b.Npix;  // [ch][pix]
  __shared__ double s_js[64], s_jb[64], s_c[128], red[2][8];
  __shared__ double s_pre[64], s_post[64];
  const int Nel = b.Nel;
  // traces with the cross-element factors
  for (int e = 0; e < Nel; e++) {
    const int n = b.en[e], f = b.efirst[e];
    const double *sm = jr.dsmall + jr.blk_off[ilm * Nel + e];
    const double *bg = jr.dbig + jr.blk_off[ilm * Nel + e];
    double ts = 0.0, tb = 0.0;
    for (int idx = tid; idx < b.nch * n * n; idx += blockDim.x) {
      const int ch = idx / (n * n), rem = idx % (n * n), bcol = rem / n, a = rem % n;  // small(a,bcol)
      const double p = pin[(int64_t)ch * b.Npix + (f + bcol) * b.Nrad + f + a];       // Psub(bcol,a)
      ts += sm[idx] * p;
      if (e > 0) tb += bg[idx] * p;
    }
    for (int o = 16; o > 0; o >>= 1) {
      ts += __shfl_down_sync(0xffffffffu, ts, o);
      tb += __shfl_down_sync(0xffffffffu, tb, o);
    }
    if ((tid & 31) == 0) {
      red[0][tid >> 5] = ts;
      red[1][tid >> 5] = tb;
    }
    __syncthreads();
    if (tid == 0) {
      double a = 0, c = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
        a += red[0][w];
        c += red[1][w];
      }
      s_js[e] = fac * a;
      s_jb[e] = fac * c;
    }
    __syncthreads();
  }
  if (tid == 0) {
    // element e receives  (sum of big traces of outer elements) * small_e
    //                   + (sum of small traces of inner elements) * big_e
    double run = 0.0;
    for (int e = 0; e < Nel; e++) {
      s_pre[e] = run;
      run += s_js[e];
    }
    run = 0.0;
    for (int e = Nel - 1; e >= 0; e--) {
      s_post[e] = run;
      run += s_jb[e];
    }
  }
  __syncthreads();
  for (int e = 0; e < Nel; e++) {
    const int n = b.en[e], f = b.efirst[e], nn = n * n;
    const double *sm = jr.dsmall + jr.blk_off[ilm * Nel + e];
    const double *bg = jr.dbig + jr.blk_off[ilm * Nel + e];
    const double *Bf = jr.dB + jr.B_off[ilm * Nel + e];
    const double *sg = jr.dsigma + jr.sig_off[ilm * Nel + e];
    const int rank = jr.rank[ilm * Nel + e];
    const int ldB = b.nch * nn;
    // c_p = sigma_p sum_ch s_ch sum_ab B[ch*nn+ab,p] Psub_ch(a,b)
    for (int p = tid >> 5; p < rank; p += blockDim.x >> 5) {
      double s = 0.0;
      for (int idx = tid & 31; idx < ldB; idx += 32) {
        const int ch = idx / nn, ab = idx % nn, a = ab % n, bcol = ab / n;
        const double sgn = (b.nch == 2 && ch == 1) ? -1.0 : 1.0;
        s += sgn * Bf[(int64_t)p * ldB + idx] * pin[(int64_t)ch * b.Npix + (f + a) * b.Nrad + f + bcol];
      }
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) s_c[p] = sg[p] * s;
    }
    __syncthreads();
    for (int idx = tid; idx < ldB; idx += blockDim.x) {
      const int ch = idx / nn, ab = idx % nn, a = ab % n, bcol = ab / n;
      const double sgn = (b.nch == 2 && ch == 1) ? -1.0 : 1.0;
      double s = 0.0;
      for (int p = 0; p < rank; p++) s += Bf[(int64_t)p * ldB + idx] * s_c[p];
      double v = fac * sgn * s + s_post[e] * sm[idx];
      if (e > 0) v += s_pre[e] * bg[idx];
      out[(int64_t)((f + a) * b.Nrad + f + bcol) * nq + L * b.nch + ch] += v;
    }
    __syncthreads();
  }
}

// Coulomb unpack: dense J[(ang i, r), (ang j, c)] = Jsec[sp=(sj,si)][pix=(r,c)][pos_j*NP + pos_i].
// One CTA per ACTIVE angular block (list blocks[] = angi | angj << 16); the caller zero-fills J first.
static __global__ void k_unpack_J(BasisDev b, const int *__restrict__ ang_sec, const int *__restrict__ ang_pos,
                           const int *__restrict__ blocks, const double *__restrict__ Jsec, double *__restrict__ J,
                           int64_t ld) {
  const int angi = blocks[blockIdx.x] & 0xffff, angj = blocks[blockIdx.x] >> 16;
  const int si = b.ang_skip[angi], sj = b.ang_skip[angj];
  const int ni = b.Nrad - si, nj = b.Nrad - sj;
  double *dst = J + b.ang_off[angi] + (int64_t)b.ang_off[angj] * ld;
  const int sp = ang_sec[angj] * b.ns + ang_sec[angi];
  const double *src = Jsec + (int64_t)sp * b.Npix * b.NB + ang_pos[angj] * b.NP + ang_pos[angi];
  for (int idx = threadIdx.x; idx < ni * nj; idx += blockDim.x) {
    const int r = idx % ni + si, c = idx / ni + sj;
    dst[(r - si) + (int64_t)(c - sj) * ld] = src[(int64_t)(r * b.Nrad + c) * b.NB];
  }
}

}  // namespace dev
}  // namespace hfq
