"""ctypes front end of the C oracle (oracle/csrc/jk_oracle.c).  Test infrastructure only.

Takes the integral caches as plain arrays (from the numpy oracle or exported from the
product's tables), so the CPU baseline and the GPU path can be timed on identical inputs.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import gaunt as _gaunt

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _host_tag():
    """Hash of this machine's CPU feature flags: the -march=native build is per host."""
    import hashlib
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    return hashlib.sha1(flags.encode()).hexdigest()[:10]


def lib():
    """The C oracle, compiled on first use with -O3 -march=native for THIS host (the timed CPU arm must not
    run a generic build, and a native build must not travel to a different CPU)."""
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "csrc", "jk_oracle.c")
        so = os.path.join(_HERE, "_build", "libjk_oracle.%s.so" % _host_tag())
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", _HERE, so])
        _lib = ctypes.CDLL(so)
        _lib.jk_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return lib().jk_num_threads()


def use_all_cores():
    """Use every core this process may run on (torchrun exports OMP_NUM_THREADS=1).  Returns the count."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().jk_set_num_threads(int(n))
    return num_threads()


def _ptrs(mats):
    keep = [np.ascontiguousarray(np.asarray(m, dtype=np.float64).ravel(order="F")) for m in mats]
    arr = (ctypes.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
    return arr, keep


def _i(v):
    return np.ascontiguousarray(v, dtype=np.int32)


class DiatomicCaches:
    """Flat caches for the C oracle.  blocks[ilm*Nel+iel] = (small(2,n,n), big(2,n,n), B, sigma)."""

    def __init__(self, Nrad, efirst, en, lval, mval, lmL, lmM, pref, blocks):
        self.Nrad, self.efirst, self.en = int(Nrad), _i(efirst), _i(en)
        self.lval, self.mval, self.lmL, self.lmM = _i(lval), _i(mval), _i(lmL), _i(lmM)
        self.pref = np.ascontiguousarray(pref, dtype=np.float64)
        self.Nang, self.Nel, self.nlm = len(self.lval), len(self.en), len(self.lmL)
        self.NL = int(self.lmL.max()) + 1
        self.P0, self._k0 = _ptrs([b[0][0] for b in blocks])
        self.P2, self._k1 = _ptrs([b[0][1] for b in blocks])
        self.Q0, self._k2 = _ptrs([b[1][0] for b in blocks])
        self.Q2, self._k3 = _ptrs([b[1][1] for b in blocks])
        self.B, self._k4 = _ptrs([b[2] for b in blocks])
        self.S, self._k5 = _ptrs([b[3] for b in blocks])
        self.rank = _i([b[2].shape[1] for b in blocks])
        g0, g2 = _gaunt.coupling_tables(self.lval, self.mval, self.NL, True)
        self.g0 = np.ascontiguousarray(g0); self.g2 = np.ascontiguousarray(g2)
        LM = set()
        for i in range(self.Nang):
            for j in range(self.Nang):
                M = int(self.mval[j] - self.mval[i])
                for L in range(max(abs(int(self.lval[j] - self.lval[i])) - 2, abs(M)), int(self.lval[j] + self.lval[i]) + 3):
                    LM.add((L, M))
        LM = sorted(LM)
        self.LML, self.LMM = _i([p[0] for p in LM]), _i([p[1] for p in LM])

    def pure_idx(self):
        idx = []
        for i, m in enumerate(self.mval):
            idx += list(range(i * self.Nrad + (0 if m == 0 else 1), (i + 1) * self.Nrad))
        return np.array(idx)

    def expand(self, P):
        nd = self.Nang * self.Nrad
        out = np.zeros((nd, nd), order="F")
        pi = self.pure_idx()
        out[np.ix_(pi, pi)] = P
        return out

    def exchange_blocks(self, Pdummy, jangs, kangs):
        """Selected output blocks K.block(jang,kang) (Nrad x Nrad each), reference sign."""
        Pd = np.asfortranarray(Pdummy)
        jangs, kangs = _i(jangs), _i(kangs)
        out = np.zeros((len(jangs), self.Nrad, self.Nrad))
        v = ctypes.c_void_p
        lib().jk_diatomic_exchange_blocks(
            self.Nang, self.Nrad, self.Nel, self.NL, v(self.efirst.ctypes.data), v(self.en.ctypes.data),
            v(self.lval.ctypes.data), v(self.mval.ctypes.data), v(self.g0.ctypes.data), v(self.g2.ctypes.data),
            self.nlm, v(self.lmL.ctypes.data), v(self.lmM.ctypes.data), v(self.pref.ctypes.data),
            self.P0, self.P2, self.Q0, self.Q2, self.B, self.S, v(self.rank.ctypes.data), v(Pd.ctypes.data),
            len(jangs), v(jangs.ctypes.data), v(kangs.ctypes.data), v(out.ctypes.data))
        return np.transpose(out, (0, 2, 1))  # stored column-major per block

    def exchange(self, P):
        """Full K (boundary removed), all output blocks."""
        Pd = self.expand(P)
        ja, ka = np.meshgrid(np.arange(self.Nang), np.arange(self.Nang), indexing="ij")
        blk = self.exchange_blocks(Pd, ja.ravel(), ka.ravel())
        nd = self.Nang * self.Nrad
        K = np.zeros((nd, nd))
        for b, (j, k) in enumerate(zip(ja.ravel(), ka.ravel())):
            K[j * self.Nrad:(j + 1) * self.Nrad, k * self.Nrad:(k + 1) * self.Nrad] = blk[b]
        pi = self.pure_idx()
        return K[np.ix_(pi, pi)]

    def coulomb(self, P, single_M=None):
        """Full J (boundary removed).  single_M: the density only has blocks with m_k - m_l == single_M (checked):
        build only those channels -- identical result, without the folds that add exact zeros."""
        Pd = self.expand(P)
        nd = self.Nang * self.Nrad
        J = np.zeros((nd, nd), order="F")
        v = ctypes.c_void_p
        if single_M is not None:
            N = self.Nrad
            for k in range(self.Nang):
                for l in range(self.Nang):
                    if self.mval[k] - self.mval[l] != single_M and Pd[k * N:(k + 1) * N, l * N:(l + 1) * N].any():
                        raise ValueError("density has a block with m_k - m_l != single_M")
            lib().jk_diatomic_coulomb_single_M(
                self.Nang, self.Nrad, self.Nel, self.NL, v(self.efirst.ctypes.data), v(self.en.ctypes.data),
                v(self.lval.ctypes.data), v(self.mval.ctypes.data), v(self.g0.ctypes.data), v(self.g2.ctypes.data),
                self.nlm, v(self.lmL.ctypes.data), v(self.lmM.ctypes.data), v(self.pref.ctypes.data),
                len(self.LML), v(self.LML.ctypes.data), v(self.LMM.ctypes.data),
                self.P0, self.P2, self.Q0, self.Q2, self.B, self.S, v(self.rank.ctypes.data), v(Pd.ctypes.data),
                v(J.ctypes.data), int(single_M))
            pi = self.pure_idx()
            return J[np.ix_(pi, pi)]
        lib().jk_diatomic_coulomb(
            self.Nang, self.Nrad, self.Nel, self.NL, v(self.efirst.ctypes.data), v(self.en.ctypes.data),
            v(self.lval.ctypes.data), v(self.mval.ctypes.data), v(self.g0.ctypes.data), v(self.g2.ctypes.data),
            self.nlm, v(self.lmL.ctypes.data), v(self.lmM.ctypes.data), v(self.pref.ctypes.data),
            len(self.LML), v(self.LML.ctypes.data), v(self.LMM.ctypes.data),
            self.P0, self.P2, self.Q0, self.Q2, self.B, self.S, v(self.rank.ctypes.data), v(Pd.ctypes.data),
            v(J.ctypes.data))
        pi = self.pure_idx()
        return J[np.ix_(pi, pi)]

    def coulomb_timing_sample(self, Pdummy, stride):
        """Seconds of a 1/stride sample of the (serial) Coulomb build -- see jk_oracle.c; no result."""
        import time
        Pd = np.asfortranarray(Pdummy)
        nd = self.Nang * self.Nrad
        if getattr(self, "_jscratch", None) is None or self._jscratch.shape[0] != nd:
            self._jscratch = np.zeros((nd, nd), order="F")
        v = ctypes.c_void_p
        t0 = time.perf_counter()
        lib().jk_diatomic_coulomb_timing_sample(
            self.Nang, self.Nrad, self.Nel, self.NL, v(self.efirst.ctypes.data), v(self.en.ctypes.data),
            v(self.lval.ctypes.data), v(self.mval.ctypes.data), v(self.g0.ctypes.data), v(self.g2.ctypes.data),
            self.nlm, v(self.lmL.ctypes.data), v(self.lmM.ctypes.data), v(self.pref.ctypes.data),
            len(self.LML), v(self.LML.ctypes.data), v(self.LMM.ctypes.data),
            self.P0, self.P2, self.Q0, self.Q2, self.B, self.S, v(self.rank.ctypes.data), v(Pd.ctypes.data),
            v(self._jscratch.ctypes.data), int(stride))
        return time.perf_counter() - t0

    @classmethod
    def from_oracle(cls, b):
        Nel = b.radial.Nel()
        blocks = []
        for ilm in range(len(b.lm_map)):
            for e in range(Nel):
                i = ilm * Nel + e
                blocks.append((np.stack([b.disjoint_P0[i], b.disjoint_P2[i]]),
                               np.stack([b.disjoint_Q0[i], b.disjoint_Q2[i]]), b.cd_B[i], b.cd_sigma[i]))
        return cls(b.Nrad(), [b.radial.get_idx(e)[0] for e in range(Nel)], [b.radial.fem.nprim(e) for e in range(Nel)],
                   b.lval, b.mval, [p[0] for p in b.lm_map], [p[1] for p in b.lm_map], b.LMfac_abs(), blocks)

    @classmethod
    def from_npz(cls, path):
        """From the arrays written by tools/export_caches.py (the CPU reference arm's inputs)."""
        d = dict(np.load(path))   # NpzFile re-reads an array on every access: materialise once
        en, ranks = d["en"], d["ranks"]
        Nel, nlm = len(en), len(d["lmL"])
        blocks, o1, o2 = [], 0, 0
        so = 0
        for ilm in range(nlm):
            for e in range(Nel):
                n, r = int(en[e]), int(ranks[ilm * Nel + e])
                sm = d["small"][o1:o1 + 2 * n * n].reshape(2, n, n).transpose(0, 2, 1)
                bg = d["big"][o1:o1 + 2 * n * n].reshape(2, n, n).transpose(0, 2, 1)
                o1 += 2 * n * n
                B = d["B"][o2:o2 + 2 * n * n * r].reshape(2 * n * n, r, order="F")
                o2 += 2 * n * n * r
                blocks.append((sm, bg, B, d["sigma"][so:so + r]))
                so += r
        return cls(int(d["Nrad"]), d["efirst"], en, d["lval"], d["mval"], d["lmL"], d["lmM"], d["pref"], blocks)

    @classmethod
    def from_tables(cls, T):
        """From the product's exported tables (identical inputs for the CPU baseline)."""
        blocks = [T.block(ilm, e) for ilm in range(T.nlm) for e in range(T.Nel)]
        return cls(T.Nrad, T.efirst, T.en, T.lval, T.mval, T.lmL, T.lmM, T.pref, blocks)


class AtomicCaches:
    """Flat caches of the atomic basis for the C oracle (jk_atomic_exchange_blocks: TwoDBasis.cpp:879-999,
    CoulombExchangeFE.h:484-530).  blocks[L*Nel+iel] = (small(1,n,n), big(1,n,n), chol, sigma)."""

    def __init__(self, Nrad, efirst, en, lval, mval, blocks):
        self.Nrad, self.efirst, self.en = int(Nrad), _i(efirst), _i(en)
        self.lval, self.mval = _i(lval), _i(mval)
        self.Nang, self.Nel = len(self.lval), len(self.en)
        self.NL = len(blocks) // self.Nel
        self.small, self._k0 = _ptrs([b[0][0] for b in blocks])
        self.big, self._k1 = _ptrs([b[1][0] for b in blocks])
        self.chol, self._k2 = _ptrs([b[2] for b in blocks])
        self.rank = _i([b[2].shape[1] for b in blocks])
        _, g2 = _gaunt.coupling_tables(self.lval, self.mval, self.NL, False)
        self.g2 = np.ascontiguousarray(g2)

    @classmethod
    def from_tables(cls, T):
        blocks = [T.block(L, e) for L in range(T.nlm) for e in range(T.Nel)]
        return cls(T.Nrad, T.efirst, T.en, T.lval, T.mval, blocks)

    def exchange_blocks(self, P, jangs, kangs):
        Pd = np.asfortranarray(P)
        jangs, kangs = _i(jangs), _i(kangs)
        out = np.zeros((len(jangs), self.Nrad, self.Nrad))
        v = ctypes.c_void_p
        lib().jk_atomic_exchange_blocks(
            self.Nang, self.Nrad, self.Nel, self.NL, v(self.efirst.ctypes.data), v(self.en.ctypes.data),
            v(self.lval.ctypes.data), v(self.mval.ctypes.data), v(self.g2.ctypes.data), self.small, self.big, self.chol,
            v(self.rank.ctypes.data), v(Pd.ctypes.data), len(jangs), v(jangs.ctypes.data), v(kangs.ctypes.data),
            v(out.ctypes.data))
        return np.transpose(out, (0, 2, 1))

    def exchange(self, P):
        na, N = self.Nang, self.Nrad
        ja = [j for j in range(na) for k in range(na)]
        ka = [k for j in range(na) for k in range(na)]
        blk = self.exchange_blocks(P, ja, ka)
        K = np.zeros((na * N, na * N))
        for b, (j, k) in enumerate(zip(ja, ka)):
            K[j * N:(j + 1) * N, k * N:(k + 1) * N] = blk[b]
        return K
