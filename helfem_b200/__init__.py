"""helfem_b200 -- B200-native SCF Fock builds (J, K) for HelFEM bases.

Host-side mirror of the reference's operator interface for this path: the
``TwoDBasis`` classes expose ``compute_tei()``, ``coulomb(P)``, ``exchange(P)``
(and ``set_absm_symmetric`` for the diatomic basis) with the reference's names,
argument meaning and error behaviour (src/atomic/TwoDBasis.h:181,223-227,
src/diatomic/basis.h:265,300-302,344).  Everything is computed by the CUDA
library ``libhelfemqc_b200.so`` through the C ABI in include/helfem_b200.h;
there is no CPU fallback -- if the library or a GPU is missing, calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libhelfemqc_b200.so")
_lib = None

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)


class _TablesDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("nch", ctypes.c_int), ("Nrad", ctypes.c_int), ("Nel", ctypes.c_int),
                ("Nang", ctypes.c_int), ("nlm", ctypes.c_int),
                ("efirst", _c_int_p), ("en", _c_int_p), ("lval", _c_int_p), ("mval", _c_int_p),
                ("lmL", _c_int_p), ("lmM", _c_int_p), ("pref", _c_dbl_p), ("rank", _c_int_p),
                ("small_", _c_dbl_p), ("big_", _c_dbl_p), ("B", _c_dbl_p), ("sigma", _c_dbl_p),
                ("Rhalf", ctypes.c_double)]


class _TablesInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("kind", "nch", "Nrad", "Nel", "Nang", "nlm", "Nbf", "Ndummy")]


EXPORTED_SYMBOLS = [
    "hfq_last_error", "hfq_tables_atomic", "hfq_tables_atomic_yukawa", "hfq_tables_atomic_erfc",
    "hfq_tables_set_pair_tensors", "hfq_sap_table", "hfq_tables_get_pair_tensor", "hfq_erfc_phi", "hfq_tables_sadatom", "hfq_tables_sadatom_rs", "hfq_tables_diatomic", "hfq_tables_diatomic_device", "hfq_tables_from_arrays", "hfq_tables_get_info",
    "hfq_tables_get_ints", "hfq_tables_get_doubles", "hfq_tables_get_block", "hfq_tables_one_electron",
    "hfq_tables_destroy", "hfq_create", "hfq_destroy", "hfq_nbf", "hfq_set_absm_symmetric", "hfq_coulomb",
    "hfq_exchange", "hfq_coulomb_device", "hfq_exchange_device", "hfq_last_timings", "hfq_exchange_output_pattern",
    "hfq_coulomb_output_pattern",
    "hfq_tables_sadatom_batch", "hfq_coulomb_radial_batch", "hfq_syev_batch", "hfq_set_host_threads", "hfq_fock_build", "hfq_fock_build_device", "hfq_comm_unique_id", "hfq_comm_init", "hfq_comm_size", "hfq_shard_assign",
    "hfq_coulomb_exchange", "hfq_coulomb_exchange_device", "hfq_grid_attach", "hfq_grid_npoints", "hfq_grid_density", "hfq_grid_fxc", "hfq_eval_fxc",
]


def lib():
    """Load libhelfemqc_b200.so (built in-tree by helfem_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise RuntimeError("libhelfemqc_b200.so is missing: run `python -m helfem_b200.build` "
                           "(there is no CPU fallback for the Fock-build path)")
    L = ctypes.CDLL(_LIBPATH)
    L.hfq_last_error.restype = ctypes.c_char_p
    vp, ci, cd, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int64
    L.hfq_tables_atomic.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci, cd, ci, cd, ci]
    L.hfq_tables_atomic_yukawa.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci, cd, ci, cd, ci, cd]
    L.hfq_tables_atomic_erfc.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci, cd, ci, cd, ci, cd]
    L.hfq_tables_set_pair_tensors.argtypes = [vp, vp, i64]
    L.hfq_tables_get_pair_tensor.argtypes = [vp, ci, ci, ci, vp, i64]
    L.hfq_tables_get_pair_tensor.restype = i64
    L.hfq_erfc_phi.argtypes = [ci, cd, cd]
    L.hfq_erfc_phi.restype = cd
    L.hfq_sap_table.argtypes = [vp, vp, vp, ci, ci, vp, i64]
    L.hfq_sap_table.restype = i64
    L.hfq_tables_sadatom.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, cd, ci, cd, ci]
    L.hfq_tables_sadatom_rs.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, cd, ci, cd, ci, ci, cd]
    L.hfq_tables_diatomic.argtypes = [ctypes.POINTER(vp), ci, ci, cd, _c_int_p, ci, ci, ci, cd, ci, cd, ci]
    L.hfq_tables_diatomic_device.argtypes = [ctypes.POINTER(vp), ci, ci, cd, _c_int_p, ci, ci, ci, cd, ci, cd, ci, ci]
    L.hfq_tables_from_arrays.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(_TablesDesc)]
    L.hfq_tables_get_info.argtypes = [vp, ctypes.POINTER(_TablesInfo)]
    L.hfq_tables_get_ints.argtypes = [vp, ci, vp, i64]
    L.hfq_tables_get_doubles.argtypes = [vp, ci, vp, i64]
    L.hfq_tables_get_block.argtypes = [vp, ci, ci, vp, vp, vp, vp]
    L.hfq_tables_one_electron.argtypes = [vp, vp, vp, vp]
    L.hfq_tables_destroy.argtypes = [vp]
    L.hfq_tables_destroy.restype = None
    L.hfq_create.argtypes = [ctypes.POINTER(vp), vp, ci]
    L.hfq_destroy.argtypes = [vp]
    L.hfq_destroy.restype = None
    L.hfq_nbf.argtypes = [vp]
    L.hfq_set_absm_symmetric.argtypes = [vp, ci]
    L.hfq_coulomb.argtypes = [vp, vp, i64, vp, i64]
    L.hfq_exchange.argtypes = [vp, vp, i64, vp, i64]
    L.hfq_coulomb_device.argtypes = [vp, vp, i64, vp, i64, vp]
    L.hfq_exchange_device.argtypes = [vp, vp, i64, vp, i64, ci, ci, vp]
    L.hfq_last_timings.argtypes = [vp, vp, ci]
    L.hfq_exchange_output_pattern.argtypes = [vp, vp, i64, vp, i64]
    L.hfq_coulomb_output_pattern.argtypes = [vp, vp, i64, vp, i64]
    L.hfq_coulomb_exchange.argtypes = [vp, vp, i64, cd, vp, i64, vp, i64]
    L.hfq_coulomb_exchange_device.argtypes = [vp, vp, i64, cd, vp, i64, vp, i64, ci, ci, vp]
    L.hfq_tables_sadatom_batch.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, cd, ci, cd, ci]
    L.hfq_coulomb_radial_batch.argtypes = [vp, vp, vp, ci, cd, vp]
    L.hfq_syev_batch.argtypes = [vp, vp, ci, i64, vp]
    L.hfq_set_host_threads.argtypes = [ci]
    L.hfq_fock_build.argtypes = [vp, vp, i64, cd, vp, i64, vp, i64, ci, ci, vp, i64, vp, vp, cd]
    L.hfq_fock_build_device.argtypes = [vp, vp, i64, cd, vp, i64, vp, i64, ci, ci, vp, i64, vp, vp, cd, vp]
    L.hfq_comm_unique_id.argtypes = [vp]
    L.hfq_comm_init.argtypes = [vp, vp, ci, ci]
    L.hfq_comm_size.argtypes = [vp]
    L.hfq_shard_assign.argtypes = [vp, ci, ci, vp]
    L.hfq_grid_attach.argtypes = [vp, ci, ci]
    L.hfq_grid_npoints.argtypes = [vp]
    L.hfq_grid_npoints.restype = i64
    L.hfq_grid_density.argtypes = [vp, vp, i64, vp, i64, ci, vp, vp, vp, vp, vp, vp, vp]
    L.hfq_grid_fxc.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, i64, vp, i64, vp]
    L.hfq_eval_fxc.argtypes = [vp, ci, ci, vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, vp, ci, cd]
    _lib = L
    return L


def set_host_threads(n=None, world=1):
    """OpenMP threads of the host-side setup; default: this process' share of the cores it may run on."""
    if n is None:
        n = max(1, len(os.sched_getaffinity(0)) // max(1, world))
    _check(lib().hfq_set_host_threads(int(n)))
    return int(n)


class HfqError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        msg = lib().hfq_last_error().decode()
        # the reference throws std::logic_error for misuse and size mismatches
        raise (ValueError if rc == -1 else HfqError)(msg)
    return rc


def _fmat(a, n):
    a = np.asarray(a, dtype=np.float64)
    if a.shape != (n, n):
        raise ValueError("Matrix does not have expected size! Got %s, expected %i x %i!" % (a.shape, n, n))
    return np.asfortranarray(a)


class Tables:
    """Host-side basis description and integral caches (hfq_tables)."""

    def __init__(self, handle):
        self._h = handle
        info = _TablesInfo()
        _check(lib().hfq_tables_get_info(self._h, ctypes.byref(info)))
        for name, _ in _TablesInfo._fields_:
            setattr(self, name, getattr(info, name))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.hfq_tables_destroy(self._h)
            self._h = None

    @classmethod
    def atomic(cls, Z, lmax, mmax, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, nquad=0):
        h = ctypes.c_void_p()
        _check(lib().hfq_tables_atomic(ctypes.byref(h), Z, lmax, mmax, nelem, nnodes, Rmax, igrid, zexp, nquad))
        return cls(h)

    @classmethod
    def atomic_yukawa(cls, Z, lmax, mmax, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, nquad=0, lam=0.4):
        h = ctypes.c_void_p()
        _check(lib().hfq_tables_atomic_yukawa(ctypes.byref(h), Z, lmax, mmax, nelem, nnodes, Rmax, igrid, zexp, nquad, lam))
        return cls(h)

    @classmethod
    def atomic_erfc(cls, Z, lmax, mmax, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, nquad=0, mu=0.3):
        h = ctypes.c_void_p()
        _check(lib().hfq_tables_atomic_erfc(ctypes.byref(h), Z, lmax, mmax, nelem, nnodes, Rmax, igrid, zexp, nquad, mu))
        return cls(h)

    def set_pair_tensors(self, ktei_list):
        """Attach the reference's rs_ktei cache: list over (L*Nel + iel)*Nel + jel of (Ni*Nj) x (Ni*Nj) matrices
        ktei[kk*Ni + jj, ll*Ni + ii] (src/atomic/TwoDBasis.h, utils::exchange_tei)."""
        flat = np.concatenate([np.asarray(k, dtype=np.float64).reshape(-1, order="F") for k in ktei_list])
        _check(lib().hfq_tables_set_pair_tensors(self._h, flat.ctypes.data, flat.size))
        return self

    def pair_tensor(self, L, iel, jel):
        n = _check(lib().hfq_tables_get_pair_tensor(self._h, L, iel, jel, None, 0))
        if n == 0:
            return None
        out = np.empty(n)
        _check(lib().hfq_tables_get_pair_tensor(self._h, L, iel, jel, out.ctypes.data, n))
        m = int(round(np.sqrt(n)))
        return out.reshape(m, m, order="F")

    @classmethod
    def sadatom(cls, Z, lmax, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, nquad=0):
        h = ctypes.c_void_p()
        _check(lib().hfq_tables_sadatom(ctypes.byref(h), Z, lmax, nelem, nnodes, Rmax, igrid, zexp, nquad))
        return cls(h)

    @classmethod
    def sadatom_batch(cls, lmax, nbatch, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, nquad=0):
        h = ctypes.c_void_p()
        _check(lib().hfq_tables_sadatom_batch(ctypes.byref(h), lmax, nbatch, nelem, nnodes, Rmax, igrid, zexp, nquad))
        return cls(h)

    @classmethod
    def sadatom_rs(cls, Z, lmax, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, nquad=0, rs=1, param=0.4):
        h = ctypes.c_void_p()
        _check(lib().hfq_tables_sadatom_rs(ctypes.byref(h), Z, lmax, nelem, nnodes, Rmax, igrid, zexp, nquad, rs, param))
        return cls(h)

    @classmethod
    def diatomic(cls, Z1, Z2, Rbond, lmax_per_m, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=1.0, nquad=0, device=None):
        """device: None = compute_tei on the host (OpenMP); a CUDA device index = in-element kernels on that GPU."""
        h = ctypes.c_void_p()
        lm = (ctypes.c_int * len(lmax_per_m))(*[int(x) for x in lmax_per_m])
        if device is None:
            _check(lib().hfq_tables_diatomic(ctypes.byref(h), Z1, Z2, Rbond, lm, len(lmax_per_m), nelem, nnodes, Rmax,
                                             igrid, zexp, nquad))
        else:
            _check(lib().hfq_tables_diatomic_device(ctypes.byref(h), Z1, Z2, Rbond, lm, len(lmax_per_m), nelem, nnodes,
                                                    Rmax, igrid, zexp, nquad, int(device)))
        return cls(h)

    @classmethod
    def from_arrays(cls, kind, Nrad, efirst, en, lval, mval, lmL, lmM, pref, blocks, Rhalf=0.0):
        """blocks[ilm*Nel+iel] = (small, big, B, sigma) with small/big shaped (nch, n, n)
        (each n x n block column-major when flattened), B (nch*n*n, rank), sigma (rank,)."""
        nch = 2 if kind == 1 else 1
        ia = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        efirst, en, lval, mval, lmL, lmM = map(ia, (efirst, en, lval, mval, lmL, lmM))
        pref = np.ascontiguousarray(pref, dtype=np.float64)
        rank = ia([b[2].shape[1] for b in blocks])

        def cat(parts):
            return np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64).ravel() for p in parts])
                                        if parts else np.zeros(0))
        small = cat([np.concatenate([np.asarray(m).ravel(order="F") for m in b[0]]) for b in blocks])
        big = cat([np.concatenate([np.asarray(m).ravel(order="F") for m in b[1]]) for b in blocks])
        Bc = cat([np.asarray(b[2]).ravel(order="F") for b in blocks])
        sig = cat([np.asarray(b[3]) for b in blocks])
        if len(sig) == 0:
            sig = np.zeros(1)
        if len(Bc) == 0:
            Bc = np.zeros(1)
        d = _TablesDesc(kind, nch, Nrad, len(en), len(lval), len(lmL),
                        efirst.ctypes.data_as(_c_int_p), en.ctypes.data_as(_c_int_p), lval.ctypes.data_as(_c_int_p),
                        mval.ctypes.data_as(_c_int_p), lmL.ctypes.data_as(_c_int_p), lmM.ctypes.data_as(_c_int_p),
                        pref.ctypes.data_as(_c_dbl_p), rank.ctypes.data_as(_c_int_p), small.ctypes.data_as(_c_dbl_p),
                        big.ctypes.data_as(_c_dbl_p), Bc.ctypes.data_as(_c_dbl_p), sig.ctypes.data_as(_c_dbl_p),
                        float(Rhalf))
        h = ctypes.c_void_p()
        _check(lib().hfq_tables_from_arrays(ctypes.byref(h), ctypes.byref(d)))
        return cls(h)

    def ints(self, what, n):
        out = np.zeros(n, dtype=np.int32)
        _check(lib().hfq_tables_get_ints(self._h, what, out.ctypes.data, n))
        return out

    @property
    def lval(self): return self.ints(0, self.Nang)
    @property
    def mval(self): return self.ints(1, self.Nang)
    @property
    def efirst(self): return self.ints(2, self.Nel)
    @property
    def en(self): return self.ints(3, self.Nel)
    @property
    def lmL(self): return self.ints(4, self.nlm)
    @property
    def lmM(self): return self.ints(5, self.nlm)
    @property
    def ranks(self): return self.ints(6, self.nlm * self.Nel)

    @property
    def pref(self):
        out = np.zeros(self.nlm)
        _check(lib().hfq_tables_get_doubles(self._h, 0, out.ctypes.data, self.nlm))
        return out

    @property
    def bval(self):
        out = np.zeros(self.Nel + 1)
        _check(lib().hfq_tables_get_doubles(self._h, 1, out.ctypes.data, self.Nel + 1))
        return out

    def block(self, ilm, iel):
        """(small, big, B, sigma): small/big (nch, n, n), B (nch*n*n, rank)."""
        n = int(self.en[iel])
        r = int(self.ranks[ilm * self.Nel + iel])
        small = np.zeros(self.nch * n * n); big = np.zeros(self.nch * n * n)
        B = np.zeros(max(1, self.nch * n * n * r)); sig = np.zeros(max(1, r))
        _check(lib().hfq_tables_get_block(self._h, ilm, iel, small.ctypes.data, big.ctypes.data, B.ctypes.data,
                                          sig.ctypes.data))
        sm = np.stack([small[c * n * n:(c + 1) * n * n].reshape(n, n, order="F") for c in range(self.nch)])
        bg = np.stack([big[c * n * n:(c + 1) * n * n].reshape(n, n, order="F") for c in range(self.nch)])
        return sm, bg, B[:self.nch * n * n * r].reshape(self.nch * n * n, r, order="F"), sig[:r]

    def one_electron(self):
        n = self.Nbf
        S, T, V = (np.zeros((n, n), order="F") for _ in range(3))
        _check(lib().hfq_tables_one_electron(self._h, S.ctypes.data, T.ctypes.data, V.ctypes.data))
        return S, T, V


class _BasisBase:
    """Common part of the atomic / diatomic TwoDBasis mirrors."""

    def __init__(self, device=0):
        self._tables = None
        self._ctx = None
        self._device = device
        self._absm = False

    def __del__(self):
        if getattr(self, "_ctx", None) and _lib is not None:
            _lib.hfq_destroy(self._ctx)
            self._ctx = None

    # -- reference API -------------------------------------------------------
    def compute_tei(self, exchange=True):
        """Build the integral caches (host) -- TwoDBasis::compute_tei."""
        self._tables = self._make_tables()
        return self

    def Nbf(self):
        return self.tables.Nbf

    def Nrad(self):
        return self.tables.Nrad

    def Nang(self):
        return self.tables.Nang

    def coulomb(self, P):
        ctx = self._context()
        n = self.Nbf()
        Pf = _fmat(P, n)
        J = np.full((n, n), np.nan, order="F")   # every element must be defined by the call
        _check(lib().hfq_coulomb(ctx, Pf.ctypes.data, n, J.ctypes.data, n))
        return J

    def exchange(self, P):
        ctx = self._context()
        n = self.Nbf()
        Pf = _fmat(P, n)
        K = np.full((n, n), np.nan, order="F")   # every element must be defined by the call
        _check(lib().hfq_exchange(ctx, Pf.ctypes.data, n, K.ctypes.data, n))
        return K

    def coulomb_exchange(self, P, kscale=1.0):
        """(J, K) = (coulomb(P), exchange(kscale * P)) in one call (one upload, overlapped copies)."""
        ctx = self._context()
        n = self.Nbf()
        Pf = _fmat(P, n)
        J = np.full((n, n), np.nan, order="F")   # every element must be defined by the call
        K = np.full((n, n), np.nan, order="F")   # every element must be defined by the call
        _check(lib().hfq_coulomb_exchange(ctx, Pf.ctypes.data, n, kscale, J.ctypes.data, n, K.ctypes.data, n))
        return J, K

    def coulomb_exchange_device(self, dP_ptr, dJ_ptr, dK_ptr, kscale=1.0, shard=0, nshards=1, stream=None):
        n = self.Nbf()
        _check(lib().hfq_coulomb_exchange_device(self._context(), dP_ptr, n, kscale, dJ_ptr, n, dK_ptr, n, shard, nshards,
                                                 stream))

    def fock_build_device(self, dP_ptr, dJ_ptr, dK_ptr, kscale=0.5, x_func=-1, c_func=0, dH_ptr=None, thr=1e-12, stream=None):
        """XC + J + K of one restricted Fock build, device-resident (hfq_fock_build_device); returns (Exc, Nel).
        A DFT grid must be attached (DFTGrid(basis, ...))."""
        n = self.Nbf()
        exc, nel = ctypes.c_double(), ctypes.c_double()
        _check(lib().hfq_fock_build_device(self._context(), dP_ptr, n, kscale, dJ_ptr, n, dK_ptr, n, x_func, c_func, dH_ptr, n,
                                           ctypes.byref(exc), ctypes.byref(nel), thr, stream))
        return exc.value, nel.value

    # -- device-resident variants (torch CUDA tensors, column-major = transposed view) ---------
    def coulomb_device(self, dP_ptr, dJ_ptr, stream=None):
        n = self.Nbf()
        _check(lib().hfq_coulomb_device(self._context(), dP_ptr, n, dJ_ptr, n, stream))

    def exchange_device(self, dP_ptr, dK_ptr, shard=0, nshards=1, stream=None):
        n = self.Nbf()
        _check(lib().hfq_exchange_device(self._context(), dP_ptr, n, dK_ptr, n, shard, nshards, stream))

    def comm_init(self, rank=None, world=None, bcast=None):
        """Bind the context to an NCCL communicator of `world` ranks (one process per GPU): hfq_comm_init.  The
        128-byte id is created on rank 0 and distributed with `bcast(bytes_or_None) -> bytes`; by default through
        torch.distributed (plumbing only: the collective of the build itself is issued by the library)."""
        if bcast is None:
            from . import dist as _dist
            rank, world, bcast = _dist.torch_bcast()
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _check(lib().hfq_comm_unique_id(buf))
        ident = bcast(buf.raw if rank == 0 else None)
        _check(lib().hfq_comm_init(self._context(), ctypes.c_char_p(ident), rank, world))
        return self

    def exchange_output_pattern(self, coulomb=False):
        """(bf_sector[Nbf], [(row sector, col sector), ...]) of the last exchange (or coulomb) result."""
        n = self.Nbf()
        bs = np.zeros(n, dtype=np.int32)
        cap = 4 * self.tables.Nang ** 2 + 16
        pr = np.zeros(cap, dtype=np.int32)
        fn = lib().hfq_coulomb_output_pattern if coulomb else lib().hfq_exchange_output_pattern
        k = _check(fn(self._context(), bs.ctypes.data, n, pr.ctypes.data, cap))
        return bs, [(int(pr[2 * i]), int(pr[2 * i + 1])) for i in range(k)]

    def last_timings(self):
        out = np.zeros(20)
        _check(lib().hfq_last_timings(self._context(), out.ctypes.data, 20))
        keys = ["ms_pack", "ms_fold", "ms_tgemm", "ms_offdiag", "ms_unpack", "ms_total", "flops_fold", "flops_tgemm",
                "flops_offdiag", "launches", "device_bytes", "alg_fold", "alg_tgemm", "alg_offdiag",
                "launches_fold", "launches_tgemm", "launches_offdiag", "h2d_bytes", "d2h_bytes", "speculative_hits"]
        return dict(zip(keys, out))

    # -- helpers ---------------------------------------------------------------
    @property
    def tables(self):
        if self._tables is None:
            raise ValueError("Primitive teis have not been computed!\n")   # reference: std::logic_error
        return self._tables

    def _context(self):
        if self._ctx is None:
            t = self.tables
            h = ctypes.c_void_p()
            _check(lib().hfq_create(ctypes.byref(h), t._h, self._device))
            self._ctx = h
            _check(lib().hfq_set_absm_symmetric(self._ctx, int(self._absm)))
        return self._ctx

    def overlap(self):
        return self.tables.one_electron()[0]

    def kinetic(self):
        return self.tables.one_electron()[1]

    def nuclear(self):
        return self.tables.one_electron()[2]


class AtomicTwoDBasis(_BasisBase):
    """helfem::atomic::basis::TwoDBasisT<double> (src/atomic/TwoDBasis.h)."""

    def __init__(self, Z, lmax, mmax, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, nquad=0, device=0):
        super().__init__(device)
        self._args = (Z, lmax, mmax, nelem, nnodes, Rmax, igrid, zexp, nquad)

    def _make_tables(self):
        return Tables.atomic(*self._args)

    def compute_yukawa(self, lam):
        """TwoDBasisT::compute_yukawa (src/atomic/TwoDBasis.cpp:737-758): Yukawa-screened caches."""
        self._rs = _BasisBase(self._device)
        self._rs._tables = Tables.atomic_yukawa(*self._args, lam=lam)
        return self

    def compute_erfc(self, mu):
        """TwoDBasisT::compute_erfc (src/atomic/TwoDBasis.cpp:762-771): erfc-attenuated pair tensors."""
        self._rs = _BasisBase(self._device)
        self._rs._tables = Tables.atomic_erfc(*self._args, mu=mu)
        return self

    def rs_exchange(self, P):
        """TwoDBasisT::rs_exchange (src/atomic/TwoDBasis.cpp:1001-1131), kernel of the last compute_yukawa /
        compute_erfc call."""
        if getattr(self, "_rs", None) is None:
            raise ValueError("Primitive teis have not been computed!\n")
        return self._rs.exchange(P)


class DiatomicTwoDBasis(_BasisBase):
    """helfem::diatomic::basis::TwoDBasis (src/diatomic/basis.h)."""

    def __init__(self, Z1, Z2, Rbond, lmax_per_m, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=1.0, nquad=0,
                 device=0, tei_on_device=False):
        """tei_on_device: compute_tei() evaluates the in-element two-electron kernels on the GPU `device`
        (hfq_tables_diatomic_device) instead of the host."""
        super().__init__(device)
        self._args = (Z1, Z2, Rbond, list(lmax_per_m), nelem, nnodes, Rmax, igrid, zexp, nquad)
        self._tei_device = device if tei_on_device else None

    def _make_tables(self):
        return Tables.diatomic(*self._args, device=self._tei_device)

    def set_absm_symmetric(self, sym):
        self._absm = bool(sym)
        if self._ctx is not None:
            _check(lib().hfq_set_absm_symmetric(self._ctx, int(self._absm)))

    def is_absm_symmetric(self):
        return self._absm


class SadatomTwoDBasis:
    """helfem::sadatom::basis::TwoDBasis (src/sadatom/basis.h:72,110-114): spherically averaged atom.
    ``coulomb(Prad)`` takes the radial density matrix, ``exchange(cube)`` a list of per-l matrices."""

    def __init__(self, Z, lmax, nelem, nnodes=15, Rmax=40.0, igrid=4, zexp=2.0, nquad=0, device=0):
        self._args = (Z, lmax, nelem, nnodes, Rmax, igrid, zexp, nquad)
        self.lmax = lmax
        self._k = _BasisBase(device)
        self._j = _BasisBase(device)

    def compute_tei(self, exchange=True):
        Z, lmax, nelem, nnodes, Rmax, igrid, zexp, nquad = self._args
        self._k._tables = Tables.sadatom(*self._args)
        self._j._tables = Tables.atomic(Z, 0, 0, nelem, nnodes, Rmax, igrid, zexp, nquad)
        return self

    def Nrad(self):
        return self._k.tables.Nrad

    def coulomb(self, Prad):
        """J(P_in) = 4 pi J_0(P_in), src/sadatom/basis.cpp:186-207."""
        return 4.0 * np.pi * self._j.coulomb(Prad)

    def compute_yukawa(self, lam):
        """src/sadatom/basis.cpp:154-173."""
        self._rs = _BasisBase(self._k._device)
        self._rs._tables = Tables.sadatom_rs(*self._args, rs=1, param=lam)
        return self

    def compute_erfc(self, mu):
        """src/sadatom/basis.cpp:175-184."""
        self._rs = _BasisBase(self._k._device)
        self._rs._tables = Tables.sadatom_rs(*self._args, rs=2, param=mu)
        return self

    def sap_table(self, Pl_a, Pl_b=None, x_func=1):
        """effective_potential_table (src/sadatom/main.cpp:55-107): rows (nucleus, then element x node), columns
        r, rho, grad rho, lapl rho, tau, v_coul, v_xc, weight, Z_eff.  Host-side post-processing."""
        t = self._k.tables
        N = t.Nrad
        pack = lambda cube: np.ascontiguousarray(np.stack([np.asfortranarray(np.asarray(c, dtype=np.float64)).T
                                                          for c in cube]))
        if len(Pl_a) != self.lmax + 1 or (Pl_b is not None and len(Pl_b) != self.lmax + 1):
            raise ValueError("Density matrix am does not match basis set!")
        a = pack(Pl_a)
        if a.shape != (len(Pl_a), N, N):
            raise ValueError("Density matrix does not match basis set!")
        b = None if Pl_b is None else pack(Pl_b)
        need = int(_check(lib().hfq_sap_table(t._h, None, None, 0, 0, None, 0)))
        out = np.empty(need)
        rows = int(_check(lib().hfq_sap_table(t._h, a.ctypes.data, None if b is None else b.ctypes.data, len(Pl_a),
                                              x_func, out.ctypes.data, need)))
        return out.reshape(9, rows).T

    def rs_exchange(self, cube):
        """src/sadatom/basis.cpp:314-420."""
        if getattr(self, "_rs", None) is None:
            raise ValueError("Primitive teis have not been computed!\n")
        return self.exchange(cube, _eng=self._rs)

    def exchange(self, cube, _eng=None):
        """Per-l exchange blocks (reference sign, -K), src/sadatom/basis.cpp:209-312."""
        N = self.Nrad()
        if len(cube) != self.lmax + 1:
            raise ValueError("Density matrix am does not match basis set!")
        n = (self.lmax + 1) * N
        P = np.zeros((n, n), order="F")
        for l, Pl in enumerate(cube):
            Pl = np.asarray(Pl)
            if Pl.shape != (N, N):
                raise ValueError("Density matrix does not match basis set!")
            P[l * N:(l + 1) * N, l * N:(l + 1) * N] = Pl
        K = (_eng or self._k).exchange(P)
        return [np.array(K[l * N:(l + 1) * N, l * N:(l + 1) * N]) for l in range(self.lmax + 1)]


class SadatomDFTGrid:
    """helfem::sadatom::dftgrid::DFTGrid (src/sadatom/dftgrid.h:124,127): radial-only quadrature of the spherically
    averaged atom.  Densities and Fock matrices are per-l cubes (lists of Nrad x Nrad matrices)."""

    def __init__(self, basis):
        self.basis = basis
        self._g = DFTGrid(basis._k, 1, 1)
        self.N = self._g.N

    def _dense(self, cube):
        N, L = self.basis.Nrad(), self.basis.lmax + 1
        if len(cube) != L:
            raise ValueError("Density matrix am does not match basis set!")
        P = np.zeros((L * N, L * N), order="F")
        for l, Pl in enumerate(cube):
            P[l * N:(l + 1) * N, l * N:(l + 1) * N] = np.asarray(Pl)
        return P

    def _cube(self, H):
        N, L = self.basis.Nrad(), self.basis.lmax + 1
        return [np.array(H[l * N:(l + 1) * N, l * N:(l + 1) * N]) for l in range(L)]

    def density(self, Pa, Pb=None, flags=0):
        return self._g.density(self._dense(Pa), None if Pb is None else self._dense(Pb), flags)

    def fxc(self, exc, vrho, vsigma=None, vtau=None, vlapl=None, beta=True):
        Ha, Hb, e = self._g.fxc(exc, vrho, vsigma, vtau, vlapl, beta)
        return self._cube(Ha), (None if Hb is None else self._cube(Hb)), e

    def eval_Fxc(self, x_func, c_func, P, Pb=None, beta=True, thr=1e-12):
        """(H cube or (Ha, Hb), Exc, Nel) as src/sadatom/dftgrid.cpp:505-653 (built-in Slater exchange only)."""
        H, exc, nel, _ = self._g.eval_Fxc(x_func, c_func, self._dense(P), None if Pb is None else self._dense(Pb), beta, thr)
        if Pb is None:
            return self._cube(H), exc, nel
        return (self._cube(H[0]), self._cube(H[1])), exc, nel


class TablesBasis(_BasisBase):
    """A basis whose caches were produced elsewhere (e.g. by an existing HelFEM build)."""

    def __init__(self, tables, device=0):
        super().__init__(device)
        self._tables = tables

    def _make_tables(self):
        return self._tables

    def set_absm_symmetric(self, sym):
        self._absm = bool(sym)
        if self._ctx is not None:
            _check(lib().hfq_set_absm_symmetric(self._ctx, int(self._absm)))


class DFTGrid:
    """DFT quadrature grid on the GPU: helfem::atomic::dftgrid::DFTGrid (src/atomic/dftgrid.h:139-160)
    for an atomic basis, helfem::diatomic::dftgrid_purem::PureMDFTGrid
    (src/diatomic/dftgrid_purem.h:140-160) for a diatomic basis (mang is ignored there).

    ``eval_Fxc`` has the reference's argument order; functionals other than the built-in Slater
    exchange are evaluated by the caller between ``density`` and ``fxc`` (libxc layout)."""

    GRAD, TAU, LAPL = 1, 2, 4

    def __init__(self, basis, lang, mang=1):
        self.basis = basis
        self.lang, self.mang = int(lang), int(mang)
        self._select()
        self.N = int(lib().hfq_grid_npoints(basis._context()))

    def _select(self):
        # several grids may live on one basis (the reference's diatomic driver holds a 3D and a pure-m grid,
        # src/diatomic/main.cpp:329-330): attaching an existing (lang, mang) selects it
        _check(lib().hfq_grid_attach(self.basis._context(), self.lang, self.mang))

    def density(self, Pa, Pb=None, flags=0):
        self._select()
        n = self.basis.Nbf()
        Pa = _fmat(Pa, n)
        pol = Pb is not None
        if pol:
            Pb = _fmat(Pb, n)
        ns = 2 if pol else 1
        out = {"rho": np.zeros((self.N, ns)), "w": np.zeros(self.N)}
        if flags & 1:
            out["sigma"] = np.zeros((self.N, 3 if pol else 1))
        if flags & 6:
            out["tau"] = np.zeros((self.N, ns))
        if flags & 4:
            out["lapl"] = np.zeros((self.N, ns))
        nel, ekin = ctypes.c_double(), ctypes.c_double()
        ptr = lambda k: out[k].ctypes.data if k in out else None
        _check(lib().hfq_grid_density(self.basis._context(), Pa.ctypes.data, n, Pb.ctypes.data if pol else None, n, flags,
                                      ptr("rho"), ptr("sigma"), ptr("tau"), ptr("lapl"), ptr("w"), ctypes.byref(nel),
                                      ctypes.byref(ekin)))
        out["Nel"], out["Ekin"] = nel.value, ekin.value
        self._pol = pol
        return out

    def fxc(self, exc, vrho, vsigma=None, vtau=None, vlapl=None, beta=True):
        self._select()
        n = self.basis.Nbf()
        Ha = np.zeros((n, n), order="F")
        Hb = np.zeros((n, n), order="F") if self._pol else None
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        exc, vrho, vsigma, vtau, vlapl = map(c, (exc, vrho, vsigma, vtau, vlapl))
        p = lambda a: None if a is None else a.ctypes.data
        flags = (1 if vsigma is not None else 0) | (2 if vtau is not None else 0) | (4 if vlapl is not None else 0)
        e = ctypes.c_double()
        _check(lib().hfq_grid_fxc(self.basis._context(), flags, int(beta), p(exc), p(vrho), p(vsigma), p(vtau), p(vlapl),
                                  Ha.ctypes.data, n, p(Hb), n, ctypes.byref(e)))
        return Ha, Hb, e.value

    def eval_Fxc(self, x_func, c_func, P, Pb=None, beta=True, thr=1e-12):
        """Returns (H or (Ha, Hb), Exc, Nel, Ekin)."""
        self._select()
        n = self.basis.Nbf()
        Pa = _fmat(P, n)
        pol = Pb is not None
        if pol:
            Pb = _fmat(Pb, n)
        Ha = np.zeros((n, n), order="F")
        Hb = np.zeros((n, n), order="F") if pol else None
        exc, nel, ekin = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        _check(lib().hfq_eval_fxc(self.basis._context(), x_func, c_func, Pa.ctypes.data, n, Pb.ctypes.data if pol else None,
                                  n, Ha.ctypes.data, n, Hb.ctypes.data if pol else None, n, ctypes.byref(exc),
                                  ctypes.byref(nel), ctypes.byref(ekin), int(beta), thr))
        return ((Ha, Hb) if pol else Ha), exc.value, nel.value, ekin.value


AtomicDFTGrid = DFTGrid
PureMDFTGrid = DFTGrid
