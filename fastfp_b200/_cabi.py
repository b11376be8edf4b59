"""ctypes binding of ``libfastfp_b200.so`` (the C ABI of ``include/fastfp_b200.h``).

The product path is the CUDA library; there is deliberately **no CPU fallback**: if the
library is missing, or no CUDA device is visible, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libfastfp_b200.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_double_pp = C.POINTER(c_double_p)

FREQS_ON_DEVICE = 1
OUT_ON_DEVICE = 2
PARAMS_ON_DEVICE = 4

# every symbol include/fastfp_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fastfp_last_error": (C.c_char_p, []),
    "fastfp_version": (C.c_int, []),
    "fastfp_device_count": (C.c_int, []),
    "fastfp_pack_create": (
        C.c_int,
        [C.c_int, C.c_int, c_int64_p, c_int64_p, c_double_pp, c_double_pp, c_double_pp, c_double_pp,
         c_double_pp, C.c_void_p, C.POINTER(C.c_void_p)],
    ),
    "fastfp_fp_sweep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "fastfp_fp_terms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "fastfp_fe_sweep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                  C.c_int, C.c_void_p]),
    "fastfp_nmfp_pack_create": (
        C.c_int,
        [C.c_int, C.c_int, c_int64_p, c_int64_p, c_double_pp, c_double_pp, c_double_pp, c_double_pp,
         c_double_pp, c_int64_p, c_double_pp, C.c_void_p, C.POINTER(C.c_void_p)],
    ),
    "fastfp_nmfp_tile_sizes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fastfp_nmfp_stage_a": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fastfp_nmfp_stage_b": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p],
    ),
    "fastfp_nmfp_sweep": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p],
    ),
    "fastfp_powerlaw_phiinv": (
        C.c_int,
        [C.c_void_p, c_double_pp, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "fastfp_sweep_chunk_toas": (C.c_int, [C.c_int64, C.c_int]),
    "fastfp_pack_create_blockn": (
        C.c_int,
        [C.c_int, C.c_int, c_int64_p, c_int64_p, c_double_pp, c_double_pp, c_double_pp, c_double_pp, c_double_pp,
         c_double_pp, C.POINTER(C.POINTER(C.c_int32)), c_double_pp, C.POINTER(C.POINTER(C.c_ubyte)), c_int64_p,
         c_double_pp, C.c_void_p, C.POINTER(C.c_void_p)],
    ),
    "fastfp_pack_destroy": (None, [C.c_void_p]),
    "fastfp_pack_bytes": (C.c_int64, [C.c_void_p]),
    "fastfp_pack_num_pulsars": (C.c_int, [C.c_void_p]),
    "fastfp_pack_mvar_total": (C.c_int64, [C.c_void_p]),
    "fastfp_pack_set_path": (C.c_int, [C.c_void_p, C.c_int]),
    "fastfp_pack_path": (C.c_int, [C.c_void_p]),
    "fastfp_pack_factor_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "fastfp_hash64": (C.c_uint64, [C.c_void_p, C.c_int64, C.c_uint64]),
    "fastfp_hash64_many": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "fastfp_kernel_launches": (C.c_int64, []),
    "fastfp_xcy": (
        C.c_int,
        [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "fastfp_tnt": (
        C.c_int,
        [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "fastfp_nmfp_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "fastfp_nmfp_stage_ms": (C.c_int, [C.c_void_p, c_double_p]),
    "fastfp_xcy_blockn": (
        C.c_int,
        [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    "fastfp_fp64_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]),
}


class FastFpError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise FastFpError(
            f"{_LIB_PATH} is missing: build it with `python -m fastfp_b200.build` "
            "(there is no CPU fallback for the Fp hot path)"
        )
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().fastfp_last_error()
        raise FastFpError(f"fastfp_b200 error {rc}: {msg.decode() if msg else '?'}")


def require_device() -> int:
    n = load().fastfp_device_count()
    if n < 1:
        raise FastFpError("no CUDA device visible: the Fp hot path only runs on a B200 (no CPU fallback)")
    return n


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr_array(arrs: Sequence[np.ndarray]):
    arr = (c_double_p * len(arrs))()
    for i, a in enumerate(arrs):
        arr[i] = a.ctypes.data_as(c_double_p)
    return arr


def _int64_array(vals: Sequence[int]):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def _vp(x) -> C.c_void_p:
    """host ndarray / integer device address -> void*"""
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(int(x))


def _check_lists(toas, residuals, Nvecs, Ts, mats, what):
    P = len(toas)
    if P < 1 or not (len(residuals) == len(Nvecs) == len(Ts) == len(mats) == P):
        raise ValueError(f"toas, residuals, Nvecs, Ts and {what} must be lists of equal length P >= 1")
    toas, residuals, Nvecs = [as_f64(a) for a in toas], [as_f64(a) for a in residuals], [as_f64(a) for a in Nvecs]
    Ts, mats = [as_f64(a) for a in Ts], [as_f64(a) for a in mats]
    n, m = [], []
    for p in range(P):
        if Ts[p].ndim != 2:
            raise ValueError(f"Ts[{p}] must be 2-D (ntoa, nbasis)")
        np_, mp_ = Ts[p].shape
        if toas[p].shape != (np_,) or residuals[p].shape != (np_,) or Nvecs[p].shape != (np_,):
            raise ValueError(
                f"pulsar {p}: toas/residuals/Nvec must have shape ({np_},) to match Ts "
                "(a diagonal N is required, as in the reference's get_xCy)"
            )
        if mats[p].shape != (mp_, mp_):
            raise ValueError(f"pulsar {p}: {what} must have shape ({mp_}, {mp_})")
        n.append(np_)
        m.append(mp_)
    return P, n, m, toas, residuals, Nvecs, Ts, mats


def hash64(a: np.ndarray, seed: int = 0) -> int:
    """64-bit hash of every byte of a C-contiguous host array (``fastfp_hash64``)."""
    return int(load().fastfp_hash64(C.c_void_p(a.ctypes.data), a.nbytes, C.c_uint64(seed & (2**64 - 1))))


def hash64_many(arrays, seeds) -> list:
    """``hash64`` of each C-contiguous host array in one library call (one thread pool over all of them)."""
    n = len(arrays)
    if n == 0:
        return []
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrays])
    sizes = (C.c_int64 * n)(*[a.nbytes for a in arrays])
    sd = (C.c_uint64 * n)(*[s & (2**64 - 1) for s in seeds])
    out = (C.c_uint64 * n)()
    check(load().fastfp_hash64_many(ptrs, sizes, n, sd, out))
    return list(out)


class Pack:
    """Owner of one ``fastfp_pack_t*``: the device-resident packed pulsar array."""

    def __init__(self, handle, P: int, device: int, nmfp: bool, n, m):
        self._h, self.P, self.device, self.nmfp = handle, P, device, nmfp
        self.n, self.m = list(n), list(m)
        self._warn_if_not_spd()

    PATHS = {"auto": 0, "fp64": 1, "i8": 2}

    def set_path(self, path: str) -> None:
        """Kernel of the plain-Fp sweep: "auto" (default: the INT8 tensor-core kernel when the pack has digit
        planes), "fp64" (the DMMA kernel) or "i8" (raises unless every pulsar fits); ``path`` then
        reads "fp64", "i8" or "mixed" (auto, with the pulsars the tensor kernel does not take on the fp64 kernel)."""
        check(load().fastfp_pack_set_path(self._h, self.PATHS[path]))

    @property
    def path(self) -> str:
        return {1: "fp64", 2: "i8", 3: "mixed"}[load().fastfp_pack_path(self._h)]

    def factor_info(self):
        """Per-pulsar status of the one-time Cholesky (0 = fine, j+1 = pivot j not positive)."""
        info = (C.c_int32 * self.P)()
        rc = load().fastfp_pack_factor_info(self._h, info)
        if rc < 0:
            check(rc)
        return list(info)

    def _warn_if_not_spd(self):
        bad = [(p, v) for p, v in enumerate(self.factor_info()) if v]
        if bad:
            import warnings

            what = "Sigma" if not self.nmfp else "the draw-independent block of Sigma"
            warnings.warn(
                "fastfp_b200: " + what + " is not numerically symmetric positive definite for pulsar(s) "
                + ", ".join(f"{p} (pivot {v - 1})" for p, v in bad)
                + "; their terms are NaN. The sweep path factorises Sigma = L L^T (lower triangle); the "
                "reference's general LU solve is only available through get_xCy.", RuntimeWarning, stacklevel=3)

    @classmethod
    def create_fp(cls, toas, residuals, Nvecs, Ts, sigmas, device: int = 0, stream: int = 0) -> "Pack":
        P, n, m, toas, residuals, Nvecs, Ts, sigmas = _check_lists(toas, residuals, Nvecs, Ts, sigmas, "sigmas")
        lib = load()
        require_device()
        h = C.c_void_p()
        check(
            lib.fastfp_pack_create(
                device, P, _int64_array(n), _int64_array(m), _ptr_array(toas), _ptr_array(residuals),
                _ptr_array(Nvecs), _ptr_array(Ts), _ptr_array(sigmas), C.c_void_p(stream), C.byref(h),
            )
        )
        return cls(h, P, device, False, n, m)

    @classmethod
    def create_nmfp(cls, toas, residuals, Nvecs, Ts, TNTs, m_fix, phiinv_fix, device: int = 0, stream: int = 0):
        P, n, m, toas, residuals, Nvecs, Ts, TNTs = _check_lists(toas, residuals, Nvecs, Ts, TNTs, "TNTs")
        if len(m_fix) != P or len(phiinv_fix) != P:
            raise ValueError("m_fix and phiinv_fix must have one entry per pulsar")
        pf = []
        for p in range(P):
            if not 0 <= int(m_fix[p]) <= m[p]:
                raise ValueError(f"pulsar {p}: m_fix out of range")
            a = as_f64(phiinv_fix[p]).reshape(-1)
            if a.shape[0] != int(m_fix[p]):
                raise ValueError(f"pulsar {p}: phiinv_fix must have m_fix entries")
            pf.append(a if a.size else np.zeros(1))
        lib = load()
        require_device()
        h = C.c_void_p()
        check(
            lib.fastfp_nmfp_pack_create(
                device, P, _int64_array(n), _int64_array(m), _ptr_array(toas), _ptr_array(residuals),
                _ptr_array(Nvecs), _ptr_array(Ts), _ptr_array(TNTs), _int64_array(m_fix), _ptr_array(pf),
                C.c_void_p(stream), C.byref(h),
            )
        )
        return cls(h, P, device, True, n, m)

    @classmethod
    def create_blockn(cls, toas, residuals, Nvecs, Ts, mats, m_fix=None, phiinv_fix=None, device: int = 0,
                      stream: int = 0) -> "Pack":
        """Pack with a block-diagonal N (kernel ECORR) for at least one pulsar. ``mats`` are the
        sigmas (plain Fp, ``m_fix is None``) or the TNTs (nmfp), formed with the block N."""
        from . import blockn

        P = len(toas)
        if P < 1 or not (len(residuals) == len(Nvecs) == len(Ts) == len(mats) == P):
            raise ValueError("toas, residuals, Nvecs, Ts and the matrices must be lists of equal length P >= 1")
        Ts = [as_f64(T) for T in Ts]
        mats = [as_f64(a) for a in mats]
        prep, n, m = [], [], []
        lib = load()
        for p in range(P):
            if Ts[p].ndim != 2 or mats[p].shape != (Ts[p].shape[1],) * 2:
                raise ValueError(f"pulsar {p}: Ts must be (ntoa, nbasis) and the matrix (nbasis, nbasis)")
            ci = lib.fastfp_sweep_chunk_toas(Ts[p].shape[1], 1)
            if ci <= 0:
                raise ValueError(f"pulsar {p}: basis width {Ts[p].shape[1]} is not supported with a block-diagonal N")
            d = blockn.prepare(toas[p], residuals[p], Nvecs[p], Ts[p], ci)
            prep.append(d)
            n.append(d["toas"].shape[0])
            m.append(Ts[p].shape[1])
        require_device()
        pf = None
        if m_fix is not None:
            pf = [as_f64(a).reshape(-1) if len(a) else np.zeros(1) for a in phiinv_fix]
        i32pp = (C.POINTER(C.c_int32) * P)(*[d["slot_idx"].ctypes.data_as(C.POINTER(C.c_int32)) for d in prep])
        u8pp = (C.POINTER(C.c_ubyte) * P)(*[d["done_mask"].ctypes.data_as(C.POINTER(C.c_ubyte)) for d in prep])
        h = C.c_void_p()
        check(
            lib.fastfp_pack_create_blockn(
                device, P, _int64_array(n), _int64_array(m), _ptr_array([d["toas"] for d in prep]),
                _ptr_array([d["res"] for d in prep]), _ptr_array([d["res_w"] for d in prep]),
                _ptr_array([d["Nvec"] for d in prep]), _ptr_array([d["T"] for d in prep]), _ptr_array(mats),
                i32pp, _ptr_array([d["slot_val"] for d in prep]), u8pp,
                _int64_array(m_fix) if m_fix is not None else None, _ptr_array(pf) if pf is not None else None,
                C.c_void_p(stream), C.byref(h),
            )
        )
        return cls(h, P, device, m_fix is not None, n, m)

    # -- sweeps -------------------------------------------------------------------------
    def fp_sweep(self, freqs, out=None, stream: int = 0, terms: bool = False):
        """``freqs``: host ndarray or ``(device_address, F)``; ``out``: None (a host array is
        returned), a host ndarray, or an integer device address."""
        lib = load()
        flags = 0
        if isinstance(freqs, tuple):
            fptr, F = freqs
            flags |= FREQS_ON_DEVICE
        else:
            freqs = as_f64(freqs).reshape(-1)
            fptr, F = freqs, freqs.shape[0]
        ret = None
        if out is None:
            ret = np.empty((self.P, F) if terms else (F,), dtype=np.float64)
            optr = ret
        elif isinstance(out, np.ndarray):
            optr = out
        else:
            optr = out
            flags |= OUT_ON_DEVICE
        fn = lib.fastfp_fp_terms if terms else lib.fastfp_fp_sweep
        check(fn(self._h, _vp(fptr), F, _vp(optr), flags, C.c_void_p(stream)))
        return ret

    def fe_sweep(self, freqs, fplus, fcross, out=None, stream: int = 0):
        """Fe-statistic for ``S`` sky positions: ``fplus``, ``fcross`` host arrays ``(S, P)``; returns / fills
        ``(S, F)``. ``freqs`` / ``out`` as in :meth:`fp_sweep`."""
        lib = load()
        fplus, fcross = as_f64(fplus), as_f64(fcross)
        if fplus.ndim != 2 or fplus.shape != fcross.shape or fplus.shape[1] != self.P:
            raise ValueError("fplus and fcross must both have shape (n_sky, n_pulsars)")
        S = fplus.shape[0]
        flags = 0
        if isinstance(freqs, tuple):
            fptr, F = freqs
            flags |= FREQS_ON_DEVICE
        else:
            freqs = as_f64(freqs).reshape(-1)
            fptr, F = freqs, freqs.shape[0]
        ret = None
        if out is None:
            ret = np.empty((S, F), dtype=np.float64)
            optr = ret
        elif isinstance(out, np.ndarray):
            optr = out
        else:
            optr = out
            flags |= OUT_ON_DEVICE
        check(lib.fastfp_fe_sweep(self._h, _vp(fptr), F, _vp(fplus), _vp(fcross), S, _vp(optr), flags, C.c_void_p(stream)))
        return ret

    def nmfp_sweep(self, freqs, phiinv_var, D: int, out=None, stream: int = 0):
        lib = load()
        flags = 0
        if isinstance(freqs, tuple):
            fptr, F = freqs
            flags |= FREQS_ON_DEVICE
        else:
            freqs = as_f64(freqs).reshape(-1)
            fptr, F = freqs, freqs.shape[0]
        if isinstance(phiinv_var, np.ndarray):
            phiinv_var = as_f64(phiinv_var)
            pptr = phiinv_var
        else:
            pptr = phiinv_var
            flags |= PARAMS_ON_DEVICE
        ret = None
        if out is None:
            ret = np.empty((D, F), dtype=np.float64)
            optr = ret
        elif isinstance(out, np.ndarray):
            optr = out
        else:
            optr = out
            flags |= OUT_ON_DEVICE
        check(lib.fastfp_nmfp_sweep(self._h, _vp(fptr), F, _vp(pptr), D, _vp(optr), flags, C.c_void_p(stream)))
        return ret

    # the two halves of nmfp_sweep as separate calls (device pointers only): parallel.py shards them in two dimensions
    def nmfp_tile_sizes(self):
        """doubles per 32-frequency tile (all pulsars) of the two stage-A outputs"""
        z, a = C.c_int64(0), C.c_int64(0)
        check(load().fastfp_nmfp_tile_sizes(self._h, C.byref(z), C.byref(a)))
        return int(z.value), int(a.value)

    def nmfp_stage_a(self, freqs_ptr: int, F: int, z_ptr: int, a_ptr: int, stream: int = 0):
        check(load().fastfp_nmfp_stage_a(self._h, C.c_void_p(freqs_ptr), F, C.c_void_p(z_ptr), C.c_void_p(a_ptr),
                                         C.c_void_p(stream)))

    def nmfp_stage_b(self, freqs_ptr: int, F: int, z_ptr: int, a_ptr: int, tiles_per_block: int, phiinv_ptr: int,
                     D: int, out_ptr: int, stream: int = 0):
        check(load().fastfp_nmfp_stage_b(self._h, C.c_void_p(freqs_ptr), F, C.c_void_p(z_ptr), C.c_void_p(a_ptr),
                                         tiles_per_block, C.c_void_p(phiinv_ptr), D, C.c_void_p(out_ptr),
                                         C.c_void_p(stream)))

    def powerlaw_phiinv(self, Ffreqs, log10_A, gamma, curn_Ffreqs, curn_log10_A, curn_gamma, out_dev, stream=0):
        """Device-side ``get_phiinv`` of the varying block for D draws (host params in)."""
        lib = load()
        Ff = [as_f64(a) for a in Ffreqs]
        A, G = as_f64(log10_A), as_f64(gamma)
        D = A.shape[0]
        if curn_Ffreqs is not None and len(curn_Ffreqs):
            cf, cA, cG = as_f64(curn_Ffreqs), as_f64(curn_log10_A), as_f64(curn_gamma)
            args = (_vp(cf), cf.shape[0], _vp(cA), _vp(cG))
        else:
            args = (C.c_void_p(0), 0, C.c_void_p(0), C.c_void_p(0))
        check(
            lib.fastfp_powerlaw_phiinv(
                self._h, _ptr_array(Ff), _vp(A), _vp(G), D, *args, C.c_void_p(int(out_dev)), C.c_void_p(stream)
            )
        )

    def stage_timing(self, enable: bool = True) -> None:
        """Bracket the three nmfp stages of later sweeps with CUDA events (measurement aid)."""
        check(load().fastfp_nmfp_stage_timing(self._h, int(bool(enable))))

    def stage_ms(self):
        """``(stage A, factor, stage B)`` milliseconds of the last timed nmfp sweep."""
        out = (C.c_double * 3)()
        check(load().fastfp_nmfp_stage_ms(self._h, out))
        return tuple(out)

    @property
    def mvar_total(self) -> int:
        return int(load().fastfp_pack_mvar_total(self._h))

    @property
    def nbytes(self) -> int:
        return int(load().fastfp_pack_bytes(self._h))

    def close(self):
        if self._h:
            load().fastfp_pack_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tnt(Nvec, T, phiinv=None, device: int = 0, stream: int = 0) -> np.ndarray:
    """``T^T N^-1 T`` (``+ diag(phiinv)``) on the device; diagonal ``N`` given as the variance vector."""
    Nvec, T = as_f64(Nvec), as_f64(T)
    if T.ndim != 2 or Nvec.shape != (T.shape[0],):
        raise ValueError("tnt: shapes must be Nvec (n,), T (n, m)")
    ph = None
    if phiinv is not None:
        ph = as_f64(phiinv)
        if ph.shape != (T.shape[1],):
            raise ValueError("tnt: phiinv must have shape (m,)")
    lib = load()
    require_device()
    out = np.empty((T.shape[1], T.shape[1]))
    check(lib.fastfp_tnt(device, T.shape[0], T.shape[1], _vp(Nvec), _vp(T), _vp(ph) if ph is not None else None,
                         _vp(out), C.c_void_p(stream)))
    return out


def xcy(Nvec, T, sigma, x, y, device: int = 0, stream: int = 0) -> float:
    from . import blockn

    if blockn.is_block(Nvec):  # block-diagonal N: Sherman-Morrison on the host, same kernel
        nvec, T, sigma, x, y = as_f64(Nvec._nvec), as_f64(T), as_f64(sigma), as_f64(x), as_f64(y)
        if T.ndim != 2 or nvec.shape != (T.shape[0],) or x.shape != nvec.shape or y.shape != nvec.shape \
                or sigma.shape != (T.shape[1],) * 2:
            raise ValueError("get_xCy: shapes must be Nvec (n,), T (n,m), sigma (m,m), x (n,), y (n,)")
        B = blockn.BlockNvec(nvec, Nvec._slices, Nvec._jvec)
        xw, yw = as_f64(B.solve(x) * nvec), as_f64(B.solve(y) * nvec)
        lib = load()
        require_device()
        out = np.empty(1)
        check(lib.fastfp_xcy_blockn(device, T.shape[0], T.shape[1], _vp(nvec), _vp(T), _vp(sigma), _vp(x), _vp(xw),
                                    _vp(yw), _vp(out), C.c_void_p(stream)))
        return float(out[0])
    Nvec, T, sigma, x, y = as_f64(Nvec), as_f64(T), as_f64(sigma), as_f64(x), as_f64(y)
    if T.ndim != 2:
        raise ValueError("get_xCy: T must be 2-D (ntoa, nbasis)")
    n, m = T.shape
    if Nvec.shape != (n,) or x.shape != (n,) or y.shape != (n,) or sigma.shape != (m, m):
        raise ValueError("get_xCy: shapes must be Nvec (n,), T (n,m), sigma (m,m), x (n,), y (n,)")
    lib = load()
    require_device()
    out = np.empty(1)
    check(lib.fastfp_xcy(device, n, m, _vp(Nvec), _vp(T), _vp(sigma), _vp(x), _vp(y), _vp(out), C.c_void_p(stream)))
    return float(out[0])


def fp64_peak(kind: int = 0, iters: int = 20000, device: int = 0):
    lib = load()
    require_device()
    tf, ms = C.c_double(), C.c_double()
    check(lib.fastfp_fp64_peak(device, kind, iters, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def kernel_launches() -> int:
    return int(load().fastfp_kernel_launches())
