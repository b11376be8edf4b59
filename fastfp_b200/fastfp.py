"""``FastFp`` -- drop-in for the reference's ``fastfp.fastfp.FastFp`` (``fastfp/fastfp.py:22-101``).

Same constructor and call signatures; the JAX/XLA program behind ``calculate_Fp`` is replaced
by the sm_100a sweep kernel of ``libfastfp_b200.so`` reached through the C ABI. Differences a
caller can see, all additive:

* ``fgw`` may be a scalar (reference semantics, returns a float) **or** a 1-D array of
  frequencies (returns ``(F,)``) -- what the reference obtains with
  ``jax.vmap(calculate_Fp, in_axes=(0, None, None, None))`` (``examples/run_fp.py:63``);
  :func:`fastfp_b200.vmap` keeps that spelling working too.
* ``fgw`` may be a float64 CUDA ``torch.Tensor``; the result is then a CUDA tensor and nothing
  crosses PCIe (the sweep is enqueued on torch's current stream).
* ``compute_Fp`` is an alias of ``calculate_Fp`` (the name ``enterprise_extensions``' ``FpStat``
  uses, cited at ``fastfp/fastfp.py:27-28``).

No validation beyond shapes is added: NaN/Inf propagate silently exactly as in the reference.
"""
from __future__ import annotations

import numpy as np

from . import _cabi


def _fingerprint(lists) -> tuple:
    """Content key of the caller's lists: shape + a 64-bit hash of EVERY byte of every array
    (``fastfp_hash64``: memory-bandwidth class, threaded for large arrays). The reference is a pure function
    of these arguments (``fastfp/fastfp.py:52``), so the cached device pack must be rebuilt whenever any
    entry changes -- including in-place edits, which an address- or sample-based key cannot see."""
    key, arrays, seeds, slots = [], [], [], []
    for lst in lists:
        key.append(len(lst))
        for a in lst:
            parts = [a._nvec, a._jvec, np.asarray([(s.start, s.stop) for s in a._slices], dtype=np.int64)] \
                if hasattr(a, "_nvec") else [a]
            for x in parts:
                x = np.ascontiguousarray(x)
                slots.append(len(key))
                seeds.append(len(key))
                arrays.append(x)
                key.append((x.shape, x.dtype.str))
    for slot, h in zip(slots, _cabi.hash64_many(arrays, seeds)):  # one call: all arrays share the hashing threads
        key[slot] = key[slot] + (h,)
    return tuple(key)


def _shape_key(lists) -> tuple:
    """Cheap structural key (lengths and shapes only): decides whether the cached pack can even be tried."""
    key = []
    for lst in lists:
        key.append(len(lst))
        for a in lst:
            key.append((np.shape(a._nvec), len(a._slices)) if hasattr(a, "_nvec") else np.shape(a))
    return tuple(key)


_HASHER = None


def _hasher():
    """One worker thread for the content hash (the ctypes calls release the GIL)."""
    global _HASHER
    if _HASHER is None:
        from concurrent.futures import ThreadPoolExecutor

        _HASHER = ThreadPoolExecutor(max_workers=1, thread_name_prefix="fastfp-hash")
    return _HASHER


class _PackCache:
    """The device pack is a CACHE of the caller's lists; the reference is a pure function of them. Every call
    hashes every byte of the lists. To keep that off the critical path the sweep is started on the cached pack
    right away while the hash runs on a worker thread (or, for asynchronous device-resident calls, on this thread
    after the launch); if the hash shows the inputs changed, the pack is rebuilt and the sweep repeated -- the
    caller never sees a result computed from stale data."""

    _pack = None
    _pack_key = None
    _pack_shape = None

    def invalidate(self):
        """Drop the cached device pack (the next call rebuilds it)."""
        if self._pack is not None:
            self._pack.close()
        self._pack, self._pack_key, self._pack_shape = None, None, None

    def _build_pack(self, lists):  # -> _cabi.Pack
        raise NotImplementedError

    def _ensure(self, lists, force=False, key=None):
        key = _fingerprint(lists) if key is None else key
        if force or self._pack is None or key != self._pack_key:
            if self._pack is not None:
                self._pack.close()
                self._pack = None
            self._pack = self._build_pack(lists)
            self._pack_key, self._pack_shape = key, _shape_key(lists)
        return self._pack

    def _run_verified(self, lists, run, asynchronous):
        """``run(pack)`` on a pack that is verified to match ``lists`` byte for byte."""
        if self._pack is None or _shape_key(lists) != self._pack_shape:
            return run(self._ensure(lists))
        if asynchronous:  # the launch returns at once: hash here while the GPU works
            res = run(self._pack)
            key = _fingerprint(lists)
        else:             # the call blocks until the result is on the host: hash on the worker meanwhile
            fut = _hasher().submit(_fingerprint, lists)
            try:
                res = run(self._pack)
            finally:
                key = fut.result()
        if key == self._pack_key:
            return res
        return run(self._ensure(lists, key=key))  # the inputs changed: rebuild, repeat


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class FastFp(_PackCache):
    """Fp-statistic (Ellis, Siemens & Creighton 2012) for a list of pulsars.

    :param psrs: objects with ``.toas`` and ``.residuals`` (seconds) -- all the reference reads
        (``fastfp/fastfp.py:44-45``)
    :param pta: stored and never used in compute, as in the reference (``fastfp.py:42``)
    :param device: CUDA device ordinal (extension; default 0 or ``LOCAL_RANK``)
    :param path: which kernel sweeps (extension): ``"auto"`` (default; also from ``FASTFP_B200_PATH``) = the
        INT8 tensor-core kernel (``tcgen05``, exact digit-plane product) when every pulsar fits its tile,
        else the fp64 DMMA kernel; ``"fp64"`` / ``"i8"`` force one (``"i8"`` raises if the pack cannot take it).
        Both meet the same parity bar; see DESIGN.md.
    """

    def __init__(self, psrs, pta=None, device=None, path=None):
        self.psrs = psrs
        self.pta = pta
        self.toas = [np.asarray(psr.toas, dtype=np.float64) for psr in psrs]
        self.residuals = [np.asarray(psr.residuals, dtype=np.float64) for psr in psrs]
        if device is None:
            import os

            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = int(device)
        import os as _os

        self.path = path if path is not None else _os.environ.get("FASTFP_B200_PATH", "auto")
        if self.path not in ("auto", "fp64", "i8", "prefer-i8"):
            raise ValueError("path must be 'auto', 'fp64', 'i8' or 'prefer-i8'")

    # -- packing (one-time, frequency-independent precompute on the device) -----------------
    def _build_pack(self, lists):
        from . import blockn

        Nvecs, Ts, sigmas = lists
        if any(blockn.is_block(N) for N in Nvecs):  # block-diagonal N (kernel ECORR)
            pack = _cabi.Pack.create_blockn(self.toas, self.residuals, Nvecs, Ts, sigmas, device=self.device)
        else:
            pack = _cabi.Pack.create_fp(self.toas, self.residuals, Nvecs, Ts, sigmas, device=self.device)
        if self.path == "prefer-i8":  # the tensor kernel where the pack can take it, silently the fp64 one otherwise
            try:
                pack.set_path("i8")
            except _cabi.FastFpError:
                pass
        elif self.path != "auto":
            pack.set_path(self.path)
        return pack

    def prepare(self, Nvecs, Ts, sigmas, force=False):
        """Upload and pre-reduce the per-pulsar arrays. The pack is cached and keyed on the full contents
        of the three lists (every byte is hashed on each call), so passing different arrays -- or the same
        arrays edited in place -- rebuilds it; ``force=True`` rebuilds unconditionally."""
        return self._ensure((Nvecs, Ts, sigmas), force=force)

    def __call__(self, fgw, Nvecs, Ts, sigmas):
        """Callable method (reference ``fastfp.py:47-49``)."""
        return self.calculate_Fp(fgw, Nvecs, Ts, sigmas)

    def calculate_Fp(self, fgw, Nvecs, Ts, sigmas):
        """Fp at ``fgw`` (reference ``fastfp.py:51-92``); see the module docstring for the
        batched forms of ``fgw``."""
        lists = (Nvecs, Ts, sigmas)
        if _is_cuda_tensor(fgw):
            import torch

            if fgw.dtype != torch.float64:
                raise TypeError("fgw tensor must be float64 (the reference enables jax x64)")
            if fgw.device.index != self.device:
                raise ValueError(f"fgw is on {fgw.device}, the pack on cuda:{self.device}")
            f = fgw.contiguous().reshape(-1)
            out = torch.empty(f.shape[0], dtype=torch.float64, device=f.device)
            stream = torch.cuda.current_stream(f.device).cuda_stream

            def run(pack):
                pack.fp_sweep((f.data_ptr(), f.shape[0]), out=out.data_ptr(), stream=stream)
                return out

            return self._run_verified(lists, run, asynchronous=True).reshape(fgw.shape)
        f = np.asarray(fgw, dtype=np.float64)
        res = self._run_verified(lists, lambda pack: pack.fp_sweep(f.reshape(-1)), asynchronous=False)
        return np.float64(res[0]) if f.ndim == 0 else res.reshape(f.shape)

    compute_Fp = calculate_Fp

    def per_pulsar_terms(self, fgw, Nvecs, Ts, sigmas):
        """``0.5 * N^T M^-1 N`` per pulsar, ``(P, F)`` -- the summands of ``fastfp.py:90``."""
        pack = self.prepare(Nvecs, Ts, sigmas)
        return pack.fp_sweep(np.atleast_1d(np.asarray(fgw, dtype=np.float64)), terms=True)

    # pytree protocol of the reference (fastfp.py:94-101), kept so code that flattens the
    # object keeps working; there is no tracing here.
    def tree_flatten(self):
        return (), (self.psrs, self.pta)

    @classmethod
    def tree_unflatten(cls, aux_data, children):
        return cls(*aux_data, *children)
