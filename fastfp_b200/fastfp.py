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
    key = []
    for lst in lists:
        key.append(len(lst))
        for a in lst:
            parts = [a._nvec, a._jvec, np.asarray([(s.start, s.stop) for s in a._slices], dtype=np.int64)] \
                if hasattr(a, "_nvec") else [a]
            for x in parts:
                x = np.ascontiguousarray(x)
                key.append((x.shape, x.dtype.str, _cabi.hash64(x, seed=len(key))))
    return tuple(key)


def _is_cuda_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class FastFp(object):
    """Fp-statistic (Ellis, Siemens & Creighton 2012) for a list of pulsars.

    :param psrs: objects with ``.toas`` and ``.residuals`` (seconds) -- all the reference reads
        (``fastfp/fastfp.py:44-45``)
    :param pta: stored and never used in compute, as in the reference (``fastfp.py:42``)
    :param device: CUDA device ordinal (extension; default 0 or ``LOCAL_RANK``)
    """

    def __init__(self, psrs, pta=None, device=None):
        self.psrs = psrs
        self.pta = pta
        self.toas = [np.asarray(psr.toas, dtype=np.float64) for psr in psrs]
        self.residuals = [np.asarray(psr.residuals, dtype=np.float64) for psr in psrs]
        if device is None:
            import os

            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = int(device)
        self._pack = None
        self._pack_key = None

    # -- packing (one-time, frequency-independent precompute on the device) -----------------
    def invalidate(self):
        """Drop the cached device pack (the next call rebuilds it)."""
        if self._pack is not None:
            self._pack.close()
        self._pack, self._pack_key = None, None

    def prepare(self, Nvecs, Ts, sigmas, force=False):
        """Upload and pre-reduce the per-pulsar arrays. The pack is cached and keyed on the full contents
        of the three lists (every byte is hashed on each call), so passing different arrays -- or the same
        arrays edited in place -- rebuilds it; ``force=True`` rebuilds unconditionally."""
        key = _fingerprint((Nvecs, Ts, sigmas))
        if force or self._pack is None or key != self._pack_key:
            if self._pack is not None:
                self._pack.close()
            from . import blockn

            if any(blockn.is_block(N) for N in Nvecs):  # block-diagonal N (kernel ECORR)
                self._pack = _cabi.Pack.create_blockn(self.toas, self.residuals, Nvecs, Ts, sigmas, device=self.device)
            else:
                self._pack = _cabi.Pack.create_fp(self.toas, self.residuals, Nvecs, Ts, sigmas, device=self.device)
            self._pack_key = key
        return self._pack

    def __call__(self, fgw, Nvecs, Ts, sigmas):
        """Callable method (reference ``fastfp.py:47-49``)."""
        return self.calculate_Fp(fgw, Nvecs, Ts, sigmas)

    def calculate_Fp(self, fgw, Nvecs, Ts, sigmas):
        """Fp at ``fgw`` (reference ``fastfp.py:51-92``); see the module docstring for the
        batched forms of ``fgw``."""
        pack = self.prepare(Nvecs, Ts, sigmas)
        if _is_cuda_tensor(fgw):
            import torch

            if fgw.dtype != torch.float64:
                raise TypeError("fgw tensor must be float64 (the reference enables jax x64)")
            if fgw.device.index != self.device:
                raise ValueError(f"fgw is on {fgw.device}, the pack on cuda:{self.device}")
            f = fgw.contiguous().reshape(-1)
            out = torch.empty(f.shape[0], dtype=torch.float64, device=f.device)
            stream = torch.cuda.current_stream(f.device).cuda_stream
            pack.fp_sweep((f.data_ptr(), f.shape[0]), out=out.data_ptr(), stream=stream)
            return out.reshape(fgw.shape)
        f = np.asarray(fgw, dtype=np.float64)
        res = pack.fp_sweep(f.reshape(-1))
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
