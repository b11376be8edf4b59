"""``FastFe`` -- the Fe-statistic on the B200 engine.

The reference lists the Fe-statistic as a to-do (``README.md:23``); ``enterprise_extensions.frequentist.FeStat`` is the
implementation users have today. Fe (Ellis, Siemens & Creighton 2012) is the coherent Earth-term counterpart of Fp: for
a sky position ``(gwtheta, gwphi)`` the four templates of pulsar ``p`` are ``[F+ s, F+ c, Fx s, Fx c]`` with the same
``s, c = sin, cos(((2 pi) f) t)`` as ``FastFp.calculate_Fp`` (``fastfp/fastfp.py:78-79``) and the antenna patterns
``F+_p, Fx_p``; the statistic is ``1/2 N^T M^-1 N`` with the 4-vector ``N`` and the 4x4 matrix ``M`` summed over
pulsars. Every entry is one of the five inner products the Fp sweep already forms per (pulsar, frequency), so a sky
scan costs one sweep plus a small combine kernel (``fastfp_fe_sweep``).
"""
from __future__ import annotations

import numpy as np

from .fastfp import FastFp, _is_cuda_tensor


def antenna_pattern(pos, gwtheta, gwphi):
    """``(F+, Fx)`` of a pulsar at unit vector ``pos`` for a source at polar angle ``gwtheta`` and azimuth ``gwphi``
    (the convention of ``enterprise.signals.utils.create_gw_antenna_pattern``). ``gwtheta`` / ``gwphi`` may be
    arrays (broadcast); ``pos`` is ``(3,)`` or ``(P, 3)`` -> result ``(..., P)``."""
    pos = np.atleast_2d(np.asarray(pos, dtype=np.float64))
    th, ph = np.broadcast_arrays(np.asarray(gwtheta, dtype=np.float64), np.asarray(gwphi, dtype=np.float64))
    m = np.stack((np.sin(ph), -np.cos(ph), np.zeros_like(ph)), axis=-1)
    n = np.stack((-np.cos(th) * np.cos(ph), -np.cos(th) * np.sin(ph), np.sin(th)), axis=-1)
    om = np.stack((-np.sin(th) * np.cos(ph), -np.sin(th) * np.sin(ph), -np.cos(th)), axis=-1)
    mp, npos, op = m @ pos.T, n @ pos.T, om @ pos.T
    fplus = 0.5 * (mp ** 2 - npos ** 2) / (1.0 + op)
    fcross = mp * npos / (1.0 + op)
    return fplus, fcross


class FastFe(FastFp):
    """Fe-statistic for a list of pulsars; the pulsars additionally need ``.pos`` (unit vector, as
    ``enterprise.pulsar.Pulsar`` provides). Same packing and caching as :class:`FastFp`."""

    def __init__(self, psrs, pta=None, device=None, path=None):
        super().__init__(psrs, pta=pta, device=device, path=path)
        self.pos = np.stack([np.asarray(psr.pos, dtype=np.float64) for psr in psrs])

    def calculate_Fe(self, fgw, gwtheta, gwphi, Nvecs, Ts, sigmas):
        """Fe at frequency ``fgw`` (scalar or ``(F,)``, host array or CUDA tensor) and sky position(s)
        ``gwtheta``, ``gwphi`` (scalars or ``(S,)``): returns a scalar, ``(F,)``, ``(S,)`` or ``(S, F)``."""
        th, ph = np.asarray(gwtheta, dtype=np.float64), np.asarray(gwphi, dtype=np.float64)
        sky_batched = th.ndim > 0 or ph.ndim > 0
        th, ph = np.broadcast_arrays(np.atleast_1d(th), np.atleast_1d(ph))
        fplus, fcross = antenna_pattern(self.pos, th, ph)  # (S, P)
        lists = (Nvecs, Ts, sigmas)
        if _is_cuda_tensor(fgw):
            import torch

            if fgw.dtype != torch.float64:
                raise TypeError("fgw tensor must be float64")
            if fgw.device.index != self.device:
                raise ValueError(f"fgw is on {fgw.device}, the pack on cuda:{self.device}")
            f = fgw.contiguous().reshape(-1)
            out = torch.empty((fplus.shape[0], f.shape[0]), dtype=torch.float64, device=f.device)
            stream = torch.cuda.current_stream(f.device).cuda_stream

            def run(pack):
                pack.fe_sweep((f.data_ptr(), f.shape[0]), fplus, fcross, out=out.data_ptr(), stream=stream)
                return out

            res = self._run_verified(lists, run, asynchronous=True)
            res = res if fgw.ndim else res[:, 0]
            return res if sky_batched else res[0]
        f = np.asarray(fgw, dtype=np.float64)
        res = self._run_verified(lists, lambda pack: pack.fe_sweep(f.reshape(-1), fplus, fcross), asynchronous=False)
        if f.ndim == 0:
            res = res[:, 0]
        if not sky_batched:
            res = res[0]
        return np.float64(res) if np.ndim(res) == 0 else res

    compute_Fe = calculate_Fe
