"""Drop-in for the hot-path part of the reference's ``fastfp/utils.py``.

* :func:`get_xCy` -- ``fastfp/utils.py:26-54``, run on the device through the C ABI.
* :func:`get_mats_fp` / :func:`get_mats_nmfp` -- ``fastfp/utils.py:57-101``: they only *collect*
  matrices from an ``enterprise`` PTA-like object (anything with ``get_phiinv``, ``get_TNT``,
  ``get_ndiag``, ``get_basis``); the return orders differ exactly as in the reference.
* :func:`compute_TNTs` / :func:`compute_sigmas` -- device-side ``T^T N^-1 T (+ diag(phiinv))`` from the
  raw basis (no reference counterpart, ``enterprise`` does this on the host).
* :func:`initialize_pta` -- ``fastfp/utils.py:104-163``: a pass-through to the third-party ``enterprise``
  packages where they are installed; it raises a clear error when they are absent.
"""
from __future__ import annotations

import numpy as np

from . import _cabi


def get_xCy(Nvec, T, sigma, x, y, device: int = 0):
    """``x^T C^-1 y`` with ``C = N + T B T^T`` (reference ``utils.py:49-54``).

    ``Nvec`` is the white-noise variance vector, as in the reference; beyond the reference
    (``utils.py:29-31`` excludes it) a block-diagonal ``N`` is accepted as a
    :class:`fastfp_b200.BlockNvec` or an enterprise ``ShermanMorrison`` object. Anything else
    raises ``ValueError``."""
    return np.float64(_cabi.xcy(Nvec, T, sigma, x, y, device=device))


def get_mats_fp(pta, noise):
    """``(Nvecs, Ts, sigmas)`` for the plain Fp-statistic (reference ``utils.py:57-78``)."""
    phiinvs = pta.get_phiinv(noise)
    TNTs = pta.get_TNT(noise)
    Nvecs = pta.get_ndiag(noise)
    Ts = pta.get_basis(noise)
    sigmas = [np.asarray(TNT, dtype=np.float64) + np.diag(np.asarray(phiinv, dtype=np.float64))
              for TNT, phiinv in zip(TNTs, phiinvs)]
    return Nvecs, Ts, sigmas


def get_mats_nmfp(pta, noise):
    """``(TNTs, Nvecs, Ts)`` for the noise-marginalised statistic (reference ``utils.py:81-101``;
    note the order differs from :func:`get_mats_fp`)."""
    TNTs = pta.get_TNT(noise)
    Nvecs = pta.get_ndiag(noise)
    Ts = pta.get_basis(noise)
    return TNTs, Nvecs, Ts


def compute_TNTs(Nvecs, Ts, device: int = 0):
    """``T^T N^-1 T`` per pulsar on the device -- what ``pta.get_TNT`` supplies to
    :func:`get_mats_nmfp` (reference ``utils.py:97``) -- for callers that hold only the raw
    ``(Nvec, T)``. A block-diagonal ``N`` (:class:`fastfp_b200.BlockNvec` / enterprise
    ``ShermanMorrison``) is handled by its Sherman-Morrison correction on the host."""
    from . import blockn

    out = []
    for Nvec, T in zip(Nvecs, Ts):
        if blockn.is_block(Nvec):
            T = np.asarray(T, dtype=np.float64)
            B = blockn.BlockNvec(np.asarray(Nvec._nvec, dtype=np.float64), Nvec._slices, Nvec._jvec)
            out.append(T.T @ B.solve(T))
        else:
            out.append(_cabi.tnt(Nvec, T, device=device))
    return out


def compute_sigmas(Nvecs, Ts, phiinvs, device: int = 0):
    """``Sigma = T^T N^-1 T + diag(phiinv)`` per pulsar (reference ``utils.py:76``) from the raw
    ``(Nvec, T, phiinv)``; the diagonal is added on the device for diagonal ``N``."""
    from . import blockn

    out = []
    for Nvec, T, ph in zip(Nvecs, Ts, phiinvs):
        if blockn.is_block(Nvec):
            out.append(compute_TNTs([Nvec], [T], device=device)[0] + np.diag(np.asarray(ph, dtype=np.float64)))
        else:
            out.append(_cabi.tnt(Nvec, T, phiinv=ph, device=device))
    return out


def initialize_pta(psrs, noise, inc_cp=True, rn_comps=30, gwb_comps=30, simple_wn=True, inc_ecorr=False,
                   select="backend"):
    """Model construction with ``enterprise`` (reference ``utils.py:104-163``; same arguments, defaults and signal
    order ``timing model + white noise + red noise (+ common red noise)``, so ``pta.get_basis`` column order
    matches the phi layouts of :class:`fastfp_b200.RN_container`). A thin pass-through to the third-party
    packages: it is only available where ``enterprise`` and ``enterprise_extensions`` are installed, and raises
    ``NotImplementedError`` with guidance otherwise. Nothing of the hot path depends on
    it: any object with ``get_phiinv`` / ``get_TNT`` / ``get_ndiag`` / ``get_basis`` works with ``get_mats_*``."""
    try:
        from enterprise.signals import gp_signals, parameter, signal_base, white_signals
        from enterprise_extensions import blocks, model_utils
    except ImportError as exc:
        raise NotImplementedError(
            "initialize_pta builds an enterprise PTA (third-party model construction) and needs the `enterprise` "
            "and `enterprise_extensions` packages, which are not installed here; construct the PTA with "
            "enterprise/fastfp and hand it to get_mats_fp / get_mats_nmfp, or pass the matrices directly"
        ) from exc
    span = model_utils.get_tspan(psrs)
    if simple_wn:      # EFAC fixed to 1: simulated data sets
        white = white_signals.MeasurementNoise(efac=parameter.Constant(1.0))
    else:              # per-backend EFAC/EQUAD, optionally ECORR as a Gaussian process on an epoch basis
        white = blocks.white_noise_block(inc_ecorr=bool(inc_ecorr), gp_ecorr=bool(inc_ecorr), select=select)
    signal = gp_signals.TimingModel(use_svd=True) + white + blocks.red_noise_block(Tspan=span, components=rn_comps)
    if inc_cp:
        signal = signal + blocks.common_red_noise_block(Tspan=span, components=gwb_comps)
    pta = signal_base.PTA([signal(psr) for psr in psrs])
    pta.set_default_params(noise)
    return pta
