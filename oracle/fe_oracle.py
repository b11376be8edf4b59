"""CPU ORACLE for the Fe-statistic -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED BY THE REFERENCE: gabefreedman/fastfp has no Fe implementation (it is a README to-do,
``README.md:23``), so there is no reference output to pin against. This module restates the published statistic
(Ellis, Siemens & Creighton 2012, ApJ 756:175, section 3; the structure of ``enterprise_extensions.frequentist.FeStat``,
which is not installed here) on top of the reference's OWN inner product ``get_xCy`` (``fastfp/utils.py:49-54``, restated
in ``oracle/fp_oracle.py`` and pinned there), with the reference's basis conventions for the sin/cos templates
(``fastfp/fastfp.py:77-79``: phase ``((2 pi) f) t``, prefactor ``f^(-1/3)``).
"""
from __future__ import annotations

import numpy as np

from .fp_oracle import get_xCy


def antenna_pattern(pos, gwtheta, gwphi):
    """F+, Fx for a pulsar at unit vector ``pos``, source at (gwtheta, gwphi): the geometric definitions
    m = (sin phi, -cos phi, 0), n = (-cos theta cos phi, -cos theta sin phi, sin theta), Omega = -(source direction);
    F+ = ((m.p)^2 - (n.p)^2) / (2 (1 + Omega.p)), Fx = (m.p)(n.p) / (1 + Omega.p)."""
    m = np.array([np.sin(gwphi), -np.cos(gwphi), 0.0])
    n = np.array([-np.cos(gwtheta) * np.cos(gwphi), -np.cos(gwtheta) * np.sin(gwphi), np.sin(gwtheta)])
    om = np.array([-np.sin(gwtheta) * np.cos(gwphi), -np.sin(gwtheta) * np.sin(gwphi), -np.cos(gwtheta)])
    fplus = 0.5 * (np.dot(m, pos) ** 2 - np.dot(n, pos) ** 2) / (1 + np.dot(om, pos))
    fcross = np.dot(m, pos) * np.dot(n, pos) / (1 + np.dot(om, pos))
    return fplus, fcross


def calculate_Fe(fgw, gwtheta, gwphi, toas, residuals, positions, Nvecs, Ts, sigmas):
    """One frequency, one sky position: literal per-pulsar loop with 4 templates, 4 + 16 inner products."""
    N = np.zeros(4)
    M = np.zeros((4, 4))
    for Nvec, T, sigma, toa, resid, pos in zip(Nvecs, Ts, sigmas, toas, residuals, positions):
        fplus, fcross = antenna_pattern(pos, gwtheta, gwphi)
        s = 1 / fgw ** (1 / 3) * np.sin(2 * np.pi * fgw * toa)
        c = 1 / fgw ** (1 / 3) * np.cos(2 * np.pi * fgw * toa)
        A = np.stack((fplus * s, fplus * c, fcross * s, fcross * c))
        N += np.array([get_xCy(Nvec, T, sigma, resid, A[i]) for i in range(4)])
        M += np.array([[get_xCy(Nvec, T, sigma, A[i], A[j]) for j in range(4)] for i in range(4)])
    return 0.5 * np.dot(N, np.linalg.solve(M, N))
