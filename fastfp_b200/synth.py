"""Seeded synthetic pulsar-timing-array inputs (no ``enterprise`` needed).

This is the recipe of SURVEY.md §8(d): it produces exactly the objects the reference's
hot path consumes -- per-pulsar ``toas``/``residuals`` (what ``FastFp.__init__`` copies,
reference ``fastfp/fastfp.py:44-45``), and the lists ``Nvecs``, ``Ts``, ``sigmas``/``TNTs``
that ``get_mats_fp`` / ``get_mats_nmfp`` return (reference ``fastfp/utils.py:72-76, 97-99``).

Everything is float64 NumPy on the host; it is *input construction*, not part of the
measured path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Dict, List, Optional

import numpy as np

from . import constants as const

SEED0 = 20240607
MJD0_SECONDS = 53000.0 * 86400.0
SPAN_YEARS = 15.0


def powerlaw_phi(Ffreqs: np.ndarray, log10_A: float, gamma: float) -> np.ndarray:
    """Power-law prior variances, same expression and operation order as the reference's
    ``RN_container._powerlaw`` (``fastfp/nmfp.py:226-234``). Used here only to *draw* the
    injected red noise and to build the fixed-noise ``sigmas`` of the plain-Fp configs."""
    df = np.diff(np.concatenate((np.array([0.0]), Ffreqs[::2])))
    return (
        Ffreqs ** (-gamma)
        * (10.0**log10_A) ** 2
        / 12.0
        / np.pi**2
        * const.fyr ** (gamma - 3)
        * np.repeat(df, 2)
    )


@dataclass
class SynthPTA:
    """A synthetic array of pulsars plus everything the Fp / nmfp entry points take."""

    psrs: list  # duck-typed pulsars: .name .toas .residuals .Mmat .backend_flags
    Nvecs: List[np.ndarray]
    Ts: List[np.ndarray]
    TNTs: List[np.ndarray]
    phis: List[np.ndarray]
    sigmas: List[np.ndarray]
    noise: Dict[str, float]  # "<psr>_red_noise_log10_A", "<psr>_red_noise_gamma", "gw_*"
    Tspan: float
    Ffreqs: Optional[np.ndarray]  # (2*ncomps,) repeat(k/Tspan, 2) or None for white-only
    n_tm: List[int] = field(default_factory=list)
    ncomps: int = 30
    inc_cp: bool = True

    @property
    def P(self) -> int:
        return len(self.psrs)

    # the four accessors of an ``enterprise`` PTA that ``get_mats_fp`` / ``get_mats_nmfp`` call
    # (reference ``fastfp/utils.py:72-75, 97-99``)
    def get_phiinv(self, noise=None):
        return [1.0 / phi for phi in self.phis]

    def get_TNT(self, noise=None):
        return self.TNTs

    def get_ndiag(self, noise=None):
        return self.Nvecs

    def get_basis(self, noise=None):
        return self.Ts

    @property
    def toas(self):
        return [q.toas for q in self.psrs]

    @property
    def residuals(self):
        return [q.residuals for q in self.psrs]


def _timing_basis(t: np.ndarray, n_tm: int) -> np.ndarray:
    """Left singular vectors of a toy timing-model design matrix (mirrors
    ``TimingModel(use_svd=True)``, reference ``fastfp/utils.py:146``)."""
    tm = 0.5 * (t[0] + t[-1])
    span = max(t[-1] - t[0], 1.0)
    tau = (t - tm) / span
    w1 = 2.0 * np.pi * (t - tm) / const.yr
    cols = [np.ones_like(t), tau, tau**2]
    k = 1
    while len(cols) < n_tm:
        cols.append(np.sin(k * w1))
        cols.append(np.cos(k * w1))
        cols.append(tau * np.sin(k * w1))
        cols.append(tau * np.cos(k * w1))
        cols.append(tau ** (k + 2))
        k += 1
    M = np.stack(cols[:n_tm], axis=1)
    M = M / np.linalg.norm(M, axis=0)
    U, _, _ = np.linalg.svd(M, full_matrices=False)
    return np.ascontiguousarray(U)


def _sky_position(rng) -> np.ndarray:
    """isotropic unit vector (``enterprise.pulsar.Pulsar.pos``)"""
    z = rng.uniform(-1.0, 1.0)
    ph = rng.uniform(0.0, 2.0 * np.pi)
    r = np.sqrt(1.0 - z * z)
    return np.array([r * np.cos(ph), r * np.sin(ph), z])


def make_pta(
    P: int,
    n,
    n_tm=12,
    ncomps: int = 30,
    white_only: bool = False,
    inc_cp: bool = True,
    seed: int = SEED0,
    epoch: int = 0,
    nbackends: int = 2,
) -> SynthPTA:
    """Build a synthetic PTA.

    ``n`` and ``n_tm`` may be ints or per-pulsar sequences (ragged arrays are the norm in
    the reference: its pulsar axis is a Python loop, ``fastfp/fastfp.py:72-74``).
    ``white_only=True`` gives config C1: ``T`` holds only the timing-model columns.
    ``epoch > 1`` clusters the TOAs into observing epochs of ``epoch`` TOAs 0.2 s apart (what ECORR models),
    epochs alternating between ``nbackends`` receiver backends (``psr.backend_flags``); see :func:`with_ecorr`.
    """
    ns = [int(n)] * P if np.isscalar(n) else [int(v) for v in n]
    ntms = [int(n_tm)] * P if np.isscalar(n_tm) else [int(v) for v in n_tm]
    assert len(ns) == P and len(ntms) == P

    rngs = [np.random.default_rng(seed + p) for p in range(P)]
    if epoch > 1:
        toas, flags = [], []
        for r, k in zip(rngs, ns):
            nep = -(-k // epoch)
            t0 = MJD0_SECONDS + np.sort(r.uniform(0.0, SPAN_YEARS * const.yr, size=nep))
            t0 = t0 + 10.0 * np.arange(nep)  # keep epochs at least 10 s apart
            toas.append((t0[:, None] + 0.2 * np.arange(epoch)[None, :]).reshape(-1)[:k])
            flags.append(np.repeat(np.array([f"be{e % nbackends}" for e in range(nep)]), epoch)[:k])
    else:
        toas = [MJD0_SECONDS + np.sort(r.uniform(0.0, SPAN_YEARS * const.yr, size=k)) for r, k in zip(rngs, ns)]
        flags = [np.array(["synth"] * k) for k in ns]
    Tspan = float(max(t.max() for t in toas) - min(t.min() for t in toas))
    Ffreqs = None if white_only else np.repeat(np.arange(1, ncomps + 1) / Tspan, 2)

    noise: Dict[str, float] = {"gw_log10_A": float(np.log10(2e-15)), "gw_gamma": 13.0 / 3.0}
    psrs, Nvecs, Ts, TNTs, phis, sigmas = [], [], [], [], [], []
    for p in range(P):
        rng, t, ntm = rngs[p], toas[p], ntms[p]
        name = f"J{p:04d}+0000"
        sig = rng.uniform(1e-7, 1e-6, size=t.size)
        Nvec = sig**2
        U = _timing_basis(t, ntm)
        log10_A = float(rng.uniform(-15.0, -13.0))
        gamma = float(rng.uniform(1.0, 6.0))
        noise[f"{name}_red_noise_log10_A"] = log10_A
        noise[f"{name}_red_noise_gamma"] = gamma
        white = sig * rng.standard_normal(t.size)
        if white_only:
            T = U
            phi = np.ones(ntm) * 1e40
            r = white
        else:
            F = np.empty((t.size, 2 * ncomps))
            arg = 2.0 * np.pi * t[:, None] * Ffreqs[None, ::2]
            F[:, ::2] = np.sin(arg)
            F[:, 1::2] = np.cos(arg)
            T = np.concatenate((U, F), axis=1)
            phi_rn = powerlaw_phi(Ffreqs, log10_A, gamma)
            if inc_cp:
                phi_rn = phi_rn + powerlaw_phi(Ffreqs, noise["gw_log10_A"], noise["gw_gamma"])
            phi = np.concatenate((np.ones(ntm) * 1e40, phi_rn))
            a = rng.standard_normal(2 * ncomps) * np.sqrt(phi_rn)
            r = white + F @ a
        r = r - U @ (U.T @ r)  # timing-model fit removed
        T = np.ascontiguousarray(T)
        TNT = T.T @ (T / Nvec[:, None])
        TNT = 0.5 * (TNT + TNT.T)
        psrs.append(
            SimpleNamespace(
                name=name,
                toas=t,
                residuals=r,
                Mmat=U,
                backend_flags=flags[p],
                pos=_sky_position(rng),
            )
        )
        Nvecs.append(Nvec)
        Ts.append(T)
        TNTs.append(TNT)
        phis.append(phi)
        sigmas.append(TNT + np.diag(1.0 / phi))  # reference fastfp/utils.py:76
    return SynthPTA(psrs, Nvecs, Ts, TNTs, phis, sigmas, noise, Tspan, Ffreqs, ntms, ncomps, inc_cp and not white_only)


def with_ecorr(pta: SynthPTA, kernel: bool = False, seed: int = SEED0 + 555):
    """ECORR on top of a PTA made with ``epoch > 1``: per (pulsar, backend) a ``log10_ecorr`` in ``[-7, -6]`` added
    to ``pta.noise`` under the key :class:`fastfp_b200.GPEcorr_container` looks up (``fastfp/nmfp.py:447-450``).
    Returns ``(Nvecs, Ts, TNTs, phis)`` for

    * ``kernel=False`` -- the GP form the reference implements: ``T = [tm | U | Fourier]`` with the epoch-indicator
      columns ``U`` (backend by backend), ``phi = [1e40 | 10^(2 log10_ecorr) | red noise]``, diagonal ``N``;
    * ``kernel=True`` -- the block-diagonal ``N`` (:class:`fastfp_b200.BlockNvec`) with the unchanged basis."""
    from . import model

    rng = np.random.default_rng(seed)
    Nvecs, Ts, TNTs, phis = [], [], [], []
    for p, q in enumerate(pta.psrs):
        for val in np.unique(q.backend_flags):
            pta.noise[f"{q.name}_basis_ecorr_{val}_log10_ecorr"] = float(rng.uniform(-7.0, -6.0))
        ntm = pta.n_tm[p]
        if kernel:
            B = model.kernel_ecorr_blocks(q, pta.Nvecs[p], pta.noise)
            TNT = pta.Ts[p].T @ B.solve(pta.Ts[p])
            Nvecs.append(B); Ts.append(pta.Ts[p]); phis.append(pta.phis[p])
        else:
            U = model.ecorr_basis_by_backend(q)
            w = model.ecorr_weights_by_backend(q)
            jv = np.concatenate([wi * 10.0 ** (2.0 * pta.noise[f"{q.name}_basis_ecorr_{val}_log10_ecorr"])
                                 for wi, val in zip(w, np.unique(q.backend_flags))]) if U.shape[1] else np.zeros(0)
            T = np.ascontiguousarray(np.concatenate((pta.Ts[p][:, :ntm], U, pta.Ts[p][:, ntm:]), axis=1))
            TNT = T.T @ (T / pta.Nvecs[p][:, None])
            Nvecs.append(pta.Nvecs[p]); Ts.append(T)
            phis.append(np.concatenate((pta.phis[p][:ntm], jv, pta.phis[p][ntm:])))
        TNTs.append(0.5 * (TNT + TNT.T))
    return Nvecs, Ts, TNTs, phis


def fp_freqs(F: int) -> np.ndarray:
    """The plain-Fp frequency grid of ``examples/run_fp.py:59``."""
    return np.linspace(2e-9, 3e-7, F)


def nmfp_freqs(F: int, Tspan: float) -> np.ndarray:
    """The nmfp frequency grid of ``examples/run_nmfp.py:247``."""
    return np.arange(1, F + 1) / Tspan


def draw_samples(pta: SynthPTA, D: int, seed: int = SEED0 + 100003) -> Dict[str, np.ndarray]:
    """``D`` stand-in MCMC draws as the dict-of-``(D,)``-arrays that ``map_params``
    builds (reference ``examples/run_nmfp.py:174-186, 256-261``)."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for psr in pta.psrs:
        out[f"{psr.name}_red_noise_log10_A"] = rng.uniform(-15.0, -13.0, size=D)
        out[f"{psr.name}_red_noise_gamma"] = rng.uniform(1.0, 6.0, size=D)
    out["gw_log10_A"] = rng.uniform(-15.0, -14.0, size=D)
    out["gw_gamma"] = rng.uniform(3.0, 5.5, size=D)
    return out


# The BASELINE.json configs made concrete.
CONFIGS = {
    "C1": dict(P=10, n=1000, n_tm=12, white_only=True, F=1, D=1),
    "C2": dict(P=45, n=5000, n_tm=12, white_only=False, F=10_000, D=1),
    "C3": dict(P=45, n=5000, n_tm=12, white_only=False, F=1_000, D=1_000),
    "C4": dict(P=68, n=10_000, n_tm=12, white_only=False, F=1_000_000, D=1),
    "C5": dict(P=68, n=10_000, n_tm=12, white_only=False, F=10_000, D=10_000),
}


def make_config(name: str, **over) -> SynthPTA:
    c = dict(CONFIGS[name])
    c.update(over)
    return make_pta(c["P"], c["n"], c["n_tm"], white_only=c["white_only"])
