"""Convergence diagnostics of the drjacoby_output object (host side, numpy).

Mirrors the reference's R post-processing, which is also host code (a "next" row):
  gelman_rubin   R/diagnostic.R:17-38   rounded sqrt(V/W) over equal-length chains
  ess            R/main.R:521-526       coda::effectiveSize = N var(x) / spec0, spec0 from an AR(p)
                                        fit (Yule-Walker, AIC order selection, p <= 10 log10 N)
  dic_gelman     R/main.R:548-554       mean(D) + 0.5 var(D), D = -2 loglike (cold rung, sampling)
"""
from __future__ import annotations

import numpy as np


def gelman_rubin(draws, chains: int | None = None, samples: int | None = None) -> float:
    """draws: array [chains, samples] of one parameter (sampling phase)."""
    x = np.asarray(draws, dtype=np.float64)
    if x.ndim != 2:
        raise ValueError("draws must be [chains, samples]")
    m, n = x.shape
    if (chains is not None and chains != m) or (samples is not None and samples != n):
        raise ValueError("chains/samples do not match the shape of draws")
    if m <= 1:
        raise ValueError("chains must be greater than 1")
    if n <= 1:
        raise ValueError("samples must be greater than 1")
    grand = x.mean()
    means = x.mean(axis=1)
    W = x.var(axis=1, ddof=1).sum() / m
    Bv = n / (m - 1) * np.square(means - grand).sum()
    V = (1.0 - 1.0 / n) * W + Bv / n
    return float(np.round(np.sqrt(V / W), 4))


def _levinson(r: np.ndarray):
    """Levinson-Durbin on autocovariances r[0..p]: (list of AR coefficient vectors, innovations variances)."""
    p = r.size - 1
    v = np.empty(p + 1)
    v[0] = r[0]
    coefs = [np.zeros(0)]
    a = np.zeros(0)
    for k in range(1, p + 1):
        num = r[k] - (a @ r[k - 1:0:-1] if k > 1 else 0.0)
        kappa = num / v[k - 1]
        a = np.append(a - kappa * a[::-1], kappa)
        v[k] = v[k - 1] * (1.0 - kappa * kappa)
        coefs.append(a)
    return coefs, v


def spectrum0_ar(x) -> tuple[float, int]:
    """coda::spectrum0.ar for one series: (spectral density at zero, selected AR order)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    n = x.size
    t = np.arange(1, n + 1, dtype=np.float64)
    tc = t - t.mean()
    slope = (tc @ (x - x.mean())) / (tc @ tc)
    resid = (x - x.mean()) - slope * tc
    if np.isclose(resid.std(ddof=1), 0.0, rtol=0.0, atol=1.5e-8):
        return 0.0, 0
    pmax = min(n - 1, int(np.floor(10 * np.log10(n))))
    xc = x - x.mean()
    # autocovariances with divisor n via FFT-free direct products (pmax is at most ~50)
    r = np.array([xc[: n - k] @ xc[k:] for k in range(pmax + 1)]) / n
    coefs, v = _levinson(r)
    aic = n * np.log(v) + 2.0 * np.arange(pmax + 1) + 2.0
    order = int(np.argmin(aic))
    var_pred = v[order] * n / (n - (order + 1))
    return float(var_pred / (1.0 - coefs[order].sum()) ** 2), order


def ess(x) -> float:
    """coda::effectiveSize of one series."""
    x = np.asarray(x, dtype=np.float64).ravel()
    spec, _ = spectrum0_ar(x)
    return 0.0 if spec == 0.0 else float(x.size * x.var(ddof=1) / spec)


def gelman_rubin_from_moments(chain_mean, chain_var, samples: int) -> float:
    """gelman_rubin() from per-chain means and sample variances (equal chain lengths)."""
    m = np.asarray(chain_mean, dtype=np.float64)
    v = np.asarray(chain_var, dtype=np.float64)
    chains = m.size
    if chains <= 1:
        raise ValueError("chains must be greater than 1")
    W = v.sum() / chains
    Bv = samples / (chains - 1) * np.square(m - m.mean()).sum()
    V = (1.0 - 1.0 / samples) * W + Bv / samples
    return float(np.round(np.sqrt(V / W), 4))


def ess_from_autocov(autocov_sum, n: int) -> float:
    """coda::effectiveSize from the autocovariance SUMS sum_j (x_j - mean)(x_{j+l} - mean), l = 0..order.max,
    of a series of length n (the sums come from the device, drj_session_summary)."""
    r = np.asarray(autocov_sum, dtype=np.float64) / n
    if not r[0] > 0.0:
        return 0.0
    pmax = r.size - 1
    coefs, v = _levinson(r)
    aic = n * np.log(v) + 2.0 * np.arange(pmax + 1) + 2.0
    order = int(np.argmin(aic))
    var_pred = v[order] * n / (n - (order + 1))
    spec0 = var_pred / (1.0 - coefs[order].sum()) ** 2
    var1 = r[0] * n / (n - 1)
    return float(n * var1 / spec0)


def dic_gelman(loglike) -> float:
    dev = -2.0 * np.asarray(loglike, dtype=np.float64).ravel()
    return float(dev.mean() + 0.5 * dev.var(ddof=1))
