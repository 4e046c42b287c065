"""ORACLE (test infrastructure): numpy restatement of the reference's convergence diagnostics.

  gelman_rubin   R/diagnostic.R:17-38
  ess            R/main.R:521-526 -> coda::effectiveSize (coda is an unpinned dependency of the
                 reference, DESCRIPTION:21): spectrum0.ar -> stats::ar (Yule-Walker, AIC order
                 selection, order.max = floor(10*log10(N))), ESS = N*var(x)/spec0.
  dic_gelman     R/main.R:548-554
Pinned by docs/articles/example.html:359 (Rhat 1.0019 / 1.0034) and :384 (ESS 1023.8251 / 942.6932).
Only tests/ and bench.py's checker legs import this file.
"""
from __future__ import annotations

import numpy as np


def gelman_rubin(draws: np.ndarray) -> float:
    """draws: [chains, samples] sampling-phase draws of one parameter."""
    chains, samples = draws.shape
    all_mean = draws.mean()
    chain_mean = draws.mean(axis=1)
    chain_var = draws.var(axis=1, ddof=1)
    W = chain_var.sum() / chains
    B = samples / (chains - 1) * ((chain_mean - all_mean) ** 2).sum()
    V = (1 - 1 / samples) * W + (1 / samples) * B
    return float(np.round(np.sqrt(V / W), 4))


def _ar_yw(x: np.ndarray):
    """stats::ar(x, aic=TRUE, method='yule-walker') -> (coefficients, var.pred, order)."""
    n = x.size
    order_max = min(n - 1, int(np.floor(10 * np.log10(n))))
    xc = x - x.mean()
    r = np.array([np.dot(xc[: n - k], xc[k:]) / n for k in range(order_max + 1)])
    # Levinson-Durbin
    vars_ = np.empty(order_max + 1)
    vars_[0] = r[0]
    coefs = [np.zeros(0)]
    phi = np.zeros(0)
    for k in range(1, order_max + 1):
        acc = r[k] - np.dot(phi, r[k - 1:0:-1]) if k > 1 else r[1]
        refl = acc / vars_[k - 1]
        phi = np.concatenate([phi - refl * phi[::-1], [refl]])
        vars_[k] = vars_[k - 1] * (1 - refl * refl)
        coefs.append(phi.copy())
    aic = n * np.log(vars_) + 2 * np.arange(order_max + 1) + 2
    order = int(np.argmin(aic))
    var_pred = vars_[order] * n / (n - (order + 1))
    return coefs[order], var_pred, order


def ess(x: np.ndarray) -> float:
    """coda::effectiveSize of one series."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    z = np.arange(1, n + 1, dtype=np.float64)
    A = np.stack([np.ones(n), z], 1)
    resid = x - A @ np.linalg.lstsq(A, x, rcond=None)[0]
    if np.isclose(resid.std(ddof=1), 0.0, atol=1.5e-8):
        return 0.0
    ar, var_pred, _ = _ar_yw(x)
    spec0 = var_pred / (1 - ar.sum()) ** 2
    return float(n * x.var(ddof=1) / spec0)


def dic_gelman(loglike_cold_sampling: np.ndarray) -> float:
    dev = -2.0 * np.asarray(loglike_cold_sampling, dtype=np.float64).ravel()
    return float(dev.mean() + 0.5 * dev.var(ddof=1))
