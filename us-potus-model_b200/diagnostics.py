"""Convergence diagnostics for the benchmark metric (ESS/sec).  The reference never looks at
n_eff / R-hat; these follow Stan's published definitions
(stan/analyze/mcmc/compute_effective_sample_size.hpp; Vehtari et al. 2021 for the rank-normalised
bulk ESS) so that the numbers mean what rstan::monitor's would."""
from __future__ import annotations

import numpy as np


def _autocov_fft(x: np.ndarray) -> np.ndarray:
    """Biased autocovariance of each row (Stan's autocovariance: divide by n)."""
    n = x.shape[-1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=-1, keepdims=True)
    f = np.fft.rfft(xc, m, axis=-1)
    ac = np.fft.irfft(f * np.conj(f), m, axis=-1)[..., :n]
    return ac / n


def ess(x: np.ndarray) -> float:
    """Stan's multi-chain effective sample size.  x: [chains, draws]."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    c, n = x.shape
    if n < 4 or not np.all(np.isfinite(x)):
        return float("nan")
    acov = _autocov_fft(x)
    chain_var = acov[:, 0] * n / (n - 1.0)
    mean_var = chain_var.mean()
    var_plus = mean_var * (n - 1.0) / n
    if c > 1:
        var_plus += x.mean(axis=1).var(ddof=1)
    if var_plus <= 0:
        return float("nan")
    rho_hat = np.zeros(n)
    t = 1
    rho_even = 1.0
    rho_hat[0] = rho_even
    rho_odd = 1 - (mean_var - acov[:, 1].mean()) / var_plus
    rho_hat[1] = rho_odd
    # Geyer's initial positive sequence
    while t < n - 4 and (rho_even + rho_odd) > 0:
        rho_even = 1 - (mean_var - acov[:, t + 1].mean()) / var_plus
        rho_odd = 1 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if rho_even + rho_odd >= 0:
            rho_hat[t + 1] = rho_even
            rho_hat[t + 2] = rho_odd
        t += 2
    max_t = t
    if rho_even > 0:
        rho_hat[max_t + 1] = rho_even
    # initial monotone sequence
    t = 1
    while t <= max_t - 3:
        if rho_hat[t + 1] + rho_hat[t + 2] > rho_hat[t - 1] + rho_hat[t]:
            rho_hat[t + 1] = (rho_hat[t - 1] + rho_hat[t]) / 2
            rho_hat[t + 2] = rho_hat[t + 1]
        t += 2
    tau = -1 + 2 * rho_hat[:max_t].sum() + rho_hat[max_t + 1]
    tau = max(tau, 1.0 / np.log10(c * n))
    return float(c * n / tau)


def _split(x: np.ndarray) -> np.ndarray:
    c, n = x.shape
    h = n // 2
    return np.concatenate([x[:, :h], x[:, n - h:]], axis=0)


def _rank_normalise(x: np.ndarray) -> np.ndarray:
    from scipy.stats import norm, rankdata
    r = rankdata(x.reshape(-1), method="average").reshape(x.shape)
    return norm.ppf((r - 0.375) / (x.size + 0.25))


def ess_bulk(x: np.ndarray) -> float:
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None, :]
    return ess(_rank_normalise(_split(x)))


def rhat(x: np.ndarray) -> float:
    """Split R-hat.  x: [chains, draws]."""
    x = _split(np.asarray(x, dtype=np.float64))
    c, n = x.shape
    w = x.var(axis=1, ddof=1).mean()
    b = n * x.mean(axis=1).var(ddof=1)
    return float(np.sqrt(((n - 1) / n * w + b / n) / w))


def summarize_monitor(monitor: np.ndarray) -> dict:
    """monitor: [chains, draws, K] -> min/median ESS (classic + bulk) and max R-hat over the K scalars."""
    K = monitor.shape[-1]
    e = np.array([ess(monitor[:, :, k]) for k in range(K)])
    eb = np.array([ess_bulk(monitor[:, :, k]) for k in range(K)])
    rh = np.array([rhat(monitor[:, :, k]) for k in range(K)])
    return dict(ess_min=float(np.nanmin(e)), ess_median=float(np.nanmedian(e)), ess_bulk_min=float(np.nanmin(eb)),
                ess_bulk_median=float(np.nanmedian(eb)), rhat_max=float(np.nanmax(rh)), ess=e, ess_bulk=eb, rhat=rh)
