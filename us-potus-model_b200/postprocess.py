"""Host-side restatement of the reports' post-processing -- the CHECKER of the on-device
version (csrc/potus_post.cu, `PotusFit.summary()` / `potus_postprocess`), which is what the product path uses.

Restates (numpy, on the election-day slice the sampler's `monitor` buffer holds for EVERY draw):
  README.Rmd:206-220   per-state mean / 2.5% / 97.5% / P(win) of predicted_score[, T, s]
  README.Rmd:230-248   national vote per draw = state_weights-weighted mean of the state shares
  README.Rmd:271-300   electoral-college simulation: dem_ev = sum(ev * (share > 0.5)) per draw
  README.Rmd:378-390   Brier scores (EV-weighted, unweighted) and states called correctly
The actual-winner lists are the ones hard-coded at README.Rmd:381 (2008), :892 (2012), :1416 (2016).
"""
from __future__ import annotations

import numpy as np

DEM_WINNERS = {
    2008: ('CA', 'NV', 'OR', 'WA', 'CO', 'NM', 'MN', 'IL', 'VA', 'DC', 'MD', 'DE', 'NJ', 'CT', 'RI', 'MA', 'NH', 'VT', 'NY', 'HI', 'ME',
           'MI', 'IA', 'OH', 'PA', 'WI', 'FL', 'NC', 'IN'),
    2012: ('CA', 'NV', 'OR', 'WA', 'CO', 'NM', 'MN', 'IL', 'VA', 'DC', 'MD', 'DE', 'NJ', 'CT', 'RI', 'MA', 'NH', 'VT', 'NY', 'HI', 'ME',
           'MI', 'IA', 'OH', 'PA', 'WI', 'FL'),
    2016: ('CA', 'NV', 'OR', 'WA', 'CO', 'NM', 'MN', 'IL', 'VA', 'DC', 'MD', 'DE', 'NJ', 'CT', 'RI', 'MA', 'NH', 'VT', 'NY', 'HI', 'ME'),
}
# published by the reference (README.md:75,169,260): ev_wtd_brier, unwtd_brier, states_correct
PUBLISHED_BRIER = {2008: (0.0321964, 0.0289902, 49), 2012: (0.0324297, 0.0193188, 50), 2016: (0.0725679, 0.0508319, 48)}


def election_day_shares(monitor: np.ndarray) -> np.ndarray:
    """monitor [chains, draws, S+1] (logit scale, last column national_mu_b_average) -> shares [n, S]."""
    m = np.asarray(monitor)
    return 1.0 / (1.0 + np.exp(-m.reshape(-1, m.shape[-1])[:, :-1]))


def state_table(shares: np.ndarray, states) -> dict:
    return dict(state=[str(s) for s in states], mean=shares.mean(0), low=np.quantile(shares, 0.025, axis=0),
                high=np.quantile(shares, 0.975, axis=0), prob=(shares > 0.5).mean(0))


def national_vote(shares: np.ndarray, state_weights) -> dict:
    nat = shares @ np.asarray(state_weights)
    return dict(mean=float(nat.mean()), low=float(np.quantile(nat, 0.025)), high=float(np.quantile(nat, 0.975)),
                prob=float((nat > 0.5).mean()), draws=nat)


def electoral_college(shares: np.ndarray, ev) -> dict:
    dem_ev = (shares > 0.5) @ np.asarray(ev)
    return dict(mean=float(dem_ev.mean()), median=float(np.median(dem_ev)), low=float(np.quantile(dem_ev, 0.025)),
                high=float(np.quantile(dem_ev, 0.975)), prob=float((dem_ev >= 270).mean()), draws=dem_ev)


def brier_scores(prob: np.ndarray, states, ev, year: int) -> dict:
    actual = np.array([1.0 if str(s) in DEM_WINNERS[year] else 0.0 for s in states])
    diff = (actual - np.asarray(prob)) ** 2
    w = np.asarray(ev, float) / np.sum(ev)
    return dict(ev_wtd_brier=float(np.sum(diff * w)), unwtd_brier=float(diff.mean()),
                states_correct=int(np.sum(np.round(prob) == actual)))
