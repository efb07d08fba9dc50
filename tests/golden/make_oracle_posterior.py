"""Long fp64 oracle runs (C oracle, Stan-semantics NUTS, recursive tree) on the three data lists:
election-day posterior summaries used (a) to validate the oracle against the reference's published
tables (README.md:83-136,179-232,279-332 -> readme_tables.json) and (b) as golden vectors for the
GPU sampler.  Output: oracle_posterior_{year}.json (per-state mean/sd/quantiles of
inv_logit(mu_b[,T]), MCSE, sampler diagnostics).   usage: make_oracle_posterior.py [years...]"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
import orc  # noqa: E402

years = [int(a) for a in sys.argv[1:]] or [2016, 2012, 2008]
for year in years:
    data = pkg.load_npz(os.path.join(HERE, f"datalist_{year}.npz"))
    om = orc.OracleModel(data)
    chains = 8
    r = om.sample(chains=chains, iter_warmup=500, iter_sampling=500, seed=1843, threads=8, tree_mode=0)
    mon = r["monitor"]  # [chains, 500, S+1] logit scale
    p = 1.0 / (1.0 + np.exp(-mon))
    # national row as the reports compute it (README.Rmd:230-248): state_weights-weighted mean of the state shares per draw
    p[:, :, -1] = p[:, :, :-1] @ np.asarray(data["state_weights"])
    flat = p.reshape(-1, p.shape[-1])
    ess = np.array([pkg.diagnostics.ess(p[:, :, k]) for k in range(p.shape[-1])])
    out = dict(
        year=year, chains=chains, iter_warmup=500, iter_sampling=500, seed=1843,
        states=[str(s) for s in data["_state_names"]] + ["\u2013"],  # README's label of the national row
        mean=flat.mean(0).tolist(), sd=flat.std(0, ddof=1).tolist(),
        q025=np.quantile(flat, 0.025, axis=0).tolist(), q975=np.quantile(flat, 0.975, axis=0).tolist(),
        q05=np.quantile(flat, 0.05, axis=0).tolist(), q95=np.quantile(flat, 0.95, axis=0).tolist(),
        prob=(flat > 0.5).mean(0).tolist(), ess=ess.tolist(), mcse=(flat.std(0, ddof=1) / np.sqrt(ess)).tolist(),
        stepsize=r["stepsize"].tolist(), n_leapfrog=r["n_leapfrog"].tolist(), seconds=r["seconds"],
        mean_treedepth_sampling=float(r["stats"][:, 500:, 3].mean()), mean_accept_sampling=float(r["stats"][:, 500:, 1].mean()),
        divergent_sampling=int(r["stats"][:, 500:, 5].sum()),
    )
    json.dump(out, open(os.path.join(HERE, f"oracle_posterior_{year}.json"), "w"), indent=0)
    tab = {row["state"]: row for row in json.load(open(os.path.join(HERE, "readme_tables.json")))[str(year)]}
    dm = max(abs(out["mean"][i] - tab[s]["mean"]) for i, s in enumerate(out["states"]) if s in tab)
    dl = max(abs(out["q025"][i] - tab[s]["low"]) for i, s in enumerate(out["states"]) if s in tab)
    dh = max(abs(out["q975"][i] - tab[s]["high"]) for i, s in enumerate(out["states"]) if s in tab)
    print(year, f"oracle {r['seconds']:.0f}s, {int(r['n_leapfrog'].sum())} leapfrogs; vs README: max|dmean| {dm:.4f} max|dlow| {dl:.4f} max|dhigh| {dh:.4f};"
          f" ESS min {ess.min():.0f}; eps {np.round(r['stepsize'], 4)}", flush=True)
