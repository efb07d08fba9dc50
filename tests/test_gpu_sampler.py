"""GPU sampler tests (-m gpu): decision-level agreement with the oracle's iterative NUTS, Stan-semantics
invariants, sharding invariance, output contract, and statistical parity with the reference's published
tables (README.md:279-332 etc. -> tests/golden/readme_tables.json)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_first_transitions_follow_the_oracle(pkg, orc_mod, datalists, cuda_lib):
    """Same Philox streams, same algorithm: until fp32-vs-fp64 round-off is amplified by the chaotic dynamics,
    the device and the fp64 oracle (tree_mode=1) take identical decisions.  Checked on the first 12 warm-up
    iterations (depths up to 10): tree depth, n_leapfrog and divergence equal; lp within 0.1 on the first 5 iterations (far-out inits, |lp| ~ 1.2e6); step size within 1% and accept_stat within 0.02 on the first 8."""
    d = datalists[2016]
    fit = pkg.cmdstan_model().sample(data=d, seed=1843, chains=2, iter_warmup=12, iter_sampling=0, keep_per_chain=0)
    sp = fit.sampler_params()
    om = orc_mod.OracleModel(d)
    r = om.sample(chains=2, iter_warmup=12, iter_sampling=0, seed=1843, threads=2, tree_mode=1)
    for c in range(2):
        assert np.array_equal(sp["treedepth__"][c], r["stats"][c, :, 3]), (sp["treedepth__"][c], r["stats"][c, :, 3])
        assert np.array_equal(sp["n_leapfrog__"][c], r["stats"][c, :, 4])
        assert np.array_equal(sp["divergent__"][c], r["stats"][c, :, 5])
        dlp = np.abs(sp["lp__"][c] - r["stats"][c, :, 0])
        assert dlp[:5].max() < 0.1, dlp   # later iterations: a flipped multinomial pick changes lp by O(10) this far from stationarity
        assert np.abs(sp["stepsize__"][c] / r["stats"][c, :, 2] - 1)[:8].max() < 0.01
        assert np.abs(sp["accept_stat__"][c] - r["stats"][c, :, 1])[:8].max() < 0.02


def test_sharding_does_not_change_any_chain(pkg, datalists, cuda_lib):
    """Chains are keyed by global id: 6 chains in one sampler == 2 samplers of 3 with chain_id_offset 0 and 3
    (bit-identical draws).  This is the multi-GPU partition property, checked on one device."""
    d = datalists[2008]
    m = pkg.cmdstan_model()
    kw = dict(data=d, seed=7, iter_warmup=25, iter_sampling=6, keep_per_chain=2)
    full = m.sample(chains=6, **kw)
    a = m.sample(chains=3, chain_id_offset=0, **kw)
    b = m.sample(chains=3, chain_id_offset=3, **kw)
    th = full.theta().reshape(6, 2, -1)
    assert np.array_equal(th[:3], a.theta().reshape(3, 2, -1)) and np.array_equal(th[3:], b.theta().reshape(3, 2, -1))
    assert np.array_equal(full.monitor()[3:], b.monitor())


def test_output_contract(pkg, orc_mod, datalists, cuda_lib):
    """Shapes/layout of what rstan::extract consumers index (final_2016.R:556,568,597,622,647,682,708) and
    consistency of the transformed parameters with the oracle's constrain() at the same theta."""
    d = datalists[2016]
    fit = pkg.cmdstan_model("poll_model_2020.stan").sample(data=d, seed=3, chains=4, iter_warmup=30, iter_sampling=8, keep_per_chain=4)
    ex = fit.extract(["mu_b", "mu_c", "mu_m", "mu_pop", "polling_bias", "e_bias", "predicted_score"])
    n = 16
    assert ex["mu_b"].shape == (n, 51, 254) and ex["predicted_score"].shape == (n, 254, 51)
    assert ex["mu_c"].shape == (n, 161) and ex["mu_m"].shape == (n, 3) and ex["mu_pop"].shape == (n, 3)
    assert ex["polling_bias"].shape == (n, 51) and ex["e_bias"].shape == (n, 254)
    assert np.allclose(ex["predicted_score"], 1 / (1 + np.exp(-np.transpose(ex["mu_b"], (0, 2, 1)))), atol=1e-6)
    th = fit.theta()
    om = orc_mod.OracleModel(d)
    for k in (0, 7, 15):
        c = om.constrain(th[k])
        assert np.abs(c["mu_b"] - ex["mu_b"][k]).max() < 2e-5
        assert np.abs(c["mu_c"] - ex["mu_c"][k]).max() < 1e-6 and np.abs(c["e_bias"] - ex["e_bias"][k]).max() < 1e-6
        assert np.abs(c["polling_bias"] - ex["polling_bias"][k]).max() < 1e-5
        assert np.abs(c["mu_m"] - ex["mu_m"][k]).max() < 1e-6 and np.abs(c["mu_pop"] - ex["mu_pop"][k]).max() < 1e-6
    mon = fit.monitor()   # every sampling iteration; kept draws are the thinned subset (iterations 0, thin, 2 thin, ... as CmdStan's `thin`)
    assert mon.shape == (4, 8, 52)
    kept_T = ex["mu_b"][:, :, 253].reshape(4, 4, 51)
    assert np.allclose(mon[:, 0::2, :51], kept_T, atol=1e-6)
    sp = fit.sampler_params()
    assert set(sp) == set(pkg.model.SAMPLER_PARAMS) and sp["lp__"].shape == (4, 38)
    lp_o = np.array([om.logp_grad(th[k])[0] for k in range(n)])
    lp_kept = sp["lp__"][:, 30:][:, 0::2].reshape(-1)
    assert np.abs(lp_kept - lp_o).max() < 0.05
    assert fit.model_name == "poll_model_2020"
    with pytest.raises(KeyError):
        fit.extract("mu_a")
    # f3: CmdStan CSV files (what rstan::read_stan_csv parses, final_2016.R:543) hold the same draws
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        paths = fit.save_csvfiles(td, chains=[0, 3])
        back = pkg.stancsv.read_stan_csv(paths)
    assert back["names"] == pkg.stancsv.column_names(d) and len(back["names"]) == 43360
    for j, c in enumerate((0, 3)):
        for k in range(4):
            r = c * 4 + k
            # files carry 6 significant figures (CmdStan's default sig_figs)
            assert np.allclose(back["draws"]["mu_b"][k, j], ex["mu_b"][r], rtol=1e-5, atol=1e-6)
            assert np.allclose(back["draws"]["predicted_score"][k, j], ex["predicted_score"][r], rtol=1e-5, atol=1e-6)
            assert np.allclose(back["draws"]["raw_mu_b_T"][k, j], th[r, :51], rtol=1e-5, atol=1e-6)
            assert abs(back["sampler_params"]["lp__"][k, j] - (sp["lp__"][c, 30 + 2 * k] + np.log(0.02))) < 6.0   # 6 significant figures of -1.17e6
    im = fit.inv_metric()
    assert im.shape == (4, 15098) and (im > 0).all() and np.isfinite(im).all()
    assert np.allclose(back["inv_metric"][1], im[3], rtol=1e-4)
    assert abs(back["stepsize"][0] - sp["stepsize__"][0, -1]) < 1e-6


def test_no_mode_variant_runs_and_hides_full_only_pars(pkg, datalists, cuda_lib):
    fit = pkg.cmdstan_model("poll_model_2020_no_mode_adjustment.stan").sample(data=datalists[2012], chains=3, iter_warmup=20,
                                                                              iter_sampling=4, keep_per_chain=1)
    assert fit.extract("mu_b").shape == (3, 51, 251)
    with pytest.raises(KeyError):
        fit.extract("e_bias")
    assert fit.model_name.endswith("no_mode_adjustment")


def test_stan_adaptation_invariants(pkg, datalists, cuda_lib):
    """Step size frozen after warm-up; windows end at 99/149/249/449 for 500 warm-up iterations is checked on the
    short-warm-up rule instead (100 -> 15/75/10): metric update at iteration 89 resets the step-size search.
    accept_stat targets adapt_delta; no divergences post warm-up; depth <= max_treedepth; n_leapfrog <= 2^depth+1."""
    d = datalists[2016]
    fit = pkg.cmdstan_model().sample(data=d, seed=5, chains=16, iter_warmup=100, iter_sampling=30, keep_per_chain=1, max_treedepth=9)
    sp = fit.sampler_params()
    eps = sp["stepsize__"]
    assert np.all(eps[:, 100:] == eps[:, 100:101])
    assert np.all(sp["treedepth__"] <= 9) and np.all(sp["n_leapfrog__"] <= 2 ** 9 - 1 + 2 ** 9)
    assert np.all((sp["accept_stat__"] >= 0) & (sp["accept_stat__"] <= 1))
    acc = sp["accept_stat__"][:, 100:].mean()
    assert 0.6 < acc < 0.995
    assert sp["divergent__"][:, 100:].sum() == 0
    assert np.all(np.isfinite(sp["energy__"])) and np.all(np.isfinite(sp["lp__"]))


def test_posterior_matches_reference_tables_2016(pkg, datalists, cuda_lib):
    """End-to-end statistical parity with the reference's published election-day table (README.md:279-332:
    per-state mean / 2.5% / 97.5% of inv_logit(mu_b[,T]), 3 d.p., from 6x500 rstan draws).  Tolerances are the
    reference's own run-to-run spread (README vs model_reports/v4_cov_error_rewrite.html):
    |dmean| <= 0.003, |d interval end| <= 0.012.  Also against the long fp64 oracle run within 4 MCSE + 5e-4."""
    d = datalists[2016]
    fit = pkg.cmdstan_model().sample(data=d, seed=1843, chains=296, iter_warmup=500, iter_sampling=200, keep_per_chain=1)
    st = fit.stats
    assert st["n_divergent_sampling"] == 0
    p = 1 / (1 + np.exp(-fit.monitor().reshape(-1, 52)))
    p[:, 51] = p[:, :51] @ d["state_weights"]   # national row as README.Rmd:230-248 computes it
    names = [str(s) for s in d["_state_names"]] + ["\u2013"]  # README labels the national row with an en dash
    tab = {r["state"]: r for r in json.load(open(os.path.join(GOLDEN, "readme_tables.json")))["2016"]}
    mean, lo, hi = p.mean(0), np.quantile(p, 0.025, axis=0), np.quantile(p, 0.975, axis=0)
    dm = max(abs(mean[i] - tab[s]["mean"]) for i, s in enumerate(names))
    dl = max(abs(lo[i] - tab[s]["low"]) for i, s in enumerate(names))
    dh = max(abs(hi[i] - tab[s]["high"]) for i, s in enumerate(names))
    print(f"2016 vs README: max|dmean| {dm:.4f} |dlow| {dl:.4f} |dhigh| {dh:.4f}; eps {st['mean_stepsize']:.4f} depth {st['mean_treedepth']:.2f}")
    assert dm <= 0.003 and dl <= 0.012 and dh <= 0.012
    ora = json.load(open(os.path.join(GOLDEN, "oracle_posterior_2016.json")))
    z = np.abs(mean - np.array(ora["mean"])) / (4 * np.array(ora["mcse"]) + 5e-4)
    assert z.max() <= 1.0, z.max()
    # the north_star's 90% intervals (5% / 95%; the README prints 2.5 / 97.5 only): device vs the long fp64 oracle run.
    # Tolerance: Monte-Carlo error of a tail quantile ~ 2.1 x MCSE of the mean (density at the 5% point of a normal),
    # so 4 x 2.1 x mcse + 1e-3 for the oracle's own 4000-draw quantile error
    q05, q95 = np.quantile(p, 0.05, axis=0), np.quantile(p, 0.95, axis=0)
    tolq = 8.4 * np.array(ora["mcse"]) + 1e-3
    dq = max(np.abs(q05 - np.array(ora["q05"])).max(), np.abs(q95 - np.array(ora["q95"])).max())
    print(f"2016 90% interval ends vs oracle: max|d| {dq:.4f}")
    assert np.all(np.abs(q05 - np.array(ora["q05"])) <= tolq) and np.all(np.abs(q95 - np.array(ora["q95"])) <= tolq)
    ess = pkg.diagnostics.ess(p[:, 9].reshape(296, 200))
    assert ess > 0.2 * p.shape[0]
    # the reports' headline numbers (README.md:260): Brier scores and states called, from OUR draws
    sh = pkg.postprocess.election_day_shares(fit.monitor())
    tab_ = pkg.postprocess.state_table(sh, d["_state_names"])
    b = pkg.postprocess.brier_scores(tab_["prob"], d["_state_names"], d["_ev_state"], 2016)
    pub = pkg.postprocess.PUBLISHED_BRIER[2016]
    print("2016 Brier", b, "published", pub)
    assert abs(b["ev_wtd_brier"] - pub[0]) < 0.006 and abs(b["unwtd_brier"] - pub[1]) < 0.004 and abs(b["states_correct"] - pub[2]) <= 1


def test_posterior_matches_reference_tables_2008_no_mode(pkg, datalists, cuda_lib):
    """Config 1 of BASELINE.json (the 2008 backtest, no-mode model) vs README.md:83-136."""
    d = datalists[2008]
    fit = pkg.cmdstan_model("poll_model_2020_no_mode_adjustment.stan").sample(data=d, seed=1843, chains=148, iter_warmup=500,
                                                                              iter_sampling=200, keep_per_chain=1)
    p = 1 / (1 + np.exp(-fit.monitor().reshape(-1, 52)))
    p[:, 51] = p[:, :51] @ d["state_weights"]   # national row as README.Rmd:230-248 computes it
    names = [str(s) for s in d["_state_names"]] + ["\u2013"]  # README labels the national row with an en dash
    tab = {r["state"]: r for r in json.load(open(os.path.join(GOLDEN, "readme_tables.json")))["2008"]}
    mean, lo, hi = p.mean(0), np.quantile(p, 0.025, axis=0), np.quantile(p, 0.975, axis=0)
    dm = max(abs(mean[i] - tab[s]["mean"]) for i, s in enumerate(names))
    dl = max(abs(lo[i] - tab[s]["low"]) for i, s in enumerate(names))
    dh = max(abs(hi[i] - tab[s]["high"]) for i, s in enumerate(names))
    print(f"2008 vs README: max|dmean| {dm:.4f} |dlow| {dl:.4f} |dhigh| {dh:.4f}")
    assert dm <= 0.003 and dl <= 0.012 and dh <= 0.012


def test_posterior_matches_reference_tables_2012_no_mode(pkg, datalists, cuda_lib):
    """The 2012 backtest (no-mode model, final_2012.R) vs README.md:179-232, and vs the long fp64 oracle run."""
    d = datalists[2012]
    fit = pkg.cmdstan_model("poll_model_2020_no_mode_adjustment.stan").sample(data=d, seed=1843, chains=148, iter_warmup=500,
                                                                              iter_sampling=200, keep_per_chain=1)
    p = 1 / (1 + np.exp(-fit.monitor().reshape(-1, 52)))
    p[:, 51] = p[:, :51] @ d["state_weights"]
    names = [str(s) for s in d["_state_names"]] + ["\u2013"]
    tab = {r["state"]: r for r in json.load(open(os.path.join(GOLDEN, "readme_tables.json")))["2012"]}
    mean, lo, hi = p.mean(0), np.quantile(p, 0.025, axis=0), np.quantile(p, 0.975, axis=0)
    dm = max(abs(mean[i] - tab[s]["mean"]) for i, s in enumerate(names))
    dl = max(abs(lo[i] - tab[s]["low"]) for i, s in enumerate(names))
    dh = max(abs(hi[i] - tab[s]["high"]) for i, s in enumerate(names))
    print(f"2012 vs README: max|dmean| {dm:.4f} |dlow| {dl:.4f} |dhigh| {dh:.4f}")
    assert dm <= 0.003 and dl <= 0.012 and dh <= 0.012
    ora = json.load(open(os.path.join(GOLDEN, "oracle_posterior_2012.json")))
    z = np.abs(mean - np.array(ora["mean"])) / (4 * np.array(ora["mcse"]) + 5e-4)
    assert z.max() <= 1.0, z.max()
    q05, q95 = np.quantile(p, 0.05, axis=0), np.quantile(p, 0.95, axis=0)   # 90% interval, as in the 2016 test
    tolq = 8.4 * np.array(ora["mcse"]) + 1.5e-3                                # (148 x 200 device draws here)
    assert np.all(np.abs(q05 - np.array(ora["q05"])) <= tolq) and np.all(np.abs(q95 - np.array(ora["q95"])) <= tolq)


def test_same_seed_same_binary_is_bit_reproducible(pkg, datalists, cuda_lib):
    """Two runs of one binary with one seed must agree bit for bit (no atomics, fixed reduction trees).  Added after an
    unchanged build was seen to give two different n_leapfrog totals (profiles/r01_d_ab4.log)."""
    d = datalists[2016]
    runs = []
    for _ in range(3):
        fit = pkg.cmdstan_model("poll_model_2020.stan").sample(data=d, seed=1843, chains=148, iter_warmup=40, iter_sampling=10,
                                                               keep_per_chain=1)
        sp = fit.sampler_params()
        runs.append({k: v.copy() for k, v in sp.items()})
        fit.close()
    for r in runs[1:]:
        for k in runs[0]:
            assert np.array_equal(runs[0][k], r[k]), f"{k} differs between identical runs"


@pytest.mark.parametrize("year", [2012, 2008])
def test_no_mode_draw_record_offsets(pkg, orc_mod, datalists, cuda_lib, year, tmp_path):
    """The no-mode data lists still carry M and Pop (final_2012.R:501-540, final_2008.R:504-545) although the model has no
    mu_m / mu_pop blocks.  The draw record must be read where the kernel wrote it: polling_bias, mu_c and theta of the
    2012/2008 fits against the oracle's constrain() at the same theta (polling_bias is extracted at final_2012.R:623,
    final_2008.R:627), and the CmdStan CSV of the no-mode variant round-trips."""
    d = datalists[year]
    assert int(d["M"]) >= 1 and int(d["Pop"]) >= 1 and "poll_mode_state" not in d
    fit = pkg.cmdstan_model("poll_model_2020_no_mode_adjustment.stan").sample(data=d, seed=11, chains=3, iter_warmup=25,
                                                                              iter_sampling=4, keep_per_chain=2)
    ex = fit.extract(["mu_b", "mu_c", "polling_bias"])
    th = fit.theta()
    om = orc_mod.OracleModel(d)
    assert th.shape == (6, om.D) and np.all(np.isfinite(th)) and np.abs(th).max() < 50
    for k in range(6):
        c = om.constrain(th[k])
        assert np.abs(c["mu_b"] - ex["mu_b"][k]).max() < 2e-5
        assert np.abs(c["mu_c"] - ex["mu_c"][k]).max() < 1e-6
        assert np.abs(c["polling_bias"] - ex["polling_bias"][k]).max() < 1e-5
    sp = fit.sampler_params()
    lp_o = np.array([om.logp_grad(th[k])[0] for k in range(6)])
    assert np.abs(sp["lp__"][:, 25:][:, 0::2].reshape(-1) - lp_o).max() < 0.05
    paths = fit.save_csvfiles(str(tmp_path))
    back = pkg.stancsv.read_stan_csv(paths)
    for c in range(3):
        for k in range(2):
            assert np.allclose(back["draws"]["polling_bias"][k, c], ex["polling_bias"][2 * c + k], rtol=1e-5, atol=1e-6)
            assert np.allclose(back["draws"]["raw_polling_bias"][k, c], th[2 * c + k, -51:], rtol=1e-5, atol=1e-6)


def test_no_metric_adaptation_below_20_warmup_iterations_on_device(pkg, orc_mod, datalists, cuda_lib):
    """iter_warmup < 20: Stan performs no variance adaptation; the inverse metric stays 1 and sampling runs at exp(x_bar)
    of one uninterrupted dual-averaging run.  (The kernel used to fire a window end on the last warm-up iteration with zero
    samples: metric 1e-3, step size 1.)  Decision-level agreement with the oracle through warm-up AND the first sampling
    iterations pins the hand-over."""
    d = datalists[2016]
    fit = pkg.cmdstan_model().sample(data=d, seed=1843, chains=2, iter_warmup=12, iter_sampling=3, keep_per_chain=1)
    assert np.all(fit.inv_metric() == 1.0)
    sp = fit.sampler_params()
    r = orc_mod.OracleModel(d).sample(chains=2, iter_warmup=12, iter_sampling=3, seed=1843, threads=2, tree_mode=1)
    for c in range(2):
        assert np.all(sp["stepsize__"][c, 12:] == sp["stepsize__"][c, 12])
        assert abs(sp["stepsize__"][c, 12] / r["stepsize"][c] - 1) < 0.05   # exp(x_bar) of 12 noisy fp32-vs-fp64 accept_stats
        assert np.array_equal(sp["treedepth__"][c, :12], r["stats"][c, :12, 3])


def test_adaptation_matches_oracle_distribution(pkg, datalists, cuda_lib):
    """SURVEY 8(a) row a12 against the oracle, past the first window: 296 device chains x 500 warm-up iterations vs the
    committed 64-chain fp64 oracle run (tests/golden/oracle_adaptation_2016.npz, make_oracle_adaptation.py; independent
    chain ids, same law).  (i) per-coordinate mean over chains of log(inverse metric): |device - oracle| within 5 standard
    errors of the difference (+1.5% slack), for every one of the 15098 coordinates; (ii) the final step size: means within
    4 SE, spread within a factor 1.5; (iii) the dual-averaging recursion replayed from the recorded accept_stat__ reproduces
    every recorded step size when -- and only when -- it is restarted exactly after the window ends 99/149/249/449 (Stan
    defaults 75/25/50), the restart value being a power-of-two multiple (init_stepsize); (iv) frozen after warm-up; (v) the
    mean step size per iteration tracks the oracle's through the whole warm-up."""
    d = datalists[2016]
    ora = np.load(os.path.join(GOLDEN, "oracle_adaptation_2016.npz"))
    C = 296
    fit = pkg.cmdstan_model().sample(data=d, seed=1843, chains=C, iter_warmup=500, iter_sampling=4, keep_per_chain=1)
    im = fit.inv_metric()
    lm, ls = np.log(im).mean(0), np.log(im).std(0, ddof=1)
    n_o = int(ora["chains"])
    se = np.sqrt(ls ** 2 / C + ora["log_inv_metric_sd"].astype(np.float64) ** 2 / n_o)
    z = np.abs(lm - ora["log_inv_metric_mean"]) / (5 * se + 0.015)    # 15098 coordinates: the largest of that many |N(0,1)| is ~4.1
    print(f"a12: log inv-metric max z {z.max():.3f} (coordinate {int(z.argmax())}); median ratio {np.exp(np.median(lm - ora['log_inv_metric_mean'])):.4f}")
    assert z.max() <= 1.0, (z.max(), int(z.argmax()))
    assert abs(np.median(lm - ora["log_inv_metric_mean"])) < 0.01
    sp = fit.sampler_params()
    eps_f = sp["stepsize__"][:, 500]
    eo = ora["stepsize"]
    se_e = np.sqrt(eps_f.var(ddof=1) / C + eo.var(ddof=1) / n_o)
    print(f"a12: final eps device {eps_f.mean():.5f} +- {eps_f.std(ddof=1):.5f}, oracle {eo.mean():.5f} +- {eo.std(ddof=1):.5f}")
    assert abs(eps_f.mean() - eo.mean()) <= 4 * se_e + 1e-4
    assert 1 / 1.5 < eps_f.std(ddof=1) / eo.std(ddof=1) < 1.5
    # (iii) replay Stan's dual averaging from the RECORDED accept_stat__ with restarts exactly after iterations 99/149/249/449:
    # every recorded step size must be reproduced (1e-3), the value after a window end must be the replayed value times an
    # integer power of two (init_stepsize doubles / halves), and the frozen sampling step size must be exp(x_bar) of the last run
    eps = sp["stepsize__"].astype(np.float64)
    acc = np.minimum(1.0, sp["accept_stat__"].astype(np.float64))
    mu = np.log(10 * eps[:, 0]); cnt = 0; sbar = np.zeros(C); xbar = np.zeros(C)
    worst, worst_k = 0.0, 0.0
    for it in range(500):
        cnt += 1
        eta = 1.0 / (cnt + 10)
        sbar = (1 - eta) * sbar + eta * (0.8 - acc[:, it])
        x = mu - sbar * np.sqrt(cnt) / 0.05
        xe = cnt ** -0.75
        xbar = (1 - xe) * xbar + xe * x
        if it in (99, 149, 249, 449):
            k = np.log2(eps[:, it + 1] / np.exp(x))
            worst_k = max(worst_k, np.abs(k - np.rint(k)).max())
            mu = np.log(10 * eps[:, it + 1]); cnt = 0; sbar = np.zeros(C); xbar = np.zeros(C)
        elif it < 499:
            worst = max(worst, np.abs(eps[:, it + 1] / np.exp(x) - 1).max())
    print(f"a12: dual-averaging replay: worst relative step-size mismatch {worst:.2e}; worst distance of log2(restart ratio) from an integer {worst_k:.2e}")
    assert worst < 1e-3 and worst_k < 2e-3
    assert np.abs(eps[:, 500] / np.exp(xbar) - 1).max() < 1e-3
    eps = sp["stepsize__"][:, :500]
    assert np.all(sp["stepsize__"][:, 500:] == sp["stepsize__"][:, 500:501])
    by_iter = np.abs(np.log(eps.mean(0)[100:] / ora["stepsize_by_iter"][100:500]))   # the whole adaptation path, not just its end:
    print(f"a12: mean step size by iteration vs oracle: median |log ratio| {np.median(by_iter):.3f}, max {by_iter.max():.3f}")
    assert np.median(by_iter) < 0.06 and by_iter.max() < 0.3     # per-iteration eps has ~30% spread; oracle mean is over 64 chains


def test_in_library_multi_gpu_is_the_same_chains(pkg, datalists, cuda_lib):
    """PotusConfig.n_gpus: ONE process, chains sharded over the devices, one ncclAllGather of the kept draws
    inside potus_run.  Every chain must be bit-identical to the single-device run (RNG streams are keyed by global chain id)
    and every output must come back in global chain order.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on one."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = datalists[2008]
    m = pkg.cmdstan_model()
    kw = dict(data=d, seed=7, iter_warmup=25, iter_sampling=6, keep_per_chain=2)
    one = m.sample(chains=7, **kw)                 # 7 chains over 2 devices: shards of 4 and 3 (unequal: the gather pads)
    two = m.sample(chains=7, n_gpus=2, **kw)
    assert two.stats["n_draws_kept"] == 14 and two.stats["gpu_launches"] == 2 * one.stats["gpu_launches"] and two.stats["seconds_gather"] > 0
    assert np.array_equal(one.theta(), two.theta())
    assert np.array_equal(one.extract("mu_b"), two.extract("mu_b")) and np.array_equal(one.extract("polling_bias"), two.extract("polling_bias"))
    assert np.array_equal(one.monitor(), two.monitor())
    a, b = one.sampler_params(), two.sampler_params()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(one.inv_metric(), two.inv_metric())
    assert one.stats["n_leapfrog_total"] == two.stats["n_leapfrog_total"]


@pytest.mark.parametrize("force_stream", [False, True])
def test_sampling_phase_transitions_follow_the_oracle(pkg, orc_mod, datalists, cuda_lib, force_stream):
    """The STATIONARY phase, decision by decision: chains seeded (potus_set_state) with the committed adapted oracle states
    (tests/golden/oracle_adapted_states_2016.npz: position after 500 warm-up iterations, step size ~0.014, adapted metric) run
    sampling transitions on the device and in the fp64 oracle with the same Philox streams.  These are the depth-8 trajectories
    that end through an in-subtree U-turn -- the exit path almost every production transition takes.  First transition: tree
    depth, n_leapfrog, divergence identical and accept_stat within 0.01 for every chain; over 3 transitions at most one chain
    may have left the oracle's path (fp32 round-off amplified over ~750 leapfrogs)."""
    d = datalists[2016]
    st = np.load(os.path.join(GOLDEN, "oracle_adapted_states_2016.npz"))
    C, n_it = 8, 3
    state = dict(theta=st["q"][:C].astype(np.float64), stepsize=st["stepsize"][:C], inv_metric=st["inv_metric"][:C].astype(np.float64))
    fit = pkg.cmdstan_model().sample(data=d, seed=99, chains=C, iter_warmup=0, iter_sampling=n_it, keep_per_chain=n_it, state=state,
                                     force_stream=force_stream)
    sp = fit.sampler_params()
    om = orc_mod.OracleModel(d)
    off_path = 0
    for c in range(C):
        q_o, s_o = om.transitions(state["theta"][c], state["stepsize"][c], state["inv_metric"][c], n_iter=n_it, seed=99, chain=c, tree_mode=1, iter0=0)
        assert sp["treedepth__"][c, 0] == s_o[0, 3] and sp["n_leapfrog__"][c, 0] == s_o[0, 4] and sp["divergent__"][c, 0] == s_o[0, 5] == 0
        assert abs(sp["accept_stat__"][c, 0] - s_o[0, 1]) < 0.01 and abs(sp["lp__"][c, 0] - s_o[0, 0]) < 0.05
        assert np.all(sp["stepsize__"][c] == np.float32(state["stepsize"][c]))
        off_path += not np.array_equal(sp["n_leapfrog__"][c], s_o[:, 4])
        assert np.all(sp["treedepth__"][c] >= 7)      # stationary trajectories: ~pi/eps leapfrogs
    assert off_path <= 1, off_path
    assert np.all(sp["n_leapfrog__"] < 2 ** sp["treedepth__"] + 0.5)   # in-subtree exits happen: not every tree is complete
    th = fit.theta().reshape(C, n_it, -1)
    assert np.abs(th[:, 0] - np.stack([om.transitions(state["theta"][c], state["stepsize"][c], state["inv_metric"][c], n_iter=1, seed=99, chain=c,
                                                      tree_mode=1, iter0=0)[0][0] for c in range(C)])).max() < 5e-3


def test_on_device_postprocessing_matches_the_host_restatement(pkg, datalists, cuda_lib):
    """potus_postprocess (csrc/potus_post.cu: SURVEY 8(f) row f2 on the device, over every sampling iteration) against the host
    numpy restatement of the reports' computations (postprocess.py: README.Rmd:206-300) and diagnostics.py (Stan ESS / split
    R-hat) on the same monitor buffer: means/sd 1e-6, quantiles 2e-6 (exact order statistics of fp32 shares), P(win) and the
    electoral-college numbers exact up to draws whose share is within 1e-6 of 0.5, ESS 1e-6 relative, R-hat 1e-9; and the
    device-formed predicted_score equals inv_logit(mu_b)' of the kept draws."""
    d = datalists[2016]
    fit = pkg.cmdstan_model().sample(data=d, seed=3, chains=24, iter_warmup=150, iter_sampling=60, keep_per_chain=3)
    ev = d["_ev_state"].astype(float)
    sm = fit.summary(ev=ev, ev_threshold=270.0)
    mon = fit.monitor()
    sh = pkg.postprocess.election_day_shares(mon)
    tab = pkg.postprocess.state_table(sh, d["_state_names"])
    assert np.abs(sm["states"]["mean"] - tab["mean"]).max() < 1e-6 and np.abs(sm["states"]["sd"] - sh.std(0, ddof=1)).max() < 1e-6
    assert np.abs(sm["states"]["q025"] - tab["low"]).max() < 2e-6 and np.abs(sm["states"]["q975"] - tab["high"]).max() < 2e-6
    for q, nm in ((0.05, "q05"), (0.5, "q50"), (0.95, "q95")):
        assert np.abs(sm["states"][nm] - np.quantile(sh, q, axis=0)).max() < 2e-6
    edge = (np.abs(sh - 0.5) < 1e-6).mean(0)
    assert np.all(np.abs(sm["states"]["prob"] - tab["prob"]) <= edge + 1e-12)
    nat = pkg.postprocess.national_vote(sh, d["state_weights"])
    assert abs(sm["national"]["mean"] - nat["mean"]) < 1e-6 and abs(sm["national"]["q025"] - nat["low"]) < 2e-6 and abs(sm["national"]["q975"] - nat["high"]) < 2e-6
    assert abs(sm["national"]["prob"] - nat["prob"]) <= (np.abs(nat["draws"] - 0.5) < 1e-6).mean() + 1e-12
    ec = pkg.postprocess.electoral_college(sh, ev)
    n_edge = (np.abs(sh - 0.5) < 1e-6).any(1).mean()
    assert abs(sm["electoral_votes"]["mean"] - ec["mean"]) <= 60 * n_edge + 1e-6 and abs(sm["electoral_votes"]["prob"] - ec["prob"]) <= n_edge + 1e-12
    assert abs(sm["electoral_votes"]["q50"] - ec["median"]) <= (1 if n_edge > 0 else 1e-9)
    e_host = np.array([pkg.diagnostics.ess(mon[:, :, k]) for k in range(52)])
    r_host = np.array([pkg.diagnostics.rhat(mon[:, :, k]) for k in range(52)])
    assert np.abs(sm["ess"] / e_host - 1).max() < 1e-6, np.abs(sm["ess"] / e_host - 1).max()
    assert np.abs(sm["rhat"] - r_host).max() < 1e-9
    assert np.abs(sm["monitor_mean"] - mon.reshape(-1, 52).mean(0)).max() < 1e-6
    ps, mu = fit.extract("predicted_score"), fit.extract("mu_b")
    assert ps.shape == (72, 254, 51) and np.abs(ps - 1 / (1 + np.exp(-np.transpose(mu, (0, 2, 1))))).max() < 1e-6
    st, sp = fit.stats, fit.sampler_params()
    assert st["n_leapfrog_total"] == int(sp["n_leapfrog__"].sum()) and st["n_leapfrog_sampling"] == int(sp["n_leapfrog__"][:, 150:].sum())
    assert abs(st["mean_accept_stat"] - sp["accept_stat__"][:, 150:].mean()) < 1e-6 and abs(st["mean_treedepth"] - sp["treedepth__"][:, 150:].mean()) < 1e-9
