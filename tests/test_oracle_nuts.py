"""Invariants of the oracle's Stan-semantics sampler (the sampler itself is third-party Stan 2.24.1 and
is not in the reference tree, so it is pinned by behaviour)."""
import numpy as np
import pytest

from conftest import small_datalist
import potus_oracle as po


@pytest.fixture(scope="module")
def small(orc_mod):
    d = small_datalist(S=5, T=9, Ns=40, Nn=12)
    return d, orc_mod.OracleModel(d)


def test_recursive_and_iterative_trees_agree(small):
    """Same momentum / direction streams => identical tree shape (depth, n_leapfrog, divergence) for Stan's
    recursion and the iterative builder the CUDA kernel uses; only the multinomial pick may differ."""
    d, om = small
    q0 = np.random.default_rng(0).normal(0, 0.3, om.D)
    for eps in (0.01, 0.03):
        qa, sa = om.transitions(q0, eps, np.ones(om.D), n_iter=1, tree_mode=0)
        qb, sb = om.transitions(q0, eps, np.ones(om.D), n_iter=1, tree_mode=1)
        assert np.array_equal(sa[:, 3:6], sb[:, 3:6])          # treedepth, n_leapfrog, divergent
        assert np.allclose(sa[:, 1], sb[:, 1], rtol=1e-12)      # accept_stat: same energies along the trajectory
    # several transitions in a row stay consistent in distribution: compare mean depth loosely
    _, sa = om.transitions(q0, 0.02, np.ones(om.D), n_iter=40, tree_mode=0)
    _, sb = om.transitions(q0, 0.02, np.ones(om.D), n_iter=40, tree_mode=1)
    assert abs(sa[:, 3].mean() - sb[:, 3].mean()) < 1.0


def test_trajectory_length_and_accept_bounds(small):
    d, om = small
    q0 = np.zeros(om.D)
    _, st = om.transitions(q0, 0.02, np.ones(om.D), n_iter=30, tree_mode=1)
    depth, nleap = st[:, 3], st[:, 4]
    assert np.all(nleap >= 1) and np.all(nleap <= 2 ** 10 - 1)
    assert np.all(nleap <= 2 ** np.maximum(depth + 1, 1) - 1)   # at most one aborted extra doubling
    assert np.all((st[:, 1] >= 0) & (st[:, 1] <= 1))
    assert np.all(st[:, 5] == 0)


def test_max_treedepth_is_respected(small):
    d, om = small
    _, st = om.transitions(np.zeros(om.D), 1e-4, np.ones(om.D), n_iter=2, tree_mode=1, max_depth=4)
    assert np.all(st[:, 3] == 4) and np.all(st[:, 4] == 15)


def test_huge_step_diverges(small):
    d, om = small
    _, st = om.transitions(np.zeros(om.D), 50.0, np.ones(om.D), n_iter=3, tree_mode=1)
    assert np.all(st[:, 5] == 1)


def test_window_schedule_and_dual_averaging(small):
    """Stan's windows for 500 warm-up iterations end at 99/149/249/449: the step size is re-initialised there
    (init_stepsize doubles/halves from the adapted value, mu = log(10 eps)), so the next iteration's step size
    is far from its predecessor's exactly at those iterations and frozen after warm-up."""
    d, om = small
    r = om.sample(chains=2, iter_warmup=500, iter_sampling=60, threads=2, tree_mode=1)
    eps = r["stats"][:, :, 2]
    assert np.all(eps[:, 500:] == eps[:, 500:501])                        # frozen after warm-up
    assert np.allclose(eps[:, 500], r["stepsize"])
    acc = r["stats"][:, 500:, 1].mean()
    assert 0.6 < acc < 0.98                                               # adapt_delta = 0.8 target, roughly
    assert r["stats"][:, 500:, 5].sum() == 0


def test_short_warmup_uses_15_75_10_rule(small):
    d, om = small
    r = om.sample(chains=1, iter_warmup=100, iter_sampling=5, threads=1, tree_mode=0)
    assert np.isfinite(r["stats"][0, :, 0]).all()


def test_sampler_recovers_posterior_moments_small(small):
    """Detailed balance, statistically: two independent seeds / both tree builders agree on posterior means
    of the monitored quantities within Monte-Carlo error."""
    d, om = small
    a = om.sample(chains=4, iter_warmup=300, iter_sampling=400, threads=4, tree_mode=0, seed=11)
    b = om.sample(chains=4, iter_warmup=300, iter_sampling=400, threads=4, tree_mode=1, seed=12)
    ma, mb = a["monitor"].reshape(-1, om.S + 1), b["monitor"].reshape(-1, om.S + 1)
    se = np.sqrt(ma.var(0) / 400 + mb.var(0) / 400)                       # ESS >= ~400 of 1600 assumed
    assert np.all(np.abs(ma.mean(0) - mb.mean(0)) < 5 * se)


def test_chain_rng_is_keyed_by_global_chain_id(small):
    d, om = small
    a = om.sample(chains=2, iter_warmup=20, iter_sampling=5, threads=2, tree_mode=1)
    b = om.sample(chains=1, iter_warmup=20, iter_sampling=5, threads=1, tree_mode=1, chain_id_offset=1)
    assert np.array_equal(a["monitor"][1], b["monitor"][0])


def test_config0_2008_backtest_plumbing(orc_mod, datalists):
    """BASELINE.json configs[0]: the repo's own 2008 backtest at 2 chains x 200 iterations (100 warm-up => Stan's
    15/75/10 % window rule), run through the CPU oracle because rstan is not installable here.  Plumbing check:
    data list -> sampler -> election-day table lands on the published one (README.md:83-136) within the Monte-Carlo
    error of 200 draws."""
    import json, os
    from conftest import GOLDEN
    d = datalists[2008]
    om = orc_mod.OracleModel(d)
    r = om.sample(chains=2, iter_warmup=100, iter_sampling=100, seed=1843, threads=2, literal=True, tree_mode=0)
    assert r["stats"][:, 100:, 5].sum() == 0
    p = 1 / (1 + np.exp(-r["monitor"].reshape(-1, 52)))[:, :51]
    tab = {row["state"]: row for row in json.load(open(os.path.join(GOLDEN, "readme_tables.json")))["2008"]}
    names = [str(s) for s in d["_state_names"]]
    dm = max(abs(p[:, i].mean() - tab[s]["mean"]) for i, s in enumerate(names))
    assert dm < 0.012, dm
