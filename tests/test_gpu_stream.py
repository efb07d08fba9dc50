"""GPU parity tests of the STREAMING kernel family (potus_stream.cu; BASELINE config 5 and every shape the resident
kernel does not hold).  Everything goes through the C-ABI; the checker is the fp64 oracle.  `force_stream=True`
(POTUS_FLAG_FORCE_STREAM) runs shapes that would fit the resident kernel through the streaming one, so the two kernel
families are also compared with each other on the reference's own 2016 list."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, small_datalist

pytestmark = pytest.mark.gpu


def _shape(pkg, name):
    if name == "toy":
        return small_datalist(S=5, T=9, Ns=40, Nn=12)
    if name == "toy_nomode_nonat":
        return small_datalist(S=7, T=5, Ns=9, Nn=0, full=False)
    if name == "T2":
        return small_datalist(S=4, T=2, Ns=6, Nn=3)
    if name == "one_state":
        return small_datalist(S=1, T=2, Ns=2, Nn=1, P=1)
    if name == "S51_T254":
        return small_datalist(S=51, T=254, Ns=300, Nn=50, P=40)
    if name == "T129_tile_edge":           # day 129 is the first row of the second tile
        return small_datalist(S=9, T=129, Ns=200, Nn=40, P=7)
    if name == "T128_exact":
        return small_datalist(S=17, T=128, Ns=150, Nn=30, P=5)
    if name == "S64_T300":
        return pkg.synthetic_datalist(S=64, T=300, N_state=3000, N_national=800, P=40)
    if name == "S128_T365":
        return pkg.synthetic_datalist(S=128, T=365, N_state=8000, N_national=2000, P=128)
    if name == "S200_T512_max_days":
        return pkg.synthetic_datalist(S=200, T=512, N_state=6000, N_national=1500, P=300)
    if name == "config5":
        return pkg.synthetic_datalist()       # S=256, T=365, N=40000+10000, P=512, seed 1843
    raise KeyError(name)


@pytest.mark.parametrize("name", ["toy", "toy_nomode_nonat", "T2", "one_state", "S51_T254", "T129_tile_edge", "T128_exact", "S64_T300",
                                  "S128_T365", "S200_T512_max_days", "config5"])
def test_stream_logp_grad_matches_oracle(pkg, orc_mod, cuda_lib, name):
    """fp32 state, fp16x2-split tcgen05 GEMMs with fp32 accumulate, fp64 energy reduction.
    Tolerance: |lp - lp_oracle| <= 1e-8 |lp| (floor 1e3), gradient max error <= 2e-6 max|grad|
    (3e-6 on the toy shapes, like the resident kernel's edge-shape test: with S = 1 a single 22-bit operand pair decides)."""
    d = _shape(pkg, name)
    om = orc_mod.OracleModel(d)
    rng = np.random.default_rng(1)
    th = np.stack([0.5 * rng.standard_normal(om.D), rng.uniform(-2, 2, om.D), np.zeros(om.D)])
    lp, g = pkg.logp_grad(d, th, force_stream=True)
    for i in range(len(th)):
        lpo, go = om.logp_grad(th[i])
        # (point 1 = U(-2,2), Stan's init range: |eta - eta_hat| > 12 for many polls, the far-tail branch; 3e-8 there)
        assert abs(lp[i] - lpo) <= (3e-8 if (i == 1 or name == "config5") else 1e-8) * max(abs(lpo), 1e3), (name, i, lp[i], lpo)
        tol = 3e-6 if int(d["S"]) < 51 else 2e-6
        assert np.abs(g[i] - go).max() <= tol * max(np.abs(go).max(), 1.0), (name, i, np.abs(g[i] - go).max(), np.abs(go).max())


def test_stream_and_resident_kernels_agree_on_the_2016_list(pkg, datalists, cuda_lib):
    """The two kernel families evaluate the same model: same lp (1e-8 relative; each is within 1e-8 of the oracle) and gradient (1e-6 max|g|) on the 2016 list,
    and the 2016 known-answer values hold through the streaming kernel too."""
    d = datalists[2016]
    D = 15098
    rng = np.random.default_rng(2)
    th = np.stack([np.zeros(D), 0.1 * np.sin(1 + 0.37 * np.arange(D)), rng.uniform(-2, 2, D)])
    lp_r, g_r = pkg.logp_grad(d, th)
    lp_s, g_s = pkg.logp_grad(d, th, force_stream=True)
    assert np.all(np.abs(lp_r - lp_s) <= 1e-8 * np.abs(lp_r))
    for i in range(3):
        assert np.abs(g_r[i] - g_s[i]).max() <= 1e-6 * np.abs(g_r[i]).max()
    kat = json.load(open(os.path.join(GOLDEN, "known_answers.json")))["2016"]
    assert abs(lp_s[0] - kat["lp_zero"]) < 1e-8 * abs(lp_s[0]) and abs(lp_s[1] - kat["lp_sin"]) < 1e-8 * abs(lp_s[1])


@pytest.mark.parametrize("name,iters", [("2016", 12), ("S64_T300", 10), ("config5", 3)])
def test_stream_first_transitions_follow_the_oracle(pkg, orc_mod, datalists, cuda_lib, name, iters):
    """Decision-level parity of the streaming NUTS with the fp64 oracle (tree_mode=1, same Philox streams): tree depth,
    n_leapfrog and divergence equal on ALL the iterations run; step size within 1% / accept_stat within 0.02 until fp32-vs-fp64
    round-off has been amplified by the dynamics (first 8 iterations on the 2016 list, first 6 on the synthetic shapes)."""
    d = datalists[2016] if name == "2016" else _shape(pkg, name)
    fit = pkg.cmdstan_model().sample(data=d, seed=1843, chains=2, iter_warmup=iters, iter_sampling=0, keep_per_chain=0, force_stream=True)
    sp = fit.sampler_params()
    r = orc_mod.OracleModel(d).sample(chains=2, iter_warmup=iters, iter_sampling=0, seed=1843, threads=2, tree_mode=1)
    # equality is required while the two arithmetics still follow the same path: all 12 iterations on the 2016 list, the first
    # 6 on the synthetic shapes (their later iterations sit on 1023-leaf trajectories from far-out inits, where one ulp in a
    # poll's linear predictor legitimately ends in a different tree a few iterations on -- seen when a re-association of that
    # sum moved the first difference of S=64 x T=300 from after iteration 10 to before it)
    k_eq = iters if name == "2016" else min(6, iters)
    for c in range(2):
        assert np.array_equal(sp["treedepth__"][c, :k_eq], r["stats"][c, :k_eq, 3]), (sp["treedepth__"][c], r["stats"][c, :, 3])
        assert np.array_equal(sp["n_leapfrog__"][c, :k_eq], r["stats"][c, :k_eq, 4])
        assert np.array_equal(sp["divergent__"][c, :k_eq], r["stats"][c, :k_eq, 5])
        k = min(8 if name == "2016" else 6, iters)
        assert np.abs(sp["stepsize__"][c] / r["stats"][c, :, 2] - 1)[:k].max() < 0.01
        assert np.abs(sp["accept_stat__"][c] - r["stats"][c, :, 1])[:k].max() < 0.02


def test_stream_output_contract_and_sharding(pkg, orc_mod, cuda_lib):
    """rstan::extract shapes on a shape only the streaming kernel holds (S=64, T=300), transformed parameters against the
    oracle's constrain(), monitor == kept draws, bit-identical chains under sharding and across repeated runs."""
    d = _shape(pkg, "S64_T300")
    m = pkg.cmdstan_model()
    kw = dict(data=d, seed=5, iter_warmup=20, iter_sampling=4, keep_per_chain=2)
    fit = m.sample(chains=4, **kw)
    ex = fit.extract(["mu_b", "mu_c", "mu_m", "mu_pop", "polling_bias", "e_bias", "predicted_score"])
    assert ex["mu_b"].shape == (8, 64, 300) and ex["predicted_score"].shape == (8, 300, 64) and ex["e_bias"].shape == (8, 300)
    th = fit.theta()
    om = orc_mod.OracleModel(d)
    assert th.shape == (8, om.D)
    for k in (0, 3, 7):
        c = om.constrain(th[k])
        assert np.abs(c["mu_b"] - ex["mu_b"][k]).max() < 2e-5
        assert np.abs(c["mu_c"] - ex["mu_c"][k]).max() < 1e-6 and np.abs(c["e_bias"] - ex["e_bias"][k]).max() < 1e-6
        assert np.abs(c["polling_bias"] - ex["polling_bias"][k]).max() < 1e-5
        assert np.abs(c["mu_m"] - ex["mu_m"][k]).max() < 1e-6 and np.abs(c["mu_pop"] - ex["mu_pop"][k]).max() < 1e-6
    mon = fit.monitor()
    assert mon.shape == (4, 4, 65)
    assert np.allclose(mon[:, 0::2, :64], ex["mu_b"][:, :, 299].reshape(4, 2, 64), atol=1e-6)
    nat = np.einsum("dk,k->d", ex["mu_b"][:, :, 299], d["state_weights"])
    assert np.allclose(mon[:, 0::2, 64].reshape(-1), nat, atol=2e-5)
    sp = fit.sampler_params()
    lp_o = np.array([om.logp_grad(th[k])[0] for k in range(8)])
    assert np.abs(sp["lp__"][:, 20:][:, 0::2].reshape(-1) - lp_o).max() < 0.1
    im = fit.inv_metric()
    assert im.shape == (4, om.D) and (im > 0).all()
    a = m.sample(chains=2, chain_id_offset=0, **kw)
    b = m.sample(chains=2, chain_id_offset=2, **kw)
    t4 = th.reshape(4, 2, -1)
    assert np.array_equal(t4[:2], a.theta().reshape(2, 2, -1)) and np.array_equal(t4[2:], b.theta().reshape(2, 2, -1))
    again = m.sample(chains=4, **kw)
    assert np.array_equal(again.theta(), th) and np.array_equal(again.sampler_params()["n_leapfrog__"], sp["n_leapfrog__"])


def test_stream_posterior_matches_reference_tables_2016(pkg, datalists, cuda_lib):
    """End-to-end statistical parity of the STREAMING kernel family on the reference's own data: the 2016 list forced
    through potus_stream_kernel vs README.md:279-332 (|dmean| <= 0.003, interval ends <= 0.012) and vs the long fp64
    oracle run (4 MCSE + 5e-4)."""
    d = datalists[2016]
    fit = pkg.cmdstan_model().sample(data=d, seed=1843, chains=148, iter_warmup=500, iter_sampling=200, keep_per_chain=1, force_stream=True)
    st = fit.stats
    assert st["n_divergent_sampling"] == 0
    p = 1 / (1 + np.exp(-fit.monitor().reshape(-1, 52)))
    p[:, 51] = p[:, :51] @ d["state_weights"]
    names = [str(s) for s in d["_state_names"]] + ["–"]
    tab = {r["state"]: r for r in json.load(open(os.path.join(GOLDEN, "readme_tables.json")))["2016"]}
    mean, lo, hi = p.mean(0), np.quantile(p, 0.025, axis=0), np.quantile(p, 0.975, axis=0)
    dm = max(abs(mean[i] - tab[s]["mean"]) for i, s in enumerate(names))
    dl = max(abs(lo[i] - tab[s]["low"]) for i, s in enumerate(names))
    dh = max(abs(hi[i] - tab[s]["high"]) for i, s in enumerate(names))
    print(f"stream 2016 vs README: max|dmean| {dm:.4f} |dlow| {dl:.4f} |dhigh| {dh:.4f}; eps {st['mean_stepsize']:.4f} depth {st['mean_treedepth']:.2f} "
          f"{st['n_leapfrog_total'] / st['seconds_total']:.3e} leapfrog/s")
    assert dm <= 0.003 and dl <= 0.012 and dh <= 0.012
    ora = json.load(open(os.path.join(GOLDEN, "oracle_posterior_2016.json")))
    z = np.abs(mean - np.array(ora["mean"])) / (4 * np.array(ora["mcse"]) + 5e-4)
    assert z.max() <= 1.0, z.max()
    assert 0.011 < st["mean_stepsize"] < 0.017 and 7.9 < st["mean_treedepth"] < 8.4


def test_shapes_the_resident_kernel_refuses_are_routed_to_the_streaming_family(pkg, orc_mod, cuda_lib):
    """The envelope of the drop-in: whatever Stan's data block allows and the resident kernel cannot hold
    goes to the streaming family WITHOUT a flag -- fractional `unadjusted_*` (poll_model_2020.stan:22-23: real in [0, 1]),
    more than 63 polls in one (state, day) cell, T = 300 (BASELINE config 2 says "T~300"), N > 1648 -- and matches the oracle."""
    rng = np.random.default_rng(8)
    cases = {}
    d = small_datalist(S=51, T=254, Ns=300, Nn=50, P=40)
    d["unadjusted_state"] = rng.uniform(0, 1, 300); d["unadjusted_national"] = rng.uniform(0, 1, 50)
    cases["fractional unadjusted"] = d
    cases["100 polls per cell"] = small_datalist(S=2, T=3, Ns=600, Nn=30, P=4)
    cases["T=300"] = small_datalist(S=51, T=300, Ns=900, Nn=200, P=60)
    cases["N=4000"] = small_datalist(S=51, T=254, Ns=3200, Nn=800, P=100)
    for name, d in cases.items():
        om = orc_mod.OracleModel(d)
        th = rng.normal(0, 0.5, (2, om.D))
        lp, g = pkg.logp_grad(d, th)          # no force_stream: potus_create / potus_logp_grad route by themselves
        for i in range(2):
            lpo, go = om.logp_grad(th[i])
            assert abs(lp[i] - lpo) <= 1e-8 * max(abs(lpo), 1e3), (name, lp[i], lpo)
            assert np.abs(g[i] - go).max() <= 3e-6 * max(np.abs(go).max(), 1.0), (name, np.abs(g[i] - go).max(), np.abs(go).max())
        fit = pkg.cmdstan_model().sample(data=d, seed=2, chains=2, iter_warmup=6, iter_sampling=2, keep_per_chain=1)
        assert np.all(np.isfinite(fit.sampler_params()["lp__"])) and fit.extract("mu_b").shape == (2, int(d["S"]), int(d["T"]))
