"""Host-side mirror of the reference's boundary: list marshalling (R's sloppy types), model-variant
selection, Stan-style diagnostics, chain sharding and the all-gather used for N>1 (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_marshal_accepts_R_style_doubles_and_ignores_extra_names(datalists):
    from us_potus_model_b200 import cabi
    d = dict(datalists[2012])
    d["state"] = d["state"].astype(np.float64)          # as.numeric(factor) arrives as REALSXP (final_2012.R:136)
    d["n_democrat_state"] = d["n_democrat_state"].astype(np.float64)
    d["S"] = 51.0
    d["sigma_a"] = 0.012; d["current_T"] = 250           # carried but unused names (final_2012.R:489,504)
    pd, keep = cabi.marshal_data(d)
    assert pd.S == 51 and pd.N_state_polls == 966 and not pd.poll_mode_state
    d["state"] = d["state"] + 0.5
    with pytest.raises(ValueError):
        cabi.marshal_data(d)
    d2 = dict(datalists[2012]); del d2["mu_b_prior"]
    with pytest.raises(KeyError):
        cabi.marshal_data(d2)


def test_variant_selection(pkg, datalists):
    m = pkg.cmdstan_model("scripts/model/poll_model_2020_no_mode_adjustment.stan")
    assert m.variant == "no_mode"
    assert pkg.cmdstan_model("scripts/model/poll_model_2020.stan").variant == "full"
    with pytest.raises(ValueError):
        pkg.cmdstan_model("scripts/deprecated/Stan/Refactored/poll_model_v14.stan")
    with pytest.raises(ValueError, match="poll_mode"):
        pkg.cmdstan_model("poll_model_2020.stan").sample(data=datalists[2008], chains=1, iter_warmup=1, iter_sampling=1)


def test_ess_iid_and_ar1(pkg):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 2000))
    e = pkg.diagnostics.ess(x)
    assert 6500 < e < 9500
    rho = 0.8
    y = np.zeros((4, 4000))
    eps = rng.standard_normal((4, 4000))
    for t in range(1, 4000):
        y[:, t] = rho * y[:, t - 1] + eps[:, t]
    e = pkg.diagnostics.ess(y)
    expect = 16000 * (1 - rho) / (1 + rho)
    assert 0.7 * expect < e < 1.4 * expect
    assert abs(pkg.diagnostics.rhat(x) - 1) < 0.01
    assert pkg.diagnostics.rhat(x + np.arange(4)[:, None]) > 1.3
    eb = pkg.diagnostics.ess_bulk(x)
    assert 6000 < eb < 9800


def test_shard_chains_partition():
    sys.path.insert(0, ROOT)
    import bench
    for total, world in ((8192, 8), (1024, 4), (10, 4), (7, 8)):
        seen = []
        for r in range(world):
            off, n = bench.shard_chains(total, world, r)
            seen += list(range(off, off + n))
        assert seen == list(range(total))


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    off, n = bench.shard_chains(6, world, rank)
    # rank-stamped payload: [local chains, keep, draw_len] with value = global chain id * 1000 + slot
    local = torch.stack([torch.full((2, 5), float((off + c) * 1000)) + torch.arange(2)[:, None] for c in range(n)])
    out = bench.allgather_draws(local)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_ordering_gloo_world2():
    """N>1 path on CPU: the all-gather returns rank-major blocks, i.e. draws ordered by global chain id."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        a = got[r].reshape(6, 2, 5)
        assert np.array_equal(a[:, 0, 0], np.arange(6) * 1000.0) and np.array_equal(a[:, 1, 0], np.arange(6) * 1000.0 + 1)


def test_brier_restatement_reproduces_published_scores(pkg, datalists):
    """Feeding the reference's own published P(win) column (README tables, 3 d.p.) through the restated Brier
    computation (README.Rmd:378-390) gives the published scores (README.md:75,169,260) to the rounding of the inputs."""
    import json
    from conftest import GOLDEN
    tabs = json.load(open(os.path.join(GOLDEN, "readme_tables.json")))
    for year, d in datalists.items():
        rows = {r["state"]: r for r in tabs[str(year)]}
        states = [str(s) for s in d["_state_names"]]
        prob = np.array([rows[s]["prob"] for s in states])
        b = pkg.postprocess.brier_scores(prob, states, d["_ev_state"], year)
        pub = pkg.postprocess.PUBLISHED_BRIER[year]
        assert abs(b["ev_wtd_brier"] - pub[0]) < 4e-4 and abs(b["unwtd_brier"] - pub[1]) < 4e-4, (year, b, pub)
        assert b["states_correct"] == pub[2]
    sh = np.random.default_rng(0).uniform(0.3, 0.7, (100, 51))
    ec = pkg.postprocess.electoral_college(sh, datalists[2016]["_ev_state"])
    assert 0 <= ec["prob"] <= 1 and datalists[2016]["_ev_state"].sum() == 538


class _StubFit:
    """Quacks like PotusFit with draws taken from the fp64 oracle (no GPU here): exercises the CSV layer."""

    def __init__(self, data, theta, chains, keep, iter_warmup, iter_sampling):
        from types import SimpleNamespace
        from oracle import potus_oracle as po
        self.data, self._theta = data, theta
        self.cfg = SimpleNamespace(chains=chains, iter_warmup=iter_warmup, iter_sampling=iter_sampling, seed=1843,
                                   adapt_delta=0.8, max_treedepth=10, init_radius=2.0, chain_id_offset=0)
        self.model_name = "poll_model_2020" if "poll_mode_state" in data else "poll_model_2020_no_mode_adjustment"
        self.stats = dict(seconds_warmup=1.5, seconds_sampling=2.5)
        self.n_draws = chains * keep
        self._cd = [po.constrained_draw(t, data) for t in theta]
        self._fw = [po.forward_closed(t, data) for t in theta]
        rng = np.random.default_rng(5)
        n_it = iter_warmup + iter_sampling
        self._sp = {k: rng.random((chains, n_it)) for k in
                    ("lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__")}
        self._im = rng.random((chains, theta.shape[1])) + 0.1

    def theta(self):
        return self._theta

    def extract(self, pars):
        return {p: np.stack([c[p] for c in self._cd]) for p in pars}

    def sampler_params(self, inc_warmup=True):
        return self._sp if inc_warmup else {k: v[:, self.cfg.iter_warmup:] for k, v in self._sp.items()}

    def inv_metric(self):
        return self._im


@pytest.mark.parametrize("year", [2016, 2008])
def test_stan_csv_layout_and_roundtrip(pkg, datalists, year, tmp_path):
    """f3: CmdStan CSV columns in Stan's order (parameters, transformed parameters, GQ; column-major, 1-based),
    values equal to the oracle's transformed parameters, and a reader round trip shaped like rstan::extract."""
    sys.path.insert(0, ROOT)
    from oracle import potus_oracle as po
    sc = pkg.stancsv
    data = datalists[year]
    S, T = int(data["S"]), int(data["T"])
    names = sc.column_names(data)
    if year == 2016:
        assert len(names) == 7 + 15098 + 15301 + 12954           # ~43k doubles per draw
        assert names[7 + 51 + 12954] == "raw_mu_c.1" and "sigma_rho" in names and "mu_e_bias" in names
    else:
        assert "sigma_rho" not in names and "mu_m.1" not in names and "e_bias.1" not in names
    assert names[:8] == ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__",
                         "raw_mu_b_T.1"]
    assert names[7 + S: 7 + S + 3] == ["raw_mu_b.1.1", "raw_mu_b.2.1", "raw_mu_b.3.1"]       # column-major
    assert names[-1] == f"predicted_score.{T}.{S}" and names[-2] == f"predicted_score.{T - 1}.{S}"
    D = po.block_layout(data)[1]
    rng = np.random.default_rng(11)
    chains, keep, nw, ns = 2, 2, 10, 8
    theta = 0.3 * rng.standard_normal((chains * keep, D))
    fit = _StubFit(data, theta, chains, keep, nw, ns)
    paths = sc.write_stan_csv(fit, str(tmp_path), sig_figs=12)
    assert [os.path.basename(p) for p in paths] == [f"{fit.model_name}-1.csv", f"{fit.model_name}-2.csv"]
    out = sc.read_stan_csv(paths)
    assert out["names"] == names
    cfgd = out["config"]
    assert cfgd["num_samples"] == "8" and cfgd["thin"] == "4" and cfgd["num_warmup"] == "10" and cfgd["seed"] == "1843"
    assert cfgd["stan_version_minor"] == "24" and cfgd["model"] == fit.model_name + "_model"
    dr = out["draws"]
    assert dr["mu_b"].shape == (keep, chains, S, T) and dr["predicted_score"].shape == (keep, chains, T, S)
    for c in range(chains):
        for k in range(keep):
            r = c * keep + k
            fw = fit._fw[r]
            np.testing.assert_allclose(dr["mu_b"][k, c], fw["mu_b"], rtol=0, atol=1e-9)
            np.testing.assert_allclose(dr["predicted_score"][k, c], po.sigmoid(fw["mu_b"]).T, atol=1e-9)
            np.testing.assert_allclose(dr["national_mu_b_average"][k, c], fw["nat_avg"], atol=1e-9)
            np.testing.assert_allclose(dr["national_polling_bias_average"][k, c], fw["nat_pb"], atol=1e-9)
            np.testing.assert_allclose(dr["raw_mu_b"][k, c], fw["Z"], atol=1e-9)
            if year == 2016:
                np.testing.assert_allclose(dr["sigma_rho"][k, c], fw["sigma_rho"], atol=1e-9)
                np.testing.assert_allclose(dr["rho_e_bias"][k, c], fw["rho"], atol=1e-9)
                np.testing.assert_allclose(dr["mu_e_bias"][k, c], fw["mu_e"], atol=1e-9)
            # kept iterations are thin*k of the sampling phase (CmdStan's thin)
            it = nw + 4 * k
            assert abs(out["sampler_params"]["stepsize__"][k, c] - fit._sp["stepsize__"][c, it]) < 1e-9
        np.testing.assert_allclose(out["inv_metric"][c], fit._im[c], rtol=1e-5)
    # the likelihood through the CSV's own logit_pi columns reproduces the oracle's log density
    th0 = theta[0]
    lp = po.logp_grad_closed(th0, data)[0]
    eta_s, eta_n = dr["logit_pi_democrat_state"][0, 0], dr["logit_pi_democrat_national"][0, 0]
    ll = (np.asarray(data["n_democrat_state"]) * eta_s - np.asarray(data["n_two_share_state"]) * po.softplus(eta_s)).sum() + \
         (np.asarray(data["n_democrat_national"]) * eta_n - np.asarray(data["n_two_share_national"]) * po.softplus(eta_n)).sum()
    par = po.split(th0, data)
    prior = -0.5 * sum((v ** 2).sum() for k, v in par.items() if k not in ("rho_e_bias",))
    if year == 2016:
        rho = po.sigmoid(par["rho_e_bias"][0])
        prior += -0.5 * ((rho - 0.7) / 0.1) ** 2 + np.log(rho * (1 - rho))
    assert abs((ll + prior) - lp) < 1e-6 * abs(lp)


def test_short_or_long_vectors_are_rejected_before_the_library_sees_them(datalists):
    """marshal_data checks every vector length against N_state_polls / N_national_polls / S (Stan: "mismatch in dimension
    declared and found in context"); without it validate() would index past the end of a short vector."""
    from us_potus_model_b200 import cabi
    base = datalists[2016]
    for key in ("state", "n_two_share_national", "unadjusted_state", "mu_b_prior"):
        d = dict(base); d[key] = np.asarray(base[key])[:-1]
        with pytest.raises(ValueError, match="mismatch in dimension.*" + key):
            cabi.marshal_data(d)
    d = dict(base); d["state_covariance_0"] = np.asarray(base["state_covariance_0"])[:50, :50]
    with pytest.raises(ValueError, match="state_covariance_0"):
        cabi.marshal_data(d)


def test_reference_arm_line_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the same metric/unit,
    `impl: reference`, both gradient forms of the C restatement timed from the committed adapted oracle states, zero copy
    bytes.  Runs here, on the host cores (a few seconds)."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-transitions", "2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "leapfrog steps/sec" and line["unit"] == "leapfrog/s" and line["higher_is_better"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value_literal"] > 0 and cb["value_collapsed"] > 0 and cb["value"] == line["value"]
    assert "adapted oracle states" in cb["sample"] and cb["ess_per_sec_committed_run"]["min"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "leapfrog/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["steps"] == 1 and line["config"]["chains"] == cb["cores"]
