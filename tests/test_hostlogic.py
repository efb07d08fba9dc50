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
