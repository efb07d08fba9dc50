"""Pins the oracle: literal .stan transcription == closed form == C restatement == finite differences,
and the known-answer values of SURVEY.md section 8(c)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, small_datalist
import potus_oracle as po


def test_known_answer_values(datalists):
    kat = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    survey = {2016: (-1174269.285425, 4609.587580, -1175521.931267, 5685.936326),
              2012: (-681552.248616, 609.979449, -682221.341046, 1526.318034),
              2008: (-638100.940496, 1786.924962, -638129.440538, 873.370008)}
    for y, d in datalists.items():
        _, D = po.block_layout(d)
        assert D == kat[str(y)]["D"]
        lp0, g0 = po.logp_grad_closed(np.zeros(D), d)
        lp1, g1 = po.logp_grad_closed(0.1 * np.sin(1 + 0.37 * np.arange(D)), d)
        s = survey[y]
        assert abs(lp0 - s[0]) < 1e-5 and abs(np.linalg.norm(g0) - s[1]) < 1e-5
        assert abs(lp1 - s[2]) < 1e-5 and abs(np.linalg.norm(g1) - s[3]) < 1e-5


@pytest.mark.parametrize("year", [2016, 2008])
def test_literal_equals_closed_form(datalists, year):
    d = datalists[year]
    _, D = po.block_layout(d)
    th = np.random.default_rng(year).uniform(-1.5, 1.5, D)
    lp_a, g_a = po.logp_grad_literal(th, d)
    lp_b, g_b = po.logp_grad_closed(th, d)
    assert abs(lp_a - lp_b) < 1e-9 * abs(lp_a)
    assert np.abs(g_a - g_b).max() < 1e-10 * np.abs(g_a).max()


@pytest.mark.parametrize("year", [2016, 2012, 2008])
def test_c_oracle_equals_numpy(datalists, orc_mod, year):
    d = datalists[year]
    om = orc_mod.OracleModel(d)
    th = np.random.default_rng(7).uniform(-2, 2, om.D)
    lp_n, g_n = po.logp_grad_closed(th, d)
    for literal in (False, True):
        lp_c, g_c = om.logp_grad(th, literal=literal)
        assert abs(lp_c - lp_n) < 1e-10 * abs(lp_n)
        assert np.abs(g_c - g_n).max() < 1e-10 * np.abs(g_n).max()
    c = om.constrain(th)
    f = po.constrained_draw(th, d)
    assert np.abs(c["mu_b"] - f["mu_b"]).max() < 1e-12 and np.abs(c["polling_bias"] - f["polling_bias"]).max() < 1e-13
    if po.is_full_model(d):
        assert np.abs(c["e_bias"] - f["e_bias"]).max() < 1e-13


@pytest.mark.parametrize("full", [True, False])
def test_finite_differences_small(full):
    d = small_datalist(full=full)
    _, D = po.block_layout(d)
    th = np.random.default_rng(1).normal(0, 0.7, D)
    _, g = po.logp_grad_closed(th, d)
    idx = np.arange(D)
    fd = po.logp_grad_fd(th, d, idx, h=1e-6)
    assert np.abs(fd - g).max() < 2e-5 * max(1.0, np.abs(g).max())
    lp_l, g_l = po.logp_grad_literal(th, d)
    assert np.abs(g_l - g).max() < 1e-10 * np.abs(g).max()


def test_edge_shapes_literal_vs_closed(orc_mod):
    # T=2 (a single walk step), no national polls, one state poll, odd/even S
    for kw in (dict(S=4, T=2, Ns=6, Nn=3), dict(S=7, T=5, Ns=9, Nn=0), dict(S=3, T=4, Ns=1, Nn=2), dict(S=6, T=3, Ns=5, Nn=1, full=False)):
        d = small_datalist(**kw)
        _, D = po.block_layout(d)
        th = np.random.default_rng(5).normal(0, 0.5, D)
        lp_l, g_l = po.logp_grad_literal(th, d)
        lp_c, g_c = po.logp_grad_closed(th, d)
        assert abs(lp_l - lp_c) < 1e-9 * max(1, abs(lp_l)) and np.abs(g_l - g_c).max() < 1e-9 * max(1, np.abs(g_l).max())
        om = orc_mod.OracleModel(d)
        lp_o, g_o = om.logp_grad(th)
        assert abs(lp_o - lp_c) < 1e-9 * max(1, abs(lp_c)) and np.abs(g_o - g_c).max() < 1e-9 * max(1, np.abs(g_c).max())


def test_raw_mu_b_last_column_is_prior_only(datalists):
    d = datalists[2008]
    blocks, D = po.block_layout(d)
    th = np.random.default_rng(2).normal(0, 1, D)
    lp, g = po.logp_grad_closed(th, d)
    S, T = 51, int(d["T"])
    last = slice(S + S * (T - 1), S + S * T)  # raw_mu_b[:, T]
    assert np.allclose(g[last], -th[last])  # never enters mu_b (poll_model_2020.stan:85-86)


def test_philox_known_answer(orc_mod):
    # Random123 kat_vectors: philox4x32-10, counter 0, key 0
    assert orc_mod.rng_words(0, 0, 0, 0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_cumulant_series_of_the_poll_term():
    """The kernels' poll term (csrc/potus_kernel.cu poll_term): with d = eta - eta_hat and p = inv_logit(eta_hat),
    log(1 - p + p e^d) - p d = sum_{k>=2} kappa_k d^k / k!  with the Bernoulli(p) cumulants kappa_2 = v, kappa_3 = v w,
    kappa_4 = v(1-6v), kappa_5 = v w(1-12v), kappa_6 = v(1-30v+120v^2), kappa_7 = v w(1-60v+360v^2),
    kappa_8 = v(1-126v+1680v^2-5040v^3), kappa_9 = v w(1-252v+5040v^2-20160v^3), kappa_10 = v(1-510v+17640v^2-151200v^3+362880v^4),
    v = p(1-p), w = 1-2p.  Restated in numpy (fp64) exactly as the device evaluates it (series through d^10 AND its term-by-term
    derivative) and checked against the closed forms on |d| < 0.4: truncation error < 7e-10 v on the gradient, and the pair is an
    exact energy / gradient pair (finite difference of the series = its derivative)."""
    def series(p, d):
        v, w, x2 = p * (1 - p), 1 - 2 * p, d * d
        e4 = (1 - 6 * v) / 24; e6 = (1 - 30 * v + 120 * v * v) / 720; e8 = (1 - 126 * v + 1680 * v ** 2 - 5040 * v ** 3) / 40320
        e10 = (1 - 510 * v + 17640 * v ** 2 - 151200 * v ** 3 + 362880 * v ** 4) / 3628800
        o5 = (1 - 12 * v) / 120; o7 = (1 - 60 * v + 360 * v * v) / 5040; o9 = (1 - 252 * v + 5040 * v ** 2 - 20160 * v ** 3) / 362880
        ev = x2 * (0.5 + x2 * (e4 + x2 * (e6 + x2 * (e8 + x2 * e10))))
        od = x2 * (1 / 6 + x2 * (o5 + x2 * (o7 + x2 * o9)))
        g = v * (ev + w * d * od)
        dev = d * (1 + x2 * (4 * e4 + x2 * (6 * e6 + x2 * (8 * e8 + x2 * 10 * e10))))
        dod = x2 * (0.5 + x2 * (5 * o5 + x2 * (7 * o7 + x2 * 9 * o9)))
        return g, v * (dev + w * dod)
    P, D = np.meshgrid(np.linspace(0.003, 0.997, 300), np.linspace(-0.4, 0.4, 401))
    g, gp = series(P, D)
    g_exact = np.log1p(P * np.expm1(D)) - P * D
    gp_exact = P * np.exp(D) / (1 - P + P * np.exp(D)) - P
    v = P * (1 - P)
    assert np.abs(g - g_exact).max() < 3e-11 and (np.abs(gp - gp_exact) / v).max() < 3e-9
    h = 1e-5
    fd = (series(P, D + h)[0] - series(P, D - h)[0]) / (2 * h)
    assert np.abs(fd - gp).max() < 1e-9
    # what the poll contributes: f = n [rho_hat d - g],  r = n [rho_hat - g'] == y - n inv_logit(eta)
    n, y, eta = 1873.0, 901.0, 0.113
    eh = np.log((y / n) / (1 - y / n)); p = 1 / (1 + np.exp(-eh)); rho = y / n - p
    gg, ggp = series(p, eta - eh)
    assert abs(n * (rho - ggp) - (y - n / (1 + np.exp(-eta)))) < 1e-7
    ll = lambda e: y * e - n * np.logaddexp(0, e)
    assert abs(n * (rho * (eta - eh) - gg) - (ll(eta) - ll(eh))) < 1e-7
