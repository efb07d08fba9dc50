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
