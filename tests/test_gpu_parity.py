"""GPU parity tests proper (run on the B200 box with -m gpu; everything goes through the C-ABI).
The checker is the fp64 oracle (oracle/), pinned in test_oracle.py; tolerances are stated per test."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, small_datalist

pytestmark = pytest.mark.gpu


def _points(D, seed):
    rng = np.random.default_rng(seed)
    return np.stack([np.zeros(D), 0.1 * np.sin(1 + 0.37 * np.arange(D)), rng.uniform(-2, 2, D), 0.5 * rng.standard_normal(D)])


@pytest.mark.parametrize("year", [2016, 2012, 2008])
def test_logp_grad_matches_oracle(pkg, orc_mod, datalists, cuda_lib, year):
    """fp32 state, fp16x2-split tensor-core GEMM (22-bit operands), fp64 energy reductions.
    Tolerance: |lp - lp_oracle| <= 1e-8 |lp| (lp ~ -1e6, so ~1e-2 absolute; observed ~5e-4);
    gradient max error <= 2e-6 max|grad| (observed ~4e-7)."""
    d = datalists[year]
    om = orc_mod.OracleModel(d)
    th = _points(om.D, year)
    lp, g = pkg.logp_grad(d, th)
    for i in range(len(th)):
        lpo, go = om.logp_grad(th[i])
        assert abs(lp[i] - lpo) <= 1e-8 * abs(lpo), (i, lp[i], lpo)
        assert np.abs(g[i] - go).max() <= 2e-6 * np.abs(go).max(), (i, np.abs(g[i] - go).max(), np.abs(go).max())


def test_known_answer_values_on_device(pkg, datalists, cuda_lib):
    kat = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    for y in (2016, 2012, 2008):
        D = kat[str(y)]["D"]
        th = np.stack([np.zeros(D), 0.1 * np.sin(1 + 0.37 * np.arange(D))])
        lp, g = pkg.logp_grad(datalists[y], th)
        assert abs(lp[0] - kat[str(y)]["lp_zero"]) < 1e-8 * abs(lp[0])
        assert abs(lp[1] - kat[str(y)]["lp_sin"]) < 1e-8 * abs(lp[1])
        assert abs(np.linalg.norm(g[0]) - kat[str(y)]["gnorm_zero"]) < 1e-5 * kat[str(y)]["gnorm_zero"]
        assert abs(np.linalg.norm(g[1]) - kat[str(y)]["gnorm_sin"]) < 1e-5 * kat[str(y)]["gnorm_sin"]


@pytest.mark.parametrize("kw", [dict(S=4, T=2, Ns=6, Nn=3), dict(S=7, T=5, Ns=9, Nn=0), dict(S=3, T=4, Ns=1, Nn=2),
                                dict(S=6, T=3, Ns=5, Nn=1, full=False), dict(S=51, T=254, Ns=300, Nn=50, P=40),
                                dict(S=50, T=253, Ns=64, Nn=64, P=3, full=False), dict(S=1, T=2, Ns=2, Nn=1, P=1)])
def test_edge_shapes_match_oracle(pkg, orc_mod, cuda_lib, kw):
    """Ragged / extreme shapes: T=2, empty national set, a single poll, odd and even S, S=51/T=254 (the tile
    limits), polls on day 1 and day T, no-mode variant."""
    d = small_datalist(**kw)
    om = orc_mod.OracleModel(d)
    th = np.random.default_rng(3).normal(0, 0.6, (3, om.D))
    lp, g = pkg.logp_grad(d, th)
    for i in range(3):
        lpo, go = om.logp_grad(th[i])
        assert abs(lp[i] - lpo) <= 1e-7 * max(abs(lpo), 1e3)
        assert np.abs(g[i] - go).max() <= 3e-6 * max(np.abs(go).max(), 1.0)


def test_far_tail_and_nonfinite_inputs(pkg, orc_mod, datalists, cuda_lib):
    """|eta - eta_hat| beyond the centred branch (inits far out) still tracks the oracle; NaN input gives NaN lp."""
    d = datalists[2008]
    om = orc_mod.OracleModel(d)
    th = np.random.default_rng(9).uniform(-6, 6, (1, om.D))
    lp, g = pkg.logp_grad(d, th)
    lpo, go = om.logp_grad(th[0])
    assert abs(lp[0] - lpo) <= 1e-6 * abs(lpo)
    assert np.abs(g[0] - go).max() <= 1e-5 * np.abs(go).max()
    th[0, 100] = np.nan
    lp, g = pkg.logp_grad(d, th)
    assert not np.isfinite(lp[0])


def test_gradient_is_consistent_with_energy_on_device(pkg, datalists, cuda_lib):
    """Size-independent property at full size: directional finite difference of the DEVICE lp equals the
    DEVICE gradient (checks that the fp32 path is a consistent energy/gradient pair)."""
    d = datalists[2016]
    D = 15098
    rng = np.random.default_rng(4)
    th = 0.3 * rng.standard_normal(D)
    v = rng.standard_normal(D); v /= np.linalg.norm(v)
    h = 2e-2
    lp, g = pkg.logp_grad(d, np.stack([th, th + h * v, th - h * v]))
    fd = (lp[1] - lp[2]) / (2 * h)
    assert abs(fd - g[0] @ v) <= 2e-3 * max(1.0, abs(fd))
