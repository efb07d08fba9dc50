"""The committed data-list fixtures carry the reference's shapes and poll totals (SURVEY.md 8(c) anchors);
when the reference checkout is present (authoring container) the builder must regenerate them exactly."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_fixture_anchors(datalists):
    kat = json.load(open(os.path.join(GOLDEN, "known_answers.json")))
    expect = {2016: (1258, 361, 254, 161), 2012: (966, 191, 251, 160), 2008: (960, 251, 247, 90)}
    for y, d in datalists.items():
        assert (d["N_state_polls"], d["N_national_polls"], d["T"], d["P"]) == expect[y]
        k = kat[str(y)]
        assert int(d["n_democrat_state"].sum()) == k["sum_y_state"] and int(d["n_two_share_state"].sum()) == k["sum_n_state"]
        assert d["state"].min() >= 1 and d["state"].max() <= 51 and d["day_state"].max() <= d["T"]
        assert np.allclose(d["state_weights"].sum(), 1.0)
        assert np.allclose(d["state_covariance_0"], d["state_covariance_0"].T)
        assert np.linalg.eigvalsh(d["state_covariance_0"]).min() > 0
    d = datalists[2016]
    assert (int(d["n_democrat_state"].sum()), int(d["n_two_share_state"].sum())) == (455258, 881209)
    assert (int(d["n_democrat_national"].sum()), int(d["n_two_share_national"].sum())) == (434638, 822534)
    assert int(d["unadjusted_state"].sum()) == 1027 and int(d["unadjusted_national"].sum()) == 241
    assert len(set(zip(d["state"], d["day_state"]))) == 1106 and len(set(d["day_national"])) == 180
    w, c0 = d["state_weights"], d["state_covariance_0"]
    nat_sd = np.sqrt(w @ c0 @ w)
    assert abs(nat_sd - 0.0498188) < 1e-6
    assert abs(d["mu_b_T_scale"] / nat_sd - 2.408728) < 1e-5 and abs(d["random_walk_scale"] / nat_sd - 0.231780) < 1e-5
    assert abs(c0[0, 0] - 0.00520530) < 1e-7 and abs(c0[0, 1] - 0.00236598) < 1e-7
    assert list(d["_state_names"][:4]) == ["AK", "AL", "AR", "AZ"]
    assert np.allclose(d["mu_b_prior"][:3], [-0.368677, -0.492543, -0.515219], atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference checkout only exists in the authoring container")
def test_builder_regenerates_fixtures(pkg, datalists):
    for y, d in datalists.items():
        fresh = pkg.build_datalist(y, "/root/reference/data")
        for k, v in d.items():
            a, b = np.asarray(v), np.asarray(fresh[k])
            assert a.shape == b.shape, k
            if a.dtype.kind in "fc":
                assert np.allclose(a, b, rtol=1e-12, atol=1e-14), k
            else:
                assert np.array_equal(a, b), k


def test_make_positive_definite_floor(pkg):
    m = np.array([[1.0, 0.9, 0.9], [0.9, 1.0, -0.9], [0.9, -0.9, 1.0]])
    assert np.linalg.eigvalsh(m).min() < 0
    p = pkg.datalist.make_positive_definite(m)
    lam = np.linalg.eigvalsh(p)
    assert lam.min() > 0 and lam.min() < 1e-12  # floored at 2*tol, not shifted


def test_synthetic_datalist_shapes(pkg):
    d = pkg.synthetic_datalist(S=12, T=20, N_state=200, N_national=50, P=9)
    assert d["state"].max() <= 12 and d["day_state"].max() <= 20 and len(d["n_democrat_national"]) == 50
    assert np.all(d["n_democrat_state"] <= d["n_two_share_state"])
