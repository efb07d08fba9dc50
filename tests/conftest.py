import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import potus_pkg
    return potus_pkg.load()


@pytest.fixture(scope="session")
def orc_mod():
    import orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def datalists(pkg):
    return {y: pkg.load_npz(os.path.join(GOLDEN, f"datalist_{y}.npz")) for y in (2016, 2012, 2008)}


@pytest.fixture(scope="session")
def cuda_lib(pkg):
    """The in-tree CUDA library; building it here is the same nvcc recipe __graft_entry__.build() runs."""
    from us_potus_model_b200 import build, cabi
    build.build()
    return cabi.load_library()


def small_datalist(S=5, T=9, Ns=40, Nn=12, P=6, full=True, seed=3, nat_days=None):
    """Synthetic data list with the real list's structure at toy size (edge-case shapes)."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((S, S))
    cov = A @ A.T / S * 0.005 + 0.002 * np.eye(S)
    w = rng.dirichlet(np.ones(S))
    d = dict(
        N_national_polls=Nn, N_state_polls=Ns, T=T, S=S, P=P, M=3, Pop=2,
        state=rng.integers(1, S + 1, Ns).astype(np.int32), state_weights=w,
        day_state=rng.integers(1, T + 1, Ns).astype(np.int32),
        day_national=(rng.integers(1, T + 1, Nn) if nat_days is None else np.asarray(nat_days)).astype(np.int32),
        poll_state=rng.integers(1, P + 1, Ns).astype(np.int32), poll_national=rng.integers(1, P + 1, Nn).astype(np.int32),
        unadjusted_state=rng.integers(0, 2, Ns).astype(float), unadjusted_national=rng.integers(0, 2, Nn).astype(float),
        n_two_share_state=rng.integers(200, 3000, Ns).astype(np.int32),
        n_two_share_national=rng.integers(500, 60000, Nn).astype(np.int32),
        sigma_measure_noise_national=0.04, sigma_measure_noise_state=0.04, mu_b_prior=rng.normal(0, 0.3, S),
        sigma_c=0.06, sigma_m=0.04, sigma_pop=0.04, sigma_e_bias=0.02, state_covariance_0=cov,
        polling_bias_scale=0.052, mu_b_T_scale=0.12, random_walk_scale=0.0115,
    )
    d["n_democrat_state"] = rng.binomial(d["n_two_share_state"], 0.5).astype(np.int32)
    d["n_democrat_national"] = rng.binomial(d["n_two_share_national"], 0.52).astype(np.int32)
    if Ns > 1:  # force polls on the first and the last day
        d["day_state"][0], d["day_state"][1] = 1, T
    if full:
        d.update(poll_mode_state=rng.integers(1, 4, Ns).astype(np.int32), poll_mode_national=rng.integers(1, 4, Nn).astype(np.int32),
                 poll_pop_state=rng.integers(1, 3, Ns).astype(np.int32), poll_pop_national=rng.integers(1, 3, Nn).astype(np.int32))
    return d
