"""Fixture for the a12 parity test (tests/test_gpu_sampler.py::test_adaptation_matches_oracle_distribution):
the fp64 oracle's warm-up adaptation on the 2016 list -- 64 chains x 500 warm-up iterations (Stan defaults:
windows end at 99/149/249/449), iterative tree (the kernel's), collapsed gradient.  Stores, per coordinate,
mean and SD over chains of the adapted inverse metric, the per-chain final step size, and the per-iteration
mean step size over chains.     python tests/golden/make_oracle_adaptation.py [chains]   (~15 min on 8 cores)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import potus_pkg  # noqa: E402
pkg = potus_pkg.load()
import orc  # noqa: E402

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
om = orc.OracleModel(d)
r = om.sample(chains=chains, iter_warmup=500, iter_sampling=0, seed=1843, threads=os.cpu_count(), tree_mode=1, save_inv_metric=True,
              chain_id_offset=100000)   # ids disjoint from the GPU run's: independent chains, same law
im = r["inv_metric"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_adaptation_2016.npz"),
                    inv_metric_mean=im.mean(0).astype(np.float32), inv_metric_sd=im.std(0, ddof=1).astype(np.float32),
                    log_inv_metric_mean=np.log(im).mean(0).astype(np.float32), log_inv_metric_sd=np.log(im).std(0, ddof=1).astype(np.float32),
                    stepsize=r["stepsize"], stepsize_by_iter=r["stats"][:, :, 2].mean(0), treedepth_by_iter=r["stats"][:, :, 3].mean(0),
                    chains=chains, seconds=r["seconds"])
print("chains", chains, "seconds", r["seconds"], "eps mean", r["stepsize"].mean(), "sd", r["stepsize"].std(ddof=1))
