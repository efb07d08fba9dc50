"""Adapted sampler states of the fp64 oracle on the 2016 list, for the CPU arm of bench.py: 8 chains x (500 warm-up + 1
sampling iteration), Stan defaults; stores each chain's position (a point of the stationary phase), its adapted inverse
metric and step size.  bench.py starts its bounded CPU samples of the SAMPLING phase from these, so that the reference arm
times stationary depth-8 trajectories instead of the first warm-up iterations.   (~4 min on 8 cores)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import potus_pkg  # noqa: E402
pkg = potus_pkg.load()
import orc  # noqa: E402

d = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
om = orc.OracleModel(d)
r = om.sample(chains=8, iter_warmup=500, iter_sampling=1, seed=1843, threads=os.cpu_count(), tree_mode=0, save_theta=True, save_inv_metric=True,
              chain_id_offset=200000)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_adapted_states_2016.npz"), q=r["theta"][:, 0, :].astype(np.float32),
                    inv_metric=r["inv_metric"].astype(np.float32), stepsize=r["stepsize"], seconds=r["seconds"],
                    n_leapfrog=r["n_leapfrog"], treedepth_last=r["stats"][:, -1, 3])
print("seconds", r["seconds"], "eps", r["stepsize"], "depth of the sampling iteration", r["stats"][:, -1, 3])
