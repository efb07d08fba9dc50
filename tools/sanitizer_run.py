"""Small run for compute-sanitizer that also reports how each transition ended (development aid).
   python tools/sanitizer_run.py [chains] [iter_warmup] [iter_sampling]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 1
data = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=1)
sp = fit.sampler_params()
depth, nleap, div = sp["treedepth__"].astype(int), sp["n_leapfrog__"].astype(int), sp["divergent__"].astype(int)
full = nleap == (2 ** depth - 1)          # every doubling completed: ended at the top level (persist / max depth)
print(f"transitions {depth.size}: divergent {int(div.sum())}; ended inside a subtree (in-loop break, n_leapfrog < 2^depth - 1) "
      f"{int((~full & (div == 0)).sum())}; ended at the top level {int((full & (div == 0)).sum())}; "
      f"leapfrogs {int(nleap.sum())}; depth histogram {np.bincount(depth.ravel()).tolist()}", flush=True)
