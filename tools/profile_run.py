"""Tiny run for ncu captures: 148 chains, a handful of early warm-up iterations (deep trees)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 148
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 1
data = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=1)
st = fit.stats
print("chains", chains, "leapfrogs", st["n_leapfrog_total"], "device s", st["seconds_total"],
      "lf/s", st["n_leapfrog_total"] / st["seconds_total"])
