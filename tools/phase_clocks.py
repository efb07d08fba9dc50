"""Per-phase cycle counts of the NUTS kernel (development aid): uses the -DPOTUS_PROF build."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import potus_pkg
pkg = potus_pkg.load()
from us_potus_model_b200 import cabi
cabi._lib = cabi.load_library(os.path.join(ROOT, "us-potus-model_b200", "lib", "libpotus_b200_prof.so"))
data = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 148
fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=40, iter_sampling=10, keep_per_chain=1)
st = fit.stats
print("lf/s", st["n_leapfrog_total"] / st["seconds_total"], "cycles per leaf at 1.9GHz:", 1.9e9 * chains * st["seconds_total"] / st["n_leapfrog_total"] if chains <= 148 else None)
