"""Per-phase cycle counts of the NUTS kernels (development aid): uses the -DPOTUS_PROF build.
   python tools/phase_clocks.py [chains] [2016|stream2016|syn] [iter_warmup] [iter_sampling]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import potus_pkg
pkg = potus_pkg.load()
from us_potus_model_b200 import cabi
cabi._lib = cabi.load_library(os.path.join(ROOT, "us-potus-model_b200", "lib", "libpotus_b200_prof.so"))
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 148
what = sys.argv[2] if len(sys.argv) > 2 else "2016"
nw = int(sys.argv[3]) if len(sys.argv) > 3 else (40 if what != "syn" else 8)
ns = int(sys.argv[4]) if len(sys.argv) > 4 else (10 if what != "syn" else 2)
data = pkg.synthetic_datalist() if what == "syn" else pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=1, force_stream=(what == "stream2016"))
st = fit.stats
print(what, "lf/s", st["n_leapfrog_total"] / st["seconds_total"], "cycles per leaf at 1.9GHz:", 1.9e9 * chains * st["seconds_total"] / st["n_leapfrog_total"] if chains <= 148 else None)
