"""A/B timing of development builds of the CUDA library (not used by tests or bench):
   python tools/ab_run.py <suffix>[,<suffix>...] [chains] [iter_warmup] [iter_sampling]
loads us-potus-model_b200/lib/libpotus_b200<suffix>.so for each suffix in a fresh process and prints leapfrog/s."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import potus_pkg
    pkg = potus_pkg.load()
    from us_potus_model_b200 import cabi
    suffix, chains, nw, ns = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    cabi._lib = cabi.load_library(os.path.join(ROOT, "us-potus-model_b200", "lib", f"libpotus_b200{suffix}.so"))
    data = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
    fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=1)
    st = fit.stats
    sp = fit.sampler_params()
    print(f"lib '{suffix}': chains {chains} leapfrogs {st['n_leapfrog_total']} device s {st['seconds_total']:.4f} "
          f"lf/s {st['n_leapfrog_total'] / st['seconds_total']:.0f} checksum lp {sp['lp__'].sum():.3f} nleap {int(sp['n_leapfrog__'].sum())}",
          flush=True)
else:
    sufs = sys.argv[1].split(",")
    rest = (sys.argv[2:] + ["148", "40", "10"][len(sys.argv) - 2:])[:3]
    for s in sufs:
        subprocess.run([sys.executable, "-u", __file__, "--one", "" if s == "default" else s, *rest], check=False)
