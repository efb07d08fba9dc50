"""A/B of streaming-kernel builds on config 5: python tools/stream_ab.py <suffix>[,<suffix>...] [chains] [nw] [ns]"""
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
    fit = pkg.cmdstan_model().sample(data=pkg.synthetic_datalist(), seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=1)
    st, sp = fit.stats, fit.sampler_params()
    print(f"lib '{suffix}': leapfrogs {st['n_leapfrog_total']} device s {st['seconds_total']:.3f} lf/s {st['n_leapfrog_total'] / st['seconds_total']:.0f} "
          f"sampling-phase lf/s {st['n_leapfrog_sampling'] / max(st['seconds_sampling'], 1e-9):.0f} checksum lp {sp['lp__'].sum():.3f}", flush=True)
else:
    rest = (sys.argv[2:] + ["148", "8", "2"][len(sys.argv) - 2:])[:3]
    for s in sys.argv[1].split(","):
        subprocess.run([sys.executable, "-u", __file__, "--one", "" if s == "default" else s, *rest], check=False)
