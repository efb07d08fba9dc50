"""Development aid: compare builds of the streaming kernel (lib/libpotus_b200<suffix>.so) on one synthetic shape:
gradient against the fp64 oracle (plain-evaluation sweep) and the first transitions' tree sizes against each other
(leaf-mode sweep).   python tools/stream_variant_check.py _prev,_d,_h,default"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] == "--one":
    import numpy as np
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import potus_pkg
    pkg = potus_pkg.load()
    import orc
    from us_potus_model_b200 import cabi
    suffix = sys.argv[2]
    cabi._lib = cabi.load_library(os.path.join(ROOT, "us-potus-model_b200", "lib", f"libpotus_b200{suffix}.so"))
    for S, T, Ns, Nn, P in ((64, 300, 3000, 800, 40), (130, 140, 1501, 333, 77)):
        data = pkg.synthetic_datalist(S=S, T=T, N_state=Ns, N_national=Nn, P=P)
        om = orc.OracleModel(data)
        rng = np.random.default_rng(1)
        th = np.stack([0.5 * rng.standard_normal(om.D), rng.uniform(-2, 2, om.D)])
        lp, g = pkg.logp_grad(data, th, force_stream=True)
        for i in range(2):
            lpo, go = om.logp_grad(th[i])
            print(f"lib '{suffix}' S{S} T{T} pt{i}: lp rel {abs(lp[i] - lpo) / abs(lpo):.2e} grad max err / max|grad| {np.abs(g[i] - go).max() / np.abs(go).max():.2e}", flush=True)
        fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=3, iter_warmup=6, iter_sampling=0, keep_per_chain=0, force_stream=True)
        sp = fit.sampler_params()
        print(f"lib '{suffix}' S{S} T{T} n_leapfrog {sp['n_leapfrog__'].astype(int).tolist()} lp[0] {np.round(sp['lp__'][0], 3).tolist()}", flush=True)
else:
    for s in sys.argv[1].split(","):
        subprocess.run([sys.executable, "-u", __file__, "--one", "" if s == "default" else s], check=False)
