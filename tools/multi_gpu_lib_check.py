"""In-library multi-GPU (PotusConfig.n_gpus): timing of one process driving N devices + one ncclAllGather.
   python tools/multi_gpu_lib_check.py [n_gpus] [chains_total] [iter_warmup] [iter_sampling]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
chains = int(sys.argv[2]) if len(sys.argv) > 2 else 296 * G
nw = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ns = int(sys.argv[4]) if len(sys.argv) > 4 else 50
data = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
for g in sorted({1, G}):
    t = time.time()
    fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains // G * g, iter_warmup=nw, iter_sampling=ns, keep_per_chain=3, n_gpus=g)
    st = fit.stats
    mu = fit.extract("mu_b")
    print(json.dumps({"n_gpus": g, "chains": chains // G * g, "iter": [nw, ns], "leapfrogs": st["n_leapfrog_total"], "device_s": st["seconds_total"],
                      "gather_s": st["seconds_gather"], "leapfrog_per_s": st["n_leapfrog_total"] / st["seconds_total"], "wall_s": time.time() - t,
                      "draws": list(mu.shape), "gathered_MB": mu.shape[0] * fit.device_buffer(0)[1] / max(mu.shape[0], 1) * 4 / 1e6 if g > 1 else 0}), flush=True)
    fit.close()
