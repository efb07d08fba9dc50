"""Ad-hoc GPU validation of the streaming kernel family (development aid; judged tests: tests/test_gpu_stream.py).
   python tools/stream_check.py [grad|sample|all] [max_case]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
import orc  # noqa: E402
from conftest import small_datalist  # noqa: E402


def cases():
    yield "toy S5 T9", small_datalist(S=5, T=9, Ns=40, Nn=12)
    yield "toy S7 T5 no national, no-mode", small_datalist(S=7, T=5, Ns=9, Nn=0, full=False)
    yield "S51 T254 small polls", small_datalist(S=51, T=254, Ns=300, Nn=50, P=40)
    yield "2016 list", pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
    yield "syn S64 T300", pkg.synthetic_datalist(S=64, T=300, N_state=3000, N_national=800, P=40)
    yield "syn S128 T365", pkg.synthetic_datalist(S=128, T=365, N_state=8000, N_national=2000, P=128)
    yield "syn S256 T365 N50k (config 5)", pkg.synthetic_datalist()


def block_report(om, data, err):
    import potus_oracle as po
    blocks, _ = po.block_layout(data)
    o, rep = 0, []
    for name, n in blocks:
        if n:
            rep.append(f"{name}:{np.abs(err[o:o + n]).max():.1e}")
        o += n
    return " ".join(rep)


def check_grad(name, data):
    om = orc.OracleModel(data)
    rng = np.random.default_rng(1)
    th = np.stack([0.5 * rng.standard_normal(om.D), rng.uniform(-2, 2, om.D)])
    t = time.time()
    lp, g = pkg.logp_grad(data, th, force_stream=True)
    dt = time.time() - t
    for i in range(len(th)):
        lpo, go = om.logp_grad(th[i])
        err = np.abs(g[i] - go)
        print(f"[{name}] D={om.D} pt{i}: lp gpu {lp[i]:.5f} oracle {lpo:.5f} rel {abs(lp[i]-lpo)/abs(lpo):.2e} | grad max|err| {err.max():.3e} "
              f"rel {err.max()/np.abs(go).max():.2e} at {int(err.argmax())}", flush=True)
        print("      ", block_report(om, data, g[i] - go), flush=True)
    print(f"[{name}] call {dt:.2f}s", flush=True)
    return om


def check_sample(name, data, chains=2, nw=8, ns=2, oracle=True):
    t = time.time()
    fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=1, force_stream=True)
    st = fit.stats
    sp = fit.sampler_params()
    print(f"[{name}] sample {chains}x({nw}+{ns}): wall {time.time()-t:.2f}s device {st['seconds_total']:.3f}s leapfrogs {st['n_leapfrog_total']} "
          f"-> {st['n_leapfrog_total']/max(st['seconds_total'],1e-9):.3e} lf/s", flush=True)
    print("   gpu depth", sp["treedepth__"][0].astype(int).tolist(), "nleap", sp["n_leapfrog__"][0].astype(int).tolist(), flush=True)
    print("   gpu eps", np.round(sp["stepsize__"][0], 5).tolist(), "acc", np.round(sp["accept_stat__"][0], 3).tolist(), flush=True)
    print("   gpu lp", np.round(sp["lp__"][0], 2).tolist(), flush=True)
    if oracle:
        om = orc.OracleModel(data)
        r = om.sample(chains=chains, iter_warmup=nw, iter_sampling=ns, seed=1843, threads=chains, tree_mode=1)
        print("   ora depth", r["stats"][0, :, 3].astype(int).tolist(), "nleap", r["stats"][0, :, 4].astype(int).tolist(), flush=True)
        print("   ora eps", np.round(r["stats"][0, :, 2], 5).tolist(), "acc", np.round(r["stats"][0, :, 1], 3).tolist(), flush=True)
        print("   ora lp", np.round(r["stats"][0, :, 0], 2).tolist(), flush=True)
        same = all(np.array_equal(sp["n_leapfrog__"][c], r["stats"][c, :, 4]) for c in range(chains))
        print("   decisions identical:", same, flush=True)
        th = fit.theta()
        c = om.constrain(th[0])
        mu = fit.extract("mu_b")[0]
        print("   mu_b vs oracle constrain max err", np.abs(c["mu_b"] - mu).max(), "polling_bias", np.abs(c["polling_bias"] - fit.extract("polling_bias")[0]).max(), flush=True)
    fit.close()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    maxc = int(sys.argv[2]) if len(sys.argv) > 2 else 99
    for i, (name, data) in enumerate(cases()):
        if i >= maxc:
            break
        if what in ("grad", "all"):
            check_grad(name, data)
        if what in ("sample", "all"):
            big = int(data["S"]) > 64
            check_sample(name, data, chains=2, nw=6 if big else 10, ns=1 if big else 2, oracle=True)
