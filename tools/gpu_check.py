"""Ad-hoc GPU validation (development aid; the judged tests live in tests/)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
import orc  # noqa: E402


def check_grad(year):
    data = pkg.load_npz(os.path.join(ROOT, "tests", "golden", f"datalist_{year}.npz"))
    om = orc.OracleModel(data)
    rng = np.random.default_rng(0)
    pts = [np.zeros(om.D), 0.1 * np.sin(1 + 0.37 * np.arange(om.D)), rng.uniform(-2, 2, om.D), rng.normal(0, 1, om.D) * 0.5]
    th = np.stack(pts)
    t = time.time()
    lp, g = pkg.logp_grad(data, th)
    dt = time.time() - t
    for i in range(len(th)):
        lpo, go = om.logp_grad(th[i])
        err = np.abs(g[i] - go)
        j = int(err.argmax())
        print(f"[{year}] pt{i}: lp gpu {lp[i]:.5f} oracle {lpo:.5f} diff {lp[i]-lpo:.3e} | grad max|err| {err.max():.3e} at {j} "
              f"(gpu {g[i][j]:.5f} ora {go[j]:.5f}) max|g| {np.abs(go).max():.3f} rel {err.max()/np.abs(go).max():.2e}")
        # per-block error report
        import potus_oracle as po
        blocks, _ = po.block_layout(data)
        o = 0
        rep = []
        for name, n in blocks:
            rep.append(f"{name}:{np.abs(err[o:o+n]).max():.1e}")
            o += n
        print("      ", " ".join(rep))
    print(f"[{year}] logp_grad call {dt:.2f}s")
    return data, om


def run_sampler(data, chains, nw, ns, keep=2):
    t = time.time()
    fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=keep)
    dt = time.time() - t
    st = fit.stats
    sp = fit.sampler_params()
    print(f"sample chains={chains} {nw}+{ns}: wall {dt:.2f}s device {st['seconds_total']:.3f}s (warmup {st['seconds_warmup']:.3f}) "
          f"leapfrogs {st['n_leapfrog_total']} -> {st['n_leapfrog_total']/max(st['seconds_total'],1e-9):.3e} lf/s")
    print("   mean eps", st["mean_stepsize"], "accept", st["mean_accept_stat"], "depth", st["mean_treedepth"], "div", st["n_divergent_sampling"])
    print("   chain0 depths", sp["treedepth__"][0].astype(int).tolist()[:60])
    print("   chain0 eps", np.round(sp["stepsize__"][0][:: max(1, (nw + ns) // 12)], 4).tolist())
    print("   chain0 lp", np.round(sp["lp__"][0][:: max(1, (nw + ns) // 12)], 1).tolist())
    print("   chain0 accept", np.round(sp["accept_stat__"][0][:: max(1, (nw + ns) // 12)], 3).tolist())
    return fit


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    data, om = check_grad(2016)
    if what in ("all", "grad"):
        check_grad(2008)
    if what in ("all", "sample"):
        fit = run_sampler(data, 8, 40, 10)
        r = om.sample(chains=2, iter_warmup=40, iter_sampling=10, threads=2, tree_mode=1)
        print("oracle chain0 depths", r["stats"][0, :, 3].astype(int).tolist())
        print("oracle chain0 eps", np.round(r["stats"][0, ::4, 2], 4).tolist())
        print("oracle chain0 lp", np.round(r["stats"][0, ::4, 0], 1).tolist())
        mu = fit.extract("mu_b")
        print("mu_b[:, :, T-1] mean over draws (first 5 states):", mu[:, :5, -1].mean(0))
        fit2 = run_sampler(data, 148, 100, 50)
