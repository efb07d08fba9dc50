"""Config-5 (S=256, T=365, N=50k) run of the streaming kernel: python tools/stream_syn_run.py [chains] [iter_warmup] [iter_sampling]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 148
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 2
data = pkg.synthetic_datalist()
t = time.time()
fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=1)
st = fit.stats
sp = fit.sampler_params()
print(f"config5 chains {chains} {nw}+{ns}: wall {time.time()-t:.2f}s device {st['seconds_total']:.3f}s (warmup {st['seconds_warmup']:.3f}) leapfrogs "
      f"{st['n_leapfrog_total']} -> {st['n_leapfrog_total']/st['seconds_total']:.4e} lf/s; sampling phase {st['n_leapfrog_sampling']/max(st['seconds_sampling'],1e-9):.4e} lf/s", flush=True)
print("  chain0 depth", sp["treedepth__"][0].astype(int).tolist(), flush=True)
print("  chain0 eps", np.round(sp["stepsize__"][0], 5).tolist(), flush=True)
print("  chain0 acc", np.round(sp["accept_stat__"][0], 3).tolist(), flush=True)
print("  chain0 lp", np.round(sp["lp__"][0], 1).tolist(), "checksum nleap", int(sp["n_leapfrog__"].sum()), "lp", float(sp["lp__"].sum()), flush=True)
print("  mean depth by iter", np.round(sp["treedepth__"].mean(0), 2).tolist(), flush=True)
if ns >= 20:
    sm = fit.summary()
    print(f"  sampling phase: divergent {st['n_divergent_sampling']}, mean accept {st['mean_accept_stat']:.3f}, mean depth {st['mean_treedepth']:.2f}, mean eps {st['mean_stepsize']:.5f}; "
          f"ESS of the {len(sm['ess'])} monitored scalars over {chains * ns} draws: min {np.nanmin(sm['ess']):.0f} median {np.nanmedian(sm['ess']):.0f}; split R-hat max {np.nanmax(sm['rhat']):.3f}; "
          f"min-ESS/s {np.nanmin(sm['ess']) / st['seconds_total']:.1f}", flush=True)
    print("  eps by iteration (mean over chains, every 10th):", np.round(sp["stepsize__"].mean(0)[::10], 5).tolist(), flush=True)
    print("  depth by iteration (mean over chains, every 10th):", np.round(sp["treedepth__"].mean(0)[::10], 2).tolist(), flush=True)
