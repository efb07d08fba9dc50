"""Small run for compute-sanitizer that also reports how each transition ended (development aid).
   python tools/sanitizer_run.py [chains] [iter_warmup] [iter_sampling] [adapted] [stream]
`adapted`: seed the chains from the committed adapted oracle states (potus_set_state, iter_warmup forced to 0), so that the
transitions are stationary depth-8 trajectories ending through the in-subtree U-turn break."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 1
adapted = "adapted" in sys.argv[4:]
stream = "stream" in sys.argv[4:]
offset = next((int(a.split("=")[1]) for a in sys.argv[4:] if a.startswith("offset=")), 0)   # global id of the first chain
data = pkg.load_npz(os.path.join(ROOT, "tests", "golden", "datalist_2016.npz"))
if "toy" in sys.argv[4:]:   # a toy list (S=7, T=5, 9 polls, no-mode): short trajectories whose trees often end by the in-subtree U-turn break
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import small_datalist
    data = small_datalist(S=7, T=5, Ns=9, Nn=0, full=False)
state = None
if adapted:
    st = np.load(os.path.join(ROOT, "tests", "golden", "oracle_adapted_states_2016.npz"))
    idx = (offset + np.arange(chains)) % st["q"].shape[0]
    state = dict(theta=st["q"][idx].astype(np.float64), stepsize=st["stepsize"][idx], inv_metric=st["inv_metric"][idx].astype(np.float64))
    nw = 0
fit = pkg.cmdstan_model().sample(data=data, seed=1843, chains=chains, iter_warmup=nw, iter_sampling=ns, keep_per_chain=1, state=state, force_stream=stream, chain_id_offset=offset)
sp = fit.sampler_params()
depth, nleap, div = sp["treedepth__"].astype(int), sp["n_leapfrog__"].astype(int), sp["divergent__"].astype(int)
full = nleap == (2 ** depth - 1)          # every doubling completed: ended at the top level (persist / max depth)
print(f"{'stream' if stream else 'resident'} kernel, {'adapted states' if adapted else 'random inits'}: transitions {depth.size}: divergent {int(div.sum())}; "
      f"ended inside a subtree (in-loop U-turn break, n_leapfrog < 2^depth - 1) {int((~full & (div == 0)).sum())}; ended at the top level "
      f"{int((full & (div == 0)).sum())}; leapfrogs {int(nleap.sum())}; depth histogram {np.bincount(depth.ravel()).tolist()}", flush=True)
