"""B200-native NUTS sampler for The Economist's dynamic multilevel poll model
(scripts/model/poll_model_2020.stan of TheEconomist/us-potus-model) -- host-side package.

Holds only what the hot path needs:
  csrc/        hand-written sm_100a CUDA (tcgen05/TMEM/TMA) + the C-ABI (include/potus_b200.h)
  cabi.py      ctypes mirror of the C-ABI (what the R .Call shim binds)
  model.py     host mirror of the reference's sampling boundary (cmdstanr `$sample()` /
               rstan::extract), final_2016.R:532-543
  datalist.py  the named data list of final_20{08,12,16}.R (the reference's host language, R, is
               not in this image)
  diagnostics.py  Stan-style ESS / R-hat for the benchmark metric
  postprocess.py  the reports' election-day summaries (state table, national vote, EV simulation, Brier)
  stancsv.py   CmdStan-format CSV writer (what rstan::read_stan_csv consumes, final_2016.R:543) + a reader for tests
  build.py     nvcc recipe
"""
from . import datalist  # noqa: F401
from .datalist import build_datalist, synthetic_datalist, load_npz, save_npz  # noqa: F401
from .model import cmdstan_model, PotusFit, logp_grad  # noqa: F401
from . import diagnostics  # noqa: F401
from . import postprocess  # noqa: F401
from . import stancsv  # noqa: F401
