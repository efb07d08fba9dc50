"""Host-side mirror of the reference's sampling boundary.

Reference call site (scripts/model/final_2016.R:532-543):

    model <- cmdstanr::cmdstan_model("scripts/model/poll_model_2020.stan", compile=TRUE, force=TRUE)
    fit   <- model$sample(data = data, seed = 1843, parallel_chains = n_cores, chains = n_chains,
                          iter_warmup = n_warmup, iter_sampling = n_sampling, refresh = n_refresh)
    out   <- rstan::read_stan_csv(fit$output_files())
    ...   rstan::extract(out, pars = "mu_b")[[1]][,,254]          (:556)

Here (Python stands in for R, which is not in this image; the R shim in r/ has the same shape):

    model = cmdstan_model("poll_model_2020.stan")
    fit   = model.sample(data=data, seed=1843, chains=1024, iter_warmup=500, iter_sampling=500)
    fit.extract("mu_b")[:, :, 253]

Everything numerical happens behind the C-ABI of include/potus_b200.h in the CUDA library; this
file only marshals the named list, forwards the cmdstanr argument names and reshapes outputs the
way rstan::extract does ([draws, dims...], draw index first).  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import cabi

_PAR_DIMS = {
    "mu_b": lambda d: (int(d["S"]), int(d["T"])),
    "mu_c": lambda d: (int(d["P"]),),
    "mu_m": lambda d: (int(d["M"]),),
    "mu_pop": lambda d: (int(d["Pop"]),),
    "polling_bias": lambda d: (int(d["S"]),),
    "e_bias": lambda d: (int(d["T"]),),
    "predicted_score": lambda d: (int(d["T"]), int(d["S"])),
}
_FULL_ONLY = ("mu_m", "mu_pop", "e_bias")
SAMPLER_PARAMS = ("lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__")


def _variant_of(stan_file: str | None) -> str | None:
    if stan_file is None:
        return None
    base = os.path.basename(stan_file)
    if "no_mode_adjustment" in base:
        return "no_mode"
    if "poll_model_2020" in base:
        return "full"
    raise ValueError(f"unknown model '{stan_file}': this library implements poll_model_2020.stan and "
                     "poll_model_2020_no_mode_adjustment.stan only")


class PotusFit:
    """What `out` is to the reference's consumers (final_2016.R:553-708, README.Rmd:206)."""

    def __init__(self, lib, handle, data, cfg, variant):
        self._lib, self._h, self.data, self.cfg, self.variant = lib, handle, data, cfg, variant
        self.model_name = "poll_model_2020" if variant == "full" else "poll_model_2020_no_mode_adjustment"
        st = cabi.PotusStats()
        cabi.check(lib, lib.potus_get_stats(handle, C.byref(st)))
        self.stats = {k: getattr(st, k) for k, _ in cabi.PotusStats._fields_}

    def _get(self, name: str) -> np.ndarray:
        n = self._lib.potus_draws_size(self._h, name.encode())
        if n == 0:
            raise KeyError(f"'{name}' is not an extractable quantity of {self.model_name}")
        out = np.empty(n, dtype=np.float64)
        cabi.check(self._lib, self._lib.potus_get_draws(self._h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), n))
        return out

    @property
    def n_draws(self) -> int:
        return int(self.stats["n_draws_kept"])

    def extract(self, pars):
        """rstan::extract(out, pars=...)[[1]] analogue: array [draws, dims...] (chains concatenated
        in chain order; rstan additionally permutes draws, which no consumer depends on)."""
        single = isinstance(pars, str)
        names = [pars] if single else list(pars)
        res = {}
        for nm in names:
            if nm not in _PAR_DIMS:
                raise KeyError(nm)
            if nm in _FULL_ONLY and self.variant != "full":
                raise KeyError(f"{nm} does not exist in {self.model_name}")
            dims = _PAR_DIMS[nm](self.data)
            flat = self._get(nm)
            res[nm] = flat.reshape((self.n_draws,) + dims, order="F")
        return res[names[0]] if single else res

    def theta(self) -> np.ndarray:
        d = int(self.stats["n_params"])
        return self._get("theta").reshape((self.n_draws, d), order="F")

    def monitor(self) -> np.ndarray:
        """[chains, iter_sampling, S+1]: mu_b[,T] and national_mu_b_average[T] for EVERY sampling iteration."""
        S = int(self.data["S"])
        a = self._get("monitor").reshape((self.cfg.iter_sampling * self.cfg.chains, S + 1), order="F")
        return a.reshape(self.cfg.chains, self.cfg.iter_sampling, S + 1)

    def sampler_params(self, inc_warmup: bool = True) -> dict:
        n_it = self.cfg.iter_warmup + self.cfg.iter_sampling
        a = self._get("sampler_params").reshape((n_it * self.cfg.chains, 7), order="F").reshape(self.cfg.chains, n_it, 7)
        if not inc_warmup:
            a = a[:, self.cfg.iter_warmup:, :]
        return {k: a[:, :, i] for i, k in enumerate(SAMPLER_PARAMS)}

    def inv_metric(self) -> np.ndarray:
        """[chains, D] adapted diagonal inverse metric (CmdStan's "Diagonal elements of inverse mass matrix")."""
        d = int(self.stats["n_params"])
        return self._get("inv_metric").reshape((self.cfg.chains, d), order="F")

    def summary(self, ev=None, ev_threshold: float = 270.0, ess: bool = True) -> dict:
        """On-device post-processing over ALL sampling iterations (potus_postprocess): the election-day state table the
        reports print (README.Rmd:206-248: mean / sd / quantiles / P(win) per state, national vote), the electoral-college
        simulation (README.Rmd:271-300) when `ev` is given, and Stan ESS / split R-hat of the monitored scalars."""
        S = int(self.data["S"])
        tab = np.zeros((S + 2, 8))
        et = np.zeros((S + 1, 3)) if ess else None
        f64p = C.POINTER(C.c_double)
        evp = None
        if ev is not None:
            eva = np.ascontiguousarray(ev, dtype=np.float64)
            if eva.shape != (S,):
                raise ValueError("ev must have one entry per state")
            evp = eva.ctypes.data_as(f64p)
        cabi.check(self._lib, self._lib.potus_postprocess(self._h, evp, float(ev_threshold), tab.ctypes.data_as(f64p),
                                                          et.ctypes.data_as(f64p) if ess else None))
        cols = ("mean", "sd", "q025", "q05", "q50", "q95", "q975", "prob")
        out = {"states": {c: tab[:S, i].copy() for i, c in enumerate(cols)}, "national": {c: float(tab[S, i]) for i, c in enumerate(cols)}}
        if ev is not None:
            out["electoral_votes"] = {c: float(tab[S + 1, i]) for i, c in enumerate(cols)}
        if ess:
            out["ess"], out["rhat"], out["monitor_mean"] = et[:, 0].copy(), et[:, 1].copy(), et[:, 2].copy()
        return out

    def save_csvfiles(self, directory: str, basename: str | None = None, chains=None) -> list:
        """cmdstanr `fit$save_output_files()` / `fit$output_files()` analogue: CmdStan-format CSVs that
        rstan::read_stan_csv parses (final_2016.R:543).  See stancsv.py."""
        from . import stancsv
        return stancsv.write_stan_csv(self, directory, basename, chains)

    def device_buffer(self, which: int):
        """(device pointer, n_floats) of a raw fp32 output buffer; bench.py wraps it for the NCCL all-gather."""
        p, n = C.c_void_p(), C.c_size_t()
        cabi.check(self._lib, self._lib.potus_device_buffer(self._h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def close(self):
        if self._h:
            self._lib.potus_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CmdStanModelB200:
    def __init__(self, stan_file: str | None = None):
        self.stan_file = stan_file
        self.variant = _variant_of(stan_file)

    def sample(self, data: dict, seed: int = 1843, chains: int = 4, parallel_chains: int | None = None,
               iter_warmup: int = 500, iter_sampling: int = 500, refresh: int | None = None,
               adapt_delta: float = 0.8, max_treedepth: int = 10, keep_per_chain: int = 0, device: int = 0,
               chain_id_offset: int = 0, init: float = 2.0, force_stream: bool = False, n_gpus: int = 1, state: dict | None = None) -> PotusFit:
        """cmdstanr `$sample()` argument names; `parallel_chains`/`refresh` are accepted and ignored
        (all chains run concurrently on the GPU).  n_gpus > 1: the library itself shards `chains` over that many devices
        in this one process and all-gathers the kept draws with NCCL (include/potus_b200.h, PotusConfig.n_gpus)."""
        lib = cabi.load_library()
        has_mode = "poll_mode_state" in data
        variant = self.variant or ("full" if has_mode else "no_mode")
        if variant == "full" and not has_mode:
            raise ValueError("poll_model_2020.stan needs poll_mode_*/poll_pop_* in the data list")
        if variant == "no_mode" and has_mode:
            data = {k: v for k, v in data.items() if not (k.startswith("poll_mode_") or k.startswith("poll_pop_"))}
        pd, keep = cabi.marshal_data(data)
        cfg = cabi.make_config(chains=chains, iter_warmup=iter_warmup, iter_sampling=iter_sampling, seed=seed,
                               keep_per_chain=keep_per_chain, max_treedepth=max_treedepth, adapt_delta=adapt_delta,
                               init_radius=init, device=device, chain_id_offset=chain_id_offset, force_stream=force_stream, n_gpus=n_gpus)
        h = C.c_void_p()
        cabi.check(lib, lib.potus_create(C.byref(pd), C.byref(cfg), C.byref(h)))
        try:
            if state is not None:   # dict(theta=[chains, D], stepsize=[chains], inv_metric=[chains, D]); needs iter_warmup == 0
                th = np.ascontiguousarray(state["theta"], dtype=np.float64)
                ep = np.ascontiguousarray(state["stepsize"], dtype=np.float64)
                im = np.ascontiguousarray(state["inv_metric"], dtype=np.float64)
                if th.shape != im.shape or th.shape[0] != chains or ep.shape != (chains,):
                    raise ValueError("state: theta and inv_metric must be [chains, D], stepsize [chains]")
                f64p = C.POINTER(C.c_double)
                cabi.check(lib, lib.potus_set_state(h, th.ctypes.data_as(f64p), ep.ctypes.data_as(f64p), im.ctypes.data_as(f64p)))
            cabi.check(lib, lib.potus_run(h))
        except Exception:
            lib.potus_destroy(h)
            raise
        del keep
        return PotusFit(lib, h, data, cfg, variant)


def cmdstan_model(stan_file: str | None = None, compile: bool = True, force: bool = False) -> CmdStanModelB200:  # noqa: A002
    """cmdstanr::cmdstan_model analogue; `compile`/`force` are accepted for signature parity (the CUDA
    library is prebuilt by __graft_entry__.build())."""
    return CmdStanModelB200(stan_file)


def logp_grad(data: dict, theta: np.ndarray, force_stream: bool = False):
    """Test hook (potus_logp_grad): lp and gradient on the device for theta[n, D] (Stan order)."""
    lib = cabi.load_library()
    pd, keep = cabi.marshal_data(data)
    th = np.ascontiguousarray(np.atleast_2d(np.asarray(theta, dtype=np.float64)))
    n, D = th.shape
    Dlib = lib.potus_num_params(C.byref(pd))
    if Dlib != D:
        raise ValueError(f"theta has {D} columns, model has {Dlib} parameters")
    lp = np.empty(n)
    g = np.empty((n, D))
    f64p = C.POINTER(C.c_double)
    cabi.check(lib, lib.potus_logp_grad_ex(C.byref(pd), th.ctypes.data_as(f64p), n, lp.ctypes.data_as(f64p), g.ctypes.data_as(f64p),
                                           1 if force_stream else 0))
    del keep
    return lp, g
