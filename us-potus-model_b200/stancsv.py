"""CmdStan-format CSV output.

The reference never touches the sampler's memory: `fit$output_files()` are CmdStan CSVs that
`rstan::read_stan_csv` parses back (final_2016.R:543, final_2012.R:569, final_2008.R:573).  Writing the same
files makes the GPU sampler a zero-edit replacement for everything downstream of that line.

Column order is Stan's: 7 sampler diagnostics, then `parameters` (poll_model_2020.stan:56-69, constrained
values), `transformed parameters` (:70-113, declaration order) and `generated quantities` (:134-140);
containers are flattened column-major with 1-based `name.i.j` labels.  The device returns, per kept draw,
the unconstrained vector and the consumed transformed parameters (mu_b, mu_c, mu_m, mu_pop, e_bias,
polling_bias); the remaining CSV columns (national_mu_b_average, national_polling_bias_average,
sigma_rho, logit_pi_democrat_*, predicted_score) are one-line functions of those and are formed here while
formatting -- nothing on this path feeds back into sampling.

`read_stan_csv` is the Python stand-in for rstan::read_stan_csv + rstan::extract used by the tests.
"""
from __future__ import annotations

import os
import re

import numpy as np

SAMPLER_COLS = ("lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__")


def _is_full(data: dict) -> bool:
    return "poll_mode_state" in data


def _dims(data: dict):
    S, T, P = int(data["S"]), int(data["T"]), int(data["P"])
    Nn, Ns = int(data["N_national_polls"]), int(data["N_state_polls"])
    full = _is_full(data)
    par = [("raw_mu_b_T", (S,)), ("raw_mu_b", (S, T)), ("raw_mu_c", (P,))]
    if full:
        par += [("raw_mu_m", (int(data["M"]),)), ("raw_mu_pop", (int(data["Pop"]),)), ("mu_e_bias", ()),
                ("rho_e_bias", ()), ("raw_e_bias", (T,))]
    par += [("raw_measure_noise_national", (Nn,)), ("raw_measure_noise_state", (Ns,)), ("raw_polling_bias", (S,))]
    tp = [("mu_b", (S, T)), ("mu_c", (P,))]
    if full:
        tp += [("mu_m", (int(data["M"]),)), ("mu_pop", (int(data["Pop"]),)), ("e_bias", (T,))]
    tp += [("polling_bias", (S,)), ("national_mu_b_average", (T,)), ("national_polling_bias_average", ())]
    if full:
        tp += [("sigma_rho", ())]
    tp += [("logit_pi_democrat_state", (Ns,)), ("logit_pi_democrat_national", (Nn,))]
    gq = [("predicted_score", (T, S))]
    return par, tp, gq


def _labels(name: str, dims: tuple) -> list[str]:
    if not dims:
        return [name]
    if len(dims) == 1:
        return [f"{name}.{i + 1}" for i in range(dims[0])]
    return [f"{name}.{i + 1}.{j + 1}" for j in range(dims[1]) for i in range(dims[0])]  # column-major


def column_names(data: dict) -> list[str]:
    par, tp, gq = _dims(data)
    out = list(SAMPLER_COLS)
    for nm, d in par + tp + gq:
        out += _labels(nm, d)
    return out


def draw_table(data: dict, theta: np.ndarray, tps: dict) -> np.ndarray:
    """[n_draws, n_columns - 7]: every parameter / transformed parameter / generated quantity of the CSV.

    theta [n, D] unconstrained (Stan order); tps: mu_b [n,S,T], mu_c [n,P], polling_bias [n,S] and, for
    the full model, mu_m, mu_pop, e_bias -- the arrays PotusFit.extract returns."""
    theta = np.atleast_2d(np.asarray(theta, float))
    n = theta.shape[0]
    full = _is_full(data)
    par, tp, gq = _dims(data)
    w = np.asarray(data["state_weights"], float)
    cols, o = [], 0
    raw = {}
    for nm, d in par:
        k = int(np.prod(d)) if d else 1
        blk = theta[:, o:o + k]
        o += k
        if nm == "mu_e_bias":       # real<offset=0, multiplier=0.02>
            blk = 0.02 * blk
        elif nm == "rho_e_bias":    # real<lower=0, upper=1>
            blk = 0.5 * (1.0 + np.tanh(0.5 * blk))
        raw[nm] = blk
        cols.append(blk)
    if o != theta.shape[1]:
        raise ValueError(f"theta has {theta.shape[1]} columns, the model has {o} parameters")
    mu_b = np.asarray(tps["mu_b"], float).reshape(n, int(data["S"]), int(data["T"]))
    mu_c = np.asarray(tps["mu_c"], float).reshape(n, -1)
    pb = np.asarray(tps["polling_bias"], float).reshape(n, -1)
    nat_avg = np.einsum("nst,s->nt", mu_b, w)
    nat_pb = pb @ w
    i0 = lambda key: np.asarray(data[key]).astype(np.int64) - 1
    st, ds, dn = i0("state"), i0("day_state"), i0("day_national")
    eta_s = mu_b[:, st, ds] + mu_c[:, i0("poll_state")] + pb[:, st] + \
        raw["raw_measure_noise_state"] * float(data["sigma_measure_noise_state"])
    eta_n = nat_avg[:, dn] + mu_c[:, i0("poll_national")] + nat_pb[:, None] + \
        raw["raw_measure_noise_national"] * float(data["sigma_measure_noise_national"])
    if full:
        mu_m = np.asarray(tps["mu_m"], float).reshape(n, -1)
        mu_pop = np.asarray(tps["mu_pop"], float).reshape(n, -1)
        e = np.asarray(tps["e_bias"], float).reshape(n, -1)
        eta_s = eta_s + mu_m[:, i0("poll_mode_state")] + mu_pop[:, i0("poll_pop_state")] + \
            np.asarray(data["unadjusted_state"], float)[None, :] * e[:, ds]
        eta_n = eta_n + mu_m[:, i0("poll_mode_national")] + mu_pop[:, i0("poll_pop_national")] + \
            np.asarray(data["unadjusted_national"], float)[None, :] * e[:, dn]
        rho = raw["rho_e_bias"][:, 0]
        sigma_rho = np.sqrt(1.0 - rho * rho) * float(data["sigma_e_bias"])
    have = dict(mu_b=mu_b.transpose(0, 2, 1).reshape(n, -1),   # column-major [S,T]: s fastest
                mu_c=mu_c, polling_bias=pb, national_mu_b_average=nat_avg,
                national_polling_bias_average=nat_pb[:, None],
                logit_pi_democrat_state=eta_s, logit_pi_democrat_national=eta_n,
                predicted_score=(0.5 * (1.0 + np.tanh(0.5 * mu_b))).reshape(n, -1))  # [T,S] col-major: t fastest
    if full:
        have.update(mu_m=mu_m, mu_pop=mu_pop, e_bias=e, sigma_rho=sigma_rho[:, None])
    for nm, _ in tp + gq:
        cols.append(have[nm])
    return np.concatenate(cols, axis=1)


def _header(model_name, chain_id, seed, num_samples, num_warmup, thin, adapt_delta, max_depth, init, note):
    lines = [
        "stan_version_major = 2", "stan_version_minor = 24", "stan_version_patch = 1",
        f"model = {model_name}_model", "method = sample (Default)", "  sample",
        f"    num_samples = {num_samples}", f"    num_warmup = {num_warmup}", "    save_warmup = 0 (Default)",
        f"    thin = {thin}", "    adapt", "      engaged = 1 (Default)",
        "      gamma = 0.050000000000000003 (Default)", f"      delta = {adapt_delta:.17g}",
        "      kappa = 0.75 (Default)", "      t0 = 10 (Default)", "      init_buffer = 75 (Default)",
        "      term_buffer = 50 (Default)", "      window = 25 (Default)", "    algorithm = hmc (Default)",
        "      hmc", "        engine = nuts (Default)", "          nuts", f"            max_depth = {max_depth}",
        "        metric = diag_e (Default)", "        metric_file =  (Default)", "        stepsize = 1 (Default)",
        "        stepsize_jitter = 0 (Default)", f"id = {chain_id}", "data", "  file = (in-memory data list)",
        f"init = {init:g}", "random", f"  seed = {seed}", "output", "  file = output.csv (Default)",
        "  diagnostic_file =  (Default)", "  refresh = 100 (Default)",
    ]
    return "".join(f"# {ln}\n" for ln in lines) + "".join(f"# {ln}\n" for ln in note)


def write_chain_csv(path, names, rows, *, model_name, chain_id, seed, num_samples, num_warmup, thin, stepsize,
                    inv_metric, adapt_delta=0.8, max_depth=10, init=2.0, elapsed=(0.0, 0.0), sig_figs=6, note=()):
    """One CmdStan output file: config comments, header, adaptation block, draws, timing trailer."""
    rows = np.atleast_2d(np.asarray(rows, float))
    if rows.shape[1] != len(names):
        raise ValueError(f"{rows.shape[1]} columns for {len(names)} names")
    with open(path, "w") as f:
        f.write(_header(model_name, chain_id, seed, num_samples, num_warmup, thin, adapt_delta, max_depth, init, note))
        f.write(",".join(names) + "\n")
        f.write("# Adaptation terminated\n")
        f.write(f"# Step size = {stepsize:.6g}\n")
        f.write("# Diagonal elements of inverse mass matrix:\n")
        f.write("# " + ", ".join(f"{v:.6g}" for v in np.asarray(inv_metric, float)) + "\n")
        np.savetxt(f, rows, fmt=f"%.{sig_figs}g", delimiter=",")
        f.write("# \n")
        f.write(f"#  Elapsed Time: {elapsed[0]:g} seconds (Warm-up)\n")
        f.write(f"#                {elapsed[1]:g} seconds (Sampling)\n")
        f.write(f"#                {elapsed[0] + elapsed[1]:g} seconds (Total)\n")
        f.write("# \n")
    return path


def write_stan_csv(fit, directory: str, basename: str | None = None, chains=None, sig_figs: int = 6) -> list[str]:
    """`fit$output_files()` analogue: one CmdStan CSV per chain holding that chain's kept draws.

    Each line is ~43 000 numbers on the 2016 list (as in the reference), so `chains` lets a caller write a
    subset.  Kept draws are the device's thinned draws (PotusConfig.keep_per_chain; 0 = every iteration):
    the header reports thin = iter_sampling // keep and num_samples = keep * thin so that
    1 + (num_samples - 1) %/% thin rows, what rstan expects, is what the file holds."""
    os.makedirs(directory, exist_ok=True)
    cfg, data = fit.cfg, fit.data
    C = int(cfg.chains)
    keep = fit.n_draws // C
    if keep < 1:
        raise ValueError("the fit holds no kept draws")
    thin = max(1, int(cfg.iter_sampling) // keep)
    names = column_names(data)
    sel = range(C) if chains is None else [int(c) for c in chains]
    sp = fit.sampler_params(inc_warmup=False)
    inv_metric = fit.inv_metric()
    theta = fit.theta()
    tp_names = ["mu_b", "mu_c", "polling_bias"] + (["mu_m", "mu_pop", "e_bias"] if _is_full(data) else [])
    tps = fit.extract(tp_names)
    it_kept = np.arange(keep) * thin                       # sampling iterations the device kept: 0, thin, 2 thin, ... as CmdStan's `thin`
    base = basename or fit.model_name
    el = (float(fit.stats.get("seconds_warmup", 0.0)), float(fit.stats.get("seconds_sampling", 0.0)))
    paths = []
    for c in sel:
        r = slice(c * keep, (c + 1) * keep)
        body = draw_table(data, theta[r], {k: v[r] for k, v in tps.items()})
        diag = np.stack([sp[k][c, it_kept] for k in SAMPLER_COLS], axis=1)
        if _is_full(data):   # Stan's lp__ carries the constant log-Jacobian log(0.02) of mu_e_bias' multiplier; the library drops constants
            diag[:, SAMPLER_COLS.index("lp__")] += np.log(0.02)
        note = (f"sampler = potus_b200 (sm_100a resident NUTS, fp32 state); iter_sampling = {int(cfg.iter_sampling)}, "
                f"kept iterations = {thin}k (k = 0, 1, ...)",)
        p = os.path.join(directory, f"{base}-{c + 1}.csv")
        write_chain_csv(p, names, np.concatenate([diag, body], axis=1), model_name=fit.model_name,
                        chain_id=c + 1 + int(getattr(cfg, "chain_id_offset", 0)), seed=int(cfg.seed),
                        num_samples=keep * thin, num_warmup=int(cfg.iter_warmup), thin=thin,
                        stepsize=float(diag[-1, 2]), inv_metric=inv_metric[c], adapt_delta=float(cfg.adapt_delta),
                        max_depth=int(cfg.max_treedepth), init=float(cfg.init_radius), elapsed=el, sig_figs=sig_figs,
                        note=note)
        paths.append(p)
    return paths


_KV = re.compile(r"^#\s*([A-Za-z_]+)\s*=\s*(\S*)")


def read_stan_csv(paths) -> dict:
    """Parse CmdStan CSVs the way rstan::read_stan_csv + extract(permuted = FALSE) see them.

    Returns {"names", "config" (first file's key = value comments), "stepsize" [chains], "inv_metric"
    [chains][D], "draws" {par: array [iterations, chains, dims...]}, "sampler_params" {name: [iter, chains]}}."""
    if isinstance(paths, (str, os.PathLike)):
        paths = [paths]
    mats, steps, metrics, names, config = [], [], [], None, {}
    for fi, p in enumerate(paths):
        rows, hdr, want_metric = [], None, False
        with open(p) as f:
            for ln in f:
                if ln.startswith("#"):
                    if want_metric:
                        metrics.append(np.array([float(x) for x in ln[1:].split(",")]))
                        want_metric = False
                        continue
                    if ln.startswith("# Step size"):
                        steps.append(float(ln.split("=")[1]))
                    elif ln.startswith("# Diagonal elements"):
                        want_metric = True
                    elif fi == 0:
                        m = _KV.match(ln)
                        if m:
                            config[m.group(1)] = m.group(2)
                    continue
                if hdr is None:
                    hdr = ln.strip().split(",")
                    continue
                rows.append(np.array(ln.split(","), dtype=float))
        if names is None:
            names = hdr
        elif names != hdr:
            raise ValueError(f"{p}: header differs from the first file's")
        mats.append(np.stack(rows) if rows else np.empty((0, len(hdr))))
    arr = np.stack(mats, axis=1)                       # [iter, chain, column]
    groups: dict = {}
    for j, nm in enumerate(names):
        base, *idx = nm.split(".")
        groups.setdefault(base, []).append((tuple(int(i) for i in idx), j))
    draws, sp = {}, {}
    for base, items in groups.items():
        cols = [j for _, j in items]
        if base in SAMPLER_COLS:
            sp[base] = arr[:, :, cols[0]]
            continue
        if not items[0][0]:
            draws[base] = arr[:, :, cols[0]]
            continue
        dims = tuple(max(ix[k] for ix, _ in items) for k in range(len(items[0][0])))
        out = np.empty(arr.shape[:2] + dims)
        for ix, j in items:
            out[(slice(None), slice(None)) + tuple(i - 1 for i in ix)] = arr[:, :, j]
        draws[base] = out
    return dict(names=names, config=config, stepsize=np.array(steps), inv_metric=metrics, draws=draws,
                sampler_params=sp)
