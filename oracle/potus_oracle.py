"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import or execute it, and only as the checker / reported baseline.

fp64 restatement of the reference's log-posterior for
    /root/reference/scripts/model/poll_model_2020.stan            (full model)
    /root/reference/scripts/model/poll_model_2020_no_mode_adjustment.stan

PARITY PIN STATUS: the reference has no unit tests or golden vectors for this path and neither
R nor Stan exists in this image, so the arithmetic cannot be checked against the reference's
own binary.  What pins this oracle:
  (1) `logp_literal` is a line-by-line transcription of the .stan file (loops as written);
      `logp_grad_closed` is an independent closed form (hand-derived gradient); the two and a
      central finite difference agree to 1e-10 relative (tests/test_oracle.py).
  (2) the known-answer values in SURVEY.md section 8(c) (lp at theta=0 and at
      theta_i = 0.1 sin(1+0.37 i) for 2016/2012/2008) are reproduced to 1e-6.
  (3) end to end: the C oracle sampler (oracle/potus_oracle.c, same formulas) reproduces the
      reference's published per-state tables (README.md:83-136,179-232,279-332) within the
      README-vs-v4-report run-to-run spread -- see tests/golden/README.md.
Below the end-to-end tables parity is therefore "unpinned" in the task's sense and says so.

Unconstrained parameter order (Stan declaration order, matrices column-major,
poll_model_2020.stan:56-69):
  raw_mu_b_T[S] | raw_mu_b[S,T] | raw_mu_c[P] | raw_mu_m[M] | raw_mu_pop[Pop] | mu_e_bias(u) |
  rho_e_bias(u) | raw_e_bias[T] | raw_measure_noise_national[Nn] | raw_measure_noise_state[Ns] |
  raw_polling_bias[S]
The no-mode variant drops raw_mu_m, raw_mu_pop, mu_e_bias, rho_e_bias, raw_e_bias.
lp drops every additive constant Stan's `~` drops, and also the constant log(0.02) Jacobian of
the offset/multiplier transform.
"""
from __future__ import annotations

import numpy as np


def is_full_model(data: dict) -> bool:
    return "poll_mode_state" in data


def block_layout(data: dict):
    """[(name, size)] in unconstrained order, and total dimension D."""
    S, T, P = int(data["S"]), int(data["T"]), int(data["P"])
    Nn, Ns = int(data["N_national_polls"]), int(data["N_state_polls"])
    blocks = [("raw_mu_b_T", S), ("raw_mu_b", S * T), ("raw_mu_c", P)]
    if is_full_model(data):
        blocks += [("raw_mu_m", int(data["M"])), ("raw_mu_pop", int(data["Pop"])), ("mu_e_bias", 1),
                   ("rho_e_bias", 1), ("raw_e_bias", T)]
    blocks += [("raw_measure_noise_national", Nn), ("raw_measure_noise_state", Ns), ("raw_polling_bias", S)]
    return blocks, sum(n for _, n in blocks)


def split(theta, data):
    blocks, D = block_layout(data)
    assert theta.shape[-1] == D, (theta.shape, D)
    out, o = {}, 0
    for name, n in blocks:
        out[name] = theta[..., o:o + n]
        o += n
    return out


def transformed_data(data: dict):
    """poll_model_2020.stan:42-55.  All three Cholesky factors are scalar multiples of chol(Sigma0)."""
    w = np.asarray(data["state_weights"], float)
    cov0 = np.asarray(data["state_covariance_0"], float)
    nat_sd = float(np.sqrt(w @ cov0 @ w))
    L0 = np.linalg.cholesky(cov0)
    return dict(nat_sd=nat_sd, L0=L0,
                a_b=float(data["polling_bias_scale"]) / nat_sd,
                a_T=float(data["mu_b_T_scale"]) / nat_sd,
                a_w=float(data["random_walk_scale"]) / nat_sd)


# ----------------------------------------------------------------------------------------------
# (1) literal transcription in torch fp64 (autograd gives an independent gradient)
# ----------------------------------------------------------------------------------------------
def logp_literal(theta, data):
    """Line-by-line poll_model_2020.stan:70-131 in torch float64.  `theta` is a 1-D torch tensor
    (requires_grad allowed).  Returns the scalar lp."""
    import torch
    dt = torch.float64
    full = is_full_model(data)
    S, T = int(data["S"]), int(data["T"])
    td = transformed_data(data)
    tt = lambda a: torch.tensor(np.array(a, dtype=float), dtype=dt)
    w = tt(data["state_weights"])
    # :48-54  three scaled covariances, three Cholesky decompositions (done literally)
    cov0 = tt(data["state_covariance_0"])
    chol = lambda scale: torch.linalg.cholesky(cov0 * (float(scale) / td["nat_sd"]) ** 2)
    L_pb, L_T, L_w = chol(data["polling_bias_scale"]), chol(data["mu_b_T_scale"]), chol(data["random_walk_scale"])
    par = split(theta, data)
    raw_mu_b = par["raw_mu_b"].reshape(T, S).T  # column-major S x T
    # :77-79
    polling_bias = L_pb @ par["raw_polling_bias"]
    nat_pb = polling_bias @ w
    # :85-86
    cols = [None] * T
    cols[T - 1] = L_T @ par["raw_mu_b_T"] + tt(data["mu_b_prior"])
    for i in range(1, T):
        cols[T - 1 - i] = L_w @ raw_mu_b[:, T - 1 - i] + cols[T - i]
    mu_b = torch.stack(cols, dim=1)
    nat_avg = mu_b.T @ w  # :87
    mu_c = par["raw_mu_c"] * float(data["sigma_c"])  # :88
    lp = torch.zeros((), dtype=dt)
    if full:
        mu_m = par["raw_mu_m"] * float(data["sigma_m"])
        mu_pop = par["raw_mu_pop"] * float(data["sigma_pop"])
        mu_e = 0.02 * par["mu_e_bias"][0]  # offset 0, multiplier 0.02 (:61)
        rho = torch.sigmoid(par["rho_e_bias"][0])  # <lower=0, upper=1> (:62)
        lp = lp + torch.log(rho) + torch.log1p(-rho)  # log-Jacobian of lub_constrain
        sig_e = float(data["sigma_e_bias"])
        e = [par["raw_e_bias"][0] * sig_e]  # :91
        sigma_rho = torch.sqrt(1 - rho * rho) * sig_e  # :92
        for t in range(1, T):
            e.append(mu_e + rho * (e[t - 1] - mu_e) + par["raw_e_bias"][t] * sigma_rho)  # :93
        e_bias = torch.stack(e)
    st = np.asarray(data["state"]) - 1
    ds = np.asarray(data["day_state"]) - 1
    dn = np.asarray(data["day_national"]) - 1
    ps = np.asarray(data["poll_state"]) - 1
    pn = np.asarray(data["poll_national"]) - 1
    # :95-112
    eta_s = (mu_b[st, ds] + mu_c[ps] + par["raw_measure_noise_state"] * float(data["sigma_measure_noise_state"])
             + polling_bias[st])
    eta_n = (nat_avg[dn] + mu_c[pn] + par["raw_measure_noise_national"] * float(data["sigma_measure_noise_national"])
             + nat_pb)
    if full:
        eta_s = eta_s + mu_m[np.asarray(data["poll_mode_state"]) - 1] + mu_pop[np.asarray(data["poll_pop_state"]) - 1] \
            + tt(data["unadjusted_state"]) * e_bias[ds]
        eta_n = eta_n + mu_m[np.asarray(data["poll_mode_national"]) - 1] + mu_pop[np.asarray(data["poll_pop_national"]) - 1] \
            + tt(data["unadjusted_national"]) * e_bias[dn]
    # :117-128 priors (constants dropped)
    for name in ("raw_mu_b_T", "raw_mu_b", "raw_mu_c", "raw_measure_noise_national", "raw_measure_noise_state",
                 "raw_polling_bias"):
        lp = lp - 0.5 * (par[name] ** 2).sum()
    if full:
        lp = lp - 0.5 * (par["raw_mu_m"] ** 2).sum() - 0.5 * (par["raw_mu_pop"] ** 2).sum() \
            - 0.5 * (par["raw_e_bias"] ** 2).sum()
        lp = lp - 0.5 * (mu_e / 0.02) ** 2 - 0.5 * ((rho - 0.7) / 0.1) ** 2
    # :130-131 binomial_logit (binomial coefficient dropped)
    sp = torch.nn.functional.softplus
    ys, ns_ = tt(data["n_democrat_state"]), tt(data["n_two_share_state"])
    yn, nn_ = tt(data["n_democrat_national"]), tt(data["n_two_share_national"])
    lp = lp + (ys * eta_s - ns_ * sp(eta_s)).sum() + (yn * eta_n - nn_ * sp(eta_n)).sum()
    return lp


def logp_grad_literal(theta: np.ndarray, data: dict):
    import torch
    th = torch.tensor(np.asarray(theta, float), dtype=torch.float64, requires_grad=True)
    lp = logp_literal(th, data)
    (g,) = torch.autograd.grad(lp, th)
    return float(lp.detach()), g.numpy()


# ----------------------------------------------------------------------------------------------
# (2) closed form: scans + two GEMMs, hand-derived gradient
# ----------------------------------------------------------------------------------------------
def softplus(x):
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    return 0.5 * (1 + np.tanh(0.5 * x))


def forward_closed(theta: np.ndarray, data: dict, td=None):
    """Transformed parameters (poll_model_2020.stan:70-113) in collapsed form."""
    td = td or transformed_data(data)
    full = is_full_model(data)
    S, T = int(data["S"]), int(data["T"])
    L0, a_b, a_T, a_w = td["L0"], td["a_b"], td["a_T"], td["a_w"]
    w = np.asarray(data["state_weights"], float)
    par = split(np.asarray(theta, float), data)
    Z = par["raw_mu_b"].reshape(T, S).T
    # W[:,t] = a_T zT + a_w sum_{u=t}^{T-2} Z[:,u];  mu_b = prior 1^T + L0 W
    W = np.empty((S, T))
    W[:, T - 1] = 0.0
    W[:, :T - 1] = a_w * np.cumsum(Z[:, T - 2::-1], axis=1)[:, ::-1]
    W += a_T * par["raw_mu_b_T"][:, None]
    mu_b = np.asarray(data["mu_b_prior"], float)[:, None] + L0 @ W
    out = dict(mu_b=mu_b, W=W, Z=Z, par=par)
    out["polling_bias"] = a_b * (L0 @ par["raw_polling_bias"])
    out["nat_pb"] = out["polling_bias"] @ w
    out["nat_avg"] = w @ mu_b
    out["mu_c"] = par["raw_mu_c"] * float(data["sigma_c"])
    if full:
        out["mu_m"] = par["raw_mu_m"] * float(data["sigma_m"])
        out["mu_pop"] = par["raw_mu_pop"] * float(data["sigma_pop"])
        mu_e = 0.02 * par["mu_e_bias"][0]
        rho = sigmoid(par["rho_e_bias"][0])
        sig_e = float(data["sigma_e_bias"])
        sig_rho = np.sqrt(1 - rho * rho) * sig_e
        e = np.empty(T)
        e[0] = par["raw_e_bias"][0] * sig_e
        for t in range(1, T):
            e[t] = mu_e + rho * (e[t - 1] - mu_e) + par["raw_e_bias"][t] * sig_rho
        out.update(mu_e=mu_e, rho=rho, sigma_rho=sig_rho, e_bias=e)
    return out


def logp_grad_closed(theta: np.ndarray, data: dict, td=None):
    td = td or transformed_data(data)
    full = is_full_model(data)
    S, T = int(data["S"]), int(data["T"])
    L0, a_b, a_T, a_w = td["L0"], td["a_b"], td["a_T"], td["a_w"]
    w = np.asarray(data["state_weights"], float)
    f = forward_closed(theta, data, td)
    par = f["par"]
    st = np.asarray(data["state"]) - 1
    ds = np.asarray(data["day_state"]) - 1
    dn = np.asarray(data["day_national"]) - 1
    ps = np.asarray(data["poll_state"]) - 1
    pn = np.asarray(data["poll_national"]) - 1
    sig_s, sig_n = float(data["sigma_measure_noise_state"]), float(data["sigma_measure_noise_national"])
    eta_s = f["mu_b"][st, ds] + f["mu_c"][ps] + sig_s * par["raw_measure_noise_state"] + f["polling_bias"][st]
    eta_n = f["nat_avg"][dn] + f["mu_c"][pn] + sig_n * par["raw_measure_noise_national"] + f["nat_pb"]
    if full:
        ms, mn = np.asarray(data["poll_mode_state"]) - 1, np.asarray(data["poll_mode_national"]) - 1
        os_, on = np.asarray(data["poll_pop_state"]) - 1, np.asarray(data["poll_pop_national"]) - 1
        us, un = np.asarray(data["unadjusted_state"], float), np.asarray(data["unadjusted_national"], float)
        eta_s = eta_s + f["mu_m"][ms] + f["mu_pop"][os_] + us * f["e_bias"][ds]
        eta_n = eta_n + f["mu_m"][mn] + f["mu_pop"][on] + un * f["e_bias"][dn]
    ys, ns_ = np.asarray(data["n_democrat_state"], float), np.asarray(data["n_two_share_state"], float)
    yn, nn_ = np.asarray(data["n_democrat_national"], float), np.asarray(data["n_two_share_national"], float)
    lp = -0.5 * float(np.sum(np.asarray(theta, float) ** 2))
    if full:
        u_mu, u_rho = par["mu_e_bias"][0], par["rho_e_bias"][0]
        rho = f["rho"]
        lp += 0.5 * u_rho ** 2  # rho's unconstrained value has no N(0,1) prior: undo, add the real terms
        lp += -0.5 * ((rho - 0.7) / 0.1) ** 2 + np.log(rho) + np.log1p(-rho)
    lp += float(np.sum(ys * eta_s - ns_ * softplus(eta_s)) + np.sum(yn * eta_n - nn_ * softplus(eta_n)))

    # ---- gradient
    r_s = ys - ns_ * sigmoid(eta_s)
    r_n = yn - nn_ * sigmoid(eta_n)
    g = {k: -v.copy() for k, v in par.items()}
    g["raw_measure_noise_state"] += sig_s * r_s
    g["raw_measure_noise_national"] += sig_n * r_n
    G = np.zeros((S, T))
    np.add.at(G, (st, ds), r_s)
    rn_day = np.zeros(T)
    np.add.at(rn_day, dn, r_n)
    G += w[:, None] * rn_day[None, :]
    H = L0.T @ G
    cH = np.cumsum(H, axis=1)
    gZ = -f["Z"].copy()
    gZ[:, :T - 1] += a_w * cH[:, :T - 1]
    g["raw_mu_b"] = gZ.T.reshape(-1)
    g["raw_mu_b_T"] += a_T * cH[:, T - 1]
    g_pb = np.zeros(S)
    np.add.at(g_pb, st, r_s)
    g_pb += w * r_n.sum()
    g["raw_polling_bias"] += a_b * (L0.T @ g_pb)
    gc = np.zeros(int(data["P"]))
    np.add.at(gc, ps, r_s)
    np.add.at(gc, pn, r_n)
    g["raw_mu_c"] += float(data["sigma_c"]) * gc
    if full:
        gm = np.zeros(int(data["M"])); np.add.at(gm, ms, r_s); np.add.at(gm, mn, r_n)
        gp = np.zeros(int(data["Pop"])); np.add.at(gp, os_, r_s); np.add.at(gp, on, r_n)
        g["raw_mu_m"] += float(data["sigma_m"]) * gm
        g["raw_mu_pop"] += float(data["sigma_pop"]) * gp
        g_e = np.zeros(T)
        np.add.at(g_e, ds, us * r_s)
        np.add.at(g_e, dn, un * r_n)
        ebar = np.empty(T)
        ebar[T - 1] = g_e[T - 1]
        for t in range(T - 2, -1, -1):
            ebar[t] = g_e[t] + rho * ebar[t + 1]
        sig_e, sig_rho, mu_e, e = float(data["sigma_e_bias"]), f["sigma_rho"], f["mu_e"], f["e_bias"]
        ze = par["raw_e_bias"]
        gze = -ze.copy()
        gze[0] += sig_e * ebar[0]
        gze[1:] += sig_rho * ebar[1:]
        g["raw_e_bias"] = gze
        d_mu_e = (1 - rho) * ebar[1:].sum()
        g["mu_e_bias"] = np.array([0.02 * d_mu_e - u_mu])
        d_rho = np.sum(ebar[1:] * ((e[:-1] - mu_e) - ze[1:] * sig_e * rho / np.sqrt(1 - rho * rho))) \
            - (rho - 0.7) / 0.01
        g["rho_e_bias"] = np.array([rho * (1 - rho) * d_rho + (1 - 2 * rho)])
    blocks, _ = block_layout(data)
    grad = np.concatenate([np.atleast_1d(g[name]) for name, _ in blocks])
    return lp, grad


def logp_grad_fd(theta, data, idx, h=1e-5):
    """Central finite differences of logp (closed form) on the coordinates `idx`."""
    td = transformed_data(data)
    out = np.empty(len(idx))
    for k, i in enumerate(idx):
        tp = np.array(theta, float); tp[i] += h
        tm = np.array(theta, float); tm[i] -= h
        out[k] = (logp_grad_closed(tp, data, td)[0] - logp_grad_closed(tm, data, td)[0]) / (2 * h)
    return out


def constrained_draw(theta: np.ndarray, data: dict, td=None) -> dict:
    """Transformed parameters + generated quantities that the reference's consumers extract
    (final_2016.R:553-708): mu_b[S,T], mu_c, mu_m, mu_pop, e_bias, polling_bias,
    predicted_score[T,S] = inv_logit(mu_b)' (poll_model_2020.stan:134-140)."""
    f = forward_closed(theta, data, td)
    out = dict(mu_b=f["mu_b"], mu_c=f["mu_c"], polling_bias=f["polling_bias"],
               predicted_score=sigmoid(f["mu_b"]).T)
    if is_full_model(data):
        out.update(mu_m=f["mu_m"], mu_pop=f["mu_pop"], e_bias=f["e_bias"])
    return out
