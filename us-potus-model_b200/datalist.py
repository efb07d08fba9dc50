"""Named data list builder: a Python restatement of the R wrangling that
precedes the sampling boundary in the reference drivers.

Follows (file:line in /root/reference):
  scripts/model/final_2016.R:73-162   poll wrangling, indices
  scripts/model/final_2016.R:166-193  state context (weights) from 2012.csv
  scripts/model/final_2016.R:198-353  state covariance / scales
  scripts/model/final_2016.R:403-414  mu_b_prior from state_priors_08_12_16.csv
  scripts/model/final_2016.R:436-514  the `data` list handed to Stan
and the 2012/2008 analogues (final_2012.R, final_2008.R) which differ only in
the poll CSV parsing, the absence of mode/population indices and the weights
(no population-growth factor).

The host language of the reference is R, which is not installed in this image,
so this module exists to (a) generate the committed fixtures under
tests/golden/ (see tests/golden/make_fixtures.py) and (b) let bench.py build
its workload without R.  The product boundary (the C-ABI in include/) takes the
same named list whoever assembled it.

R-isms reproduced on purpose: `%/%` on day differences, str_extract's first
match of "[A-z0-9 ]+", factor level = sorted unique strings, arrange() +
distinct(.keep_all=TRUE) keeping the first row, round() half-to-even,
lqmm::make.positive.definite (eigenvalue floor), CR-only CSV files.
"""
from __future__ import annotations

import io
import os
import re
from dataclasses import dataclass

import numpy as np
import pandas as pd

ADJUSTERS = ("ABC", "Washington Post", "Ipsos", "Pew", "YouGov", "NBC")  # final_2016.R:423-430


def _read_csv_any_newline(path: str, **kw) -> pd.DataFrame:
    """data/2008.csv, 2012.csv and state_region_crosswalk.csv use bare-CR line ends."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw.startswith(b"\xef\xbb\xbf"):
        raw = raw[3:]
    text = raw.decode("utf-8", errors="replace").replace("\r\n", "\n").replace("\r", "\n")
    return pd.read_csv(io.StringIO(text), **kw)


def make_positive_definite(m: np.ndarray) -> np.ndarray:
    """lqmm::make.positive.definite (third-party, not in tree; restated from its
    published source): m + V diag(max(0, 2*tol - lambda)) V^T with
    tol = d * max|lambda| * eps.  Used at final_2016.R:286,295."""
    d = m.shape[0]
    lam, vec = np.linalg.eigh(m)
    tol = d * np.max(np.abs(lam)) * np.finfo(float).eps
    tau = np.maximum(0.0, 2.0 * tol - lam)
    return m + (vec * tau) @ vec.T


def cov_matrix(n: int, sigma2: float, rho: float) -> np.ndarray:
    """final_2016.R:29-35."""
    m = np.full((n, n), rho)
    np.fill_diagonal(m, 1.0)
    s = np.sqrt(sigma2) * np.eye(n)
    return s @ m @ s


def _r_round(x):
    """R's round(): IEC 60559 half-to-even, same as numpy.rint."""
    return np.rint(x)


@dataclass
class YearSpec:
    year: int
    run_date: str
    election_day: str
    start_date: str
    polls_csv: str
    dem: str
    rep: str
    context_csv: str
    context_rep_count: str
    growth: bool
    full_model: bool


SPECS = {
    2016: YearSpec(2016, "2016-11-08", "2016-11-08", "2016-03-01", "all_polls.csv", "clinton", "trump",
                   "2012.csv", "romney_count", True, True),
    2012: YearSpec(2012, "2012-11-06", "2012-11-06", "2012-03-01", "all_polls_2012.csv", "obama", "romney",
                   "2008.csv", "mccain_count", False, False),
    2008: YearSpec(2008, "2008-11-03", "2008-11-03", "2008-03-01", "all_polls_2008.csv", "obama", "mccain",
                   "2008.csv", "mccain_count", False, False),
}


def _wrangle_polls(spec: YearSpec, data_dir: str) -> pd.DataFrame:
    run_date = pd.Timestamp(spec.run_date)
    start_date = pd.Timestamp(spec.start_date)
    path = os.path.join(data_dir, spec.polls_csv)
    if spec.year == 2016:
        ap = pd.read_csv(path)  # final_2016.R:74
        ap = ap.rename(columns={"number.of.observations": "n", "start.date": "start", "end.date": "end_"})
        ap["begin"] = pd.to_datetime(ap["start"], format="%Y-%m-%d")
        ap["end"] = pd.to_datetime(ap["end_"], format="%Y-%m-%d")
        for c in ("johnson", "mcmullin"):
            ap[c] = ap[c].astype(float)
    else:
        ap = _read_csv_any_newline(path)  # final_2012.R:76 (read_csv; dates are m/d/yy)
        ap = ap.rename(columns={"number.of.observations": "n", "start.date": "start", "end.date": "end_"})
        ap["begin"] = pd.to_datetime(ap["start"], format="%m/%d/%y")
        ap["end"] = pd.to_datetime(ap["end_"], format="%m/%d/%y")
    ap = ap[ap["end"] <= run_date].copy()  # :84 / 2012:89-90

    # t = end - (1 + as.numeric(end-begin)) %/% 2      (:90)
    span = (ap["end"] - ap["begin"]).dt.days
    ap["t"] = ap["end"] - pd.to_timedelta((1 + span) // 2, unit="D")
    keep = (ap["t"] >= start_date) & ap["t"].notna() & (ap["n"] > 1)
    if spec.year == 2016:
        keep &= ap["population"].isin(["Likely Voters", "Registered Voters", "Adults"])  # :92-94
    df = ap[keep].copy()

    # pollster mutations (:98-105)
    def _extract(s):
        m = re.search(r"[A-z0-9 ]+", s)
        return re.sub(r"\s+$", "", m.group(0)) if m else np.nan
    df["pollster"] = df["pollster"].map(_extract)
    repl = {"Fox News": "FOX", "WashPost": "Washington Post", "ABC News": "ABC"}
    if spec.year == 2016:
        repl.update({"DHM Research": "DHM", "Public Opinion Strategies": "POS"})
    df["pollster"] = df["pollster"].replace(repl)

    if spec.year == 2016:
        mode_l = df["mode"].fillna("").str.lower()
        df["mode"] = np.where(df["mode"] == "Internet", "Online poll",
                              np.where(mode_l.str.contains("live phone"), "Live phone component", "Other"))  # :112-116
        df["polltype"] = df["population"]
    else:
        rec = {"Likely Voters": 0.0, "Registered Voters": 1.0, "Adults": 2.0}
        df["polltype"] = df["population"].map(rec)  # others -> NA (as.integer of a non-number)

    dem, rep = df[spec.dem].astype(float), df[spec.rep].astype(float)
    df["two_party_sum"] = dem + rep
    df["n_dem"] = _r_round(df["n"] * dem / 100)  # :123
    df["n_rep"] = _r_round(df["n"] * rep / 100)  # :125

    # numerical indices (:129-141), computed BEFORE distinct()
    state_abb_list = list(pd.read_csv(os.path.join(data_dir, "potus_results_76_16.csv"))["state"].drop_duplicates())
    levels = ["--"] + state_abb_list
    tmin = df["t"].min()
    df["poll_day"] = (df["t"] - tmin).dt.days + 1
    idx = df["state"].map({s: i + 1 for i, s in enumerate(levels)})
    df["index_s"] = np.where(idx == 1, 52, idx - 1)
    pl = sorted(df["pollster"].dropna().unique())
    df["index_p"] = df["pollster"].map({p: i + 1 for i, p in enumerate(pl)})
    if spec.year == 2016:
        ml = sorted(df["mode"].unique())
        df["index_m"] = df["mode"].map({p: i + 1 for i, p in enumerate(ml)})
        pol = sorted(df["polltype"].unique())
        df["index_pop"] = df["polltype"].map({p: i + 1 for i, p in enumerate(pol)})

    # arrange(state, t, polltype, two_party_sum) %>% distinct(state, t, pollster, .keep_all = TRUE)  (:143-144)
    df = df.sort_values(["state", "t", "polltype", "two_party_sum"], kind="stable", na_position="last")
    df = df.drop_duplicates(subset=["state", "t", "pollster"], keep="first").reset_index(drop=True)
    return df


def build_datalist(year: int, data_dir: str = "/root/reference/data") -> dict:
    """Return the named `data` list of final_{year}.R as a dict of numpy arrays / scalars
    (1-based indices, exactly as R hands them to Stan)."""
    spec = SPECS[year]
    election_day = pd.Timestamp(spec.election_day)
    run_date = pd.Timestamp(spec.run_date)
    df = _wrangle_polls(spec, data_dir)

    # state context (final_2016.R:166-189; final_2012.R:171-193)
    ctx = _read_csv_any_newline(os.path.join(data_dir, spec.context_csv))
    if spec.growth:
        share = ctx["total_count"] * (1 + ctx["adult_pop_growth_2011_15"])
    else:
        share = ctx["total_count"].astype(float)
    ctx["share_national_vote"] = share / share.sum()
    ctx = ctx.sort_values("state", kind="stable").reset_index(drop=True)
    states = list(ctx["state"])
    state_weights = (ctx["share_national_vote"] / ctx["share_national_vote"].sum()).to_numpy()

    # covariance (final_2016.R:198-325) -- identical in the three drivers (all use the 2016 dem share)
    res = pd.read_csv(os.path.join(data_dir, "potus_results_76_16.csv"))
    res = res[res["year"] == 2016][["state", "dem"]].dropna()
    feat = {"2016": res.set_index("state")["dem"]}
    census = pd.read_csv(os.path.join(data_dir, "acs_2013_variables.csv"))
    census = census[census["state"].notna()].drop(columns=["state_fips", "pop_total", "pop_density"]).set_index("state")
    for c in census.columns:
        feat[c] = census[c]
    urb = pd.read_csv(os.path.join(data_dir, "urbanicity_index.csv")).set_index("state")
    feat["pop_density"] = urb["average_log_pop_within_5_miles"]
    wev = pd.read_csv(os.path.join(data_dir, "white_evangel_pct.csv")).set_index("state")
    feat["pct_white_evangel"] = wev["pct_white_evangel"]
    fm = pd.DataFrame(feat)  # rows = state, cols = variable
    fm = (fm - fm.min()) / (fm.max() - fm.min())  # min-max per variable (:252-254)
    fm = fm.T.dropna(axis=0)  # spread(state,value) %>% na.omit(): drop variables with any NA
    fm = fm[sorted(fm.columns)]
    assert list(fm.columns) == states, "state order mismatch between covariance features and weights"
    C = np.corrcoef(fm.to_numpy().T)  # cor() over the variables, 51x51
    C[C < 0] = 0.0
    new_C = make_positive_definite(0.75 * C + 0.25 * np.ones((51, 51)))
    new_C = make_positive_definite(new_C)
    state_covariance_0 = cov_matrix(51, 0.07 ** 2, 0.9) * new_C

    days_til_election = (election_day - run_date).days
    mu_b_T_scale = (0.03 + (10 ** -6.6) * days_til_election ** 2) * 4  # :335-341,471
    polling_bias_scale = 0.013 * 4
    random_walk_scale = 0.05 / np.sqrt(300) * 4

    # priors (:403-410)
    pri = pd.read_csv(os.path.join(data_dir, "state_priors_08_12_16.csv"), parse_dates=["date"])
    pri = pri[pri["date"] <= run_date]
    pri = pri[pri["date"] == pri.groupby("state")["date"].transform("max")]
    pri = pri.sort_values("state", kind="stable")
    assert list(pri["state"]) == states
    p = pri["pred"].to_numpy()
    mu_b_prior = np.log(p / (1 - p))

    first_day = df["begin"].min()
    T = int(round((election_day - first_day).days))
    st = df[df["index_s"] != 52]
    na = df[df["index_s"] == 52]
    unadj = (~df["pollster"].isin(ADJUSTERS)).astype(float)

    data = dict(
        N_national_polls=len(na), N_state_polls=len(st), T=T, S=51,
        P=int(df["pollster"].nunique()),
        M=int(df["mode"].nunique()) if "mode" in df else 1,
        Pop=int(df["polltype"].nunique()),
        state=st["index_s"].to_numpy(dtype=np.int32),
        state_weights=state_weights,
        day_state=st["poll_day"].to_numpy(dtype=np.int32),
        day_national=na["poll_day"].to_numpy(dtype=np.int32),
        poll_state=st["index_p"].to_numpy(dtype=np.int32),
        poll_national=na["index_p"].to_numpy(dtype=np.int32),
        unadjusted_national=unadj[na.index].to_numpy(),
        unadjusted_state=unadj[st.index].to_numpy(),
        n_democrat_national=na["n_dem"].to_numpy(dtype=np.int32),
        n_democrat_state=st["n_dem"].to_numpy(dtype=np.int32),
        n_two_share_national=(na["n_dem"] + na["n_rep"]).to_numpy(dtype=np.int32),
        n_two_share_state=(st["n_dem"] + st["n_rep"]).to_numpy(dtype=np.int32),
        sigma_measure_noise_national=0.04, sigma_measure_noise_state=0.04,
        mu_b_prior=mu_b_prior, sigma_c=0.06, sigma_m=0.04, sigma_pop=0.04, sigma_e_bias=0.02,
        state_covariance_0=state_covariance_0,
        polling_bias_scale=polling_bias_scale, mu_b_T_scale=mu_b_T_scale, random_walk_scale=random_walk_scale,
    )
    if spec.full_model:
        data.update(
            poll_mode_national=na["index_m"].to_numpy(dtype=np.int32),
            poll_mode_state=st["index_m"].to_numpy(dtype=np.int32),
            poll_pop_national=na["index_pop"].to_numpy(dtype=np.int32),
            poll_pop_state=st["index_pop"].to_numpy(dtype=np.int32),
        )
    else:
        data["sigma_a"] = 0.012  # carried but unused (final_2012.R:489,525)
    data["_state_names"] = np.array(states)
    data["_ev_state"] = ctx["ev"].to_numpy(dtype=np.int32)
    data["_year"] = year
    return data


def synthetic_datalist(S=256, T=365, N_state=40000, N_national=10000, P=512, M=3, Pop=3, seed=1843) -> dict:
    """BASELINE.json config 5 / SURVEY.md section 8(d): synthetic stress problem with the same
    structure as the real list (covariance mirrors final_2016.R:272-295,324-325)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    feat = rng.standard_normal((9, S))
    C = np.corrcoef(feat.T)
    C[C < 0] = 0
    new_C = make_positive_definite(0.75 * C + 0.25)
    cov0 = cov_matrix(S, 0.07 ** 2, 0.9) * new_C
    w = rng.dirichlet(np.ones(S))
    prior = rng.normal(0, 0.3, S)
    nat_sd = np.sqrt(w @ cov0 @ w)
    L = np.linalg.cholesky(cov0)
    a_T, a_w = 0.12 / nat_sd, (0.05 / np.sqrt(300) * 4) / nat_sd
    mu = np.empty((S, T))
    mu[:, T - 1] = prior + a_T * (L @ rng.standard_normal(S))
    for t in range(T - 2, -1, -1):
        mu[:, t] = mu[:, t + 1] + a_w * (L @ rng.standard_normal(S))
    nat = w @ mu

    def polls(N, national):
        day = rng.integers(1, T + 1, N)
        st = rng.choice(S, N, p=w) + 1
        zipf = 1.0 / np.arange(1, P + 1)
        pol = rng.choice(P, N, p=zipf / zipf.sum()) + 1
        mode = rng.integers(1, M + 1, N)
        pop = rng.integers(1, Pop + 1, N)
        un = (rng.random(N) < 0.78).astype(float)
        n = np.clip(np.rint(rng.lognormal(np.log(700), 0.6, N)), 100, 60000).astype(np.int32)
        eta = nat[day - 1] if national else mu[st - 1, day - 1]
        eta = eta + rng.normal(0, 0.05, N)
        y = rng.binomial(n, 1 / (1 + np.exp(-eta))).astype(np.int32)
        return day.astype(np.int32), st.astype(np.int32), pol.astype(np.int32), mode.astype(np.int32), pop.astype(np.int32), un, n, y

    ds, ss, ps, ms, os_, us, ns, ys = polls(N_state, False)
    dn, _, pn, mn, on, un, nn, yn = polls(N_national, True)
    return dict(
        N_national_polls=N_national, N_state_polls=N_state, T=T, S=S, P=P, M=M, Pop=Pop,
        state=ss, state_weights=w, day_state=ds, day_national=dn, poll_state=ps, poll_national=pn,
        poll_mode_state=ms, poll_mode_national=mn, poll_pop_state=os_, poll_pop_national=on,
        unadjusted_state=us, unadjusted_national=un,
        n_democrat_state=ys, n_democrat_national=yn, n_two_share_state=ns, n_two_share_national=nn,
        sigma_measure_noise_national=0.04, sigma_measure_noise_state=0.04, mu_b_prior=prior,
        sigma_c=0.06, sigma_m=0.04, sigma_pop=0.04, sigma_e_bias=0.02, state_covariance_0=cov0,
        polling_bias_scale=0.052, mu_b_T_scale=0.12, random_walk_scale=0.05 / np.sqrt(300) * 4,
    )


def save_npz(path: str, data: dict) -> None:
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in data.items()})


def load_npz(path: str) -> dict:
    z = np.load(path, allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        out[k] = v.item() if v.ndim == 0 else v
    return out
