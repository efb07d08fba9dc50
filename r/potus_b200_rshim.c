/* R .Call shim for libpotus_b200.so -- the reference-side binding of include/potus_b200.h.
 *
 * Replaces, in scripts/model/final_2016.R:532-543 (and final_2012.R:558-569, final_2008.R:562-573),
 *     model <- cmdstanr::cmdstan_model(...); fit <- model$sample(data = data, seed = 1843, ...)
 *     out   <- rstan::read_stan_csv(fit$output_files())
 * by
 *     dyn.load("potus_b200_rshim.so")
 *     out <- .Call("potus_nuts_sample", data, list(chains = 1024L, iter_warmup = 500L, iter_sampling = 500L,
 *                                                 seed = 1843, keep_per_chain = 3L))
 * `out` is a named list of REAL arrays with the dims rstan::extract(out, pars)[[1]] has
 * (mu_b [draws,S,T], predicted_score [draws,T,S], ...), see r/potus_b200.R for extract_b200().
 *
 * NOT COMPILED IN THIS IMAGE: R (Rinternals.h, libR) is absent.  Build where R exists with r/build_rshim.sh
 * (R CMD SHLIB with -Iinclude and the in-tree libpotus_b200.so on the link line).
 * Contract with R: look names up with R_NamesSymbol, ignore unknown names, accept INTSXP or integral REALSXP
 * for integer fields (R hands `state`, `poll_*`, `n_democrat_*` over as doubles, final_2016.R:436-460),
 * PROTECT every allocation, and call Rf_error() only after every C resource is released (it longjmps).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Utils.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "potus_b200.h"

/* R_CheckUserInterrupt longjmps; run it under R_ToplevelExec so that the sampler can be destroyed first */
static void check_interrupt_fn(void* dummy) { (void)dummy; R_CheckUserInterrupt(); }
static int interrupted(void) { return R_ToplevelExec(check_interrupt_fn, NULL) == FALSE; }

static SEXP list_get(SEXP lst, const char* name) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  if (names == R_NilValue) return R_NilValue;
  for (R_xlen_t i = 0; i < XLENGTH(lst); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
  return R_NilValue;
}
/* integer vector from INTSXP or integral REALSXP; caller frees */
static int32_t* as_i32(SEXP x, R_xlen_t* n, int* ok) {
  if (x == R_NilValue) { *n = 0; return NULL; }
  *n = XLENGTH(x);
  int32_t* r = (int32_t*)malloc(sizeof(int32_t) * (*n > 0 ? *n : 1));
  if (TYPEOF(x) == INTSXP) memcpy(r, INTEGER(x), sizeof(int32_t) * *n);
  else if (TYPEOF(x) == REALSXP) {
    for (R_xlen_t i = 0; i < *n; ++i) { double v = REAL(x)[i]; if (v != (double)(int32_t)v) *ok = 0; r[i] = (int32_t)v; }
  } else *ok = 0;
  return r;
}
static double* as_f64(SEXP x, R_xlen_t* n, int* ok) {
  if (x == R_NilValue) { *n = 0; return NULL; }
  *n = XLENGTH(x);
  double* r = (double*)malloc(sizeof(double) * (*n > 0 ? *n : 1));
  if (TYPEOF(x) == REALSXP) memcpy(r, REAL(x), sizeof(double) * *n);
  else if (TYPEOF(x) == INTSXP) for (R_xlen_t i = 0; i < *n; ++i) r[i] = INTEGER(x)[i];
  else *ok = 0;
  return r;
}
static double scalar(SEXP lst, const char* name, double dflt) {
  SEXP x = list_get(lst, name);
  return x == R_NilValue ? dflt : Rf_asReal(x);
}
static SEXP fetch(PotusSampler* s, const char* par, int ndim, const int* dims, int* rc) {
  size_t n = potus_draws_size(s, par);
  if (n == 0) return R_NilValue;
  SEXP a = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
  *rc = potus_get_draws(s, par, REAL(a), n);
  SEXP d = PROTECT(Rf_allocVector(INTSXP, ndim));
  for (int i = 0; i < ndim; ++i) INTEGER(d)[i] = dims[i];
  Rf_setAttrib(a, R_DimSymbol, d);
  UNPROTECT(2);
  return a;
}

SEXP potus_nuts_sample(SEXP data, SEXP config) {
  PotusData d; memset(&d, 0, sizeof d);
  void* owned[32]; int n_owned = 0, ok = 1, dim_ok = 1; R_xlen_t n; const char* bad_dim = "";
  /* want < 0: no length check.  A vector shorter than the scalar sizes say would be read past its end by the library:
   * refuse it here with Stan's own wording ("mismatch in dimension declared and found in context") */
#define IVEC(field, want) do { d.field = as_i32(list_get(data, #field), &n, &ok); owned[n_owned++] = (void*)d.field; \
    if (d.field && (want) >= 0 && n != (R_xlen_t)(want)) { dim_ok = 0; bad_dim = #field; } } while (0)
#define DVEC(field, want) do { d.field = as_f64(list_get(data, #field), &n, &ok); owned[n_owned++] = (void*)d.field; \
    if (d.field && (want) >= 0 && n != (R_xlen_t)(want)) { dim_ok = 0; bad_dim = #field; } } while (0)
  d.N_national_polls = (int32_t)scalar(data, "N_national_polls", 0); d.N_state_polls = (int32_t)scalar(data, "N_state_polls", 0);
  d.T = (int32_t)scalar(data, "T", 0); d.S = (int32_t)scalar(data, "S", 0); d.P = (int32_t)scalar(data, "P", 0);
  d.M = (int32_t)scalar(data, "M", 1); d.Pop = (int32_t)scalar(data, "Pop", 1);
  { const int Ns = d.N_state_polls, Nn = d.N_national_polls, S = d.S;
  IVEC(state, Ns); IVEC(day_state, Ns); IVEC(day_national, Nn); IVEC(poll_state, Ns); IVEC(poll_national, Nn);
  IVEC(poll_mode_state, Ns); IVEC(poll_mode_national, Nn); IVEC(poll_pop_state, Ns); IVEC(poll_pop_national, Nn);
  IVEC(n_democrat_national, Nn); IVEC(n_two_share_national, Nn); IVEC(n_democrat_state, Ns); IVEC(n_two_share_state, Ns);
  DVEC(unadjusted_national, Nn); DVEC(unadjusted_state, Ns); DVEC(mu_b_prior, S); DVEC(state_weights, S); DVEC(state_covariance_0, S * S); }
  d.sigma_c = scalar(data, "sigma_c", 0); d.sigma_m = scalar(data, "sigma_m", 0); d.sigma_pop = scalar(data, "sigma_pop", 0);
  d.sigma_measure_noise_national = scalar(data, "sigma_measure_noise_national", 0);
  d.sigma_measure_noise_state = scalar(data, "sigma_measure_noise_state", 0); d.sigma_e_bias = scalar(data, "sigma_e_bias", 0);
  d.random_walk_scale = scalar(data, "random_walk_scale", 0); d.mu_b_T_scale = scalar(data, "mu_b_T_scale", 0);
  d.polling_bias_scale = scalar(data, "polling_bias_scale", 0);
  if (d.poll_mode_state == NULL) { d.unadjusted_national = NULL; d.unadjusted_state = NULL; }  /* no-mode variant ignores them */

  PotusConfig c; memset(&c, 0, sizeof c);
  c.chains = (int32_t)scalar(config, "chains", 4); c.iter_warmup = (int32_t)scalar(config, "iter_warmup", 500);
  c.iter_sampling = (int32_t)scalar(config, "iter_sampling", 500); c.keep_per_chain = (int32_t)scalar(config, "keep_per_chain", 0);
  c.max_treedepth = (int32_t)scalar(config, "max_treedepth", 10); c.device = (int32_t)scalar(config, "device", 0);
  c.seed = (uint64_t)scalar(config, "seed", 1843); c.adapt_delta = scalar(config, "adapt_delta", 0.8);
  c.init_radius = scalar(config, "init", 2.0); c.chain_id_offset = (int32_t)scalar(config, "chain_id_offset", 0);
  c.n_gpus = (int32_t)scalar(config, "n_gpus", 1);   /* > 1: the library shards the chains over that many devices and all-gathers the draws */

  char err[512]; err[0] = 0;
  PotusSampler* s = NULL;
  int rc = (ok && dim_ok) ? potus_create(&d, &c, &s) : POTUS_ERR_INVALID_DATA;
  if (!ok) strncpy(err, "data list: integer fields must hold integers", sizeof err - 1);
  else if (!dim_ok) snprintf(err, sizeof err, "Exception: mismatch in dimension declared and found in context; processing stage=data initialization; variable name=%s", bad_dim);
  if (rc == POTUS_OK) {   /* run without blocking R's event loop: poll the device, let the user interrupt */
    rc = potus_run_begin(s);
    int done = 0, stop = 0;
    while (rc == POTUS_OK && !done && !stop) {
      usleep(100000);
      rc = potus_run_poll(s, &done);
      if (!done && interrupted()) stop = 1;
    }
    if (stop) { potus_destroy(s); for (int i = 0; i < n_owned; ++i) free(owned[i]); Rf_error("potus_b200: interrupted"); }  /* (destroy waits for the device) */
    if (rc == POTUS_OK) rc = potus_run_end(s);
  }
  if (rc != POTUS_OK && ok && dim_ok) strncpy(err, potus_last_error(), sizeof err - 1);
  SEXP out = R_NilValue;
  int nprot = 0;
  if (rc == POTUS_OK) {
    PotusStats st; potus_get_stats(s, &st);
    const int nd = st.n_draws_kept, nit = c.iter_warmup + c.iter_sampling;
    const char* names[] = {"mu_b", "mu_c", "mu_m", "mu_pop", "polling_bias", "e_bias", "predicted_score", "monitor", "sampler_params"};
    const int dims[][3] = {{nd, d.S, d.T}, {nd, d.P, 0}, {nd, d.M, 0}, {nd, d.Pop, 0}, {nd, d.S, 0}, {nd, d.T, 0}, {nd, d.T, d.S},
                           {c.chains * c.iter_sampling, d.S + 1, 0}, {c.chains * nit, 7, 0}};
    const int nds[] = {3, 2, 2, 2, 2, 2, 3, 2, 2};
    out = PROTECT(Rf_allocVector(VECSXP, 13)); ++nprot;
    SEXP nm = PROTECT(Rf_allocVector(STRSXP, 13)); ++nprot;
    for (int i = 0; i < 9 && rc == POTUS_OK; ++i) {
      SET_VECTOR_ELT(out, i, fetch(s, names[i], nds[i], dims[i], &rc));
      SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
    }
    SET_VECTOR_ELT(out, 9, Rf_mkString(d.poll_mode_state ? "poll_model_2020" : "poll_model_2020_no_mode_adjustment"));
    SET_STRING_ELT(nm, 9, Rf_mkChar("model_name"));
    if (rc == POTUS_OK) {
      /* chain_id [draws]: global chain of every kept draw (chains are concatenated in chain order) */
      const int keep = c.chains > 0 ? nd / c.chains : 0;
      SEXP cid = PROTECT(Rf_allocVector(INTSXP, nd)); ++nprot;
      for (int r = 0; r < nd; ++r) INTEGER(cid)[r] = c.chain_id_offset + (keep > 0 ? r / keep : 0) + 1;
      SET_VECTOR_ELT(out, 10, cid); SET_STRING_ELT(nm, 10, Rf_mkChar("chain_id"));
      /* timing: device seconds (total, warm-up, sampling, all-gather), leapfrogs, launches */
      SEXP tm = PROTECT(Rf_allocVector(REALSXP, 7)); ++nprot;
      SEXP tn = PROTECT(Rf_allocVector(STRSXP, 7)); ++nprot;
      const char* tnames[] = {"seconds_total", "seconds_warmup", "seconds_sampling", "seconds_gather", "n_leapfrog_total", "n_divergent_sampling", "gpu_launches"};
      const double tv[] = {st.seconds_total, st.seconds_warmup, st.seconds_sampling, st.seconds_gather, (double)st.n_leapfrog_total,
                           (double)st.n_divergent_sampling, (double)st.gpu_launches};
      for (int i = 0; i < 7; ++i) { REAL(tm)[i] = tv[i]; SET_STRING_ELT(tn, i, Rf_mkChar(tnames[i])); }
      Rf_setAttrib(tm, R_NamesSymbol, tn);
      SET_VECTOR_ELT(out, 11, tm); SET_STRING_ELT(nm, 11, Rf_mkChar("timing"));
      /* diagnostics [S+1, 3]: Stan ESS, split R-hat, mean of mu_b[,T] (logit) and national_mu_b_average[T], from the
       * on-device post-processing over every sampling iteration; state_table [S+2, 8] beside it */
      if (c.n_gpus <= 1 && c.iter_sampling >= 4) {
        SEXP dg = PROTECT(Rf_allocMatrix(REALSXP, 3, d.S + 1)); ++nprot;      /* filled row-major [S+1][3] = column-major [3, S+1] */
        SEXP tb = PROTECT(Rf_allocMatrix(REALSXP, 8, d.S + 2)); ++nprot;
        rc = potus_postprocess(s, NULL, 270.0, REAL(tb), REAL(dg));
        SEXP both = PROTECT(Rf_allocVector(VECSXP, 2)); ++nprot;
        SEXP bn = PROTECT(Rf_allocVector(STRSXP, 2)); ++nprot;
        SET_VECTOR_ELT(both, 0, dg); SET_STRING_ELT(bn, 0, Rf_mkChar("ess_rhat_mean"));
        SET_VECTOR_ELT(both, 1, tb); SET_STRING_ELT(bn, 1, Rf_mkChar("state_table"));
        Rf_setAttrib(both, R_NamesSymbol, bn);
        SET_VECTOR_ELT(out, 12, both);
      }
      SET_STRING_ELT(nm, 12, Rf_mkChar("diagnostics"));
    }
    Rf_setAttrib(out, R_NamesSymbol, nm);
    if (rc != POTUS_OK) strncpy(err, potus_last_error(), sizeof err - 1);
  }
  if (s) potus_destroy(s);
  for (int i = 0; i < n_owned; ++i) free(owned[i]);
  if (nprot) UNPROTECT(nprot);
  if (rc != POTUS_OK) Rf_error("potus_b200: %s", err);   /* all C resources are released; safe to longjmp */
  return out;
}
