/* potus_b200.h -- C-ABI of the B200-native NUTS sampler for The Economist's poll model.
 *
 * This is the drop-in boundary for the ONE hot path of TheEconomist/us-potus-model: everything
 * that happens between the R drivers handing the named `data` list to Stan and getting draws back
 *   scripts/model/final_2016.R:532-543   cmdstan_model(...)$sample(data=, seed=, chains=, ...) ;
 *                                        rstan::read_stan_csv(fit$output_files())
 *   scripts/model/final_2012.R:558-569, scripts/model/final_2008.R:562-573   (same call)
 * i.e. the log-posterior + gradient of scripts/model/poll_model_2020.stan (and the
 * _no_mode_adjustment variant), leapfrog, multinomial NUTS, Stan's warm-up adaptation, and the
 * transformed parameters / generated quantities the R consumers extract
 * (final_2016.R:556,568,597,622,647,682,708; README.Rmd:206,716,1240).
 *
 * Plain C: pointers and sizes only.  All indices in PotusData are 1-BASED exactly as R/Stan hold
 * them (poll_model_2020.stan:9-17).  No C++ exception or CUDA error crosses this boundary: every
 * entry point returns POTUS_OK or a negative status and leaves a message for potus_last_error().
 * The reference-side binding (R .Call shim) is r/potus_b200_rshim.c; see INTEGRATION.md.
 */
#ifndef POTUS_B200_H
#define POTUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define POTUS_API __attribute__((visibility("default")))
#else
#define POTUS_API
#endif

#define POTUS_OK 0
#define POTUS_ERR_INVALID_DATA (-1)   /* Stan data-block constraint violated (poll_model_2020.stan:9-23,37) */
#define POTUS_ERR_UNSUPPORTED (-2)    /* problem size outside both kernel families (resident: S<=51,T<=254,N<=1648; streaming: S<=256,T<=512) */
#define POTUS_ERR_CUDA (-3)           /* CUDA runtime / driver error, or no sm_100 device */
#define POTUS_ERR_STATE (-4)          /* call order (e.g. get_draws before run) or unknown name */
#define POTUS_ERR_INIT (-5)           /* no finite initial point after 100 attempts (Stan's rule) */

/* The named list of final_2016.R:475-514 (2012/2008: final_2012.R:501-540, final_2008.R:504-545;
 * their extra names sigma_a/M/Pop/unadjusted_* are accepted and ignored by the binding).
 * poll_mode_* / poll_pop_* == NULL selects poll_model_2020_no_mode_adjustment.stan. */
typedef struct PotusData {
  int32_t N_national_polls, N_state_polls, T, S, P, M, Pop;
  const int32_t* state;              /* [N_state_polls]    1..S              */
  const int32_t* day_state;          /* [N_state_polls]    1..T              */
  const int32_t* day_national;       /* [N_national_polls] 1..T              */
  const int32_t* poll_state;         /* [N_state_polls]    1..P              */
  const int32_t* poll_national;      /* [N_national_polls] 1..P              */
  const int32_t* poll_mode_state;    /* [N_state_polls]    1..M   or NULL    */
  const int32_t* poll_mode_national; /* [N_national_polls] 1..M   or NULL    */
  const int32_t* poll_pop_state;     /* [N_state_polls]    1..Pop or NULL    */
  const int32_t* poll_pop_national;  /* [N_national_polls] 1..Pop or NULL    */
  const int32_t* n_democrat_national;
  const int32_t* n_two_share_national;
  const int32_t* n_democrat_state;
  const int32_t* n_two_share_state;
  const double* unadjusted_national; /* in [0,1]; may be NULL for the no-mode variant */
  const double* unadjusted_state;
  const double* mu_b_prior;          /* [S] */
  const double* state_weights;       /* [S] */
  double sigma_c, sigma_m, sigma_pop;
  double sigma_measure_noise_national, sigma_measure_noise_state, sigma_e_bias;
  const double* state_covariance_0;  /* [S*S] column-major, symmetric positive definite */
  double random_walk_scale, mu_b_T_scale, polling_bias_scale;
} PotusData;

/* Mirrors the cmdstanr `$sample()` argument names used at final_2016.R:533-541 plus Stan's
 * defaults for what the reference never overrides (adapt_delta 0.8, max_treedepth 10, init 2). */
typedef struct PotusConfig {
  int32_t chains;          /* chains run by THIS sampler (this GPU)                                */
  int32_t chain_id_offset; /* global id of its first chain: RNG streams are keyed by global id,   */
                           /* so sharding chains over GPUs/ranks does not change any chain         */
  int32_t iter_warmup;     /* default 500 */
  int32_t iter_sampling;   /* default 500 */
  int32_t keep_per_chain;  /* full draws (mu_b, ...) kept per chain, evenly thinned from the       */
                           /* sampling iterations; 0 = keep all.  The 52 monitored scalars and the  */
                           /* 7 sampler diagnostics are always kept for every iteration.           */
  int32_t max_treedepth;   /* default 10 */
  int32_t device;          /* CUDA device ordinal */
  int32_t flags;           /* bit 0 (POTUS_FLAG_FORCE_STREAM): run the streaming large-S/T kernel family even when    */
                           /* the problem fits the SMEM/TMEM-resident kernel (used by the parity tests)            */
  uint64_t seed;           /* default 1843 (final_2016.R:535) */
  double adapt_delta;      /* default 0.8 */
  double init_radius;      /* default 2.0: inits ~ U(-r, r) on the unconstrained scale */
  int32_t n_gpus;          /* 0 / 1: this device only.  > 1: `chains` is the TOTAL; ONE process shards it contiguously over     */
                           /* devices [device, device + n_gpus) (the reference's parallel_chains, final_2016.R:536), replicates */
                           /* the read-only model, and potus_run ends with ONE ncclAllGather of the kept-draw buffers           */
  int32_t reserved;
} PotusConfig;

#define POTUS_FLAG_FORCE_STREAM 1

typedef struct PotusStats {
  int64_t n_leapfrog_total;      /* all chains, warm-up + sampling                           */
  int64_t n_leapfrog_sampling;
  int64_t n_divergent_sampling;
  int64_t gpu_launches;          /* kernels launched by potus_run                            */
  double seconds_total;          /* device time of potus_run (CUDA events)                   */
  double seconds_warmup;
  double seconds_sampling;
  double mean_stepsize;          /* post-warm-up, mean over chains                           */
  double mean_accept_stat;       /* sampling iterations                                      */
  double mean_treedepth;
  int32_t n_params;              /* unconstrained dimension D                                */
  int32_t n_draws_kept;          /* chains * keep_per_chain                                  */
  double seconds_gather;         /* n_gpus > 1: device time of the ncclAllGather (included in seconds_total; the other   */
                                 /* seconds_* are the max over devices)                                              */
} PotusStats;

typedef struct PotusSampler PotusSampler;

/* Validate `data` against the Stan data block, build device-side structures, draw inits. */
POTUS_API int potus_create(const PotusData* data, const PotusConfig* config, PotusSampler** out);
/* Run warm-up + sampling for all chains; blocks until done. */
POTUS_API int potus_run(PotusSampler* s);
/* The same in three steps for hosts with an event loop (the R shim polls R_CheckUserInterrupt between polls): begin enqueues all
 * device work and returns at once, poll sets *done when it has finished, end collects statistics (and runs the all-gather). */
POTUS_API int potus_run_begin(PotusSampler* s);
POTUS_API int potus_run_poll(PotusSampler* s, int* done);
POTUS_API int potus_run_end(PotusSampler* s);
/* Optional, before potus_run, with iter_warmup == 0: start every chain from a given adapted state (resume / continue a run,
 * or seed the sampling phase from another sampler's warm-up) instead of Stan's random inits + warm-up.
 * theta [chains][D] row-major, Stan unconstrained order; stepsize [chains]; inv_metric [chains][D] (diagonal of M^-1). */
POTUS_API int potus_set_state(PotusSampler* s, const double* theta, const double* stepsize, const double* inv_metric);
/* Number of doubles potus_get_draws would write for `par` (0 if unknown). */
POTUS_API size_t potus_draws_size(const PotusSampler* s, const char* par);
/* Copy kept draws of one quantity to host, shaped like rstan::extract(out, pars=par)[[1]]:
 * R column-major with the draw index fastest, chains concatenated in chain order:
 *   "mu_b" [draws,S,T]  "mu_c" [draws,P]  "mu_m" [draws,M]  "mu_pop" [draws,Pop]
 *   "polling_bias" [draws,S]  "e_bias" [draws,T]  "predicted_score" [draws,T,S]
 *   "theta" [draws,D] (unconstrained, Stan order)
 *   "monitor" [iter_sampling*chains, S+1]  every sampling iteration: mu_b[,T] and national_mu_b_average[T]
 *   "sampler_params" [(iter_warmup+iter_sampling)*chains, 7]
 *        lp__, accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__
 *   "inv_metric" [chains, D]  adapted diagonal of the inverse metric per chain (the numbers CmdStan prints
 *        under "# Diagonal elements of inverse mass matrix:"), Stan unconstrained order               */
POTUS_API int potus_get_draws(PotusSampler* s, const char* par, double* out, size_t n);
POTUS_API int potus_get_stats(PotusSampler* s, PotusStats* stats);
/* On-device post-processing over ALL chains x iter_sampling monitored draws -- what the reference's reports compute on the host from
 * rstan::extract(out, "predicted_score") (README.Rmd:206-300, final_2016.R:708-823), done where the draws are:
 *   ev [S] electoral votes or NULL; ev_threshold (270);
 *   state_table [(S+2)][8] row-major: mean, sd, 2.5%, 5%, 50%, 95%, 97.5% quantiles (exact order statistics, linear interpolation as
 *     numpy / R type 7), P(> 0.5) of inv_logit(mu_b[s,T]); row S = national vote (state_weights-weighted mean of the shares of each draw);
 *     row S+1 = democratic electoral votes of each draw, last column P(ev >= ev_threshold) (zeros when ev is NULL);
 *   ess_table [(S+1)][3] row-major or NULL: Stan effective sample size, split R-hat and mean of the monitored scalars on the logit scale
 *     (row S = national_mu_b_average[T]). */
POTUS_API int potus_postprocess(PotusSampler* s, const double* ev, double ev_threshold, double* state_table, double* ess_table);
/* Device pointers to the raw fp32 buffers (for the torch.distributed all-gather in bench.py; with n_gpus > 1, which = 0
 * is the all-gathered buffer on the first device, [n_gpus][ceil(chains/n_gpus)*keep][draw_len], and 1 / 2 are not available):
 *   which = 0: kept draws   [chains*keep][draw_len]   (draw_len floats per draw, Stan block order:
 *              mu_b | mu_c | mu_m | mu_pop | e_bias | polling_bias | theta)
 *   which = 1: monitor      [chains][iter_sampling][S+1]
 *   which = 2: sampler_params [chains][iter_warmup+iter_sampling][8]  (lp__ and energy__ centred by the
 *              data constant; potus_get_draws("sampler_params") returns them uncentred in fp64) */
POTUS_API int potus_device_buffer(PotusSampler* s, int which, void** dptr, size_t* n_floats);
POTUS_API void potus_destroy(PotusSampler* s);
POTUS_API const char* potus_last_error(void);

/* Test hook: log density (constants dropped as Stan's `~` does, log(0.02) Jacobian constant
 * excluded) and gradient for n_chains unconstrained vectors theta[n_chains][D] (Stan order),
 * evaluated by the same device code the sampler uses. */
POTUS_API int potus_logp_grad(const PotusData* data, const double* theta, int n_chains, double* lp, double* grad);
/* Same, with force_stream != 0 evaluating on the streaming kernel family (BASELINE config 5 path) whatever the size. */
POTUS_API int potus_logp_grad_ex(const PotusData* data, const double* theta, int n_chains, double* lp, double* grad, int force_stream);
/* Unconstrained dimension for a data list (15098 for the 2016 list). */
POTUS_API int potus_num_params(const PotusData* data);

#ifdef __cplusplus
}
#endif
#endif /* POTUS_B200_H */
