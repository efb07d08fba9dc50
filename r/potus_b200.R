# Drop-in for the sampling boundary of scripts/model/final_20{08,12,16}.R (lines 532-543 in final_2016.R).
#   source("r/potus_b200.R")
#   out <- potus_b200_sample(data, chains = 1024, iter_warmup = n_warmup, iter_sampling = n_sampling, seed = 1843)
#   mu_b_T_posterior_draw <- extract_b200(out, pars = "mu_b")[[1]][,,254]      # was rstan::extract(out, ...)
potus_b200_sample <- function(data, chains = 1024L, iter_warmup = 500L, iter_sampling = 500L, seed = 1843,
                              keep_per_chain = 3L, parallel_chains = NULL, refresh = NULL, adapt_delta = 0.8,
                              max_treedepth = 10L, n_gpus = 1L, lib = "r/potus_b200_rshim.so") {
  if (!is.loaded("potus_nuts_sample")) dyn.load(lib)
  .Call("potus_nuts_sample", data,
        list(chains = as.integer(chains), iter_warmup = as.integer(iter_warmup), iter_sampling = as.integer(iter_sampling),
             seed = seed, keep_per_chain = as.integer(keep_per_chain), adapt_delta = adapt_delta,
             max_treedepth = as.integer(max_treedepth), n_gpus = as.integer(n_gpus)))
}
# rstan::extract(out, pars = x) returns a named list of arrays [draws, dims...]; same shape here.
extract_b200 <- function(out, pars) out[pars]
# out also carries: out$sampler_params [iterations*chains, 7], out$monitor, out$chain_id [draws], out$timing (device seconds,
# leapfrogs), out$diagnostics$ess_rhat_mean [3, S+1] and out$diagnostics$state_table [8, S+2] (on-device post-processing over
# every sampling iteration), out$model_name.  Build the shim with  sh r/build_rshim.sh  (needs R; never compiled in this image).
