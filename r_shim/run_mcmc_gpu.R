# run_mcmc_gpu.R -- the R-side change a drjacoby maintainer would make to route run_mcmc() through
# libdrjacoby_b200.so.  It is the dispatch section of R/main.R (lines 392-402: lapply / clusterApplyLB
# over deploy_chain, each calling main_cpp -> .Call(`_drjacoby_main_cpp`, args)) replaced by ONE
# .Call for all chains.  Everything before it (input checks, pre-processing, R/main.R:233-343) and
# after it (post-processing into the drjacoby_output object, R/main.R:411-577) is unchanged, because
# the shim returns `output_raw` in exactly the shape deploy_chain() produces.
#
# NOT runnable in this repository's image (no R).  See INTEGRATION.md.

dyn.load("drjacoby_gpu_shim.so")   # built with: R CMD SHLIB drjacoby_gpu_shim.c -I../include -L../drjacoby_b200 -ldrjacoby_b200

# `model` is either list(builtin = "example") or
#   list(source = paste(readLines("my_model.cu"), collapse = "\n"), loglike = "loglike", logprior = "logprior")
# R closures are rejected: the GPU path has no CPU fallback.
run_chains_gpu <- function(args_params, init_mat, chains, model, seed = 1, replay = NULL) {
  if (is.function(model$loglike) || is.function(model$logprior)) {
    stop("R functions cannot run on the GPU path: supply a CUDA model (see cuda_template.cu)")
  }
  names(args_params$theta_min) <- rownames(init_mat)     # parameter names travel with theta_min
  args_params$init_mat <- init_mat                        # d x chains, R/main.R:322-328
  args_params$chains <- as.integer(chains)
  args_params$seed <- as.numeric(seed)
  args_params$replay <- replay                            # NULL = Philox on the device
  args_params$model <- model
  .Call("drjgpu_main", list(args_params = args_params))
}

# In run_mcmc(), R/main.R:392-402 becomes:
#
#   if (gpu) {
#     rownames(init_mat) <- df_params$name
#     output_raw <- run_chains_gpu(args_params, init_mat, chains, model = gpu_model, seed = gpu_seed)
#   } else if (!is.null(cluster)) { ... unchanged ... }
#
# To reproduce a CPU run chain for chain (same set.seed), pass the uniforms the CPU run would draw:
#   U <- 3 * sum(!skip_param) * rungs + (coupling_on && rungs > 1) * (rungs - 1)
#   replay <- runif(chains * (burnin - 1 + samples) * U)   # after the default-init runif() calls
