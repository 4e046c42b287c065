// drjacoby_b200 model template -- the CUDA analogue of the reference's
// inst/templates/cpp_template.cpp.  Fill in the two functions and pass the file (or its text) to
// drjacoby_b200.source_cuda(); then run_mcmc(..., loglike = "loglike", logprior = "logprior").
//
// What you get (generated from df_params / data / misc before compilation):
//   params[P_<name>]         parameter values in df_params order (the reference: params["<name>"])
//   data.item(DATA_<name>)   const double* to a data vector, data.len(DATA_<name>) its length
//   misc.scalar(MISC_<name>) a numeric element of misc; misc.item()/len() for vectors
//   block                    the likelihood block being evaluated (the reference: misc["block"])
//   R::dnorm, R::dlnorm, R::dgamma, R::dbeta, R::dpois, R::dbinom, R::dunif, R::dexp with R's
//   argument order, e.g. R::dnorm(x, mean, sd, true)
// Every chain x rung is one GPU thread: write plain scalar code, no shared state, no RNG calls.

__device__ double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
  // Insert loglikelihood code here:

  double loglikelihood = 0.0;
  return loglikelihood;
}

__device__ double logprior(const double* params, const DrjList& misc) {
  // Insert logprior code here:

  double logprior = 0.0;
  return logprior;
}

// NOTE: if your functions are not called "loglike" / "logprior", pass their names to run_mcmc().
