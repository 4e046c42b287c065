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

// ---------------------------------------------------------------------------------------------
// OPTIONAL: logarithms the sampler already has (HINTS).  For every parameter with a lower bound the sampler keeps
// log(theta_i - min_i), the Jacobian term of its reparameterisation.  Where min_i is exactly 0 (a standard deviation, a
// rate) that IS log(theta_i), which a scale parameter's likelihood and a log-normal or gamma prior would compute again.
// A model struct may add
//
//     static constexpr bool HINTS = true;
//     __device__ static double logprior_h(const double* params, const DrjList& misc, const double* loglo, unsigned mask) {
//       const double ls = (mask >> P_sigma & 1u) ? loglo[P_sigma] : log(params[P_sigma]);   // bit i set: loglo[i] == log(params[i])
//       return R::dunif(params[P_mu], -10.0, 10.0, true) + R::dlnorm_l(params[P_sigma], ls, 0.0, 1.0, true);
//     }
//     // per-datum models also:  prepare_h(params, misc, block, k, loglo, mask)  in place of prepare()
//
// (logprior() / prepare() stay: they serve the initial evaluation.)  Same function of the same argument: same bits.
// ---------------------------------------------------------------------------------------------
// OPTIONAL: quantities that depend on the data alone (CTA_INIT).  A model struct may add
//
//     static constexpr bool CTA_INIT = true;
//     __device__ static double* my_consts() { __shared__ double c[64]; return c; }       // shared memory the model owns
//     __device__ static void cta_init(const DrjList& data, const DrjList& misc) {        // every thread of every block, once
//       for (int j = threadIdx.x; j < 64 && j < data.len(DATA_y); j += blockDim.x) my_consts()[j] = lgamma(data.item(DATA_y)[j] + 1.0);
//     }
//
// and read my_consts() in loglike(): cta_init runs before the block's first evaluation of the model and is followed by a
// block barrier.  (The built-in logistic-growth model keeps the Stirling error and log(2 pi y) of its Poisson counts there.)
// ---------------------------------------------------------------------------------------------
// OPTIONAL: the per-datum form, for likelihoods that are a sum over many data points.
// The functions above are evaluated by ONE GPU thread per chain x rung, which walks all the data
// itself.  If instead you describe ONE term of the sum, the sampler stages the data once per thread
// block in shared memory (streamed by the TMA engine when it does not fit) and all chains x rungs of
// the block walk it together; and when a run has only a few chains (chains x rungs <= 4096) over a long
// data vector, every chain gets a whole cluster of thread blocks that divides the sum (one thread per
// chain x rung would take ~5 clock cycles per data point whatever the GPU).  Define a struct like this and `#define DRJ_USER_MODEL` to its name
// (loglike()/logprior() above are then unused and may be deleted):
//
//   struct MyModel {
//     static constexpr int D = DRJ_D, NBLOCK = DRJ_NBLOCK;     // from df_params
//     static constexpr bool DATUM_FORM = true;
//     struct Consts { double mean, half_inv_var, lognorm; };   // whatever depends on params only
//     __device__ static int datum_item(int block) { return DATA_x; }            // which data vector
//     __device__ static bool prepare(const double* params, const DrjList& misc, int block, Consts& k) {
//       k.mean = params[P_mu];  k.half_inv_var = 0.5 / (params[P_sigma] * params[P_sigma]);
//       k.lognorm = 0.918938533204672741780329736406 + log(params[P_sigma]);
//       return params[P_sigma] > 0;          // false -> loglike() below is used for this evaluation
//     }
//     __device__ static double datum(const Consts& k, double x, double acc) {   // acc += term(x)
//       const double d = x - k.mean;  return fma(d, d, acc);
//     }
//     __device__ static double finish(const Consts& k, double acc, long long n) {
//       return -((double)n * k.lognorm + k.half_inv_var * acc);
//     }
//     __device__ static double loglike(const double* params, const DrjList& data, const DrjList& misc, int block) {
//       double r = 0; for (long long i = 0; i < data.len(DATA_x); ++i) r += R::dnorm(data.item(DATA_x)[i], params[P_mu], params[P_sigma], true);
//       return r;                              // exact fallback, same value as prepare/datum/finish
//     }
//     __device__ static double logprior(const double* params, const DrjList& misc) { return 0.0; }
//   };
//   #define DRJ_USER_MODEL MyModel
//
// datum() must ADD its term to acc (the sampler keeps 8 partial sums and adds them at the end).
//
// ---------------------------------------------------------------------------------------------
// OPTIONAL: the lane form, for models with many parameters (more than 8) whose blocks are sums of terms.
// One GPU thread per chain x rung keeps theta and the block likelihoods of such a model in its own (slow) local
// memory and adds every block up alone.  If the model struct also says how to split a block over LANES threads,
// a particle is run by a group of LANES lanes of a warp, its state lives once in shared memory, and the lanes add
// their shares of a block in parallel (their partial sums are combined in a fixed order).  Add to the struct
// (DATUM_FORM = false; loglike() and logprior() as above):
//
//     static constexpr int LANES = 8;          // 4, 8, 16 or 32
//     __device__ static double loglike_lanes(const double* params, const DrjList& data, const DrjList& misc,
//                                            int block, int lane) {
//       // the terms of loglike(params, data, misc, block) that lane `lane` adds up; over lane = 0 .. LANES-1 every
//       // term exactly once, e.g. for a block that sums over a data vector:
//       double r = 0;
//       for (long long i = lane; i < data.len(block - 1); i += LANES) r += R::dnorm(data.item(block - 1)[i], params[block - 1], 1.0, true);
//       return r;
//     }
