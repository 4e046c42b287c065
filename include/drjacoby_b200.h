/*
 * drjacoby_b200.h -- C ABI of libdrjacoby_b200.so, the B200 (sm_100a) implementation of the
 * drjacoby MCMC sampling core.
 *
 * Drop-in boundary.  The reference crosses from R into native code through ONE registered .Call
 * routine, once per chain:
 *     R/RcppExports.R:4-6        main_cpp <- function(args) .Call(`_drjacoby_main_cpp`, args)
 *     src/RcppExports.cpp:15-33  RcppExport SEXP _drjacoby_main_cpp(SEXP argsSEXP) + R_init_drjacoby
 *     src/main.cpp:17-80         main_cpp(): dispatch on C++ / R user functions
 *     src/main.cpp:84-278        run_mcmc<T1,T2>(): the chain
 * This header is what a .Call shim binds instead (r_shim/drjacoby_gpu_shim.c): the
 * same inputs as the nested `args` list built at R/main.R:348-386 and unpacked by System::load
 * (src/System.cpp:7-45), flattened to plain C arrays, and the same nine outputs as the list
 * returned at src/main.cpp:268-276.  One difference in granularity: ALL chains of a run_mcmc()
 * call go down in ONE call (the reference loops chains in R, R/main.R:392-402).
 *
 * Plain C: pointers and sizes only, no C++/torch types.  All buffers are caller-owned host memory
 * unless a comment says "device".  Every function returns DRJ_OK (0) or a DRJ_ERR_* code and, on
 * error, fills *err (if non-NULL).  The library never falls back to a CPU path: without a CUDA
 * device every compute entry point returns DRJ_ERR_CUDA.
 */
#ifndef DRJACOBY_B200_H
#define DRJACOBY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRJ_ABI_VERSION 2

/* status codes; 1..4 mirror the reference's Rcpp::stop() sites */
enum {
  DRJ_OK = 0,
  DRJ_ERR_INIT_NEGINF = 1,  /* src/Particle.h:75-78  "Starting values result in -Inf ..." */
  DRJ_ERR_LOGLIKE_BAD = 2,  /* src/Particle.h:121-124 "NA, NaN or Inf in likelihood" */
  DRJ_ERR_LOGPRIOR_BAD = 3, /* src/Particle.h:125-128 "NA, NaN or Inf in prior" */
  DRJ_ERR_TRANS_TYPE = 4,   /* src/Particle.cpp:71,95,120 "trans_type invalid" */
  DRJ_ERR_ARG = 10,         /* bad argument (the checks of R/main.R:236-283 re-done defensively) */
  DRJ_ERR_MODEL = 11,       /* unknown built-in model / model does not match config */
  DRJ_ERR_COMPILE = 12,     /* NVRTC failed; err->message holds the start of the log */
  DRJ_ERR_CUDA = 13,        /* CUDA runtime error / no device */
  DRJ_ERR_MEMORY = 14,      /* outputs do not fit in device memory: lower iterations or store flags */
  DRJ_ERR_INTERRUPT = 15    /* the progress callback asked to stop (Rcpp::checkUserInterrupt, main.cpp:156) */
};

/* Flattened named list of numeric vectors: the reference's `data` (args_params$x) and `misc`
 * Rcpp::List arguments (src/System.cpp:15-18).  Item k is values[offsets[k] .. offsets[k]+lengths[k]). */
typedef struct drj_list {
  int32_t n_items;
  const char* const* names; /* [n_items], may be NULL */
  const int64_t* offsets;   /* [n_items] */
  const int64_t* lengths;   /* [n_items] */
  const double* values;
} drj_list;

enum { DRJ_RNG_PHILOX = 0, DRJ_RNG_REPLAY = 1 };

/* Everything System holds (src/System.h:10-53), batched over chains. */
typedef struct drj_config {
  int32_t abi_version;       /* DRJ_ABI_VERSION */
  int32_t device;            /* CUDA device ordinal */
  int32_t d;                 /* number of parameters (rows of df_params) */
  int32_t n_block;           /* max(unlist(df_params$block)), R/main.R:356 */
  int32_t chains;            /* chains in THIS call (a shard when several GPUs are used) */
  int32_t rungs;
  int32_t burnin;
  int32_t samples;
  int32_t coupling_on;
  int32_t save_hot_draws;    /* theta of all rungs (1) or of the cold rung only (0) */
  int32_t store_pt;          /* loglike/logprior of all rungs (1 = reference behaviour) or cold rung only (0) */
  int32_t rng_mode;          /* DRJ_RNG_* */
  double target_acceptance;
  const double* theta_min;   /* [d] */
  const double* theta_max;   /* [d] */
  const int32_t* trans_type; /* [d] 0..3, R/main.R:293 */
  const uint8_t* skip_param; /* [d] min == max, R/main.R:296 */
  const int32_t* block_ptr;  /* [d+1] CSR over df_params$block */
  const int32_t* block_idx;  /* [block_ptr[d]] 1-based block ids */
  const double* beta_raised; /* [rungs] beta_manual^alpha, ascending, last == 1 (R/main.R:268-274,335) */
  const double* theta_init;  /* [chains][d]  (init_mat of R/main.R:322-328, transposed) */
  uint64_t seed;             /* Philox key */
  int64_t chain_offset;      /* global id of chain 0 of this call: results do not depend on sharding */
  const double* replay_uniforms; /* DRJ_RNG_REPLAY: [chains][burnin-1+samples][U],
                                    U = 3*d_free*rungs + (coupling_on && rungs>1 ? rungs-1 : 0);
                                    the reference's own consumption order */
  int32_t iters_per_launch;  /* 0 = choose; iterations advanced by one kernel launch */
  int32_t chains_per_cta;    /* 0 = choose */
  int32_t flags;             /* DRJ_FLAG_* */
  int32_t layout_chains;     /* chains of the WHOLE run when this call is one shard of it (0 = `chains`): the kernel family
                                and the reduction layout of the few-chains kernels are chosen from it, so that a shard
                                reproduces its slice of the unsharded run bit for bit.  The multi-device entry points set it. */
} drj_config;

#define DRJ_FLAG_NO_TMA 1    /* stage data tiles with plain loads instead of the TMA bulk-copy engine */
#define DRJ_FLAG_STREAM_KERNEL 2   /* force the sweep-kernel variant that parks particle state in HBM across data sweeps */
#define DRJ_FLAG_RESIDENT_KERNEL 4 /* force the variant that keeps particle state in registers (default unless data is streamed) */
#define DRJ_FLAG_TEAM_KERNEL 8      /* force the cluster-per-chain kernels (few chains x long data; datum-form models, rungs <= 64) */
#define DRJ_FLAG_NO_TEAM_KERNEL 16  /* never use the cluster-per-chain kernels */
#define DRJ_FLAG_GRID_KERNEL 32     /* force the grid-per-run kernels (chains x rungs <= 256 over data far larger than L2: every SM
                                       streams 1/n_SM of the item once per sweep for all particles; datum-form models) */
#define DRJ_FLAG_NO_GRID_KERNEL 64  /* never use the grid-per-run kernels */
#define DRJ_FLAG_LANES_KERNEL 128   /* force the lanes kernel (a group of 8 or 32 lanes per particle; per-datum models and generic models with a
                                       lane form, rungs <= 32): many parameters, or few particles over short data */
#define DRJ_FLAG_NO_LANES_KERNEL 256 /* never use the lanes kernel */
#define DRJ_FLAG_SPEC_KERNEL 512    /* force the speculative kernel (one warp per chain, five updates in flight: one rung, one block,
                                       d <= 4, data <= 4096 points); results are bitwise those of one thread per particle */
#define DRJ_FLAG_NO_SPEC_KERNEL 1024 /* never use the speculative kernel */

/* A model = the user's loglike + logprior (inst/templates/cpp_template.cpp:4-18).  Either a
 * built-in model (builtin != NULL) or CUDA source compiled at run time with NVRTC for sm_100a:
 *   __device__ double <loglike_name>(const double* params, const DrjList& data, const DrjList& misc, int block);
 *   __device__ double <logprior_name>(const double* params, const DrjList& misc);
 * with P_<name>, DATA_<name>, MISC_<name> index constants generated from the name arrays.
 * R closures cannot run on the GPU and are rejected by the host wrappers. */
typedef struct drj_model_desc {
  const char* builtin;            /* e.g. "example"; NULL for a source model */
  const char* cuda_source;
  const char* loglike_name;
  const char* logprior_name;
  const char* const* param_names; /* [d] */
  const char* const* data_names;  /* [n data items] */
  const char* const* misc_names;  /* [n misc items] */
  int32_t n_data_items;
  int32_t n_misc_items;
  int32_t d;
  int32_t n_block;
} drj_model_desc;

/* The nine outputs of src/main.cpp:268-276 for every chain, chain-major, C-contiguous; plus the
 * accept counts the reference only prints (main.cpp:195,255) and the final bandwidths. */
typedef struct drj_output {
  double* loglike_burnin;      /* [chains][pt_rungs][burnin]        pt_rungs = store_pt ? rungs : 1 */
  double* logprior_burnin;     /* [chains][pt_rungs][burnin] */
  double* theta_burnin;        /* [chains][theta_rungs][burnin][d]  theta_rungs = save_hot_draws ? rungs : 1 */
  double* loglike_sampling;    /* [chains][pt_rungs][samples] */
  double* logprior_sampling;   /* [chains][pt_rungs][samples] */
  double* theta_sampling;      /* [chains][theta_rungs][samples][d] */
  int32_t* mc_accept_burnin;   /* [chains][rungs-1] raw counts */
  int32_t* mc_accept_sampling; /* [chains][rungs-1] */
  int32_t* accept_burnin;      /* [chains][rungs] Particle::accept_count at the end of burn-in; may be NULL */
  int32_t* accept_sampling;    /* [chains][rungs] ... at the end of sampling; may be NULL */
  double* t_diff;              /* [chains] seconds (whole-call wall time / chains); may be NULL */
  double* bw_final;            /* [chains][rungs][d] final proposal bandwidths; may be NULL */
} drj_output;

/* On-device summaries of the stored cold-rung sampling draws: what R/main.R:498-554 computes from the
 * full data frames (Rhat from per-chain means/variances, ESS from the autocovariances of the
 * concatenated chains, DIC from the deviance moments), without moving the draws to the host. */
typedef struct drj_summary {
  int32_t max_lag;           /* in: autocovariance lags 0..max_lag (coda uses floor(10 log10 N)) */
  int32_t n_quantiles;       /* in: number of probabilities in quantile_probs (0 = none) */
  const double* center;      /* in: [d] centre of the autocovariance sums; NULL = mean over this call's chains */
  double* chain_mean;        /* out [chains][d]; may be NULL */
  double* chain_var;         /* out [chains][d] sample variance (n-1); may be NULL */
  double* autocov_sum;       /* out [d][max_lag+1]: sum_j (x_j - c)(x_{j+l} - c), chains concatenated; may be NULL */
  double* deviance_mean_var; /* out [2]: mean and sample variance of -2 loglike; may be NULL */
  double* head;              /* out [d][max_lag]: the first max_lag draws of the concatenated series; may be NULL */
  double* tail;              /* out [d][max_lag]: the last max_lag draws (with head: lets the caller add the lag
                                products that straddle two shards when chains are split over GPUs); may be NULL */
  const double* quantile_probs; /* in [n_quantiles], each in [0,1]; may be NULL */
  double* quantiles;         /* out [d][n_quantiles]: R's quantile(x, probs) (type 7) of every parameter's cold-rung sampling
                                draws, all chains together -- exact order statistics found by a radix select on the device
                                (what summary() / the credible-interval plots compute from mcmc$output); may be NULL */
} drj_summary;

typedef struct drj_error {
  int32_t code;
  int32_t rung;      /* 0-based */
  int32_t param;     /* 0-based, -1 at init */
  int32_t iteration; /* global update index, 0 at init */
  int64_t chain;     /* global chain id */
  double theta[64];  /* offending proposed theta (first min(d,64) entries) */
  char message[512];
} drj_error;

/* progress / interrupt hook (update_progress + Rcpp::checkUserInterrupt, main.cpp:156,181-189);
 * called on the calling thread between kernel batches; return non-zero to stop the run */
typedef int (*drj_progress_fn)(void* user, int phase /*0 burn-in, 1 sampling*/, int done, int total);

typedef struct drj_model drj_model;     /* opaque */
typedef struct drj_session drj_session; /* opaque */

/* ---- library */
int drj_abi_version(void);
int drj_device_count(void);
/* Device buffers of destroyed sessions are kept in a process-wide pool (at most DRJ_POOL_GB GB, default 48; 0 turns
 * it off) and reused by the next session of the same shape; the pool is emptied when an allocation fails, or here.
 * Returns the bytes given back to the driver.  (No reference counterpart: R's allocator is the host's.) */
int64_t drj_release_cached_memory(void);
int drj_builtin_model_count(void);
const char* drj_builtin_model_name(int i);

/* ---- models */
int drj_model_create(const drj_model_desc* desc, int device, drj_model** out, drj_error* err);
void drj_model_destroy(drj_model* m);
double drj_model_compile_seconds(const drj_model* m); /* NVRTC time, 0 for built-ins / cache hits */

/* ---- the drop-in entry: replaces `chains` calls of _drjacoby_main_cpp (src/RcppExports.cpp:15-24) */
int drj_run_mcmc(const drj_config* cfg, const drj_model* model, const drj_list* data, const drj_list* misc,
                 drj_output* out, drj_progress_fn progress, void* progress_user, drj_error* err);

/* ---- session API: the same run in pieces, state resident in HBM between calls
 * (progress bars / interrupts between batches; benchmarking with inputs already on the device) */
int drj_session_create(const drj_config* cfg, const drj_model* model, const drj_list* data, const drj_list* misc,
                       drj_session** out, drj_error* err);
/* advance by n_iter updates inside the current phase; *kernel_ms (may be NULL) receives the CUDA-event
 * time of the sweep kernels of this call, measured on the session's stream */
int drj_session_advance(drj_session* s, int n_iter, float* kernel_ms, drj_error* err);
int drj_session_iterations_done(const drj_session* s); /* rows stored so far, incl. the initial row */
int drj_session_download(drj_session* s, drj_output* out, drj_error* err); /* NULL members of *out are skipped */
/* The same with every `thin`-th stored iteration only (0, thin, 2 thin, ...): the draw arrays of *out hold
 * ceil(burnin / thin) resp. ceil(samples / thin) rows per series, the counters are complete.  For runs whose
 * draws are too large to move whole ("summary / thinned modes"); the reference has no thinning
 * inside run_mcmc, its sample_chains() subsamples afterwards (R/samples.R). */
int drj_session_download_thinned(drj_session* s, int thin, drj_output* out, drj_error* err);
int drj_session_summary(drj_session* s, drj_summary* summary, drj_error* err);
int drj_session_launches(const drj_session* s, int64_t* sweep_launches, int64_t* aux_launches);
/* cumulative CUDA-event time on the session's stream: sweep kernels only / sweep kernels + output transposes */
int drj_session_timing(const drj_session* s, double* sweep_ms_total, double* stream_ms_total);
void drj_session_destroy(drj_session* s);

/* ---- one process, several GPUs: replaces the reference's only parallelism, the dispatch of chains over a PSOCK
 * cluster (R/main.R:392-402: clusterApplyLB(cl = cluster, 1:chains, deploy_chain, ...)).  Chains are split into
 * contiguous ranges, one per device; each device is driven by its own host thread (which never calls back into
 * the caller: the progress hook runs on the calling thread only, as the R API requires, SURVEY 8b "Threading");
 * every shard writes straight into its slice of the caller's chain-major buffers; the kernel family and the
 * reduction layout are fixed from the whole run's size and the Philox counters carry global chain ids, so the
 * result is bit for bit the one-device result.  devices == NULL or n_devices <= 0: all visible devices; devices
 * beyond the number of chains stay idle; an ordinal listed twice gets two chain ranges, run side by side.
 * cfg->device is ignored, cfg->chain_offset is added to every shard's offset. */
typedef struct drj_multi drj_multi; /* opaque: one session per device */
int drj_run_mcmc_multi(const drj_config* cfg, const int32_t* devices, int32_t n_devices, const drj_model* model,
                       const drj_list* data, const drj_list* misc, drj_output* out, drj_progress_fn progress,
                       void* progress_user, drj_error* err);
int drj_multi_create(const drj_config* cfg, const int32_t* devices, int32_t n_devices, const drj_model* model,
                     const drj_list* data, const drj_list* misc, drj_multi** out, drj_error* err);
int drj_multi_device_count(const drj_multi* m); /* devices actually used */
/* *kernel_ms (may be NULL): CUDA-event time of the sweep kernels of this call, max over the devices */
int drj_multi_advance(drj_multi* m, int n_iter, float* kernel_ms, drj_error* err);
int drj_multi_download_thinned(drj_multi* m, int thin, drj_output* out, drj_error* err); /* thin = 1: everything */
/* drj_session_summary over ALL chains: per-chain moments from every device, the autocovariance sums about the
 * global centre added over the devices plus the lag products that straddle two devices' chain ranges, pooled
 * deviance moments, quantiles by a radix select whose histograms are added over the devices; merged on the host
 * (a few KB per device).  head / tail refer to the whole concatenated series. */
int drj_multi_summary(drj_multi* m, drj_summary* summary, drj_error* err);
int drj_multi_launches(const drj_multi* m, int64_t* sweep_launches, int64_t* aux_launches); /* summed over devices */
void drj_multi_destroy(drj_multi* m);

/* ---- measurement helper: FP64 FMA throughput of `device` in TFLOP/s (2 flop per DFMA), the
 * roofline denominator for the compute-bound configs (MEASURED_PEAKS.json has no FP64 figure) */
int drj_fp64_peak_tflops(int device, double* tflops, drj_error* err);

#ifdef __cplusplus
}
#endif
#endif
