// drj_host.cu -- host side of libdrjacoby_b200.so: the C ABI of include/drjacoby_b200.h.
//
// Replaces (reference): the .Call entry _drjacoby_main_cpp (src/RcppExports.cpp:15-24), System::load
// (src/System.cpp:7-45), and the driver part of run_mcmc<T1,T2> (src/main.cpp:84-278): allocation of
// the output stores (:118-124), the burn-in / sampling loops (:153-251), the accept-count reset
// between phases (:208-210), and the returned list (:268-276) -- for ALL chains of a run at once.
// The per-particle work (Particle::update, coupling) is in drj_sampler.cuh.
//
// There is deliberately no CPU fallback: every compute entry point fails with DRJ_ERR_CUDA when no
// device is present.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/drjacoby_b200.h"
#include "drj_models.cuh"

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
static int set_err(drj_error* e, int code, const char* fmt, ...) {
  if (e) {
    e->code = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(e->message, sizeof(e->message), fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CUDA_TRY(call)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return set_err(err, DRJ_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(_e), \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                            \
  } while (0)

static double wall_now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------------------------------------
// auxiliary kernels
// ------------------------------------------------------------------------------------------------
// dst[n * dst_stride + dst_off + c] = src[c * N + n]   (ring [c][n]  ->  per-series [n][c])
__global__ void drj_transpose_rows(const double* __restrict__ src, long long N, int ncols, double* __restrict__ dst,
                                   long long dst_stride, long long dst_off) {
  __shared__ double tile[32][33];
  const long long n0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j;
    const long long n = n0 + threadIdx.x;
    if (c < ncols && n < N) tile[j][threadIdx.x] = src[(long long)c * N + n];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long n = n0 + j;
    const int c = c0 + threadIdx.x;
    if (c < ncols && n < N) dst[n * dst_stride + dst_off + c] = tile[threadIdx.x][j];
  }
}

// bw_final[p][i] = exp(logbw[i][p])
__global__ void drj_export_bw(const double* __restrict__ logbw, long long P, int D, double* __restrict__ out) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  for (int i = 0; i < D; ++i) out[p * D + i] = exp(logbw[(long long)i * P + p]);
}

// ---- on-device summaries of the stored draws (R/main.R:498-554 without moving the draws)
// series of chain c: x(c, t) = src[c * cs + t * ts], t < S
__device__ __forceinline__ double drj_block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];  // fixed order: deterministic
  __syncthreads();
  if (threadIdx.x == 0) sh[0] = t;
  __syncthreads();
  t = sh[0];
  return t;
}

// one CTA per (chain, series): mean and sample variance in ONE pass over the chain's S draws, as sums of
// d = x - x_0 and d^2 (shifted-data algorithm: the pivot is a draw of the same chain, so |mean - pivot| is a
// few standard deviations and the cancellation in sum d^2 - (sum d)^2 / S costs a few ulp), four independent
// loads in flight per thread
__global__ void __launch_bounds__(256) drj_chain_moments(const double* __restrict__ src, long long cs, long long ts, long long ks,
                                                        int S, int K, double* __restrict__ mean, double* __restrict__ var) {
  __shared__ double sh[8];
  const int c = blockIdx.x, k = blockIdx.y;
  const double* x = src + (long long)c * cs + (long long)k * ks;
  const double pivot = x[0];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
  const int B = blockDim.x;
  int t = threadIdx.x;
  for (; t + 3 * B < S; t += 4 * B) {
    const double d0 = x[(long long)t * ts] - pivot, d1 = x[(long long)(t + B) * ts] - pivot;
    const double d2 = x[(long long)(t + 2 * B) * ts] - pivot, d3 = x[(long long)(t + 3 * B) * ts] - pivot;
    a0 += d0; a1 += d1; a2 += d2; a3 += d3;
    q0 = fma(d0, d0, q0); q1 = fma(d1, d1, q1); q2 = fma(d2, d2, q2); q3 = fma(d3, d3, q3);
  }
  for (; t < S; t += B) { const double d = x[(long long)t * ts] - pivot; a0 += d; q0 = fma(d, d, q0); }
  const double sd = drj_block_sum((a0 + a1) + (a2 + a3), sh);
  const double sq = drj_block_sum((q0 + q1) + (q2 + q3), sh);
  if (threadIdx.x == 0) {
    const double md = sd / (double)S;
    mean[(long long)c * K + k] = pivot + md;
    var[(long long)c * K + k] = S > 1 ? fmax(sq - sd * md, 0.0) / (double)(S - 1) : 0.0;
  }
}

// partial sums of (x_j - ctr)(x_{j+l} - ctr), l = l0 .. l0+15, over the chains CONCATENATED in chain order
// (coda::effectiveSize is applied to the concatenated sampling draws, R/main.R:521-526).
// grid = (parameters x lag chunks of 16, nb).  A CTA walks tiles of DRJ_AC_TILE consecutive elements of the
// concatenated series: it stages x_j - ctr of the tile and of the tile shifted by l0 (+ halo) in shared memory
// -- ONE run of TILE + l0 + HALO elements when l0 <= TILE; elements past the end are staged as 0, so the
// inner loop needs no bounds -- and every thread takes EIGHT consecutive j against the 24-value lag window
// they share: 4 + 12 LDS.128 for 128 DFMA into 16 accumulators (with pairs of j it was 10 LDS.128 per 32 DFMA
// and shared-memory bound).  Layout in shared memory: 8 values, 2 pad (80-byte groups), so that the 16-byte
// loads of the 8 lanes of a quarter-warp, 80 bytes apart, hit distinct banks.
// History: v1 read every operand from global memory (17 loads per 16 FMA, a 64-bit division per element):
// 205 ms for 6.5e8 draws x 2 parameters x 89 lags, 60 % of the wall time of the ESS run; v2 (pairs) 45 ms.
#define DRJ_LAGCH 16
#define DRJ_AC_TILE 2048
#define DRJ_AC_HALO 32
#define DRJ_AC_PHYS(i) ((i) + 2 * ((i) >> 3))
__device__ __forceinline__ void drj_ac_stage(double* __restrict__ dst, int count, const double* __restrict__ src, long long cs,
                                             long long ts, int S, long long N, long long j_first, double ctr) {
  // dst[phys(i)] = x_{j_first + i} - ctr (0 past the end); one division per thread, then a carried (chain, t) walk.
  // Eight loads are issued before the first store: one load in flight per warp left the kernel latency-bound.
  long long j = j_first + threadIdx.x;
  long long c = j / S;
  int t = (int)(j - c * S);
  const int step_c = (int)(blockDim.x / S), step_t = (int)(blockDim.x - (long long)step_c * S);
  for (int i0 = threadIdx.x; i0 < count; i0 += 8 * blockDim.x) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      v[u] = (j < N && i0 + u * (int)blockDim.x < count) ? src[c * cs + (long long)t * ts] : ctr;
      j += blockDim.x;
      c += step_c;
      t += step_t;
      if (t >= S) { t -= S; ++c; }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * (int)blockDim.x;
      if (i < count) dst[DRJ_AC_PHYS(i)] = v[u] - ctr;
    }
  }
}
__global__ void __launch_bounds__(256) drj_autocov_partial(const double* __restrict__ src, long long cs, long long ts, int S,
                                                          long long N, const double* __restrict__ ctrs, long long ntiles, int nchunk,
                                                          double* __restrict__ partial /*[param][chunk][gridDim.y][DRJ_LAGCH]*/) {
  __shared__ double sh[8];
  __shared__ __align__(16) double xs[DRJ_AC_PHYS(2 * DRJ_AC_TILE + DRJ_AC_HALO) + 2];
  // blockIdx.x = (parameter, lag chunk), blockIdx.y = tile walker: the CTAs that stage the SAME tiles (one per
  // chunk and parameter) are neighbours in launch order and walk in step, so the tile comes from HBM once and
  // from L2 for the others
  const int chunk = blockIdx.x % nchunk, k = blockIdx.x / nchunk;
  const int l0 = chunk * DRJ_LAGCH;
  const double ctr = ctrs[k];
  const double* x = src + k;
  const bool overlap = l0 <= DRJ_AC_TILE;
  const int shift = overlap ? l0 : DRJ_AC_TILE;  // logical index of the shifted tile inside xs (a multiple of 8)
  double acc[DRJ_LAGCH];
#pragma unroll
  for (int q = 0; q < DRJ_LAGCH; ++q) acc[q] = 0.0;
  for (long long tile = blockIdx.y; tile < ntiles; tile += gridDim.y) {
    const long long j0 = tile * DRJ_AC_TILE;
    __syncthreads();
    if (overlap) drj_ac_stage(xs, DRJ_AC_TILE + l0 + DRJ_AC_HALO, x, cs, ts, S, N, j0, ctr);
    else {
      drj_ac_stage(xs, DRJ_AC_TILE, x, cs, ts, S, N, j0, ctr);
      drj_ac_stage(xs + DRJ_AC_PHYS(DRJ_AC_TILE), DRJ_AC_TILE + DRJ_AC_HALO, x, cs, ts, S, N, j0 + l0, ctr);
    }
    __syncthreads();
    for (int it = threadIdx.x; it < DRJ_AC_TILE / 8; it += blockDim.x) {
      const double2* pa = reinterpret_cast<const double2*>(&xs[10 * it]);
      const double2* pw = reinterpret_cast<const double2*>(&xs[10 * (it + (shift >> 3))]);
      double xj[8], w[24];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const double2 v = pa[u]; xj[2 * u] = v.x; xj[2 * u + 1] = v.y; }
#pragma unroll
      for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int u = 0; u < 4; ++u) { const double2 v = pw[5 * g + u]; w[8 * g + 2 * u] = v.x; w[8 * g + 2 * u + 1] = v.y; }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
#pragma unroll
        for (int q = 0; q < DRJ_LAGCH; ++q) acc[q] = fma(xj[jj], w[jj + q], acc[q]);
    }
  }
  double* out = partial + (((long long)k * nchunk + chunk) * gridDim.y + blockIdx.y) * DRJ_LAGCH;
#pragma unroll
  for (int q = 0; q < DRJ_LAGCH; ++q) {
    const double t = drj_block_sum(acc[q], sh);
    if (threadIdx.x == 0) out[q] = t;
  }
}

// grid = (lag chunks, parameters): out[param][l] = sum over the CTAs of the partial kernel, fixed order
__global__ void drj_reduce_partials(const double* __restrict__ partial, int nblocks, double* __restrict__ out, int max_lag) {
  const int q = threadIdx.x, chunk = blockIdx.x, k = blockIdx.y;
  const int l = chunk * DRJ_LAGCH + q;
  if (q < DRJ_LAGCH && l <= max_lag) {
    const double* p = partial + (((long long)k * gridDim.x + chunk) * nblocks) * DRJ_LAGCH + q;
    double t = 0.0;
    for (int b = 0; b < nblocks; ++b) t += p[(long long)b * DRJ_LAGCH];
    out[(long long)k * (max_lag + 1) + l] = t;
  }
}

// every thin-th row of every series: dst[(q * n_keep + j) * w + c] = src[(q * n + j * thin) * w + c]
__global__ void __launch_bounds__(256) drj_thin_rows(const double* __restrict__ src, double* __restrict__ dst, long long n_series,
                                                    int n, int n_keep, int w, int thin) {
  const long long total = n_series * n_keep * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / w;
    const int c = (int)(i - row * w);
    const long long q = row / n_keep;
    const int j = (int)(row - q * n_keep);
    dst[i] = src[(q * n + (long long)j * thin) * w + c];
  }
}

// the first / last L elements of the concatenated series of every parameter: out[k][i]
__global__ void drj_series_ends(const double* __restrict__ src, long long cs, long long ts, int S, long long N, int D, int L,
                                double* __restrict__ head, double* __restrict__ tail) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  for (int k = 0; k < D; ++k) {
    const long long jh = i, jt = N - L + i;
    head[(long long)k * L + i] = (jh < N) ? src[(jh / S) * cs + (jh % S) * ts + k] : 0.0;
    tail[(long long)k * L + i] = (jt >= 0) ? src[(jt / S) * cs + (jt % S) * ts + k] : 0.0;
  }
}

// ---- quantiles by radix select (R/main.R's consumers call quantile() on mcmc$output; at GPU chain counts the
// draws cannot leave the device).  One pass = one 11-bit digit of the order-preserving 64-bit key of a double:
// every draw of parameter blockIdx.y whose high `bits_done` bits equal the prefix of one of the parameter's groups
// is counted in that group's 2048-bin histogram (shared memory, warp-aggregated: the first digits -- sign and
// exponent -- put nearly all draws of a parameter in two or three bins), then the non-empty bins are added to
// the global histogram.  The host picks the bin holding each wanted rank and descends (6 passes: 11,11,11,11,11,9 bits).
#define DRJ_QBINS 2048
__device__ __forceinline__ unsigned long long drj_order_key(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__global__ void __launch_bounds__(256) drj_quantile_hist(const double* __restrict__ src, long long cs, long long ts, int S, int C,
                                                        int bits_done, int nbits, const unsigned long long* __restrict__ prefix,
                                                        const int* __restrict__ group_ptr /*[D+1]*/, unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int qh[];  // [groups of this parameter][DRJ_QBINS]
  const int k = blockIdx.y;
  const int g0 = group_ptr[k], ng = group_ptr[k + 1] - g0;
  if (ng == 0) return;
  for (int i = threadIdx.x; i < ng * DRJ_QBINS; i += blockDim.x) qh[i] = 0u;
  __syncthreads();
  const int shift = 64 - bits_done - nbits;
  const unsigned mask = (1u << nbits) - 1u;
  const int lane = threadIdx.x & 31;
  for (int c = blockIdx.x; c < C; c += gridDim.x) {
    const double* x = src + (long long)c * cs + k;
    for (int t0 = 0; t0 < S; t0 += blockDim.x) {
      const int t = t0 + threadIdx.x;
      const bool in = t < S;
      const unsigned long long key = in ? drj_order_key(x[(long long)t * ts]) : 0ULL;
      const unsigned bin = (unsigned)(key >> shift) & mask;
      for (int g = 0; g < ng; ++g) {
        const bool hit = in && (bits_done == 0 || (key >> (64 - bits_done)) == prefix[g0 + g]);
        const unsigned peers = __match_any_sync(0xffffffffu, hit ? bin : 0xffffffffu);
        if (hit && lane == __ffs(peers) - 1) atomicAdd(&qh[g * DRJ_QBINS + bin], (unsigned)__popc(peers));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ng * DRJ_QBINS; i += blockDim.x)
    if (qh[i]) atomicAdd(&hist[(long long)g0 * DRJ_QBINS + i], (unsigned long long)qh[i]);
}

// The state-independent part of every update of one launch, for runs too small to fill the machine: in such a run every
// thread is one dependent chain of ~1000 instructions per update, a third of them the counter-based draw (Philox, the
// inverse normal, log u) that does not depend on the chain at all.  This kernel computes them for all iterations of
// the launch at once with the whole GPU (same device functions, same bits), the sweep kernel reads them back.
// Element (it, free slot, particle) of z / lu at ((it * d_free + slot) * P + p); coupling log u at
// ((it * chains + chain) * (R - 1) + pair).  Counters as in the sweep kernel (drj_sampler.cuh).
__global__ void __launch_bounds__(256) drj_draw_kernel(const DrjKernelArgs a, double* __restrict__ z, double* __restrict__ lu,
                                                        double* __restrict__ clu, const int* __restrict__ free_param) {
  const long long P = a.P;
  const int R = a.rungs, DF = a.d_free;
  const unsigned k0 = (unsigned)a.seed, k1 = (unsigned)(a.seed >> 32);
  const long long per_it = P * DF, total = per_it * a.n_iter;
  const long long stride = (long long)gridDim.x * blockDim.x, first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long e = first; e < total; e += stride) {
    const long long it = e / per_it, rem = e - it * per_it;
    const int fs = (int)(rem / P);
    const long long p = rem - (long long)fs * P, chain = p / R;
    const int r = (int)(p - chain * R);
    const int t = a.iter0 + (int)it;
    double u1, u2, u3;
    if (a.rng_mode) {
      const double* q = a.replay + (chain * a.replay_T + (t - 1)) * a.replay_U + 3 * ((long long)r * DF + fs);
      u1 = q[0]; u2 = q[1]; u3 = q[2];
    } else {
      const unsigned long long g = (unsigned long long)(a.chain_offset + chain);
      DrjU4 w = drj_philox4x32_10((unsigned)g, (unsigned)t, ((unsigned)r << 16) | (unsigned)free_param[fs], (unsigned)(g >> 32) << 8, k0, k1);
      u1 = drj_u32_to_unif(w.x); u2 = drj_u32_to_unif(w.y); u3 = drj_u32_to_unif(w.z);
    }
    z[e] = drj_norm_rand(u1, u2);
    lu[e] = log(u3);
  }
  if (clu && R > 1) {
    const long long per_it2 = a.chains * (R - 1), total2 = per_it2 * a.n_iter;
    for (long long e = first; e < total2; e += stride) {
      const long long it = e / per_it2, rem = e - it * per_it2, chain = rem / (R - 1);
      const int r = (int)(rem - chain * (R - 1));
      const int t = a.iter0 + (int)it;
      double u;
      if (a.rng_mode) u = a.replay[(chain * a.replay_T + (t - 1)) * a.replay_U + 3LL * DF * R + r];
      else {
        const unsigned long long g = (unsigned long long)(a.chain_offset + chain);
        DrjU4 w = drj_philox4x32_10((unsigned)g, (unsigned)t, (unsigned)r, 1u | ((unsigned)(g >> 32) << 8), k0, k1);
        u = drj_u32_to_unif(w.x);
      }
      clu[e] = log(u);
    }
  }
}

// FP64 FMA throughput probe: 8 independent dependent-chains per thread
__global__ void __launch_bounds__(256) drj_fp64_probe(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ------------------------------------------------------------------------------------------------
// models: built-in registry + NVRTC
// ------------------------------------------------------------------------------------------------
// a source model's kernels on ONE device (a CUmodule belongs to a context)
struct JitFuncs {
  CUmodule module = nullptr;
  CUfunction init_cu = nullptr, sweep_cu = nullptr, stream_cu = nullptr, team_init_cu = nullptr, team_sweep_cu = nullptr,
             init_chain_cu = nullptr, small_cu = nullptr, grid_init_cu = nullptr, grid_sweep_cu = nullptr, lanes_sweep_cu = nullptr, spec5_cu = nullptr, spec7_cu = nullptr;
};

struct drj_model {
  std::string name;
  int D = 0, NBLOCK = 0;
  bool datum_form = false;
  bool jit = false;
  const void* init_fn = nullptr;   // built-in: __global__ function addresses
  const void* init_chain_fn = nullptr;  // initial likelihood / prior, one thread per chain
  const void* sweep_fn = nullptr;
  const void* small_fn = nullptr;  // sweep variant for runs too small to fill the machine (drj_sweep_kernel_small)
  const void* stream_fn = nullptr; // sweep variant for data streamed through shared memory (datum-form models)
  const void* team_init_fn = nullptr;   // cluster-per-chain variants (datum-form models)
  const void* team_sweep_fn = nullptr;
  const void* grid_init_fn = nullptr;   // grid-per-run variants (datum-form models)
  const void* grid_sweep_fn = nullptr;
  const void* lanes_sweep_fn = nullptr; // a group of lanes per particle (drj_lanes.cuh): per-datum models and generic models with a lane form
  int lanes = 0;                        // lanes per particle of that kernel (0 = the model has none)
  int lanes_bpp = 0;                    // its shared-memory bytes per particle
  const void* spec5_fn = nullptr;       // one CTA per chain, five / seven updates in flight (drj_spec.cuh): one block, d <= 4
  const void* spec7_fn = nullptr;
  bool spec = false;                    // the model has that kernel
  // jit: the compiled cubin, loaded into a device's context on first use there (a model serves every device:
  // drj_run_mcmc_multi drives all GPUs from one model handle)
  std::vector<char> cubin;
  mutable std::map<int, JitFuncs> per_device;
  mutable std::mutex mu;
  double compile_seconds = 0.0;
  int device = 0;
};

struct BuiltinEntry {
  const char* name;
  int D, NBLOCK;
  bool datum_form;
  const void* init_fn;
  const void* sweep_fn;
  const void* stream_fn;
  const void* team_init_fn;
  const void* team_sweep_fn;
  const void* init_chain_fn;
  const void* grid_init_fn;
  const void* grid_sweep_fn;
  const void* lanes_sweep_fn;
  int lanes, lanes_bpp;
  const void* spec5_fn;
  const void* spec7_fn;
  const void* small_fn;
};

#define DRJ_REG(NAME, ...)                                                               \
  {                                                                                      \
    NAME, __VA_ARGS__::D, __VA_ARGS__::NBLOCK, __VA_ARGS__::DATUM_FORM,                  \
        (const void*)drj_init_kernel<__VA_ARGS__>, (const void*)drj_sweep_kernel<__VA_ARGS__, false>, \
        DrjStreamKernel<__VA_ARGS__>::get(), DrjStreamKernel<__VA_ARGS__>::team_init(),                \
        DrjStreamKernel<__VA_ARGS__>::team_sweep(), (const void*)drj_init_chain_kernel<__VA_ARGS__>,   \
        DrjStreamKernel<__VA_ARGS__>::grid_init(), DrjStreamKernel<__VA_ARGS__>::grid_sweep(),         \
        DrjLanesKernel<__VA_ARGS__>::sweep(), DrjLanesKernel<__VA_ARGS__>::lanes,                      \
        DrjLanesKernel<__VA_ARGS__>::bytes_per_particle, DrjSpecKernel<__VA_ARGS__>::sweep5(),         \
        DrjSpecKernel<__VA_ARGS__>::sweep7(), (const void*)drj_sweep_kernel_small<__VA_ARGS__>         \
  }

static const BuiltinEntry BUILTINS[] = {
    DRJ_REG("example", DrjModelExample),
    DRJ_REG("example_exact", DrjModelExampleExact),
    DRJ_REG("coupling", DrjModelCoupling),
    DRJ_REG("normal_mu", DrjModelNormalMu),
    DRJ_REG("doublewell", DrjModelDoubleWell),
    DRJ_REG("multilevel", DrjModelMultilevel<5>),
    DRJ_REG("multilevel", DrjModelMultilevel<50>),
    DRJ_REG("blocks", DrjModelBlocks<6>),
    DRJ_REG("returnprior", DrjModelReturnPrior),
    DRJ_REG("testfile", DrjModelTestFile<true, true>),
    DRJ_REG("testfile_nullprior", DrjModelTestFile<true, false>),
    DRJ_REG("testfile_nulllike", DrjModelTestFile<false, true>),
    DRJ_REG("logistic", DrjModelLogistic),
    DRJ_REG("const", DrjModelConst),
    DRJ_REG("halfdomain_like", DrjModelHalfDomain<true>),
    DRJ_REG("halfdomain_prior", DrjModelHalfDomain<false>),
};
static const int N_BUILTINS = (int)(sizeof(BUILTINS) / sizeof(BUILTINS[0]));

// ---- driver entry points (no link-time dependency on libcuda: the library must load on a CPU box)
struct DriverApi {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;
  CUresult (*LaunchCooperativeKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                      CUstream, void**) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*OccupancyMaxActiveClusters)(int*, CUfunction, const CUlaunchConfig*) = nullptr;
  bool ok = false;
};
static DriverApi g_drv;
static std::mutex g_mutex;

static bool load_driver_api() {
  if (g_drv.ok) return true;
  struct { const char* sym; void** slot; } want[] = {
      {"cuModuleLoadData", (void**)&g_drv.ModuleLoadData},
      {"cuModuleGetFunction", (void**)&g_drv.ModuleGetFunction},
      {"cuModuleUnload", (void**)&g_drv.ModuleUnload},
      {"cuLaunchKernel", (void**)&g_drv.LaunchKernel},
      {"cuGetErrorString", (void**)&g_drv.GetErrorString},
      {"cuLaunchKernelEx", (void**)&g_drv.LaunchKernelEx},
      {"cuLaunchCooperativeKernel", (void**)&g_drv.LaunchCooperativeKernel},
      {"cuOccupancyMaxActiveBlocksPerMultiprocessor", (void**)&g_drv.OccupancyMaxActiveBlocksPerMultiprocessor},
      {"cuFuncSetAttribute", (void**)&g_drv.FuncSetAttribute},
      {"cuOccupancyMaxActiveClusters", (void**)&g_drv.OccupancyMaxActiveClusters},
  };
  for (auto& w : want) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(w.sym, w.slot, cudaEnableDefault, &q) != cudaSuccess || !*w.slot) return false;
  }
  g_drv.ok = true;
  return true;
}

// ---- NVRTC through dlopen (optional at load time, required only for source models)
struct NvrtcApi {
  void* handle = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  const char* (*GetErrorString)(nvrtcResult);
};
static NvrtcApi g_nvrtc;

static bool load_nvrtc(std::string& why) {
  if (g_nvrtc.handle) return true;
  const char* cands[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                         "/usr/local/cuda/lib64/libnvrtc.so"};
  void* h = nullptr;
  for (const char* c : cands) {
    h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  if (!h) { why = "libnvrtc.so.12 not found"; return false; }
#define NV_SYM(field, name)                              \
  *(void**)(&g_nvrtc.field) = dlsym(h, name);            \
  if (!g_nvrtc.field) { why = std::string("missing ") + name; dlclose(h); return false; }
  NV_SYM(CreateProgram, "nvrtcCreateProgram")
  NV_SYM(DestroyProgram, "nvrtcDestroyProgram")
  NV_SYM(CompileProgram, "nvrtcCompileProgram")
  NV_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
  NV_SYM(GetCUBIN, "nvrtcGetCUBIN")
  NV_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
  NV_SYM(GetProgramLog, "nvrtcGetProgramLog")
  NV_SYM(GetErrorString, "nvrtcGetErrorString")
#undef NV_SYM
  g_nvrtc.handle = h;
  return true;
}

// the device headers, embedded at build time (csrc/build.py writes drj_embedded_sources.inc)
static const char* const EMBED_DEVICE =
#include "drj_embedded_device.inc"
    ;
static const char* const EMBED_SAMPLER =
#include "drj_embedded_sampler.inc"
    ;
static const char* const EMBED_MODELS =
#include "drj_embedded_models.inc"
    ;
static const char* const EMBED_LANES =
#include "drj_embedded_lanes.inc"
    ;
static const char* const EMBED_SPEC =
#include "drj_embedded_spec.inc"
    ;

static uint64_t fnv1a(const std::string& s) {
  uint64_t h = 1469598103934665603ULL;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}

static bool valid_ident(const char* s) {
  if (!s || !*s) return false;
  if (!(isalpha((unsigned char)s[0]) || s[0] == '_')) return false;
  for (const char* p = s; *p; ++p)
    if (!(isalnum((unsigned char)*p) || *p == '_')) return false;
  return true;
}

// The cubin cache directory, or "" when no safe one is available (then every source model is compiled afresh).
// $DRJ_CACHE_DIR, else $XDG_CACHE_HOME/drjacoby_b200, else ~/.cache/drjacoby_b200, else /tmp/drjacoby_b200_cache_<uid>.
// The directory must be a real directory (not a symlink) owned by this user and closed to everybody else:
// a cached cubin is code that will run in this process, so a directory another local user could have
// created or can write to is never used.
static std::string cache_dir() {
  std::string d;
  if (const char* e = getenv("DRJ_CACHE_DIR")) d = e;
  else if (const char* x = getenv("XDG_CACHE_HOME")) { if (*x) { mkdir(x, 0700); d = std::string(x) + "/drjacoby_b200"; } }
  if (d.empty()) {
    if (const char* h = getenv("HOME")) {
      if (*h) { const std::string c = std::string(h) + "/.cache"; mkdir(c.c_str(), 0700); d = c + "/drjacoby_b200"; }
    }
  }
  if (d.empty()) d = std::string("/tmp/drjacoby_b200_cache_") + std::to_string((long)getuid());
  for (int attempt = 0; attempt < 2; ++attempt) {
    mkdir(d.c_str(), 0700);
    struct stat st;
    if (lstat(d.c_str(), &st) == 0 && S_ISDIR(st.st_mode) && st.st_uid == getuid() && (st.st_mode & 077) == 0) return d;
    if (attempt == 0 && !getenv("DRJ_CACHE_DIR")) d = std::string("/tmp/drjacoby_b200_cache_") + std::to_string((long)getuid());
    else break;
  }
  return std::string();
}

// second, independent 64-bit hash (different offset basis and a final avalanche): the cache file carries both
// hashes and the length of the key text, and an entry is used only if all three match
static uint64_t fnv1a_alt(const std::string& s) {
  uint64_t h = 0x9ae16a3b2f90404fULL;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; h ^= h >> 29; }
  h ^= h >> 32; h *= 0xd6e8feb86659fd93ULL; h ^= h >> 32;
  return h;
}

// Generates the translation unit for a source model: index constants from the name arrays, the
// user's code, the model struct that adapts it to the sampler contract, and the two kernels.
static std::string make_jit_source(const drj_model_desc* desc) {
  std::string s;
  s += "#include \"drj_models.cuh\"\n";
  for (int i = 0; i < desc->d; ++i)
    if (desc->param_names && valid_ident(desc->param_names[i]))
      s += "constexpr int P_" + std::string(desc->param_names[i]) + " = " + std::to_string(i) + ";\n";
  for (int i = 0; i < desc->n_data_items; ++i)
    if (desc->data_names && valid_ident(desc->data_names[i]))
      s += "constexpr int DATA_" + std::string(desc->data_names[i]) + " = " + std::to_string(i) + ";\n";
  for (int i = 0; i < desc->n_misc_items; ++i)
    if (desc->misc_names && valid_ident(desc->misc_names[i]))
      s += "constexpr int MISC_" + std::string(desc->misc_names[i]) + " = " + std::to_string(i) + ";\n";
  s += "constexpr int DRJ_D = " + std::to_string(desc->d) + ";\n";
  s += "constexpr int DRJ_NBLOCK = " + std::to_string(desc->n_block) + ";\n";
  s += "#line 1 \"user_model.cu\"\n";
  s += desc->cuda_source;
  s += "\n#line 1 \"drj_user_glue.cu\"\n";
  const std::string ll = desc->loglike_name ? desc->loglike_name : "loglike";
  const std::string lp = desc->logprior_name ? desc->logprior_name : "logprior";
  s += "#ifndef DRJ_USER_MODEL\n"
       "struct DrjUserModel {\n"
       "  static constexpr int D = DRJ_D, NBLOCK = DRJ_NBLOCK;\n"
       "  static constexpr bool DATUM_FORM = false;\n"
       "  __device__ static __forceinline__ double loglike(const double* th, const DrjList& data, const DrjList& misc, int block) {\n"
       "    return ::" + ll + "(th, data, misc, block);\n"
       "  }\n"
       "  __device__ static __forceinline__ double logprior(const double* th, const DrjList& misc) {\n"
       "    return ::" + lp + "(th, misc);\n"
       "  }\n"
       "};\n"
       "#define DRJ_USER_MODEL DrjUserModel\n"
       "#endif\n"
       "static_assert(DRJ_USER_MODEL::D == DRJ_D, \"model D does not match df_params\");\n"
       "static_assert(DRJ_USER_MODEL::NBLOCK == DRJ_NBLOCK, \"model NBLOCK does not match df_params$block\");\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT) drj_user_init(const __grid_constant__ DrjKernelArgs a) {\n"
       "  drj_init_body<DRJ_USER_MODEL, DRJ_MODE_SOLO>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT) drj_user_init_chain(const __grid_constant__ DrjKernelArgs a) {\n"
       "  drj_init_chain_body<DRJ_USER_MODEL>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT, 2) drj_user_sweep(const __grid_constant__ DrjKernelArgs a) {\n"
       "  drj_sweep_body<DRJ_USER_MODEL, false, DRJ_MODE_SOLO>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT, 1) drj_user_sweep_small(const __grid_constant__ DrjKernelArgs a) {\n"
       "  drj_sweep_body<DRJ_USER_MODEL, false, DRJ_MODE_SOLO, true>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT, 2) drj_user_sweep_stream(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DRJ_USER_MODEL::DATUM_FORM) drj_sweep_body<DRJ_USER_MODEL, true, DRJ_MODE_SOLO>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_TEAM_NT, 1) drj_user_init_team(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DRJ_USER_MODEL::DATUM_FORM) drj_init_body<DRJ_USER_MODEL, DRJ_MODE_TEAM>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_TEAM_NT, 1) drj_user_sweep_team(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DRJ_USER_MODEL::DATUM_FORM) drj_sweep_body<DRJ_USER_MODEL, false, DRJ_MODE_TEAM>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT, 1) drj_user_init_grid(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DRJ_USER_MODEL::DATUM_FORM) drj_init_body<DRJ_USER_MODEL, DRJ_MODE_GRID>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT, 1) drj_user_sweep_grid(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DRJ_USER_MODEL::DATUM_FORM) drj_sweep_body<DRJ_USER_MODEL, false, DRJ_MODE_GRID>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT, 2) drj_user_sweep_lanes(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DrjLanesTraits<DRJ_USER_MODEL>::ok) drj_lanes_sweep_body<DRJ_USER_MODEL>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(32) drj_user_sweep_spec5(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DrjSpecTraits<DRJ_USER_MODEL>::ok) drj_spec_sweep_body<DRJ_USER_MODEL, 5>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(1 << DRJ_SPEC_DEEP) drj_user_sweep_spec7(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DrjSpecTraits<DRJ_USER_MODEL>::ok) drj_spec_sweep_body<DRJ_USER_MODEL, DRJ_SPEC_DEEP>(a);\n"
       "}\n"
       "extern \"C\" __global__ void drj_user_traits(int* out) {\n"
       "  out[0] = DRJ_USER_MODEL::DATUM_FORM ? 1 : 0;\n"
       "  out[1] = DrjLanesKernel<DRJ_USER_MODEL>::lanes;\n"
       "  out[2] = DrjLanesKernel<DRJ_USER_MODEL>::bytes_per_particle;\n"
       "  out[3] = DrjSpecTraits<DRJ_USER_MODEL>::ok ? 1 : 0;\n"
       "}\n";
  return s;
}

// load the model's cubin into the current device's context (once per device) and look its kernels up
static int jit_functions(const drj_model* m, int device, const JitFuncs** out, drj_error* err) {
  std::lock_guard<std::mutex> lock(m->mu);
  auto it = m->per_device.find(device);
  if (it != m->per_device.end()) { *out = &it->second; return DRJ_OK; }
  if (!load_driver_api()) return set_err(err, DRJ_ERR_CUDA, "CUDA driver entry points unavailable");
  JitFuncs f;
  CUresult cr = g_drv.ModuleLoadData(&f.module, m->cubin.data());
  if (cr != CUDA_SUCCESS) {
    const char* es = nullptr;
    g_drv.GetErrorString(cr, &es);
    return set_err(err, DRJ_ERR_CUDA, "cuModuleLoadData: %s", es ? es : "?");
  }
  if (g_drv.ModuleGetFunction(&f.init_cu, f.module, "drj_user_init") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.sweep_cu, f.module, "drj_user_sweep") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.stream_cu, f.module, "drj_user_sweep_stream") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.small_cu, f.module, "drj_user_sweep_small") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.team_init_cu, f.module, "drj_user_init_team") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.team_sweep_cu, f.module, "drj_user_sweep_team") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.grid_init_cu, f.module, "drj_user_init_grid") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.grid_sweep_cu, f.module, "drj_user_sweep_grid") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.lanes_sweep_cu, f.module, "drj_user_sweep_lanes") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.spec5_cu, f.module, "drj_user_sweep_spec5") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.spec7_cu, f.module, "drj_user_sweep_spec7") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&f.init_chain_cu, f.module, "drj_user_init_chain") != CUDA_SUCCESS) {
    g_drv.ModuleUnload(f.module);
    return set_err(err, DRJ_ERR_COMPILE, "kernels not found in the compiled module");
  }
  *out = &(m->per_device[device] = f);
  return DRJ_OK;
}

static int jit_compile(const drj_model_desc* desc, drj_model* m, drj_error* err) {
  const double t0 = wall_now();
  const std::string src = make_jit_source(desc);
  const std::string key_text = src + "|" + EMBED_DEVICE + "|" + EMBED_SAMPLER + "|" + EMBED_MODELS + "|" + EMBED_LANES + "|" + EMBED_SPEC + "|sm_100a|v9";
  const uint64_t h1 = fnv1a(key_text), h2 = fnv1a_alt(key_text), klen = (uint64_t)key_text.size();
  char hex[40];
  snprintf(hex, sizeof(hex), "%016llx%08llx", (unsigned long long)h1, (unsigned long long)(h2 & 0xffffffffULL));
  const std::string dir = cache_dir();
  const std::string path = dir.empty() ? std::string() : dir + "/" + hex + ".cubin";
  std::vector<char>& cubin = m->cubin;
  cubin.clear();
  // cache file = { magic, h1, h2, key length } + cubin; anything that does not match is ignored and rewritten
  const uint64_t magic = 0x324e494255434a44ULL;  // "DJCUBIN2"
  if (!path.empty())
    if (FILE* f = fopen(path.c_str(), "rb")) {
      uint64_t hdr[4] = {0, 0, 0, 0};
      struct stat st;
      if (fstat(fileno(f), &st) == 0 && st.st_uid == getuid() && (st.st_mode & 022) == 0 && st.st_size > (off_t)sizeof(hdr) &&
          fread(hdr, sizeof(hdr), 1, f) == 1 && hdr[0] == magic && hdr[1] == h1 && hdr[2] == h2 && hdr[3] == klen) {
        cubin.resize((size_t)st.st_size - sizeof(hdr));
        if (fread(cubin.data(), 1, cubin.size(), f) != cubin.size()) cubin.clear();
      }
      fclose(f);
    }
  if (cubin.empty()) {
    std::string why;
    if (!load_nvrtc(why)) return set_err(err, DRJ_ERR_COMPILE, "NVRTC unavailable: %s", why.c_str());
    nvrtcProgram prog;
    const char* hdr_src[] = {EMBED_DEVICE, EMBED_SAMPLER, EMBED_MODELS, EMBED_LANES, EMBED_SPEC};
    const char* hdr_name[] = {"drj_device.cuh", "drj_sampler.cuh", "drj_models.cuh", "drj_lanes.cuh", "drj_spec.cuh"};
    nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.c_str(), "drj_jit_model.cu", 5, hdr_src, hdr_name);
    if (r != NVRTC_SUCCESS) return set_err(err, DRJ_ERR_COMPILE, "nvrtcCreateProgram: %s", g_nvrtc.GetErrorString(r));
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"};
    r = g_nvrtc.CompileProgram(prog, 4, opts);
    if (r != NVRTC_SUCCESS) {
      size_t n = 0;
      g_nvrtc.GetProgramLogSize(prog, &n);
      std::string log(n, '\0');
      if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
      g_nvrtc.DestroyProgram(&prog);
      return set_err(err, DRJ_ERR_COMPILE, "NVRTC compile failed: %.440s", log.c_str());
    }
    size_t n = 0;
    g_nvrtc.GetCUBINSize(prog, &n);
    cubin.resize(n);
    g_nvrtc.GetCUBIN(prog, cubin.data());
    g_nvrtc.DestroyProgram(&prog);
    m->compile_seconds = wall_now() - t0;
    if (!path.empty()) {
      const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
      if (FILE* f = fopen(tmp.c_str(), "wb")) {
        const uint64_t hdr[4] = {magic, h1, h2, klen};
        const bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1 && fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
        fclose(f);
        if (ok) { chmod(tmp.c_str(), 0600); rename(tmp.c_str(), path.c_str()); }
        else unlink(tmp.c_str());
      }
    }
  }
  const JitFuncs* jf = nullptr;
  int rc = jit_functions(m, m->device, &jf, err);
  if (rc) return rc;
  {  // does the user's model use the per-datum form?  does it have a lanes kernel?
    CUfunction traits = nullptr;
    int* d_flag = nullptr;
    int h_flag[4] = {0, 0, 0, 0};
    if (g_drv.ModuleGetFunction(&traits, jf->module, "drj_user_traits") == CUDA_SUCCESS &&
        cudaMalloc(&d_flag, sizeof(h_flag)) == cudaSuccess) {
      void* params[] = {(void*)&d_flag};
      if (g_drv.LaunchKernel(traits, 1, 1, 1, 1, 1, 1, 0, nullptr, params, nullptr) == CUDA_SUCCESS &&
          cudaMemcpy(h_flag, d_flag, sizeof(h_flag), cudaMemcpyDeviceToHost) == cudaSuccess) {
        m->datum_form = h_flag[0] != 0;
        m->lanes = h_flag[1];
        m->lanes_bpp = h_flag[2];
        m->spec = h_flag[3] != 0;
      }
      cudaFree(d_flag);
    }
  }
  return DRJ_OK;
}

extern "C" int drj_abi_version(void) { return DRJ_ABI_VERSION; }

extern "C" int drj_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int drj_builtin_model_count(void) { return N_BUILTINS; }
extern "C" const char* drj_builtin_model_name(int i) { return (i >= 0 && i < N_BUILTINS) ? BUILTINS[i].name : nullptr; }

extern "C" int drj_model_create(const drj_model_desc* desc, int device, drj_model** out, drj_error* err) {
  if (!desc || !out) return set_err(err, DRJ_ERR_ARG, "drj_model_create: NULL argument");
  *out = nullptr;
  if (desc->builtin) {
    for (int i = 0; i < N_BUILTINS; ++i) {
      const BuiltinEntry& b = BUILTINS[i];
      if (strcmp(b.name, desc->builtin) != 0) continue;
      if (desc->d > 0 && (b.D != desc->d || (desc->n_block > 0 && b.NBLOCK != desc->n_block))) continue;
      drj_model* m = new drj_model();
      m->name = b.name; m->D = b.D; m->NBLOCK = b.NBLOCK; m->datum_form = b.datum_form;
      m->init_fn = b.init_fn; m->sweep_fn = b.sweep_fn; m->stream_fn = b.stream_fn; m->device = device;
      m->team_init_fn = b.team_init_fn; m->team_sweep_fn = b.team_sweep_fn; m->init_chain_fn = b.init_chain_fn;
      m->grid_init_fn = b.grid_init_fn; m->grid_sweep_fn = b.grid_sweep_fn;
      m->lanes_sweep_fn = b.lanes_sweep_fn; m->lanes = b.lanes; m->lanes_bpp = b.lanes_bpp;
      m->spec5_fn = b.spec5_fn; m->spec7_fn = b.spec7_fn; m->spec = b.spec5_fn != nullptr;
      m->small_fn = b.small_fn;
      *out = m;
      return DRJ_OK;
    }
    // a sized family (multilevel<G>, blocks<G>) not pre-instantiated: specialise it with NVRTC
    std::string fam = desc->builtin;
    std::string glue;
    if (fam == "multilevel" && desc->d > 0)
      glue = "#define DRJ_USER_MODEL DrjModelMultilevel<" + std::to_string(desc->d) + ">\n";
    else if (fam == "blocks" && desc->d > 3)
      glue = "#define DRJ_USER_MODEL DrjModelBlocks<" + std::to_string(desc->d - 3) + ">\n";
    else
      return set_err(err, DRJ_ERR_MODEL, "unknown built-in model '%s' (d=%d, n_block=%d)", desc->builtin, desc->d, desc->n_block);
    drj_model_desc jd = *desc;
    jd.builtin = nullptr;
    jd.cuda_source = glue.c_str();
    jd.param_names = nullptr; jd.data_names = nullptr; jd.misc_names = nullptr;
    jd.n_block = desc->d + 1 - (fam == "blocks" ? 3 : 0);
    return drj_model_create(&jd, device, out, err);
  }
  if (!desc->cuda_source) return set_err(err, DRJ_ERR_MODEL, "model has neither a built-in name nor CUDA source");
  if (desc->d < 1 || desc->n_block < 1) return set_err(err, DRJ_ERR_ARG, "source model needs d >= 1 and n_block >= 1");
  // the two names are spliced into the generated translation unit: identifiers only
  if ((desc->loglike_name && !valid_ident(desc->loglike_name)) || (desc->logprior_name && !valid_ident(desc->logprior_name)))
    return set_err(err, DRJ_ERR_ARG, "loglike / logprior must be plain C identifiers (the names of __device__ functions)");
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return set_err(err, DRJ_ERR_CUDA, "no CUDA device %d", device); }
  cudaFree(0);  // make sure the primary context exists before the driver API is used
  std::lock_guard<std::mutex> lock(g_mutex);
  drj_model* m = new drj_model();
  m->name = "jit"; m->D = desc->d; m->NBLOCK = desc->n_block; m->jit = true; m->device = device;
  int rc = jit_compile(desc, m, err);
  if (rc != DRJ_OK) { delete m; return rc; }
  *out = m;
  return DRJ_OK;
}

extern "C" void drj_model_destroy(drj_model* m) {
  if (!m) return;
  if (g_drv.ok) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : m->per_device)
      if (kv.second.module && cudaSetDevice(kv.first) == cudaSuccess) g_drv.ModuleUnload(kv.second.module);
    cudaSetDevice(cur);
  }
  delete m;
}

extern "C" double drj_model_compile_seconds(const drj_model* m) { return m ? m->compile_seconds : 0.0; }

// ------------------------------------------------------------------------------------------------
// session
// ------------------------------------------------------------------------------------------------
// ---- device-memory pool.  A session's buffers (tens of GB of stored draws for a long run) go back to a
// process-wide pool when the session is destroyed and are handed to the next session that asks for the same
// size, instead of cudaFree / cudaMalloc: a series of runs of one shape (the usual way run_mcmc is used, and
// what bench.py does) then costs no allocation time, and -- measured -- the memory-bound summary kernels of the
// next run no longer slow down several-fold now and then right after tens of GB were freed and re-allocated.
// At most DRJ_POOL_GB (default 48) GB per device, everything is released when an allocation fails or
// drj_release_cached_memory() is called.
struct PoolBlock { void* p; size_t bytes; int device; };
static std::vector<PoolBlock> g_pool;
static std::map<int, size_t> g_pool_bytes;  // cached bytes per device (the cap is per device: each GPU has its own 180 GB)
static std::mutex g_pool_mutex;

static size_t pool_cap() {
  static size_t cap = [] {
    const char* e = getenv("DRJ_POOL_GB");
    const double gb = e ? atof(e) : 48.0;
    return gb > 0 ? (size_t)(gb * 1073741824.0) : (size_t)0;
  }();
  return cap;
}
static size_t pool_release(int device) {  // device < 0: all devices
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  size_t freed = 0;
  int cur = 0;
  cudaGetDevice(&cur);
  for (size_t i = 0; i < g_pool.size();) {
    if (device < 0 || g_pool[i].device == device) {
      cudaSetDevice(g_pool[i].device);
      cudaFree(g_pool[i].p);
      freed += g_pool[i].bytes;
      g_pool_bytes[g_pool[i].device] -= g_pool[i].bytes;
      g_pool[i] = g_pool.back();
      g_pool.pop_back();
    } else ++i;
  }
  cudaSetDevice(cur);
  return freed;
}
static size_t pool_cached(int device) {
  std::lock_guard<std::mutex> lk(g_pool_mutex);
  size_t n = 0;
  for (const PoolBlock& b : g_pool) if (b.device == device) n += b.bytes;
  return n;
}
// size of the block that serves a request: small requests are rounded up to a power of two (>= 256 bytes) so that the
// few dozen small buffers of a session find a block of their class whatever the run's shape; large ones are kept exact
// (a series of runs of one shape asks for the same sizes).  EVERY device allocation of a session goes through the pool:
// cudaMalloc / cudaFree take a process-wide lock (and cudaFree synchronises the device), so with one host thread per GPU
// the ~50 small allocations and frees of each session serialised across the devices (8 GPUs: +0.18 s on a 0.22 s run).
static size_t pool_size(size_t bytes) {
  if (bytes >= (1u << 20)) return bytes;
  size_t c = 256;
  while (c < bytes) c <<= 1;
  return c;
}
static cudaError_t pool_alloc(int device, size_t bytes, void** out) {
  bytes = pool_size(bytes);
  {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (size_t i = 0; i < g_pool.size(); ++i)
      if (g_pool[i].device == device && g_pool[i].bytes == bytes) {
        *out = g_pool[i].p;
        g_pool_bytes[device] -= bytes;
        g_pool[i] = g_pool.back();
        g_pool.pop_back();
        return cudaSuccess;
      }
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) {  // give the cached blocks back to the driver and try again
    cudaGetLastError();
    if (pool_release(device) > 0) e = cudaMalloc(out, bytes);
  }
  return e;
}
static void pool_free(int device, void* p, size_t bytes) {
  if (!p) return;
  bytes = pool_size(bytes);
  {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    if (g_pool_bytes[device] + bytes <= pool_cap()) {
      g_pool.push_back({p, bytes, device});
      g_pool_bytes[device] += bytes;
      return;
    }
  }
  cudaFree(p);
}

extern "C" int64_t drj_release_cached_memory(void) { return (int64_t)pool_release(-1); }

struct drj_session {
  drj_config cfg;
  const drj_model* model = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  double sweep_ms_total = 0.0, stream_ms_total = 0.0;  // CUDA-event time: sweep kernels / sweep + transposes
  DrjKernelArgs a;
  std::vector<std::pair<void*, size_t>> allocs;  // device buffers of the session (returned to the pool on destroy)
  int D = 0, B = 0, R = 0, C = 0;
  long long P = 0, Ppt = 0, Pth = 0;
  int d_free = 0;
  int tb = 1;        // iterations per launch (ring depth)
  int cpc = 1, nthreads = 32, grid = 1;
  int t_done = 0;    // updates done (0 .. burnin-1+samples)
  int T_total = 0;
  // device outputs
  double *fin_ll_b = nullptr, *fin_lp_b = nullptr, *fin_th_b = nullptr;
  double *fin_ll_s = nullptr, *fin_lp_s = nullptr, *fin_th_s = nullptr;
  int *mc_b = nullptr, *mc_s = nullptr, *acc_b = nullptr;
  unsigned long long* err_key = nullptr;
  int64_t sweep_launches = 0, aux_launches = 0;
  bool stream_variant = false;  // data too large for shared memory: use the streaming sweep kernel
  double *draw_z = nullptr, *draw_lu = nullptr, *draw_clu = nullptr;  // draws computed ahead (DrjKernelArgs::draws)
  int* d_free_param = nullptr;  // free slot -> parameter index
  bool small = false;           // the sweep kernel of runs too small to fill the machine (draws ahead, parallel coupling decisions)
  bool team = false;            // few chains x long data: cluster-per-chain kernels (drj_sweep_item_team)
  int team_ctas = 1;            // CTAs per cluster
  bool team_ctas_forced = false;
  bool gridk = false;           // few particles x data >> L2: grid-per-run kernels (drj_sweep_item_grid), cooperative launch
  unsigned int* grid_bar = nullptr;
  bool lanes = false;           // a group of lanes per particle (drj_lanes.cuh); the init kernels stay one thread per particle
  int lanes_cpc = 0, lanes_grid = 0, lanes_threads = 0;
  unsigned lanes_smem = 0;
  bool spec = false;            // one CTA per chain, 5 or 7 updates in flight (drj_spec.cuh); same init kernels, same data buffer
  int spec_depth = 0;
  const JitFuncs* jf = nullptr; // source model: its kernels in this device's context
  unsigned dyn_smem = 0;        // dynamic shared memory of the model kernels: the data tile buffer
  double t_create = 0.0;
  bool failed = false;
  // summary scratch, kept for the life of the session (grow-only): cudaMalloc / cudaFree on every
  // drj_session_summary call cost up to ~0.5 s now and then next to tens of GB of stored draws
  void* scratch[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

static cudaError_t scratch_get(drj_session* s, int slot, size_t bytes, double** out) {
  bytes = std::max<size_t>(bytes, 8);
  if (s->scratch_bytes[slot] < bytes) {
    if (s->scratch[slot]) pool_free(s->cfg.device, s->scratch[slot], s->scratch_bytes[slot]);
    s->scratch[slot] = nullptr;
    s->scratch_bytes[slot] = 0;
    cudaError_t e = pool_alloc(s->cfg.device, bytes, &s->scratch[slot]);
    if (e != cudaSuccess) return e;
    s->scratch_bytes[slot] = bytes;
  }
  *out = (double*)s->scratch[slot];
  return cudaSuccess;
}

template <class T>
static cudaError_t dalloc(drj_session* s, T** p, size_t n) {
  void* q = nullptr;
  const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  cudaError_t e = pool_alloc(s->cfg.device, bytes, &q);
  if (e == cudaSuccess) { s->allocs.push_back({q, bytes}); *p = (T*)q; }
  return e;
}

template <class T>
static cudaError_t upload(drj_session* s, const T** dev, const T* host, size_t n) {
  T* q = nullptr;
  cudaError_t e = dalloc(s, &q, n);
  if (e != cudaSuccess) return e;
  *dev = q;
  if (n) e = cudaMemcpyAsync(q, host, n * sizeof(T), cudaMemcpyHostToDevice, s->stream);
  return e;
}

// data / misc list -> device, every item 16-byte aligned and the buffer padded by one tile so the
// TMA engine can always fetch whole tiles
static cudaError_t upload_list(drj_session* s, const drj_list* l, DrjList* out) {
  const int n = l ? l->n_items : 0;
  std::vector<long long> off(std::max(n, 1), 0), len(std::max(n, 1), 0);
  long long total = 0;
  for (int k = 0; k < n; ++k) {
    off[k] = total;
    len[k] = l->lengths[k];
    total += (len[k] + 1) & ~1LL;
  }
  double* vals = nullptr;
  cudaError_t e = dalloc(s, &vals, (size_t)total + DRJ_TILE + 2);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(vals, 0, ((size_t)total + DRJ_TILE + 2) * sizeof(double), s->stream);
  if (e != cudaSuccess) return e;
  for (int k = 0; k < n; ++k)
    if (len[k] > 0) {
      e = cudaMemcpyAsync(vals + off[k], l->values + l->offsets[k], (size_t)len[k] * sizeof(double),
                          cudaMemcpyHostToDevice, s->stream);
      if (e != cudaSuccess) return e;
    }
  const long long *doff = nullptr, *dlen = nullptr;
  e = upload(s, &doff, off.data(), off.size());
  if (e != cudaSuccess) return e;
  e = upload(s, &dlen, len.data(), len.size());
  if (e != cudaSuccess) return e;
  // the host vectors must outlive the async copies
  e = cudaStreamSynchronize(s->stream);
  out->values = vals; out->offsets = doff; out->lengths = dlen; out->n_items = n;
  return e;
}

static int cu_fail(drj_error* err, const char* what, CUresult cr) {
  const char* es = nullptr;
  g_drv.GetErrorString(cr, &es);
  return set_err(err, DRJ_ERR_CUDA, "%s: %s", what, es ? es : "?");
}

// initial likelihood and prior, one thread per chain (not for the cluster / grid kernels, whose init kernel does it itself)
static int launch_init_chain(drj_session* s, drj_error* err) {
  const drj_model* m = s->model;
  void* params[] = {(void*)&s->a};
  const unsigned threads = (unsigned)std::min(128, ((s->C + 31) / 32) * 32);
  const unsigned grid = (unsigned)((s->C + threads - 1) / threads);
  if (m->jit) {
    CUresult cr = g_drv.LaunchKernel(s->jf->init_chain_cu, grid, 1, 1, threads, 1, 1, s->dyn_smem, (CUstream)s->stream, params, nullptr);
    if (cr != CUDA_SUCCESS) return cu_fail(err, "cuLaunchKernel", cr);
  } else {
    CUDA_TRY(cudaLaunchKernel(m->init_chain_fn, dim3(grid), dim3(threads), params, s->dyn_smem, s->stream));
  }
  s->aux_launches++;
  return DRJ_OK;
}

static int launch_model_kernel(drj_session* s, bool sweep, drj_error* err) {
  const drj_model* m = s->model;
  void* params[] = {(void*)&s->a};
  if (sweep && s->a.draws) {  // the draws of this launch's iterations, ahead of the sweep on the same stream
    const long long work = s->P * s->d_free * (long long)s->a.n_iter;
    const int blocks = (int)std::min<long long>((work + 255) / 256, 8LL * 148);
    drj_draw_kernel<<<std::max(1, blocks), 256, 0, s->stream>>>(s->a, s->draw_z, s->draw_lu, s->draw_clu, s->d_free_param);
    CUDA_TRY(cudaGetLastError());
    s->aux_launches++;
  }
  const bool stream = sweep && s->stream_variant;
  if (s->gridk) {  // the whole grid is one team: cooperative launch (all CTAs co-resident), barrier counter at 0
    CUDA_TRY(cudaMemsetAsync(s->grid_bar, 0, sizeof(unsigned int), s->stream));
    if (m->jit) {
      CUresult cr = g_drv.LaunchCooperativeKernel(sweep ? s->jf->grid_sweep_cu : s->jf->grid_init_cu, (unsigned)s->grid, 1, 1,
                                                  (unsigned)s->nthreads, 1, 1, s->dyn_smem, (CUstream)s->stream, params);
      if (cr != CUDA_SUCCESS) return cu_fail(err, "cuLaunchCooperativeKernel", cr);
    } else {
      CUDA_TRY(cudaLaunchCooperativeKernel(sweep ? m->grid_sweep_fn : m->grid_init_fn, dim3(s->grid), dim3(s->nthreads), params,
                                           s->dyn_smem, s->stream));
    }
    return DRJ_OK;
  }
  if (s->team) {  // one cluster of team_ctas CTAs per chain
    if (m->jit) {
      CUlaunchAttribute at[1];
      at[0].id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
      at[0].value.clusterDim.x = (unsigned)s->team_ctas; at[0].value.clusterDim.y = 1; at[0].value.clusterDim.z = 1;
      CUlaunchConfig lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDimX = (unsigned)s->grid; lc.gridDimY = 1; lc.gridDimZ = 1;
      lc.blockDimX = (unsigned)s->nthreads; lc.blockDimY = 1; lc.blockDimZ = 1;
      lc.sharedMemBytes = s->dyn_smem; lc.hStream = (CUstream)s->stream; lc.attrs = at; lc.numAttrs = 1;
      CUresult cr = g_drv.LaunchKernelEx(&lc, sweep ? s->jf->team_sweep_cu : s->jf->team_init_cu, params, nullptr);
      if (cr != CUDA_SUCCESS) return cu_fail(err, "cuLaunchKernelEx", cr);
    } else {
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)s->team_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cudaLaunchConfig_t lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDim = dim3(s->grid); lc.blockDim = dim3(s->nthreads);
      lc.dynamicSmemBytes = s->dyn_smem; lc.stream = s->stream; lc.attrs = at; lc.numAttrs = 1;
      CUDA_TRY(cudaLaunchKernelExC(&lc, sweep ? m->team_sweep_fn : m->team_init_fn, params));
    }
    return DRJ_OK;
  }
  if (s->spec && sweep) {  // one CTA of 2^depth threads per chain; the data tile buffer is the one-thread-per-particle kernels'
    const unsigned nt = 1u << (s->spec_depth == 7 ? DRJ_SPEC_DEEP : 5);
    if (m->jit) {
      CUresult cr = g_drv.LaunchKernel(s->spec_depth == 7 ? s->jf->spec7_cu : s->jf->spec5_cu, (unsigned)s->C, 1, 1, nt, 1, 1, s->dyn_smem,
                                       (CUstream)s->stream, params, nullptr);
      if (cr != CUDA_SUCCESS) return cu_fail(err, "cuLaunchKernel", cr);
    } else {
      CUDA_TRY(cudaLaunchKernel(s->spec_depth == 7 ? m->spec7_fn : m->spec5_fn, dim3(s->C), dim3(nt), params, s->dyn_smem, s->stream));
    }
    return DRJ_OK;
  }
  if (s->lanes && sweep) {  // its own geometry; the init kernels of the session use the one-thread-per-particle one
    DrjKernelArgs a2 = s->a;
    a2.chains_per_cta = s->lanes_cpc;
    void* p2[] = {(void*)&a2};
    if (m->jit) {
      CUresult cr = g_drv.LaunchKernel(s->jf->lanes_sweep_cu, (unsigned)s->lanes_grid, 1, 1, (unsigned)s->lanes_threads, 1, 1, s->lanes_smem,
                                       (CUstream)s->stream, p2, nullptr);
      if (cr != CUDA_SUCCESS) return cu_fail(err, "cuLaunchKernel", cr);
    } else {
      CUDA_TRY(cudaLaunchKernel(m->lanes_sweep_fn, dim3(s->lanes_grid), dim3(s->lanes_threads), p2, s->lanes_smem, s->stream));
    }
    return DRJ_OK;
  }
  const bool small = sweep && !stream && s->small;
  if (m->jit) {
    CUresult cr = g_drv.LaunchKernel(!sweep ? s->jf->init_cu : (stream ? s->jf->stream_cu : (small ? s->jf->small_cu : s->jf->sweep_cu)), s->grid, 1, 1, s->nthreads, 1, 1,
                                     s->dyn_smem, (CUstream)s->stream, params, nullptr);
    if (cr != CUDA_SUCCESS) return cu_fail(err, "cuLaunchKernel", cr);
  } else {
    CUDA_TRY(cudaLaunchKernel(!sweep ? m->init_fn : (stream ? m->stream_fn : (small ? m->small_fn : m->sweep_fn)), dim3(s->grid), dim3(s->nthreads),
                              params, s->dyn_smem, s->stream));
  }
  return DRJ_OK;
}

// cudaGetDeviceProperties is slow (it queries the driver for ~100 attributes under a process-wide lock): once per device
static cudaError_t device_props(int device, cudaDeviceProp* out) {
  static std::mutex mu;
  static std::map<int, cudaDeviceProp> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(device);
  if (it == cache.end()) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return e;
    it = cache.emplace(device, p).first;
  }
  *out = it->second;
  return cudaSuccess;
}

static int prop_sms(const drj_session* s) {
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, s->cfg.device);
  return n;
}

// the one-thread-per-particle kernels: opt in to more than 48 KB of dynamic shared memory when the tile buffer
// plus the per-parameter tables need it (a per-datum model with many parameters)
static int solo_configure(drj_session* s, drj_error* err) {
  if ((size_t)s->dyn_smem + 24u * (size_t)s->D + 16384u <= 48u * 1024u) return DRJ_OK;  // static + dynamic within the default limit
  // (the attribute belongs to the function, not to the session: it only ever grows, so that a session created later
  // with a smaller buffer does not pull it below what an earlier, still open session launches with)
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, unsigned> granted;
  std::lock_guard<std::mutex> lock(mu);
  const drj_model* m = s->model;
  auto need = [&](const void* f) {
    unsigned& g = granted[std::make_pair(s->cfg.device, f)];
    if (g >= s->dyn_smem) return 0u;
    g = s->dyn_smem;
    return g;
  };
  if (m->jit) {
    for (CUfunction f : {s->jf->init_cu, s->jf->init_chain_cu, s->jf->sweep_cu, s->jf->small_cu, s->jf->stream_cu, s->jf->spec5_cu, s->jf->spec7_cu})
      // (a module's functions die with it and their addresses may be reused: no record kept, set every time)
      if (f && g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)s->dyn_smem) != CUDA_SUCCESS)
        return set_err(err, DRJ_ERR_CUDA, "cuFuncSetAttribute failed (dynamic shared memory %u bytes)", s->dyn_smem);
  } else {
    for (const void* f : {m->init_fn, m->init_chain_fn, m->sweep_fn, m->small_fn, m->stream_fn, m->spec5_fn, m->spec7_fn})
      if (f)
        if (const unsigned b = need(f)) CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b));
  }
  return DRJ_OK;
}

// GRID kernels: opt in to the dynamic shared memory of the chunk pipeline; grid = every CTA the device can hold at once
static int grid_configure(drj_session* s, int& grid_stages, drj_error* err) {
  const drj_model* m = s->model;
  int per_sm = 0;
  {  // as many 64 KB stages as fit next to the kernels' static shared memory (a user model's Consts table is part of it)
    int stat = 0, optin = 0;
    if (m->jit) {
      int a = 0, b = 0;
      CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuFuncGetAttribute", (void**)&FuncGetAttribute, cudaEnableDefault, &q) == cudaSuccess && FuncGetAttribute) {
        FuncGetAttribute(&a, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, s->jf->grid_sweep_cu);
        FuncGetAttribute(&b, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, s->jf->grid_init_cu);
      }
      stat = std::max(a, b);
    } else {
      cudaFuncAttributes fa;
      CUDA_TRY(cudaFuncGetAttributes(&fa, m->grid_sweep_fn));
      stat = (int)fa.sharedSizeBytes;
      CUDA_TRY(cudaFuncGetAttributes(&fa, m->grid_init_fn));
      stat = std::max(stat, (int)fa.sharedSizeBytes);
    }
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->cfg.device);
    const long long chunk_bytes = (long long)DRJ_GRID_CHUNK * sizeof(double);
    const int fit = (int)(((long long)optin - stat - 1024 - 64) / chunk_bytes);
    if (fit < 2) return set_err(err, DRJ_ERR_CUDA, "the grid kernels need two %lld-byte stages next to %d bytes of static shared memory", chunk_bytes, stat);
    grid_stages = std::min(grid_stages, fit);
    s->dyn_smem = (unsigned)(((long long)grid_stages * DRJ_GRID_CHUNK + 8) * sizeof(double));
  }
  if (m->jit) {
    for (CUfunction f : {s->jf->grid_init_cu, s->jf->grid_sweep_cu})
      if (g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)s->dyn_smem) != CUDA_SUCCESS)
        return set_err(err, DRJ_ERR_CUDA, "cuFuncSetAttribute failed for the grid kernels");
    int a = 0, b = 0;
    if (g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&a, s->jf->grid_sweep_cu, s->nthreads, s->dyn_smem) != CUDA_SUCCESS ||
        g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&b, s->jf->grid_init_cu, s->nthreads, s->dyn_smem) != CUDA_SUCCESS)
      return set_err(err, DRJ_ERR_CUDA, "cuOccupancyMaxActiveBlocksPerMultiprocessor failed for the grid kernels");
    per_sm = std::min(a, b);
  } else {
    for (const void* f : {m->grid_init_fn, m->grid_sweep_fn})
      CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->dyn_smem));
    int a = 0, b = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, m->grid_sweep_fn, s->nthreads, s->dyn_smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, m->grid_init_fn, s->nthreads, s->dyn_smem));
    per_sm = std::min(a, b);
  }
  if (per_sm < 1) return set_err(err, DRJ_ERR_CUDA, "the grid kernels do not fit on an SM (%u bytes of shared memory)", s->dyn_smem);
  int ctas = prop_sms(s);  // one CTA per SM streams 1/n_SM of the item
  if (const char* e = getenv("DRJ_GRID_CTAS")) {
    const int v = atoi(e);
    if (v >= 1) ctas = std::min(v, prop_sms(s) * per_sm);
  }
  s->grid = ctas;
  return DRJ_OK;
}

// TEAM kernels: opt in to the dynamic shared memory of the tile pipeline and to clusters of 16 CTAs, then
// shrink the cluster until the device can co-schedule at least one
static int team_configure(drj_session* s, drj_error* err) {
  const drj_model* m = s->model;
  for (int which = 0; which < 2; ++which) {
    if (m->jit) {
      CUfunction f = which ? s->jf->team_sweep_cu : s->jf->team_init_cu;
      CUresult cr = g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)s->dyn_smem);
      if (cr == CUDA_SUCCESS) cr = g_drv.FuncSetAttribute(f, CU_FUNC_ATTRIBUTE_NON_PORTABLE_CLUSTER_SIZE_ALLOWED, 1);
      if (cr != CUDA_SUCCESS) return set_err(err, DRJ_ERR_CUDA, "cuFuncSetAttribute failed for the cluster kernels");
    } else {
      const void* f = which ? m->team_sweep_fn : m->team_init_fn;
      CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->dyn_smem));
      CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
  }
  while (s->team_ctas > 1) {
    int nclusters = 0;
    if (m->jit) {
      CUlaunchAttribute at[1];
      at[0].id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
      at[0].value.clusterDim.x = (unsigned)s->team_ctas; at[0].value.clusterDim.y = 1; at[0].value.clusterDim.z = 1;
      CUlaunchConfig lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDimX = (unsigned)s->team_ctas; lc.gridDimY = 1; lc.gridDimZ = 1;
      lc.blockDimX = (unsigned)s->nthreads; lc.blockDimY = 1; lc.blockDimZ = 1;
      lc.sharedMemBytes = s->dyn_smem; lc.attrs = at; lc.numAttrs = 1;
      if (g_drv.OccupancyMaxActiveClusters(&nclusters, s->jf->team_sweep_cu, &lc) != CUDA_SUCCESS) nclusters = 0;
    } else {
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)s->team_ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cudaLaunchConfig_t lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDim = dim3(s->team_ctas); lc.blockDim = dim3(s->nthreads);
      lc.dynamicSmemBytes = s->dyn_smem; lc.attrs = at; lc.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&nclusters, m->team_sweep_fn, &lc) != cudaSuccess) { cudaGetLastError(); nclusters = 0; }
    }
    // all chains' clusters resident at once (a 16-CTA cluster needs a whole GPC: at most 7-8 fit), or at
    // least one when the size was forced
    if (nclusters >= (s->team_ctas_forced ? 1 : std::min(s->C, prop_sms(s)))) break;
    s->team_ctas /= 2;
  }
  return DRJ_OK;
}

static int transpose_to(drj_session* s, const double* ring, long long N, int ncols, double* dst, long long dst_stride,
                        long long dst_off, drj_error* err) {
  if (ncols <= 0 || N <= 0) return DRJ_OK;
  dim3 block(32, 8);
  for (int c0 = 0; c0 < ncols; c0 += 32 * 65535) {
    const int nc = std::min(ncols - c0, 32 * 65535);
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((nc + 31) / 32));
    drj_transpose_rows<<<grid, block, 0, s->stream>>>(ring + (long long)c0 * N, N, nc, dst, dst_stride, dst_off + c0);
    s->aux_launches++;
  }
  CUDA_TRY(cudaGetLastError());
  return DRJ_OK;
}

// move `nt` ring rows into the final per-series layout at column `col0` of phase `sampling`
static int flush_ring(drj_session* s, int nt, bool sampling, int col0, drj_error* err) {
  const int L = sampling ? s->cfg.samples : s->cfg.burnin;
  int rc;
  if ((rc = transpose_to(s, s->a.ring_ll, s->Ppt, nt, sampling ? s->fin_ll_s : s->fin_ll_b, L, col0, err))) return rc;
  if ((rc = transpose_to(s, s->a.ring_lp, s->Ppt, nt, sampling ? s->fin_lp_s : s->fin_lp_b, L, col0, err))) return rc;
  // theta ring [nt][D][Pth] is [nt*D][Pth]: column (t*D + i) -> dst[(n*L + col0 + t)*D + i]
  return transpose_to(s, s->a.ring_theta, s->Pth, nt * s->D, sampling ? s->fin_th_s : s->fin_th_b,
                      (long long)L * s->D, (long long)col0 * s->D, err);
}

static const char* code_message(int code) {
  switch (code) {
    case DRJ_ERR_INIT_NEGINF:
      return "Starting values result in -Inf in likelihood or prior. Consider setting inital values in the parameters data.frame.";
    case DRJ_ERR_LOGLIKE_BAD: return "NA, NaN or Inf in likelihood";
    case DRJ_ERR_LOGPRIOR_BAD: return "NA, NaN or Inf in prior";
    case DRJ_ERR_TRANS_TYPE: return "trans_type invalid";
    default: return "error";
  }
}

// after a synchronised launch: did any particle report an error?  (launch_iter0 = global update index of
// the launch's first iteration, 0 for init)
static int check_device_error(drj_session* s, int launch_iter0, drj_error* err) {
  unsigned long long key = 0;
  CUDA_TRY(cudaMemcpy(&key, s->err_key, sizeof(key), cudaMemcpyDeviceToHost));
  if (key == ~0ULL) return DRJ_OK;
  const int it = (int)(key >> 44);
  const long long chain = (long long)((key >> 20) & 0xFFFFFF);
  const int rung = (int)((key >> 10) & 0x3FF);
  const int param = (int)(key & 0x3FF) - 1;
  const long long p = chain * s->R + rung;
  int code = 0;
  CUDA_TRY(cudaMemcpy(&code, s->a.err_code + p, sizeof(int), cudaMemcpyDeviceToHost));
  s->failed = true;
  if (err) {
    memset(err, 0, sizeof(*err));
    err->code = code; err->rung = rung; err->param = param;
    err->iteration = (param < 0) ? 0 : launch_iter0 + it;
    err->chain = s->cfg.chain_offset + chain;
    for (int i = 0; i < s->D && i < 64; ++i)
      cudaMemcpy(&err->theta[i], s->a.err_theta + (long long)i * s->P + p, sizeof(double), cudaMemcpyDeviceToHost);
    std::string th;
    for (int i = 0; i < s->D && i < 8; ++i) { char b[40]; snprintf(b, sizeof(b), "%s%.6g", i ? " " : "", err->theta[i]); th += b; }
    snprintf(err->message, sizeof(err->message), "%s (chain %lld, rung %d, parameter %d, iteration %d; theta = %s%s)",
             code_message(code), (long long)err->chain + 1, rung + 1, param + 1, err->iteration, th.c_str(),
             s->D > 8 ? " ..." : "");
  }
  return code ? code : DRJ_ERR_CUDA;
}

static int validate(const drj_config* c, const drj_model* m, const drj_list* data, drj_error* err) {
  if (!c || !m) return set_err(err, DRJ_ERR_ARG, "NULL config or model");
  if (c->abi_version != DRJ_ABI_VERSION) return set_err(err, DRJ_ERR_ARG, "ABI version mismatch (%d != %d)", c->abi_version, DRJ_ABI_VERSION);
  if (c->d < 1 || c->chains < 1 || c->rungs < 1 || c->burnin < 1 || c->samples < 1 || c->n_block < 1)
    return set_err(err, DRJ_ERR_ARG, "d, chains, rungs, burnin, samples and n_block must be positive");
  if (c->d != m->D || c->n_block != m->NBLOCK)
    return set_err(err, DRJ_ERR_MODEL, "model '%s' has d=%d, n_block=%d but the config has d=%d, n_block=%d", m->name.c_str(),
                   m->D, m->NBLOCK, c->d, c->n_block);
  if (c->rungs > DRJ_NT) return set_err(err, DRJ_ERR_ARG, "rungs > %d not supported", DRJ_NT);
  if (m->datum_form && c->d > 256) return set_err(err, DRJ_ERR_ARG, "per-datum (DATUM_FORM) models support at most 256 parameters; use the generic form");
  if (c->layout_chains < 0) return set_err(err, DRJ_ERR_ARG, "layout_chains must be >= 0");
  if (c->d > 1022) return set_err(err, DRJ_ERR_ARG, "d > 1022 not supported");
  if (c->chains >= (1 << 24)) return set_err(err, DRJ_ERR_ARG, "more than 2^24-1 chains per call: shard the run");
  if (!(c->target_acceptance >= 0.0 && c->target_acceptance <= 1.0)) return set_err(err, DRJ_ERR_ARG, "target_acceptance must be in [0,1]");
  if (!c->theta_min || !c->theta_max || !c->trans_type || !c->skip_param || !c->block_ptr || !c->block_idx ||
      !c->beta_raised || !c->theta_init)
    return set_err(err, DRJ_ERR_ARG, "NULL array in config");
  for (int i = 0; i < c->d; ++i) {
    if (c->trans_type[i] < 0 || c->trans_type[i] > 3) return set_err(err, DRJ_ERR_TRANS_TYPE, "trans_type invalid");
    for (int j = c->block_ptr[i]; j < c->block_ptr[i + 1]; ++j)
      if (c->block_idx[j] < 1 || c->block_idx[j] > c->n_block) return set_err(err, DRJ_ERR_ARG, "block id out of range");
  }
  for (int r = 0; r < c->rungs; ++r) {
    if (!(c->beta_raised[r] >= 0.0 && c->beta_raised[r] <= 1.0)) return set_err(err, DRJ_ERR_ARG, "beta_raised must be in [0,1]");
    if (r > 0 && !(c->beta_raised[r] > c->beta_raised[r - 1])) return set_err(err, DRJ_ERR_ARG, "beta_raised must be increasing");
  }
  if (c->beta_raised[c->rungs - 1] != 1.0) return set_err(err, DRJ_ERR_ARG, "last beta_raised must be 1");
  if (c->rng_mode == DRJ_RNG_REPLAY && !c->replay_uniforms) return set_err(err, DRJ_ERR_ARG, "replay mode needs replay_uniforms");
  if (c->rng_mode != DRJ_RNG_REPLAY && c->rng_mode != DRJ_RNG_PHILOX) return set_err(err, DRJ_ERR_ARG, "unknown rng_mode");
  if (m->datum_form && (!data || data->n_items < 1)) return set_err(err, DRJ_ERR_ARG, "model '%s' needs a data item", m->name.c_str());
  return DRJ_OK;
}

extern "C" void drj_session_destroy(drj_session* s) {
  if (!s) return;
  cudaSetDevice(s->cfg.device);
  cudaStreamSynchronize(s->stream);  // nothing of this session may still be running when its buffers are reused
  for (auto& pb : s->allocs) pool_free(s->cfg.device, pb.first, pb.second);
  for (int k = 0; k < 12; ++k) if (s->scratch[k]) pool_free(s->cfg.device, s->scratch[k], s->scratch_bytes[k]);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->ev2) cudaEventDestroy(s->ev2);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

static int session_create_impl(drj_session* s, const drj_config* cfg, const drj_model* model, const drj_list* data,
                               const drj_list* misc, drj_error* err) {
  s->cfg = *cfg;
  s->model = model;
  s->t_create = wall_now();
  const int D = s->D = cfg->d, B = s->B = cfg->n_block, R = s->R = cfg->rungs, C = s->C = cfg->chains;
  const long long P = s->P = (long long)C * R;
  s->Ppt = cfg->store_pt ? P : C;
  s->Pth = cfg->save_hot_draws ? P : C;
  s->T_total = cfg->burnin - 1 + cfg->samples;
  if (cudaSetDevice(cfg->device) != cudaSuccess) { cudaGetLastError(); return set_err(err, DRJ_ERR_CUDA, "no CUDA device %d", cfg->device); }
  CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&s->ev0));
  CUDA_TRY(cudaEventCreate(&s->ev1));
  CUDA_TRY(cudaEventCreate(&s->ev2));
  if (model->jit) {
    cudaFree(0);  // the primary context of this device must exist before the driver API is used
    int rc = jit_functions(model, cfg->device, &s->jf, err);
    if (rc) return rc;
  }

  // ---- launch geometry: all rungs of a chain in one CTA, as many chains per CTA as fit in DRJ_NT
  // threads, but no more than needed to give every SM a few CTAs
  cudaDeviceProp prop;
  CUDA_TRY(device_props(cfg->device, &prop));
  const int max_cpc = std::max(1, DRJ_NT / R);
  int cpc = cfg->chains_per_cta > 0 ? std::min(cfg->chains_per_cta, max_cpc) : max_cpc;
  if (cfg->chains_per_cta <= 0) {
    // ~128-thread CTAs: with ~128 registers per thread four of them fill an SM's register file, and
    // a finer CTA grain shortens the last, partially filled wave (all CTAs do equal work, so the
    // tail is a whole-CTA quantum).  Fewer chains per CTA still when there are few chains, so that
    // every SM gets work.
    cpc = std::min(cpc, std::max(1, 128 / R));
    const long long want_ctas = 4LL * prop.multiProcessorCount;
    const int fit = (int)std::max<long long>(1, (C + want_ctas - 1) / want_ctas);
    cpc = std::min(cpc, fit);
    cpc = std::max(cpc, std::min(max_cpc, (32 + R - 1) / R));  // but never less than one full warp
    if (const char* e = getenv("DRJ_CHAINS_PER_CTA")) {            // tuning override
      const int v = atoi(e);
      if (v > 0) cpc = std::min(v, max_cpc);
    }
  }
  s->cpc = cpc;
  s->nthreads = std::min(DRJ_NT, ((cpc * R + 31) / 32) * 32);
  s->grid = (C + cpc - 1) / cpc;

  // ---- sizes and the memory check
  s->d_free = 0;
  std::vector<int> free_slot(D);
  for (int i = 0; i < D; ++i) free_slot[i] = cfg->skip_param[i] ? -1 : s->d_free++;
  const long long U = 3LL * s->d_free * R + ((cfg->coupling_on && R > 1) ? (R - 1) : 0);
  const double bytes_final = 8.0 * ((double)s->Ppt * 2 + (double)s->Pth * D) * ((double)cfg->burnin + cfg->samples);
  const double bytes_state = 8.0 * (double)P * (6.0 * D + B + 4);
  const double bytes_replay = cfg->rng_mode == DRJ_RNG_REPLAY ? 8.0 * (double)C * s->T_total * U : 0.0;
  const double ring_row = 8.0 * ((double)s->Ppt * 2 + (double)s->Pth * D);
  int tb = cfg->iters_per_launch > 0 ? cfg->iters_per_launch : (int)std::min(1024.0, std::max(1.0, 268435456.0 / ring_row));
  tb = std::max(1, std::min(tb, std::max(1, s->T_total)));
  tb = std::min(tb, (1 << 20) - 1);
  tb = std::min(tb, std::max(1, 4000000 / (D * B + 1)));  // (grid kernels: the barrier counter of a launch stays below 2^32)
  s->tb = tb;
  size_t free_b = 0, total_b = 0;
  CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  const double need = bytes_final + bytes_state + bytes_replay + ring_row * tb + 64e6;
  free_b += pool_cached(cfg->device);  // blocks in the library's own pool are reused or released on demand
  if (need > 0.92 * (double)free_b)
    return set_err(err, DRJ_ERR_MEMORY,
                   "run needs %.1f GB of device memory (outputs %.1f GB) but %.1f GB are free: reduce iterations, "
                   "chains per call, or turn off store_pt / save_hot_draws",
                   need / 1e9, bytes_final / 1e9, (double)free_b / 1e9);

  // datum-form model with a long data item -> the sweep-kernel variant that parks the particle state in
  // HBM across data sweeps (mandatory when the item is streamed, n > 2 tiles; measured faster from
  // n ~ 2048 on even when the data is resident in shared memory: tools/n_sweep.py)
  if (model->datum_form && data && (model->jit ? s->jf->stream_cu != nullptr : model->stream_fn != nullptr))
    for (int k = 0; k < data->n_items; ++k)
      if (data->lengths[k] >= DRJ_TILE) s->stream_variant = true;
  if (model->datum_form && (model->jit ? s->jf->stream_cu != nullptr : model->stream_fn != nullptr)) {
    if (cfg->flags & DRJ_FLAG_STREAM_KERNEL) s->stream_variant = true;    // tuning / testing overrides
    if (cfg->flags & DRJ_FLAG_RESIDENT_KERNEL) s->stream_variant = false;
  }

  // few chains over a long data item -> one thread-block cluster per chain, the sweep divided over the
  // cluster's threads (drj_sweep_item_team).  With P = chains x rungs threads the one-thread-per-particle
  // kernels need ~5 cycles per datum per sweep whatever P is, until P fills the machine (~75k threads).
  // The family is chosen from the WHOLE run's particle count when this call is a shard of a larger run
  // (cfg->layout_chains): the families associate the likelihood sum differently, so every shard must use the one
  // the unsharded run would use to stay bit-identical with it.
  const long long layout_C = cfg->layout_chains > 0 ? std::max(cfg->layout_chains, C) : C;
  const long long layout_P = layout_C * R;
  int team_stages = 3;  // x 64 KB tiles
  int grid_stages = 3;  // x 64 KB chunks (fewer if the model's static shared memory leaves less room: grid_configure)
  {
    long long longest = 0;
    if (data) for (int k = 0; k < data->n_items; ++k) longest = std::max<long long>(longest, data->lengths[k]);
    // few particles over data far larger than L2 -> the grid-per-run kernels: every SM streams 1/n_SM of the item
    // ONCE per sweep and accumulates for all particles (the HBM-bound regime; with more particles than ~16 the sweep
    // is FP64-bound and a cluster per chain does as well)
    const bool can_grid = model->datum_form && data && layout_P <= DRJ_GRID_MAXP &&
                          (model->jit ? (s->jf->grid_sweep_cu != nullptr && g_drv.LaunchCooperativeKernel != nullptr)
                                      : model->grid_sweep_fn != nullptr);
    bool want_grid = longest >= (1LL << 20) && layout_P <= 16;
    if (cfg->flags & DRJ_FLAG_GRID_KERNEL) want_grid = true;
    if (cfg->flags & (DRJ_FLAG_NO_GRID_KERNEL | DRJ_FLAG_TEAM_KERNEL)) want_grid = false;
    if (const char* e = getenv("DRJ_GRID")) want_grid = (e[0] == '1') ? true : (e[0] == '0' ? false : want_grid);
    if (can_grid && want_grid) {
      int coop = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg->device);
      if (coop) {
        s->gridk = true;
        s->stream_variant = false;
        if (const char* e = getenv("DRJ_GRID_STAGES")) grid_stages = std::max(2, std::min(DRJ_GRID_MAXSTAGES, atoi(e)));
        s->cpc = C;
        s->nthreads = DRJ_NT;
      }
    }
    const bool can = !s->gridk && model->datum_form && data && R <= DRJ_TEAM_MAXR &&
                     (model->jit ? (s->jf->team_sweep_cu != nullptr && g_drv.LaunchKernelEx != nullptr) : model->team_sweep_fn != nullptr);
    bool want = longest > 2LL * DRJ_TILE && layout_P <= 4096;  // measured cross-over, tools/team_bench.py
    if (cfg->flags & DRJ_FLAG_TEAM_KERNEL) want = true;
    if (cfg->flags & DRJ_FLAG_NO_TEAM_KERNEL) want = false;
    if (const char* e = getenv("DRJ_TEAM")) want = (e[0] == '1') ? true : (e[0] == '0' ? false : want);
    if (can && want) {
      s->team = true;
      s->stream_variant = false;
      const long long ntiles = std::max<long long>(1, (longest + DRJ_TEAM_TILE - 1) / DRJ_TEAM_TILE);
      int K = 1;  // upper bound; team_configure() shrinks it until all chains' clusters are co-resident
      while (K * 2 <= DRJ_TEAM_VRANKS && (long long)K * 2 * C <= prop.multiProcessorCount && K * 2 <= ntiles) K *= 2;
      if (const char* e = getenv("DRJ_TEAM_CTAS")) {
        const int v = atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) { K = v; s->team_ctas_forced = true; }
      }
      if (const char* e = getenv("DRJ_TEAM_STAGES")) team_stages = std::max(2, std::min(3, atoi(e)));
      s->team_ctas = K;
      s->cpc = DRJ_TEAM_NT / R;  // slots per rung in a CTA
      s->nthreads = DRJ_TEAM_NT;
    }
  }

  // one CTA per chain, five or seven updates in flight (drj_spec.cuh): a run of few single-rung chains of a small model -- the
  // reference's everyday shape -- is a serial recurrence that one thread per chain runs at ~3 us per update.  The
  // results are bitwise those of the one-thread-per-particle kernels, so the choice may depend on this shard's size.
  {
    long long longest = 0;
    if (data) for (int k = 0; k < data->n_items; ++k) longest = std::max<long long>(longest, data->lengths[k]);
    bool can = model->spec && !s->team && !s->gridk && R == 1 && B == 1 && D <= 4 &&
               (model->jit ? s->jf->spec5_cu != nullptr : model->spec5_fn != nullptr) &&
               (!model->datum_form || longest <= 2LL * DRJ_TILE);
    int n_free = 0;
    for (int i = 0; i < D; ++i) {
      if (cfg->skip_param[i]) continue;
      ++n_free;
      can = can && (cfg->block_ptr[i + 1] - cfg->block_ptr[i] == 1);  // every free parameter lists the one block
    }
    can = can && n_free >= 1;
    bool want = C <= 1024;
    if (cfg->flags & DRJ_FLAG_SPEC_KERNEL) want = true;
    if (cfg->flags & (DRJ_FLAG_NO_SPEC_KERNEL | DRJ_FLAG_LANES_KERNEL)) want = false;
    if (const char* e = getenv("DRJ_SPEC")) want = (e[0] == '1') ? true : (e[0] == '0' ? false : want);
    if (can && want) {
      s->spec = true;
      s->stream_variant = false;
      // seven updates in flight (four warps per chain, one per scheduler) while the chains leave the SMs room for it,
      // five (one warp per chain) for more chains
      s->spec_depth = (C <= 2LL * prop.multiProcessorCount) ? 7 : 5;
      if (const char* e = getenv("DRJ_SPEC_DEPTH")) { const int v = atoi(e); if (v == 5 || v == 7) s->spec_depth = v; }
    }
  }

  // a group of lanes per particle (drj_lanes.cuh): generic models that declare a lane form (many parameters: the state moves
  // from local to shared memory and the blocks are summed lane-parallel), and per-datum models when the run has too few
  // particles to fill the machine with one thread each and the data is short (longer data: the cluster kernels above).
  // Chosen from the WHOLE run's particle count, like the other families (the likelihood sums associate differently).
  {
    long long longest = 0;
    if (data) for (int k = 0; k < data->n_items; ++k) longest = std::max<long long>(longest, data->lengths[k]);
    const int Lm = model->lanes;
    const bool can = Lm > 0 && !s->team && !s->gridk && !s->spec && R <= std::min(32, DRJ_NT / std::max(1, Lm)) &&
                     (model->jit ? s->jf->lanes_sweep_cu != nullptr : model->lanes_sweep_fn != nullptr);
    bool want = model->datum_form ? (layout_P <= 2048 && longest <= 2LL * DRJ_TILE) : true;
    if (cfg->flags & DRJ_FLAG_LANES_KERNEL) want = true;
    if (cfg->flags & DRJ_FLAG_NO_LANES_KERNEL) want = false;
    if (const char* e = getenv("DRJ_LANES")) want = (e[0] == '1') ? true : (e[0] == '0' ? false : want);
    if (can && want) {
      int stat = 0, optin = 0, per_sm = 0;
      if (model->jit) {
        CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuFuncGetAttribute", (void**)&FuncGetAttribute, cudaEnableDefault, &qr) == cudaSuccess && FuncGetAttribute)
          FuncGetAttribute(&stat, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, s->jf->lanes_sweep_cu);
        else stat = 16384;
      } else {
        cudaFuncAttributes fa;
        CUDA_TRY(cudaFuncGetAttributes(&fa, model->lanes_sweep_fn));
        stat = (int)fa.sharedSizeBytes;
      }
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
      cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, cfg->device);
      const long long idx_bytes = 4LL * cfg->block_ptr[D] + 16;
      const int np_max = DRJ_NT / Lm;
      // two CTAs per SM when a whole chain fits in half of an SM's shared memory, else one
      int cpc_l = 0;
      for (long long budget : {(long long)per_sm / 2 - stat - 1024 - idx_bytes, (long long)optin - stat - 1024 - idx_bytes}) {
        const int np_fit = (int)std::min<long long>(np_max, budget / model->lanes_bpp);
        cpc_l = np_fit / R;
        if (cpc_l >= 1) break;
      }
      if (cpc_l >= 1) {
        if (cfg->chains_per_cta > 0) cpc_l = std::min(cpc_l, cfg->chains_per_cta);
        else cpc_l = std::min(cpc_l, std::max(1, (int)((C + 2LL * prop.multiProcessorCount - 1) / (2LL * prop.multiProcessorCount))));  // few chains: spread them over the SMs
        s->lanes = true;
        s->stream_variant = false;
        s->lanes_cpc = cpc_l;
        s->lanes_threads = ((cpc_l * R * Lm + 31) / 32) * 32;
        s->lanes_grid = (C + cpc_l - 1) / cpc_l;
        s->lanes_smem = (unsigned)((long long)(s->lanes_threads / Lm) * model->lanes_bpp + idx_bytes);
        if (model->jit) {
          if (g_drv.FuncSetAttribute(s->jf->lanes_sweep_cu, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)s->lanes_smem) != CUDA_SUCCESS)
            return set_err(err, DRJ_ERR_CUDA, "cuFuncSetAttribute failed for the lanes kernel (%u bytes of shared memory)", s->lanes_smem);
        } else {
          CUDA_TRY(cudaFuncSetAttribute(model->lanes_sweep_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->lanes_smem));
        }
      }
    }
  }

  // dynamic shared memory = the data tile buffer of datum-form models: the longest item if every item
  // stays resident (<= 2 tiles), else two tiles; + 8 doubles of slack for the pipelined prefetch
  if (model->datum_form && data) {
    long long longest = 0;
    for (int k = 0; k < data->n_items; ++k) longest = std::max<long long>(longest, data->lengths[k]);
    // (the kernels keep the data resident only for single-block models: DrjTiles::init)
    const bool resident = model->NBLOCK == 1 && longest <= 2 * DRJ_TILE;
    const long long doubles = (resident ? ((longest + 1) & ~1LL) : 2LL * DRJ_TILE) + 8;
    s->dyn_smem = (unsigned)(doubles * sizeof(double));
    if (s->team) {
      s->dyn_smem = (unsigned)(((long long)team_stages * DRJ_TEAM_TILE + 8) * sizeof(double));
      int rc = team_configure(s, err);
      if (rc) return rc;
      s->grid = C * s->team_ctas;
    } else if (s->gridk) {
      s->dyn_smem = (unsigned)(((long long)grid_stages * DRJ_GRID_CHUNK + 8) * sizeof(double));
      int rc = grid_configure(s, grid_stages, err);
      if (rc) return rc;
    } else {
      int rc = solo_configure(s, err);
      if (rc) return rc;
    }
  }

  if (const char* e = getenv("DRJ_CARVEOUT")) {  // tuning: shared-memory carve-out (percent) of the one-thread-per-particle sweep kernel
    const int pct = atoi(e);
    if (!s->team && !s->gridk && !model->jit && pct >= 0 && pct <= 100)
      cudaFuncSetAttribute(s->stream_variant ? model->stream_fn : model->sweep_fn, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  }

  // ---- uploads
  DrjKernelArgs& a = s->a;
  memset(&a, 0, sizeof(a));
  {  // chains inside one warp (rungs | 32) and no CTA-wide streamed sweep: per-warp synchronisation
    bool streamed = false;
    if (model->datum_form && data)
      for (int k = 0; k < data->n_items; ++k) streamed = streamed || data->lengths[k] > 2 * DRJ_TILE;
    // Per-warp mode puts one coupling walker in EVERY warp (about 25 x (R-1) issue slots per iteration at
    // 1/32 lane efficiency) instead of packing the CTA's walkers into one warp, so it only pays when the
    // update phase is long compared with the walk (measured: K4 +6 %, the 2-parameter n = 100 sweep -6 %).
    int d_free_n = 0;
    for (int i = 0; i < D; ++i) d_free_n += cfg->skip_param[i] ? 0 : 1;
    // (a per-datum model with several blocks never keeps its data resident: its sweeps take the tile path with
    // CTA-wide barriers and a single producer thread, so its warps must not run ahead of each other)
    const bool cta_wide = streamed || s->stream_variant || s->team || s->gridk || (model->datum_form && model->NBLOCK > 1);
    a.warp_local = (32 % R == 0 && !cta_wide && 6 * d_free_n > R - 1) ? 1 : 0;
    if (const char* e = getenv("DRJ_WARP_LOCAL")) if (e[0] == '1' && 32 % R == 0 && !cta_wide) a.warp_local = 1;
    if (const char* e = getenv("DRJ_NO_WARP_LOCAL")) if (e[0] == '1') a.warp_local = 0;
  }
  a.chains = C; a.rungs = R; a.chains_per_cta = s->cpc; a.team_ctas = s->team_ctas; a.team_stages = team_stages; a.coupling_on = cfg->coupling_on ? 1 : 0;
  a.grid_stages = grid_stages;
  a.grid_slots = (int)std::max<long long>(1, DRJ_NT / std::max<long long>(1, layout_P));  // from the whole run's size: shards stay bit-identical
  a.grid_tile = (layout_P >= 2 && layout_P <= 8) ? 1 : 0;  // (likewise from the whole run's size)
  {
    bool trivial = model->NBLOCK == 1;
    for (int i = 0; i < D && trivial; ++i)
      if (!cfg->skip_param[i]) trivial = cfg->block_ptr[i + 1] - cfg->block_ptr[i] == 1 && cfg->block_idx[cfg->block_ptr[i]] == 1;
    a.blocks_trivial = trivial ? 1 : 0;
    unsigned m = 0u;
    for (int i = 0; i < D && i < 32; ++i)
      if ((cfg->trans_type[i] == 2 || cfg->trans_type[i] == 3) && cfg->theta_min[i] == 0.0) m |= 1u << i;
    if (const char* e = getenv("DRJ_NO_HINTS")) if (e[0] == '1') m = 0u;
    a.lo_is_log = m;
  }
  if (const char* e = getenv("DRJ_GRID_TILE")) { if (e[0] == '0') a.grid_tile = 0; else if (e[0] == '1' && layout_P <= 8) a.grid_tile = 1; }
  a.save_hot_draws = cfg->save_hot_draws ? 1 : 0; a.store_pt = cfg->store_pt ? 1 : 0;
  a.rng_mode = cfg->rng_mode; a.d_free = s->d_free;
  a.use_tma = (cfg->flags & DRJ_FLAG_NO_TMA) ? 0 : 1;
  if (const char* e = getenv("DRJ_NO_TMA")) if (e[0] == '1') a.use_tma = 0;
  a.chain_offset = cfg->chain_offset; a.seed = cfg->seed; a.target_acceptance = cfg->target_acceptance; a.P = P;
  CUDA_TRY(upload(s, &a.theta_min, cfg->theta_min, D));
  CUDA_TRY(upload(s, &a.theta_max, cfg->theta_max, D));
  CUDA_TRY(upload(s, &a.trans_type, cfg->trans_type, D));
  CUDA_TRY(upload(s, &a.free_slot, free_slot.data(), D));
  std::vector<int> free_param;
  for (int i = 0; i < D; ++i) if (free_slot[i] >= 0) free_param.push_back(i);
  {
    // draws computed ahead for the one-thread-per-particle kernels when the run is too small to fill the machine (at
    // most two warps per scheduler: every update is then pure latency, and the draw is a third of it); from ~2 warps per
    // scheduler on the sweep hides that latency itself and the extra pass through HBM would only cost
    cudaDeviceProp prop3;
    const bool solo = !s->team && !s->gridk && !s->lanes && !s->spec && !s->stream_variant;
    const double draw_bytes = 16.0 * ((double)s->tb + 1) * s->d_free * (double)P + 8.0 * (double)s->tb * (double)C * std::max(R - 1, 1);
    bool want = solo && s->d_free > 0 && device_props(cfg->device, &prop3) == cudaSuccess &&
                P <= 32LL * 8 * prop3.multiProcessorCount && draw_bytes <= 1.0e9;
    if (const char* e = getenv("DRJ_DRAWS")) want = (e[0] == '1') ? (solo && s->d_free > 0 && draw_bytes <= 8.0e9) : (e[0] == '0' ? false : want);
    if (want) {
      const size_t rows = ((size_t)s->tb + 1) * (size_t)s->d_free * (size_t)P;
      CUDA_TRY(dalloc(s, &s->draw_z, rows));
      CUDA_TRY(dalloc(s, &s->draw_lu, rows));
      CUDA_TRY(cudaMemsetAsync(s->draw_z, 0, rows * sizeof(double), s->stream));   // (the padded row is read, never used)
      CUDA_TRY(cudaMemsetAsync(s->draw_lu, 0, rows * sizeof(double), s->stream));
      if (cfg->coupling_on && R > 1) CUDA_TRY(dalloc(s, &s->draw_clu, (size_t)s->tb * (size_t)C * (size_t)(R - 1)));
      const int* fp = nullptr;
      CUDA_TRY(upload(s, &fp, free_param.data(), free_param.size()));
      s->d_free_param = const_cast<int*>(fp);
      a.draws = 1; a.draw_z = s->draw_z; a.draw_lu = s->draw_lu; a.draw_clu = s->draw_clu;
    }
    // (the same small runs: coupling decisions for every possible carried replica at once, drj_sampler.cuh)
    bool par = solo && R > 1 && R <= 32 && cfg->coupling_on && device_props(cfg->device, &prop3) == cudaSuccess &&
               P <= 32LL * 8 * prop3.multiProcessorCount;
    if (const char* e = getenv("DRJ_COUPLE_PAR")) par = (e[0] == '1') ? (solo && R > 1 && R <= 32 && cfg->coupling_on) : (e[0] == '0' ? false : par);
    a.couple_par = par ? 1 : 0;
    s->small = a.draws || a.couple_par;
  }
  CUDA_TRY(upload(s, &a.block_ptr, cfg->block_ptr, D + 1));
  CUDA_TRY(upload(s, &a.block_idx, cfg->block_idx, cfg->block_ptr[D]));
  CUDA_TRY(upload(s, &a.beta_raised, cfg->beta_raised, R));
  CUDA_TRY(upload(s, &a.theta_init, cfg->theta_init, (size_t)C * D));
  CUDA_TRY(cudaStreamSynchronize(s->stream));  // free_slot is a local
  CUDA_TRY(upload_list(s, data, &a.data));
  CUDA_TRY(upload_list(s, misc, &a.misc));
  if (cfg->rng_mode == DRJ_RNG_REPLAY) {
    CUDA_TRY(upload(s, &a.replay, cfg->replay_uniforms, (size_t)C * s->T_total * U));
    a.replay_T = s->T_total; a.replay_U = U;
  }
  // ---- state
  CUDA_TRY(dalloc(s, &a.theta, (size_t)D * P));
  CUDA_TRY(dalloc(s, &a.phi, (size_t)D * P));
  CUDA_TRY(dalloc(s, &a.logbw, (size_t)D * P));
  CUDA_TRY(dalloc(s, &a.bw_index, (size_t)D * P));
  CUDA_TRY(dalloc(s, &a.adj_lo, (size_t)D * P));
  CUDA_TRY(dalloc(s, &a.adj_hi, (size_t)D * P));
  CUDA_TRY(dalloc(s, &a.ll_block, (size_t)B * P));
  CUDA_TRY(dalloc(s, &a.ll, (size_t)P));
  CUDA_TRY(dalloc(s, &a.lp, (size_t)P));
  CUDA_TRY(dalloc(s, &a.accept, (size_t)P));
  CUDA_TRY(dalloc(s, &s->acc_b, (size_t)P));
  CUDA_TRY(dalloc(s, &s->mc_b, (size_t)C * std::max(R - 1, 1)));
  CUDA_TRY(dalloc(s, &s->mc_s, (size_t)C * std::max(R - 1, 1)));
  CUDA_TRY(cudaMemsetAsync(s->mc_b, 0, sizeof(int) * (size_t)C * std::max(R - 1, 1), s->stream));
  CUDA_TRY(cudaMemsetAsync(s->mc_s, 0, sizeof(int) * (size_t)C * std::max(R - 1, 1), s->stream));
  CUDA_TRY(cudaMemsetAsync(s->acc_b, 0, sizeof(int) * (size_t)P, s->stream));
  a.mc_accept = s->mc_b;
  CUDA_TRY(dalloc(s, &a.err_code, (size_t)P));
  CUDA_TRY(dalloc(s, &a.err_theta, (size_t)D * P));
  CUDA_TRY(dalloc(s, &s->err_key, 1));
  a.err_key = s->err_key;
  if (s->gridk) {
    CUDA_TRY(dalloc(s, &a.grid_part, (size_t)2 * DRJ_GRID_VRANKS * P));
    CUDA_TRY(dalloc(s, &s->grid_bar, 1));
    a.grid_bar = s->grid_bar;
  }
  CUDA_TRY(cudaMemsetAsync(a.err_code, 0, sizeof(int) * (size_t)P, s->stream));
  CUDA_TRY(cudaMemsetAsync(s->err_key, 0xFF, sizeof(unsigned long long), s->stream));
  // ---- rings and final outputs
  CUDA_TRY(dalloc(s, &a.ring_ll, (size_t)tb * s->Ppt));
  CUDA_TRY(dalloc(s, &a.ring_lp, (size_t)tb * s->Ppt));
  CUDA_TRY(dalloc(s, &a.ring_theta, (size_t)tb * D * s->Pth));
  CUDA_TRY(dalloc(s, &s->fin_ll_b, (size_t)s->Ppt * cfg->burnin));
  CUDA_TRY(dalloc(s, &s->fin_lp_b, (size_t)s->Ppt * cfg->burnin));
  CUDA_TRY(dalloc(s, &s->fin_th_b, (size_t)s->Pth * cfg->burnin * D));
  CUDA_TRY(dalloc(s, &s->fin_ll_s, (size_t)s->Ppt * cfg->samples));
  CUDA_TRY(dalloc(s, &s->fin_lp_s, (size_t)s->Ppt * cfg->samples));
  CUDA_TRY(dalloc(s, &s->fin_th_s, (size_t)s->Pth * cfg->samples * D));

  // ---- Particle::init + init_like for every chain x rung, row 0 of the burn-in outputs
  a.n_iter = 1; a.iter0 = 0;
  CUDA_TRY(dalloc(s, &a.init_llb, (size_t)B * C));
  CUDA_TRY(dalloc(s, &a.init_lp, (size_t)C));
  int rc = (s->team || s->gridk) ? DRJ_OK : launch_init_chain(s, err);
  if (rc) return rc;
  rc = launch_model_kernel(s, false, err);
  if (rc) return rc;
  s->aux_launches++;
  if ((rc = flush_ring(s, 1, false, 0, err))) return rc;
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  if ((rc = check_device_error(s, 0, err))) return rc;
  if (cfg->burnin == 1) a.mc_accept = s->mc_s;  // no burn-in updates: straight into the sampling phase
  return DRJ_OK;
}

extern "C" int drj_session_create(const drj_config* cfg, const drj_model* model, const drj_list* data,
                                  const drj_list* misc, drj_session** out, drj_error* err) {
  if (!out) return set_err(err, DRJ_ERR_ARG, "NULL out");
  *out = nullptr;
  int rc = validate(cfg, model, data, err);
  if (rc) return rc;
  drj_session* s = new drj_session();
  rc = session_create_impl(s, cfg, model, data, misc, err);
  if (rc) { drj_session_destroy(s); return rc; }
  *out = s;
  return DRJ_OK;
}

extern "C" int drj_session_advance(drj_session* s, int n_iter, float* kernel_ms, drj_error* err) {
  if (!s) return set_err(err, DRJ_ERR_ARG, "NULL session");
  if (s->failed) return set_err(err, DRJ_ERR_ARG, "session already failed");
  if (kernel_ms) *kernel_ms = 0.f;
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  const int nb = s->cfg.burnin - 1;  // burn-in updates
  int left = std::min(n_iter, s->T_total - s->t_done);
  while (left > 0) {
    const bool sampling = s->t_done >= nb;
    const int phase_left = sampling ? (s->T_total - s->t_done) : (nb - s->t_done);
    const int nt = std::min(std::min(left, phase_left), s->tb);
    s->a.n_iter = nt;
    s->a.iter0 = s->t_done + 1;
    s->a.mc_accept = sampling ? s->mc_s : s->mc_b;
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    int rc = launch_model_kernel(s, true, err);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    s->sweep_launches++;
    // burn-in row index = update index (row 0 = initial values); sampling column = t - burnin
    const int col0 = sampling ? (s->t_done - nb) : (s->t_done + 1);
    if ((rc = flush_ring(s, nt, sampling, col0, err))) return rc;
    CUDA_TRY(cudaEventRecord(s->ev2, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    {
      float ms = 0.f, ms2 = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
      CUDA_TRY(cudaEventElapsedTime(&ms2, s->ev0, s->ev2));
      if (kernel_ms) *kernel_ms += ms;
      s->sweep_ms_total += ms;
      s->stream_ms_total += ms2;
    }
    if ((rc = check_device_error(s, s->t_done + 1, err))) return rc;
    s->t_done += nt;
    left -= nt;
    if (!sampling && s->t_done == nb) {
      // end of burn-in: keep the counts, reset accept_count of all rungs (main.cpp:208-210)
      CUDA_TRY(cudaMemcpyAsync(s->acc_b, s->a.accept, sizeof(int) * (size_t)s->P, cudaMemcpyDeviceToDevice, s->stream));
      CUDA_TRY(cudaMemsetAsync(s->a.accept, 0, sizeof(int) * (size_t)s->P, s->stream));
    }
  }
  return DRJ_OK;
}

extern "C" int drj_session_iterations_done(const drj_session* s) { return s ? s->t_done + 1 : 0; }

extern "C" int drj_session_launches(const drj_session* s, int64_t* sweep_launches, int64_t* aux_launches) {
  if (!s) return DRJ_ERR_ARG;
  if (sweep_launches) *sweep_launches = s->sweep_launches;
  if (aux_launches) *aux_launches = s->aux_launches;
  return DRJ_OK;
}

extern "C" int drj_session_timing(const drj_session* s, double* sweep_ms_total, double* stream_ms_total) {
  if (!s) return DRJ_ERR_ARG;
  if (sweep_ms_total) *sweep_ms_total = s->sweep_ms_total;
  if (stream_ms_total) *stream_ms_total = s->stream_ms_total;
  return DRJ_OK;
}

// D = -2 loglike over `C` chains of `S` draws each: mean and sample variance from the per-chain moments, chains in order
static void deviance_from_chain_moments(const double* lm, const double* lv, long long C, int S, double* out2) {
  const long long N = C * S;
  double gm = 0.0;
  for (long long c = 0; c < C; ++c) gm += lm[c];
  gm /= (double)C;
  double m2 = 0.0;
  for (long long c = 0; c < C; ++c) m2 += lv[c] * (S - 1) + (double)S * (lm[c] - gm) * (lm[c] - gm);
  out2[0] = -2.0 * gm;
  out2[1] = N > 1 ? 4.0 * m2 / (double)(N - 1) : 0.0;
}

// ---- quantile select, shared by the one-device and the multi-device summaries
struct QTarget { int param; long long rank; unsigned long long prefix; long long below; };  // `below` = draws with a smaller prefix

// one histogram pass on one device: groups = distinct (param, prefix) pairs, ordered by param
static int session_quantile_pass(drj_session* s, int bits_done, int nbits, const std::vector<unsigned long long>& prefix,
                                 const std::vector<int>& group_ptr, std::vector<unsigned long long>& hist, drj_error* err) {
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  const int D = s->D, C = s->C, S = s->cfg.samples;
  const int thR = s->cfg.save_hot_draws ? s->R : 1;
  const double* th = s->fin_th_s + (long long)(thR - 1) * S * D;
  const long long th_cs = (long long)thR * S * D, th_ts = D;
  const int ng = (int)prefix.size();
  int max_per_param = 0;
  for (int k = 0; k < D; ++k) max_per_param = std::max(max_per_param, group_ptr[k + 1] - group_ptr[k]);
  double *d_prefix = nullptr, *d_ptr = nullptr, *d_hist = nullptr;
  CUDA_TRY(scratch_get(s, 9, sizeof(unsigned long long) * (size_t)std::max(ng, 1), &d_prefix));
  CUDA_TRY(scratch_get(s, 10, sizeof(int) * (size_t)(D + 1), &d_ptr));
  CUDA_TRY(scratch_get(s, 11, sizeof(unsigned long long) * (size_t)std::max(ng, 1) * DRJ_QBINS, &d_hist));
  CUDA_TRY(cudaMemcpyAsync(d_prefix, prefix.data(), sizeof(unsigned long long) * (size_t)ng, cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(d_ptr, group_ptr.data(), sizeof(int) * (size_t)(D + 1), cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * (size_t)ng * DRJ_QBINS, s->stream));
  const size_t smem = sizeof(unsigned int) * (size_t)max_per_param * DRJ_QBINS;
  CUDA_TRY(cudaFuncSetAttribute(drj_quantile_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 49152)));
  const int per_sm = smem > 100 * 1024 ? 1 : 2;
  const int nb = std::max(1, std::min(C, prop_sms(s) * per_sm));
  for (int k0 = 0; k0 < D; k0 += 65535) {
    const int nk = std::min(D - k0, 65535);
    drj_quantile_hist<<<dim3(nb, nk), 256, smem, s->stream>>>(th + k0, th_cs, th_ts, S, C, bits_done, nbits, (const unsigned long long*)d_prefix,
                                                             (const int*)d_ptr + k0, (unsigned long long*)d_hist);
    s->aux_launches++;
  }
  CUDA_TRY(cudaGetLastError());
  hist.assign((size_t)ng * DRJ_QBINS, 0ULL);
  CUDA_TRY(cudaMemcpyAsync(hist.data(), d_hist, sizeof(unsigned long long) * hist.size(), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  return DRJ_OK;
}

// R's quantile(x, probs) (type 7: h = (N - 1) p, x[floor h] + (h - floor h)(x[floor h + 1] - x[floor h])) of every
// parameter over all draws of all sessions; `pass` runs one histogram pass on every session and returns the sum
template <class PassFn>
static int select_quantiles(int D, long long N, const double* probs, int nq, double* out, PassFn pass, drj_error* err) {
  if (nq <= 0 || !probs || !out) return DRJ_OK;
  if (nq > 8) return set_err(err, DRJ_ERR_ARG, "at most 8 quantile probabilities per call");
  for (int j = 0; j < nq; ++j)
    if (!(probs[j] >= 0.0 && probs[j] <= 1.0)) return set_err(err, DRJ_ERR_ARG, "quantile probabilities must be in [0,1]");
  std::vector<QTarget> tg;  // two order statistics per probability
  std::vector<double> frac((size_t)nq);
  for (int k = 0; k < D; ++k)
    for (int j = 0; j < nq; ++j) {
      const double h = (double)(N - 1) * probs[j];
      const long long lo = (long long)std::floor(h);
      frac[j] = h - (double)lo;
      tg.push_back({k, lo, 0ULL, 0});
      tg.push_back({k, std::min(lo + 1, N - 1), 0ULL, 0});
    }
  static const int digit_bits[6] = {11, 11, 11, 11, 11, 9};
  int bits_done = 0;
  for (int pass_i = 0; pass_i < 6; ++pass_i) {
    const int nbits = digit_bits[pass_i];
    // distinct (param, prefix) groups, ordered by parameter
    std::vector<unsigned long long> prefix;
    std::vector<int> group_ptr((size_t)D + 1, 0), group_of(tg.size(), -1);
    for (int k = 0; k < D; ++k) {
      group_ptr[k] = (int)prefix.size();
      for (size_t i = 0; i < tg.size(); ++i) {
        if (tg[i].param != k) continue;
        int g = -1;
        for (int q = group_ptr[k]; q < (int)prefix.size(); ++q) if (prefix[q] == tg[i].prefix) { g = q; break; }
        if (g < 0) { g = (int)prefix.size(); prefix.push_back(tg[i].prefix); }
        group_of[i] = g;
      }
    }
    group_ptr[D] = (int)prefix.size();
    std::vector<unsigned long long> hist;
    int rc = pass(bits_done, nbits, prefix, group_ptr, hist);
    if (rc) return rc;
    for (size_t i = 0; i < tg.size(); ++i) {
      const unsigned long long* h = hist.data() + (size_t)group_of[i] * DRJ_QBINS;
      long long cum = tg[i].below;
      int bin = 0;
      const int nbins = 1 << nbits;
      for (; bin < nbins; ++bin) {
        if (cum + (long long)h[bin] > tg[i].rank) break;
        cum += (long long)h[bin];
      }
      if (bin == nbins) return set_err(err, DRJ_ERR_CUDA, "quantile select lost rank %lld (histogram total %lld)", tg[i].rank, cum);
      tg[i].below = cum;
      tg[i].prefix = (tg[i].prefix << nbits) | (unsigned long long)bin;
    }
    bits_done += nbits;
  }
  auto value_of = [](unsigned long long key) {
    const unsigned long long b = (key >> 63) ? (key & 0x7fffffffffffffffULL) : ~key;
    double v;
    memcpy(&v, &b, sizeof(v));
    return v;
  };
  for (int k = 0; k < D; ++k)
    for (int j = 0; j < nq; ++j) {
      const double lo = value_of(tg[(size_t)(k * nq + j) * 2].prefix), hi = value_of(tg[(size_t)(k * nq + j) * 2 + 1].prefix);
      out[(size_t)k * nq + j] = lo + frac[j] * (hi - lo);
    }
  return DRJ_OK;
}

// everything but the quantiles; lm_out / lv_out (may be NULL) receive the per-chain mean and variance of the
// cold-rung sampling log-likelihood, from which the deviance moments of any set of chains follow exactly
static int session_summary_core(drj_session* s, drj_summary* sm, std::vector<double>* lm_out, std::vector<double>* lv_out, drj_error* err) {
  if (!s || !sm) return set_err(err, DRJ_ERR_ARG, "NULL session or summary");
  if (sm->max_lag < 0 || sm->max_lag > 4096) return set_err(err, DRJ_ERR_ARG, "max_lag must be in [0, 4096]");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  const int D = s->D, C = s->C, S = s->cfg.samples;
  const int thR = s->cfg.save_hot_draws ? s->R : 1, ptR = s->cfg.store_pt ? s->R : 1;
  // cold rung, sampling phase: theta[(c*thR + thR-1)*S + t][k], loglike[(c*ptR + ptR-1)*S + t]
  const double* th = s->fin_th_s + (long long)(thR - 1) * S * D;
  const long long th_cs = (long long)thR * S * D, th_ts = D;
  const double* ll = s->fin_ll_s + (long long)(ptR - 1) * S;
  const long long ll_cs = (long long)ptR * S;
  const long long N = (long long)C * S;
  const int L = sm->max_lag;
  double *d_mean = nullptr, *d_var = nullptr, *d_part = nullptr, *d_ac = nullptr, *d_lm = nullptr, *d_lv = nullptr;
  const int nchunk = L / DRJ_LAGCH + 1;
  const long long ntiles = (N + DRJ_AC_TILE - 1) / DRJ_AC_TILE;
  const int nb = (int)std::max<long long>(1, std::min<long long>(ntiles, 4096 / ((long long)nchunk * D) + 1));
  double* d_ctr = nullptr;
  cudaError_t e = cudaSuccess;
  std::vector<double> mean((size_t)C * D), var((size_t)C * D), lm(C), lv(C), ac((size_t)D * (L + 1)), ctr(D);
#define SUM_TRY(call) do { e = (call); if (e != cudaSuccess) { CUDA_TRY(e); } } while (0)
  SUM_TRY(scratch_get(s, 0, sizeof(double) * (size_t)C * D, &d_mean));
  SUM_TRY(scratch_get(s, 1, sizeof(double) * (size_t)C * D, &d_var));
  SUM_TRY(scratch_get(s, 2, sizeof(double) * (size_t)C, &d_lm));
  SUM_TRY(scratch_get(s, 3, sizeof(double) * (size_t)C, &d_lv));
  SUM_TRY(scratch_get(s, 4, sizeof(double) * (size_t)D * nchunk * nb * DRJ_LAGCH, &d_part));
  SUM_TRY(scratch_get(s, 5, sizeof(double) * (size_t)D * (L + 1), &d_ac));
  SUM_TRY(scratch_get(s, 6, sizeof(double) * (size_t)D, &d_ctr));
  // only what the caller asked for (each part is a pass over the stored draws)
  const bool want_theta_moments = sm->chain_mean || sm->chain_var || (sm->autocov_sum && !sm->center);
  const bool want_dev = sm->deviance_mean_var || lm_out || lv_out;
  if (want_theta_moments) {
    drj_chain_moments<<<dim3(C, D), 256, 0, s->stream>>>(th, th_cs, th_ts, 1, S, D, d_mean, d_var);
    s->aux_launches++;
    SUM_TRY(cudaMemcpyAsync(mean.data(), d_mean, sizeof(double) * mean.size(), cudaMemcpyDeviceToHost, s->stream));
    SUM_TRY(cudaMemcpyAsync(var.data(), d_var, sizeof(double) * var.size(), cudaMemcpyDeviceToHost, s->stream));
  }
  if (want_dev) {
    drj_chain_moments<<<dim3(C, 1), 256, 0, s->stream>>>(ll, ll_cs, 1, 0, S, 1, d_lm, d_lv);
    s->aux_launches++;
    SUM_TRY(cudaMemcpyAsync(lm.data(), d_lm, sizeof(double) * lm.size(), cudaMemcpyDeviceToHost, s->stream));
    SUM_TRY(cudaMemcpyAsync(lv.data(), d_lv, sizeof(double) * lv.size(), cudaMemcpyDeviceToHost, s->stream));
  }
  SUM_TRY(cudaStreamSynchronize(s->stream));
  if (sm->chain_mean) memcpy(sm->chain_mean, mean.data(), sizeof(double) * mean.size());
  if (sm->chain_var) memcpy(sm->chain_var, var.data(), sizeof(double) * var.size());
  if (lm_out) *lm_out = lm;
  if (lv_out) *lv_out = lv;
  if (sm->deviance_mean_var) deviance_from_chain_moments(lm.data(), lv.data(), C, S, sm->deviance_mean_var);
  if (sm->autocov_sum) {
    for (int k = 0; k < D; ++k) {
      if (sm->center) ctr[k] = sm->center[k];
      else { double t = 0.0; for (int c = 0; c < C; ++c) t += mean[(size_t)c * D + k]; ctr[k] = t / C; }
    }
    SUM_TRY(cudaMemcpyAsync(d_ctr, ctr.data(), sizeof(double) * (size_t)D, cudaMemcpyHostToDevice, s->stream));
    for (int k0 = 0; k0 < D; k0 += 65535) {
      const int nk = std::min(D - k0, 65535);
      drj_autocov_partial<<<dim3(nchunk * nk, nb), 256, 0, s->stream>>>(th + k0, th_cs, th_ts, S, N, d_ctr + k0, ntiles, nchunk,
                                                                        d_part + (size_t)k0 * nchunk * nb * DRJ_LAGCH);
      drj_reduce_partials<<<dim3(nchunk, nk), 32, 0, s->stream>>>(d_part + (size_t)k0 * nchunk * nb * DRJ_LAGCH, nb,
                                                                  d_ac + (size_t)k0 * (L + 1), L);
      s->aux_launches += 2;
    }
    SUM_TRY(cudaMemcpyAsync(ac.data(), d_ac, sizeof(double) * ac.size(), cudaMemcpyDeviceToHost, s->stream));
    SUM_TRY(cudaStreamSynchronize(s->stream));
    memcpy(sm->autocov_sum, ac.data(), sizeof(double) * ac.size());
  }
  if ((sm->head || sm->tail) && L > 0) {
    double *d_head = nullptr, *d_tail = nullptr;
    SUM_TRY(scratch_get(s, 7, sizeof(double) * (size_t)D * L, &d_head));
    SUM_TRY(scratch_get(s, 8, sizeof(double) * (size_t)D * L, &d_tail));
    drj_series_ends<<<(L + 127) / 128, 128, 0, s->stream>>>(th, th_cs, th_ts, S, N, D, L, d_head, d_tail);
    s->aux_launches++;
    std::vector<double> hh((size_t)D * L), tt((size_t)D * L);
    e = cudaMemcpyAsync(hh.data(), d_head, sizeof(double) * hh.size(), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tt.data(), d_tail, sizeof(double) * tt.size(), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    SUM_TRY(e);
    if (sm->head) memcpy(sm->head, hh.data(), sizeof(double) * hh.size());
    if (sm->tail) memcpy(sm->tail, tt.data(), sizeof(double) * tt.size());
  }
  SUM_TRY(cudaGetLastError());
#undef SUM_TRY
  return DRJ_OK;
}

extern "C" int drj_session_summary(drj_session* s, drj_summary* sm, drj_error* err) {
  int rc = session_summary_core(s, sm, nullptr, nullptr, err);
  if (rc) return rc;
  if (sm->n_quantiles > 0 && sm->quantiles) {
    auto pass = [&](int bits_done, int nbits, const std::vector<unsigned long long>& prefix, const std::vector<int>& group_ptr,
                    std::vector<unsigned long long>& hist) { return session_quantile_pass(s, bits_done, nbits, prefix, group_ptr, hist, err); };
    rc = select_quantiles(s->D, (long long)s->C * s->cfg.samples, sm->quantile_probs, sm->n_quantiles, sm->quantiles, pass, err);
  }
  return rc;
}

// draws to the host; thin > 1 keeps the stored iterations 0, thin, 2 thin, ... of every series (gathered on the
// device into a compact buffer first), so the caller's arrays hold ceil(n / thin) rows per series
static int copy_series(drj_session* s, double* dst, const double* src, long long n_series, int n, int w, int thin, drj_error* err) {
  if (!dst) return DRJ_OK;
  if (thin <= 1) {
    CUDA_TRY(cudaMemcpy(dst, src, sizeof(double) * (size_t)n_series * n * w, cudaMemcpyDeviceToHost));
    return DRJ_OK;
  }
  const int n_keep = (n + thin - 1) / thin;
  const size_t bytes = sizeof(double) * (size_t)n_series * n_keep * w;
  double* tmp = nullptr;
  CUDA_TRY(pool_alloc(s->cfg.device, std::max<size_t>(bytes, 8), (void**)&tmp));
  const long long total = n_series * n_keep * w;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 16));
  drj_thin_rows<<<grid, 256, 0, s->stream>>>(src, tmp, n_series, n, n_keep, w, thin);
  s->aux_launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  pool_free(s->cfg.device, tmp, std::max<size_t>(bytes, 8));
  CUDA_TRY(e);
  return DRJ_OK;
}

extern "C" int drj_session_download(drj_session* s, drj_output* out, drj_error* err) {
  return drj_session_download_thinned(s, 1, out, err);
}

extern "C" int drj_session_download_thinned(drj_session* s, int thin, drj_output* out, drj_error* err) {
  if (!s || !out) return set_err(err, DRJ_ERR_ARG, "NULL session or output");
  if (thin < 1) return set_err(err, DRJ_ERR_ARG, "thin must be >= 1");
  CUDA_TRY(cudaSetDevice(s->cfg.device));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  const drj_config& c = s->cfg;
  int rc;
  if ((rc = copy_series(s, out->loglike_burnin, s->fin_ll_b, s->Ppt, c.burnin, 1, thin, err))) return rc;
  if ((rc = copy_series(s, out->logprior_burnin, s->fin_lp_b, s->Ppt, c.burnin, 1, thin, err))) return rc;
  if ((rc = copy_series(s, out->theta_burnin, s->fin_th_b, s->Pth, c.burnin, s->D, thin, err))) return rc;
  if ((rc = copy_series(s, out->loglike_sampling, s->fin_ll_s, s->Ppt, c.samples, 1, thin, err))) return rc;
  if ((rc = copy_series(s, out->logprior_sampling, s->fin_lp_s, s->Ppt, c.samples, 1, thin, err))) return rc;
  if ((rc = copy_series(s, out->theta_sampling, s->fin_th_s, s->Pth, c.samples, s->D, thin, err))) return rc;
#define D2H(dst, src, n) \
  if (dst) CUDA_TRY(cudaMemcpy(dst, src, (n), cudaMemcpyDeviceToHost))
  if (s->R > 1) {
    D2H(out->mc_accept_burnin, s->mc_b, sizeof(int) * (size_t)s->C * (s->R - 1));
    D2H(out->mc_accept_sampling, s->mc_s, sizeof(int) * (size_t)s->C * (s->R - 1));
  }
  D2H(out->accept_burnin, s->acc_b, sizeof(int) * (size_t)s->P);
  D2H(out->accept_sampling, s->a.accept, sizeof(int) * (size_t)s->P);
#undef D2H
  if (out->bw_final) {
    double* tmp = nullptr;
    CUDA_TRY(pool_alloc(s->cfg.device, sizeof(double) * (size_t)s->P * s->D, (void**)&tmp));
    drj_export_bw<<<(unsigned)((s->P + 255) / 256), 256, 0, s->stream>>>(s->a.logbw, s->P, s->D, tmp);
    s->aux_launches++;
    cudaError_t e = cudaMemcpyAsync(out->bw_final, tmp, sizeof(double) * (size_t)s->P * s->D, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    pool_free(s->cfg.device, tmp, sizeof(double) * (size_t)s->P * s->D);
    CUDA_TRY(e);
  }
  if (out->t_diff) {
    const double dt = (wall_now() - s->t_create) / std::max(1, s->C);
    for (int ch = 0; ch < s->C; ++ch) out->t_diff[ch] = dt;
  }
  return DRJ_OK;
}

// ------------------------------------------------------------------------------------------------
// the drop-in entry
// ------------------------------------------------------------------------------------------------
extern "C" int drj_run_mcmc(const drj_config* cfg, const drj_model* model, const drj_list* data, const drj_list* misc,
                            drj_output* out, drj_progress_fn progress, void* progress_user, drj_error* err) {
  if (err) memset(err, 0, sizeof(*err));
  if (!out) return set_err(err, DRJ_ERR_ARG, "NULL output");
  drj_session* s = nullptr;
  int rc = drj_session_create(cfg, model, data, misc, &s, err);
  if (rc) return rc;
  const int nb = cfg->burnin - 1;
  // burn-in then sampling, in ~1 % batches when a progress hook is given (main.cpp:181-189, 241-249)
  for (int phase = 0; phase < 2 && rc == DRJ_OK; ++phase) {
    const int total = phase == 0 ? nb : cfg->samples;
    const int step = progress ? std::max(1, (total + 99) / 100) : std::max(total, 1);
    for (int done = 0; done < total && rc == DRJ_OK;) {
      const int n = std::min(step, total - done);
      rc = drj_session_advance(s, n, nullptr, err);
      done += n;
      if (rc == DRJ_OK && progress && progress(progress_user, phase, done + (phase == 0 ? 1 : 0), phase == 0 ? cfg->burnin : cfg->samples))
        rc = set_err(err, DRJ_ERR_INTERRUPT, "interrupted by the progress callback");
    }
  }
  if (rc == DRJ_OK) rc = drj_session_download(s, out, err);
  drj_session_destroy(s);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// one process, several GPUs (replaces the PSOCK-cluster dispatch of R/main.R:392-402)
// ------------------------------------------------------------------------------------------------
struct drj_multi {
  drj_config cfg;                       // the whole run
  std::vector<int> devices;             // one per shard
  std::vector<int> first;               // first chain of shard k; first[n] = chains
  std::vector<drj_session*> sessions;
  int D = 0, R = 0;
  long long U = 0, T = 0;
  double t_create = 0.0;
};

// run fn(k) for every shard on its own host thread (shard 0 on the calling thread); returns the error of the shard
// whose failure comes first in the reference's order (lowest iteration, then lowest chain)
template <class Fn>
static int for_each_shard(int n, drj_error* err, Fn fn) {
  std::vector<int> rc((size_t)n, DRJ_OK);
  std::vector<drj_error> errs((size_t)n);
  for (auto& e : errs) memset(&e, 0, sizeof(e));
  std::vector<std::thread> th;
  for (int k = 1; k < n; ++k) th.emplace_back([&, k] { rc[k] = fn(k, &errs[k]); });
  rc[0] = fn(0, &errs[0]);
  for (auto& t : th) t.join();
  int best = -1;
  for (int k = 0; k < n; ++k) {
    if (rc[k] == DRJ_OK) continue;
    if (best < 0 || errs[k].iteration < errs[best].iteration ||
        (errs[k].iteration == errs[best].iteration && errs[k].chain < errs[best].chain))
      best = k;
  }
  if (best < 0) return DRJ_OK;
  if (err) *err = errs[best];
  return rc[best];
}

static void shard_output(const drj_multi* m, int k, int thin, const drj_output* out, drj_output* o) {
  const drj_config& c = m->cfg;
  const long long c0 = m->first[k];
  const long long ptR = c.store_pt ? c.rungs : 1, thR = c.save_hot_draws ? c.rungs : 1;
  const long long nb = (c.burnin + thin - 1) / thin, ns = (c.samples + thin - 1) / thin;
  memset(o, 0, sizeof(*o));
#define OFF(field, stride) o->field = out->field ? out->field + c0 * (stride) : nullptr
  OFF(loglike_burnin, ptR * nb); OFF(logprior_burnin, ptR * nb); OFF(theta_burnin, thR * nb * c.d);
  OFF(loglike_sampling, ptR * ns); OFF(logprior_sampling, ptR * ns); OFF(theta_sampling, thR * ns * c.d);
  OFF(mc_accept_burnin, (long long)(c.rungs - 1)); OFF(mc_accept_sampling, (long long)(c.rungs - 1));
  OFF(accept_burnin, (long long)c.rungs); OFF(accept_sampling, (long long)c.rungs);
  OFF(t_diff, 1LL); OFF(bw_final, (long long)c.rungs * c.d);
#undef OFF
}

extern "C" void drj_multi_destroy(drj_multi* m) {
  if (!m) return;
  for (drj_session* s : m->sessions) drj_session_destroy(s);
  delete m;
}

extern "C" int drj_multi_device_count(const drj_multi* m) { return m ? (int)m->sessions.size() : 0; }

extern "C" int drj_multi_create(const drj_config* cfg, const int32_t* devices, int32_t n_devices, const drj_model* model,
                                const drj_list* data, const drj_list* misc, drj_multi** out, drj_error* err) {
  if (err) memset(err, 0, sizeof(*err));
  if (!out) return set_err(err, DRJ_ERR_ARG, "NULL out");
  *out = nullptr;
  int rc = validate(cfg, model, data, err);
  if (rc) return rc;
  std::vector<int> devs;
  if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
  else {
    const int n = drj_device_count();
    if (n < 1) return set_err(err, DRJ_ERR_CUDA, "no CUDA device");
    for (int i = 0; i < n; ++i) devs.push_back(i);
  }
  const int visible = drj_device_count();
  // (a device may be listed more than once: its chain ranges then run side by side on it)
  for (size_t i = 0; i < devs.size(); ++i)
    if (devs[i] < 0 || devs[i] >= visible) return set_err(err, DRJ_ERR_CUDA, "no CUDA device %d", devs[i]);
  const int n = (int)std::min<long long>((long long)devs.size(), cfg->chains);  // no empty shards: spare devices stay idle
  drj_multi* m = new drj_multi();
  m->cfg = *cfg;
  m->t_create = wall_now();
  m->D = cfg->d; m->R = cfg->rungs;
  m->devices.assign(devs.begin(), devs.begin() + n);
  m->first.resize((size_t)n + 1);
  for (int k = 0; k <= n; ++k) {  // contiguous ranges, sizes differing by at most one
    const long long base = cfg->chains / n, rem = cfg->chains % n;
    m->first[k] = (int)(k * base + std::min<long long>(k, rem));
  }
  int d_free = 0;
  for (int i = 0; i < cfg->d; ++i) d_free += cfg->skip_param[i] ? 0 : 1;
  m->U = 3LL * d_free * cfg->rungs + ((cfg->coupling_on && cfg->rungs > 1) ? (cfg->rungs - 1) : 0);
  m->T = cfg->burnin - 1 + cfg->samples;
  m->sessions.assign((size_t)n, nullptr);
  rc = for_each_shard(n, err, [&](int k, drj_error* e) {
    drj_config c = *cfg;
    c.device = m->devices[k];
    c.chains = m->first[k + 1] - m->first[k];
    c.chain_offset = cfg->chain_offset + m->first[k];
    c.layout_chains = std::max(cfg->layout_chains, cfg->chains);
    c.theta_init = cfg->theta_init + (long long)m->first[k] * cfg->d;
    if (cfg->rng_mode == DRJ_RNG_REPLAY) c.replay_uniforms = cfg->replay_uniforms + (long long)m->first[k] * m->T * m->U;
    return drj_session_create(&c, model, data, misc, &m->sessions[k], e);
  });
  if (rc) { drj_multi_destroy(m); return rc; }
  *out = m;
  return DRJ_OK;
}

extern "C" int drj_multi_advance(drj_multi* m, int n_iter, float* kernel_ms, drj_error* err) {
  if (!m) return set_err(err, DRJ_ERR_ARG, "NULL multi-device run");
  const int n = (int)m->sessions.size();
  std::vector<float> ms((size_t)n, 0.f);
  int rc = for_each_shard(n, err, [&](int k, drj_error* e) { return drj_session_advance(m->sessions[k], n_iter, &ms[k], e); });
  if (kernel_ms) *kernel_ms = *std::max_element(ms.begin(), ms.end());
  return rc;
}

extern "C" int drj_multi_download_thinned(drj_multi* m, int thin, drj_output* out, drj_error* err) {
  if (!m || !out) return set_err(err, DRJ_ERR_ARG, "NULL multi-device run or output");
  if (thin < 1) return set_err(err, DRJ_ERR_ARG, "thin must be >= 1");
  int rc = for_each_shard((int)m->sessions.size(), err, [&](int k, drj_error* e) {
    drj_output o;
    shard_output(m, k, thin, out, &o);
    return drj_session_download_thinned(m->sessions[k], thin, &o, e);
  });
  if (rc == DRJ_OK && out->t_diff) {  // whole-call wall time / chains, as the one-device entry reports it
    const double dt = (wall_now() - m->t_create) / std::max(1, m->cfg.chains);
    for (int ch = 0; ch < m->cfg.chains; ++ch) out->t_diff[ch] = dt;
  }
  return rc;
}

extern "C" int drj_multi_launches(const drj_multi* m, int64_t* sweep_launches, int64_t* aux_launches) {
  if (!m) return DRJ_ERR_ARG;
  int64_t a = 0, b = 0;
  for (const drj_session* s : m->sessions) { a += s->sweep_launches; b += s->aux_launches; }
  if (sweep_launches) *sweep_launches = a;
  if (aux_launches) *aux_launches = b;
  return DRJ_OK;
}

extern "C" int drj_multi_summary(drj_multi* m, drj_summary* sm, drj_error* err) {
  if (!m || !sm) return set_err(err, DRJ_ERR_ARG, "NULL multi-device run or summary");
  if (sm->max_lag < 0 || sm->max_lag > 4096) return set_err(err, DRJ_ERR_ARG, "max_lag must be in [0, 4096]");
  const int n = (int)m->sessions.size(), D = m->D, L = sm->max_lag, S = m->cfg.samples;
  const long long C = m->cfg.chains;
  // ---- pass 1: per-chain moments of theta and of the log-likelihood, every device into its slice
  const bool need_mean = sm->chain_mean || sm->chain_var || (sm->autocov_sum && !sm->center);
  std::vector<double> mean, var, lm((size_t)C), lv((size_t)C);
  if (need_mean) { mean.resize((size_t)C * D); var.resize((size_t)C * D); }
  const bool want_dev = sm->deviance_mean_var != nullptr;
  if (need_mean || want_dev) {
    int rc = for_each_shard(n, err, [&](int k, drj_error* e) {
      drj_summary one;
      memset(&one, 0, sizeof(one));
      if (need_mean) { one.chain_mean = mean.data() + (size_t)m->first[k] * D; one.chain_var = var.data() + (size_t)m->first[k] * D; }
      std::vector<double> a, b;
      int r = session_summary_core(m->sessions[k], &one, want_dev ? &a : nullptr, want_dev ? &b : nullptr, e);
      if (r == DRJ_OK && want_dev) {
        std::copy(a.begin(), a.end(), lm.begin() + m->first[k]);
        std::copy(b.begin(), b.end(), lv.begin() + m->first[k]);
      }
      return r;
    });
    if (rc) return rc;
  }
  if (sm->chain_mean) memcpy(sm->chain_mean, mean.data(), sizeof(double) * mean.size());
  if (sm->chain_var) memcpy(sm->chain_var, var.data(), sizeof(double) * var.size());
  if (want_dev) deviance_from_chain_moments(lm.data(), lv.data(), C, S, sm->deviance_mean_var);
  // ---- pass 2: autocovariance sums about the GLOBAL centre (the same expression as the one-device path: mean of the
  // chain means in chain order), added over the devices in shard order, plus the lag products that straddle two shards
  if (sm->autocov_sum || ((sm->head || sm->tail) && L > 0)) {
    std::vector<double> ctr((size_t)D, 0.0);
    for (int k = 0; k < D; ++k) {
      if (sm->center) ctr[k] = sm->center[k];
      else if (!mean.empty()) { double t = 0.0; for (long long c = 0; c < C; ++c) t += mean[(size_t)c * D + k]; ctr[k] = t / (double)C; }
    }
    const size_t nac = (size_t)D * (L + 1), nht = (size_t)D * std::max(L, 1);
    std::vector<double> ac((size_t)n * nac, 0.0), hd((size_t)n * nht, 0.0), tl((size_t)n * nht, 0.0);
    int rc = for_each_shard(n, err, [&](int k, drj_error* e) {
      drj_summary one;
      memset(&one, 0, sizeof(one));
      one.max_lag = L;
      one.center = ctr.data();
      if (sm->autocov_sum) one.autocov_sum = ac.data() + (size_t)k * nac;
      if (L > 0) { one.head = hd.data() + (size_t)k * nht; one.tail = tl.data() + (size_t)k * nht; }
      return session_summary_core(m->sessions[k], &one, nullptr, nullptr, e);
    });
    if (rc) return rc;
    if (sm->autocov_sum) {
      for (size_t i = 0; i < nac; ++i) {
        double t = 0.0;
        for (int k = 0; k < n; ++k) t += ac[(size_t)k * nac + i];
        sm->autocov_sum[i] = t;
      }
      // x_j among the last `lag` draws of shard k pairs with x_{j+lag} among the first `lag` draws of shard k+1
      // (a shard holds at least `samples` >= 1 draws; lags longer than a shard would reach a third shard: the
      // concatenated window below covers that too)
      for (int p = 0; p < D && L > 0; ++p)
        for (int k = 0; k + 1 < n; ++k) {
          // window = last min(L, len_k) draws of shards <= k (walking back) | first min(L, len) draws of shards > k
          std::vector<double> left, right;
          for (int q = k; q >= 0 && (int)left.size() < L; --q) {
            const long long len = (long long)(m->first[q + 1] - m->first[q]) * S;
            const int take = (int)std::min<long long>(L - (long long)left.size(), std::min<long long>(len, L));
            const double* t = tl.data() + (size_t)q * nht + (size_t)p * L;  // last L draws of shard q (zero-padded at the front if shorter)
            for (int i = 0; i < take; ++i) left.push_back(t[L - 1 - i] - ctr[p]);  // reversed: left[0] is the draw just before the boundary
          }
          for (int q = k + 1; q < n && (int)right.size() < L; ++q) {
            const long long len = (long long)(m->first[q + 1] - m->first[q]) * S;
            const int take = (int)std::min<long long>(L - (long long)right.size(), std::min<long long>(len, L));
            const double* h = hd.data() + (size_t)q * nht + (size_t)p * L;
            for (int i = 0; i < take; ++i) right.push_back(h[i] - ctr[p]);
          }
          // pairs (left[a], right[b]) at lag a + b + 1; only pairs whose LEFT element lies in shard k itself are new at
          // this boundary (pairs reaching further back are counted at the earlier boundary)
          const long long len_k = (long long)(m->first[k + 1] - m->first[k]) * S;
          const int own = (int)std::min<long long>((long long)left.size(), len_k);
          for (int a = 0; a < own; ++a)
            for (int b = 0; b < (int)right.size() && a + b + 1 <= L; ++b)
              sm->autocov_sum[(size_t)p * (L + 1) + a + b + 1] += left[a] * right[b];
        }
    }
    if (L > 0) {  // ends of the whole concatenated series (for a caller that shards further, e.g. over nodes)
      if (sm->head) memcpy(sm->head, hd.data(), sizeof(double) * nht);
      if (sm->tail) memcpy(sm->tail, tl.data() + (size_t)(n - 1) * nht, sizeof(double) * nht);
    }
  }
  // ---- quantiles: radix select with the histograms of all devices added per pass
  if (sm->n_quantiles > 0 && sm->quantiles) {
    auto pass = [&](int bits_done, int nbits, const std::vector<unsigned long long>& prefix, const std::vector<int>& group_ptr,
                    std::vector<unsigned long long>& hist) {
      std::vector<std::vector<unsigned long long>> part((size_t)n);
      int rc = for_each_shard(n, err, [&](int k, drj_error* e) {
        return session_quantile_pass(m->sessions[k], bits_done, nbits, prefix, group_ptr, part[k], e);
      });
      if (rc) return rc;
      hist.assign(part[0].size(), 0ULL);
      for (int k = 0; k < n; ++k)
        for (size_t i = 0; i < hist.size(); ++i) hist[i] += part[k][i];
      return (int)DRJ_OK;
    };
    int rc = select_quantiles(D, C * S, sm->quantile_probs, sm->n_quantiles, sm->quantiles, pass, err);
    if (rc) return rc;
  }
  return DRJ_OK;
}

extern "C" int drj_run_mcmc_multi(const drj_config* cfg, const int32_t* devices, int32_t n_devices, const drj_model* model,
                                  const drj_list* data, const drj_list* misc, drj_output* out, drj_progress_fn progress,
                                  void* progress_user, drj_error* err) {
  if (err) memset(err, 0, sizeof(*err));
  if (!out) return set_err(err, DRJ_ERR_ARG, "NULL output");
  drj_multi* m = nullptr;
  int rc = drj_multi_create(cfg, devices, n_devices, model, data, misc, &m, err);
  if (rc) return rc;
  const int nb = cfg->burnin - 1;
  // the progress hook runs here, on the calling thread, between batches (the worker threads of a batch have joined)
  for (int phase = 0; phase < 2 && rc == DRJ_OK; ++phase) {
    const int total = phase == 0 ? nb : cfg->samples;
    const int step = progress ? std::max(1, (total + 99) / 100) : std::max(total, 1);
    for (int done = 0; done < total && rc == DRJ_OK;) {
      const int n = std::min(step, total - done);
      rc = drj_multi_advance(m, n, nullptr, err);
      done += n;
      if (rc == DRJ_OK && progress && progress(progress_user, phase, done + (phase == 0 ? 1 : 0), phase == 0 ? cfg->burnin : cfg->samples))
        rc = set_err(err, DRJ_ERR_INTERRUPT, "interrupted by the progress callback");
    }
  }
  if (rc == DRJ_OK) rc = drj_multi_download_thinned(m, 1, out, err);
  drj_multi_destroy(m);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// FP64 peak probe
// ------------------------------------------------------------------------------------------------
extern "C" int drj_fp64_peak_tflops(int device, double* tflops, drj_error* err) {
  if (!tflops) return set_err(err, DRJ_ERR_ARG, "NULL tflops");
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return set_err(err, DRJ_ERR_CUDA, "no CUDA device %d", device); }
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  CUDA_TRY(cudaMalloc(&out, sizeof(double) * (size_t)blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    drj_fp64_probe<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) { cudaFree(out); return set_err(err, DRJ_ERR_CUDA, "probe failed: %s", cudaGetErrorString(e)); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
    if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return DRJ_OK;
}
