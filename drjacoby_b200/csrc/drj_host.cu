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
struct drj_model {
  std::string name;
  int D = 0, NBLOCK = 0;
  bool datum_form = false;
  bool jit = false;
  const void* init_fn = nullptr;   // built-in: __global__ function addresses
  const void* init_chain_fn = nullptr;  // initial likelihood / prior, one thread per chain
  const void* sweep_fn = nullptr;
  const void* stream_fn = nullptr; // sweep variant for data streamed through shared memory (datum-form models)
  const void* team_init_fn = nullptr;   // cluster-per-chain variants (datum-form models)
  const void* team_sweep_fn = nullptr;
  CUmodule module = nullptr;       // jit
  CUfunction init_cu = nullptr, sweep_cu = nullptr, stream_cu = nullptr, team_init_cu = nullptr, team_sweep_cu = nullptr,
             init_chain_cu = nullptr;
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
};

#define DRJ_REG(NAME, ...)                                                               \
  {                                                                                      \
    NAME, __VA_ARGS__::D, __VA_ARGS__::NBLOCK, __VA_ARGS__::DATUM_FORM,                  \
        (const void*)drj_init_kernel<__VA_ARGS__>, (const void*)drj_sweep_kernel<__VA_ARGS__, false>, \
        DrjStreamKernel<__VA_ARGS__>::get(), DrjStreamKernel<__VA_ARGS__>::team_init(),                \
        DrjStreamKernel<__VA_ARGS__>::team_sweep(), (const void*)drj_init_chain_kernel<__VA_ARGS__>    \
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

static std::string cache_dir() {
  const char* e = getenv("DRJ_CACHE_DIR");
  std::string d = e ? e : (std::string("/tmp/drjacoby_b200_cache_") + std::to_string((long)getuid()));
  mkdir(d.c_str(), 0700);
  return d;
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
       "  drj_init_body<DRJ_USER_MODEL, false>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT) drj_user_init_chain(const __grid_constant__ DrjKernelArgs a) {\n"
       "  drj_init_chain_body<DRJ_USER_MODEL>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT, 2) drj_user_sweep(const __grid_constant__ DrjKernelArgs a) {\n"
       "  drj_sweep_body<DRJ_USER_MODEL, false, false>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_NT, 2) drj_user_sweep_stream(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DRJ_USER_MODEL::DATUM_FORM) drj_sweep_body<DRJ_USER_MODEL, true, false>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_TEAM_NT, 1) drj_user_init_team(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DRJ_USER_MODEL::DATUM_FORM) drj_init_body<DRJ_USER_MODEL, true>(a);\n"
       "}\n"
       "extern \"C\" __global__ void __launch_bounds__(DRJ_TEAM_NT, 1) drj_user_sweep_team(const __grid_constant__ DrjKernelArgs a) {\n"
       "  if constexpr (DRJ_USER_MODEL::DATUM_FORM) drj_sweep_body<DRJ_USER_MODEL, false, true>(a);\n"
       "}\n"
       "extern \"C\" __global__ void drj_user_traits(int* out) { out[0] = DRJ_USER_MODEL::DATUM_FORM ? 1 : 0; }\n";
  return s;
}

static int jit_compile(const drj_model_desc* desc, drj_model* m, drj_error* err) {
  const double t0 = wall_now();
  const std::string src = make_jit_source(desc);
  const std::string key_text = src + "|" + EMBED_DEVICE + "|" + EMBED_SAMPLER + "|" + EMBED_MODELS + "|sm_100a|v5";
  char hex[32];
  snprintf(hex, sizeof(hex), "%016llx", (unsigned long long)fnv1a(key_text));
  const std::string path = cache_dir() + "/" + hex + ".cubin";
  std::vector<char> cubin;
  if (FILE* f = fopen(path.c_str(), "rb")) {
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n > 0) {
      cubin.resize(n);
      if (fread(cubin.data(), 1, n, f) != (size_t)n) cubin.clear();
    }
    fclose(f);
  }
  if (cubin.empty()) {
    std::string why;
    if (!load_nvrtc(why)) return set_err(err, DRJ_ERR_COMPILE, "NVRTC unavailable: %s", why.c_str());
    nvrtcProgram prog;
    const char* hdr_src[] = {EMBED_DEVICE, EMBED_SAMPLER, EMBED_MODELS};
    const char* hdr_name[] = {"drj_device.cuh", "drj_sampler.cuh", "drj_models.cuh"};
    nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.c_str(), "drj_jit_model.cu", 3, hdr_src, hdr_name);
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
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    if (FILE* f = fopen(tmp.c_str(), "wb")) {
      fwrite(cubin.data(), 1, cubin.size(), f);
      fclose(f);
      rename(tmp.c_str(), path.c_str());
    }
  }
  if (!load_driver_api()) return set_err(err, DRJ_ERR_CUDA, "CUDA driver entry points unavailable");
  CUresult cr = g_drv.ModuleLoadData(&m->module, cubin.data());
  if (cr != CUDA_SUCCESS) {
    const char* es = nullptr;
    g_drv.GetErrorString(cr, &es);
    return set_err(err, DRJ_ERR_CUDA, "cuModuleLoadData: %s", es ? es : "?");
  }
  if (g_drv.ModuleGetFunction(&m->init_cu, m->module, "drj_user_init") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&m->sweep_cu, m->module, "drj_user_sweep") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&m->stream_cu, m->module, "drj_user_sweep_stream") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&m->team_init_cu, m->module, "drj_user_init_team") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&m->team_sweep_cu, m->module, "drj_user_sweep_team") != CUDA_SUCCESS ||
      g_drv.ModuleGetFunction(&m->init_chain_cu, m->module, "drj_user_init_chain") != CUDA_SUCCESS)
    return set_err(err, DRJ_ERR_COMPILE, "kernels not found in the compiled module");
  {  // does the user's model use the per-datum form?
    CUfunction traits = nullptr;
    int* d_flag = nullptr;
    int h_flag = 0;
    if (g_drv.ModuleGetFunction(&traits, m->module, "drj_user_traits") == CUDA_SUCCESS &&
        cudaMalloc(&d_flag, sizeof(int)) == cudaSuccess) {
      void* params[] = {(void*)&d_flag};
      if (g_drv.LaunchKernel(traits, 1, 1, 1, 1, 1, 1, 0, nullptr, params, nullptr) == CUDA_SUCCESS &&
          cudaMemcpy(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess)
        m->datum_form = h_flag != 0;
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
  if (m->module && g_drv.ok) g_drv.ModuleUnload(m->module);
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
// Blocks >= 1 MB only, at most DRJ_POOL_GB (default 48) GB per process, everything is released when an
// allocation fails or drj_release_cached_memory() is called.
struct PoolBlock { void* p; size_t bytes; int device; };
static std::vector<PoolBlock> g_pool;
static size_t g_pool_bytes = 0;
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
      g_pool_bytes -= g_pool[i].bytes;
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
static cudaError_t pool_alloc(int device, size_t bytes, void** out) {
  if (bytes >= (1u << 20)) {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    for (size_t i = 0; i < g_pool.size(); ++i)
      if (g_pool[i].device == device && g_pool[i].bytes == bytes) {
        *out = g_pool[i].p;
        g_pool_bytes -= bytes;
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
  if (bytes >= (1u << 20)) {
    std::lock_guard<std::mutex> lk(g_pool_mutex);
    if (g_pool_bytes + bytes <= pool_cap()) {
      g_pool.push_back({p, bytes, device});
      g_pool_bytes += bytes;
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
  bool team = false;            // few chains x long data: cluster-per-chain kernels (drj_sweep_item_team)
  int team_ctas = 1;            // CTAs per cluster
  bool team_ctas_forced = false;
  unsigned dyn_smem = 0;        // dynamic shared memory of the model kernels: the data tile buffer
  double t_create = 0.0;
  bool failed = false;
  // summary scratch, kept for the life of the session (grow-only): cudaMalloc / cudaFree on every
  // drj_session_summary call cost up to ~0.5 s now and then next to tens of GB of stored draws
  void* scratch[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t scratch_bytes[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

static cudaError_t scratch_get(drj_session* s, int slot, size_t bytes, double** out) {
  bytes = std::max<size_t>(bytes, 8);
  if (s->scratch_bytes[slot] < bytes) {
    if (s->scratch[slot]) cudaFree(s->scratch[slot]);
    s->scratch[slot] = nullptr;
    s->scratch_bytes[slot] = 0;
    cudaError_t e = cudaMalloc(&s->scratch[slot], bytes);
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

// initial likelihood and prior, one thread per chain (not for the cluster kernels, whose init kernel does it itself)
static int launch_init_chain(drj_session* s, drj_error* err) {
  const drj_model* m = s->model;
  void* params[] = {(void*)&s->a};
  const unsigned threads = (unsigned)std::min(128, ((s->C + 31) / 32) * 32);
  const unsigned grid = (unsigned)((s->C + threads - 1) / threads);
  if (m->jit) {
    CUresult cr = g_drv.LaunchKernel(m->init_chain_cu, grid, 1, 1, threads, 1, 1, s->dyn_smem, (CUstream)s->stream, params, nullptr);
    if (cr != CUDA_SUCCESS) {
      const char* es = nullptr;
      g_drv.GetErrorString(cr, &es);
      return set_err(err, DRJ_ERR_CUDA, "cuLaunchKernel: %s", es ? es : "?");
    }
  } else {
    CUDA_TRY(cudaLaunchKernel(m->init_chain_fn, dim3(grid), dim3(threads), params, s->dyn_smem, s->stream));
  }
  s->aux_launches++;
  return DRJ_OK;
}

static int launch_model_kernel(drj_session* s, bool sweep, drj_error* err) {
  const drj_model* m = s->model;
  void* params[] = {(void*)&s->a};
  const bool stream = sweep && s->stream_variant;
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
      CUresult cr = g_drv.LaunchKernelEx(&lc, sweep ? m->team_sweep_cu : m->team_init_cu, params, nullptr);
      if (cr != CUDA_SUCCESS) {
        const char* es = nullptr;
        g_drv.GetErrorString(cr, &es);
        return set_err(err, DRJ_ERR_CUDA, "cuLaunchKernelEx: %s", es ? es : "?");
      }
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
  if (m->jit) {
    CUresult cr = g_drv.LaunchKernel(!sweep ? m->init_cu : (stream ? m->stream_cu : m->sweep_cu), s->grid, 1, 1, s->nthreads, 1, 1, s->dyn_smem,
                                     (CUstream)s->stream, params, nullptr);
    if (cr != CUDA_SUCCESS) {
      const char* es = nullptr;
      g_drv.GetErrorString(cr, &es);
      return set_err(err, DRJ_ERR_CUDA, "cuLaunchKernel: %s", es ? es : "?");
    }
  } else {
    CUDA_TRY(cudaLaunchKernel(!sweep ? m->init_fn : (stream ? m->stream_fn : m->sweep_fn), dim3(s->grid), dim3(s->nthreads),
                              params, s->dyn_smem, s->stream));
  }
  return DRJ_OK;
}

static int prop_sms(const drj_session* s) {
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, s->cfg.device);
  return n;
}

// TEAM kernels: opt in to the dynamic shared memory of the tile pipeline and to clusters of 16 CTAs, then
// shrink the cluster until the device can co-schedule at least one
static int team_configure(drj_session* s, drj_error* err) {
  const drj_model* m = s->model;
  for (int which = 0; which < 2; ++which) {
    if (m->jit) {
      CUfunction f = which ? m->team_sweep_cu : m->team_init_cu;
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
      if (g_drv.OccupancyMaxActiveClusters(&nclusters, m->team_sweep_cu, &lc) != CUDA_SUCCESS) nclusters = 0;
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
  for (void* p : s->scratch) if (p) cudaFree(p);
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
  if (model->jit && !load_driver_api()) return set_err(err, DRJ_ERR_CUDA, "CUDA driver entry points unavailable");

  // ---- launch geometry: all rungs of a chain in one CTA, as many chains per CTA as fit in DRJ_NT
  // threads, but no more than needed to give every SM a few CTAs
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
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
  if (model->datum_form && data && (model->jit ? model->stream_cu != nullptr : model->stream_fn != nullptr))
    for (int k = 0; k < data->n_items; ++k)
      if (data->lengths[k] >= DRJ_TILE) s->stream_variant = true;
  if (model->datum_form && (model->jit ? model->stream_cu != nullptr : model->stream_fn != nullptr)) {
    if (cfg->flags & DRJ_FLAG_STREAM_KERNEL) s->stream_variant = true;    // tuning / testing overrides
    if (cfg->flags & DRJ_FLAG_RESIDENT_KERNEL) s->stream_variant = false;
  }

  // few chains over a long data item -> one thread-block cluster per chain, the sweep divided over the
  // cluster's threads (drj_sweep_item_team).  With P = chains x rungs threads the one-thread-per-particle
  // kernels need ~5 cycles per datum per sweep whatever P is, until P fills the machine (~75k threads).
  int team_stages = 3;  // x 64 KB tiles
  {
    long long longest = 0;
    if (data) for (int k = 0; k < data->n_items; ++k) longest = std::max<long long>(longest, data->lengths[k]);
    const bool can = model->datum_form && data && R <= DRJ_TEAM_MAXR &&
                     (model->jit ? (model->team_sweep_cu != nullptr && g_drv.LaunchKernelEx != nullptr) : model->team_sweep_fn != nullptr);
    bool want = longest > 2LL * DRJ_TILE && P <= 4096;  // measured cross-over, tools/team_bench.py
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
    }
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
    a.warp_local = (32 % R == 0 && !streamed && !s->stream_variant && !s->team && 6 * d_free_n > R - 1) ? 1 : 0;
    if (const char* e = getenv("DRJ_WARP_LOCAL")) if (e[0] == '1' && 32 % R == 0 && !streamed && !s->stream_variant && !s->team) a.warp_local = 1;
    if (const char* e = getenv("DRJ_NO_WARP_LOCAL")) if (e[0] == '1') a.warp_local = 0;
  }
  a.chains = C; a.rungs = R; a.chains_per_cta = s->cpc; a.team_ctas = s->team_ctas; a.team_stages = team_stages; a.coupling_on = cfg->coupling_on ? 1 : 0;
  a.save_hot_draws = cfg->save_hot_draws ? 1 : 0; a.store_pt = cfg->store_pt ? 1 : 0;
  a.rng_mode = cfg->rng_mode; a.d_free = s->d_free;
  a.use_tma = (cfg->flags & DRJ_FLAG_NO_TMA) ? 0 : 1;
  if (const char* e = getenv("DRJ_NO_TMA")) if (e[0] == '1') a.use_tma = 0;
  a.chain_offset = cfg->chain_offset; a.seed = cfg->seed; a.target_acceptance = cfg->target_acceptance; a.P = P;
  CUDA_TRY(upload(s, &a.theta_min, cfg->theta_min, D));
  CUDA_TRY(upload(s, &a.theta_max, cfg->theta_max, D));
  CUDA_TRY(upload(s, &a.trans_type, cfg->trans_type, D));
  CUDA_TRY(upload(s, &a.free_slot, free_slot.data(), D));
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
  int rc = s->team ? DRJ_OK : launch_init_chain(s, err);
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

extern "C" int drj_session_summary(drj_session* s, drj_summary* sm, drj_error* err) {
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
  if (want_theta_moments) {
    drj_chain_moments<<<dim3(C, D), 256, 0, s->stream>>>(th, th_cs, th_ts, 1, S, D, d_mean, d_var);
    s->aux_launches++;
    SUM_TRY(cudaMemcpyAsync(mean.data(), d_mean, sizeof(double) * mean.size(), cudaMemcpyDeviceToHost, s->stream));
    SUM_TRY(cudaMemcpyAsync(var.data(), d_var, sizeof(double) * var.size(), cudaMemcpyDeviceToHost, s->stream));
  }
  if (sm->deviance_mean_var) {
    drj_chain_moments<<<dim3(C, 1), 256, 0, s->stream>>>(ll, ll_cs, 1, 0, S, 1, d_lm, d_lv);
    s->aux_launches++;
    SUM_TRY(cudaMemcpyAsync(lm.data(), d_lm, sizeof(double) * lm.size(), cudaMemcpyDeviceToHost, s->stream));
    SUM_TRY(cudaMemcpyAsync(lv.data(), d_lv, sizeof(double) * lv.size(), cudaMemcpyDeviceToHost, s->stream));
  }
  SUM_TRY(cudaStreamSynchronize(s->stream));
  if (sm->chain_mean) memcpy(sm->chain_mean, mean.data(), sizeof(double) * mean.size());
  if (sm->chain_var) memcpy(sm->chain_var, var.data(), sizeof(double) * var.size());
  if (sm->deviance_mean_var) {
    // D = -2 loglike over all chains: mean and sample variance from the per-chain moments (equal lengths)
    double gm = 0.0;
    for (int c = 0; c < C; ++c) gm += lm[c];
    gm /= C;
    double m2 = 0.0;
    for (int c = 0; c < C; ++c) m2 += lv[c] * (S - 1) + (double)S * (lm[c] - gm) * (lm[c] - gm);
    sm->deviance_mean_var[0] = -2.0 * gm;
    sm->deviance_mean_var[1] = N > 1 ? 4.0 * m2 / (double)(N - 1) : 0.0;
  }
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
  CUDA_TRY(cudaMalloc(&tmp, std::max<size_t>(bytes, 8)));
  const long long total = n_series * n_keep * w;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((total + 255) / 256, 148 * 16));
  drj_thin_rows<<<grid, 256, 0, s->stream>>>(src, tmp, n_series, n, n_keep, w, thin);
  s->aux_launches++;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, s->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
  cudaFree(tmp);
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
    CUDA_TRY(cudaMalloc(&tmp, sizeof(double) * (size_t)s->P * s->D));
    drj_export_bw<<<(unsigned)((s->P + 255) / 256), 256, 0, s->stream>>>(s->a.logbw, s->P, s->D, tmp);
    s->aux_launches++;
    cudaError_t e = cudaMemcpyAsync(out->bw_final, tmp, sizeof(double) * (size_t)s->P * s->D, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(tmp);
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
