// Development micro-benchmark #2: inner-loop variants of the 8-accumulator datum pattern at the
// occupancy the sweep kernel really has (128-thread CTAs, 4 per SM = 16 warps/SM).
#include <cstdio>
#include <cuda_runtime.h>
#define TILE 2048

__device__ __forceinline__ void dat(double x, double a, double b, double& acc) { double z = fma(x, a, b); acc = fma(z, z, acc); }

// V1: as in drj_accum_span (unroll 2 groups of 8)
template <int UNR>
__global__ void __launch_bounds__(128, 4) v1(const double* __restrict__ g, double* out, int reps, double a0, double b0) {
  __shared__ __align__(16) double buf[TILE];
  for (int j = threadIdx.x; j < TILE; j += blockDim.x) buf[j] = g[j];
  __syncthreads();
  const double a = a0 + 1e-9 * threadIdx.x, b = b0 - 1e-9 * threadIdx.x;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const double2* t2 = reinterpret_cast<const double2*>(buf);
  for (int r = 0; r < reps; ++r) {
#pragma unroll UNR
    for (int j = 0; j < TILE / 8; ++j) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 v = t2[4 * j + q];
        dat(v.x, a, b, acc[2 * q]);
        dat(v.y, a, b, acc[2 * q + 1]);
      }
    }
  }
  double s = 0;
  for (int k = 0; k < 8; ++k) s += acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// V2: explicit register double buffering: load group j+1 while computing group j
__global__ void __launch_bounds__(128, 4) v2(const double* __restrict__ g, double* out, int reps, double a0, double b0) {
  __shared__ __align__(16) double buf[TILE + 8];
  for (int j = threadIdx.x; j < TILE + 8; j += blockDim.x) buf[j] = g[j % TILE];
  __syncthreads();
  const double a = a0 + 1e-9 * threadIdx.x, b = b0 - 1e-9 * threadIdx.x;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const double2* t2 = reinterpret_cast<const double2*>(buf);
  for (int r = 0; r < reps; ++r) {
    double2 c0 = t2[0], c1 = t2[1], c2 = t2[2], c3 = t2[3];
#pragma unroll 2
    for (int j = 0; j < TILE / 8; ++j) {
      const double2 n0 = t2[4 * j + 4], n1 = t2[4 * j + 5], n2 = t2[4 * j + 6], n3 = t2[4 * j + 7];
      dat(c0.x, a, b, acc[0]); dat(c0.y, a, b, acc[1]); dat(c1.x, a, b, acc[2]); dat(c1.y, a, b, acc[3]);
      dat(c2.x, a, b, acc[4]); dat(c2.y, a, b, acc[5]); dat(c3.x, a, b, acc[6]); dat(c3.y, a, b, acc[7]);
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    }
  }
  double s = 0;
  for (int k = 0; k < 8; ++k) s += acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// V3: 16 accumulators
__global__ void __launch_bounds__(128, 4) v3(const double* __restrict__ g, double* out, int reps, double a0, double b0) {
  __shared__ __align__(16) double buf[TILE];
  for (int j = threadIdx.x; j < TILE; j += blockDim.x) buf[j] = g[j];
  __syncthreads();
  const double a = a0 + 1e-9 * threadIdx.x, b = b0 - 1e-9 * threadIdx.x;
  double acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0;
  const double2* t2 = reinterpret_cast<const double2*>(buf);
  for (int r = 0; r < reps; ++r) {
#pragma unroll 2
    for (int j = 0; j < TILE / 16; ++j) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double2 v = t2[8 * j + q];
        dat(v.x, a, b, acc[2 * q]);
        dat(v.y, a, b, acc[2 * q + 1]);
      }
    }
  }
  double s = 0;
  for (int k = 0; k < 16; ++k) s += acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// V4: two-pass per group: all z first, then all accumulates (max distance between dependent DFMAs)
__global__ void __launch_bounds__(128, 4) v4(const double* __restrict__ g, double* out, int reps, double a0, double b0) {
  __shared__ __align__(16) double buf[TILE];
  for (int j = threadIdx.x; j < TILE; j += blockDim.x) buf[j] = g[j];
  __syncthreads();
  const double a = a0 + 1e-9 * threadIdx.x, b = b0 - 1e-9 * threadIdx.x;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const double2* t2 = reinterpret_cast<const double2*>(buf);
  for (int r = 0; r < reps; ++r) {
#pragma unroll 4
    for (int j = 0; j < TILE / 8; ++j) {
      double z[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double2 v = t2[4 * j + q];
        z[2 * q] = fma(v.x, a, b);
        z[2 * q + 1] = fma(v.y, a, b);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = fma(z[q], z[q], acc[q]);
    }
  }
  double s = 0;
  for (int k = 0; k < 8; ++k) s += acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// V5: no shared memory at all (x synthesised in registers): upper bound of the DFMA pattern itself
__global__ void __launch_bounds__(128, 4) v5(const double* __restrict__ g, double* out, int reps, double a0, double b0) {
  const double a = a0 + 1e-9 * threadIdx.x, b = b0 - 1e-9 * threadIdx.x;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double x[8];
  for (int k = 0; k < 8; ++k) x[k] = g[k] + threadIdx.x;
  for (int r = 0; r < reps; ++r) {
#pragma unroll 4
    for (int j = 0; j < TILE / 8; ++j) {
#pragma unroll
      for (int q = 0; q < 8; ++q) dat(x[q], a, b, acc[q]);
    }
  }
  double s = 0;
  for (int k = 0; k < 8; ++k) s += acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double *g, *out;
  cudaMalloc(&g, TILE * 8); cudaMalloc(&out, sizeof(double) * sms * 64 * 1024);
  cudaMemset(g, 0, TILE * 8);
  const int reps = 400;
  for (int cps : {1, 2, 4}) {  // CTAs (of 128 threads) per SM
    const int blocks = sms * cps, th = 128;
    const double fl = 4.0 * TILE * reps * (double)blocks * th;
#define T(name, call) { float ms = time_ms([&] { call; }); printf("%-28s CTAs/SM %d: %.2f TFLOP/s\n", name, cps, fl / ms / 1e9); }
    T("v1 unroll1", (v1<1><<<blocks, th>>>(g, out, reps, 0.5, -1.5)))
    T("v1 unroll2 (current)", (v1<2><<<blocks, th>>>(g, out, reps, 0.5, -1.5)))
    T("v1 unroll4", (v1<4><<<blocks, th>>>(g, out, reps, 0.5, -1.5)))
    T("v1 unroll8", (v1<8><<<blocks, th>>>(g, out, reps, 0.5, -1.5)))
    T("v2 reg double buffer", (v2<<<blocks, th>>>(g, out, reps, 0.5, -1.5)))
    T("v3 16 accumulators", (v3<<<blocks, th>>>(g, out, reps, 0.5, -1.5)))
    T("v4 z-then-acc unroll4", (v4<<<blocks, th>>>(g, out, reps, 0.5, -1.5)))
    T("v5 no LDS (upper bound)", (v5<<<blocks, th>>>(g, out, reps, 0.5, -1.5)))
  }
  return 0;
}
