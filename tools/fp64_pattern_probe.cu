// Micro-benchmark (development tool, not part of the library): FP64 pipe utilisation of the
// per-datum normal-likelihood pattern  z = fma(x,a,b); acc_k = fma(z,z,acc_k)  read from shared
// memory with broadcast LDS, as a function of accumulators per thread and warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pattern_probe fp64_pattern_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#define TILE 2048

template <int ACC>
__global__ void pattern(const double* __restrict__ g, double* out, int reps, double a0, double b0) {
  __shared__ __align__(16) double buf[TILE];
  for (int j = threadIdx.x; j < TILE; j += blockDim.x) buf[j] = g[j];
  __syncthreads();
  const double a = a0 + 1e-9 * threadIdx.x, b = b0 - 1e-9 * threadIdx.x;
  double acc[ACC];
#pragma unroll
  for (int k = 0; k < ACC; ++k) acc[k] = 0.0;
  const double2* t2 = reinterpret_cast<const double2*>(buf);
  for (int r = 0; r < reps; ++r) {
#pragma unroll 4
    for (int j = 0; j < TILE / 2; j += ACC / 2 > 0 ? ACC / 2 : 1) {
      if (ACC == 1) {
        double2 v = t2[j];
        double z = fma(v.x, a, b); acc[0] = fma(z, z, acc[0]);
        z = fma(v.y, a, b); acc[0] = fma(z, z, acc[0]);
      } else {
#pragma unroll
        for (int k = 0; k < ACC / 2; ++k) {
          double2 v = t2[j + k];
          double z0 = fma(v.x, a, b), z1 = fma(v.y, a, b);
          acc[2 * k] = fma(z0, z0, acc[2 * k]);
          acc[2 * k + 1] = fma(z1, z1, acc[2 * k + 1]);
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ACC; ++k) s += acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void chains8(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// same but all-register operands (no immediates / uniform registers)
__global__ void chains8r(double* out, int iters, double a0, double b0) {
  double a = a0 + 1e-12 * threadIdx.x, b = b0 + 1e-12 * threadIdx.x;
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
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
  cudaMalloc(&g, TILE * 8); cudaMalloc(&out, sizeof(double) * sms * 32 * 1024);
  cudaMemset(g, 0, TILE * 8);
  printf("device %s, %d SMs\n", p.name, sms);
  for (int wps : {4, 8, 16, 32, 64}) {           // warps per SM
    const int threads = 256, blocks = sms * (wps * 32 / threads > 0 ? wps * 32 / threads : 1);
    const int th = wps * 32 < threads ? wps * 32 : threads;
    const int iters = 2048;
    float ms = time_ms([&] { chains8<<<blocks, th>>>(out, iters, 0.999999, 1e-7); });
    float ms2 = time_ms([&] { chains8r<<<blocks, th>>>(out, iters, 0.999999, 1e-7); });
    double fl = 2.0 * 64 * iters * (double)blocks * th;
    printf("chains8   warps/SM %2d: %.2f TFLOP/s (imm/uniform operands)  %.2f TFLOP/s (register operands)\n", wps,
           fl / ms / 1e9, fl / ms2 / 1e9);
  }
  const int reps = 200;
#define RUN(ACC)                                                                                   \
  for (int wps : {8, 16, 24, 32, 64}) {                                                            \
    const int threads = 256;                                                                       \
    const int blocks = sms * (wps * 32 / threads);                                                 \
    float ms = time_ms([&] { pattern<ACC><<<blocks, threads>>>(g, out, reps, 0.5, -1.5); });       \
    double fl = 4.0 * TILE * reps * (double)blocks * threads;                                      \
    printf("pattern acc=%d warps/SM %2d: %.2f TFLOP/s\n", ACC, wps, fl / ms / 1e9);                 \
  }
  RUN(1) RUN(2) RUN(4) RUN(8)
  return 0;
}
