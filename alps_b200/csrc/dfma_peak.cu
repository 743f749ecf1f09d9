// alps_b200: FP64 FMA micro-benchmark -- the roofline denominator for the quadrature kernel
// (MEASURED_PEAKS.json has no FP64 entry).  Not part of the hot path.
#include "common.cuh"
#include "kernels.h"

namespace alps {

constexpr int PEAK_ILP = 16;
constexpr int PEAK_ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma_peak(double a, double b, double* out) {
  double acc[PEAK_ILP];
#pragma unroll
  for (int i = 0; i < PEAK_ILP; i++) acc[i] = (double)(threadIdx.x + i);
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int i = 0; i < PEAK_ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < PEAK_ILP; i++) s += acc[i];
  if (s == 123.456) out[0] = s;   // never true; keeps the loop alive
}

// Same FMA count, but every DFMA reads three fresh 64-bit operands (no operand-reuse-cache hit between
// consecutive instructions): measures the register-read limit that a register-tiled FP64 kernel sees.
__global__ void __launch_bounds__(256) k_dfma_peak_noreuse(double a, double b, double* out) {
  double acc[PEAK_ILP], x[8], y[8];
#pragma unroll
  for (int i = 0; i < PEAK_ILP; i++) acc[i] = (double)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    x[i] = a + 1e-12 * (threadIdx.x + i);
    y[i] = b + 1e-13 * (threadIdx.x + 3 * i);
  }
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int i = 0; i < PEAK_ILP; i++) acc[i] = fma(x[i & 7], y[(3 * i + 1) & 7], acc[i]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < PEAK_ILP; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}

double run_dfma_peak_noreuse(cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* d_out = nullptr;
  cudaMalloc(&d_out, sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8 * 4;
  k_dfma_peak_noreuse<<<blocks, 256, 0, st>>>(0.999999, 1e-9, d_out);
  double best = 0.0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0, st);
    k_dfma_peak_noreuse<<<blocks, 256, 0, st>>>(0.999999, 1e-9, d_out);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * PEAK_ILP * (double)PEAK_ITERS * 256.0 * blocks / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return best;
}

double run_dfma_peak(cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* d_out = nullptr;
  cudaMalloc(&d_out, sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 8 * 4;
  k_dfma_peak<<<blocks, 256, 0, st>>>(0.999999, 1e-9, d_out);   // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0, st);
    k_dfma_peak<<<blocks, 256, 0, st>>>(0.999999, 1e-9, d_out);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * PEAK_ILP * (double)PEAK_ITERS * 256.0 * blocks;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return best;
}

}  // namespace alps
