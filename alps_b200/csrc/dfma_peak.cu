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

// The same FP64 units driven through the tensor-pipe instruction (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4):
// 256 FMAs per warp instruction -- the roofline denominator of the DMMA quadrature variants.
__global__ void __launch_bounds__(256) k_dmma_peak(double a, double b, double* out) {
  double c[PEAK_ILP][2];
#pragma unroll
  for (int i = 0; i < PEAK_ILP; i++) {
    c[i][0] = (double)(threadIdx.x + i);
    c[i][1] = (double)(threadIdx.x - i);
  }
  const double fa = a + 1e-9 * threadIdx.x, fb = b + 1e-9 * threadIdx.x;
  for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
    for (int i = 0; i < PEAK_ILP; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(fa), "d"(fb));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < PEAK_ILP; i++) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

double run_dmma_peak(cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* d_out = nullptr;
  cudaMalloc(&d_out, sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = sms * 2;
  k_dmma_peak<<<blocks, 256, 0, st>>>(0.999999, 1e-9, d_out);
  double best = 0.0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0, st);
    k_dmma_peak<<<blocks, 256, 0, st>>>(0.999999, 1e-9, d_out);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double tf = 2.0 * 256.0 * PEAK_ILP * (double)PEAK_ITERS * 8.0 * blocks / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return best;
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
