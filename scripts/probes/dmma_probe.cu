// FP64 tensor-core (DMMA, mma.sync f64) throughput probe on sm_100a, next to the DFMA figure.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double a, double b, double* out) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x + i; c1[i] = threadIdx.x - i; }
  double fa = a + 1e-9 * threadIdx.x, fb = b + 1e-9 * threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(fa), "d"(fb));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

// m16n8k16: A 8 regs, B 4 regs, C 4 regs
template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double a, double b, double* out) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + i + j;
  double fa[8], fb[4];
#pragma unroll
  for (int j = 0; j < 8; j++) fa[j] = a + 1e-9 * (threadIdx.x + j);
#pragma unroll
  for (int j = 0; j < 4; j++) fb[j] = b + 1e-9 * (threadIdx.x + j);
  for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(fa[0]), "d"(fa[1]), "d"(fa[2]), "d"(fa[3]), "d"(fa[4]), "d"(fa[5]), "d"(fa[6]), "d"(fa[7]),
                     "d"(fb[0]), "d"(fb[1]), "d"(fb[2]), "d"(fb[3]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double a, double b, double* out) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}

template <typename F>
double timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best * 1e-3;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* d; cudaMalloc(&d, 8);
  for (int wps = 4; wps <= 32; wps *= 2) {          // warps per SM
    const int bpsm = (wps * 32 + 255) / 256; const int thr = wps * 32 < 256 ? wps * 32 : 256;
    const int blocks = sms * bpsm;
    const double nw = (double)blocks * thr / 32;
    double t;
    t = timeit([&] { k_dfma<16><<<blocks, thr>>>(0.999999, 1e-9, d); });
    printf("warps/SM %2d  DFMA ilp16      %7.2f TFLOP/s\n", wps, 2.0 * 16 * ITERS * 32 * nw / t / 1e12);
    t = timeit([&] { k_dmma884<4><<<blocks, thr>>>(0.999999, 1e-9, d); });
    printf("warps/SM %2d  DMMA 884  x4    %7.2f TFLOP/s\n", wps, 2.0 * 256 * 4 * ITERS * nw / t / 1e12);
    t = timeit([&] { k_dmma884<16><<<blocks, thr>>>(0.999999, 1e-9, d); });
    printf("warps/SM %2d  DMMA 884  x16   %7.2f TFLOP/s\n", wps, 2.0 * 256 * 16 * ITERS * nw / t / 1e12);
    t = timeit([&] { k_dmma16816<4><<<blocks, thr>>>(0.999999, 1e-9, d); });
    printf("warps/SM %2d  DMMA 16816 x4   %7.2f TFLOP/s\n", wps, 2.0 * 2048 * 4 * (ITERS / 4) * nw / t / 1e12);
    t = timeit([&] { k_dmma16816<8><<<blocks, thr>>>(0.999999, 1e-9, d); });
    printf("warps/SM %2d  DMMA 16816 x8   %7.2f TFLOP/s\n", wps, 2.0 * 2048 * 8 * (ITERS / 4) * nw / t / 1e12);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
