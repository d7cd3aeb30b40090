// DFMA / FFMA vector peak of the box (SURVEY 8d asks for measured fp64 / fp32 peaks; MEASURED_PEAKS.json has only HBM and
// bf16).  8 independent FMA chains per thread, 1024 threads x 2 CTAs per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void __launch_bounds__(1024) fma_kernel(T* out, int iters, T a, T b) {
  T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
      x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
template <typename T>
double run(const char* name, int iters) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int blocks = sms * 2;
  T* out; cudaMalloc(&out, sizeof(T) * blocks * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  fma_kernel<T><<<blocks, 1024>>>(out, iters, (T)1.0000001, (T)1e-9);
  cudaEventRecord(a);
  for (int r = 0; r < 5; r++) fma_kernel<T><<<blocks, 1024>>>(out, iters, (T)1.0000001, (T)1e-9);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double flops = 2.0 * 64 * iters * 1024.0 * blocks * 5;
  double tf = flops / (ms * 1e-3) / 1e12;
  printf("{\"pipe\": \"%s\", \"tflops\": %.2f, \"ms\": %.3f, \"sms\": %d}\n", name, tf, ms, sms);
  cudaFree(out);
  return tf;
}
int main() { run<double>("fp64_dfma", 20000); run<float>("fp32_ffma", 40000); return 0; }
