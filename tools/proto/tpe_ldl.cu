// Prototype measurement for DESIGN.md section 6: thread-per-env 18x18 LDL^T factorisation + solve, the solver core of
// one Newton iteration, with lanes = envs.  Matrices live in a global SoA array [entry][env] (coalesced), each
// thread factorises its own matrix.  Variants: ROLLED (dynamic indices, local-memory array) and UNROLLED (compile-time
// indices, ptxas keeps what it can in registers).   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tpe_ldl.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int NV = 18, NTRI = NV * (NV + 1) / 2;

template <bool UNROLL>
__global__ void __launch_bounds__(128) ldl_kernel(const double* __restrict__ Hin, const double* __restrict__ gin,
                                                  double* __restrict__ xout, int n) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double H[NTRI], y[NV], dinv[NV];
  if (UNROLL) {
#pragma unroll
    for (int t = 0; t < NTRI; t++) H[t] = Hin[(size_t)t * n + e];
#pragma unroll
    for (int i = 0; i < NV; i++) y[i] = gin[(size_t)i * n + e];
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double inv = 1.0 / fmax(H[k * (k + 1) / 2 + k], 1e-15);
      dinv[k] = inv;
#pragma unroll
      for (int i = k + 1; i < NV; i++) {
        double t = H[i * (i + 1) / 2 + k], l = t * inv;
#pragma unroll
        for (int j = k + 1; j <= i; j++) H[i * (i + 1) / 2 + j] -= l * H[j * (j + 1) / 2 + k];
        y[i] -= l * y[k];
      }
    }
#pragma unroll
    for (int i = 0; i < NV; i++) y[i] *= dinv[i];
#pragma unroll
    for (int k = NV - 1; k > 0; k--)
#pragma unroll
      for (int i = 0; i < k; i++) y[i] -= H[k * (k + 1) / 2 + i] * dinv[i] * y[k];
#pragma unroll
    for (int i = 0; i < NV; i++) xout[(size_t)i * n + e] = y[i];
  } else {
    for (int t = 0; t < NTRI; t++) H[t] = Hin[(size_t)t * n + e];
    for (int i = 0; i < NV; i++) y[i] = gin[(size_t)i * n + e];
#pragma unroll 1
    for (int k = 0; k < NV; k++) {
      double inv = 1.0 / fmax(H[k * (k + 1) / 2 + k], 1e-15);
      dinv[k] = inv;
#pragma unroll 1
      for (int i = k + 1; i < NV; i++) {
        double t = H[i * (i + 1) / 2 + k], l = t * inv;
#pragma unroll 1
        for (int j = k + 1; j <= i; j++) H[i * (i + 1) / 2 + j] -= l * H[j * (j + 1) / 2 + k];
        y[i] -= l * y[k];
      }
    }
    for (int i = 0; i < NV; i++) y[i] *= dinv[i];
#pragma unroll 1
    for (int k = NV - 1; k > 0; k--)
#pragma unroll 1
      for (int i = 0; i < k; i++) y[i] -= H[k * (k + 1) / 2 + i] * dinv[i] * y[k];
    for (int i = 0; i < NV; i++) xout[(size_t)i * n + e] = y[i];
  }
}

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 1 << 20;
  double *H, *g, *x;
  cudaMalloc(&H, (size_t)NTRI * n * 8); cudaMalloc(&g, (size_t)NV * n * 8); cudaMalloc(&x, (size_t)NV * n * 8);
  // SPD matrices: diag 10 + i, off-diagonals 0.01
  double* h = (double*)malloc((size_t)NTRI * n * 8);
  for (int i = 0, t = 0; i < NV; i++) for (int j = 0; j <= i; j++, t++) for (int e = 0; e < n; e++) h[(size_t)t * n + e] = i == j ? 10.0 + i + 1e-3 * (e % 7) : 0.01 * ((i + j) % 3);
  cudaMemcpy(H, h, (size_t)NTRI * n * 8, cudaMemcpyHostToDevice);
  cudaMemset(g, 0, (size_t)NV * n * 8);
  for (int variant = 0; variant < 2; variant++) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int reps = 20;
    for (int r = 0; r < reps + 3; r++) {
      if (r == 3) cudaEventRecord(a);
      if (variant == 0) ldl_kernel<false><<<(n + 127) / 128, 128>>>(H, g, x, n);
      else ldl_kernel<true><<<(n + 127) / 128, 128>>>(H, g, x, n);
    }
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaError_t err = cudaGetLastError();
    printf("{\"variant\": \"%s\", \"n\": %d, \"ms_per_launch\": %.4f, \"factor_solves_per_s\": %.4g, \"GBps_in\": %.1f, \"err\": \"%s\"}\n",
           variant ? "unrolled" : "rolled", n, ms / reps, (double)n * reps / (ms * 1e-3), (double)(NTRI + 2 * NV) * 8 * n * reps / (ms * 1e-3) / 1e9, cudaGetErrorString(err));
  }
  return 0;
}
