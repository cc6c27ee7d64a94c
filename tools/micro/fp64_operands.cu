// FP64 issue rate against the number of distinct 64-bit REGISTER source operands (register-file read bandwidth):
// the DFMA peak microbenchmark takes two of its three sources from uniform registers.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_operands.cu -o /tmp/fp64_operands
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) kern(double *out, int iters, double a, double b) {
  double x[8], y[8], z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = threadIdx.x * 1e-3 + i;
    y[i] = 0.999999 + 1e-9 * (threadIdx.x + i);
    z[i] = 1e-9 * (i + 1 + threadIdx.x);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) x[i] = fma(x[i], a, b);                        // 1 register source
        if (MODE == 1) x[i] = fma(x[i], y[i], b);                     // 2
        if (MODE == 2) x[i] = fma(x[i], y[i], z[i]);                  // 3
        if (MODE == 3) x[i] = __dmul_rn(x[i], y[i]);                  // DMUL, 2
        if (MODE == 4) x[i] = __dadd_rn(x[i], z[i]);                  // DADD, 2
        if (MODE == 5) x[i] = fma(x[i], y[(i + r) & 7], z[(i + 3 * r) & 7]);  // 3, rotating (no .reuse pattern)
        if (MODE == 6) x[i] = __dmul_rn(x[i], a);                     // DMUL, 1
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i] + y[i] + z[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name) {
  const int blocks = 148 * 4, iters = 4000;
  double *out;
  cudaMalloc(&out, blocks * 256 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    kern<MODE><<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(out);
  const double per = best * 1e-3 * 1.965e9 / (64.0 * iters * (blocks / 148.0) * 2.0);
  printf("%-44s %.3f ms   %.2f cycles per warp instruction per SMSP\n", name, best, per);
}

int main() {
  run<0>("DFMA  R, c, c   (1 register source)");
  run<1>("DFMA  R, R, c   (2)");
  run<2>("DFMA  R, R, R   (3)");
  run<5>("DFMA  R, R, R   (3, rotating operands)");
  run<3>("DMUL  R, R      (2)");
  run<6>("DMUL  R, c      (1)");
  run<4>("DADD  R, R      (2)");
  return 0;
}
