// FP64 issue interval for MIXED instruction streams (DMUL / DADD / DFMA alternating on independent chains), with
// register and uniform operands as in the run kernel.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_mix.cu -o /tmp/fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128) kern(double *out, int iters, double a, double b) {
  double x[9], y[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    x[i] = 1.0 + threadIdx.x * 1e-3 + i;
    y[i] = 0.999999 + 1e-9 * (threadIdx.x + i);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 9; i += 3) {
        if (MODE == 0) {  // all DFMA (reference)
          x[i] = fma(x[i], y[i], b);
          x[i + 1] = fma(x[i + 1], y[i + 1], b);
          x[i + 2] = fma(x[i + 2], y[i + 2], b);
        }
        if (MODE == 1) {  // DMUL, DADD, DFMA alternating, register operands
          x[i] = __dmul_rn(x[i], y[i]);
          x[i + 1] = __dadd_rn(x[i + 1], y[i + 1]);
          x[i + 2] = fma(x[i + 2], y[i + 2], b);
        }
        if (MODE == 2) {  // DMUL by uniform, DADD register, DFMA with uniform coefficient
          x[i] = __dmul_rn(x[i], a);
          x[i + 1] = __dadd_rn(x[i + 1], y[i + 1]);
          x[i + 2] = fma(x[i + 2], a, y[i + 2]);
        }
        if (MODE == 3) {  // DMUL, DADD only
          x[i] = __dmul_rn(x[i], y[i]);
          x[i + 1] = __dadd_rn(x[i + 1], y[i + 1]);
          x[i + 2] = __dmul_rn(x[i + 2], a);
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) s += x[i] + y[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int warps_per_smsp) {
  const int blocks = 148 * warps_per_smsp, iters = 4000;
  double *out;
  cudaMalloc(&out, (size_t)blocks * 128 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    kern<MODE><<<blocks, 128>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(out);
  printf("%-52s warps/SMSP %d: %.2f cycles per warp instruction per SMSP\n", name, warps_per_smsp,
         best * 1e-3 * 1.965e9 / (72.0 * iters * warps_per_smsp));
}

int main() {
  for (int w : {4, 5, 6, 8}) {
    run<0>("DFMA only", w);
    run<1>("DMUL / DADD / DFMA alternating (register operands)", w);
    run<2>("DMUL / DADD / DFMA alternating (uniform coefficients)", w);
    run<3>("DMUL / DADD / DMUL", w);
  }
  return 0;
}
