// FP64 pipe utilisation against instruction-level parallelism per warp and resident warps per SMSP: ILP independent
// DFMA chains per thread (two register sources each), W warps per scheduler.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_ilp.cu -o /tmp/fp64_ilp
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void kern(double *out, int iters, double b) {
  double x[ILP], y[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) {
    x[i] = threadIdx.x * 1e-3 + i;
    y[i] = 0.999999 + 1e-9 * (threadIdx.x + i);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 24 / ILP; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], y[i], b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
double run(int warps_per_smsp) {
  const int threads = 128, blocks = 148 * warps_per_smsp, iters = 20000;  // 128 threads = 1 warp per SMSP
  double *out;
  cudaMalloc(&out, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    kern<ILP><<<blocks, threads>>>(out, iters, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(out);
  return best * 1e-3 * 1.965e9 / (24.0 * iters * warps_per_smsp);  // cycles per warp-DFMA per SMSP
}

int main() {
  printf("cycles per warp DFMA per SMSP (2.0 = pipe saturated)\n warps/SMSP   ILP1   ILP2   ILP3   ILP4   ILP6\n");
  for (int w : {1, 2, 3, 4, 5, 6, 8}) {
    printf(" %6d     %6.2f %6.2f %6.2f %6.2f %6.2f\n", w, run<1>(w), run<2>(w), run<3>(w), run<4>(w), run<6>(w));
  }
  return 0;
}
