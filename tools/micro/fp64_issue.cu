// Does a warp-wide FP64 instruction block the SMSP's issue port for its two pipe cycles, or can an integer-pipe
// instruction issue in between?  K independent integer LOP3/IADD instructions per DFMA, 8 independent chains per
// thread; nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_issue.cu -o /tmp/fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void __launch_bounds__(256) kern(double *out, unsigned *iout, int iters, double a, double b, unsigned m) {
  double x[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = threadIdx.x * 1e-3 + i;
    u[i] = threadIdx.x * 7u + i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[i] = fma(x[i], a, b);
#pragma unroll
        for (int k = 0; k < K; ++k) u[(i + k) & 7] = (u[(i + k) & 7] ^ m) + (unsigned)(k + r + 1);
      }
    }
  }
  double s = 0;
  unsigned v = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s += x[i];
    v ^= u[i];
  }
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (v == 0x12345u) iout[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

template <int K>
float run(int blocks, int iters) {
  double *out;
  unsigned *iout;
  cudaMalloc(&out, blocks * 256 * 8);
  cudaMalloc(&iout, blocks * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    kern<K><<<blocks, 256>>>(out, iout, iters, 0.999999, 1e-9, 0x5bd1e995u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(out);
  cudaFree(iout);
  return best;
}

int main() {
  const int iters = 4000;
  for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
    const int blocks = 148 * warps_per_smsp / 2;  // 256 threads = 8 warps = 2 per SMSP
    if (blocks < 148) continue;
    const float t0 = run<0>(blocks, iters), t1 = run<1>(blocks, iters), t2 = run<2>(blocks, iters), t4 = run<4>(blocks, iters);
    const double dfma = 64.0 * iters;  // DFMA per thread
    const double cyc0 = t0 * 1e-3 * 1.965e9 / (dfma * (blocks / 148.0) * 2.0);  // cycles per warp-DFMA per SMSP
    printf("warps/SMSP %d: ms K=0 %.3f  K=1 %.3f  K=2 %.3f  K=4 %.3f   ratio %.2f %.2f %.2f   cycles per DFMA (K=0) %.2f\n",
           blocks / 148 * 2, t0, t1, t2, t4, t1 / t0, t2 / t0, t4 / t0, cyc0);
  }
  return 0;
}
