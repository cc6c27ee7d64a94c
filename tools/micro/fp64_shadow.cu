// What does a non-FP64 instruction cost when it is issued between FP64 instructions (the run kernels' mix)?
//
// Per round: 8 independent FP64 instructions (one per chain) and NI integer-side instructions of one class spread
// evenly between them, all independent of the FP64 chains.  If the FP64 pipe (2 cycles per warp instruction) were the
// only limit, a round would take 16 cycles per warp per scheduler up to NI = 8 (one free issue slot per FP64
// instruction), i.e. max(2 D, D + I); if the issue port were blocked for both FP64 cycles it would take 16 + NI
// (2 D + I).  The table printed by main() gives cycles per round for every (FP64 form, integer class, NI).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/fp64_shadow.cu -o /tmp/fp64_shadow
//   cuobjdump -sass /tmp/fp64_shadow | python tools/micro/shadow_mix.py      # static mix of every kernel's loop
#include <cstdio>
#include <cuda_runtime.h>

// FOP: 0 DFMA R,R,UR   1 DMUL R,UR   2 DADD R,R   3 DFMA R,R,R
// IOP: 0 LOP3 R,R,R    1 IADD3 R,R,R  2 SEL R,R,P (predicate loop-invariant)   3 ISETP R,R + SEL (2 instructions)
//      4 IMAD R,R,R (FMA pipe)   5 SHF + LOP3 (1-2 register sources, 2 instructions)   6 VIMNMX R,R
// (tools/micro/shadow_mix.py prints the static mix of every loop: check it before reading the table)
template <int FOP, int IOP, int NI>
__global__ void __launch_bounds__(128) kern(double *out, unsigned *iout, int iters, double a, double b, unsigned m) {
  double x[8], y[8], z[8];
  unsigned u[8], v[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    x[i] = 1.0 + threadIdx.x * 1e-3 + i;
    y[i] = 0.999999 + 1e-9 * (threadIdx.x + i);
    z[i] = 1e-9 * (i + 1 + threadIdx.x);
    u[i] = threadIdx.x * 7u + i;
    v[i] = threadIdx.x * 13u + 3u * i + m;
    w[i] = threadIdx.x * 29u + 5u * i + (m >> 3);
  }
  const unsigned pi = (threadIdx.x * m) & 4u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (FOP == 0) x[i] = fma(x[i], y[i], b);
        if (FOP == 1) x[i] = __dmul_rn(x[i], a);
        if (FOP == 2) x[i] = __dadd_rn(x[i], z[i]);
        if (FOP == 3) x[i] = fma(x[i], y[i], z[i]);
        constexpr int per = IOP == 3 || IOP == 5 ? 2 : 1;  // instructions per integer "operation"
#pragma unroll
        for (int k = NI * i / 8 / per; k < NI * (i + 1) / 8 / per; ++k) {
          // operands rotate so that consecutive operations on one u[] never see the same (v, w) pair, and every
          // class but ISETP+SEL / IMAD is inline PTX: neither NVVM nor ptxas can fold or hoist the chain
          const int j = (k + r) & 7, j2 = (k + 3 * r + 3) & 7, j3 = (k + 5 * r + 1) & 7;
          if (IOP == 0) asm("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(u[j]) : "r"(v[j2]), "r"(w[j3]));
          if (IOP == 1) asm("{ .reg .u32 s; add.u32 s, %0, %1; add.u32 %0, s, %2; }" : "+r"(u[j]) : "r"(v[j2]), "r"(w[j3]));
          if (IOP == 2) asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.b32 %0, %1, %2, q; }" : "+r"(u[j]) : "r"(v[j2]), "r"(u[(j + 1) & 7]), "r"(pi));
          if (IOP == 3) u[j] = u[j] < v[j2] ? w[j3] : u[j];
          if (IOP == 4) u[j] = u[j] * v[j2] + w[j3];
          if (IOP == 5) asm("{ .reg .u32 s; shr.u32 s, %0, 7; xor.b32 %0, %0, s; }" : "+r"(u[j]));
          if (IOP == 6) asm("max.u32 %0, %0, %1;" : "+r"(u[j]) : "r"(v[j2]));
        }
      }
    }
  }
  double s = 0;
  unsigned q = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s += x[i] + y[i] + z[i];
    q ^= u[i] + v[i] + w[i];
  }
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (q == 0x12345u) iout[blockIdx.x * blockDim.x + threadIdx.x] = q;
}

// The commit pattern of the run kernel: integer instructions that DEPEND on FP64 results and feed FP64 instructions:
// hi word -> ISETP -> two FSEL (a 64-bit select) -> next DFMA of the same chain.  CH chains per thread.
template <int CH, bool WITH_SELECT>
__global__ void __launch_bounds__(128) kern_dep(double *out, int iters, double a, double b) {
  double x[CH], y[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    x[i] = 1.0 + threadIdx.x * 1e-3 + i;
    y[i] = 0.999999 + 1e-9 * (threadIdx.x + i);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const double xn = fma(x[i], y[i], b);
        if (WITH_SELECT) {
          const bool keep = __double2hiint(xn) < 0x40590000;  // < 100.0: integer-pipe compare on the high word
          x[i] = keep ? xn : y[i];
        } else {
          x[i] = xn;
        }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += x[i] + y[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double *g_out;
static unsigned *g_iout;
static const int WARPS = 5;  // resident warps per scheduler, as in the run kernel (5 CTAs x 128 threads per SM)

template <class F>
static float time_best(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

static double g_mhz = 1965.0;
template <int FOP, int IOP, int NI>
static double cyc() {
  const int iters = 2000, blocks = 148 * WARPS;
  const float ms = time_best([&] { kern<FOP, IOP, NI><<<blocks, 128>>>(g_out, g_iout, iters, 0.999999, 1e-9, 0x5bd1e995u); });
  // cycles per round (8 FP64 + NI others) per warp per scheduler
  return ms * 1e-3 * g_mhz * 1e6 / (4.0 * iters * WARPS);
}
template <int FOP, int IOP>
static void row(const char *fname, const char *iname) {
  printf("%-14s + %-34s NI=0 %6.2f  2 %6.2f  4 %6.2f  8 %6.2f  12 %6.2f  16 %6.2f\n", fname, iname, cyc<FOP, IOP, 0>(),
         cyc<FOP, IOP, 2>(), cyc<FOP, IOP, 4>(), cyc<FOP, IOP, 8>(), cyc<FOP, IOP, 12>(), cyc<FOP, IOP, 16>());
}
template <int FOP>
static void block(const char *fname) {
  row<FOP, 0>(fname, "LOP3 R,R,R");
  row<FOP, 1>(fname, "IADD3 R,R,R");
  row<FOP, 2>(fname, "SEL R,R,P");
  row<FOP, 3>(fname, "ISETP R,R + SEL R,R,P");
  row<FOP, 4>(fname, "IMAD R,R,R");
  row<FOP, 5>(fname, "SHF R,imm + LOP3 R,R");
  row<FOP, 6>(fname, "VIMNMX R,R");
}

template <int CH, bool SEL>
static double cyc_dep() {
  const int iters = 2000, blocks = 148 * WARPS;
  const float ms = time_best([&] { kern_dep<CH, SEL><<<blocks, 128>>>(g_out, iters, 0.999999, 1e-9); });
  return ms * 1e-3 * g_mhz * 1e6 / (4.0 * iters * WARPS * CH);  // cycles per DFMA (+ its select) per warp per scheduler
}

int main() {
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  if (khz > 0) g_mhz = khz / 1000.0;
  cudaMalloc(&g_out, (size_t)148 * WARPS * 128 * 8);
  cudaMalloc(&g_iout, (size_t)148 * WARPS * 128 * 4);
  printf("# cycles per round of 8 FP64 instructions + NI other instructions, per warp per scheduler, %d warps per scheduler,\n"
         "# SM clock %.0f MHz (pipe-bound: 16; '2 D + I' would be 16 + NI; 'max(2 D, D + I)' 16 up to NI = 8, then 8 + NI)\n",
         WARPS, g_mhz);
  block<0>("DFMA R,R,UR");
  block<1>("DMUL R,UR");
  block<2>("DADD R,R");
  block<3>("DFMA R,R,R");
  printf("# dependent select: DFMA -> hi word -> ISETP -> 2 x FSEL -> next DFMA of the chain; cycles per DFMA\n");
  printf("chains 3: plain %5.2f  with select %5.2f\n", cyc_dep<3, false>(), cyc_dep<3, true>());
  printf("chains 6: plain %5.2f  with select %5.2f\n", cyc_dep<6, false>(), cyc_dep<6, true>());
  printf("chains 8: plain %5.2f  with select %5.2f\n", cyc_dep<8, false>(), cyc_dep<8, true>());
  return 0;
}
