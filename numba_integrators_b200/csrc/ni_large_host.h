// ni_large_host.h -- host-side interface of the CTA-per-trajectory kernels (ni_large_*.cu).
#pragma once
#include <cuda_runtime.h>

#include "ni_core.cuh"

struct NiLargePlan {
  const void *k_init = nullptr;
  const void *k_run[2] = {nullptr, nullptr};  // [record]
  int threads = 0, cpt = 0;
  size_t smem_bytes = 0;
  double *tol_d = nullptr;  // rtol[n] followed by atol[n] (owned by the ensemble)
};

struct NiLargeKernels {  // every instantiation for one stencil RHS
  int P, Q;
  const void *init[2][5];        // [method][log2 cpt]
  const void *run[2][2][2][5];   // [method][advanced][record][log2 cpt]
};
const NiLargeKernels *ni_large_kernels_heat();
const NiLargeKernels *ni_large_kernels_burgers();

// CTA shape for a state dimension n <= 4096: components per thread = 1 << lg, threads, dynamic shared memory.
bool ni_large_dims(int n, int method, int *lg, int *threads, size_t *smem_bytes, bool component = false);
// Returns false when (rhs_id, n) has no CTA-per-trajectory kernel.
bool ni_large_builtin(int rhs_id, int n, int method, int advanced, NiLargePlan *plan, int *P, int *Q);
