// ni_large_host.h -- host-side interface of the CTA-per-trajectory kernels (ni_large.cu).
#pragma once
#include <cuda_runtime.h>

#include "ni_core.cuh"

struct NiLargePlan {
  const void *k_init = nullptr;
  const void *k_run[2] = {nullptr, nullptr};  // [record]
  int threads = 0;
  size_t smem_bytes = 0;
  int rhs_id = 0;
};

// Returns false when (rhs_id, n) has no CTA-per-trajectory kernel.
bool ni_large_builtin(int rhs_id, int n, int method, int advanced, NiLargePlan *plan, int *P, int *Q);
// Allocates what the plan needs on the current device (tolerance vectors, ...) and opts in to large shared memory.
bool ni_large_prepare(NiLargePlan *plan, NiArgs *a, int n, int device, int sms);
int ni_large_set_tolerances(NiLargePlan *plan, NiArgs *a, const double *rtol, const double *atol, int n, cudaStream_t s);
