// ni_large.cu -- launch plan of the CTA-per-trajectory kernels (ni_large.cuh): picks the
// components-per-thread variant and the CTA size for a state dimension n <= 4096.
#include "ni_large.cuh"
#include "ni_large_host.h"
#include "ni_host.h"
#include "ni_rhs_builtin.cuh"

NI_VISIBLE_INTERNAL bool ni_large_dims(int n, int method, int *lg_out, int *threads, size_t *smem_bytes, bool component) {
  if (n < 1 || n > NI_LARGE_MAX_N) return false;
  int lg = 0;  // components per thread = 1 << lg: the smallest that fits NI_LARGE_MAX_THREADS threads
  while ((n + (1 << lg) - 1) / (1 << lg) > NI_LARGE_MAX_THREADS) ++lg;
  const int cpt = 1 << lg;
  *lg_out = lg;
  *threads = ((n + cpt - 1) / cpt + 31) / 32 * 32;
  *smem_bytes = ni_large_smem_doubles(method, *threads, cpt, component) * sizeof(double);
  return true;
}

NI_VISIBLE_INTERNAL bool ni_large_builtin(int rhs_id, int n, int method, int advanced, NiLargePlan *plan, int *P,
                                          int *Q) {
  const NiLargeKernels *k = rhs_id == NI_RHS_HEAT1D ? ni_large_kernels_heat()
                            : rhs_id == NI_RHS_BURGERS1D ? ni_large_kernels_burgers() : nullptr;
  int lg = 0, threads = 0;
  if (!k || !ni_large_dims(n, method, &lg, &threads, &plan->smem_bytes)) return false;
  plan->cpt = 1 << lg;
  plan->threads = threads;
  plan->k_init = k->init[method][lg];
  plan->k_run[0] = k->run[method][advanced ? 1 : 0][0][lg];
  plan->k_run[1] = k->run[method][advanced ? 1 : 0][1][lg];
  *P = k->P;
  *Q = k->Q;
  return true;
}
