// ni_large.cu -- CTA-per-trajectory kernels for large state dimension (placeholder until the
// method-of-lines kernels land).
#include "ni_large_host.h"
#include "ni_host.h"

NI_VISIBLE_INTERNAL bool ni_large_builtin(int, int, int, int, NiLargePlan *, int *, int *) { return false; }
NI_VISIBLE_INTERNAL bool ni_large_prepare(NiLargePlan *, NiArgs *, int, int, int) { return true; }
NI_VISIBLE_INTERNAL int ni_large_set_tolerances(NiLargePlan *, NiArgs *, const double *, const double *, int, cudaStream_t) { return 0; }
