// Ahead-of-time sm_100a instantiation of the small-system kernels for the built-in RHS "riccati".
#include "ni_rhs_builtin.cuh"
#include "ni_host.h"
using NiRhsThis = NiRhsRiccati;
NI_DEFINE_SMALL_KSET(ni_kset_riccati, NiRhsThis)
