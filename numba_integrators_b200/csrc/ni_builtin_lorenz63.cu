// Ahead-of-time sm_100a instantiation of the small-system kernels for the built-in RHS "lorenz63".
#include "ni_rhs_builtin.cuh"
#include "ni_host.h"
using NiRhsThis = NiRhsLorenz63;
NI_DEFINE_SMALL_KSET(ni_kset_lorenz63, NiRhsThis)
