// Ahead-of-time sm_100a instantiation of the small-system kernels for the built-in RHS "exponential1".
#include "ni_rhs_builtin.cuh"
#include "ni_host.h"
using NiRhsThis = NiRhsExponential<1>;
NI_DEFINE_SMALL_KSET(ni_kset_exponential1, NiRhsThis)
