// Ahead-of-time sm_100a instantiation of the small-system kernels for the built-in RHS "harmonic".
#include "ni_rhs_builtin.cuh"
#include "ni_host.h"
using NiRhsThis = NiRhsHarmonic;
NI_DEFINE_SMALL_KSET(ni_kset_harmonic, NiRhsThis)
