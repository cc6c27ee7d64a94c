// Ahead-of-time sm_100a instantiation of the small-system kernels for the built-in RHS "vanderpol".
#include "ni_rhs_builtin.cuh"
#include "ni_host.h"
using NiRhsThis = NiRhsVanDerPol;
NI_DEFINE_SMALL_KSET(ni_kset_vanderpol, NiRhsThis)
