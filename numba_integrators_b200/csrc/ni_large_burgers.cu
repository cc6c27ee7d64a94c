// Ahead-of-time sm_100a instantiation of the CTA-per-trajectory kernels: viscous Burgers equation.
#include "ni_large_inst.h"
NI_DEFINE_LARGE_KERNELS(ni_large_kernels_burgers, NiStencilBurgers)
