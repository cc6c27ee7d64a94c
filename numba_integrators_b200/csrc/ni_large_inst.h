// Instantiates every (method, advanced, record, components-per-thread) variant for one stencil RHS.
#pragma once
#include "ni_large.cuh"
#include "ni_large_host.h"
#include "ni_host.h"

// 16 components per thread exist only when the CTA is limited to 256 threads (n / 256 = 16 for n = 4096)
#if NI_LARGE_MAX_THREADS <= 256
#define NI_L_16(...) __VA_ARGS__
#else
#define NI_L_16(...) nullptr
#endif
#define NI_L_INIT(RHS, M) \
  {(const void *)&ni_k_large_init<RHS, M, 1>, (const void *)&ni_k_large_init<RHS, M, 2>, \
   (const void *)&ni_k_large_init<RHS, M, 4>, (const void *)&ni_k_large_init<RHS, M, 8>, \
   NI_L_16((const void *)&ni_k_large_init<RHS, M, 16>)}
#define NI_L_RUN(RHS, M, A, R) \
  {(const void *)&ni_k_large_run<RHS, M, A, R, 1>, (const void *)&ni_k_large_run<RHS, M, A, R, 2>, \
   (const void *)&ni_k_large_run<RHS, M, A, R, 4>, (const void *)&ni_k_large_run<RHS, M, A, R, 8>, \
   NI_L_16((const void *)&ni_k_large_run<RHS, M, A, R, 16>)}
#define NI_DEFINE_LARGE_KERNELS(FN, RHS)                                                             \
  NI_VISIBLE_INTERNAL const NiLargeKernels *FN() {                                                   \
    static const NiLargeKernels k = {                                                                \
        RHS::P,                                                                                      \
        RHS::Q,                                                                                      \
        {NI_L_INIT(RHS, NI_RK23), NI_L_INIT(RHS, NI_RK45)},                                          \
        {{{NI_L_RUN(RHS, NI_RK23, 0, false), NI_L_RUN(RHS, NI_RK23, 0, true)},               \
          {NI_L_RUN(RHS, NI_RK23, 1, false), NI_L_RUN(RHS, NI_RK23, 1, true)}},                \
         {{NI_L_RUN(RHS, NI_RK45, 0, false), NI_L_RUN(RHS, NI_RK45, 0, true)},               \
          {NI_L_RUN(RHS, NI_RK45, 1, false), NI_L_RUN(RHS, NI_RK45, 1, true)}}}};              \
    return &k;                                                                                       \
  }
