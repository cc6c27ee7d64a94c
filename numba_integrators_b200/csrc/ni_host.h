// ni_host.h -- internal glue between the C-ABI implementation (ni_b200.cu) and the
// translation units that instantiate the built-in kernels.
#pragma once
#include "ni_core.cuh"

#define NI_KIND_SMALL 0  // thread-per-trajectory, component-major state
#define NI_KIND_LARGE 1  // CTA-per-trajectory, trajectory-major state

struct NiKernelSet {
  int kind;
  int n, P, Q;                  // n < 0: any dimension (large kernels take n at run time)
  const void *init[2];          // [method]
  const void *run[2][3][2][2];  // [method][semantics: basic, advanced, scipy][record][fast arithmetic]
};

#define NI_VISIBLE_INTERNAL __attribute__((visibility("hidden")))

#define NI_KR(RHS, M, A, R) {(const void *)&ni_k_run<RHS, M, A, R, false>, (const void *)&ni_k_run<RHS, M, A, R, true>}
// Fills a kernel set with every (method, advanced, record, arithmetic) instantiation for an RHS functor.
#define NI_DEFINE_SMALL_KSET(FN, RHS)                                                         \
  NI_VISIBLE_INTERNAL const NiKernelSet *FN() {                                               \
    static const NiKernelSet ks = {                                                           \
        NI_KIND_SMALL,                                                                        \
        RHS::N,                                                                               \
        RHS::P,                                                                               \
        RHS::Q,                                                                               \
        {(const void *)&ni_k_init<RHS, NI_RK23>, (const void *)&ni_k_init<RHS, NI_RK45>},     \
        {{{NI_KR(RHS, NI_RK23, 0, false), NI_KR(RHS, NI_RK23, 0, true)},                      \
          {NI_KR(RHS, NI_RK23, 1, false), NI_KR(RHS, NI_RK23, 1, true)},                      \
          {NI_KR(RHS, NI_RK23, 2, false), NI_KR(RHS, NI_RK23, 2, true)}},                     \
         {{NI_KR(RHS, NI_RK45, 0, false), NI_KR(RHS, NI_RK45, 0, true)},                      \
          {NI_KR(RHS, NI_RK45, 1, false), NI_KR(RHS, NI_RK45, 1, true)},                      \
          {NI_KR(RHS, NI_RK45, 2, false), NI_KR(RHS, NI_RK45, 2, true)}}}};                   \
    return &ks;                                                                               \
  }

const NiKernelSet *ni_kset_harmonic();
const NiKernelSet *ni_kset_lorenz63();
const NiKernelSet *ni_kset_vanderpol();
const NiKernelSet *ni_kset_exponential1();
const NiKernelSet *ni_kset_riccati();
const NiKernelSet *ni_kset_readme_adv();
const NiKernelSet *ni_kset_blowup();
