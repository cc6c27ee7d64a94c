// ni_rhs_builtin.cuh -- built-in right-hand sides for the small-system kernels.
//
// Each functor: N (state dim), P (parameters per trajectory), Q (auxiliary outputs) and
//   eval(t, y[N], p[P], dy[N], aux[Q]).
// The expression order (and the absence of FMA contraction: __dmul_rn/__dadd_rn) is the
// same as in oracle/ni_oracle.c:rhs_eval and in the numba functions oracle/gen_golden.py
// hands to the reference, so that all three produce identical bits.
#pragma once
#ifndef NI_RHS_BUILTIN_CUH
#define NI_RHS_BUILTIN_CUH
#include "ni_core.cuh"

#define NI_RHS_HARMONIC 0
#define NI_RHS_LORENZ63 1
#define NI_RHS_VANDERPOL 2
#define NI_RHS_HEAT1D 3
#define NI_RHS_BURGERS1D 4
#define NI_RHS_EXPONENTIAL 5
#define NI_RHS_RICCATI 6
#define NI_RHS_README_ADV 7
#define NI_RHS_BLOWUP 8
#define NI_RHS_COUNT 9

// A = NiA<FAST>: strict (unfused, the reference's arithmetic) or fast (contracted multiply-adds)

// y0' = y1, y1' = -y0   (reference.py:46-49 "sine", README.md:54-57)
struct NiRhsHarmonic {
  static constexpr int N = 2, P = 0, Q = 0;
  template <bool FAST>
  NI_DEV static void eval(double, const double *y, const double *, double *dy, double *) {
    dy[0] = y[1];
    dy[1] = -y[0];
  }
};

// Lorenz-63, p = (sigma, rho, beta), aux = x^2 + y^2 + z^2
struct NiRhsLorenz63 {
  static constexpr int N = 3, P = 3, Q = 1;
  template <bool FAST>
  NI_DEV static void eval(double, const double *y, const double *p, double *dy, double *aux) {
    using A = NiA<FAST>;
    const double x = y[0], yy = y[1], z = y[2];
    dy[0] = A::mul(p[0], A::sub(yy, x));
    dy[1] = A::msub(x, A::sub(p[1], z), yy);
    dy[2] = A::msub(x, yy, A::mul(p[2], z));
    aux[0] = A::madd(z, z, A::madd(yy, yy, A::mul(x, x)));  // strict: (x*x + y*y) + z*z
  }
};

// Van der Pol, p = (mu)
struct NiRhsVanDerPol {
  static constexpr int N = 2, P = 1, Q = 0;
  template <bool FAST>
  NI_DEV static void eval(double, const double *y, const double *p, double *dy, double *) {
    using A = NiA<FAST>;
    dy[0] = y[1];
    dy[1] = A::msub(A::mul(p[0], A::nmadd(y[0], y[0], 1.0)), y[1], y[0]);  // (mu (1 - y0^2)) y1 - y0
  }
};

// y' = y   (reference.py:54-57), any small N
template <int N_>
struct NiRhsExponential {
  static constexpr int N = N_, P = 0, Q = 0;
  template <bool FAST>
  NI_DEV static void eval(double, const double *y, const double *, double *dy, double *) {
#pragma unroll
    for (int i = 0; i < N; ++i) dy[i] = y[i];
  }
};

// y' = 2y/x - x^2 y^2   (reference.py:38-41)
struct NiRhsRiccati {
  static constexpr int N = 1, P = 0, Q = 0;
  template <bool FAST>
  NI_DEV static void eval(double t, const double *y, const double *, double *dy, double *) {
    using A = NiA<FAST>;
    dy[0] = A::nmadd(A::mul(t, t), A::mul(y[0], y[0]), A::mul(2.0, y[0]) / t);
  }
};

// README.md:80-85 Advanced example: p = (a, b0, b1); aux = a*y1; dy = (aux + b0, -y0 + b1)
struct NiRhsReadmeAdvanced {
  static constexpr int N = 2, P = 3, Q = 1;
  template <bool FAST>
  NI_DEV static void eval(double, const double *y, const double *p, double *dy, double *aux) {
    using A = NiA<FAST>;
    const double a = A::mul(p[0], y[1]);
    dy[0] = A::add(a, p[1]);
    dy[1] = A::add(-y[0], p[2]);
    aux[0] = a;
  }
};

// y' = y^2 (finite-time blow-up: exercises the step-too-small exit, _basic.py:179-180)
struct NiRhsBlowup {
  static constexpr int N = 1, P = 0, Q = 0;
  template <bool FAST>
  NI_DEV static void eval(double, const double *y, const double *, double *dy, double *) {
    dy[0] = NiA<FAST>::mul(y[0], y[0]);
  }
};
#endif  // NI_RHS_BUILTIN_CUH
