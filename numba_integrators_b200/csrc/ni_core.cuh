// ni_core.cuh -- device side of the B200 ODE-ensemble engine (small systems).
//
// One thread integrates one trajectory; the RK stage vectors k1..k7 live in
// registers (every index below is a compile-time constant), the ensemble state
// is structure-of-arrays in HBM (component-major, stride N) so that loads and
// stores of a warp are coalesced, and a persistent grid pulls trajectories from
// an atomic work queue so lanes that finish early are refilled instead of
// idling (heterogeneous step counts, BASELINE config C4).
//
// This header is compiled two ways: ahead of time by nvcc for the built-in
// right-hand sides (ni_b200.cu) and at load time by NVRTC with a user-supplied
// RHS snippet.  It therefore includes no host headers.
//
// Algorithm (reference = /root/reference/src/numba_integrators):
//   _basic.py:114-180  _step           -> ni_small_run (attempt loop)
//   _advanced.py:126-195 step_advanced -> same kernel, ADV = true (h = h_abs*direction)
//   _basic.py:33-89    select_initial_step -> ni_small_init
//   _basic.py:203-242  RK.__init__     -> ni_small_init
//   _aux.py:71-75      norm            -> sequential unfused sum, /n, sqrt
//   _aux.py:77-115     tableaux        -> NiTab<>
//   _aux.py:16-19      SAFETY / MIN_FACTOR / MAX_FACTOR
// Arithmetic order follows SURVEY.md appendix B (what numba + OpenBLAS dgemv
// produce on the reference host) so that results are bit-identical to the
// reference wherever `pow` agrees with glibc: tableau contractions are FMA
// chains in dgemv order, every element-wise expression is unfused
// (__dmul_rn/__dadd_rn are never contracted, whatever -fmad says).
#pragma once
#ifndef NI_CORE_CUH
#define NI_CORE_CUH

#define NI_RK23 0
#define NI_RK45 1

#define NI_ST_RUNNING 0
#define NI_ST_FINISHED 1
#define NI_ST_FAILED 2     // step size fell below 8 ulp(t)   (_basic.py:179-180)
#define NI_ST_NONFINITE 3  // error norm is NaN (the reference would spin forever)

#define NI_MAX_SMALL_N 64
#ifndef NI_BLOCK
#define NI_BLOCK 128
#endif
// resident CTAs per SM the register allocator must allow for the run kernel.  n = 3 (Lorenz RK45 needs 94 registers):
// 5 x 128 threads without spills beat 6 x 128 at 80 registers with a few (+3 % on C3; before the error-norm tail
// became one basic block it was the other way round); n <= 2 fits 6 CTAs without spilling; larger systems keep all
// their stage vectors in registers and take what they need
#ifndef NI_MIN_BLOCKS
#define NI_MIN_BLOCKS(n) ((n) <= 2 ? 6 : (n) == 3 ? 5 : 1)
#endif

#ifndef NI_LARGE_MAX_THREADS
#define NI_LARGE_MAX_THREADS 256  // CTA-per-trajectory kernels (ni_large.cuh): threads per CTA (<= 16 components each)
#endif
#define NI_LARGE_MAX_N 4096  // largest state dimension of the CTA-per-trajectory kernels

#ifndef NI_LEAN
#define NI_LEAN 1  // 0: every launch takes the general run body (A/B experiments)
#endif

// A/B switches of the round-2 register-port work (profiles/r02_port_model.txt): all on in the
// product; `tools/variant.sh <tag> -DNI_EXP_...=0` builds a library without one of them for tools/exp_c3.sh.
#ifndef NI_EXP_ABSMAX_FP
#define NI_EXP_ABSMAX_FP 1
#endif
#ifndef NI_EXP_NACC
#define NI_EXP_NACC 1
#endif
#define NI_SINK_MAX 8  // destinations of the in-kernel terminal sink (the GPUs of one box)
#define NI_FLAG_REC_OVERFLOW 1ull
#define NI_REC_INVALID 0xffffffffu  // rec_traj value of a reserved-but-unused record slot

#define NI_DEV __device__ __forceinline__
#define NI_HD __host__ __device__

#include "ni_math.cuh"

// Kernel arguments (passed by value).  All per-trajectory arrays have N entries;
// y, klast are [n][N], params [P][N], aux [Q][N].
struct NiArgs {
  long long N;
  double *t, *h_abs, *step_size, *y, *klast, *aux;
  double *y_rows;               // small kernels: a row-major (N x n) copy of y kept current by init / retire, so that
                                // the host reads the state without a transposition kernel; may be NULL
  double *params;
  double *t_bound, *direction;
  int *status, *n_acc, *n_rej;
  unsigned char *running;       // result of the last step() call per trajectory
  const double *y0_rows;        // init only: N x n row-major
  const double *t0, *tb0;       // init only: initial time / bound, 1 or N entries
  int t0_stride, tb_stride;     // 0 (scalar) or 1 (per trajectory)
  const double *params_rows;    // init only: N x P row-major
  double max_step, first_step;
  const double *max_step_v, *first_step_v;  // per-trajectory values (N entries) that override the scalars, or NULL
  double rtol[NI_MAX_SMALL_N], atol[NI_MAX_SMALL_N];
  int n;                        // large kernels: state dimension (run-time)
  const double *rtol_d, *atol_d;  // large kernels: tolerance vectors in HBM (n entries)
  int tol_uniform;              // large kernels: every component has rtol[0] / atol[0]
  int atol_floor;               // small kernels: every atol[i] is finite and >= 2^-200, so every error scale is (the
                                // lean run body's division guard then only watches the upper end)
  int compat;                   // 0: the reference's step semantics; 1: scipy.integrate semantics
  unsigned char *step_rejected; // compat == 1: "a rejection happened in the current step" (scipy's step_rejected)
  unsigned long long *queue;    // [0] work cursor, [1] flags, [2] #retired still RUNNING, [3] #last step accepted,
                                // [4] #trajectories parked for the next compaction pass, [5] #retired FAILED / NONFINITE
  // compaction of the tail: once the queue is empty a warp with fewer than `park_below` busy
  // lanes parks its trajectories (state written back, index appended to `park_list`) and leaves; the next pass runs
  // the parked trajectories of all warps in full warps: its work items are `work_list[0 .. *work_count)`.
  const unsigned int *work_list;           // NULL: the work items are the trajectories 0 .. N-1
  const unsigned long long *work_count;    // device counter written by the previous pass (its queue[4])
  unsigned int *park_list;
  int park_below;                          // 0: never park
  long long max_attempts;       // per trajectory per launch (bounds kernel time)
  long long max_accepts;        // 1 = lock-step ni.step(), huge = fast-forward
  // recording of accepted steps (append log, SoA, capacity rec_cap records)
  unsigned long long *rec_cursor;
  long long rec_cap;
  long long rec_chunk;          // slots a warp reserves per atomicAdd (0: exactly as many as it needs)
  unsigned int *rec_traj;       // trajectory of the record in a slot (NI_REC_INVALID: reserved, never written); the step
                                // number is not stored: the records of one trajectory appear in step order
  double *rec_t, *rec_y, *rec_aux;  // rec_y [n][rec_cap], rec_aux [Q][rec_cap]
  // terminal sink (thread-per-trajectory kernels): a trajectory that leaves the run kernel for good also stores its
  // packed terminal record (t, y, aux, n_accepted, n_rejected, status; padded to 16-byte vectors) to row
  // sink_row0 + idx of up to NI_SINK_MAX device buffers -- this GPU's slot in the receive buffer of every GPU of the
  // job, peer memory over NVLink: the final gather overlaps the integration instead of following it
  unsigned int *sink[NI_SINK_MAX];
  int sink_n;                   // 0: no sink
  long long sink_row0;
  // stiffness probe: with probe_iters > 0 the INIT kernel does not initialise anything; it estimates the
  // spectral radius of df/dy at every trajectory's current (t, y) by a nonlinear power iteration and writes it to probe_out
  int probe_iters;
  double *probe_out;            // [N]
  // ff2cond: built-in predicate parameters
  int cond_kind;                // 0: no predicate; != 0: ff2cond launch of a kernel compiled with a predicate
  int cond_index;
  double cond_value;
  unsigned char *cond_hit;
  double cond_params[8];        // cond_kind == 5: parameters of the user predicate
};

// max_step / first_step of trajectory idx (_basic.py:292-299: per-object arguments of the reference's factories)
NI_DEV double ni_max_step(const NiArgs &a, long long idx) { return a.max_step_v ? a.max_step_v[idx] : a.max_step; }
NI_DEV double ni_first_step(const NiArgs &a, long long idx) { return a.first_step_v ? a.first_step_v[idx] : a.first_step; }

// 16-byte vectors of one terminal-sink row: the packed record of include/ni_b200.h (2 (1 + n + Q) + 3 words), padded
NI_HD constexpr int ni_sink_vectors(int n, int Q) { return (2 * (1 + n + Q) + 3 + 3) / 4; }
// The terminal record of trajectory idx -- read back from the state arrays its lane has just written (or that the
// init kernel wrote, for a trajectory that was finished before the launch) -- to every sink buffer.  Rare path.
template <int n, int Q>
NI_DEV void ni_sink_write(const NiArgs &a, long long idx) {
  if (a.sink_n <= 0) return;
  constexpr int DW = 2 * (1 + n + Q), V = ni_sink_vectors(n, Q);
  unsigned w[V * 4];
#pragma unroll
  for (int k = 0; k < V * 4; ++k) w[k] = 0u;
  auto put = [&](int c, double v) {
    w[2 * c] = (unsigned)__double2loint(v);
    w[2 * c + 1] = (unsigned)__double2hiint(v);
  };
  put(0, a.t[idx]);
#pragma unroll
  for (int i = 0; i < n; ++i) put(1 + i, a.y[i * a.N + idx]);
#pragma unroll
  for (int i = 0; i < Q; ++i) put(1 + n + i, a.aux[i * a.N + idx]);
  w[DW] = (unsigned)a.n_acc[idx];
  w[DW + 1] = (unsigned)a.n_rej[idx];
  w[DW + 2] = (unsigned)a.status[idx];
  // (unrolled over the destinations: a run-time index into the kernel arguments would copy them to local memory)
#pragma unroll
  for (int r = 0; r < NI_SINK_MAX; ++r) {
    if (r < a.sink_n) {
      uint4 *dst = reinterpret_cast<uint4 *>(a.sink[r]) + (a.sink_row0 + idx) * V;
#pragma unroll
      for (int k = 0; k < V; ++k) dst[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
  }
}

// Arithmetic policy.  Strict (FAST = false): every element-wise expression is evaluated unfused, as numba
// evaluates it in the reference (bit-identical results).  Fast: multiply-adds are
// contracted; agreement with the reference is then to tolerance only.
template <bool FAST>
struct NiA {
  NI_DEV static double mul(double a, double b) { return __dmul_rn(a, b); }
  NI_DEV static double add(double a, double b) { return __dadd_rn(a, b); }
  NI_DEV static double sub(double a, double b) { return __dadd_rn(a, -b); }
  NI_DEV static double madd(double a, double b, double c) {  // a*b + c
    if constexpr (FAST)
      return fma(a, b, c);
    else
      return __dadd_rn(__dmul_rn(a, b), c);
  }
  NI_DEV static double msub(double a, double b, double c) {  // a*b - c
    if constexpr (FAST)
      return fma(a, b, -c);
    else
      return __dadd_rn(__dmul_rn(a, b), -c);
  }
  NI_DEV static double nmadd(double a, double b, double c) {  // c - a*b
    if constexpr (FAST)
      return fma(-a, b, c);
    else
      return __dadd_rn(c, -__dmul_rn(a, b));
  }
};

// sign(x) for translated Python right-hand sides (np.sign)
NI_DEV double ni_sign(double x) { return (double)((x > 0.0) - (x < 0.0)); }

// ---------------------------------------------------------------- utilities
template <int I>
struct NiInt {
  static constexpr int value = I;
};
template <int B, int E, class F>
NI_DEV void ni_static_for(F &&f) {
  if constexpr (B < E) {
    f(NiInt<B>{});
    ni_static_for<B + 1, E>(f);
  }
}

// ---------------------------------------------------------------- tableaux
// Same rational expressions as _aux.py:77-115, folded in IEEE double at compile time.
template <int METHOD>
struct NiTab;

template <>
struct NiTab<NI_RK23> {
  static constexpr int S = 3;
  static constexpr int ORDER = 2;
  NI_HD static constexpr double a(int s, int j) {
    return s == 1 ? (j == 0 ? 1.0 / 2 : 0.0) : s == 2 ? (j == 1 ? 3.0 / 4 : 0.0) : 0.0;
  }
  NI_HD static constexpr double b(int j) { return j == 0 ? 2.0 / 9 : j == 1 ? 1.0 / 3 : j == 2 ? 4.0 / 9 : 0.0; }
  NI_HD static constexpr double c(int j) { return j == 1 ? 1.0 / 2 : j == 2 ? 3.0 / 4 : 0.0; }
  NI_HD static constexpr double e(int j) {
    return j == 0 ? 5.0 / 72 : j == 1 ? -1.0 / 12 : j == 2 ? -1.0 / 9 : j == 3 ? 1.0 / 8 : 0.0;
  }
};

template <>
struct NiTab<NI_RK45> {
  static constexpr int S = 6;
  static constexpr int ORDER = 4;
  NI_HD static constexpr double a(int s, int j) {
    switch (s * 8 + j) {
      case 1 * 8 + 0: return 1.0 / 5;
      case 2 * 8 + 0: return 3.0 / 40;
      case 2 * 8 + 1: return 9.0 / 40;
      case 3 * 8 + 0: return 44.0 / 45;
      case 3 * 8 + 1: return -56.0 / 15;
      case 3 * 8 + 2: return 32.0 / 9;
      case 4 * 8 + 0: return 19372.0 / 6561;
      case 4 * 8 + 1: return -25360.0 / 2187;
      case 4 * 8 + 2: return 64448.0 / 6561;
      case 4 * 8 + 3: return -212.0 / 729;
      case 5 * 8 + 0: return 9017.0 / 3168;
      case 5 * 8 + 1: return -355.0 / 33;
      case 5 * 8 + 2: return 46732.0 / 5247;
      case 5 * 8 + 3: return 49.0 / 176;
      case 5 * 8 + 4: return -5103.0 / 18656;
      default: return 0.0;
    }
  }
  NI_HD static constexpr double b(int j) {
    switch (j) {
      case 0: return 35.0 / 384;
      case 2: return 500.0 / 1113;
      case 3: return 125.0 / 192;
      case 4: return -2187.0 / 6784;
      case 5: return 11.0 / 84;
      default: return 0.0;
    }
  }
  NI_HD static constexpr double c(int j) {
    switch (j) {
      case 1: return 1.0 / 5;
      case 2: return 3.0 / 10;
      case 3: return 4.0 / 5;
      case 4: return 8.0 / 9;
      case 5: return 1.0;
      default: return 0.0;
    }
  }
  NI_HD static constexpr double e(int j) {
    switch (j) {
      case 0: return -71.0 / 57600;
      case 2: return 71.0 / 16695;
      case 3: return -71.0 / 1920;
      case 4: return 17253.0 / 339200;
      case 5: return -22.0 / 525;
      case 6: return 1.0 / 40;
      default: return 0.0;
    }
  }
};

// The same coefficients in constant memory: FP64 instructions take them as c[bank][offset] operands,
// which saves the two UMOVs per use that 64-bit immediates cost.  Index: row*8 + column, rows as in
// ni_coef (1..S-1 = A, S = B, S+1 = E), row 0 = C.
#define NI_TAB_ROW(T, r) T(r, 0), T(r, 1), T(r, 2), T(r, 3), T(r, 4), T(r, 5), T(r, 6), T(r, 7)
NI_HD constexpr double ni_tab23(int r, int j) {
  using T = NiTab<NI_RK23>;
  return r == 0 ? T::c(j) : r < T::S ? T::a(r, j) : r == T::S ? T::b(j) : r == T::S + 1 ? T::e(j) : 0.0;
}
NI_HD constexpr double ni_tab45(int r, int j) {
  using T = NiTab<NI_RK45>;
  return r == 0 ? T::c(j) : r < T::S ? T::a(r, j) : r == T::S ? T::b(j) : r == T::S + 1 ? T::e(j) : 0.0;
}
__constant__ double ni_ctab23[5 * 8] = {NI_TAB_ROW(ni_tab23, 0), NI_TAB_ROW(ni_tab23, 1), NI_TAB_ROW(ni_tab23, 2),
                                        NI_TAB_ROW(ni_tab23, 3), NI_TAB_ROW(ni_tab23, 4)};
__constant__ double ni_ctab45[8 * 8] = {NI_TAB_ROW(ni_tab45, 0), NI_TAB_ROW(ni_tab45, 1), NI_TAB_ROW(ni_tab45, 2),
                                        NI_TAB_ROW(ni_tab45, 3), NI_TAB_ROW(ni_tab45, 4), NI_TAB_ROW(ni_tab45, 5),
                                        NI_TAB_ROW(ni_tab45, 6), NI_TAB_ROW(ni_tab45, 7)};
// run-time value of coefficient (ROW, J); ROW = 0 gives C[J]
template <int METHOD, int ROW, int J>
NI_DEV double ni_cval() {
  if constexpr (METHOD == NI_RK45)
    return ni_ctab45[ROW * 8 + J];
  else
    return ni_ctab23[ROW * 8 + J];
}

// Coefficient of row ROW, column j.  ROW in 1..S-1 -> A[ROW][j]; ROW == S -> B[j]; ROW == S+1 -> E[j].
template <int METHOD, int ROW>
NI_HD constexpr double ni_coef(int j) {
  using T = NiTab<METHOD>;
  return ROW < T::S ? T::a(ROW, j) : ROW == T::S ? T::b(j) : T::e(j);
}

// sum_{j<COLS} K_j * coef(ROW, j) for ONE component, in the order OpenBLAS dgemv_n uses for an
// (n x COLS) matrix with n <= 3 rows (oracle/ni_oracle.c:dot_cols is the CPU
// twin): COLS <= 3 is a plain FMA chain; otherwise fma(K0,c0,K1*c1) + fma(K2,c2,K3*c3) (n >= 2) or
// the chain seeded with K1*c1 (n == 1, ONE_ROW), followed by the FMA tail j = 4...  `k(NiInt<j>{})`
// returns K_j and is only called for non-zero coefficients: fma(x, 0, acc) == acc for finite x, so
// the structural zeros of the tableau are skipped (the reference does not skip them; results differ
// only for non-finite K).
template <int METHOD, int ROW, int COLS, bool ONE_ROW, class G>
NI_DEV double ni_dot_col(G &&k) {
  double acc = 0.0;
  auto term = [&](auto jj, double acc_in, bool have) -> double {
    constexpr int j = decltype(jj)::value;
    const double cv = ni_cval<METHOD, ROW, j>();
    return have ? fma(k(jj), cv, acc_in) : __dmul_rn(k(jj), cv);
  };
  if constexpr (COLS <= 3) {
    bool have = false;
    ni_static_for<0, COLS>([&](auto jj) {
      if constexpr (ni_coef<METHOD, ROW>(decltype(jj)::value) != 0.0) {
        acc = term(jj, acc, have);
        have = true;
      }
    });
  } else if constexpr (ONE_ROW) {
    bool have = false;
    if constexpr (ni_coef<METHOD, ROW>(1) != 0.0) {
      acc = term(NiInt<1>{}, acc, false);
      have = true;
    }
    if constexpr (ni_coef<METHOD, ROW>(0) != 0.0) {
      acc = term(NiInt<0>{}, acc, have);
      have = true;
    }
    if constexpr (ni_coef<METHOD, ROW>(2) != 0.0) {
      acc = term(NiInt<2>{}, acc, have);
      have = true;
    }
    if constexpr (ni_coef<METHOD, ROW>(3) != 0.0) acc = term(NiInt<3>{}, acc, have);
  } else {
    auto pair = [&](auto ii, auto jj) -> double {  // fma(K_i, c_i, rn(K_j * c_j))
      constexpr int i = decltype(ii)::value, j = decltype(jj)::value;
      constexpr double ci = ni_coef<METHOD, ROW>(i), cj = ni_coef<METHOD, ROW>(j);
      if constexpr (ci == 0.0 && cj == 0.0)
        return 0.0;
      else if constexpr (cj == 0.0)
        return __dmul_rn(k(ii), ni_cval<METHOD, ROW, i>());
      else if constexpr (ci == 0.0)
        return __dmul_rn(k(jj), ni_cval<METHOD, ROW, j>());
      else
        return fma(k(ii), ni_cval<METHOD, ROW, i>(), __dmul_rn(k(jj), ni_cval<METHOD, ROW, j>()));
    };
    acc = __dadd_rn(pair(NiInt<0>{}, NiInt<1>{}), pair(NiInt<2>{}, NiInt<3>{}));
  }
  if constexpr (COLS > 4) {
    ni_static_for<4, COLS>([&](auto jj) {
      if constexpr (ni_coef<METHOD, ROW>(decltype(jj)::value) != 0.0) acc = term(jj, acc, true);
    });
  }
  return acc;
}

// out[i] = sum_{j<COLS} K[j][i] * coef(j) for the register-resident stage matrix of the small kernels
template <int METHOD, int ROW, int COLS, int n, int KR>
NI_DEV void ni_dot(const double (&K)[KR][n], double (&out)[n]) {
#pragma unroll
  for (int i = 0; i < n; ++i)
    out[i] = ni_dot_col<METHOD, ROW, COLS, n == 1>([&](auto jj) { return K[decltype(jj)::value][i]; });
}

// _aux.py:71-75
template <int n>
NI_DEV double ni_rms(const double (&x)[n]) {
  double acc = __dmul_rn(x[0], x[0]);
#pragma unroll
  for (int i = 1; i < n; ++i) acc = __dadd_rn(acc, __dmul_rn(x[i], x[i]));
  if constexpr (n == 3)
    return sqrt(ni_div3(acc));
  else
    return sqrt(acc / (double)n);
}

// x ** error_exponent in select_initial_step (_basic.py:87): same correctly-rounded evaluation as the
// controller where it applies, the library pow outside that range
template <int K>
NI_DEV double ni_pow_init(double x) {
  if (x > 7.888609052210118e-31 && x < 1.2676506002282294e+30) return ni_pow_core<K>(x);
  return pow(x, -1.0 / K);
}

// |nextafter(t, dir*inf) - t|  (_basic.py:139), for finite t.  Integer sign tests instead of FP64 compares: this
// runs once per attempt in the branch-free hot loop.
NI_DEV double ni_ulp_eps(double t, double dir) {
  const int th = __double2hiint(t), tl = __double2loint(t);
  const int s = (th ^ __double2hiint(dir)) >> 31;  // 0: the neighbour towards dir*inf is further from zero; -1: nearer
  const long long nb = __double_as_longlong(t) + (long long)(s | 1);
  const double d = __dadd_rn(__longlong_as_double(nb), -t);
  const double e = __hiloint2double(__double2hiint(d) & 0x7fffffff, __double2loint(d));  // fabs on the integer pipe
  const bool zero = ((th << 1) | tl) == 0;         // t == +-0: the smallest denormal, whatever the direction
  return zero ? __longlong_as_double(1LL) : e;
}

// Comparisons of the hot loop on the integer pipe (the FP64 pipe is the kernel's bound; a DSETP costs it as much as
// a DFMA).  For doubles that are >= +0 and not NaN (+inf allowed) the order of the values is the order of their bit
// patterns as signed 64-bit integers.
NI_DEV bool ni_lt_pos(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b); }
// max(|a|, |b|); differs from fmax(fabs(a), fabs(b)) only for NaN operands (it returns the NaN, fmax the number)
NI_DEV double ni_absmax(double a, double b) {
  // (the sign is cleared on the high WORD: nvcc turns a 64-bit `& 0x7fff...` back into an FP64-pipe |x|)
  const unsigned ah = (unsigned)__double2hiint(a) & 0x7fffffffu, bh = (unsigned)__double2hiint(b) & 0x7fffffffu;
  const unsigned long long x = ((unsigned long long)ah << 32) | (unsigned)__double2loint(a),
                           y = ((unsigned long long)bh << 32) | (unsigned)__double2loint(b);
  return __longlong_as_double((long long)(x > y ? x : y));
}

// ---------------------------------------------------------------- one attempt (shared by every run-kernel variant)
// Stages, y_new, the FSAL evaluation at (tn, y_new), the error estimate, and -- strict arithmetic -- error / scale,
// the RMS norm and the controller power as ONE basic block (_basic.py:151-169, :174/:178): the n quotients, the root
// and the power are the guarded branch-free sequences of ni_math.cuh, so their dependent chains interleave (the
// operators' slow-path branches would serialise them).  `out != 0` reports an operand outside the guarded ranges (a
// zero, denormal or non-finite error: the caller redoes the tail with the operators from `num`).  Fast arithmetic:
// `en` is the MEAN SQUARE error and the caller takes the controller power (ni_pow_fast*).
// kl = K[0] = K[-1] of the previous attempt: stale after a rejection (_basic.py:151).
template <class Rhs, int METHOD, bool FAST, bool LITE = false>
NI_DEV void ni_rk_attempt(const NiArgs &a, double t, double tn, double h, const double (&y)[Rhs::N],
                          const double (&kl)[Rhs::N], const double (&p)[Rhs::P > 0 ? Rhs::P : 1], double (&yn)[Rhs::N],
                          double (&kn)[Rhs::N], double (&auxn)[Rhs::Q > 0 ? Rhs::Q : 1], double (&num)[Rhs::N],
                          double &en, double &pw, bool &en_lt1, unsigned &out) {
  constexpr int n = Rhs::N, Q = Rhs::Q;
  using Tab = NiTab<METHOD>;
  constexpr int S = Tab::S;
  double K[S + 1][n];
#pragma unroll
  for (int i = 0; i < n; ++i) K[0][i] = kl[i];
  ni_static_for<1, S>([&](auto ss) {
    constexpr int s = decltype(ss)::value;
    double d[n], ys[n], auxd[Q > 0 ? Q : 1];
    ni_dot<METHOD, s, s, n>(K, d);
#pragma unroll
    for (int i = 0; i < n; ++i) ys[i] = NiA<FAST>::madd(d[i], h, y[i]);
    Rhs::template eval<FAST>(NiA<FAST>::madd(ni_cval<METHOD, 0, s>(), h, t), ys, p, K[s], auxd);
  });
  double d[n];
  ni_dot<METHOD, S, S, n>(K, d);
#pragma unroll
  for (int i = 0; i < n; ++i) yn[i] = NiA<FAST>::madd(h, d[i], y[i]);
  Rhs::template eval<FAST>(tn, yn, p, K[S], auxn);
  double ev[n];
  ni_dot<METHOD, S + 1, S + 1, n>(K, ev);
  if constexpr (FAST) {
    double msq = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const double sc = fma(fmax(fabs(y[i]), fabs(yn[i])), a.rtol[i], a.atol[i]);
      const double e = __dmul_rn(ev[i], __dmul_rn(h, ni_rcp_fast(sc)));
      msq = fma(e, e, msq);
      kn[i] = K[S][i];
      num[i] = 0.0;
    }
    en = __dmul_rn(msq, 1.0 / n);  // error_norm**2 (same < 1, == 0, NaN tests)
    pw = 0.0;
    en_lt1 = false;
    out = 0u;
  } else {
    double q[n];
    NiDivGuard guard = ni_guard_begin();
    int den_hi = 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      num[i] = __dmul_rn(ev[i], h);
#if NI_EXP_ABSMAX_FP
      // max(|y|, |y_new|) by ONE FP64 compare and a select of the raw operand (the multiply takes |.| for free): two
      // port cycles + two selects, where the integer-pipe version needs two sign masks and a two-word compare first.
      // The FP64 pipe has the slack (the kernel is bound by the register ports, section 3.2 of DESIGN.md).  A NaN in
      // y_new makes the compare false and is selected, i.e. propagates like np.maximum.
      const double big = LITE ? (fabs(y[i]) > fabs(yn[i]) ? y[i] : yn[i]) : ni_absmax(y[i], yn[i]);
      const double den = __dadd_rn(a.atol[i], __dmul_rn(fabs(big), a.rtol[i]));
#else
      const double den = __dadd_rn(a.atol[i], __dmul_rn(ni_absmax(y[i], yn[i]), a.rtol[i]));
#endif
      if constexpr (LITE) {
        const int hb = __double2hiint(den);
        den_hi = hb > den_hi ? hb : den_hi;
      } else {
        ni_guard_add(guard, num[i], den);
      }
      q[i] = ni_div_core(num[i], den);
      kn[i] = K[S][i];
    }
    double acc = __dmul_rn(q[0], q[0]);
#pragma unroll
    for (int i = 1; i < n; ++i) acc = __dadd_rn(acc, __dmul_rn(q[i], q[i]));
    // acc / n (_aux.py:75): exact for n = 1, 2; Markstein for n = 3; the guard on ms below covers acc otherwise
    const double ms = n == 1 ? acc : n == 2 ? __dmul_rn(acc, 0.5) : n == 3 ? ni_div3(acc) : ni_div_core(acc, (double)n);
    // LITE (the caller knows that every scale is >= 2^-200: NiArgs::atol_floor): only the upper end of the scales is
    // watched -- see ni_guard_lite_out in ni_math.cuh for why the errors need no test of their own
    out = (LITE ? ni_guard_lite_out(den_hi) : ni_guard_out(guard)) | ni_sqrt_out(ms);
    en = ni_sqrt_core(ms);
    pw = ni_pow_ctrl_ms<Tab::ORDER + 1>(ms, en);
    en_lt1 = (unsigned)__double2hiint(en) < 0x3ff00000u;  // en > 0 here
  }
}

// ---------------------------------------------------------------- predicates (ff2cond)
// The predicate of ni.ff2cond is compiled INTO a copy of the run kernel (NVRTC, ni_ensemble_ff2cond[_user]): the
// generated translation unit defines NI_HAVE_USER_COND and ni_user_cond after this header.  The ahead-of-time kernels
// carry no predicate code at all -- a run-time predicate switch cost the hot loop of every run ~15 issue slots per
// attempt and forced y[] through local memory (dynamic component index).
#ifdef NI_HAVE_USER_COND
__device__ __forceinline__ bool ni_user_cond(double t, const double *y, const double *aux, const double *cp);
#define NI_COND 1
#else
#define NI_COND 0
#endif

// ---------------------------------------------------------------- init kernel
// RK.__init__ (_basic.py:203-242): FSAL seed K[-1] = f(t0, y0), direction, error exponent,
// select_initial_step (_basic.py:33-89) unless first_step is given.  Also transposes the
// row-major host layout into the component-major device layout.
// Stiffness probe (a mode of the init kernel, so that every right-hand side -- built-in, NVRTC, Python-translated -- has
// it without another kernel): rho ~ spectral radius of J = df/dy at the trajectory's current (t, y) by the nonlinear
// power iteration of the stabilised-RK codes (v <- f(y + eps v) - f(y), rho_k = |v_new| / (eps |v|)), started from
// v = f(y).  h_abs * rho against the method's real stability interval (RK45 3.3, RK23 2.5) tells whether the
// controller's step is limited by stability rather than accuracy (Hairer / Wanner IV.2: the test DOPRI5 applies).
template <class Rhs>
NI_DEV void ni_small_probe_body(const NiArgs &a) {
  constexpr int n = Rhs::N, P = Rhs::P, Q = Rhs::Q;
  const long long N = a.N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < N;
       idx += (long long)gridDim.x * blockDim.x) {
    double y[n], f0[n], v[n], y1[n], f1[n], p[P > 0 ? P : 1], aux[Q > 0 ? Q : 1];
#pragma unroll
    for (int i = 0; i < n; ++i) y[i] = a.y[i * N + idx];
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = a.params[i * N + idx];
    const double t = a.t[idx];
    Rhs::template eval<false>(t, y, p, f0, aux);
    double yn2 = 0.0, vn2 = 0.0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      v[i] = f0[i];
      yn2 += y[i] * y[i];
      vn2 += v[i] * v[i];
    }
    if (!(vn2 > 0.0)) {  // an equilibrium (or a non-finite right-hand side): probe along (1, ..., 1)
#pragma unroll
      for (int i = 0; i < n; ++i) v[i] = 1.0;
      vn2 = (double)n;
    }
    const double ynorm = sqrt(yn2);
    double rho = 0.0;
    for (int it = 0; it < a.probe_iters; ++it) {
      const double vnorm = sqrt(vn2);
      const double eps = 1.4901161193847656e-08 * (1.0 + ynorm) / vnorm;  // sqrt(machine epsilon), relative to |y|
#pragma unroll
      for (int i = 0; i < n; ++i) y1[i] = y[i] + eps * v[i];
      Rhs::template eval<false>(t, y1, p, f1, aux);
      double dn2 = 0.0;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        v[i] = f1[i] - f0[i];
        dn2 += v[i] * v[i];
      }
      rho = sqrt(dn2) / (eps * vnorm);
      if (!(dn2 > 0.0)) break;  // J v = 0 (or not finite): rho stays what the last iteration found
      vn2 = dn2;
    }
    a.probe_out[idx] = rho;
  }
}

template <class Rhs, int METHOD>
NI_DEV void ni_small_init_body(const NiArgs &a) {
  constexpr int n = Rhs::N, P = Rhs::P, Q = Rhs::Q;
  const long long N = a.N;
  if (a.probe_iters > 0) {
    ni_small_probe_body<Rhs>(a);
    return;
  }
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < N;
       idx += (long long)gridDim.x * blockDim.x) {
    double y[n], f0[n], p[P > 0 ? P : 1], aux[Q > 0 ? Q : 1];
#pragma unroll
    for (int i = 0; i < n; ++i) y[i] = a.y0_rows[idx * n + i];
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = a.params_rows[idx * P + i];
    const double t0 = a.t0[idx * a.t0_stride], tb = a.tb0[idx * a.tb_stride];
    Rhs::template eval<false>(t0, y, p, f0, aux);
    const double dir = tb != t0 ? (tb > t0 ? 1.0 : -1.0) : 1.0;
    double h_abs;
    const double first_step = ni_first_step(a, idx);
    if (first_step != 0.0) {
      h_abs = fabs(first_step);
    } else if (a.compat != 0) {
      // scipy.integrate._ivp.common.select_initial_step: no fastmath contractions, h0 capped by the interval,
      // h1 = (0.01 / max(d1, d2)) ** (1 / (order + 1)), result capped by the interval and max_step
      const double interval = fabs(__dadd_rn(tb, -t0));
      double scale[n], tmp[n];
#pragma unroll
      for (int i = 0; i < n; ++i) scale[i] = __dadd_rn(a.atol[i], __dmul_rn(fabs(y[i]), a.rtol[i]));
#pragma unroll
      for (int i = 0; i < n; ++i) tmp[i] = y[i] / scale[i];
      const double d0 = ni_rms<n>(tmp);
#pragma unroll
      for (int i = 0; i < n; ++i) tmp[i] = f0[i] / scale[i];
      const double d1 = ni_rms<n>(tmp);
      double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : __dmul_rn(0.01, d0) / d1;
      h0 = fmin(h0, interval);
      const double hd = __dmul_rn(h0, dir);
      double y1[n], f1[n], auxd[Q > 0 ? Q : 1];
#pragma unroll
      for (int i = 0; i < n; ++i) y1[i] = __dadd_rn(y[i], __dmul_rn(hd, f0[i]));
      Rhs::template eval<false>(__dadd_rn(t0, hd), y1, p, f1, auxd);
#pragma unroll
      for (int i = 0; i < n; ++i) tmp[i] = __dadd_rn(f1[i], -f0[i]) / scale[i];
      const double d2 = ni_rms<n>(tmp) / h0;
      const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, __dmul_rn(h0, 1e-3))
                                                     : pow(0.01 / fmax(d1, d2), 1.0 / (NiTab<METHOD>::ORDER + 1));
      h_abs = interval == 0.0 ? 0.0 : fmin(fmin(__dmul_rn(100.0, h0), h1), fmin(interval, ni_max_step(a, idx)));
    } else {
      double scale[n], tmp[n];
#pragma unroll
      for (int i = 0; i < n; ++i) scale[i] = fma(fabs(y[i]), a.rtol[i], a.atol[i]);
#pragma unroll
      for (int i = 0; i < n; ++i) tmp[i] = y[i] / scale[i];
      const double d0 = ni_rms<n>(tmp);
#pragma unroll
      for (int i = 0; i < n; ++i) tmp[i] = f0[i] / scale[i];
      const double d1 = ni_rms<n>(tmp);
      const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : __dmul_rn(0.01, d0) / d1;
      const double hd = __dmul_rn(h0, dir);
      double y1[n], f1[n], auxd[Q > 0 ? Q : 1];
#pragma unroll
      for (int i = 0; i < n; ++i) y1[i] = fma(hd, f0[i], y[i]);
      Rhs::template eval<false>(__dadd_rn(t0, hd), y1, p, f1, auxd);
#pragma unroll
      for (int i = 0; i < n; ++i) tmp[i] = __dadd_rn(f1[i], -f0[i]) / scale[i];
      const double d2 = ni_rms<n>(tmp) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15)
        h1 = fmax(1e-6, __dmul_rn(h0, 1e-3));
      else
        h1 = ni_pow_init<NiTab<METHOD>::ORDER + 1>(__dmul_rn(fmax(d1, d2), 100.0));
      h_abs = fmin(__dmul_rn(100.0, h0), h1);
    }
    a.t[idx] = t0;
    a.t_bound[idx] = tb;
    a.direction[idx] = dir;
    a.h_abs[idx] = h_abs;
    a.step_size[idx] = __dmul_rn(dir, h_abs);
#pragma unroll
    for (int i = 0; i < n; ++i) {
      a.y[i * N + idx] = y[i];
      a.klast[i * N + idx] = f0[i];
    }
    if (a.y_rows) {
#pragma unroll
      for (int i = 0; i < n; ++i) a.y_rows[idx * n + i] = y[i];
    }
#pragma unroll
    for (int i = 0; i < P; ++i) a.params[i * N + idx] = p[i];
#pragma unroll
    for (int i = 0; i < Q; ++i) a.aux[i * N + idx] = aux[i];
    a.status[idx] = NI_ST_RUNNING;
    a.n_acc[idx] = 0;
    a.n_rej[idx] = 0;
    a.running[idx] = 1;
  }
}

// ---------------------------------------------------------------- run kernel
// Persistent grid; every lane owns one trajectory at a time and performs one step ATTEMPT per
// loop iteration (accept/reject only changes which values are committed, so a warp never
// diverges inside the stage loop).  A lane retires its trajectory when it finished, failed,
// used up max_accepts / max_attempts, hit the ff2cond predicate or found the record log full;
// it then pulls the next trajectory index from the queue.
// SEM: 0 = the reference's basic solvers (_step), 1 = Advanced (step_advanced: h = h_abs*direction),
// 2 = scipy.integrate step semantics (NI_MODE_SCIPY)
//
// Loop structure: the hot part (one attempt + controller + commit for all 32 lanes) is branch-free and is
// entered every iteration; lanes whose trajectory is done raise `need` and the warp takes the SERVICE branch
// (store the retired state, pop the next trajectory from the queue, reload) only when some lane needs it --
// about once per 18 iterations on C3.  Lanes without work keep executing the arithmetic on their stale
// registers with every commit masked off.
//
// MAXS: the launch has a finite max_step.  With the default max_step = inf the clamp of _basic.py:146-147 can never
// fire; compiled out it saves ten issue cycles per attempt, so ni_small_run_body carries both bodies and picks one per
// launch (the scipy variant reads max_step elsewhere too and keeps the general body).
template <class Rhs, int METHOD, int SEM, bool REC, bool FAST, bool MAXS>
NI_DEV void ni_small_run_impl(const NiArgs &a) {
  constexpr bool ADV = SEM >= 1, SP = SEM == 2, COND = NI_COND != 0;
  constexpr int n = Rhs::N, P = Rhs::P, Q = Rhs::Q;
  // The auxiliary output of an accepted step is only observable when the step is recorded, when a predicate reads it,
  // or when the trajectory leaves the kernel.  Without record log and predicate it is therefore evaluated ONCE, at
  // retirement, from the committed (t, y) -- the same function of the same operands as the FSAL evaluation that
  // produced K[-1] (_advanced.py:181), hence the same bits -- instead of once per attempt.
  constexpr bool LAZY_AUX = Q > 0 && !REC && !COND;
  using Tab = NiTab<METHOD>;
  constexpr int S = Tab::S;
  constexpr unsigned FULL = 0xffffffffu;
  const long long N = a.N;
  const unsigned lane = threadIdx.x & 31u;
  const long long n_work = a.work_list ? (long long)*a.work_count : N;
  const int max_att = a.max_attempts > 0x7fffffffLL ? 0x7fffffff : (int)a.max_attempts;
  const int max_acc = a.max_accepts > 0x7fffffffLL ? 0x7fffffff : (int)a.max_accepts;

  long long idx = 0;
  bool active = false, exhausted = false, need = true, last_accept = false, parked = false, srej = false;
  bool compact = false;  // this lane's trajectory goes to the next compaction pass
  double t = 0, h_abs = 0, tb = 0, dir = 1, step_h = 0;
  double mxs = a.max_step;  // max_step of the lane's trajectory (MAXS / scipy bodies)
  double y[n], kl[n], p[P > 0 ? P : 1], aux[Q > 0 ? Q : 1];
#pragma unroll
  for (int i = 0; i < n; ++i) y[i] = kl[i] = 0.0;
#pragma unroll
  for (int i = 0; i < (P > 0 ? P : 1); ++i) p[i] = 0.0;
#pragma unroll
  for (int i = 0; i < (Q > 0 ? Q : 1); ++i) aux[i] = 0.0;
  int status = NI_ST_RUNNING;
  int att_left = 0, acc_left = 0;
  // REC: this warp's reserved chunk of the record log: slots [rc_base + rc_off, rc_base + rc_len) are free.  The
  // chunk start is 64-bit and changes once per chunk; the per-attempt bookkeeping runs on the 32-bit offsets.
  long long rc_base = 0;
  int rc_off = 0, rc_len = 0;
  const unsigned lanes_below = (1u << lane) - 1u;

  for (;;) {
    // ================================================================ SERVICE (rare)
    if (__any_sync(FULL, need)) {
      if (need) {
        if (active) {  // retire: write the trajectory back
          // fast-forward: the terminal step() call of `while step(): pass` (_basic.py:135-136) overwrites step_size
          if (status == NI_ST_FINISHED) step_h = ADV ? h_abs : __dmul_rn(dir, h_abs);
          a.t[idx] = t;
          a.h_abs[idx] = h_abs;
          a.step_size[idx] = step_h;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            a.y[i * N + idx] = y[i];
            a.klast[i * N + idx] = kl[i];
          }
          if (a.y_rows) {
#pragma unroll
            for (int i = 0; i < n; ++i) a.y_rows[idx * n + i] = y[i];
          }
          // the two budgets are the counters: accepted = accepts used, rejected = counted attempts - accepted
          const int n_accepted = max_acc - acc_left, n_counted = max_att - att_left;
          if constexpr (LAZY_AUX) {
            // (t, y) changed in this residency <=> a step was accepted or the failure exit committed its attempt
            if (n_accepted > 0 || status == NI_ST_FAILED) {
              double kd[n];
              Rhs::template eval<FAST>(t, y, p, kd, aux);
#pragma unroll
              for (int i = 0; i < Q; ++i) a.aux[i * N + idx] = aux[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < Q; ++i) a.aux[i * N + idx] = aux[i];
          }
          a.status[idx] = status;
          a.n_acc[idx] += n_accepted;
          a.n_rej[idx] += n_counted - n_accepted;
          a.running[idx] = last_accept ? 1 : 0;
          if (SP) a.step_rejected[idx] = srej ? 1 : 0;
          if (compact) {
            a.park_list[atomicAdd(a.queue + 4, 1ull)] = (unsigned int)idx;
            compact = false;
          } else if (status == NI_ST_RUNNING && !parked) {
            atomicAdd(a.queue + 2, 1ull);
          }
          if (status >= NI_ST_FAILED) atomicAdd(a.queue + 5, 1ull);
          if (last_accept) atomicAdd(a.queue + 3, 1ull);
          if (status != NI_ST_RUNNING) ni_sink_write<n, Q>(a, idx);
          active = false;
        }
        // refill: pop trajectories (warp-aggregated) until one is runnable or the queue is empty
        while (!active && !exhausted) {
          const unsigned grp = __activemask();
          const int leader = __ffs(grp) - 1;
          unsigned long long base = 0;
          if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(grp));
          base = __shfl_sync(grp, base, leader);
          idx = (long long)(base + __popc(grp & ((1u << lane) - 1u)));
          if (idx >= n_work || (REC && (((volatile unsigned long long *)a.queue)[1] & NI_FLAG_REC_OVERFLOW))) {
            exhausted = true;
            break;
          }
          if (a.work_list) idx = a.work_list[idx];
          status = a.status[idx];
          if (COND && a.cond_hit[idx]) status = -1;  // ff2cond: already stopped at its predicate
          if (status != NI_ST_RUNNING) {
            if (status >= 0) {
              a.running[idx] = 0;  // step() on a finished / failed solver returns False
              ni_sink_write<n, Q>(a, idx);
            }
            continue;
          }
          t = a.t[idx];
          h_abs = a.h_abs[idx];
          tb = a.t_bound[idx];
          dir = a.direction[idx];
          if (__dmul_rn(dir, __dadd_rn(t, -tb)) >= 0.0) {  // t_bound has been reached (_basic.py:135-136)
            a.step_size[idx] = ADV ? h_abs : __dmul_rn(dir, h_abs);  // _advanced.py:151 / _basic.py:136
            a.status[idx] = NI_ST_FINISHED;
            a.running[idx] = 0;
            ni_sink_write<n, Q>(a, idx);
            continue;
          }
          step_h = a.step_size[idx];
          if (MAXS || SP) mxs = ni_max_step(a, idx);
#pragma unroll
          for (int i = 0; i < n; ++i) {
            y[i] = a.y[i * N + idx];
            kl[i] = a.klast[i * N + idx];
          }
#pragma unroll
          for (int i = 0; i < P; ++i) p[i] = a.params[i * N + idx];
          if constexpr (!LAZY_AUX) {
#pragma unroll
            for (int i = 0; i < Q; ++i) aux[i] = a.aux[i * N + idx];
          }
          att_left = max_att;
          acc_left = max_acc;
          last_accept = false;
          parked = false;
          srej = SP ? a.step_rejected[idx] != 0 : false;
          status = NI_ST_RUNNING;
          active = true;
        }
        need = false;
      }
      if (a.park_below > 0) {
        // compaction: the queue is empty and this warp is mostly idle -- hand its trajectories to the next pass,
        // where they share full warps with the leftovers of the other warps (the state between two attempts is
        // exactly t, y, h_abs, K[-1]: the regular retire path above writes it back on the next turn)
        const int nact = __popc(__ballot_sync(FULL, active));
        if (nact > 0 && nact < a.park_below && __any_sync(FULL, exhausted)) {
          exhausted = true;
          if (active) {
            need = true;
            compact = true;
          }
          continue;
        }
      }
      if (__all_sync(FULL, !active)) break;
    }

    // ================================================================ HOT: one attempt for every lane
    const bool act = active;
    // ---- step entry (_basic.py:139-143); recomputed per attempt: between retries t does not change and the
    // clamp is a no-op (a retry only happens with h_abs >= min_step)
    const double eps = ni_ulp_eps(t, dir);
    const double min_step = __dmul_rn(SP ? 10.0 : 8.0, eps);  // scipy: 10 ulp
    // (h_abs, min_step, max_step are >= +0: integer-pipe comparisons, ni_lt_pos)
    if (SP) {
      // first attempt of a step <=> no rejection seen in it yet (also right after resuming a parked trajectory)
      if (!srej) h_abs = h_abs > mxs ? mxs : (h_abs < min_step ? min_step : h_abs);
    } else {
      h_abs = ni_lt_pos(h_abs, min_step) ? min_step : h_abs;
    }
    // ---- one attempt (_basic.py:145-169)
    const double h_pre = h_abs;
    const bool sp_fail = SP && h_abs < min_step;  // scipy tests the step size BEFORE every attempt, commits nothing
    if (!SP && MAXS && ni_lt_pos(mxs, h_abs)) h_abs = __dadd_rn(mxs, -eps);
    // h_abs * direction (_advanced.py:163) with direction = +-1: a sign flip on the integer pipe, bit for bit the product
    double h = (ADV || SP) ? __hiloint2double(__double2hiint(h_abs) ^ (__double2hiint(dir) & (int)0x80000000), __double2loint(h_abs))
                           : h_abs;
    double tn = __dadd_rn(t, h);
    bool reach;  // direction * (t_new - t_bound) >= 0: this attempt ends on t_bound
    {
      // sign tests on the bits of t_new - t_bound instead of the product with direction (= +-1)
      const double dtb = __dadd_rn(tn, -tb);
      const bool zero = ((__double2hiint(dtb) << 1) | __double2loint(dtb)) == 0;
      const bool same = (__double2hiint(dtb) ^ __double2hiint(dir)) >= 0;
      reach = zero || same;
      const bool clip = SP ? (same && !zero) : reach;
      tn = clip ? tb : tn;
      if (clip || SP) {  // scipy recomputes h = t_new - t on every attempt
        h = __dadd_rn(tn, -t);
        h_abs = __hiloint2double(__double2hiint(h) & 0x7fffffff, __double2loint(h));  // fabs on the integer pipe
      }
    }
    double en, pw, kn[n], yn[n], auxn[Q > 0 ? Q : 1], num[n];  // error norm (fast: its square), en ** error_exponent
    bool en_lt1, en_nan = false;
    unsigned out;
    ni_rk_attempt<Rhs, METHOD, FAST>(a, t, tn, h, y, kl, p, yn, kn, auxn, num, en, pw, en_lt1, out);
    if constexpr (FAST) {
      pw = ni_pow_fast<2 * (Tab::ORDER + 1)>(en);
      en_lt1 = en < 1.0;
      en_nan = en != en;
    } else {
      if (act && out != 0u) {  // an operand outside the guarded ranges: the tail again, with the operators
        double ev[n];
#pragma unroll
        for (int i = 0; i < n; ++i)
          ev[i] = num[i] / __dadd_rn(a.atol[i], __dmul_rn(fmax(fabs(y[i]), fabs(yn[i])), a.rtol[i]));
        en = ni_rms<n>(ev);
        pw = ni_pow_ctrl<Tab::ORDER + 1>(en);
        en_lt1 = en < 1.0;
        en_nan = en != en;
      }
    }
    const bool ok = act && en_lt1 && !sp_fail;

    // ---- record log: one slot per accepting lane.  A warp reserves `rec_chunk` slots with ONE atomicAdd
    // and hands them out locally (a single global cursor saturates the L2 atomic unit at ~0.8 G
    // reservations/s: config C2 was bound by it); rec_chunk == 0 reserves exactly popc(mask) per iteration
    // (lock-step ni.step, where a chunk would be abandoned after one use).  Slots of an abandoned chunk are
    // marked NI_REC_INVALID so that the drain can skip them.
    long long slot = -1;
    bool rec_full = false;  // this lane's record does not fit into the log any more
    if (REC) {
      const unsigned m = __ballot_sync(FULL, ok);
      const int cnt = __popc(m), rank = __popc(m & lanes_below);
      const int off = rc_off + rank;
      if (rc_off + cnt <= rc_len) {  // the usual case: the chunk has room for every accepting lane (also cnt == 0)
        slot = rc_base + off;        // < rec_cap by construction of rc_len
        rc_off += cnt;
      } else {  // warp-uniform: take the rest of the chunk, reserve the next one for the lanes that do not fit
        const int room = rc_len - rc_off;
        const long long want = a.rec_chunk > 0 ? a.rec_chunk : (long long)(cnt - room);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(a.rec_cursor, (unsigned long long)want);
        const long long nb = (long long)__shfl_sync(FULL, base, 0);
        slot = rank < room ? rc_base + off : nb + (rank - room);
        rec_full = slot >= a.rec_cap;
        const long long avail = a.rec_cap - nb;  // slots of the new chunk that exist
        rc_base = nb;
        rc_len = (int)(avail < 0 ? 0 : (avail < want ? avail : want));
        rc_off = cnt - room < rc_len ? cnt - room : rc_len;
      }
    }

    // ---- controller + commit (_basic.py:171-180), branch-free: every update is a select on the lane's flags
    {
      const bool nan = act && en_nan;
      // park: log full -> undo this attempt (the state before an attempt is exactly t, y, h_abs, K[-1]);
      // scipy's too-small-step exit also commits nothing of the attempt
      const bool park = (REC && ok && rec_full) || (act && sp_fail);
      const bool upd = act && !park;  // this attempt counts
      // SAFETY * en ** exp; en == 0 gives MAX_FACTOR either way (:172)
      const double sp = __dmul_rn(NI_MC(0, 0.9), pw);
      // accepted: min(MAX_FACTOR, sp) (en == 0 saturates ni_pow_ctrl, i.e. MAX_FACTOR, :172); rejected:
      // max(MIN_FACTOR, sp).  sp > 0 and never NaN (ni_pow_ctrl saturates): integer-pipe comparisons
      // An accepted attempt has en < 1, i.e. sp > 0.9 > MIN_FACTOR, and a rejected one en >= 1, i.e. sp <= 0.9 <
      // MAX_FACTOR: the other clamp is a no-op either way, so both are applied unconditionally.  10.0 has a zero low
      // word: sp < 10 <=> hi(sp) < hi(10).
      double fac = __double2hiint(sp) < 0x40240000 ? sp : 10.0;
      fac = ni_lt_pos(0.2, fac) ? fac : 0.2;
      if (SP && ok && srej) fac = fmin(1.0, fac);  // scipy: no growth in a step that saw a rejection
      const double h_new = __dmul_rn(h_abs, fac);
      // :179-180: commits the REJECTED t, y (h_new, min_step >= +0; a NaN h_new implies `nan`)
      const bool fail = upd && !SP && !ok && !nan && ni_lt_pos(h_new, min_step);
      const bool accepted = ok && !park;
      const bool take = accepted || fail;
      h_abs = upd ? (nan ? h_abs : h_new) : h_pre;
      t = take ? tn : t;
      step_h = (take || nan) ? h : step_h;
      const bool newk = upd && !(SP && !ok);  // scipy keeps f(t, y) of the accepted point on a rejection
#pragma unroll
      for (int i = 0; i < n; ++i) {
        y[i] = take ? yn[i] : y[i];
        // without record log and scipy semantics newk == act, and a lane without work does not care about its K[-1]
        if constexpr (SP || REC)
          kl[i] = newk ? kn[i] : kl[i];
        else
          kl[i] = kn[i];
      }
      if constexpr (!LAZY_AUX) {
#pragma unroll
        for (int i = 0; i < Q; ++i) aux[i] = take ? auxn[i] : aux[i];
      }
      acc_left -= (int)accepted;
      att_left -= (int)upd;
      last_accept = accepted;  // only read when the lane retires its trajectory, i.e. while `act`
      srej = upd ? !ok : srej;
      if (REC && accepted) {
        const long long c = a.rec_cap;
        a.rec_traj[slot] = (unsigned int)idx;
        a.rec_t[slot] = t;
#pragma unroll
        for (int i = 0; i < n; ++i) a.rec_y[i * c + slot] = y[i];
#pragma unroll
        for (int i = 0; i < Q; ++i) a.rec_aux[i * c + slot] = aux[i];
      }
      // ---- exits: raise `need` for the service branch
      bool done = park || fail || nan;
      if (upd && att_left <= 0) done = true;
      if (accepted && acc_left <= 0) done = true;
      if (REC && park && !sp_fail) atomicOr(a.queue + 1, NI_FLAG_REC_OVERFLOW);
      if (fail || (act && sp_fail)) status = NI_ST_FAILED;
      if (nan) status = NI_ST_NONFINITE;
#if NI_COND
      if (accepted && ni_user_cond(t, y, aux, a.cond_params)) {
        a.cond_hit[idx] = 1;
        parked = true;
        done = true;
      }
#endif
      if (accepted && !done && reach) {  // t == t_bound now: the attempt was clipped to it or landed on it
        // fast-forward: this was the last step (the retire path emulates the terminal step() call)
        status = NI_ST_FINISHED;
        last_accept = false;
        done = true;
      }
      need = done;
    }
  }
  if (REC) {  // slots reserved but never written
    for (int o = rc_off + (int)lane; o < rc_len; o += 32) a.rec_traj[rc_base + o] = NI_REC_INVALID;
  }
}

// ---------------------------------------------------------------- run kernel, lean variant: cold block
// Cold block of the lean run body: the attempt of this lane needs one of the rarely needed pieces of the controller --
// operands outside the guarded division / root ranges (the tail again, with the operators), the NaN latch (the
// reference would spin forever: status NONFINITE) or the step-size floor of the failure exit (_basic.py:179-180) --
// recomputed with the general formulas of ni_small_run_impl.  Out of line and by value, so that the hot loop pays one
// predicated branch for it and keeps its flags in predicate registers.
// flags out: 1 = accepted, 4 = failure exit (commit the rejected t, y), 8 = leave with `status`.
struct NiLeanCold {
  double h_new;
  int flags, status;
};
template <int n>
struct NiLeanColdIn {
  double num[n], y[n], yn[n];
};
template <int n, int METHOD, bool FAST>
static __device__ __noinline__ NiLeanCold ni_lean_cold(NiLeanColdIn<n> v, const double *rtol, const double *atol, double t,
                                                       double dir, double h_abs, double en, double pw, int ok_in, unsigned out) {
  using Tab = NiTab<METHOD>;
  bool ok = ok_in != 0, nan = false;
  if constexpr (FAST) {
    if (out != 0u) {  // (en holds the mean square error) outside [2^-100, 2^100) or NaN: the saturating root
      pw = ni_pow_fast<2 * (Tab::ORDER + 1)>(en);
      ok = en < 1.0;
      nan = en != en;
    }
  } else {
    if (out != 0u) {
      double ev[n];
#pragma unroll
      for (int i = 0; i < n; ++i)
        ev[i] = v.num[i] / __dadd_rn(atol[i], __dmul_rn(fmax(fabs(v.y[i]), fabs(v.yn[i])), rtol[i]));
      en = ni_rms<n>(ev);
      pw = ni_pow_ctrl<Tab::ORDER + 1>(en);
      ok = en < 1.0;
      nan = en != en;
    }
  }
  const double sp = __dmul_rn(0.9, pw);
  const double fac = fmax(0.2, fmin(10.0, sp));  // en == 0 saturates the power, i.e. MAX_FACTOR (:172)
  NiLeanCold r;
  r.h_new = nan ? h_abs : __dmul_rn(h_abs, fac);
  const double min_step = __dmul_rn(8.0, ni_ulp_eps(t, dir));
  const bool fail = !ok && !nan && r.h_new < min_step;  // :179-180
  r.status = nan ? NI_ST_NONFINITE : fail ? NI_ST_FAILED : NI_ST_RUNNING;
  r.flags = (ok && !nan ? 1 : 0) | (fail ? 4 : 0) | (fail || nan ? 8 : 0);
  return r;
}

// ---------------------------------------------------------------- run kernel, lean variant
// The same algorithm for the launches that need no per-attempt bookkeeping: reference semantics (SEM 0 / 1), no record
// log or a chunked one, no predicate, max_step = inf, no attempt or accept budget (fast-forward; lock-step ni.step
// takes the general body).  This is every benchmark launch.
//
// Why a second body: the thread-per-trajectory kernel is bound by REGISTER-FILE READ BANDWIDTH, not by the FP64 pipe
// (tools/micro/fp64_shadow.cu, profiles/r02_issue_model.txt: a scheduler reads two 32-bit register operands per lane
// per cycle; an FP64 instruction with two 64-bit register sources uses all of it for its two pipe cycles, so every
// register operand of every other instruction adds half a cycle).  The general body above spends ~160 such operands per
// attempt on selects, 64-bit integer compares and counters.  Here everything that is not needed on the usual path
// is moved to places that run rarely:
//  * step entry (_basic.py:139-156: 8 ulp(t) floor, t + h, clip to t_bound) is evaluated EXACTLY only where it can
//    matter, found by two conservative integer tests: "t + h may have reached t_bound" (a monotone key of the high word
//    of t + h against the key of t_bound, 2 instructions; the lane then takes a short rare block with the exact clip)
//    and "h_abs may be below 8 ulp(t)" (high word of h_abs against a power-of-two bound from the exponents of t and
//    t_bound, 1 instruction; the lane then takes the cold block and the exact entry of the service branch); otherwise
//    the entry is h = +-h_abs, tn = t + h, bit for bit what the exact entry computes when neither clamp applies;
//  * the failure exit (:179-180), the NaN latch and operands outside the guarded division / root ranges share ONE
//    out-of-line cold block behind the second conservative test; a lane that takes it commits its attempt there;
//  * attempts are counted by a warp-uniform iteration counter (rejected = iterations in residency - accepted), status /
//    step_size / running are derived at retirement, lanes without work are not masked: they run on benign values and
//    cannot raise a flag.
template <class Rhs, int METHOD, bool ADV, bool FAST, bool REC>
NI_DEV void ni_small_run_lean(const NiArgs &a) {
  constexpr int n = Rhs::N, P = Rhs::P, Q = Rhs::Q;
  using Tab = NiTab<METHOD>;
  constexpr unsigned FULL = 0xffffffffu;
  const long long N = a.N;
  const unsigned lane = threadIdx.x & 31u;

  long long idx = 0;
  bool active = false, exhausted = false, svc = true;
  double t = 0, h_abs = 0, tb = 0, dir = 1, tn = 0, step_h = 0;
  double hs = 0;  // the signed step of the coming attempt, +-h_abs: carried from the step entry to the attempt
  double y[n], kl[n], p[P > 0 ? P : 1];
#pragma unroll
  for (int i = 0; i < n; ++i) y[i] = kl[i] = 0.0;
#pragma unroll
  for (int i = 0; i < (P > 0 ? P : 1); ++i) p[i] = 0.0;
  int status = NI_ST_RUNNING, nacc = 0, acc0 = 0, rej0 = 0;
  unsigned iter = 0, it0 = 0;  // hot-loop passes of this warp; the pass count at which the lane's trajectory was loaded
  // REC: this warp's chunk of the record log: slots [rc_base + rc_off, rc_base + rc_len) are free.  An attempt only
  // starts with room for all 32 lanes (the service branch renews the chunk), so an accepted step always has its slot and
  // a full log stops trajectories BETWEEN attempts: nothing has to be undone.
  long long rc_base = 0;
  int rc_off = 0, rc_len = 0;
  bool logfull = false;
  const unsigned lanes_below = (1u << lane) - 1u;
  int kx = 0, tbk = 0x7fffffff;  // reach key: (hi(tn) ^ kx) >= tbk  <=  direction * (tn - t_bound) >= 0
  int xs = 0;                    // hi(h_abs) <u xs  <=  h_abs may be below 8 ulp(t) somewhere between t and t_bound

  const long long n_work = a.work_list ? (long long)*a.work_count : N;
  // write the lane's trajectory back.  mode 0: it has finished / failed; 1: it moves to the next compaction pass;
  // 2: the record log is full (it stays RUNNING until the host has drained the log)
  auto retire = [&](int mode) {
        const bool parking = mode == 1;
        const int n_attempts = (int)(iter - it0);
        if (status == NI_ST_RUNNING && mode == 0) {
          status = NI_ST_FINISHED;
          // the terminal step() call of `while step(): pass` (_basic.py:135-136) overwrites step_size
          step_h = ADV ? h_abs : __dmul_rn(dir, h_abs);
        }
        a.t[idx] = t;
        a.h_abs[idx] = h_abs;
        a.step_size[idx] = step_h;
#pragma unroll
        for (int i = 0; i < n; ++i) {
          a.y[i * N + idx] = y[i];
          a.klast[i * N + idx] = kl[i];
        }
        if (a.y_rows) {
#pragma unroll
          for (int i = 0; i < n; ++i) a.y_rows[idx * n + i] = y[i];
        }
        if constexpr (Q > 0) {
          // the auxiliary output is the one of the FSAL evaluation at the committed (t, y) (_advanced.py:181): same
          // function, same operands, same bits, evaluated once here instead of once per attempt
          if (nacc > 0 || status == NI_ST_FAILED) {
            double kd[n], aux[Q];
            Rhs::template eval<FAST>(t, y, p, kd, aux);
#pragma unroll
            for (int i = 0; i < Q; ++i) a.aux[i * N + idx] = aux[i];
          }
        }
        a.status[idx] = status;
        a.n_acc[idx] = acc0 + nacc;  // (acc0 / rej0: the counters as loaded, so that the retire path only stores)
        a.n_rej[idx] = rej0 + n_attempts - nacc;
        a.running[idx] = 0;
        if (parking) a.park_list[atomicAdd(a.queue + 4, 1ull)] = (unsigned int)idx;
        if (mode == 2) atomicAdd(a.queue + 2, 1ull);
        if (status >= NI_ST_FAILED) atomicAdd(a.queue + 5, 1ull);
        if (status != NI_ST_RUNNING) ni_sink_write<n, Q>(a, idx);
        active = false;
  };

  for (;;) {
    // ================================================================ SERVICE (rare)
    if (__any_sync(FULL, svc)) {
      // exit: the cold block latched a failure, or the last accepted step ended on t_bound (the reference's own test
      // at the top of its next step() call, _basic.py:135)
      if (active && (status != NI_ST_RUNNING || __dmul_rn(dir, __dadd_rn(t, -tb)) >= 0.0)) retire(0);
      // refill: pop trajectories (warp-aggregated) until one is runnable or the queue is empty
      while (!active && !exhausted) {
        const unsigned grp = __activemask();
        const int leader = __ffs(grp) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)__popc(grp));
        base = __shfl_sync(grp, base, leader);
        idx = (long long)(base + __popc(grp & ((1u << lane) - 1u)));
        if (idx >= n_work || (REC && (((volatile unsigned long long *)a.queue)[1] & NI_FLAG_REC_OVERFLOW))) {
          exhausted = true;
          break;
        }
        if (a.work_list) idx = a.work_list[idx];
        // every load of the trajectory is issued before the first one is needed: one memory round trip per refill
        // (a fast-forward run almost never pops a trajectory that is not RUNNING)
        status = a.status[idx];
        t = a.t[idx];
        h_abs = a.h_abs[idx];
        tb = a.t_bound[idx];
        dir = a.direction[idx];
        step_h = a.step_size[idx];
        acc0 = a.n_acc[idx];
        rej0 = a.n_rej[idx];
#pragma unroll
        for (int i = 0; i < n; ++i) {
          y[i] = a.y[i * N + idx];
          kl[i] = a.klast[i * N + idx];
        }
#pragma unroll
        for (int i = 0; i < P; ++i) p[i] = a.params[i * N + idx];
        if (status != NI_ST_RUNNING) {
          a.running[idx] = 0;  // step() on a finished / failed solver returns False
          ni_sink_write<n, Q>(a, idx);
          continue;
        }
        if (__dmul_rn(dir, __dadd_rn(t, -tb)) >= 0.0) {  // t_bound has been reached (_basic.py:135-136)
          a.step_size[idx] = ADV ? h_abs : __dmul_rn(dir, h_abs);  // _advanced.py:151 / _basic.py:136
          a.status[idx] = NI_ST_FINISHED;
          a.running[idx] = 0;
          ni_sink_write<n, Q>(a, idx);
          continue;
        }
        // reach key of this trajectory: hi(tn) ^ kx is monotone in dir * tn over the sign regime in which t_bound can
        // be reached, and compares as "reached" outside it
        const int tbh = __double2hiint(tb);
        if (__double2hiint(dir) >= 0) {
          kx = tbh >= 0 ? 0 : 0x7fffffff;
        } else {
          kx = tbh < 0 ? (int)0x80000000 : -1;
        }
        tbk = tbh ^ kx;
        // step-size floor test of this residency: t moves from here towards t_bound, so the larger of the two
        // exponents bounds 8 ulp(t) for every attempt (64 ulp of the larger magnitude: conservative by 8x or more)
        const int ea = __double2hiint(t) & 0x7ff00000, eb = tbh & 0x7ff00000;
        const int xe = (ea > eb ? ea : eb) - (46 << 20);
        xs = xe > 0x00100000 ? xe : 0x00100000;
        nacc = 0;
        it0 = iter;
        active = true;
      }
      if constexpr (REC) {
        // (after the refill: a warp without work takes no chunk away from the others)
        if (rc_len - rc_off < 32 && __any_sync(FULL, active)) {  // (warp-uniform) renew the chunk: leftover slots are marked unused
          for (int o = rc_off + (int)lane; o < rc_len; o += 32) a.rec_traj[rc_base + o] = NI_REC_INVALID;
          const long long want = a.rec_chunk;
          unsigned long long base = 0;
          if (lane == 0) base = atomicAdd(a.rec_cursor, (unsigned long long)want);
          rc_base = (long long)__shfl_sync(FULL, base, 0);
          const long long avail = a.rec_cap - rc_base;
          rc_off = 0;
          rc_len = (int)(avail < 0 ? 0 : (avail < want ? avail : want));
          if (rc_len < 32) {  // the log is full: no attempt starts, every trajectory of the warp waits for the drain
            for (int o = (int)lane; o < rc_len; o += 32) a.rec_traj[rc_base + o] = NI_REC_INVALID;
            rc_len = 0;
            if (lane == 0) atomicOr(a.queue + 1, NI_FLAG_REC_OVERFLOW);
            logfull = true;
          }
        }
        if (logfull || (((volatile unsigned long long *)a.queue)[1] & NI_FLAG_REC_OVERFLOW)) {
          if (logfull && active) retire(2);
          exhausted = true;  // no new trajectory starts against a full log
        }
      }
      if (a.park_below > 0) {
        // compaction: the queue is empty and this warp is mostly idle -- its trajectories go to the next pass, where
        // they share full warps with the leftovers of the other warps (the state between two attempts is exactly
        // t, y, h_abs, K[-1])
        const int nact = __popc(__ballot_sync(FULL, active));
        if (nact > 0 && nact < a.park_below && __any_sync(FULL, exhausted)) {
          if (active) retire(1);
          exhausted = true;
        }
      }
      if (active) {
        // ---- exact step entry (_basic.py:139-156) of the coming attempt
        const double eps = ni_ulp_eps(t, dir);
        const double min_step = __dmul_rn(8.0, eps);
        h_abs = ni_lt_pos(h_abs, min_step) ? min_step : h_abs;
        const double h = ADV ? __hiloint2double(__double2hiint(h_abs) ^ (__double2hiint(dir) & (int)0x80000000), __double2loint(h_abs))
                             : h_abs;
        tn = __dadd_rn(t, h);
        const double dtb = __dadd_rn(tn, -tb);
        const bool zero = ((__double2hiint(dtb) << 1) | __double2loint(dtb)) == 0;
        const bool same = (__double2hiint(dtb) ^ __double2hiint(dir)) >= 0;
        if (zero || same) {  // direction * (t_new - t_bound) >= 0: this attempt ends on t_bound
          tn = tb;
          // (the clipped step t_bound - t has the sign of `direction`: the step is again +-h_abs)
          h_abs = fabs(__dadd_rn(tn, -t));
        }
        hs = ADV ? __hiloint2double(__double2hiint(h_abs) ^ (__double2hiint(dir) & (int)0x80000000), __double2loint(h_abs)) : h_abs;
      } else {
        // a lane without work keeps executing the arithmetic: give it values that raise no flag
        t = 0.0;
        h_abs = 0.0;
        hs = 0.0;
        tn = 0.0;
        tb = 0.0;
        kx = 0;
        tbk = 0x7fffffff;
        xs = 0;
      }
      if (__all_sync(FULL, !active)) break;
    }

    // ================================================================ HOT: one attempt for every lane
    // h = h_abs * direction (_advanced.py:163; a sign flip: direction = +-1) or h_abs (_basic.py:148); also after a
    // clip to t_bound, where h_abs = |t_bound - t|
    const double h = hs;
    double en, pw, kn[n], yn[n], auxn[Q > 0 ? Q : 1], num[n];
    bool ok;
    unsigned out;
    ni_rk_attempt<Rhs, METHOD, FAST, true>(a, t, tn, h, y, kl, p, yn, kn, auxn, num, en, pw, ok, out);
    if constexpr (FAST) {
      // (en holds the mean square error) the root without its range checks; outside [2^-100, 2^100) or NaN: cold block
      pw = ni_pow_fast_core<2 * (Tab::ORDER + 1)>(en);
      ok = (unsigned)__double2hiint(en) < 0x3ff00000u;
      out = (unsigned)(__double2hiint(en) - 0x39b00000) >= (0x46300000u - 0x39b00000u) ? 1u : 0u;
    }
    // ---- controller + commit (_basic.py:171-180)
    // SAFETY * en ** exp > 0 and never NaN here; both clamps unconditionally (an accepted attempt has factor > 0.9 >
    // MIN_FACTOR, a rejected one <= 0.9 < MAX_FACTOR), on the integer pipe: 10.0 has a zero low word, so
    // sp >= 10 <=> hi(sp) >= hi(10); sp < 0.2 is a 64-bit integer compare of non-negative doubles
    const double sp = __dmul_rn(NI_MC(0, 0.9), pw);
    // (both clamps stay in line: the first rejection after the ten-fold growth from a small initial step usually has
    // an error norm above 1845, i.e. MIN_FACTOR -- sending it to the cold block cost C3 3 %)
    double fac = (unsigned)__double2hiint(sp) >= 0x40240000u ? 10.0 : sp;
#if NI_EXP_ABSMAX_FP
    fac = fac < NI_MC(2, 0.2) ? 0.2 : fac;  // (one FP64 compare with ONE register source: cheaper than the two-word integer compare)
#else
    fac = ni_lt_pos(fac, 0.2) ? 0.2 : fac;
#endif
    double h_new = __dmul_rn(h_abs, fac);
    // Lanes that need the cold block commit there; the hot commit below runs on predicates that never leave the
    // predicate registers (a flag merged across the cold call is materialised as a byte: six more instructions).
    // Lanes without work may raise `rare` (their benign values give a zero error): they take the branch and skip the
    // call -- testing `active` up here cost two instructions in every pass, the detour only costs the tail of a run.
    const bool rare = (out != 0u) | ((unsigned)__double2hiint(h_new) < (unsigned)xs);
    bool cold = false, cold_acc = false;
    if (rare && active) {
      NiLeanColdIn<n> v;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        v.num[i] = num[i];
        v.y[i] = y[i];
        v.yn[i] = yn[i];
      }
      // (|h| is the h_abs this attempt started from: passing h keeps the old h_abs dead after the multiply above)
      const NiLeanCold c = ni_lean_cold<n, METHOD, FAST>(v, a.rtol, a.atol, t, dir, fabs(h), en, pw, ok ? 1 : 0, out);
      h_new = c.h_new;
      if (c.flags & 5) {  // accepted, or the failure exit: commit the attempt (the rejected t, y when failing)
        t = tn;
#pragma unroll
        for (int i = 0; i < n; ++i) y[i] = yn[i];
      }
#if NI_EXP_NACC
      nacc += (c.flags & 1) - (int)((unsigned)(__double2hiint(en) - 0x3ff00000) >> 31);  // (undo the count below)
#else
      nacc += c.flags & 1;
#endif
      if (c.flags & 8) {
        status = c.status;
        step_h = h;  // step_size of the exit
      }
      if (REC && (c.flags & 1)) {
        cold_acc = true;
        step_h = h;
      }
      cold = true;  // the coming step entry is evaluated exactly in the service branch (8 ulp floor)
    }
    const bool take = ok && !rare;
    t = take ? tn : t;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      y[i] = take ? yn[i] : y[i];
      kl[i] = kn[i];  // also after a rejection: the stale FSAL stage
    }
    h_abs = h_new;
#if NI_EXP_NACC
    // error_norm < 1 as the sign bit of hi(en) - hi(1.0), added with one LEA.HI (a predicated increment costs three
    // instructions once ptxas has turned it into a select); a lane that took the cold block corrected its count there
    nacc += (int)((unsigned)(__double2hiint(en) - 0x3ff00000) >> 31);
#else
    nacc += take ? 1 : 0;
#endif
    ++iter;
    if constexpr (REC) {
      // one slot per accepting lane, handed out by ballot rank: consecutive lanes write consecutive slots
      if (take) step_h = h;  // (a trajectory may leave RUNNING when the log fills up: its step_size must be current)
      const bool recme = take | cold_acc;
      const unsigned m = __ballot_sync(FULL, recme);
      if (recme) {
        const long long slot = rc_base + (rc_off + __popc(m & lanes_below)), c = a.rec_cap;
        a.rec_traj[slot] = (unsigned int)idx;
        a.rec_t[slot] = t;
#pragma unroll
        for (int i = 0; i < n; ++i) a.rec_y[i * c + slot] = y[i];
#pragma unroll
        for (int i = 0; i < Q; ++i) a.rec_aux[i * c + slot] = auxn[i];
      }
      rc_off += __popc(m);
    }
    // ---- step entry of the next attempt (_basic.py:145-156): h = +-h_abs, tn = t + h; the clip to t_bound is evaluated
    // exactly, but only by the lanes whose t + h may have reached t_bound (conservative key test: the last attempt of a
    // trajectory, 0.2 % of the lane-attempts); the 8 ulp floor of a step entry (:139-143) is the service branch's, which
    // the cold block requests.  A trajectory whose last step ended on t_bound asks for the service branch here.
    {
      hs = ADV ? __hiloint2double(__double2hiint(h_abs) ^ (__double2hiint(dir) & (int)0x80000000), __double2loint(h_abs)) : h_abs;
      tn = __dadd_rn(t, hs);
      svc = REC ? (cold | (rc_len - rc_off < 32)) : cold;
      if ((__double2hiint(tn) ^ kx) >= tbk) {
        if (__dmul_rn(dir, __dadd_rn(t, -tb)) >= 0.0) {  // (:135) t_bound has been reached: the trajectory leaves
          svc = true;
        } else {
          const double dtb = __dadd_rn(tn, -tb);
          const bool zero = ((__double2hiint(dtb) << 1) | __double2loint(dtb)) == 0;
          const bool same = (__double2hiint(dtb) ^ __double2hiint(dir)) >= 0;
          if (zero || same) {  // direction * (t_new - t_bound) >= 0: the attempt ends on t_bound
            tn = tb;
            // (the clipped step t_bound - t has the sign of `direction`: the next h is again +-h_abs)
            h_abs = fabs(__dadd_rn(tn, -t));
            hs = ADV ? __hiloint2double(__double2hiint(h_abs) ^ (__double2hiint(dir) & (int)0x80000000), __double2loint(h_abs)) : h_abs;
          }
        }
      }
    }
  }
  if constexpr (REC) {  // slots reserved but never written
    for (int o = rc_off + (int)lane; o < rc_len; o += 32) a.rec_traj[rc_base + o] = NI_REC_INVALID;
  }
}

template <class Rhs, int METHOD, int SEM, bool REC, bool FAST = false>
NI_DEV void ni_small_run_body(const NiArgs &a) {
  const bool no_max_step = __double_as_longlong(a.max_step) == 0x7ff0000000000000LL && a.max_step_v == nullptr;
  if constexpr (NI_LEAN && SEM != 2 && NI_COND == 0) {
    // fast-forward without an attempt budget and with max_step = inf (recording: chunked slot reservation, i.e. an
    // ensemble of at least 65 536 trajectories); lock-step ni.step and everything else take the general body
    if (no_max_step && a.atol_floor && a.max_attempts >= 0x7fffffffLL && a.max_accepts >= 0x7fffffffLL && (!REC || a.rec_chunk >= 32))
      ni_small_run_lean<Rhs, METHOD, SEM == 1, FAST, REC>(a);
    else
      ni_small_run_impl<Rhs, METHOD, SEM, REC, FAST, true>(a);
  } else {
    if (SEM != 2 && no_max_step)
      ni_small_run_impl<Rhs, METHOD, SEM, REC, FAST, false>(a);
    else
      ni_small_run_impl<Rhs, METHOD, SEM, REC, FAST, true>(a);
  }
}

// ---------------------------------------------------------------- kernel entry points
template <class Rhs, int METHOD>
__global__ void __launch_bounds__(NI_BLOCK) ni_k_init(const NiArgs a) {
  ni_small_init_body<Rhs, METHOD>(a);
}
template <class Rhs, int METHOD, int SEM, bool REC, bool FAST>
__global__ void __launch_bounds__(NI_BLOCK, NI_MIN_BLOCKS(Rhs::N)) ni_k_run(const NiArgs a) {
  ni_small_run_body<Rhs, METHOD, SEM, REC, FAST>(a);
}
#endif  // NI_CORE_CUH
