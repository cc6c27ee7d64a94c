/*
 * ni_oracle.c -- CPU restatement of the reference's adaptive RK23/RK45 stepper.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
 * legs).  Nothing under numba_integrators_b200/ may import, link or execute it.
 *
 * Parity pin: every function here is checked bit-for-bit (n <= 3) or to
 * 1e-12 (n >= 4) against the live reference (numba 0.65 / OpenBLAS SkylakeX /
 * glibc 2.39) by oracle/gen_golden.py -> tests/golden/ (npz files) and
 * tests/test_oracle_golden.py.
 *
 * Reference lines restated (all under /root/reference/src/numba_integrators/):
 *   _aux.py:16-19     SAFETY / MIN_FACTOR / MAX_FACTOR
 *   _aux.py:71-75     norm (RMS)
 *   _aux.py:77-115    RK23 / RK45 tableaux
 *   _basic.py:33-89   select_initial_step      (_advanced.py:39-96 twin)
 *   _basic.py:114-180 _step                    (_advanced.py:126-195 twin)
 *   _basic.py:203-242 RK.__init__              (_advanced.py:200-247 twin)
 *   _API.py:22-49     step / ff2t / ff2cond    (drivers, restated in oracle.py)
 *
 * Arithmetic order: the reference evaluates tableau
 * contractions with OpenBLAS dgemv (FMA chains in a fixed order) and all
 * element-wise expressions unfused.  Build with -ffp-contract=off -mfma so the
 * only fused operations are the explicit fma() calls below.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define NI_SAFETY 0.9     /* _aux.py:16 */
#define NI_MIN_FACTOR 0.2 /* _aux.py:18 */
#define NI_MAX_FACTOR 10.0 /* _aux.py:19 (int 10 promoted to float64) */

enum { NI_RK23 = 0, NI_RK45 = 1 };
enum {
  NI_RHS_HARMONIC = 0,    /* y0' = y1, y1' = -y0   (reference.py:46 "sine") */
  NI_RHS_LORENZ63 = 1,    /* p = (sigma, rho, beta), aux = x^2+y^2+z^2 */
  NI_RHS_VANDERPOL = 2,   /* p = (mu) */
  NI_RHS_HEAT1D = 3,      /* p = (k), Dirichlet-0 3-point Laplacian */
  NI_RHS_BURGERS1D = 4,   /* p = (nu/dx^2, 1/(2dx)) */
  NI_RHS_EXPONENTIAL = 5, /* y' = y                (reference.py:54) */
  NI_RHS_RICCATI = 6,     /* y' = 2y/x - x^2 y^2   (reference.py:38) */
  NI_RHS_README_ADV = 7,  /* README.md:80-85: p=(a,b0,b1): aux=a*y1; dy=(aux+b0, -y0+b1) */
  NI_RHS_BLOWUP = 8,      /* y' = y^2 (failure-exit fixture) */
  NI_RHS_HEAT1D_AUX = 9,  /* heat1d with aux = (sum_j y_j in index order, y[n/2]): oracle/gen_golden.py f_heat_aux */
  NI_RHS_COUNT
};
enum { NI_ST_RUNNING = 0, NI_ST_FINISHED = 1, NI_ST_FAILED = 2, NI_ST_NONFINITE = 3 };

/* ---- tableaux, written as the same rational expressions as _aux.py:77-115 ---- */
static const double RK23_A[3][3] = {{0, 0, 0}, {1.0 / 2, 0, 0}, {0, 3.0 / 4, 0}};
static const double RK23_B[3] = {2.0 / 9, 1.0 / 3, 4.0 / 9};
static const double RK23_C[3] = {0, 1.0 / 2, 3.0 / 4};
static const double RK23_E[4] = {5.0 / 72, -1.0 / 12, -1.0 / 9, 1.0 / 8};

static const double RK45_A[6][5] = {
    {0., 0., 0., 0., 0.},
    {1.0 / 5, 0., 0., 0., 0.},
    {3.0 / 40, 9.0 / 40, 0., 0., 0.},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0., 0.},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK45_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK45_C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
static const double RK45_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

typedef struct {
  int rhs_id, method, advanced, n, P, Q, S; /* S = n_stages */
  double t, t_bound, direction, max_step, error_exponent, h_abs, step_size;
  double *y, *K, *rtol, *atol, *params, *aux; /* K is (S+1) x n, C order */
  double *y_old, *y_tmp, *dot, *aux_tmp;      /* scratch */
  int64_t n_acc, n_rej, nfev;
  int status;
} ni_oracle_solver;

/* ------------------------------------------------------------------ RHS ---- */
static void rhs_eval(const ni_oracle_solver *s, double t, const double *y, double *dy, double *aux) {
  const double *p = s->params;
  const int n = s->n;
  switch (s->rhs_id) {
  case NI_RHS_HARMONIC:
    dy[0] = y[1];
    dy[1] = -y[0];
    break;
  case NI_RHS_LORENZ63: {
    const double x = y[0], yy = y[1], z = y[2];
    dy[0] = p[0] * (yy - x);
    dy[1] = x * (p[1] - z) - yy;
    dy[2] = x * yy - p[2] * z;
    if (aux && s->Q > 0) aux[0] = (x * x + yy * yy) + z * z;
    break;
  }
  case NI_RHS_VANDERPOL:
    dy[0] = y[1];
    dy[1] = (p[0] * (1.0 - y[0] * y[0])) * y[1] - y[0];
    break;
  case NI_RHS_HEAT1D:
  case NI_RHS_HEAT1D_AUX: {
    double tot = 0.0;
    for (int j = 0; j < n; ++j) {
      const double ym = j > 0 ? y[j - 1] : 0.0, yp = j + 1 < n ? y[j + 1] : 0.0;
      dy[j] = p[0] * ((ym - 2.0 * y[j]) + yp);
      tot += y[j];
    }
    if (aux && s->Q > 1) {
      aux[0] = tot;
      aux[1] = y[n / 2];
    }
    break;
  }
  case NI_RHS_BURGERS1D:
    for (int j = 0; j < n; ++j) {
      const double ym = j > 0 ? y[j - 1] : 0.0, yp = j + 1 < n ? y[j + 1] : 0.0;
      dy[j] = p[0] * ((ym - 2.0 * y[j]) + yp) - (y[j] * p[1]) * (yp - ym);
    }
    break;
  case NI_RHS_EXPONENTIAL:
    for (int j = 0; j < n; ++j) dy[j] = y[j];
    break;
  case NI_RHS_RICCATI:
    for (int j = 0; j < n; ++j) dy[j] = (2.0 * y[j]) / t - (t * t) * (y[j] * y[j]);
    break;
  case NI_RHS_README_ADV: {
    const double a = p[0] * y[1];
    dy[0] = a + p[1];
    dy[1] = -y[0] + p[2];
    if (aux && s->Q > 0) aux[0] = a;
    break;
  }
  case NI_RHS_BLOWUP:
    for (int j = 0; j < n; ++j) dy[j] = y[j] * y[j];
    break;
  default:
    for (int j = 0; j < n; ++j) dy[j] = NAN;
  }
}

/* _aux.py:71-75: sqrt(sum(x*x)/x.size), sequential and unfused (appendix B) */
static double rms_norm(const double *x, int n) {
  double acc = 0.0;
  for (int i = 0; i < n; ++i) acc += x[i] * x[i];
  return sqrt(acc / (double)n);
}

/* out[i] = sum_{j<s} K[j][i] * v[j] in the order OpenBLAS dgemv_n produces for a
 * (n x s) matrix with n <= 3 rows (appendix B).  n >= 4 uses the same order
 * (tolerance parity only there: the reference's vector micro-kernel order is
 * host-CPU dependent). */
static void dot_cols(const double *K, int n, int s, const double *v, double *out) {
  for (int i = 0; i < n; ++i) {
    double acc;
    if (s <= 3) {
      acc = 0.0;
      for (int j = 0; j < s; ++j) acc = fma(K[j * n + i], v[j], acc);
    } else {
      if (n == 1) {
        acc = K[1 * n + i] * v[1];
        acc = fma(K[0 * n + i], v[0], acc);
        acc = fma(K[2 * n + i], v[2], acc);
        acc = fma(K[3 * n + i], v[3], acc);
      } else {
        const double a = fma(K[0 * n + i], v[0], K[1 * n + i] * v[1]);
        const double b = fma(K[2 * n + i], v[2], K[3 * n + i] * v[3]);
        acc = a + b;
      }
      for (int j = 4; j < s; ++j) acc = fma(K[j * n + i], v[j], acc);
    }
    out[i] = acc;
  }
}

/* _basic.py:33-89 / _advanced.py:39-96 (compiled fastmath=True in the reference:
 * scale and y1 are contracted to FMAs, the rest is literal -- appendix B) */
static double select_initial_step(ni_oracle_solver *s, const double *f0) {
  const int n = s->n;
  double *scale = s->dot, *tmp = s->y_tmp, *f1 = s->y_old;
  for (int i = 0; i < n; ++i) scale[i] = fma(fabs(s->y[i]), s->rtol[i], s->atol[i]);
  for (int i = 0; i < n; ++i) tmp[i] = s->y[i] / scale[i];
  const double d0 = rms_norm(tmp, n);
  for (int i = 0; i < n; ++i) tmp[i] = f0[i] / scale[i];
  const double d1 = rms_norm(tmp, n);
  const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
  const double hd = h0 * s->direction;
  for (int i = 0; i < n; ++i) tmp[i] = fma(hd, f0[i], s->y[i]);
  rhs_eval(s, s->t + hd, tmp, f1, NULL);
  s->nfev += 1;
  for (int i = 0; i < n; ++i) tmp[i] = (f1[i] - f0[i]) / scale[i];
  const double d2 = rms_norm(tmp, n) / h0;
  double h1;
  if (d1 <= 1e-15 && d2 <= 1e-15)
    h1 = fmax(1e-6, h0 * 1e-3);
  else
    h1 = pow(fmax(d1, d2) * 100.0, s->error_exponent);
  return fmin(100.0 * h0, h1);
}

/* ------------------------------------------------------------ lifecycle ---- */
void ni_oracle_destroy(ni_oracle_solver *s) {
  if (!s) return;
  free(s->y);
  free(s);
}

/* RK.__init__ (_basic.py:203-242) / _RK_Advanced.__init__ (_advanced.py:200-247) */
ni_oracle_solver *ni_oracle_create(int rhs_id, int method, int advanced, int n, int P, int Q, double t0,
                                   const double *y0, const double *params, double t_bound, double max_step,
                                   const double *rtol, const double *atol, double first_step) {
  if (n < 1 || rhs_id < 0 || rhs_id >= NI_RHS_COUNT || (method != NI_RK23 && method != NI_RK45)) return NULL;
  ni_oracle_solver *s = (ni_oracle_solver *)calloc(1, sizeof(*s));
  if (!s) return NULL;
  s->rhs_id = rhs_id;
  s->method = method;
  s->advanced = advanced;
  s->n = n;
  s->P = P;
  s->Q = Q;
  s->S = method == NI_RK45 ? 6 : 3;
  const int S = s->S;
  const size_t total = (size_t)n * (1 + (S + 1) + 2 + 3) + (size_t)P + 2 * (size_t)Q + 8;
  double *buf = (double *)calloc(total, sizeof(double));
  if (!buf) {
    free(s);
    return NULL;
  }
  s->y = buf;
  s->K = s->y + n;
  s->rtol = s->K + (size_t)(S + 1) * n;
  s->atol = s->rtol + n;
  s->y_old = s->atol + n;
  s->y_tmp = s->y_old + n;
  s->dot = s->y_tmp + n;
  s->params = s->dot + n;
  s->aux = s->params + P;
  s->aux_tmp = s->aux + Q;
  memcpy(s->y, y0, sizeof(double) * n);
  memcpy(s->rtol, rtol, sizeof(double) * n);
  memcpy(s->atol, atol, sizeof(double) * n);
  if (P > 0 && params) memcpy(s->params, params, sizeof(double) * P);
  s->t = t0;
  s->t_bound = t_bound;
  s->max_step = max_step;
  rhs_eval(s, s->t, s->y, s->K + (size_t)S * n, s->aux); /* K[-1] = fun(t0, y0)  :232 */
  s->nfev = 1;
  s->direction = t_bound != t0 ? (t_bound > t0 ? 1.0 : -1.0) : 1.0; /* :233 */
  s->error_exponent = -1.0 / ((method == NI_RK45 ? 4 : 2) + 1);     /* :234 */
  if (first_step == 0.0)
    s->h_abs = select_initial_step(s, s->K + (size_t)S * n); /* :236-239 */
  else
    s->h_abs = fabs(first_step); /* :241 */
  s->step_size = s->direction * s->h_abs; /* :242 */
  s->status = NI_ST_RUNNING;
  return s;
}

/* _step (_basic.py:114-180) / step_advanced (_advanced.py:126-195).
 * Returns 1 if a step was accepted ("running"), 0 otherwise. */
int ni_oracle_step(ni_oracle_solver *s) {
  const int n = s->n, S = s->S;
  const double dir = s->direction;
  const double *A = s->method == NI_RK45 ? &RK45_A[0][0] : &RK23_A[0][0];
  const int lda = s->method == NI_RK45 ? 5 : 3;
  const double *B = s->method == NI_RK45 ? RK45_B : RK23_B;
  const double *C = s->method == NI_RK45 ? RK45_C : RK23_C;
  const double *E = s->method == NI_RK45 ? RK45_E : RK23_E;
  double *K = s->K;

  if (s->status == NI_ST_FAILED || s->status == NI_ST_NONFINITE) return 0; /* latched (documented deviation) */
  if (dir * (s->t - s->t_bound) >= 0) { /* :135-136 */
    s->step_size = s->advanced ? s->h_abs : dir * s->h_abs; /* _advanced.py:151 vs _basic.py:136 */
    s->status = NI_ST_FINISHED;
    return 0;
  }
  s->status = NI_ST_RUNNING;
  const double t_old = s->t;
  memcpy(s->y_old, s->y, sizeof(double) * n);
  const double *y_old = s->y_old;
  const double eps = fabs(nextafter(t_old, dir * INFINITY) - t_old); /* :139 */
  const double min_step = 8 * eps;                                   /* :140 */
  double h_abs = s->h_abs;
  if (h_abs < min_step) h_abs = min_step; /* :142-143 */

  for (;;) {
    if (h_abs > s->max_step) h_abs = s->max_step - eps; /* :146-147 */
    double h = s->advanced ? h_abs * dir : h_abs;       /* _advanced.py:163 vs _basic.py:148 */
    double t = t_old + h;
    memcpy(K, K + (size_t)S * n, sizeof(double) * n); /* K[0] = K[-1]  :151 (stale after a reject) */
    if (dir * (t - s->t_bound) >= 0) {                /* :153-156 */
      t = s->t_bound;
      h = t - t_old;
      h_abs = fabs(h);
    }
    for (int st = 1; st < S; ++st) { /* :158-160 */
      dot_cols(K, n, st, A + st * lda, s->dot);
      for (int i = 0; i < n; ++i) s->y_tmp[i] = y_old[i] + s->dot[i] * h;
      rhs_eval(s, t_old + C[st] * h, s->y_tmp, K + (size_t)st * n, NULL);
    }
    dot_cols(K, n, S, B, s->dot); /* :162 */
    for (int i = 0; i < n; ++i) s->y_tmp[i] = y_old[i] + h * s->dot[i];
    rhs_eval(s, t, s->y_tmp, K + (size_t)S * n, s->aux_tmp); /* :164 (FSAL evaluation) */
    s->nfev += S;
    dot_cols(K, n, S + 1, E, s->dot); /* :166-169 */
    for (int i = 0; i < n; ++i)
      s->dot[i] = s->dot[i] * h / (s->atol[i] + fmax(fabs(y_old[i]), fabs(s->y_tmp[i])) * s->rtol[i]);
    const double error_norm = rms_norm(s->dot, n);

    if (error_norm < 1) { /* :171-175 */
      h_abs *= (error_norm == 0) ? NI_MAX_FACTOR : fmin(NI_MAX_FACTOR, NI_SAFETY * pow(error_norm, s->error_exponent));
      s->t = t;
      memcpy(s->y, s->y_tmp, sizeof(double) * n);
      if (s->Q > 0) memcpy(s->aux, s->aux_tmp, sizeof(double) * s->Q);
      s->h_abs = h_abs;
      s->step_size = h;
      s->n_acc += 1;
      return 1;
    }
    if (!(error_norm >= 1)) { /* NaN: the reference would spin forever; latch instead */
      s->n_rej += 1;
      s->status = NI_ST_NONFINITE;
      s->h_abs = h_abs;
      s->step_size = h;
      return 0;
    }
    h_abs *= fmax(NI_MIN_FACTOR, NI_SAFETY * pow(error_norm, s->error_exponent)); /* :177-178 */
    s->n_rej += 1;
    if (h_abs < min_step) { /* :179-180: commits the REJECTED t, y */
      s->t = t;
      memcpy(s->y, s->y_tmp, sizeof(double) * n);
      if (s->Q > 0) memcpy(s->aux, s->aux_tmp, sizeof(double) * s->Q);
      s->h_abs = h_abs;
      s->step_size = h;
      s->status = NI_ST_FAILED;
      return 0;
    }
  }
}

/* ---- accessors for the ctypes wrapper (oracle.py) ---- */
double ni_oracle_get_t(const ni_oracle_solver *s) { return s->t; }
double ni_oracle_get_t_bound(const ni_oracle_solver *s) { return s->t_bound; }
void ni_oracle_set_t_bound(ni_oracle_solver *s, double tb) { s->t_bound = tb; }
double ni_oracle_get_h_abs(const ni_oracle_solver *s) { return s->h_abs; }
void ni_oracle_set_h_abs(ni_oracle_solver *s, double h) { s->h_abs = h; }
double ni_oracle_get_step_size(const ni_oracle_solver *s) { return s->step_size; }
double ni_oracle_get_direction(const ni_oracle_solver *s) { return s->direction; }
int ni_oracle_get_status(const ni_oracle_solver *s) { return s->status; }
int64_t ni_oracle_get_n_acc(const ni_oracle_solver *s) { return s->n_acc; }
int64_t ni_oracle_get_n_rej(const ni_oracle_solver *s) { return s->n_rej; }
int64_t ni_oracle_get_nfev(const ni_oracle_solver *s) { return s->nfev; }
void ni_oracle_get_y(const ni_oracle_solver *s, double *out) { memcpy(out, s->y, sizeof(double) * s->n); }
void ni_oracle_get_aux(const ni_oracle_solver *s, double *out) {
  if (s->Q > 0) memcpy(out, s->aux, sizeof(double) * s->Q);
}
void ni_oracle_get_K(const ni_oracle_solver *s, double *out) {
  memcpy(out, s->K, sizeof(double) * (size_t)(s->S + 1) * s->n);
}

/* Tableau export so tests can compare the constants with _aux.py:77-115 */
void ni_oracle_tableau(int method, double *A, double *B, double *C, double *E) {
  if (method == NI_RK45) {
    memcpy(A, RK45_A, sizeof(RK45_A));
    memcpy(B, RK45_B, sizeof(RK45_B));
    memcpy(C, RK45_C, sizeof(RK45_C));
    memcpy(E, RK45_E, sizeof(RK45_E));
  } else {
    memcpy(A, RK23_A, sizeof(RK23_A));
    memcpy(B, RK23_B, sizeof(RK23_B));
    memcpy(C, RK23_C, sizeof(RK23_C));
    memcpy(E, RK23_E, sizeof(RK23_E));
  }
}

/* ---- ensemble driver: "the reference's solver looped per trajectory and
 * parallelised across the host cores" (BASELINE.json north_star).  Each
 * trajectory is `while step(): pass` (optionally recording accepted steps).
 * y0 is N x n row-major, params N x P; t0/t_bound have N entries if
 * per_traj_time != 0, else one.  rec_* (optional) receive up to rec_cap
 * accepted steps PER TRAJECTORY: rec_t[N][rec_cap], rec_y[N][rec_cap][n],
 * rec_aux[N][rec_cap][Q]; n_acc tells how many were written. */
typedef struct {
  int rhs_id, method, advanced, n, P, Q;
  int64_t N;
  const double *t0, *y0, *params, *t_bound;
  int per_traj_time;
  double max_step;
  const double *rtol, *atol;
  double first_step;
  double *t_out, *y_out, *aux_out, *h_abs_out;
  int64_t *n_acc, *n_rej;
  int32_t *status;
  int64_t rec_cap;
  double *rec_t, *rec_y, *rec_aux;
  int64_t next; /* shared work cursor (chunks of 64 trajectories) */
  int err;
} ens_job;

static void *ens_worker(void *arg) {
  ens_job *j = (ens_job *)arg;
  const int n = j->n, P = j->P, Q = j->Q;
  for (;;) {
    const int64_t lo = __atomic_fetch_add(&j->next, 64, __ATOMIC_RELAXED);
    if (lo >= j->N) break;
    const int64_t hi = lo + 64 < j->N ? lo + 64 : j->N;
    for (int64_t i = lo; i < hi; ++i) {
      const double ti0 = j->per_traj_time ? j->t0[i] : j->t0[0];
      const double tb = j->per_traj_time ? j->t_bound[i] : j->t_bound[0];
      ni_oracle_solver *s =
          ni_oracle_create(j->rhs_id, j->method, j->advanced, n, P, Q, ti0, j->y0 + i * n,
                           P > 0 ? j->params + i * P : NULL, tb, j->max_step, j->rtol, j->atol, j->first_step);
      if (!s) {
        __atomic_store_n(&j->err, 1, __ATOMIC_RELAXED);
        continue;
      }
      int64_t k = 0;
      while (ni_oracle_step(s)) {
        if (j->rec_cap > 0 && k < j->rec_cap) {
          if (j->rec_t) j->rec_t[i * j->rec_cap + k] = s->t;
          if (j->rec_y) memcpy(j->rec_y + (i * j->rec_cap + k) * n, s->y, sizeof(double) * n);
          if (j->rec_aux && Q > 0) memcpy(j->rec_aux + (i * j->rec_cap + k) * Q, s->aux, sizeof(double) * Q);
        }
        ++k;
      }
      if (j->t_out) j->t_out[i] = s->t;
      if (j->y_out) memcpy(j->y_out + i * n, s->y, sizeof(double) * n);
      if (j->aux_out && Q > 0) memcpy(j->aux_out + i * Q, s->aux, sizeof(double) * Q);
      if (j->h_abs_out) j->h_abs_out[i] = s->h_abs;
      if (j->n_acc) j->n_acc[i] = s->n_acc;
      if (j->n_rej) j->n_rej[i] = s->n_rej;
      if (j->status) j->status[i] = s->status;
      ni_oracle_destroy(s);
    }
  }
  return NULL;
}

int ni_oracle_max_threads(void) {
  long c = sysconf(_SC_NPROCESSORS_ONLN);
  return c > 0 ? (int)c : 1;
}

int ni_oracle_run_ensemble(int rhs_id, int method, int advanced, int n, int P, int Q, int64_t N,
                           const double *t0, const double *y0, const double *params, const double *t_bound,
                           int per_traj_time, double max_step, const double *rtol, const double *atol,
                           double first_step, int nthreads, double *t_out, double *y_out, double *aux_out,
                           double *h_abs_out, int64_t *n_acc, int64_t *n_rej, int32_t *status, int64_t rec_cap,
                           double *rec_t, double *rec_y, double *rec_aux) {
  ens_job j = {rhs_id, method, advanced, n, P, Q, N, t0, y0, params, t_bound, per_traj_time, max_step, rtol, atol,
               first_step, t_out, y_out, aux_out, h_abs_out, n_acc, n_rej, status, rec_cap, rec_t, rec_y, rec_aux,
               0, 0};
  if (nthreads <= 0) nthreads = ni_oracle_max_threads();
  if (nthreads > 1024) nthreads = 1024;
  if (nthreads == 1) {
    ens_worker(&j);
    return j.err;
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  int started = 0;
  for (int k = 0; k < nthreads; ++k)
    if (pthread_create(&th[started], NULL, ens_worker, &j) == 0) ++started;
  if (started == 0) ens_worker(&j);
  for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
  free(th);
  return j.err;
}
