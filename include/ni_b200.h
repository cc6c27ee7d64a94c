/*
 * ni_b200.h -- C ABI of the B200-native ODE-ensemble engine (libni_b200.so).
 *
 * The reference (Limespy/numba-integrators) has no FFI: its boundary is the Python call
 * surface `ni.RK45 / ni.RK23 / ni.Advanced / ni.step / ni.ff2t / ni.ff2cond`
 * (src/numba_integrators/__init__.py:3, _API.py:22-49).  The entry points below are what
 * that surface binds to when the jitclass solver object (_basic.py:182-271,
 * _advanced.py:197-278) is replaced by a device-resident ensemble of N trajectories; the
 * Python facade in numba_integrators_b200/ calls them through ctypes (INTEGRATION.md shows
 * the stub).  Plain pointers and sizes only -- no torch / numpy types.
 *
 * Conventions
 *   - every function returns 0 (NI_OK) or a negative NI_ERR_* code; ni_last_error() gives the
 *     message of the calling thread's last failure.  No exceptions cross the boundary.
 *   - host buffers are caller-owned and only read/written during the call (copies go through
 *     the ensemble's CUDA stream; pass pinned memory for asynchronous copies).
 *   - device memory is owned by the handle.  One handle = one ensemble on one GPU; calls on
 *     one handle must be serialised by the caller, different handles are independent.
 *   - row-major host layouts: y0 / y / klast are N x n, params N x P, aux N x Q.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     NI_ERR_CUDA.
 */
#ifndef NI_B200_H
#define NI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NI_OK 0
#define NI_ERR_INVALID (-1)     /* bad argument */
#define NI_ERR_CUDA (-2)        /* CUDA runtime error (no device, launch failure, ...) */
#define NI_ERR_NVRTC (-3)       /* run-time compilation of a user RHS failed (see ni_module_log) */
#define NI_ERR_UNSUPPORTED (-4) /* no kernel for this (rhs, n, method) combination */
#define NI_ERR_NOMEM (-5)
#define NI_ERR_REC_OVERFLOW (-6) /* record log full: drain it and call run/step again */

/* methods: reference `Solvers.RK23 / Solvers.RK45` (_basic.py:335-339), tableaux _aux.py:77-115 */
#define NI_RK23 0
#define NI_RK45 1

/* built-in right-hand sides (p = per-trajectory parameters, aux = auxiliary output) */
#define NI_RHS_HARMONIC 0    /* y0'=y1, y1'=-y0                    (reference.py:46-49)  */
#define NI_RHS_LORENZ63 1    /* p=(sigma,rho,beta), aux=x^2+y^2+z^2                      */
#define NI_RHS_VANDERPOL 2   /* p=(mu)                                                    */
#define NI_RHS_HEAT1D 3      /* p=(k): dy_j = k (y_{j-1} - 2 y_j + y_{j+1}), Dirichlet 0 */
#define NI_RHS_BURGERS1D 4   /* p=(nu/dx^2, 1/(2dx))                                      */
#define NI_RHS_EXPONENTIAL 5 /* y' = y                             (reference.py:54-57)  */
#define NI_RHS_RICCATI 6     /* y' = 2y/x - x^2 y^2                (reference.py:38-41)  */
#define NI_RHS_README_ADV 7  /* README.md:80-85, p=(a,b0,b1), aux=a*y1                    */
#define NI_RHS_BLOWUP 8      /* y' = y^2 (step-too-small exit)                            */

/* per-trajectory status codes (the reference only has step() -> False, _basic.py:135,180) */
#define NI_ST_RUNNING 0
#define NI_ST_FINISHED 1  /* a step() call found direction*(t - t_bound) >= 0 */
#define NI_ST_FAILED 2    /* h_abs fell below 8 ulp(t) (_basic.py:179-180); latched */
#define NI_ST_NONFINITE 3 /* error norm NaN (the reference would loop forever); latched */

/* fields for ni_ensemble_get / ni_ensemble_set / ni_ensemble_device_ptr.
 * Names follow the jitclass attributes (_basic.py:182-199). */
#define NI_F_T 0          /* double[N]                                            */
#define NI_F_Y 1          /* double[N][n]  (device: [n][N] component-major)       */
#define NI_F_AUX 2        /* double[N][Q]  `auxiliary` (_advanced.py:276-278)     */
#define NI_F_H_ABS 3      /* double[N]                                            */
#define NI_F_STEP_SIZE 4  /* double[N]                                            */
#define NI_F_KLAST 5      /* double[N][n]  K[-1], the FSAL stage                  */
#define NI_F_STATUS 6     /* int32[N]                                             */
#define NI_F_N_ACCEPTED 7 /* int32[N]                                             */
#define NI_F_N_REJECTED 8 /* int32[N]                                             */
#define NI_F_RUNNING 9    /* uint8[N]  result of the last step() per trajectory   */
#define NI_F_T_BOUND 10   /* double[N]  (settable: _API.py:33,37)                 */
#define NI_F_DIRECTION 11 /* double[N]                                            */
#define NI_F_PARAMS 12    /* double[N][P]                                         */
#define NI_F_COND_HIT 13  /* uint8[N]  ff2cond predicate hit                      */
/* record log buffers (ni_ensemble_device_ptr only; capacity entries, see the recording section) */
#define NI_F_REC_TRAJ 20  /* uint32[cap]                                          */
#define NI_F_REC_STEP 21  /* (not stored any more: NI_ERR_UNSUPPORTED; ni_ensemble_record_drain numbers the steps) */
#define NI_F_REC_T 22     /* double[cap]                                          */
#define NI_F_REC_Y 23     /* double [n][cap] (layout 0) or [cap][n] (layout 1)    */
#define NI_F_REC_AUX 24   /* double [Q][cap] (layout 0) or [cap][Q] (layout 1)    */

/* predicates for ni_ensemble_ff2cond (checked after every accepted step, _API.py:41-49) */
#define NI_COND_Y_GT 1
#define NI_COND_Y_LT 2
#define NI_COND_AUX_GT 3
#define NI_COND_AUX_LT 4

typedef struct ni_module ni_module;     /* one compiled (rhs, n, P, Q, method, advanced) kernel set */
typedef struct ni_ensemble ni_ensemble; /* N trajectories resident on one GPU */

int ni_version(void);
const char *ni_last_error(void);
int ni_device_count(int *count);

/* ---- modules -------------------------------------------------------------------------
 * Replaces numba's JIT specialisation of the solver on the user's RHS
 * (`nbODEtype`, _aux.py:64-65; `Advanced(parameters_signature, auxiliary_signature, solver)`,
 * _advanced.py:280-368).  `mode` is a bit mask: NI_MODE_ADVANCED selects step_advanced semantics
 * (h = h_abs*direction, _advanced.py:163) instead of _step's (h = h_abs, _basic.py:148). */
#define NI_MODE_ADVANCED 1 /* step_advanced semantics (h = h_abs*direction), parameters + auxiliary */
#define NI_MODE_FAST 2     /* opt-in fast arithmetic: contracted multiply-adds, approximate reciprocal and
                              controller root (agrees with the reference to tolerance, not bit for bit) */
#define NI_MODE_SCIPY 4    /* opt-in scipy.integrate.RK23/RK45 step semantics instead of the reference's: fresh FSAL
                              stage on retries, no growth after a rejection, 10 ulp minimum step tested before each
                              attempt, h = t_new - t, scipy's select_initial_step (SURVEY.md A.2 lists the differences) */
int ni_module_builtin(int rhs_id, int n, int method, int mode, ni_module **out);
/* `rhs_body` is the body of
 *   __device__ void rhs(double t, const double* y, const double* p, double* dy, double* aux)
 * in CUDA C++; it is fused into the stepper kernels with NVRTC for sm_100a. */
int ni_module_compile(const char *rhs_body, int n, int P, int Q, int method, int mode, const char *nvrtc_options,
                      ni_module **out);
/* Large systems (n <= 4096) whose RHS is a 3-point stencil: `stencil_body` is the body of
 *   __device__ double rhs_j(double t, double ym, double yc, double yp, const double* p, int j, int n)
 * returning dy_j from y_{j-1}, y_j, y_{j+1} (zero outside 0..n-1); fused into the CTA-per-trajectory kernels. */
int ni_module_compile_stencil(const char *stencil_body, int n, int P, int method, int mode,
                              const char *nvrtc_options, ni_module **out);
/* Large systems (n <= 4096) with arbitrary (dense) coupling, written component-wise: `component_body` is the body of
 *   __device__ double rhs_j(double t, const double* y, int j, int n, const double* p)
 * returning dy_j; y is the whole stage vector, staged in shared memory by the CTA that owns the trajectory. */
int ni_module_compile_component(const char *component_body, int n, int P, int method, int mode,
                                const char *nvrtc_options, ni_module **out);
/* General form of the two calls above: component = 0 (3-point stencil) / 1 (component-wise), plus Q auxiliary outputs
 * (the second return value of an Advanced right-hand side, _advanced.py:22-26 / :181) computed from the WHOLE committed
 * state by `aux_body`, the body of
 *   __device__ void aux(double t, const double* y, int n, const double* p, double* aux)
 * (evaluated by one thread when a trajectory leaves the kernel, per recorded step, and before a predicate). */
int ni_module_compile_large(int component, const char *rhs_body, const char *aux_body, int n, int P, int Q, int method,
                            int mode, const char *nvrtc_options, ni_module **out);
int ni_module_info(const ni_module *m, int *n, int *P, int *Q, int *method, int *mode, int *n_stages);
/* device layout of the y / klast / aux / params buffers behind ni_ensemble_device_ptr:
 * 0 = component-major [n][N] (thread-per-trajectory kernels), 1 = trajectory-major [N][n] (CTA-per-trajectory) */
int ni_module_layout(const ni_module *m, int *layout);
const char *ni_module_log(const ni_module *m);
int ni_module_destroy(ni_module *m);

/* ---- ensembles -----------------------------------------------------------------------
 * ni_ensemble_create + ni_ensemble_init == the solver constructor RK.__init__
 * (_basic.py:203-242): FSAL seed K[-1] = f(t0, y0), direction, select_initial_step
 * (_basic.py:33-89) unless first_step != 0.  t0 / t_bound have N entries when the matching
 * `*_per_traj` flag is non-zero, else 1.  rtol / atol have n entries (_aux.py:134-140). */
int ni_ensemble_create(const ni_module *m, int device, int64_t N, ni_ensemble **out);
int ni_ensemble_destroy(ni_ensemble *e);
int ni_ensemble_set_stream(ni_ensemble *e, void *cuda_stream); /* run on the caller's stream */
int ni_ensemble_upload(ni_ensemble *e, const double *t0, int t0_per_traj, const double *y0, const double *params,
                       const double *t_bound, int t_bound_per_traj);    /* host -> HBM staging (async) */
int ni_ensemble_reset(ni_ensemble *e, const double *rtol, const double *atol, double max_step,
                      double first_step);                               /* init kernel on resident inputs (async) */
int ni_ensemble_init(ni_ensemble *e, const double *t0, int t0_per_traj, const double *y0, const double *params,
                     const double *t_bound, int t_bound_per_traj, const double *rtol, const double *atol,
                     double max_step, double first_step);               /* upload + reset + sync */

/* Per-trajectory max_step / first_step: the reference takes both per solver object (_basic.py:292-299, 323-330), so an
 * ensemble of N objects may carry N different values.  Host arrays of N doubles (copied asynchronously: keep them valid
 * until the next synchronising call) that override the scalars of ni_ensemble_reset / ni_ensemble_init from the next
 * reset on; NULL restores the scalar.  first_step entries equal to 0 select the initial step automatically. */
int ni_ensemble_set_step_limits(ni_ensemble *e, const double *max_step, const double *first_step);

/* ni.step(solver) on every trajectory (_API.py:22-24): at most one accepted step each.
 * *n_true = number of trajectories whose step() returned True. */
int ni_ensemble_step(ni_ensemble *e, int64_t *n_true);
/* `while ni.step(solver): pass` on every trajectory.  *n_unfinished = trajectories still
 * RUNNING when the call returns (0 unless the record log filled up -> NI_ERR_REC_OVERFLOW).  Trajectories that ended
 * FAILED / NONFINITE are not "unfinished": ni_ensemble_poll counts them, NI_F_STATUS names them. */
int ni_ensemble_run(ni_ensemble *e, int64_t *n_unfinished);
int ni_ensemble_run_async(ni_ensemble *e); /* the same launch(es), no host synchronisation: ni_ensemble_poll tells how it ended */
int ni_ensemble_sync(ni_ensemble *e);
/* Synchronises and reports how the most recent run / run_async / step ended: *n_unfinished = trajectories still RUNNING
 * (attempt budget used up, record log full), *n_failed = trajectories that left with status FAILED (step size below
 * 8 ulp(t), _basic.py:179-180) or NONFINITE; returns NI_ERR_REC_OVERFLOW when the record log filled up.  Either
 * pointer may be NULL. */
int ni_ensemble_poll(ni_ensemble *e, int64_t *n_unfinished, int64_t *n_failed);
/* ni.ff2t (_API.py:27-39).  is_not_last: uint8[N] host buffer or NULL. */
int ni_ensemble_ff2t(ni_ensemble *e, double t_end, uint8_t *is_not_last);
/* ni.ff2cond (_API.py:41-49) with a built-in threshold predicate NI_COND_* (y[index] or aux[index] against `value`);
 * hit: uint8[N] or NULL.  Like a user predicate it is fused into a copy of the run kernel (NVRTC, first use of a
 * (kind, index) pair per process; `value` is a kernel argument). */
int ni_ensemble_ff2cond(ni_ensemble *e, int cond_kind, int index, double value, uint8_t *hit, int64_t *n_hit);

/* ni.ff2cond with a user predicate (_API.py:41-49 takes any njit callable): `cond_body` is the body of
 *   __device__ bool cond(double t, const double* y, const double* aux, const double* cp)
 * fused (NVRTC) into this ensemble's run kernel and evaluated after every accepted step; cp = cond_params
 * (0..8 doubles).  CTA-per-trajectory kernels evaluate it on one thread, `y` being the whole state vector staged in
 * shared memory. */
int ni_ensemble_ff2cond_user(ni_ensemble *e, const char *cond_body, const double *cond_params, int n_params,
                             uint8_t *hit, int64_t *n_hit);

int ni_ensemble_get(ni_ensemble *e, int field, void *host_out);
int ni_ensemble_set(ni_ensemble *e, int field, const void *host_in); /* NI_F_T_BOUND, NI_F_H_ABS */
int ni_ensemble_device_ptr(ni_ensemble *e, int field, void **dptr);  /* zero-copy view (NCCL gather) */
int ni_ensemble_set_max_attempts(ni_ensemble *e, int64_t per_launch);
/* Tail compaction of the thread-per-trajectory kernels (no counterpart in the reference: one object = one trajectory).
 * With a threshold set, a fast-forward run of an ensemble that fills the GPU at least twice is three launches: once the
 * work queue is empty, a warp with fewer than `park_below` busy lanes writes its trajectories back (the state between
 * two attempts is exactly t, y, h_abs, K[-1]) and appends them to a list that the next launch integrates in full warps;
 * results are bit-identical to a single launch.  OFF by default: on B200 it measured 2 % SLOWER at 1.25 M Lorenz
 * trajectories per GPU and neutral at 10 M (the parked trajectories wait for the launch boundary).
 * park_below: 0 disables, 1..32 sets the threshold, < 0 leaves it.  parked_last_run (may be NULL): trajectories that
 * changed launch in the most recent run. */
int ni_ensemble_compaction(ni_ensemble *e, int park_below, int64_t *parked_last_run);

/* ---- terminal records: the payload of the final multi-GPU gather ------------------------
 * after the integration the only cross-GPU traffic is the gather of terminal
 * t, y, auxiliary and the step statistics (the reference has no counterpart: one jitclass object = one trajectory,
 * _basic.py:182-199 lists the same fields).  ni_ensemble_pack_terminal writes ONE packed row per trajectory,
 *   double t, y[n], aux[Q]; int32 n_accepted, n_rejected, status        (8 (1 + n + Q) + 12 bytes, 4-byte aligned)
 * into a caller-owned DEVICE buffer of N rows, asynchronously on the ensemble's stream, so that a caller holding a
 * communicator moves every shard with a single all-gather / gather of equal-sized rows
 * (numba_integrators_b200/distributed.py does it with torch.distributed over NCCL; the library itself carries no
 * communicator: one process drives one GPU).  ni_ensemble_get_terminal = pack + one device-to-host copy + sync. */
int ni_ensemble_terminal_bytes(const ni_ensemble *e, int64_t *bytes_per_trajectory);
int ni_ensemble_pack_terminal(ni_ensemble *e, void *dst_device);
int ni_ensemble_get_terminal(ni_ensemble *e, void *host_out);
/* Packing with the gather fused in: every packed word is stored to n_dst (<= 16) device buffers at once -- this GPU's
 * slot in the receive buffer of every GPU of the job, peer memory reached over NVLink / NVSwitch (the caller maps the
 * peers' buffers, e.g. torch.distributed._symmetric_memory, and synchronises the GPUs afterwards).  The transfer
 * overlaps the packing; no staging copy and no library collective on the path. */
int ni_ensemble_pack_terminal_peers(ni_ensemble *e, void *const *dst_device, int n_dst);
/* The gather fused into the INTEGRATION (thread-per-trajectory kernels): with a sink set, every trajectory that leaves
 * a run kernel for good (finished, failed, non-finite; also one that was finished before the launch) stores its terminal
 * record -- the packed record above, zero-padded to a multiple of 16 bytes: ni_ensemble_terminal_sink_bytes, 64 for
 * Lorenz-63 -- to row first_row + i of each of the n_dst (<= 8) device buffers, as 16-byte vectors.  The buffers are this
 * GPU's slot in the receive buffer of every GPU of the job (peer memory over NVLink / NVSwitch, mapped by the caller):
 * the one collective of the path (reference: none -- single process) then costs a barrier after the run,
 * not a transfer.  n_dst = 0 switches the sink off.  The setting is part of the launch arguments of every later run. */
int ni_ensemble_terminal_sink_bytes(const ni_ensemble *e, int64_t *bytes_per_trajectory);
/* Stiffness probe (the reference has none): rho_host[i] = estimate of the spectral radius of df/dy at
 * trajectory i's current (t, y), by `iterations` steps of the nonlinear power iteration v <- f(y + eps v) - f(y) started
 * from v = f(y) (a mode of the init kernel: works for every right-hand side, nothing is modified).  h_abs * rho near the
 * method's real stability interval (RK45 ~3.3, RK23 ~2.5) means the step size is limited by stability, not accuracy:
 * the problem is stiff at this tolerance and an explicit pair is the wrong tool.  Synchronous. */
int ni_ensemble_stiffness(ni_ensemble *e, int iterations, double *rho_host);
/* Launch order of the run kernels' work queue: the persistent grid pops order[0], order[1], ... instead of 0, 1, ...
 * (order = a permutation of 0..N-1 in host memory, copied; NULL restores the identity).  Results do not depend on it
 * -- a trajectory's arithmetic never depends on what shares its warp -- but the END of a run does: the last
 * trajectories popped finish alone, so the run ends one trajectory-duration after the queue is empty.  Handing out the
 * expensive trajectories first (longest-processing-time-first, e.g. by the step counts of the previous run or leg of the
 * same ensemble, or by a known cost parameter such as mu of a Van der Pol sweep) leaves the cheap ones for the tail. */
int ni_ensemble_set_order(ni_ensemble *e, const uint32_t *order_host);
/* The same with the cost taken from the ensemble itself, entirely on the device: trajectories in decreasing order of the
 * attempts (n_accepted + n_rejected) they have taken so far -- the previous run, or the previous leg of an integration in
 * legs -- by a counting sort into 2048 buckets (about a millisecond for 10 M trajectories; the order inside a bucket is
 * unspecified).  ni_ensemble_get_order copies the current order (the identity if none is set) to the host. */
int ni_ensemble_order_by_attempts(ni_ensemble *e);
int ni_ensemble_get_order(ni_ensemble *e, uint32_t *order_host);
/* Guard mode (memory-safety evidence where compute-sanitizer is not available): an ensemble created while the
 * environment variable NI_B200_GUARD=1 is set keeps every device buffer between two 64 KiB guard zones filled with 0xA5.
 * *n_buffers = guarded buffers (0 when the mode is off), *bad_bytes = guard bytes that no longer hold the fill value,
 * i.e. out-of-bounds writes of any kernel since creation (0 expected).  Synchronises the ensemble's stream. */
int ni_ensemble_guard_check(ni_ensemble *e, int64_t *n_buffers, int64_t *bad_bytes);
int ni_ensemble_set_terminal_sink(ni_ensemble *e, void *const *dst_device, int n_dst, int64_t first_row);


/* ---- recording of accepted steps into an HBM append log --------------------------------
 * (README.md:62-69 appends solver.t / solver.y after every accepted step.)  A record is (trajectory, t, y[n], aux[Q]):
 * 8 (1 + n + Q) bytes of payload + a 4-byte trajectory tag.  The log stores no step numbers: the records of one
 * trajectory appear in step order (slot order), interleaved with those of other trajectories.  record_count gives the
 * number of log SLOTS in use (an upper bound: warps reserve slots in chunks); record_drain needs buffers for that many
 * records, compacts the log on the device (window by window through a fixed staging buffer: unused slots dropped, order
 * kept), numbers the steps (`step`, may be NULL: the k-th record of a trajectory in this batch is step
 * n_accepted - records_in_batch + k) and returns the number of real records in *n_out. */
int ni_ensemble_record_enable(ni_ensemble *e, int64_t capacity);
int ni_ensemble_record_count(ni_ensemble *e, int64_t *count);
int ni_ensemble_record_drain(ni_ensemble *e, uint32_t *traj, uint32_t *step, double *t, double *y, double *aux,
                             int64_t max_records, int64_t *n_out);

/* device time of the most recent init / run kernels (CUDA events on the ensemble's stream) */
int ni_ensemble_timing(ni_ensemble *e, float *ms_init, float *ms_run, int64_t *n_launches);

/* measured FP64 (DFMA) peak of `device` in TFLOP/s: register-resident FMA chains, best of
 * `repeats` launches of `iters` x 128 FMAs per thread (roofline denominator for bench.py). */
int ni_fp64_peak(int device, int iters, int repeats, double *tflops_best);

/* device self-check of the branch-free arithmetic of the error-norm / controller tail (csrc/ni_math.cuh):
 * `count` random operands inside the guarded ranges; which = 0: ni_div_core against `/`, 1: ni_sqrt_core against
 * sqrt(), 2 / 3: ni_pow_ctrl_ms<5> / <3> (seeded from the mean square) against ni_pow_ctrl on the root.
 * *mismatches = results that differ in any bit (0 expected for 0 and 1; a few per 10^5 at most for 2 and 3).
 * which = 5: ni_div_core over the operand range the lean run body's guard admits (scales in [2^-200, 2^1020), any
 * error): violations of "bit-identical where the IEEE quotient is normal, non-finite where it overflows, negligible
 * otherwise".  which = 4: *mismatches = the largest |1 - b r| of the reciprocal r that ni_div_core multiplies with, in
 * units of 2^-64 (the quotient is correctly rounded unless it lies within ~|1 - b r| * 2^-52 of a rounding boundary). */
int ni_devmath_check(int device, int which, int64_t count, uint64_t seed, int64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* NI_B200_H */
