// ni_large.cuh -- one CTA per trajectory for large state dimension (method-of-lines systems).
//
// BASELINE config C5: n = 4096 heat / Burgers, RK45.  The working set of one attempt is
// y, K0..K6, y_new = 9 vectors x 32 KB = 288 KB -- more than the 227 KB of shared memory and more
// than the 256 KB register file, so it is split: every thread owns CPT CONTIGUOUS components and
// keeps y, K0 (the FSAL stage), K2 and the stage input in registers, the remaining stage vectors
// live in shared memory in a [slot][thread] layout (conflict-free 8-byte accesses).  Nothing of the
// stage loop touches HBM: global traffic is y0 in, terminal state out, plus recorded steps.
//
// The right-hand side is a 3-point stencil functor
//     Rhs::eval(t, y_left, y_centre, y_right, p, j, n) -> dy_j  (Dirichlet 0 outside 0..n-1)
// so neighbours inside a thread come from registers; only the two edge values per thread go through
// shared memory (double-buffered: one __syncthreads per RHS evaluation).  The RMS error norm is a
// warp-shuffle reduction followed by a fixed-order sum of the per-warp partials.
//
// Semantics are those of ni_core.cuh (reference _basic.py:114-180 / _advanced.py:126-195); tableau
// contractions use the same per-component order as the small kernels and the CPU oracle, the norm
// uses a tree order (the reference's own n >= 4 order is host-BLAS/LLVM dependent, SURVEY.md
// appendix B, so parity for n >= 4 is 1e-12 relative + identical step counts, not bitwise).
#pragma once
#ifndef NI_LARGE_CUH
#define NI_LARGE_CUH
#include "ni_core.cuh"


// p = (k): dy_j = k (y_{j-1} - 2 y_j + y_{j+1})
struct NiStencilHeat {
  static constexpr int P = 1, Q = 0;
  static constexpr bool COMPONENT = false;
  template <bool FAST>
  NI_DEV static double eval(double, double ym, double yc, double yp, const double *p, int, int) {
    using A = NiA<FAST>;
    // (ym - 2*yc) + yp.  2*yc is exact, so fma(-2, yc, ym) IS RN(ym - 2*yc), bit for bit, in one instruction
    return A::mul(p[0], A::add(fma(-2.0, yc, ym), yp));
  }
};
// p = (nu/dx^2, 1/(2dx)): dy_j = p0 (y_{j-1} - 2 y_j + y_{j+1}) - (y_j p1)(y_{j+1} - y_{j-1})
struct NiStencilBurgers {
  static constexpr int P = 2, Q = 0;
  static constexpr bool COMPONENT = false;
  template <bool FAST>
  NI_DEV static double eval(double, double ym, double yc, double yp, const double *p, int, int) {
    using A = NiA<FAST>;
    const double diff = A::mul(p[0], A::add(fma(-2.0, yc, ym), yp));  // exact 2*yc: see NiStencilHeat
    return A::nmadd(A::mul(yc, p[1]), A::sub(yp, ym), diff);  // diff - (yc*p1)*(yp - ym)
  }
};

// Which stage vectors stay in registers: K0 always (read by every stage), plus the stages in NI_LARGE_KREG_MASK45 /
// NI_LARGE_KREG_MASK23 (bit j = K_j).  A CTA of 256 threads x 255 registers keeps y, K0, K1, K2, K3, K5 and the stage
// input / output of an n = 4096 system (16 components per thread) in registers -- K5 is born after the last read of
// K1, so the two share registers -- and only K4 (3 reads per attempt) lives in shared memory.  Measured on C5
// (n = 4096, RK45), accepted steps/s: 512 threads x 128 registers with K2 in registers 1.95e7; 256 x 255 with K2
// 1.95e7, with K2 K3 2.08e7, with K1 K2 K3 2.11e7, with K1 K2 K3 K5 2.19e7 (K2 K3 K5: the same).
#ifndef NI_LARGE_KREG_MASK45
#define NI_LARGE_KREG_MASK45 (NI_LARGE_MAX_THREADS <= 256 ? 46 : NI_LARGE_MAX_THREADS <= 512 ? 4 : 0)
#endif
#ifndef NI_LARGE_KREG_MASK23
#define NI_LARGE_KREG_MASK23 (NI_LARGE_MAX_THREADS <= 256 ? 6 : 0)
#endif
template <int METHOD>
NI_HD constexpr bool ni_large_kreg(int j) {
  return j > 0 && j < 8 && (((METHOD == NI_RK45 ? (NI_LARGE_KREG_MASK45) : (NI_LARGE_KREG_MASK23)) >> j) & 1);
}
template <int METHOD, int J>
struct NiKInReg {
  static constexpr bool value = (J == 0) || ni_large_kreg<METHOD>(J);
};
// register array of stage vector J (meaningful when NiKInReg<METHOD, J> holds and J > 0), and how many there are
template <int METHOD>
struct NiLargeRegs {
  NI_HD static constexpr int index(int j) {
    int r = 0;
    for (int k = 1; k < j; ++k) r += ni_large_kreg<METHOD>(k) ? 1 : 0;
    return r;
  }
  static constexpr int COUNT = index(NiTab<METHOD>::S) > 0 ? index(NiTab<METHOD>::S) : 1;
};
// shared-memory slot of stage vector J (1..S-1; the new FSAL stage K_S never leaves registers) and the slot count
template <int METHOD>
struct NiLargeSlots {
  NI_HD static constexpr int slot(int j) {
    int s = 0;
    for (int k = 1; k < j; ++k) s += ni_large_kreg<METHOD>(k) ? 0 : 1;
    return s;
  }
  static constexpr int KSLOTS = slot(NiTab<METHOD>::S), TOTAL = KSLOTS > 0 ? KSLOTS : 1;
};

// dynamic shared memory of one CTA, in doubles.  Stage vectors are laid out [slot][component][NI_LARGE_MAX_THREADS]
// whatever the CTA size, so that every access is `thread base + immediate offset`.
// Component right-hand sides (dense coupling) additionally stage the whole stage vector: + threads * cpt doubles;
// stencil kernels carry a padded buffer of MAX_THREADS * (cpt + 1) doubles in which a warp transposes its part of a
// recorded row (the record log wants coalesced stores, the registers hold cpt consecutive components per thread).
NI_HD constexpr size_t ni_large_smem_doubles(int method, int threads, int cpt, bool component = false) {
  return (size_t)(method == NI_RK45 ? NiLargeSlots<NI_RK45>::TOTAL : NiLargeSlots<NI_RK23>::TOTAL) *
             NI_LARGE_MAX_THREADS * cpt + 4 * (size_t)NI_LARGE_MAX_THREADS + 2 * 32 + 8 +
         (component ? (size_t)threads * cpt : (size_t)NI_LARGE_MAX_THREADS * (cpt + 1));
}

struct NiLargeCtx {
  double *vec;    // TOTAL slots of [CPT][T]
  double *edge;   // [2 parity][2 sides][T]
  double *part;   // [2 parity][32] per-warp partial sums
  long long *bcast;
  double *ys;     // component right-hand sides only: the full stage vector [n]; stencil kernels: the transposition
                  // buffer of recorded rows, [warp][32][cpt + 1]
  int T, tid, n, j0;
  bool full;  // n == T * CPT: no padding components, masks are skipped
};


// Component index of local slot i.  Stencil kernels: CPT contiguous components per thread (neighbours in
// registers).  Component kernels: strided ownership (tid, tid + T, ...): conflict-free staging of the stage vector.
template <class Rhs>
NI_DEV int ni_large_j(const NiLargeCtx &c, int i) {
  if constexpr (Rhs::COMPONENT)
    return c.tid + i * c.T;
  else
    return c.j0 + i;
}

// neighbour exchange: returns y_{j0-1} and y_{j0+CPT} of this thread for the stage vector `v`.  Two staging buffers
// alternate, so one CTA barrier per exchange is enough.  Measured alternatives on C5: per-warp sequence flags +
// __threadfence_block so that a warp only waits for its two neighbour warps -- 2.6x SLOWER than BAR.SYNC (the fences
// and spinning lanes cost more than the barrier stalls they remove); a split barrier (mbarrier arrive after
// publishing the edges, inner components, try_wait, edge components) -- 2 % slower (more live registers, and the
// arrive / try_wait pair costs more than BAR.SYNC when the warps run in step anyway).
template <int CPT>
NI_DEV void ni_large_edges(const NiLargeCtx &c, const double (&v)[CPT], int parity, double &left, double &right) {
  double *eL = c.edge + (parity * 2 + 0) * NI_LARGE_MAX_THREADS, *eR = c.edge + (parity * 2 + 1) * NI_LARGE_MAX_THREADS;
  eL[c.tid] = v[0];
  eR[c.tid] = v[CPT - 1];
  __syncthreads();
  left = c.tid > 0 ? eR[c.tid - 1] : 0.0;
  right = c.tid + 1 < c.T ? eL[c.tid + 1] : 0.0;
}

// dy for the CPT components of this thread; components >= n stay exactly zero
template <class Rhs, int CPT, bool FAST = false>
NI_DEV void ni_large_rhs(NiLargeCtx &c, double t, const double (&v)[CPT], const double *p, int parity,
                         double (&out)[CPT]) {
  if constexpr (Rhs::COMPONENT) {
    // dense coupling: stage the whole vector in shared memory, every thread evaluates its own components from it
    (void)parity;
    __syncthreads();  // the previous stage vector is no longer read
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int j = c.tid + i * c.T;
      if (j < c.n) c.ys[j] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int j = c.tid + i * c.T;
      out[i] = j < c.n ? Rhs::template eval<FAST>(t, (const double *)c.ys, j, c.n, p) : 0.0;
    }
  } else {
    double left, right;
    ni_large_edges<CPT>(c, v, parity, left, right);
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const double ym = i == 0 ? left : v[i - 1];
      const double yp = i == CPT - 1 ? right : v[i + 1];
      out[i] = Rhs::template eval<FAST>(t, ym, v[i], yp, p, c.j0 + i, c.n);
    }
    if (!c.full) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) out[i] = (c.j0 + i < c.n) ? out[i] : 0.0;
    }
  }
}

// sum over the block of `x`; all threads get it.  Fixed order: lanes by butterfly, then the per-warp partials by a
// pairwise tree over all NI_LARGE_MAX_THREADS / 32 slots (slots of absent warps hold 0: ni_large_ctx clears them).
// Every thread sums redundantly, so this sits on the serial path of an attempt: the tree is 4 dependent additions
// where a running sum over 16 warps was 15.
NI_DEV double ni_large_block_sum(const NiLargeCtx &c, double x, int parity) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = __dadd_rn(x, __shfl_xor_sync(0xffffffffu, x, o));
  constexpr int NW = NI_LARGE_MAX_THREADS / 32;
  double *part = c.part + parity * 32;
  if ((c.tid & 31) == 0) part[c.tid >> 5] = x;
  __syncthreads();
  double v[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) v[w] = part[w];
#pragma unroll
  for (int st = 1; st < NW; st <<= 1) {
#pragma unroll
    for (int k = 0; k + st < NW; k += 2 * st) v[k] = __dadd_rn(v[k], v[k + st]);
  }
  return v[0];
}

template <int CPT>
NI_DEV double ni_large_rms(const NiLargeCtx &c, const double (&x)[CPT], int parity) {
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < CPT; ++i) acc = __dadd_rn(acc, __dmul_rn(x[i], x[i]));
  return sqrt(ni_large_block_sum(c, acc, parity) / (double)c.n);
}

NI_DEV NiLargeCtx ni_large_ctx(const NiArgs &a, double *smem, int cpt, int total_slots) {
  NiLargeCtx c;
  c.T = blockDim.x;
  c.tid = threadIdx.x;
  c.n = a.n;
  c.j0 = c.tid * cpt;
  c.full = a.n == (int)blockDim.x * cpt;
  c.vec = smem;
  c.edge = smem + (size_t)total_slots * NI_LARGE_MAX_THREADS * cpt;
  c.part = c.edge + 4 * NI_LARGE_MAX_THREADS;
  c.bcast = (long long *)(c.part + 64);
  c.ys = (double *)(c.bcast + 8);
  for (int k = c.tid; k < 64; k += c.T) c.part[k] = 0.0;  // slots of absent warps stay zero (ni_large_block_sum)
  __syncthreads();
  return c;
}

// The state vector staged contiguously in shared memory, for the two consumers that see a CTA-per-trajectory system as
// a whole: the auxiliary function (_advanced.py:181: the second return value of the right-hand side at the committed
// state) and the ff2cond predicate (_API.py:41-49).  Both are user code over (t, y[0..n), p), evaluated by thread 0.
// The stage slots are dead between two attempts, so the vector borrows them (n <= T * CPT doubles always fit one slot).
template <class Rhs, int CPT>
NI_DEV const double *ni_large_stage(NiLargeCtx &c, const double (&y)[CPT]) {
  double *dst = Rhs::COMPONENT ? c.ys : c.vec;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int j = ni_large_j<Rhs>(c, i);
    if (j < c.n) dst[j] = y[i];
  }
  __syncthreads();
  return dst;
}
// auxiliary output of trajectory idx at (t, y) -> a.aux[idx][0..Q)
template <class Rhs, int CPT>
NI_DEV void ni_large_store_aux(const NiArgs &a, NiLargeCtx &c, long long idx, double t, const double (&y)[CPT],
                               const double *p) {
  if constexpr (Rhs::Q > 0) {
    const double *ys = ni_large_stage<Rhs, CPT>(c, y);
    if (c.tid == 0) {
      double aux[Rhs::Q];
      Rhs::aux(t, ys, c.n, p, aux);
#pragma unroll
      for (int i = 0; i < Rhs::Q; ++i) a.aux[idx * Rhs::Q + i] = aux[i];
    }
  }
}

// ---------------------------------------------------------------- init (RK.__init__ + select_initial_step)
// Stiffness probe of the CTA-per-trajectory systems (see ni_small_probe_body): the same power iteration with the
// norms taken across the CTA.
template <class Rhs, int CPT>
NI_DEV void ni_large_probe_body(const NiArgs &a, NiLargeCtx &c) {
  constexpr int P = Rhs::P;
  const int n = c.n;
  for (long long idx = blockIdx.x; idx < a.N; idx += gridDim.x) {
    double y[CPT], f0[CPT], v[CPT], y1[CPT], f1[CPT], p[P > 0 ? P : 1];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int j = ni_large_j<Rhs>(c, i);
      y[i] = j < n ? a.y[idx * n + j] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = a.params[idx * P + i];
    const double t = a.t[idx];
    int parity = 0;
    ni_large_rhs<Rhs, CPT, false>(c, t, y, p, parity, f0);
    parity ^= 1;
    double ys = 0.0, vs = 0.0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      v[i] = ni_large_j<Rhs>(c, i) < n ? f0[i] : 0.0;
      ys += y[i] * y[i];
      vs += v[i] * v[i];
    }
    const double ynorm = sqrt(ni_large_block_sum(c, ys, 0));
    double vn2 = ni_large_block_sum(c, vs, 1);
    if (!(vn2 > 0.0)) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) v[i] = ni_large_j<Rhs>(c, i) < n ? 1.0 : 0.0;
      vn2 = (double)n;
    }
    double rho = 0.0;
    for (int it = 0; it < a.probe_iters; ++it) {
      const double vnorm = sqrt(vn2);
      const double eps = 1.4901161193847656e-08 * (1.0 + ynorm) / vnorm;
#pragma unroll
      for (int i = 0; i < CPT; ++i) y1[i] = y[i] + eps * v[i];
      ni_large_rhs<Rhs, CPT, false>(c, t, y1, p, parity, f1);
      parity ^= 1;
      double ds = 0.0;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        v[i] = ni_large_j<Rhs>(c, i) < n ? f1[i] - f0[i] : 0.0;
        ds += v[i] * v[i];
      }
      const double dn2 = ni_large_block_sum(c, ds, it & 1);
      rho = sqrt(dn2) / (eps * vnorm);
      if (!(dn2 > 0.0)) break;  // (uniform across the CTA)
      vn2 = dn2;
    }
    if (c.tid == 0) a.probe_out[idx] = rho;
    __syncthreads();
  }
}

template <class Rhs, int METHOD, int CPT>
NI_DEV void ni_large_init_body(const NiArgs &a, double *smem) {
  constexpr int P = Rhs::P;
  NiLargeCtx c = ni_large_ctx(a, smem, CPT, NiLargeSlots<METHOD>::TOTAL);
  const int n = c.n;
  if (a.probe_iters > 0) {
    ni_large_probe_body<Rhs, CPT>(a, c);
    return;
  }
  for (long long idx = blockIdx.x; idx < a.N; idx += gridDim.x) {
    double y[CPT], f0[CPT], p[P > 0 ? P : 1], rt[CPT], at[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int j = ni_large_j<Rhs>(c, i);
      y[i] = j < n ? a.y0_rows[idx * n + j] : 0.0;
      rt[i] = j < n ? a.rtol_d[j] : 0.0;
      at[i] = j < n ? a.atol_d[j] : 1.0;
    }
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = a.params_rows[idx * P + i];
    const double t0 = a.t0[idx * a.t0_stride], tb = a.tb0[idx * a.tb_stride];
    ni_large_rhs<Rhs, CPT, false>(c, t0, y, p, 0, f0);
    const double dir = tb != t0 ? (tb > t0 ? 1.0 : -1.0) : 1.0;
    double h_abs;
    const double first_step = ni_first_step(a, idx);
    if (first_step == 0.0) {
      // reference: numba fastmath contracts `scale` and `y1` to FMAs; scipy semantics
      // (a.compat): unfused, h0 capped by the interval, h1 = (0.01 / max(d1, d2)) ** (1 / (order + 1)), result capped
      // by the interval and max_step
      const bool sp = a.compat != 0;
      const double interval = fabs(__dadd_rn(tb, -t0));
      double scale[CPT], tmp[CPT];
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        scale[i] = sp ? __dadd_rn(at[i], __dmul_rn(fabs(y[i]), rt[i])) : fma(fabs(y[i]), rt[i], at[i]);
        tmp[i] = y[i] / scale[i];
      }
      const double d0 = ni_large_rms<CPT>(c, tmp, 0);
#pragma unroll
      for (int i = 0; i < CPT; ++i) tmp[i] = f0[i] / scale[i];
      const double d1 = ni_large_rms<CPT>(c, tmp, 1);
      double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : __dmul_rn(0.01, d0) / d1;
      if (sp) h0 = fmin(h0, interval);
      const double hd = __dmul_rn(h0, dir);
      double y1[CPT], f1[CPT];
#pragma unroll
      for (int i = 0; i < CPT; ++i) y1[i] = sp ? __dadd_rn(y[i], __dmul_rn(hd, f0[i])) : fma(hd, f0[i], y[i]);
      ni_large_rhs<Rhs, CPT, false>(c, __dadd_rn(t0, hd), y1, p, 1, f1);
#pragma unroll
      for (int i = 0; i < CPT; ++i) tmp[i] = __dadd_rn(f1[i], -f0[i]) / scale[i];
      const double d2 = ni_large_rms<CPT>(c, tmp, 0) / h0;
      double h1;
      if (d1 <= 1e-15 && d2 <= 1e-15)
        h1 = fmax(1e-6, __dmul_rn(h0, 1e-3));
      else if (sp)
        h1 = pow(0.01 / fmax(d1, d2), 1.0 / (NiTab<METHOD>::ORDER + 1));
      else
        h1 = ni_pow_init<NiTab<METHOD>::ORDER + 1>(__dmul_rn(fmax(d1, d2), 100.0));
      h_abs = fmin(__dmul_rn(100.0, h0), h1);
      if (sp) h_abs = interval == 0.0 ? 0.0 : fmin(h_abs, fmin(interval, ni_max_step(a, idx)));
    } else {
      h_abs = fabs(first_step);
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int j = ni_large_j<Rhs>(c, i);
      if (j < n) {
        a.y[idx * n + j] = y[i];
        a.klast[idx * n + j] = f0[i];
      }
    }
    if (c.tid == 0) {
      a.t[idx] = t0;
      a.t_bound[idx] = tb;
      a.direction[idx] = dir;
      a.h_abs[idx] = h_abs;
      a.step_size[idx] = __dmul_rn(dir, h_abs);
#pragma unroll
      for (int i = 0; i < P; ++i) a.params[idx * P + i] = p[i];
      a.status[idx] = NI_ST_RUNNING;
      a.n_acc[idx] = 0;
      a.n_rej[idx] = 0;
      a.running[idx] = 1;
    }
    ni_large_store_aux<Rhs, CPT>(a, c, idx, t0, y, p);  // the auxiliary output of the FSAL seed evaluation
    __syncthreads();  // edge / partial buffers are reused by the next trajectory
  }
}

// ---------------------------------------------------------------- run
// SEM: 0 basic (_step), 1 Advanced (h = h_abs*direction), 2 scipy.integrate semantics; FAST: opt-in fast arithmetic
// (both as in ni_small_run_body; the AOT library carries SEM 0/1 strict, the other variants are built with NVRTC)
template <class Rhs, int METHOD, int SEM, bool REC, int CPT, bool FAST = false>
NI_DEV void ni_large_run_body(const NiArgs &a, double *smem) {
  constexpr bool ADV = SEM >= 1, SP = SEM == 2;
  constexpr int P = Rhs::P;
  using Tab = NiTab<METHOD>;
  using Slots = NiLargeSlots<METHOD>;
  constexpr int S = Tab::S;
  NiLargeCtx c = ni_large_ctx(a, smem, CPT, Slots::TOTAL);
  const int n = c.n, T = c.T, tid = c.tid;
  // shared stage vector J, component i of this thread
  double *const svb = c.vec + tid;
  auto sv = [&](int slot, int i) -> double & { return svb[(slot * CPT + i) * NI_LARGE_MAX_THREADS]; };

  for (;;) {
    // ---- next trajectory (one queue pop per CTA)
    if (tid == 0) {
      long long idx = (long long)atomicAdd(a.queue, 1ull);
      if (a.work_list) idx = idx < (long long)*a.work_count ? (long long)a.work_list[idx] : a.N;  // ni_ensemble_set_order
      if (REC && (((volatile unsigned long long *)a.queue)[1] & NI_FLAG_REC_OVERFLOW)) idx = a.N;
      c.bcast[0] = idx;
      if (REC) c.bcast[3] = c.bcast[4] = c.bcast[5] = 0;  // this trajectory holds no record slots yet
    }
    __syncthreads();
    const long long idx = c.bcast[0];
    __syncthreads();
    if (idx >= a.N) break;
    int status = a.status[idx];
    if (status != NI_ST_RUNNING) {
      if (tid == 0) a.running[idx] = 0;
      continue;
    }
#if NI_COND
    if (a.cond_hit[idx]) continue;  // ff2cond: already stopped at its predicate
#endif
    bool parked = false;  // stopped by the ff2cond predicate: still RUNNING, but not "unfinished"
    double t = a.t[idx], h_abs = a.h_abs[idx], step_h = a.step_size[idx];
    const double tb = a.t_bound[idx], dir = a.direction[idx], mxs = ni_max_step(a, idx);
    double y[CPT], k0[CPT], kr[NiLargeRegs<METHOD>::COUNT][CPT], p[P > 0 ? P : 1];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int j = ni_large_j<Rhs>(c, i);
      y[i] = j < n ? a.y[idx * n + j] : 0.0;
      k0[i] = j < n ? a.klast[idx * n + j] : 0.0;
#pragma unroll
      for (int r = 0; r < NiLargeRegs<METHOD>::COUNT; ++r) kr[r][i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < P; ++i) p[i] = a.params[idx * P + i];
    int nacc = a.n_acc[idx], nrej = a.n_rej[idx];
    long long att_left = a.max_attempts, acc_left = a.max_accepts;
    bool last_accept = false, fresh = true, done = false;
    bool srej = SP ? a.step_rejected[idx] != 0 : false;  // scipy: a rejection happened in the current step
    double eps = 0, min_step = 0;
    // REC: record slots are reserved in chunks of 8 per trajectory (one atomicAdd per chunk instead of per accepted
    // step; lock-step calls take one at a time) and live in shared memory: bcast[3] / bcast[5] = the next free slot,
    // read by the even / odd accepted steps (thread 0 writes the entry the NEXT step reads, so no thread can see a
    // half-updated one), bcast[4] = end of the chunk.  A trajectory only attempts a step with a slot in hand
    // (`have_slot` here and after every accepted step), so an accepted step always has its row and a full log stops the
    // trajectory BETWEEN attempts: nothing is undone and nothing extra stays live across the attempt (the kernel runs
    // at the 255-register limit).  Slots left over when the trajectory leaves are tagged NI_REC_INVALID.
    auto have_slot = [&]() -> bool {  // (uniform across the CTA)
      __syncthreads();  // thread 0's update of the slot entries at the accepted step just recorded
      const int rp = 3 + 2 * (nacc & 1);
      long long s_next = c.bcast[rp];
      if (s_next >= c.bcast[4]) {  // a new chunk
        __syncthreads();
        if (tid == 0) {
          const long long k = a.max_accepts >= 0x7fffffffLL ? 8 : 1;
          const long long base = (long long)atomicAdd(a.rec_cursor, (unsigned long long)k);
          c.bcast[rp] = base;
          c.bcast[4] = base + k;
        }
        __syncthreads();
        s_next = c.bcast[rp];
      }
      if (s_next >= a.rec_cap) {  // the log is full: the trajectory waits for the drain (still RUNNING)
        if (tid == 0) atomicOr(a.queue + 1, NI_FLAG_REC_OVERFLOW);
        return false;
      }
      return true;
    };
    if constexpr (REC) {
      if (!(__dmul_rn(dir, __dadd_rn(t, -tb)) >= 0.0) && !have_slot()) done = true;
    }

    while (!done) {
      if (fresh) {  // step entry (_basic.py:135-143)
        if (__dmul_rn(dir, __dadd_rn(t, -tb)) >= 0.0) {
          step_h = ADV ? h_abs : __dmul_rn(dir, h_abs);
          status = NI_ST_FINISHED;
          last_accept = false;
          break;
        }
        eps = ni_ulp_eps(t, dir);
        min_step = __dmul_rn(SP ? 10.0 : 8.0, eps);
        if (SP) {
          if (!srej) h_abs = h_abs > mxs ? mxs : (h_abs < min_step ? min_step : h_abs);
        } else if (h_abs < min_step) {
          h_abs = min_step;
        }
        fresh = false;
      }
      // ---- one attempt (_basic.py:145-169)
      if (SP && h_abs < min_step) {  // scipy tests the step size before every attempt and commits nothing
        status = NI_ST_FAILED;
        last_accept = false;
        break;
      }
      if (!SP && h_abs > mxs) h_abs = __dadd_rn(mxs, -eps);
      double h = (ADV || SP) ? __dmul_rn(h_abs, dir) : h_abs;
      double tn = __dadd_rn(t, h);
      {
        const double cross = __dmul_rn(dir, __dadd_rn(tn, -tb));
        const bool clip = SP ? cross > 0.0 : cross >= 0.0;
        if (clip) tn = tb;
        if (clip || SP) {
          h = __dadd_rn(tn, -t);
          h_abs = fabs(h);
        }
      }
      // K_j of component i: register-resident or shared
      auto kget = [&](auto jj, int i) -> double {
        constexpr int j = decltype(jj)::value;
        if constexpr (j == 0)
          return k0[i];
        else if constexpr (NiKInReg<METHOD, j>::value)
          return kr[NiLargeRegs<METHOD>::index(j)][i];
        else
          return sv(Slots::slot(j), i);
      };
      int parity = 0;
      ni_static_for<1, S>([&](auto ss) {
        constexpr int s = decltype(ss)::value;
        double ys[CPT], ks[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          const double d = ni_dot_col<METHOD, s, s, false>([&](auto jj) { return kget(jj, i); });
          ys[i] = NiA<FAST>::madd(d, h, y[i]);
        }
        ni_large_rhs<Rhs, CPT, FAST>(c, NiA<FAST>::madd(ni_cval<METHOD, 0, s>(), h, t), ys, p, parity, ks);
        parity ^= 1;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          if constexpr (NiKInReg<METHOD, s>::value)
            kr[NiLargeRegs<METHOD>::index(s)][i] = ks[i];
          else
            sv(Slots::slot(s), i) = ks[i];
        }
      });
      double yn[CPT], kn[CPT];
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        const double d = ni_dot_col<METHOD, S, S, false>([&](auto jj) { return kget(jj, i); });
        yn[i] = NiA<FAST>::madd(h, d, y[i]);
      }
      ni_large_rhs<Rhs, CPT, FAST>(c, tn, yn, p, parity, kn);
      parity ^= 1;
      // ---- error / scale and its sum of squares for the CPT components of this thread (_basic.py:166-169).
      // Three passes over the components, each free of branches, so that the CPT dot products and the CPT quotients
      // interleave: one component after the other (dot, tolerance test, operator `/` with its slow-path branch) is
      // a chain of ~30 dependent FP64 operations per component and was a quarter of the kernel time on C5.  Strict
      // arithmetic uses the guarded branch-free quotient of ni_math.cuh; a thread whose operands leave the guarded
      // range (a denormal or non-finite error; a zero-error padding component is fine) redoes them with the operator.
      double sq = 0.0;
      auto error_terms = [&](auto uniform) {
        constexpr bool UNI = decltype(uniform)::value != 0;
        double num[CPT], den[CPT], e[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          const double ed = ni_dot_col<METHOD, S + 1, S + 1, false>([&](auto jj) {
            if constexpr (decltype(jj)::value == S)
              return kn[i];
            else
              return kget(jj, i);
          });
          num[i] = __dmul_rn(ed, h);
        }
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          const int j = ni_large_j<Rhs>(c, i);
          double rt, at;
          if constexpr (UNI) {  // uniform tolerances come from the constant bank; a padding component is 0 / atol
            rt = a.rtol[0];
            at = a.atol[0];
          } else {
            const int jc = j < n ? j : 0;
            rt = j < n ? __ldg(a.rtol_d + jc) : 0.0;
            at = j < n ? __ldg(a.atol_d + jc) : 1.0;
          }
          // integer max of the magnitudes (no FP64 compare); a NaN operand gives a NaN scale where fmax would have
          // dropped it, but then the error is NaN as well and the attempt ends in NI_ST_NONFINITE either way
          const double big = FAST ? fmax(fabs(y[i]), fabs(yn[i])) : ni_absmax(y[i], yn[i]);
          den[i] = FAST ? fma(big, rt, at) : __dadd_rn(at, __dmul_rn(big, rt));
        }
        if constexpr (FAST) {
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            e[i] = __dmul_rn(num[i], ni_rcp_fast(den[i]));
            sq = fma(e[i], e[i], sq);
          }
        } else {
          NiDivGuard guard = ni_guard_begin();
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            ni_guard_add(guard, num[i], den[i]);
            e[i] = ni_div_core(num[i], den[i]);
          }
          if (ni_guard_out(guard) != 0u) {
#pragma unroll
            for (int i = 0; i < CPT; ++i) e[i] = num[i] / den[i];
          }
          // pairwise tree (log2 CPT dependent additions; the threads' sums are then combined by ni_large_block_sum)
#pragma unroll
          for (int i = 0; i < CPT; ++i) e[i] = __dmul_rn(e[i], e[i]);
#pragma unroll
          for (int st = 1; st < CPT; st <<= 1) {
#pragma unroll
            for (int k = 0; k + st < CPT; k += 2 * st) e[k] = __dadd_rn(e[k], e[k + st]);
          }
          sq = e[0];
        }
      };
      if (a.tol_uniform && (c.full || a.atol[0] > 0.0))  // (padding components need atol > 0: theirs is 0 / atol)
        error_terms(NiInt<1>{});
      else
        error_terms(NiInt<0>{});
      // strict: the RMS norm; fast: the mean square (the controller then takes its 2K-th root, ni_pow_fast)
      const double ssq = ni_large_block_sum(c, sq, parity & 1);
      double en, pw;
      bool ok, nan;
      if constexpr (FAST) {
        en = ssq / (double)n;
        pw = ni_pow_fast<2 * (Tab::ORDER + 1)>(en);
        ok = en < 1.0;
        nan = !ok && !(en >= 1.0);
      } else {
        // uniform across the CTA: quotient, root and controller power branch-free under the guards, the operators
        // otherwise (see ni_small_run_body)
        // (n a power of two: the reciprocal is exact; a non-finite, zero or denormal ssq fails the guard on msq)
        const double msq = (n & (n - 1)) == 0 ? __dmul_rn(ssq, 1.0 / (double)n) : ni_div_core(ssq, (double)n);
        en = ni_sqrt_core(msq);
        pw = ni_pow_ctrl_ms<Tab::ORDER + 1>(msq, en);
        ok = (unsigned)__double2hiint(en) < 0x3ff00000u;
        nan = false;
        if (ni_sqrt_out(msq) != 0u) {
          en = sqrt(ssq / (double)n);
          pw = ni_pow_ctrl<Tab::ORDER + 1>(en);
          ok = en < 1.0;
          nan = !ok && !(en >= 1.0);
        }
      }
      --att_left;

      // ---- controller + commit: uniform across the CTA
      constexpr bool park = false;  // (record-log back-pressure acts before the attempt)
      long long slot = -1;
      if (REC && ok) slot = c.bcast[3 + 2 * (nacc & 1)];
      const double sp = __dmul_rn(0.9, pw);
      double fac = ok ? (en == 0.0 ? 10.0 : fmin(10.0, sp)) : fmax(0.2, sp);
      if (SP && ok && srej) fac = fmin(1.0, fac);  // scipy: no growth in a step that saw a rejection
      const double h_new = __dmul_rn(h_abs, fac);
      const bool fail = !SP && !ok && !nan && h_new < min_step;
      const bool take = (ok && !park) || fail;
      h_abs = nan ? h_abs : h_new;
      if (take) {
        t = tn;
#pragma unroll
        for (int i = 0; i < CPT; ++i) y[i] = yn[i];
      }
      if (take || nan) step_h = h;
      if (!park && !(SP && !ok)) {  // scipy keeps f(t, y) of the accepted point on a rejection
#pragma unroll
        for (int i = 0; i < CPT; ++i) k0[i] = kn[i];  // K[0] = K[-1] of the next attempt (stale after a reject)
      }
      if (!park) srej = !ok;
      const bool accepted = ok && !park;
      bool cond_hit = false;
      if (accepted) {
        ++nacc;
        --acc_left;
        if (REC) {
          // The recorded row (32 776 B per step for n = 4096).  Stencil layout: a thread owns CPT consecutive components,
          // so a store straight from the registers touches 32 different 128-byte lines per warp instruction.  Each warp
          // transposes its 32 * CPT consecutive components through a padded shared-memory buffer ([lane][CPT + 1]:
          // conflict-free both ways, every access is `per-lane base + immediate`, only __syncwarp is needed) and writes
          // them with fully coalesced stores.  A bulk copy (cp.async.bulk) of the row would need the LINEAR image in
          // shared memory, i.e. the conflict-ridden pattern (lane stride CPT * 8 = 128 bytes = all 32 banks).
          if constexpr (!Rhs::COMPONENT) {
            const int lane = tid & 31, w0 = (tid & ~31) * CPT;
            double *const wb = c.ys + (tid >> 5) * (32 * (CPT + 1));
            double *const wr = wb + lane * (CPT + 1);
#pragma unroll
            for (int i = 0; i < CPT; ++i) wr[i] = y[i];
            __syncwarp();
            // element g = k * 32 + lane of the warp's chunk sits at [g / CPT][g % CPT]
            const double *const rd = wb + (lane / CPT) * (CPT + 1) + lane % CPT;
            double *const row = a.rec_y + slot * n + w0 + lane;
            if (c.full) {
#pragma unroll
              for (int k = 0; k < CPT; ++k) row[k * 32] = rd[k * (32 / CPT) * (CPT + 1)];
            } else {
#pragma unroll
              for (int k = 0; k < CPT; ++k)
                if (w0 + k * 32 + lane < n) row[k * 32] = rd[k * (32 / CPT) * (CPT + 1)];
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int i = 0; i < CPT; ++i)
              if (ni_large_j<Rhs>(c, i) < n) a.rec_y[slot * n + ni_large_j<Rhs>(c, i)] = y[i];
          }
          if (tid == 0) {
            a.rec_traj[slot] = (unsigned int)idx;
            a.rec_t[slot] = t;
            c.bcast[3 + 2 * (nacc & 1)] = slot + 1;  // (nacc already counts this step: the entry the next one reads)
          }
        }
        if constexpr ((REC && Rhs::Q > 0) || NI_COND != 0) {
          // the recorded auxiliary output / the predicate see the committed state as a whole (thread 0)
          const double *ys = ni_large_stage<Rhs, CPT>(c, y);
          if (tid == 0) {
            double aux[Rhs::Q > 0 ? Rhs::Q : 1];
            aux[0] = 0.0;
            if constexpr (Rhs::Q > 0) {
              Rhs::aux(t, ys, n, p, aux);
              if (REC) {
#pragma unroll
                for (int i = 0; i < Rhs::Q; ++i) a.rec_aux[slot * Rhs::Q + i] = aux[i];
              }
            }
#if NI_COND
            c.bcast[2] = ni_user_cond(t, ys, aux, a.cond_params) ? 1 : 0;
#endif
          }
#if NI_COND
          __syncthreads();
          cond_hit = c.bcast[2] != 0;
#endif
        }
      } else if (!park) {
        ++nrej;
      }
      fresh = accepted;
      last_accept = accepted;
      if (fail) status = NI_ST_FAILED;
      if (nan) status = NI_ST_NONFINITE;
      done = park || fail || nan;  // (separate statements: see the nvcc note in ni_small_run_body)
      if (att_left <= 0) done = true;
      if (accepted && acc_left <= 0) done = true;
      if (cond_hit) {  // ff2cond: the trajectory stops here (_API.py:45-49)
        if (tid == 0) a.cond_hit[idx] = 1;
        parked = true;
        done = true;
      }
      if (accepted && !done && __dmul_rn(dir, __dadd_rn(t, -tb)) >= 0.0) {
        step_h = ADV ? h_abs : __dmul_rn(dir, h_abs);  // terminal step() of `while step(): pass`
        status = NI_ST_FINISHED;
        last_accept = false;
        done = true;
      }
      if constexpr (REC) {
        if (accepted && !done && !have_slot()) done = true;  // (a full log: leave between two attempts)
      }
    }
    if (REC && tid == 0) {  // reserved, never used
      const long long end = c.bcast[4] < a.rec_cap ? c.bcast[4] : a.rec_cap;
      for (long long sl = c.bcast[3 + 2 * (nacc & 1)]; sl < end; ++sl) a.rec_traj[sl] = NI_REC_INVALID;
    }

    // ---- retire
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int j = ni_large_j<Rhs>(c, i);
      if (j < n) {
        a.y[idx * n + j] = y[i];
        a.klast[idx * n + j] = k0[i];
      }
    }
    ni_large_store_aux<Rhs, CPT>(a, c, idx, t, y, p);  // (_advanced.py:181: the auxiliary output at the committed state)
    if (tid == 0) {
      a.t[idx] = t;
      a.h_abs[idx] = h_abs;
      a.step_size[idx] = step_h;
      a.status[idx] = status;
      a.n_acc[idx] = nacc;
      a.n_rej[idx] = nrej;
      a.running[idx] = last_accept ? 1 : 0;
      if (SP) a.step_rejected[idx] = srej ? 1 : 0;
      if (status == NI_ST_RUNNING && !parked) atomicAdd(a.queue + 2, 1ull);
      if (status >= NI_ST_FAILED) atomicAdd(a.queue + 5, 1ull);
      if (last_accept) atomicAdd(a.queue + 3, 1ull);
    }
  }
}

template <class Rhs, int METHOD, int CPT>
__global__ void __launch_bounds__(NI_LARGE_MAX_THREADS, 1) ni_k_large_init(const NiArgs a) {
  extern __shared__ double ni_smem[];
  ni_large_init_body<Rhs, METHOD, CPT>(a, ni_smem);
}
template <class Rhs, int METHOD, int SEM, bool REC, int CPT>
__global__ void __launch_bounds__(NI_LARGE_MAX_THREADS, 1) ni_k_large_run(const NiArgs a) {
  extern __shared__ double ni_smem[];
  ni_large_run_body<Rhs, METHOD, SEM, REC, CPT, false>(a, ni_smem);
}
#endif  // NI_LARGE_CUH
