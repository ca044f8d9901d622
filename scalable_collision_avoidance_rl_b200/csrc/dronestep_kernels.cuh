// dronestep_kernels.cuh -- fused sm_100a kernels for the drone_env.step() hot path.
//
// Work item = one ROW of one FRAME: agent i of environment e at one time slice.
// A CTA owns G whole environments and TC consecutive time slices of them
// (G * n * TC <= blockDim.x rows, one row per thread):
//
//   * step_kernel     TC = 1: drones.step() / rewards() for E environments.
//   * rollout_kernel  T fused steps walked in chunks of TC slices.  The
//       integrator is a single integrator (x <- x + dt u, drone_env.py:78-79,235)
//       and the actions of a rollout are given up front, so the positions of all TC
//       slices of a chunk are produced first (sequential, bit-exact additions) and
//       the TC frames are then evaluated CONCURRENTLY: time becomes a parallel
//       axis, which is what fills 148 SMs when E * n is only ~4e4.
//       Early termination (drone_env.py:248-256) is resolved after the frames are
//       evaluated and before anything is stored: results of slices behind a
//       finishing slice are dropped, so the observable behaviour is exactly that
//       of stepping one slice at a time.
//
// A row is evaluated in two passes over the other agents of its frame (positions
// staged in shared memory; the n x n matrix never exists in memory):
//   pass 1  squared distance against a per-agent threshold: every pair that is
//           provably clipped to d_safety (drone_env.py:318 min(.., d_safety[i]))
//           contributes log(1) = 0, no collision, and a Delta-disk count that is a
//           per-agent constant -- no sqrt.  The step kernel does it in fp64 (7
//           instructions per pair); the rollout kernel on the packed-f32 pipe
//           (FADD2 / FMUL2 / FFMA2, two pairs per instruction) against a threshold
//           with a margin: the mask only has to be a superset of the live pairs;
//   pass 2  only the NEAR pairs take the exact path: sqrt, clip, zero rule,
//           division, log, collision test, k-nearest insert.
// In the rollout kernel the near pairs of the rows of a warp are compacted into a
// shared-memory work list (row-contiguous, ascending j) and evaluated by all
// lanes of the warp, one pair per lane per round, so that the expensive part
// (sqrt / div / log) is load balanced instead of paying for the row with the most
// neighbours in every warp; each row then folds its own segment of results in
// ascending j, the summation order of the reference (drone_env.py:282-283).
// The k+1 nearest are kept in (distance, index) lexicographic order, which is the
// stable argsort order (drone_env.py:338); clipped agents all tie at d_safety and
// enter in index order.
//
// Reference semantics (file:line in the reference's drone_env.py):
//   integrate 227-238 | distance_data 295-334 | rewards 260-293 |
//   localized_states 336-401 | termination 248-256.
// Arithmetic on the decision chain (integrate, distance, clip, thresholds) uses
// the *_rn intrinsics so that ptxas cannot contract it into FMAs the reference
// does not perform; the one FMA the reference's BLAS ddot does perform
// (sqrt(fma(dy,dy,dx*dx)), see oracle/drone_oracle.c) is written explicitly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define DS_DEVICE_CODE 1
#else
#define DS_DEVICE_CODE 0
#endif
#define DS_HD __host__ __device__ __forceinline__

namespace ds {

constexpr int kMaxK = 16;

// ---------------------------------------------------------------- arithmetic
// (host branches exist only so that tools/row_check can run the row logic on the
//  CPU against the oracle; the library itself never computes on the host.)
DS_HD double add_rn(double a, double b)
{
#if DS_DEVICE_CODE
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
DS_HD double sub_rn(double a, double b)
{
#if DS_DEVICE_CODE
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
DS_HD double mul_rn(double a, double b)
{
#if DS_DEVICE_CODE
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
DS_HD double fma_rn(double a, double b, double c)
{
#if DS_DEVICE_CODE
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
DS_HD double div_rn(double a, double b)
{
#if DS_DEVICE_CODE
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
DS_HD double sqrt_rn(double a)
{
#if DS_DEVICE_CODE
    return __dsqrt_rn(a);
#else
    return sqrt(a);
#endif
}

// ---- natural log of the barrier term (drone_env.py:331).  The reference calls np.log, whose
// last bit is platform dependent; what the contract fixes is the VALUE to 1e-5.  log_r is a
// table-driven fp64 log within 2 ulp of libm, or 2e-18 absolute where the two table terms cancel
// (checked in tests/test_row_logic_host.py):
//   x = 2^e * m,  m in [sqrt(1/2), sqrt(2));  the top 7 mantissa bits of m pick (rc, lc) with
//   rc ~ 1/centre of the interval and lc = -log(rc) (computed in long double on the host);
//   r = fma(m, rc, -1)  (|r| <= 2^-8, single rounding);  log x = e ln2 + lc + log1p(r),
//   log1p by its Taylor series through r^7 (truncation 2^-67).  The interval that contains 1 has
//   rc = 1, lc = 0, so results near log(1) = 0 keep full RELATIVE accuracy.
// Anything that is not a positive normal number (0, denormal, negative, inf, NaN) takes libm's log.
struct LogTabEntry { double rc, lc; };
constexpr int kLogTabSize = 128;
constexpr int kLogTabOffset = 0x3ff00000 - 0x3fe6a09e;   // high-word shift that centres m on 1

DS_HD long long double_bits(double a)
{
#if DS_DEVICE_CODE
    return __double_as_longlong(a);
#else
    long long b; memcpy(&b, &a, sizeof b); return b;
#endif
}
DS_HD double bits_double(long long b)
{
#if DS_DEVICE_CODE
    return __longlong_as_double(b);
#else
    double a; memcpy(&a, &b, sizeof a); return a;
#endif
}
// (out of line: libm's log is ~200 instructions that the callers' hot loops should not carry)
#if DS_DEVICE_CODE
__device__ __noinline__ double log_unusual(double x) { return log(x); }
#else
inline double log_unusual(double x) { return log(x); }
#endif
DS_HD double log_r(double x, const LogTabEntry *__restrict__ tab)
{
    const long long b = double_bits(x);
    const int hi = (int)(b >> 32);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log_unusual(x);
    const int hs = hi + kLogTabOffset;
    const int e = (hs >> 20) - 1023;
    const LogTabEntry t = tab[(hs >> 13) & (kLogTabSize - 1)];
    const double m = bits_double(b - ((long long)e << 52));
    const double r = fma_rn(m, t.rc, -1.0);
    double p = fma_rn(r, 1.0 / 7.0, -1.0 / 6.0);
    p = fma_rn(r, p, 1.0 / 5.0);
    p = fma_rn(r, p, -1.0 / 4.0);
    p = fma_rn(r, p, 1.0 / 3.0);
    p = fma_rn(r, p, -0.5);
    const double sres = fma_rn(mul_rn(r, r), p, r);
    return fma_rn((double)e, 0.6931471805599453094, add_rn(t.lc, sres));
}
// Host-side table fill (also used by tests/rowcheck): entry idx covers the doubles whose shifted
// high word has mantissa field [idx << 13, (idx + 1) << 13).
inline void fill_log_table(LogTabEntry *tab)
{
    for (int idx = 0; idx < kLogTabSize; ++idx) {
        const long long a = (long long)(0x3ff00000 + (idx << 13) - kLogTabOffset) << 32;
        double lo, hi;
        const long long a2 = a + ((long long)0x2000 << 32);
        memcpy(&lo, &a, sizeof lo); memcpy(&hi, &a2, sizeof hi);
        if (lo <= 1.0 && 1.0 < hi) { tab[idx].rc = 1.0; tab[idx].lc = 0.0; continue; }
        const double rc = 1.0 / (0.5 * (lo + hi));
        tab[idx].rc = rc;
        tab[idx].lc = (double)(-logl((long double)rc));
    }
}
DS_HD float add_rn(float a, float b)
{
#if DS_DEVICE_CODE
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
DS_HD float sub_rn(float a, float b)
{
#if DS_DEVICE_CODE
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
DS_HD float mul_rn(float a, float b)
{
#if DS_DEVICE_CODE
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
DS_HD float fma_rn(float a, float b, float c)
{
#if DS_DEVICE_CODE
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
DS_HD float div_rn(float a, float b)
{
#if DS_DEVICE_CODE
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
DS_HD float sqrt_rn(float a)
{
#if DS_DEVICE_CODE
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}
DS_HD float log_r(float a, const LogTabEntry *) { return logf(a); }

DS_HD int lowest_bit(unsigned m)   // index of the lowest set bit, m != 0
{
#if DS_DEVICE_CODE
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

template <typename Real> struct vec2_of;
template <> struct vec2_of<double> { using type = double2; };
template <> struct vec2_of<float> { using type = float2; };

template <typename Real> DS_HD Real real_inf();
template <> DS_HD double real_inf<double>() { return (double)INFINITY; }
template <> DS_HD float real_inf<float>() { return INFINITY; }

// np.nan_to_num (drone_env.py:287-288)
DS_HD double nan_to_num(double v)
{
    if (v != v) return 0.0;
    if (v > 1.7976931348623157e308) return 1.7976931348623157e308;
    if (v < -1.7976931348623157e308) return -1.7976931348623157e308;
    return v;
}
DS_HD float nan_to_num(float v)
{
    if (v != v) return 0.0f;
    if (v > 3.4028234663852886e38f) return 3.4028234663852886e38f;
    if (v < -3.4028234663852886e38f) return -3.4028234663852886e38f;
    return v;
}

// ---------------------------------------------------------------- arguments
struct Consts {              // device arrays, length n (xF: 2n); Real typed unless noted
    const void *xF, *d_safety, *delta, *radius, *log_ds;
    const void *thr2;        // Real: pairs with |xi-xj|^2 >= thr2[i] are provably clipped to d_safety[i]
    const int *clipcnt;      // int: #{j != i : d_safety[i] <= delta[j]}  (Delta-disk count of clipped pairs)
    const LogTabEntry *logtab;   // kLogTabSize entries (log_r)
};

struct StepArgs {
    int E, n, k, simplify, G, do_integrate, log_mode, max_steps;
    int ctrl;                // 0: actions given; 1 / 2: the reference's proportional / gradient controller
    double u_max;            // gradient controller's clip (drone_env.py:612)
    void *ctrl_out;          // non-null: only compute the controller's actions -> Real [E][n][2], no step
    Consts c;
    double dt, q, b, goal_tol, sentinel, zero_eps, ghost;
    const void *act;         // Real [E][n][2]
    void *pos, *vel, *r, *tr, *z;
    int *Ni, *ncoll;
    uint8_t *fin;
    int *t;
};

struct RolloutArgs {
    StepArgs s;
    int T, TC, n_actions;
    int L;                   // capacity of the CTA's near-pair work list
    unsigned mulA, mulN;     // ceil(2^32 / (G n)), ceil(2^32 / n) (0 when the divisor is 1): tid -> (slice, env, agent)
    int inline_rows;         // near pairs: 0 one CTA-wide work list, 1 rows evaluate their own (n <= 32), 2 warp-local lists
    const void *actions;     // Real [T][E][n][2] or null
    const uint8_t *aidx;     // u8 [T][E][n]
    const void *atable;      // Real [n_actions][2]
    void *pos_tr, *vel_tr, *r_tr, *tr_tr, *z_tr;
    int *Ni_tr, *ncoll_tr;
    uint8_t *fin_tr;
    double *agg;
    uint8_t *done;
};

template <typename Real> struct ParamsR {
    Real dt, q, b, goal_tol, sentinel, zero_eps, ghost;
    int log_mode, simplify, k;
    DS_HD explicit ParamsR(const StepArgs &a)
        : dt((Real)a.dt), q((Real)a.q), b((Real)a.b), goal_tol((Real)a.goal_tol),
          sentinel((Real)a.sentinel), zero_eps((Real)a.zero_eps), ghost((Real)a.ghost),
          log_mode(a.log_mode), simplify(a.simplify), k(a.k) {}
};

template <typename Real> struct AgentConst {   // per-thread (row i) constants
    Real xF, yF, ds, delta, radius, log_ds, thr2;
    int clipcnt;
};

template <typename Real>
DS_HD AgentConst<Real> load_agent_const(const Consts &c, int i)
{
    AgentConst<Real> a;
    a.xF = ((const Real *)c.xF)[2 * i];
    a.yF = ((const Real *)c.xF)[2 * i + 1];
    a.ds = ((const Real *)c.d_safety)[i];
    a.delta = ((const Real *)c.delta)[i];
    a.radius = ((const Real *)c.radius)[i];
    a.log_ds = ((const Real *)c.log_ds)[i];
    a.thr2 = ((const Real *)c.thr2)[i];
    a.clipcnt = c.clipcnt[i];
    return a;
}

// ---------------------------------------------------------------- one pair (i, j), j != i
// distance_data() for one entry of the pair matrix (drone_env.py:314-332).  A pair whose d_ij is
// exactly d_safety[i] is CLIPPED: d_ij_norm = 1, log = 0, no collision -- logd stays 0.
template <typename Real> struct PairOut {
    Real d, logd;
    bool in_disk, coll;
};

template <typename Real>
DS_HD void eval_pair(PairOut<Real> &po, Real xi, Real yi, Real xj, Real yj, Real ds_i, Real rad_i,
                     Real rad_j, Real delta_j, Real log_ds_i, const ParamsR<Real> &P,
                     const LogTabEntry *__restrict__ tab)
{
    const Real dx = sub_rn(xi, xj), dy = sub_rn(yi, yj);
    const Real dist = sqrt_rn(fma_rn(dy, dy, mul_rn(dx, dx)));     // :318 (BLAS ddot)
    const Real raw = sub_rn(sub_rn(dist, rad_i), rad_j);           // :318
    Real d = (ds_i < raw) ? ds_i : raw;                            // python min(raw, d_safety[i])
    d = (d == (Real)0) ? P.zero_eps : d;                           // :319-320
    po.d = d;
    po.in_disk = d <= delta_j;                                     // :328 deltas[j]
    // not clipped: d_ij_norm != 1, the barrier term is live (:321,327,330-332).  Straight-line:
    // near pairs are almost never clipped, so the log is evaluated for all of them (on 1 where its
    // value is not used) and selected.
    const bool live = d != ds_i;
    bool coll;
    Real lg;
    if (P.log_mode == 0) {
        const Real dn = div_rn(ds_i, d);
        coll = dn <= (Real)0;
        lg = log_r((live && !coll) ? dn : (Real)1, tab);
    } else if (P.log_mode == 1) {
        coll = (ds_i > (Real)0) ? (d < (Real)0) : (ds_i == (Real)0);
        lg = sub_rn(log_ds_i, log_r((live && !coll) ? fabs(d) : fabs(ds_i), tab));
    } else {
        // DS_LOG_RCP: log(d_safety / d) = -log(d * (1 / d_safety)); the reciprocal is rounded once
        // (the rollout kernels get it precomputed; same value)
        coll = (ds_i > (Real)0) ? (d < (Real)0) : (ds_i == (Real)0);
        lg = -log_r((live && !coll) ? mul_rn(d, div_rn((Real)1, ds_i)) : (Real)1, tab);
    }
    po.coll = live && coll;
    po.logd = live ? (coll ? P.sentinel : lg) : (Real)0;
}

// Near-pair work-list entry (rollout kernel).  Before evaluation: row | j << 10 | i << 20.
// After evaluation the same word holds j and the flags of the pair; (d, logd) sit in a parallel
// array.  adj = [in_disk] - [d_safety[i] <= delta[j]] + 1 corrects the Delta-disk count, which
// starts from "every pair is clipped".
constexpr unsigned kEntSkip = 0xffffffffu;                         // slot of a row that overflowed
DS_HD unsigned pack_entry(int row, int j, int i) { return (unsigned)row | ((unsigned)j << 10) | ((unsigned)i << 20); }
DS_HD unsigned pack_result(int j, bool in_disk, bool coll, int adj1)
{
    return (unsigned)j | (in_disk ? 1u << 10 : 0u) | (coll ? 1u << 11 : 0u) | ((unsigned)adj1 << 12);
}

// ---------------------------------------------------------------- one row of the pair matrix
// Everything rewards() derives for agent i (drone_env.py:260-293) from the staged
// positions of its frame.
template <typename Real, int K> struct RowResult {
    static constexpr int CAP = (K >= 0 ? K : kMaxK) + 1;
    Real r, tr;          // localized / global reward
    Real zx, zy;         // x_i - xF_i
    int ncoll;           // collisions in this row
    int in_range;        // agents inside the Delta disk, minus itself (:346)
    bool at_goal;        // ||xF_i - x_i|| <= goal_tol (:249-251)
    Real td[CAP];        // k+1 smallest d_ij, ascending, ties -> lowest index
    int tj[CAP];
};

// Offer (d, j) to the k+1 smallest kept in (d, j) lexicographic order -- the order
// a stable argsort of row i produces (np.argsort row, :338).  Returns false when
// the candidate does not make the list.
template <typename Real, int K>
DS_HD bool topk_offer(RowResult<Real, K> &o, int kk, Real d, int j)
{
    constexpr int CAP = RowResult<Real, K>::CAP;
    if (!(d < o.td[kk] || (d == o.td[kk] && j < o.tj[kk]))) return false;
    o.td[kk] = d; o.tj[kk] = j;
#pragma unroll
    for (int m = CAP - 1; m > 0; --m) {
        if (m <= kk) {
            const bool lt = o.td[m] < o.td[m - 1] || (o.td[m] == o.td[m - 1] && o.tj[m] < o.tj[m - 1]);
            if (lt) {
                const Real tdv = o.td[m]; o.td[m] = o.td[m - 1]; o.td[m - 1] = tdv;
                const int tjv = o.tj[m]; o.tj[m] = o.tj[m - 1]; o.tj[m - 1] = tjv;
            }
        }
    }
    return true;
}

// Running state of a row while its pairs are folded in ascending j.
template <typename Real> struct RowAcc {
    Real sum_local, sum_all;   // :282, :283
    int ncoll, cnt_nd;
    bool self_unclipped;
};

// j == i (:323-325): dist = 0, d_ii = min(-2 l_i, d_safety[i]), d_norm = 1 -> no barrier term.
template <typename Real, int K>
DS_HD void row_begin(RowResult<Real, K> &o, RowAcc<Real> &acc, int kk, int i, const AgentConst<Real> &c)
{
    constexpr int CAP = RowResult<Real, K>::CAP;
#pragma unroll
    for (int m = 0; m < CAP; ++m) { o.td[m] = real_inf<Real>(); o.tj[m] = 0; }
    acc.sum_local = 0; acc.sum_all = 0; acc.ncoll = 0;
    const Real raw_ii = sub_rn(sub_rn((Real)0, c.radius), c.radius);
    const Real d_ii = (c.ds < raw_ii) ? c.ds : raw_ii;
    acc.cnt_nd = c.clipcnt + ((d_ii <= c.delta) ? 1 : 0);               // :328 deltas[j], j == i
    acc.self_unclipped = (d_ii != c.ds);
    if (d_ii < c.ds) topk_offer<Real, K>(o, kk, d_ii, i);
}

// Fold one evaluated near pair (ascending j: summation order of :282-283).  Returns true when the
// pair is NOT clipped (d_ij != d_safety[i]).
template <typename Real, int K>
DS_HD bool row_fold(RowResult<Real, K> &o, RowAcc<Real> &acc, int kk, const AgentConst<Real> &c, int j,
                    Real d, Real logd, bool in_disk, bool coll, int adj)
{
    acc.cnt_nd += adj;                                                  // replaces the clipped-pair count
    if (d != c.ds) {
        acc.ncoll += coll ? 1 : 0;
        acc.sum_all = add_rn(acc.sum_all, logd);                                              // :283
        acc.sum_local = add_rn(acc.sum_local, mul_rn(logd, in_disk ? (Real)1 : (Real)0));     // :282
        if (d < c.ds) topk_offer<Real, K>(o, kk, d, j);
        return true;
    }
    return false;
}

// Goal cost, rewards, Delta-disk count and termination flag of the row (:249-251,272-288,346).
template <typename Real, int K>
DS_HD void row_end(RowResult<Real, K> &o, const RowAcc<Real> &acc, Real xi, Real yi, const AgentConst<Real> &c,
                   const ParamsR<Real> &P)
{
    const Real gx = sub_rn(c.xF, xi), gy = sub_rn(c.yF, yi);
    const Real nrm = sqrt_rn(add_rn(mul_rn(gx, gx), mul_rn(gy, gy)));     // :249,276 (axis norm, unfused)
    const Real goal = mul_rn(P.q, mul_rn(nrm, nrm));                      // :276
    o.r = -nan_to_num(add_rn(goal, mul_rn(P.b, acc.sum_local)));          // :282,287
    o.tr = -nan_to_num(add_rn(goal, mul_rn(P.b, acc.sum_all)));           // :283,288
    o.zx = -gx; o.zy = -gy;                                               // :357
    o.ncoll = acc.ncoll;
    o.in_range = acc.cnt_nd - 1;                                          // :346
    o.at_goal = nrm <= P.goal_tol;
}

// Whole row in one thread: both passes inline (step kernel; rollout rows whose near pairs did not
// fit the CTA's work list).
template <typename Real, int K>
DS_HD void eval_row(RowResult<Real, K> &o, int n, int i, Real xi, Real yi,
                    const AgentConst<Real> &c,
                    const typename vec2_of<Real>::type *__restrict__ s_pos,
                    const Real *__restrict__ s_delta,
                    const Real *__restrict__ s_radius,
                    const ParamsR<Real> &P, const LogTabEntry *__restrict__ tab, int cs = 1)
{
    // s_delta / s_radius: per-agent constants with element stride cs (1: plain arrays; 2: the
    // rollout kernel's packed (radius, delta) pairs)
    using V2 = typename vec2_of<Real>::type;
    const int kk = (K >= 0) ? K : P.k;
    RowAcc<Real> acc;
    row_begin<Real, K>(o, acc, kk, i, c);
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int jn = (n - j0 < 32) ? (n - j0) : 32;
        // pass 1: which agents of this block of 32 are NOT provably clipped
        unsigned near = 0;
#pragma unroll 4
        for (int jj = 0; jj < jn; ++jj) {
            const V2 pj = s_pos[j0 + jj];
            const Real dx = sub_rn(xi, pj.x), dy = sub_rn(yi, pj.y);
            const Real d2 = fma_rn(dy, dy, mul_rn(dx, dx));
            near |= ((d2 >= c.thr2) ? 0u : 1u) << jj;                   // NaN -> near (exact path)
        }
        const unsigned selfbit = ((unsigned)(i - j0) < 32u) ? (1u << (i - j0)) : 0u;
        near &= ~selfbit;
        unsigned unclipped = acc.self_unclipped ? selfbit : 0u;         // agents with d_ij != d_safety[i]
        // pass 2: exact evaluation of the near pairs, ascending j
        while (near) {
            const int jj = lowest_bit(near);
            near &= near - 1;
            const int j = j0 + jj;
            const V2 pj = s_pos[j];
            const Real dl = s_delta[j * cs];
            PairOut<Real> po;
            eval_pair<Real>(po, xi, yi, pj.x, pj.y, c.ds, c.radius, s_radius[j * cs], dl, c.log_ds, P, tab);
            const int adj = (po.in_disk ? 1 : 0) - ((c.ds <= dl) ? 1 : 0);
            if (row_fold<Real, K>(o, acc, kk, c, j, po.d, po.logd, po.in_disk, po.coll, adj))
                unclipped |= 1u << jj;
        }
        // clipped agents tie at exactly d_safety[i]: offered in index order until one is refused
        unsigned cm = ((jn == 32) ? 0xffffffffu : ((1u << jn) - 1u)) & ~unclipped;
        while (cm) {
            const int jj = lowest_bit(cm);
            if (!topk_offer<Real, K>(o, kk, c.ds, j0 + jj)) break;
            cm &= cm - 1;
        }
    }
    row_end<Real, K>(o, acc, xi, yi, c, P);
}

// k-nearest selection for the work-list / near-mask paths.  Neighbour candidates (d < d_safety[i])
// arrive in ASCENDING j, so a strict "d < slot" insertion keeps equal distances in index order --
// the stable argsort order (np.argsort row, :338) -- with one compare per slot.  Only the row's own
// entry (j = i) is out of order; it is merged afterwards with the full (d, j) rule.  CAP = k + 1
// neighbours are tracked because self may fall outside the first k + 1 (coincident agents, or a
// larger neighbour radius).
template <typename Real, int K>
DS_HD void topk_insert_sorted(RowResult<Real, K> &o, Real d, int j)
{
    constexpr int CAP = RowResult<Real, K>::CAP;
    bool lt[CAP];
#pragma unroll
    for (int m = 0; m < CAP; ++m) lt[m] = d < o.td[m];             // false for d = +inf / NaN
#pragma unroll
    for (int m = CAP - 1; m > 0; --m) {
        o.td[m] = lt[m - 1] ? o.td[m - 1] : (lt[m] ? d : o.td[m]);
        o.tj[m] = lt[m - 1] ? o.tj[m - 1] : (lt[m] ? j : o.tj[m]);
    }
    o.td[0] = lt[0] ? d : o.td[0];
    o.tj[0] = lt[0] ? j : o.tj[0];
}

// Merge the row's own entry (d_ii, i) into the sorted neighbours.
template <typename Real, int K>
DS_HD void topk_merge_self(RowResult<Real, K> &o, Real d_ii, int i)
{
    constexpr int CAP = RowResult<Real, K>::CAP;
    bool lt[CAP];
#pragma unroll
    for (int m = 0; m < CAP; ++m) lt[m] = d_ii < o.td[m] || (d_ii == o.td[m] && i < o.tj[m]);
#pragma unroll
    for (int m = CAP - 1; m > 0; --m) {
        o.td[m] = lt[m - 1] ? o.td[m - 1] : (lt[m] ? d_ii : o.td[m]);
        o.tj[m] = lt[m - 1] ? o.tj[m - 1] : (lt[m] ? i : o.tj[m]);
    }
    o.td[0] = lt[0] ? d_ii : o.td[0];
    o.tj[0] = lt[0] ? i : o.tj[0];
}

// Free slots <- lowest clipped indices (all at exactly d_safety[i]: index order).  nreal = real
// candidates found (self included); cm = clipped agents j < 32; scan(hi) tells whether agent
// hi >= 32 is unclipped (rare path: >= 32 - k of the first 32 agents unclipped).
template <typename Real, int K, typename ScanF>
DS_HD void topk_fill_clipped(RowResult<Real, K> &o, int kk, int n, int nreal, unsigned cm, Real ds, ScanF scan)
{
    constexpr int CAP = RowResult<Real, K>::CAP;
    if (n < 32) cm &= (1u << n) - 1u;
    int hi = 32;
#pragma unroll
    for (int m = 0; m < CAP; ++m) {
        if (m <= kk && m >= nreal) {
            int cj = n;
            if (cm) {
                cj = lowest_bit(cm);
                cm &= cm - 1;
            } else {
                for (; hi < n && cj == n; ++hi)
                    if (!scan(hi)) cj = hi;
            }
            if (cj < n) { o.td[m] = ds; o.tj[m] = cj; }
        }
    }
}

// The same row from its segment of the evaluated work list (rollout kernel): ent[q] / res[q],
// q < cnt, ascending j.
template <typename Real, int K>
DS_HD void eval_row_from_list(RowResult<Real, K> &o, int n, int i, Real xi, Real yi,
                              const AgentConst<Real> &c, const unsigned *__restrict__ ent,
                              const typename vec2_of<Real>::type *__restrict__ res, int cnt,
                              const ParamsR<Real> &P)
{
    using V2 = typename vec2_of<Real>::type;
    constexpr int CAP = RowResult<Real, K>::CAP;
    const int kk = (K >= 0) ? K : P.k;
#pragma unroll
    for (int m = 0; m < CAP; ++m) { o.td[m] = real_inf<Real>(); o.tj[m] = 0; }
    RowAcc<Real> acc;
    acc.sum_local = 0; acc.sum_all = 0; acc.ncoll = 0;
    int nreal = 0;
    // j == i (:323-325): dist = 0, d_ii = min(-2 l_i, d_safety[i]), d_norm = 1 -> no barrier term
    const Real raw_ii = sub_rn(sub_rn((Real)0, c.radius), c.radius);
    const Real d_ii = (c.ds < raw_ii) ? c.ds : raw_ii;
    acc.cnt_nd = c.clipcnt + ((d_ii <= c.delta) ? 1 : 0);               // :328 deltas[j], j == i
    acc.self_unclipped = (d_ii != c.ds);
    unsigned unclipped_lo = (acc.self_unclipped && i < 32) ? (1u << i) : 0u;   // agents j < 32 only
    for (int q = 0; q < cnt; ++q) {
        const unsigned w = ent[q];
        const V2 dv = res[q];
        const int j = (int)(w & 1023u);
        acc.cnt_nd += (int)((w >> 12) & 3u) - 1;                        // replaces the clipped-pair count
        // a clipped pair carries logd = +0 and no collision flag: the sums take it unconditionally
        const bool live = dv.x != c.ds;                                 // not clipped (NaN included)
        unclipped_lo |= (live && j < 32) ? (1u << j) : 0u;
        acc.ncoll += (int)((w >> 11) & 1u);
        acc.sum_all = add_rn(acc.sum_all, dv.y);                                                    // :283
        acc.sum_local = add_rn(acc.sum_local, mul_rn(dv.y, ((w >> 10) & 1u) ? (Real)1 : (Real)0));  // :282
        const bool cand = dv.x < c.ds;                                  // real candidate for the k nearest
        topk_insert_sorted<Real, K>(o, cand ? dv.x : real_inf<Real>(), j);
        nreal += cand ? 1 : 0;
    }
    if (d_ii < c.ds) { topk_merge_self<Real, K>(o, d_ii, i); ++nreal; }
    topk_fill_clipped<Real, K>(o, kk, n, nreal, ~unclipped_lo, c.ds, [&](int hi) {
        bool unclipped = (hi == i) && acc.self_unclipped;
        for (int q = 0; q < cnt; ++q)
            if ((int)(ent[q] & 1023u) == hi && res[q].x != c.ds) unclipped = true;
        return unclipped;
    });
    row_end<Real, K>(o, acc, xi, yi, c, P);
}

// The same row when n <= 32 and the row evaluates its own near pairs (dense small-n frames, where
// a work list costs more than the imbalance it removes): `near` = pass-1 mask of the row.
// cpair[j] = (radius_j, delta_j).
template <typename Real, int K>
DS_HD void eval_row_near32(RowResult<Real, K> &o, int n, int i, Real xi, Real yi, const AgentConst<Real> &c,
                           unsigned near, const typename vec2_of<Real>::type *__restrict__ s_pos,
                           const typename vec2_of<Real>::type *__restrict__ cpair, const ParamsR<Real> &P,
                           const LogTabEntry *__restrict__ tab)
{
    using V2 = typename vec2_of<Real>::type;
    constexpr int CAP = RowResult<Real, K>::CAP;
    const int kk = (K >= 0) ? K : P.k;
#pragma unroll
    for (int m = 0; m < CAP; ++m) { o.td[m] = real_inf<Real>(); o.tj[m] = 0; }
    RowAcc<Real> acc;
    acc.sum_local = 0; acc.sum_all = 0; acc.ncoll = 0;
    int nreal = 0;
    const Real raw_ii = sub_rn(sub_rn((Real)0, c.radius), c.radius);   // j == i (:323-325)
    const Real d_ii = (c.ds < raw_ii) ? c.ds : raw_ii;
    acc.cnt_nd = c.clipcnt + ((d_ii <= c.delta) ? 1 : 0);
    acc.self_unclipped = (d_ii != c.ds);
    unsigned unclipped = acc.self_unclipped ? (1u << i) : 0u;
    while (near) {                                                      // ascending j (:282-283)
        const int j = lowest_bit(near);
        near &= near - 1;
        const V2 pj = s_pos[j], cj = cpair[j];
        PairOut<Real> po;
        eval_pair<Real>(po, xi, yi, pj.x, pj.y, c.ds, c.radius, cj.x, cj.y, c.log_ds, P, tab);
        acc.cnt_nd += (po.in_disk ? 1 : 0) - ((c.ds <= cj.y) ? 1 : 0);
        unclipped |= (po.d != c.ds) ? (1u << j) : 0u;
        acc.ncoll += po.coll ? 1 : 0;
        acc.sum_all = add_rn(acc.sum_all, po.logd);
        acc.sum_local = add_rn(acc.sum_local, mul_rn(po.logd, po.in_disk ? (Real)1 : (Real)0));
        const bool cand = po.d < c.ds;
        topk_insert_sorted<Real, K>(o, cand ? po.d : real_inf<Real>(), j);
        nreal += cand ? 1 : 0;
    }
    if (d_ii < c.ds) { topk_merge_self<Real, K>(o, d_ii, i); ++nreal; }
    topk_fill_clipped<Real, K>(o, kk, n, nreal, ~unclipped, c.ds, [](int) { return true; });
    row_end<Real, K>(o, acc, xi, yi, c, P);
}

// Baseline controllers of the reference for agent i of one frame:
//   mode 1  proportional_control (drone_env.py:655-679): u = xF - x, norm capped at 1;
//   mode 2  gradient_control (drone_env.py:612-653): u = clip(-(2 (x - xF) - 0.1 sum_j push_ij), +-u_max),
//           push_ij = (x_i - x_j) / (d_ij ||x_i - x_j||) over j != i with d_ij = ||x_i - x_j|| - l_i - l_j
//           <= d_safety[i] (global knowledge; no zero rule: d_ij = 0 divides by zero like the reference).
// Same operation order as the reference (np.linalg.norm of a 1-D vector = sqrt(ddot) = sqrt(fma)).
template <typename Real>
DS_HD void control_action(int mode, int n, int i, Real xi, Real yi, Real xF, Real yF, Real ds_i, Real rad_i,
                          const typename vec2_of<Real>::type *__restrict__ s_pos,
                          const Real *__restrict__ s_radius, int cs, Real u_max, Real &ux, Real &uy)
{
    using V2 = typename vec2_of<Real>::type;
    if (mode == 1) {
        ux = sub_rn(xF, xi); uy = sub_rn(yF, yi);                                  // :669-671, k_gain = 1
        const Real nrm = sqrt_rn(fma_rn(uy, uy, mul_rn(ux, ux)));                  // :673
        if (nrm > (Real)1) { ux = div_rn(ux, nrm); uy = div_rn(uy, nrm); }         // :674-676, u_max = 1
        return;
    }
    const Real t1x = mul_rn((Real)2, sub_rn(xi, xF)), t1y = mul_rn((Real)2, sub_rn(yi, yF));   // :636
    Real t2x = 0, t2y = 0;
    for (int j = 0; j < n; ++j) {
        if (j == i) continue;
        const V2 pj = s_pos[j];
        const Real dx = sub_rn(xi, pj.x), dy = sub_rn(yi, pj.y);
        const Real nrm = sqrt_rn(fma_rn(dy, dy, mul_rn(dx, dx)));
        const Real dij = sub_rn(sub_rn(nrm, rad_i), s_radius[j * cs]);             // :644
        if (dij <= ds_i) {                                                         // :646
            const Real den = mul_rn(dij, nrm);                                     // :647
            t2x = add_rn(t2x, div_rn(dx, den)); t2y = add_rn(t2y, div_rn(dy, den));
        }
    }
    const Real gx = sub_rn(t1x, mul_rn((Real)0.1, t2x)), gy = sub_rn(t1y, mul_rn((Real)0.1, t2y));   // :649
    ux = -gx; uy = -gy;                                                            // :650 np.clip
    ux = ux < -u_max ? -u_max : (ux > u_max ? u_max : ux);
    uy = uy < -u_max ? -u_max : (uy > u_max ? u_max : uy);
}

// Write z_i (k+1 rows) and Ni_i (drone_env.py:344-397) for global agent index g.  The in-range
// neighbours occupy slots 1..min(in_range, k) in order, so Ni[kth] is either tj[kth] or -1.
template <typename Real, int K>
DS_HD void write_obs(const RowResult<Real, K> &o, int i, Real xi, Real yi,
                     const AgentConst<Real> &c,
                     const typename vec2_of<Real>::type *__restrict__ s_pos,
                     const typename vec2_of<Real>::type *__restrict__ s_vel,
                     const Real *__restrict__ s_radius,
                     const ParamsR<Real> &P, Real *__restrict__ z, int *__restrict__ Ni,
                     size_t g, int cs = 1)
{
    using V2 = typename vec2_of<Real>::type;
    const int kk = (K >= 0) ? K : P.k;
    constexpr int CAP = RowResult<Real, K>::CAP;
    int *nl = Ni + g * (size_t)(kk + 1);
    nl[0] = i;
    Real ghx = 0, ghy = 0;
    if (o.in_range < kk) {                                            // ghost rows (:383-386)
        const Real zn = sqrt_rn(fma_rn(o.zy, o.zy, mul_rn(o.zx, o.zx)));
        ghx = mul_rn(mul_rn(div_rn(o.zx, zn), c.delta), P.ghost);
        ghy = mul_rn(mul_rn(div_rn(o.zy, zn), c.delta), P.ghost);
    }
    if (P.simplify) {
        V2 *zr = reinterpret_cast<V2 *>(z + g * (size_t)(kk + 1) * 2);
        V2 v; v.x = o.zx; v.y = o.zy;
        zr[0] = v;
#pragma unroll
        for (int kth = 1; kth < CAP; ++kth) {
            if (kth <= kk) {
                const int j = o.tj[kth];
                const bool in_r = kth <= o.in_range;                  // :362-368
                const V2 pj = s_pos[j];
                v.x = in_r ? sub_rn(pj.x, xi) : ghx;
                v.y = in_r ? sub_rn(pj.y, yi) : ghy;
                zr[kth] = v;
                nl[kth] = in_r ? j : -1;
            }
        }
    } else {
        Real *zr = z + g * (size_t)(kk + 1) * 5;
        const V2 vi = s_vel[i];
        zr[0] = o.zx; zr[1] = o.zy; zr[2] = vi.x; zr[3] = vi.y; zr[4] = c.radius;   // :355-357
#pragma unroll
        for (int kth = 1; kth < CAP; ++kth) {
            if (kth <= kk) {
                const int j = o.tj[kth];
                const bool in_r = kth <= o.in_range;
                const V2 pj = s_pos[j];
                const V2 vj = s_vel[j];                               // zj = state[j,:].copy() (:367,385)
                Real *row = zr + kth * 5;
                row[0] = in_r ? sub_rn(pj.x, xi) : ghx;
                row[1] = in_r ? sub_rn(pj.y, yi) : ghy;
                row[2] = vj.x; row[3] = vj.y; row[4] = s_radius[j * cs];
                nl[kth] = in_r ? j : -1;
            }
        }
    }
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------- shared-memory carve-up
// CTA-wide: delta[n], radius[n].  Per item (TC*G*n): act/vel, pos, r, tr.  Per agent of a
// slice (G*n): chunk start position, last executed velocity.  Per frame (TC*G): collision
// count, not-at-goal flag, per-frame means.  Per environment (G): alive, t, executed slices.
template <typename Real> struct CtaSmem {
    using V2 = typename vec2_of<Real>::type;
    Real *delta, *radius;
    V2 *act, *pos, *p0, *vfin;
    Real *r, *tr;
    double *mr, *mtr;
    int *cnt, *notgoal, *mc, *alive, *tenv, *nexec;
    __host__ __device__ static size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }
    __host__ __device__ static size_t bytes(int n, int G, int TC)
    {
        const size_t A = (size_t)G * n, I = A * TC, F = (size_t)G * TC;
        return align16(2 * n * sizeof(Real)) + 2 * I * sizeof(V2) + 2 * A * sizeof(V2) +
               align16(2 * I * sizeof(Real)) + 2 * F * sizeof(double) + align16(3 * F * sizeof(int)) +
               align16(3 * (size_t)G * sizeof(int));
    }
    __device__ CtaSmem(unsigned char *base, int n, int G, int TC)
    {
        const size_t A = (size_t)G * n, I = A * TC, F = (size_t)G * TC;
        unsigned char *p = base;
        delta = reinterpret_cast<Real *>(p); radius = delta + n; p += align16(2 * n * sizeof(Real));
        act = reinterpret_cast<V2 *>(p); pos = act + I; p0 = pos + I; vfin = p0 + A;
        p += 2 * I * sizeof(V2) + 2 * A * sizeof(V2);
        r = reinterpret_cast<Real *>(p); tr = r + I; p += align16(2 * I * sizeof(Real));
        mr = reinterpret_cast<double *>(p); mtr = mr + F; p += 2 * F * sizeof(double);
        cnt = reinterpret_cast<int *>(p); notgoal = cnt + F; mc = notgoal + F; p += align16(3 * F * sizeof(int));
        alive = reinterpret_cast<int *>(p); tenv = alive + G; nexec = tenv + G;
    }
};

// ---------------------------------------------------------------- step kernel
// drones.step() / rewards() for E environments, one launch; CTA = G environments.
template <typename Real, int K, int NT>
__global__ void __launch_bounds__(NT)
step_kernel(const StepArgs a)
{
    using V2 = typename vec2_of<Real>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, G = a.G;
    CtaSmem<Real> sm(smem_raw, n, G, 1);
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        sm.delta[idx] = ((const Real *)a.c.delta)[idx];
        sm.radius[idx] = ((const Real *)a.c.radius)[idx];
    }
    const ParamsR<Real> P(a);
    const int lt = threadIdx.x;
    const int le = lt / n, i = lt - le * n;
    const long long e = (long long)blockIdx.x * G + le;
    const bool active = (lt < G * n) && (e < a.E);
    const size_t g = active ? (size_t)e * n + i : 0;

    Real xi = 0, yi = 0;
    AgentConst<Real> c{};
    V2 u_ctrl{};
    if (a.ctrl) {
        // closed loop: the action is a function of the CURRENT positions of the whole environment
        if (active) sm.pos[lt] = reinterpret_cast<const V2 *>(a.pos)[g];
        __syncthreads();
        if (active) {
            c = load_agent_const<Real>(a.c, i);
            const V2 p = sm.pos[lt];
            control_action<Real>(a.ctrl, n, i, p.x, p.y, c.xF, c.yF, c.ds, c.radius, sm.pos + le * n, sm.radius, 1,
                                 (Real)a.u_max, u_ctrl.x, u_ctrl.y);
        }
        if (a.ctrl_out) {                                  // ds_control: the actions are the result
            if (active) reinterpret_cast<V2 *>(a.ctrl_out)[g] = u_ctrl;
            return;
        }
        __syncthreads();                                   // everybody has read the old positions
    }
    if (active) {
        c = load_agent_const<Real>(a.c, i);
        V2 p = reinterpret_cast<const V2 *>(a.pos)[g];
        V2 v;
        if (a.do_integrate) {
            const V2 u = a.ctrl ? u_ctrl : reinterpret_cast<const V2 *>(a.act)[g];
            p.x = add_rn(p.x, mul_rn(P.dt, u.x));     // A = I, B = dt I (:78-79,235)
            p.y = add_rn(p.y, mul_rn(P.dt, u.y));
            v = u;                                     // :238
            reinterpret_cast<V2 *>(a.pos)[g] = p;
            reinterpret_cast<V2 *>(a.vel)[g] = v;
        } else {
            v = reinterpret_cast<const V2 *>(a.vel)[g];
        }
        xi = p.x; yi = p.y;
        sm.pos[lt] = p;
        sm.act[lt] = v;
        if (i == 0) { sm.cnt[le] = 0; sm.notgoal[le] = 0; }
    }
    __syncthreads();
    if (active) {
        RowResult<Real, K> o;
        eval_row<Real, K>(o, n, i, xi, yi, c, sm.pos + le * n, sm.delta, sm.radius, P, a.c.logtab);
        reinterpret_cast<Real *>(a.r)[g] = o.r;
        reinterpret_cast<Real *>(a.tr)[g] = o.tr;
        write_obs<Real, K>(o, i, xi, yi, c, sm.pos + le * n, sm.act + le * n, sm.radius, P,
                           reinterpret_cast<Real *>(a.z), a.Ni, g);
        if (o.ncoll) atomicAdd(&sm.cnt[le], o.ncoll);
        if (!o.at_goal) sm.notgoal[le] = 1;
    }
    __syncthreads();
    if (active && i == 0) {
        a.ncoll[e] = sm.cnt[le];                                           // :284
        if (a.do_integrate) {
            const int tt = a.t[e];
            a.fin[e] = (sm.notgoal[le] == 0 || tt >= a.max_steps - 1) ? 1 : 0;   // :251
            a.t[e] = tt + 1;                                               // :256
        }
    }
}

// ---------------------------------------------------------------- closed-loop rollout kernel
// T closed-loop steps in ONE launch: every step's action comes from one of the reference's
// baseline controllers (control_action) evaluated on the CURRENT state, so the steps of an
// environment are sequential and the time-parallel chunks of rollout_kernel do not apply; a CTA
// owns G whole environments, keeps their state on chip for the whole call and loops over time
// (five barriers per step, no launch or host round trip in between).  Recording, finished codes,
// done flags and episode sums follow ds_rollout.
template <typename Real, int K, int NT>
__global__ void __launch_bounds__(NT)
rollout_control_kernel(const RolloutArgs ra)
{
    using V2 = typename vec2_of<Real>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepArgs &a = ra.s;
    const int n = a.n, G = a.G, T = ra.T;
    CtaSmem<Real> sm(smem_raw, n, G, 1);
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        sm.delta[idx] = ((const Real *)a.c.delta)[idx];
        sm.radius[idx] = ((const Real *)a.c.radius)[idx];
    }
    const ParamsR<Real> P(a);
    const int kk = (K >= 0) ? K : a.k;
    const int cols = a.simplify ? 2 : 5;
    const int lt = threadIdx.x;
    const int le = lt / n, i = lt - le * n;
    const long long e = (long long)blockIdx.x * G + le;
    const bool active = (lt < G * n) && (e < a.E);
    const size_t g = active ? (size_t)e * n + i : 0;
    const size_t EN = (size_t)a.E * n;
    AgentConst<Real> c{};
    V2 p{}, v{};
    if (active) {
        c = load_agent_const<Real>(a.c, i);
        p = reinterpret_cast<const V2 *>(a.pos)[g];
        v = reinterpret_cast<const V2 *>(a.vel)[g];
        if (i == 0) { sm.alive[le] = (ra.done[e] == 0) ? 1 : 0; sm.tenv[le] = a.t[e]; sm.nexec[le] = 0; }
    }
    Real sum_r = 0, sum_tr = 0;
    int sum_c = 0;
    for (int step = 0; step < T; ++step) {
        if (active) sm.pos[lt] = p;
        __syncthreads();                                   // current positions, alive / t of this step
        const bool alive = active && sm.alive[le] != 0;
        V2 u{};
        if (alive) {
            if (a.ctrl)
                control_action<Real>(a.ctrl, n, i, p.x, p.y, c.xF, c.yF, c.ds, c.radius, sm.pos + le * n, sm.radius, 1,
                                     (Real)a.u_max, u.x, u.y);
            else                                           // actions given (ds_rollout_policy: the actors' draws)
                u = reinterpret_cast<const V2 *>(ra.actions)[(size_t)step * EN + g];
        }
        __syncthreads();                                   // everybody has read the old positions
        if (alive) {
            p.x = add_rn(p.x, mul_rn(P.dt, u.x));          // A = I, B = dt I (:78-79,235)
            p.y = add_rn(p.y, mul_rn(P.dt, u.y));
            v = u;                                         // :238
            sm.pos[lt] = p;
            sm.act[lt] = v;
            if (i == 0) { sm.cnt[le] = 0; sm.notgoal[le] = 0; }
        }
        __syncthreads();
        RowResult<Real, K> o;
        const size_t at = (size_t)step * EN + g;
        if (alive) {
            eval_row<Real, K>(o, n, i, p.x, p.y, c, sm.pos + le * n, sm.delta, sm.radius, P, a.c.logtab);
            if (o.ncoll) atomicAdd(&sm.cnt[le], o.ncoll);
            if (!o.at_goal) sm.notgoal[le] = 1;
            if (ra.pos_tr) reinterpret_cast<V2 *>(ra.pos_tr)[at] = p;
            if (ra.vel_tr) reinterpret_cast<V2 *>(ra.vel_tr)[at] = v;
            if (ra.r_tr) reinterpret_cast<Real *>(ra.r_tr)[at] = o.r;
            if (ra.tr_tr) reinterpret_cast<Real *>(ra.tr_tr)[at] = o.tr;
            if (ra.z_tr)
                write_obs<Real, K>(o, i, p.x, p.y, c, sm.pos + le * n, sm.act + le * n, sm.radius, P,
                                   reinterpret_cast<Real *>(ra.z_tr), ra.Ni_tr, at);
            sum_r = add_rn(sum_r, o.r); sum_tr = add_rn(sum_tr, o.tr);
        }
        __syncthreads();                                   // collision count / not-at-goal complete
        const size_t fe = (size_t)step * a.E + (size_t)e;
        if (alive) {
            const int nc = sm.cnt[le];
            const int tt = sm.tenv[le];
            const bool fin = (sm.notgoal[le] == 0) || (tt >= a.max_steps - 1);        // :248-254
            if (fin || step == T - 1) {
                // last executed step of the call: the step()-style outputs in the live buffers
                reinterpret_cast<Real *>(a.r)[g] = o.r;
                reinterpret_cast<Real *>(a.tr)[g] = o.tr;
                write_obs<Real, K>(o, i, p.x, p.y, c, sm.pos + le * n, sm.act + le * n, sm.radius, P,
                                   reinterpret_cast<Real *>(a.z), a.Ni, g);
            }
            if (i == 0) {
                sum_c += nc;
                if (ra.ncoll_tr) ra.ncoll_tr[fe] = nc;
                if (ra.fin_tr) ra.fin_tr[fe] = fin ? 1 : 0;
                if (fin || step == T - 1) { a.ncoll[e] = nc; a.fin[e] = fin ? 1 : 0; }
            }
        } else if (active && i == 0 && ra.fin_tr) {
            ra.fin_tr[fe] = 2;
        }
        __syncthreads();                                   // every row has read alive / t / flags of this step
        if (alive && i == 0) {
            const bool fin = (sm.notgoal[le] == 0) || (sm.tenv[le] >= a.max_steps - 1);
            sm.tenv[le] += 1;                              // :256
            sm.nexec[le] += 1;
            if (fin) sm.alive[le] = 0;
        }
    }
    // final state, episode sums (reduced per environment in a fixed order), done flags
    __syncthreads();
    if (active) {
        reinterpret_cast<V2 *>(a.pos)[g] = p;
        reinterpret_cast<V2 *>(a.vel)[g] = v;
        sm.r[lt] = sum_r; sm.tr[lt] = sum_tr;
    }
    __syncthreads();
    if (active && i == 0) {
        a.t[e] = sm.tenv[le];
        if (sm.nexec[le] > 0) {
            if (sm.alive[le] == 0) ra.done[e] = 1;
            double sr = 0, st = 0;
            for (int j = 0; j < n; ++j) { sr += (double)sm.r[le * n + j]; st += (double)sm.tr[le * n + j]; }
            double *ag4 = ra.agg + (size_t)e * 4;
            ag4[0] += sr / n; ag4[1] += st / n; ag4[2] += (double)sum_c; ag4[3] += (double)sm.nexec[le];
        }
    }
}

// ---------------------------------------------------------------- rollout kernel
// Shared memory of a rollout CTA.  Per CTA: constants [n], log table, work list [L].  Per row
// (I = TC*G*n), double buffered over chunks: action, position.  Per agent of a slice
// (A = G*n): final position / velocity.  Per frame (F = TC*G): collision count.  Per
// environment (G): bit s of ngbits = "slice s has an agent that is not at its goal"; alive and t
// per chunk parity.
template <typename Real> struct RoSmem {
    using V2 = typename vec2_of<Real>::type;
    V2 *cA;                  // [n] (d_safety, log d_safety)
    V2 *cB;                  // [n] (radius, delta)
    V2 *cF;                  // [n] end point
    float *cTf;              // [n] pass-1 threshold of the packed-f32 filter (squared, with margin)
    int *cC;                 // [n] Delta-disk count of the clipped pairs
    Real *cT;                // [n] exact pass-1 threshold (rows that evaluate themselves)
    float4 *pf_;             // two buffers of F * hp: (x_2q, x_2q+1, y_2q, y_2q+1) of each frame, f32
    int hp_, F_;
    LogTabEntry *logtab;
    V2 *act_, *pos_, *p0, *vfin, *res;   // act_: three, pos_: two buffers of I rows (chunk c mod 3 / mod 2)
    V2 *accv;                // [I] running (sum r, sum true_r) of the row over the call
    int *accc;               // [I] running collision count (rows with i == 0)
    unsigned *ent;
    int I_, G_;
    int *cnt, *alive_, *tenv_, *steps, *lcount;  // alive_ / tenv_: per chunk parity
    unsigned *ngbits;
    __device__ V2 *act(int b) const { return act_ + b * I_; }
    __device__ V2 *pos(int b) const { return pos_ + b * I_; }
    __device__ float4 *pf(int b) const { return pf_ + b * F_ * hp_; }
    __device__ int *alive(int b) const { return alive_ + b * G_; }
    __device__ int *tenv(int b) const { return tenv_ + b * G_; }
    __host__ __device__ static size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }
    __host__ __device__ static size_t bytes(int n, int G, int TC, int L)
    {
        const size_t A = (size_t)G * n, I = A * TC, F = (size_t)G * TC;
        return align16((3 * (size_t)n + 6 * I + 2 * A + (size_t)L) * sizeof(V2)) + kLogTabSize * sizeof(LogTabEntry) +
               align16(n * sizeof(Real)) + align16(n * sizeof(int)) + align16(n * sizeof(float)) +
               2 * F * ((n + 1) / 2) * sizeof(float4) + align16(I * sizeof(int)) +
               align16((size_t)L * sizeof(unsigned)) + align16(F * sizeof(int)) +
               align16((6 * (size_t)G + 1) * sizeof(int));
    }
    __device__ RoSmem(unsigned char *base, int n, int G, int TC, int L)
    {
        const size_t A = (size_t)G * n, I = A * TC, F = (size_t)G * TC;
        unsigned char *p = base;
        I_ = (int)I; G_ = G;
        cA = reinterpret_cast<V2 *>(p); cB = cA + n; cF = cB + n;
        act_ = cF + n; pos_ = act_ + 3 * I;
        p0 = pos_ + 2 * I; vfin = p0 + A; accv = vfin + A; res = accv + I;
        p += align16((3 * (size_t)n + 6 * I + 2 * A + (size_t)L) * sizeof(V2));   // (float32: an odd count of float2 would
        logtab = reinterpret_cast<LogTabEntry *>(p); p += kLogTabSize * sizeof(LogTabEntry);   // leave the float4 frames misaligned)
        hp_ = (n + 1) / 2; F_ = (int)F;
        pf_ = reinterpret_cast<float4 *>(p); p += 2 * F * hp_ * sizeof(float4);
        cT = reinterpret_cast<Real *>(p); p += align16(n * sizeof(Real));
        cC = reinterpret_cast<int *>(p); p += align16(n * sizeof(int));
        cTf = reinterpret_cast<float *>(p); p += align16(n * sizeof(float));
        accc = reinterpret_cast<int *>(p); p += align16(I * sizeof(int));
        ent = reinterpret_cast<unsigned *>(p); p += align16((size_t)L * sizeof(unsigned));
        cnt = reinterpret_cast<int *>(p); p += align16(F * sizeof(int));
        alive_ = reinterpret_cast<int *>(p); tenv_ = alive_ + 2 * G;
        ngbits = reinterpret_cast<unsigned *>(tenv_ + 2 * G); steps = tenv_ + 3 * G; lcount = tenv_ + 4 * G;
    }
};

// Ampere-style asynchronous global -> shared copy of one action (LDGSTS): the action stream of
// chunk c + 2 is in flight while chunk c is evaluated, without holding it in registers.
__device__ __forceinline__ void cp_async_action(double2 *dst, const double2 *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_action(float2 *dst, const float2 *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

// Pass 1 over one block of <= 32 agents of the row's frame: bit jj of the result = pair
// (i, j0 + jj) is NOT provably clipped.  near <=> !(d2 >= thr2) <=> the sign bit of thr2 - d2 is
// clear (+0 and NaN count as near: the exact path is always right), shifted into the mask with
// one funnel shift per pair.
__device__ __forceinline__ int sign_word(double t) { return __double2hiint(t); }
__device__ __forceinline__ int sign_word(float t) { return __float_as_int(t); }
template <typename Real>
__device__ __forceinline__ unsigned pass1_block(const typename vec2_of<Real>::type *__restrict__ fp, int jn,
                                                Real px, Real py, Real thr2)
{
    using V2 = typename vec2_of<Real>::type;
    unsigned m = 0;
#pragma unroll 4
    for (int jj = 0; jj < jn; ++jj) {
        const V2 pj = fp[jj];
        const Real dx = sub_rn(px, pj.x), dy = sub_rn(py, pj.y);
        const Real d2 = fma_rn(dy, dy, mul_rn(dx, dx));
        m = __funnelshift_l((unsigned)sign_word(sub_rn(thr2, d2)), m, 1);   // m = m << 1 | sign
    }
    // m: NOT-near bits, agent jj at bit jn - 1 - jj
    return __brev(~m) >> (32 - jn);
}

// Pass 1 of the rollout kernel on Blackwell's packed-f32 pipe (FADD2 / FMUL2 / FFMA2: two pairs per
// instruction, 4 instructions per pair instead of 7 and nothing on the FP64 pipe).  The near mask
// only has to be a SUPERSET of the pairs that are not clipped -- the exact fp64 path handles a
// clipped pair correctly -- so an f32 distance against a threshold with a margin far above the
// f32 rounding of coordinates below 1024 is enough; agents with non-finite or larger coordinates
// are stored as NaN, which makes every pair they are in "near" (sign bit of thr - d2 clear).
// fp: float4 per two agents (x0, x1, y0, y1); nx, ny: MINUS the row's own f32 position.
__device__ __forceinline__ unsigned pass1_block_f32x2(const float4 *__restrict__ fp, int jn, float nx, float ny,
                                                      float thr2f)
{
    const float2 NX = make_float2(nx, nx), NY = make_float2(ny, ny);
    const float2 TH = make_float2(thr2f, thr2f), M1 = make_float2(-1.0f, -1.0f);
    const int np = (jn + 1) >> 1;
    unsigned m = 0;
#pragma unroll 4
    for (int q = 0; q < np; ++q) {
        const float4 pq = fp[q];
        const float2 a = __fadd2_rn(make_float2(pq.x, pq.y), NX);
        const float2 b = __fadd2_rn(make_float2(pq.z, pq.w), NY);
        const float2 t = __ffma2_rn(__ffma2_rn(b, b, __fmul2_rn(a, a)), M1, TH);   // thr2 - d2
        m = __funnelshift_l(__float_as_uint(t.x), m, 1);
        m = __funnelshift_l(__float_as_uint(t.y), m, 1);
    }
    // m: NOT-near bits of 2 np agents, agent jj at bit 2 np - 1 - jj (a pad agent sits at 1e30: not near)
    m = __brev(~m) >> (32 - 2 * np);
    return (jn < 32) ? (m & ((1u << jn) - 1u)) : m;
}

// Which slices of a chunk execute (drone_env.py:248-256): the episode ends at the first slice
// whose agents are all at their goals (bit clear in ngbits) or whose t reaches max_steps - 1.
// Returns the number of executed slices; *env_fin = the last of them finished the episode.
DS_HD int executed_slices(unsigned ngbits, int tt, int nsl, int max_steps, bool *env_fin)
{
    const unsigned full = (nsl >= 32) ? 0xffffffffu : ((1u << nsl) - 1u);
    const unsigned atgoal = ~ngbits & full;
    const int fg = atgoal ? lowest_bit(atgoal) : nsl;        // first slice with everybody at goal
    int ft = max_steps - 1 - tt;                             // first slice at the time limit
    ft = ft < 0 ? 0 : ft;
    const int first = fg < ft ? fg : ft;
    *env_fin = first < nsl;
    return *env_fin ? first + 1 : nsl;
}

// T fused steps, TC <= 32 time slices per chunk evaluated concurrently (see the header comment).
// NB = number of 32-agent blocks whose near masks a row keeps in registers (n <= 32 NB);
// NB == 0: any n, masks in local memory.  Element indices are 32 bit: the host splits a call
// whose T * E * n would not fit.
//
// Positions of chunk c + 1 are produced during chunk c by one thread per agent: sequential,
// bit-exact single-integrator steps (A = I, B = dt I, drone_env.py:78-79,235) continuing from the
// LAST slice of chunk c.  That is speculative only in appearance: if the episode ends inside
// chunk c the environment stops stepping and the positions are never looked at.
// One chunk (default: warp-local work lists, TWO barriers):
//   (c) every row: packed-f32 pass 1 over its frame -> near mask; a warp scan gives the row a
//       contiguous segment of its WARP's slice of the work list; entries written; the action
//       stream of chunk c + 2 starts its cp.async                                      | __syncwarp
//   (d) the warp's lanes: one near pair per lane per round (eval_pair)                 | __syncwarp
//   (e) every row folds its segment and finishes the row; collision count and not-at-goal bit
//       posted to its frame / environment                                              | barrier
//   (g) every row of an executed slice stores its outputs and adds to its running episode sums;
//       one thread per environment advances t / alive; one thread per agent integrates chunk
//       c + 1                                                                          | barrier
// ra.inline_rows selects the alternatives kept for comparison: 0 = one CTA-wide list (one shared
// atomic per warp, barriers around (d), the next chunk integrated during (d) by a warp that takes
// no pairs; -1...-2 %), 1 = rows evaluate their own near pairs (n <= 32; same instruction count at
// config-3 density, worse when the rows of a warp differ more).
// Register cap of the 256-thread instantiations.  The kernel is latency bound, so resident warps
// matter, but a cap that makes ptxas spill costs more than it gains: measured on B200 (config 3,
// 160-thread CTAs) 96 registers / 4 CTAs per SM = 1.30e10 agent-steps/s, 104-112 / 3 CTAs = 1.16e10,
// 80 with 50 B of spills / 5 CTAs = 1.09e10, 64-72 with 200-370 B of spills = 0.97-1.01e10.
#ifndef DS_RO_MAXNREG
#define DS_RO_MAXNREG 96
#endif
#define DS_RO_BOUNDS(NT) __launch_bounds__(NT) __maxnreg__((NT <= 256) ? DS_RO_MAXNREG : 64)
template <typename Real, int K, int NT, int NB>
__global__ void DS_RO_BOUNDS(NT)
rollout_kernel(const RolloutArgs ra)
{
    using V2 = typename vec2_of<Real>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepArgs &a = ra.s;
    const int n = a.n, G = a.G, TC = ra.TC, E = a.E, T = ra.T, L = ra.L;
    RoSmem<Real> sm(smem_raw, n, G, TC, L);
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        V2 v;
        v.x = ((const Real *)a.c.d_safety)[idx]; v.y = ((const Real *)a.c.log_ds)[idx]; sm.cA[idx] = v;
        v.x = ((const Real *)a.c.radius)[idx]; v.y = ((const Real *)a.c.delta)[idx]; sm.cB[idx] = v;
        sm.cF[idx] = reinterpret_cast<const V2 *>(a.c.xF)[idx];
        sm.cT[idx] = ((const Real *)a.c.thr2)[idx];
        sm.cC[idx] = a.c.clipcnt[idx];
        {   // (sqrt(thr2) + 2e-3)^2 rounded up: margin >> f32 rounding of |coordinates| < 1024
            const double th = sqrt((double)sm.cT[idx]) + 2e-3;
            sm.cTf[idx] = __double2float_ru(th * th);
        }
    }
    if (sizeof(Real) == 8)
        for (int idx = threadIdx.x; idx < kLogTabSize; idx += blockDim.x) sm.logtab[idx] = a.c.logtab[idx];
    const ParamsR<Real> P(a);
    const int A = G * n;                              // agents per slice in this CTA
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    // (exact multiply-high division: ptxas re-derives these per chunk rather than keep them live)
    const int s = ra.mulA ? (int)__umulhi((unsigned)tid, ra.mulA) : tid, ag = tid - s * A;   // slice in the chunk, agent slot
    const int le = ra.mulN ? (int)__umulhi((unsigned)ag, ra.mulN) : ag, i = ag - le * n;
    const int e = blockIdx.x * G + le;
    const bool active = (s < TC) && (e < E);
    const unsigned g = active ? (unsigned)e * n + i : 0u;
    const unsigned EN = (unsigned)E * n;
    const int fr = s * G + le;                        // frame slot of this thread
    const V2 *atab = reinterpret_cast<const V2 *>(ra.atable);
    const bool agent_thread = tid < A && e < E;       // s == 0: owns agent ag across the call
    const bool env_thread = agent_thread && i == 0;   // owns environment le across the call
    // lanes of this warp that hold rows of the same frame; the first of them posts for the frame
    unsigned fmask;
    {
        const int lo = (lane - i > 0) ? lane - i : 0, hi = (lane + (n - 1 - i) < 31) ? lane + (n - 1 - i) : 31;
        fmask = ((hi - lo == 31) ? 0xffffffffu : ((1u << (hi - lo + 1)) - 1u)) << lo;
    }
    const bool frame_poster = (i == 0 || lane == 0);

    if (agent_thread) {
        sm.p0[tid] = reinterpret_cast<const V2 *>(a.pos)[g];
        sm.vfin[tid] = reinterpret_cast<const V2 *>(a.vel)[g];
        if (i == 0) { sm.alive(0)[le] = (ra.done[e] == 0) ? 1 : 0; sm.tenv(0)[le] = a.t[e]; }
    }
    // running episode sums live in shared memory (registers are what limits residency)
    if (active) { V2 z2; z2.x = 0; z2.y = 0; sm.accv[tid] = z2; sm.accc[tid] = 0; }
    if (env_thread) sm.steps[le] = 0;

    // action of this row at element index `at` -> act(b)[tid].  Explicit actions go global -> shared
    // asynchronously (no registers); index mode keeps the prefetched u8 index in one register.
    const bool direct = ra.actions != nullptr;
    // sequential integration of agent `tid` through nsl slices: pos(buf)[q] = start + dt u_0 .. u_q
    auto integrate = [&](V2 p, int abuf, int pbuf, int nsl) {
        const V2 *ua = sm.act(abuf) + tid;
        V2 *pa = sm.pos(pbuf) + tid;
        // f32 copy for pass 1: component (i & 1) of x / y in float4 (i >> 1) of frame q * G + le
        float *pfa = reinterpret_cast<float *>(sm.pf(pbuf) + le * sm.hp_ + (i >> 1)) + (i & 1);
        const int fstride = G * sm.hp_ * 4;                  // floats per slice
        const bool pad = (i == n - 1) && (n & 1);            // odd n: the missing partner never is near
        for (int q = 0; q < nsl; ++q) {
            const V2 uq = ua[q * A];
            p.x = add_rn(p.x, mul_rn(P.dt, uq.x));
            p.y = add_rn(p.y, mul_rn(P.dt, uq.y));
            pa[q * A] = p;
            const bool ok = fabs(p.x) < (Real)1024 && fabs(p.y) < (Real)1024;   // false for NaN / inf too
            pfa[q * fstride] = ok ? (float)p.x : __int_as_float(0x7fc00000);
            pfa[q * fstride + 2] = ok ? (float)p.y : __int_as_float(0x7fc00000);
            if (pad) { pfa[q * fstride + 1] = 1e30f; pfa[q * fstride + 3] = 1e30f; }
        }
    };
    unsigned at = (unsigned)s * EN + g;               // element index of this row at slice t0 + s
    unsigned fe = (unsigned)s * E + (unsigned)e;
    // stage chunk 0 (synchronously) and chunk 1 (asynchronously / index prefetched)
    unsigned ai_next = 0;
    if (active && s < T)
        sm.act(0)[tid] = direct ? reinterpret_cast<const V2 *>(ra.actions)[at] : atab[ra.aidx[at]];
    if (active && TC + s < T) {
        if (direct) cp_async_action(&sm.act(1)[tid], reinterpret_cast<const V2 *>(ra.actions) + at + (unsigned)TC * EN);
        else sm.act(1)[tid] = atab[ra.aidx[at + (unsigned)TC * EN]];
    }
    cp_async_commit();
    if (active && !direct && 2 * TC + s < T) ai_next = ra.aidx[at + 2u * (unsigned)TC * EN];
    if (tid == 0) *sm.lcount = 0;
    __syncthreads();
    if (agent_thread && sm.alive(0)[le] != 0) integrate(sm.p0[tid], 0, 0, (T < TC) ? T : TC);
    __syncthreads();

    // dense small-n frames: every row evaluates its own near pairs, two barriers per chunk
    const bool inl = (NB == 1) && ra.inline_rows == 1;
    // warp-local work lists: list build, pair evaluation and fold of a warp's rows need no CTA barrier
    const bool wl = ra.inline_rows == 2;
    const int LW = L / (int)(blockDim.x >> 5);
    int ab = 0, ab1 = 1, ab2 = 2;                           // action buffers of chunk c, c + 1, c + 2
    for (int t0 = 0, buf = 0; t0 < T; t0 += TC, buf ^= 1) {
        const int nsl = (T - t0 < TC) ? (T - t0) : TC;      // slices in this chunk
        const bool in_chunk = active && s < nsl;
        const bool env_alive = active && sm.alive(buf)[le] != 0;
        const bool valid = in_chunk && env_alive;
        const bool more = t0 + TC < T;                       // there is a chunk after this one
        // (c) pass 1
        V2 p{};
        constexpr int NBR = (NB > 0) ? NB : 32;
        unsigned near[NBR];
        int ncnt = 0;
        const V2 *fpos = sm.pos(buf) + (tid - i);           // positions of this row's frame
        if (env_thread) sm.ngbits[le] = 0;
        if (valid) {
            p = fpos[i];
            if (i == 0) sm.cnt[fr] = 0;
            const float4 *ffr = sm.pf(buf) + fr * sm.hp_;    // packed f32 positions of this row's frame
            const float thr2f = sm.cTf[i];
            const float nx = -reinterpret_cast<const float *>(ffr + (i >> 1))[i & 1];
            const float ny = -reinterpret_cast<const float *>(ffr + (i >> 1))[2 + (i & 1)];
            if (NB > 0) {
#pragma unroll
                for (int bk = 0; bk < NBR; ++bk) {
                    const int j0 = bk * 32;
                    unsigned m = 0;
                    if (j0 < n) {
                        m = pass1_block_f32x2(ffr + j0 / 2, (n - j0 < 32) ? (n - j0) : 32, nx, ny, thr2f);
                        if ((unsigned)(i - j0) < 32u) m &= ~(1u << (i - j0));
                    }
                    near[bk] = m;
                    ncnt += __popc(m);
                }
            } else {
                for (int j0 = 0, bk = 0; j0 < n; j0 += 32, ++bk) {
                    unsigned m = pass1_block_f32x2(ffr + j0 / 2, (n - j0 < 32) ? (n - j0) : 32, nx, ny, thr2f);
                    if ((unsigned)(i - j0) < 32u) m &= ~(1u << (i - j0));
                    near[bk] = m;
                    ncnt += __popc(m);
                }
            }
        }
        // segment of the work list: exclusive scan over the warp
        int seg = 0, wbase = 0, wtotal = 0;
        bool listed = true;
        if (!inl) {
            int incl = ncnt;
#pragma unroll
            for (int w = 1; w < 32; w <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, w);
                if (lane >= w) incl += v;
            }
            int lend;                                        // end of the list region this row may use
            if (wl) {                                        // warp-local list: a fixed slice per warp
                wbase = (tid >> 5) * LW;
                lend = wbase + LW;
                wtotal = __shfl_sync(0xffffffffu, incl, 31);
            } else {                                         // CTA-wide list: one atomic per warp
                if (lane == 31 && incl > 0) wbase = atomicAdd(sm.lcount, incl);
                wbase = __shfl_sync(0xffffffffu, wbase, 31);
                lend = L;
            }
            seg = wbase + incl - ncnt;
            listed = ncnt == 0 || seg + ncnt <= lend;        // otherwise the row evaluates itself in (e)
            if (valid && ncnt > 0) {
                unsigned *ep = sm.ent + seg;
                if (listed) {
                    const unsigned base = (unsigned)tid | ((unsigned)i << 20);
                    const int nb = (NB > 0) ? NBR : (n + 31) / 32;
#pragma unroll
                    for (int bk = 0; bk < nb; ++bk) {
                        unsigned m = near[bk];
                        while (m) {
                            const int jj = lowest_bit(m);
                            m &= m - 1;
                            *ep++ = base | ((unsigned)(bk * 32 + jj) << 10);
                        }
                    }
                } else {
                    for (int q = seg; q < lend && q < seg + ncnt; ++q) sm.ent[q] = kEntSkip;
                }
            }
        }
        // actions of chunk c + 2 on their way to act(ab2); chunk c + 1's (issued one chunk ago) are
        // complete after the wait and visible to the integration after the next barrier
        if (active && t0 + 2 * TC + s < T) {
            if (direct) {
                cp_async_action(&sm.act(ab2)[tid], reinterpret_cast<const V2 *>(ra.actions) + at + 2u * (unsigned)TC * EN);
            } else {
                sm.act(ab2)[tid] = atab[ai_next];
                if (t0 + 3 * TC + s < T) ai_next = ra.aidx[at + 3u * (unsigned)TC * EN];
            }
        }
        cp_async_commit();
        cp_async_wait_but_one();
        if (wl) __syncwarp(); else if (!inl) __syncthreads();
        // (d) pass 2: one near pair per thread per round
        if (!inl) {
            // CTA-wide list: the warp(s) of the agent threads integrate the next chunk in this phase,
            // so when the CTA has other warps the pairs go to those.  Warp-local list: every warp
            // works through its own slice.
            const int w0 = (!wl && more && A <= 32 && blockDim.x > 64) ? 32 : 0;
            const int M = wl ? ((wtotal < LW) ? wtotal : LW) : ((*sm.lcount < L) ? *sm.lcount : L);
            const int q0 = wl ? wbase + lane : tid - w0, qe = wl ? wbase + M : M, qs = wl ? 32 : (int)blockDim.x - w0;
            for (int q = q0; q < qe && tid >= w0; q += qs) {
                const unsigned w = sm.ent[q];
                if (w == kEntSkip) continue;
                const int row = (int)(w & 1023u), j = (int)((w >> 10) & 1023u), ri = (int)(w >> 20);
                const V2 pi = sm.pos(buf)[row], pj = sm.pos(buf)[row - ri + j];
                const V2 ca = sm.cA[ri], cbi = sm.cB[ri], cbj = sm.cB[j];
                PairOut<Real> po;
                eval_pair<Real>(po, pi.x, pi.y, pj.x, pj.y, ca.x, cbi.x, cbj.x, cbj.y, ca.y, P, sm.logtab);
                V2 dv; dv.x = po.d; dv.y = po.logd;
                sm.res[q] = dv;
                sm.ent[q] = pack_result(j, po.in_disk, po.coll, (po.in_disk ? 1 : 0) - ((ca.x <= cbj.y) ? 1 : 0) + 1);
            }
        }
        if (wl) {
            __syncwarp();
        } else if (!inl) {
            if (agent_thread && more && env_alive)
                integrate(sm.pos(buf)[(TC - 1) * A + tid], ab1, buf ^ 1, (T - t0 - TC < TC) ? (T - t0 - TC) : TC);
            __syncthreads();
        }
        // (e) rows
        RowResult<Real, K> o;
        bool notgoal = false;
        if (valid) {
            AgentConst<Real> c;
            const V2 ca = sm.cA[i], cb = sm.cB[i], cf = sm.cF[i];
            c.xF = cf.x; c.yF = cf.y; c.ds = ca.x; c.log_ds = ca.y; c.radius = cb.x; c.delta = cb.y;
            c.thr2 = sm.cT[i]; c.clipcnt = sm.cC[i];
            p = fpos[i];
            if (inl)
                eval_row_near32<Real, K>(o, n, i, p.x, p.y, c, near[0], fpos, sm.cB, P, sm.logtab);
            else if (listed)
                eval_row_from_list<Real, K>(o, n, i, p.x, p.y, c, sm.ent + seg, sm.res + seg, ncnt, P);
            else
                eval_row<Real, K>(o, n, i, p.x, p.y, c, fpos, &sm.cB[0].y, &sm.cB[0].x, P, sm.logtab, 2);
            if (o.ncoll) atomicAdd(&sm.cnt[fr], o.ncoll);
            notgoal = !o.at_goal;
        }
        {
            // one atomic per (frame, warp): any row of the frame in this warp not at its goal
            const unsigned bal = __ballot_sync(0xffffffffu, notgoal);
            if (frame_poster && active && (bal & fmask)) atomicOr(&sm.ngbits[le], 1u << s);
        }
        if (tid == 0) *sm.lcount = 0;                        // every thread has read it in (d)
        __syncthreads();
        // two-barrier modes: the next chunk's actions are staged by now; integrate here
        if ((inl || wl) && agent_thread && more && env_alive)
            integrate(sm.pos(buf)[(TC - 1) * A + tid], ab1, buf ^ 1, (T - t0 - TC < TC) ? (T - t0 - TC) : TC);
        // (g) stores
        int ne = 0;
        bool env_fin = false;
        if (env_alive) ne = executed_slices(sm.ngbits[le], sm.tenv(buf)[le], nsl, a.max_steps, &env_fin);
        if (valid) {
            if (s < ne) {
                const int nc = sm.cnt[fr];
                const bool fin = env_fin && s == ne - 1;
                const V2 *fvel = sm.act(ab) + (tid - i);
                AgentConst<Real> c{};
                { const V2 cb = sm.cB[i]; c.radius = cb.x; c.delta = cb.y; }
                if (ra.pos_tr) reinterpret_cast<V2 *>(ra.pos_tr)[at] = p;
                if (ra.vel_tr) reinterpret_cast<V2 *>(ra.vel_tr)[at] = fvel[i];        // :238
                if (ra.r_tr) reinterpret_cast<Real *>(ra.r_tr)[at] = o.r;
                if (ra.tr_tr) reinterpret_cast<Real *>(ra.tr_tr)[at] = o.tr;
                if (ra.z_tr)
                    write_obs<Real, K>(o, i, p.x, p.y, c, fpos, fvel, &sm.cB[0].x, P,
                                       reinterpret_cast<Real *>(ra.z_tr), ra.Ni_tr, at, 2);
                if (fin || t0 + s == T - 1) {
                    // last executed step of the call: leave the step()-style outputs in the live buffers
                    reinterpret_cast<Real *>(a.r)[g] = o.r;
                    reinterpret_cast<Real *>(a.tr)[g] = o.tr;
                    write_obs<Real, K>(o, i, p.x, p.y, c, fpos, fvel, &sm.cB[0].x, P,
                                       reinterpret_cast<Real *>(a.z), a.Ni, g, 2);
                    if (i == 0) { a.ncoll[e] = nc; a.fin[e] = fin ? 1 : 0; }
                }
                {
                    V2 acc = sm.accv[tid];
                    acc.x = add_rn(acc.x, o.r); acc.y = add_rn(acc.y, o.tr);
                    sm.accv[tid] = acc;
                }
                if (i == 0) {
                    sm.accc[tid] += nc;
                    if (ra.ncoll_tr) ra.ncoll_tr[fe] = nc;
                    if (ra.fin_tr) ra.fin_tr[fe] = fin ? 1 : 0;
                }
            } else if (i == 0 && ra.fin_tr) {
                ra.fin_tr[fe] = 2;
            }
        } else if (in_chunk && i == 0 && ra.fin_tr) {
            ra.fin_tr[fe] = 2;
        }
        // one thread per agent: state after the last executed slice, when the call ends here
        if (agent_thread && env_alive && (env_fin || !more)) {
            sm.p0[tid] = sm.pos(buf)[(ne - 1) * A + tid];
            sm.vfin[tid] = sm.act(ab)[(ne - 1) * A + tid];
        }
        // one thread per environment: t / alive of the next chunk (other parity)
        if (env_thread) {
            sm.steps[le] += ne;
            sm.tenv(buf ^ 1)[le] = sm.tenv(buf)[le] + ne;
            sm.alive(buf ^ 1)[le] = (env_alive && !env_fin) ? 1 : 0;
        }
        at += (unsigned)TC * EN; fe += (unsigned)TC * E;
        { const int t3 = ab; ab = ab1; ab1 = ab2; ab2 = t3; }
        __syncthreads();
    }
    const int fbuf = ((T + TC - 1) / TC) & 1;                // parity the last chunk wrote
    // episode sums (train_problem.py:98-100): sum over the call of mean_i r, mean_i true_r, the
    // collision counts and the steps, reduced per environment in a fixed order
    if (agent_thread) {
        reinterpret_cast<V2 *>(a.pos)[g] = sm.p0[tid];
        reinterpret_cast<V2 *>(a.vel)[g] = sm.vfin[tid];
    }
    if (env_thread) {
        a.t[e] = sm.tenv(fbuf)[le];
        const int steps = sm.steps[le];
        if (steps > 0) {
            if (sm.alive(fbuf)[le] == 0) ra.done[e] = 1;
            double sr = 0, st = 0, sc = 0;
            for (int q = 0; q < TC; ++q) {
                const V2 *rv = sm.accv + q * A + le * n;
                for (int j = 0; j < n; ++j) { sr += (double)rv[j].x; st += (double)rv[j].y; }
                sc += (double)sm.accc[q * A + le * n];
            }
            double *ag4 = ra.agg + (size_t)e * 4;
            ag4[0] += sr / n; ag4[1] += st / n; ag4[2] += sc; ag4[3] += (double)steps;
        }
    }
}

// ---------------------------------------------------------------- returns / advantages
// SURVEY.md section 8f row 2 -- what the reference's learners do with the reward and Ni
// trajectories of an episode (replaces the host-side ExperienceBuffers walk):
//   returns     G_i(t) = G_i(t+1) * discount + r_i(t), G_i(last) = r_i(last)   (SAC_agents.py:304-310,
//               108-113), multiply then add, unfused, backwards from the last executed step;
//   advantages  A_i(t) = sum over j in N_i(t), in list order, of (G_j(t) - V_i(t))   (:333-345),
//               V = the critic's baseline (0 when none is given).
// One thread per agent walks its environment's episode backwards; the G of a step are exchanged
// through shared memory (double buffered: one barrier per step).  Inputs are loaded U = 4 steps
// at a time so that the memory latency is paid once per block of steps, not per step.  A step is
// executed iff its finished code (ds_rollout: 0 running, 1 finished here, 2 not executed) is not
// 2; not-executed steps get zeros.  HBM bound: 37 B per agent-step (f64, no baseline).
#ifndef DS_RET_U
#define DS_RET_U 4        // steps whose inputs are in flight together (U = 8: +10% at config 3, -18% at the HBM point)
#endif
struct ReturnsArgs {
    int E, n, k, T, G;
    double discount;
    const void *r_tr, *base;
    const int *Ni_tr;
    const uint8_t *fin_tr;
    void *ret, *adv;
    uint8_t *cnt;
};

template <typename Real, int KP1>      // KP1 = k + 1 neighbour slots held in registers (0: any k)
__global__ void __launch_bounds__(256) returns_kernel(const ReturnsArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Real *sG = reinterpret_cast<Real *>(smem_raw);                 // [2][G * n]
    constexpr int U = DS_RET_U;
    const int n = a.n, A = a.G * n, tid = threadIdx.x;
    const int le = tid / n, i = tid - le * n;
    const int e = blockIdx.x * a.G + le;
    const bool active = tid < A && e < a.E;
    const size_t EN = (size_t)a.E * n;
    const size_t g = active ? (size_t)e * n + i : 0;
    const int kp1 = (KP1 > 0) ? KP1 : a.k + 1;
    const Real disc = (Real)a.discount;
    const Real *r_tr = reinterpret_cast<const Real *>(a.r_tr);
    const Real *base = reinterpret_cast<const Real *>(a.base);
    Real *ret = reinterpret_cast<Real *>(a.ret), *adv = reinterpret_cast<Real *>(a.adv);
    Real Gn = 0;
    bool started = false;
    int par = 0;
    for (int tb = a.T; tb > 0; tb -= U) {                          // steps tb-1 .. tb-U (those >= 0)
        Real r[U], v[U];
        int code[U];
        int nj[U][KP1 > 0 ? KP1 : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = tb - 1 - u;
            code[u] = 2; r[u] = 0; v[u] = 0;
            if (active && t >= 0) {
                const size_t at = (size_t)t * EN + g;
                code[u] = a.fin_tr[(size_t)t * a.E + e];
                r[u] = r_tr[at];
                if (base) v[u] = base[at];
                if (KP1 > 0) {
#pragma unroll
                    for (int m = 0; m < (KP1 > 0 ? KP1 : 1); ++m) nj[u][m] = a.Ni_tr[at * KP1 + m];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = tb - 1 - u;
            if (t < 0) break;                                      // uniform
            const bool exec = active && code[u] != 2;
            Real Gt = 0;
            if (exec) {
                Gt = started ? add_rn(mul_rn(Gn, disc), r[u]) : r[u];       // :306-309
                Gn = Gt; started = true;
            }
            if (active) sG[par * A + tid] = Gt;
            __syncthreads();
            if (active) {
                const size_t at = (size_t)t * EN + g;
                Real sum = 0;                                               // :339
                int c = 0;
                if (exec) {
                    const Real *sg = sG + par * A + le * n;
                    if (KP1 > 0) {
#pragma unroll
                        for (int m = 0; m < (KP1 > 0 ? KP1 : 1); ++m) {
                            const int j = nj[u][m];
                            if (j >= 0) { sum = add_rn(sum, sub_rn(sg[j], v[u])); ++c; }   // :344-345
                        }
                    } else {
                        for (int m = 0; m < kp1; ++m) {
                            const int j = a.Ni_tr[at * kp1 + m];
                            if (j >= 0) { sum = add_rn(sum, sub_rn(sg[j], v[u])); ++c; }
                        }
                    }
                }
                ret[at] = Gt;
                adv[at] = sum;
                if (a.cnt) a.cnt[at] = (uint8_t)c;
            }
            par ^= 1;
        }
    }
}

// ---------------------------------------------------------------- device-side reset
// SURVEY.md section 8f row 4 -- init_agents' random start (drone_env.py:193-205): n DISTINCT nodes of
// the lattice {(idx * pitch, jdx * pitch)}, idx-major, drawn uniformly in order (random.sample).
// Python's Mersenne stream cannot be reproduced here, so this path is stream-independent by
// design: Philox4x32-10 keyed by (seed), counter (environment, draw block, stream id); every draw
// is an unbiased index (Lemire's multiply-shift with rejection), duplicates within an environment
// are redrawn -- which is exactly the distribution of random.sample.  One warp per environment;
// its picks sit in shared memory for the (lane-parallel) duplicate test.  oracle/np_oracle.py restates the sampler
// bit for bit (tests/test_gpu_parity.py::test_device_reset_*).
struct ResetArgs {
    int E, n, d0, d1, real_bytes;
    unsigned seed_lo, seed_hi, stream;
    double pitch;
    void *pos, *vel;
    int *t;
    uint8_t *fin;
};

DS_HD void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4])
{
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <typename Real>
__global__ void __launch_bounds__(128) reset_random_kernel(const ResetArgs a)
{
    // one WARP per environment: the draws are sequential (the accepted sequence is what defines the
    // result), the duplicate test against the earlier picks is spread over the lanes
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int *picks = reinterpret_cast<int *>(smem_raw) + (size_t)w * a.n;
    const int e = blockIdx.x * (blockDim.x >> 5) + w;
    if (e >= a.E) return;                                  // whole warps leave together
    const unsigned L = (unsigned)a.d0 * (unsigned)a.d1;
    const unsigned thresh = (0u - L) % L;                  // draws whose low product word is below are biased
    unsigned rnd[4];
    unsigned block = 0;
    int have = 0;
    using V2 = typename vec2_of<Real>::type;
    V2 *pos = reinterpret_cast<V2 *>(a.pos) + (size_t)e * a.n;
    V2 *vel = reinterpret_cast<V2 *>(a.vel) + (size_t)e * a.n;
    for (int i = 0; i < a.n;) {
        if (have == 0) { philox4x32_10((unsigned)e, block++, a.stream, 0u, a.seed_lo, a.seed_hi, rnd); have = 4; }
        const unsigned long long m = (unsigned long long)rnd[4 - have] * L;
        --have;
        if ((unsigned)m < thresh) continue;                // Lemire: reject for exact uniformity
        const int node = (int)(m >> 32);
        bool dup = false;
        for (int q = lane; q < i; q += 32) dup |= (picks[q] == node);
        if (__any_sync(0xffffffffu, dup)) continue;        // without replacement
        if (lane == 0) {
            picks[i] = node;
            V2 pv, zv;
            pv.x = (Real)mul_rn((double)(node / a.d1), a.pitch);       // [idx * delta_l, jdx * delta_l] (:199)
            pv.y = (Real)mul_rn((double)(node % a.d1), a.pitch);
            zv.x = 0; zv.y = 0;
            pos[i] = pv; vel[i] = zv;
        }
        __syncwarp();
        ++i;
    }
    if (lane == 0) {
        if (a.t) a.t[e] = 0;
        if (a.fin) a.fin[e] = 0;
    }
}

// The same reset with the observation of the start state (init_agents' rewards() call, drone_env.py:
// 208-210) in the SAME launch, for n <= 32 and k = 2: after the draws the lanes of the environment's
// warp are its agents and evaluate their rows with the step kernel's code (eval_row / write_obs) on
// the positions the warp has just written to shared memory.  Bit-identical to reset_random_kernel
// followed by step_kernel(do_integrate = 0); one launch less per episode.
struct ResetObsArgs {
    ResetArgs r;
    StepArgs s;
};

template <typename Real>
__global__ void __launch_bounds__(128) reset_observe_kernel(const ResetObsArgs ra)
{
    using V2 = typename vec2_of<Real>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ResetArgs &a = ra.r;
    const StepArgs &sa = ra.s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, n = a.n;
    int *picks = reinterpret_cast<int *>(smem_raw) + (size_t)w * n;
    V2 *spos = reinterpret_cast<V2 *>(smem_raw + CtaSmem<Real>::align16(sizeof(int) * (size_t)n * (blockDim.x >> 5))) + (size_t)w * n;
    const int e = blockIdx.x * (blockDim.x >> 5) + w;
    if (e >= a.E) return;                                  // whole warps leave together
    const unsigned L = (unsigned)a.d0 * (unsigned)a.d1;
    const unsigned thresh = (0u - L) % L;
    unsigned rnd[4];
    unsigned block = 0;
    int have = 0;
    V2 *pos = reinterpret_cast<V2 *>(a.pos) + (size_t)e * n;
    V2 *vel = reinterpret_cast<V2 *>(a.vel) + (size_t)e * n;
    for (int i = 0; i < n;) {
        if (have == 0) { philox4x32_10((unsigned)e, block++, a.stream, 0u, a.seed_lo, a.seed_hi, rnd); have = 4; }
        const unsigned long long m = (unsigned long long)rnd[4 - have] * L;
        --have;
        if ((unsigned)m < thresh) continue;
        const int node = (int)(m >> 32);
        bool dup = false;
        for (int q = lane; q < i; q += 32) dup |= (picks[q] == node);
        if (__any_sync(0xffffffffu, dup)) continue;
        if (lane == 0) {
            picks[i] = node;
            V2 pv, zv;
            pv.x = (Real)mul_rn((double)(node / a.d1), a.pitch);
            pv.y = (Real)mul_rn((double)(node % a.d1), a.pitch);
            zv.x = 0; zv.y = 0;
            pos[i] = pv; vel[i] = zv; spos[i] = pv;
        }
        __syncwarp();
        ++i;
    }
    if (lane == 0) {
        if (a.t) a.t[e] = 0;
        if (a.fin) a.fin[e] = 0;
    }
    // ---- rewards() on the start state: lane = agent
    const ParamsR<Real> P(sa);
    int nc = 0;
    if (lane < n) {
        const int i = lane;
        const AgentConst<Real> c = load_agent_const<Real>(sa.c, i);
        const V2 p = spos[i];
        RowResult<Real, 2> o;
        eval_row<Real, 2>(o, n, i, p.x, p.y, c, spos, reinterpret_cast<const Real *>(sa.c.delta),
                          reinterpret_cast<const Real *>(sa.c.radius), P, sa.c.logtab);
        const size_t g = (size_t)e * n + i;
        reinterpret_cast<Real *>(sa.r)[g] = o.r;
        reinterpret_cast<Real *>(sa.tr)[g] = o.tr;
        // zero velocities: the 5-column observation reads them from the frame's velocity array
        write_obs<Real, 2>(o, i, p.x, p.y, c, spos, reinterpret_cast<const V2 *>(a.vel) + (size_t)e * n,
                           reinterpret_cast<const Real *>(sa.c.radius), P, reinterpret_cast<Real *>(sa.z), sa.Ni, g);
        nc = o.ncoll;
    }
    nc = __reduce_add_sync(0xffffffffu, nc);
    if (lane == 0) sa.ncoll[e] = nc;                       // :284
}

// Deterministic sum over environments of agg[E][4] -> out[0..3]; out[4] = E.
__global__ void __launch_bounds__(1024) reduce_agg_kernel(const double *__restrict__ agg, int E,
                                                          double *__restrict__ out)
{
    __shared__ double s[4][1024];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int e = threadIdx.x; e < E; e += 1024) {
        const double4 v = reinterpret_cast<const double4 *>(agg)[e];
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    }
    s[0][threadIdx.x] = a0; s[1][threadIdx.x] = a1; s[2][threadIdx.x] = a2; s[3][threadIdx.x] = a3;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w)
            for (int c = 0; c < 4; ++c) s[c][threadIdx.x] += s[c][threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x < 4) out[threadIdx.x] = s[threadIdx.x][0];
    if (threadIdx.x == 4) out[4] = (double)E;
}
#endif  // __CUDACC__

}  // namespace ds
