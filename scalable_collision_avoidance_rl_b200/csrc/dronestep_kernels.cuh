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
//           per-agent constant -- 6 instructions, no sqrt;
//   pass 2  only the NEAR pairs take the exact path: sqrt, clip, zero rule,
//           division, log, collision test, k-nearest insert.
// In the rollout kernel the near pairs of all rows of the CTA are compacted into a
// shared-memory work list (row-contiguous, ascending j) and evaluated by ALL
// threads of the CTA, one pair per thread per round, so that the expensive part
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
DS_HD double log_r(double x, const LogTabEntry *__restrict__ tab)
{
    const long long b = double_bits(x);
    const int hi = (int)(b >> 32);
    if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return log(x);
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
    if (d == (Real)0) d = P.zero_eps;                              // :319-320
    po.d = d;
    po.in_disk = d <= delta_j;                                     // :328 deltas[j]
    po.logd = (Real)0;
    po.coll = false;
    if (d != ds_i) {
        // not clipped: d_ij_norm != 1, the barrier term is live (:321,327,330-332)
        if (P.log_mode == 0) {
            const Real dn = div_rn(ds_i, d);
            po.coll = dn <= (Real)0;
            po.logd = po.coll ? P.sentinel : log_r(dn, tab);
        } else {
            po.coll = (ds_i > (Real)0) ? (d < (Real)0) : (ds_i == (Real)0);
            po.logd = po.coll ? P.sentinel : sub_rn(log_ds_i, log_r(fabs(d), tab));
        }
    }
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

// Branch-free insertion of a REAL candidate (d < d_safety[i]) into the sorted list: slot m takes
// the old slot m-1 when the candidate sorts before it, the candidate when it sorts before the old
// slot m only.  Slots beyond k hold further sorted candidates and are never read.
template <typename Real, int K>
DS_HD void topk_insert(RowResult<Real, K> &o, Real d, int j)
{
    constexpr int CAP = RowResult<Real, K>::CAP;
    bool lt[CAP];
#pragma unroll
    for (int m = 0; m < CAP; ++m) lt[m] = d < o.td[m] || (d == o.td[m] && j < o.tj[m]);
#pragma unroll
    for (int m = CAP - 1; m > 0; --m) {
        o.td[m] = lt[m - 1] ? o.td[m - 1] : (lt[m] ? d : o.td[m]);
        o.tj[m] = lt[m - 1] ? o.tj[m - 1] : (lt[m] ? j : o.tj[m]);
    }
    o.td[0] = lt[0] ? d : o.td[0];
    o.tj[0] = lt[0] ? j : o.tj[0];
}

// The same row from its segment of the evaluated work list (rollout kernel): ent[q] / res[q],
// q < cnt, ascending j.  Real candidates (d < d_safety[i], self included) are sorted by
// (d, j); the clipped agents all sit at exactly d_safety[i] behind them, so the free slots are
// filled with the lowest clipped indices directly.
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
    Real sum_local = 0, sum_all = 0;
    int ncoll = 0, nreal = 0;
    // j == i (:323-325): dist = 0, d_ii = min(-2 l_i, d_safety[i]), d_norm = 1 -> no barrier term
    const Real raw_ii = sub_rn(sub_rn((Real)0, c.radius), c.radius);
    const Real d_ii = (c.ds < raw_ii) ? c.ds : raw_ii;
    int cnt_nd = c.clipcnt + ((d_ii <= c.delta) ? 1 : 0);               // :328 deltas[j], j == i
    const bool self_unclipped = (d_ii != c.ds);
    if (d_ii < c.ds) { o.td[0] = d_ii; o.tj[0] = i; nreal = 1; }
    unsigned unclipped_lo = (self_unclipped && i < 32) ? (1u << i) : 0u;   // agents j < 32 only
    for (int q = 0; q < cnt; ++q) {
        const unsigned w = ent[q];
        const V2 dv = res[q];
        const int j = (int)(w & 1023u);
        cnt_nd += (int)((w >> 12) & 3u) - 1;                            // replaces the clipped-pair count
        if (dv.x != c.ds) {                                             // not clipped (NaN included)
            if (j < 32) unclipped_lo |= 1u << j;
            ncoll += (int)((w >> 11) & 1u);
            sum_all = add_rn(sum_all, dv.y);                                                    // :283
            sum_local = add_rn(sum_local, mul_rn(dv.y, ((w >> 10) & 1u) ? (Real)1 : (Real)0));  // :282
            if (dv.x < c.ds) { topk_insert<Real, K>(o, dv.x, j); ++nreal; }
        }
    }
    // free slots <- lowest clipped indices (d = d_safety[i] for all of them: index order)
    unsigned cm = ~unclipped_lo;
    if (n < 32) cm &= (1u << n) - 1u;
    int hi = 32;                                                        // next index to test when n > 32
#pragma unroll
    for (int m = 0; m < CAP; ++m) {
        if (m <= kk && m >= nreal) {
            int cj = n;
            if (cm) {
                cj = lowest_bit(cm);
                cm &= cm - 1;
            } else {
                for (; hi < n && cj == n; ++hi) {                       // rare: >= 32 - k agents unclipped
                    bool unclipped = (hi == i) && self_unclipped;
                    for (int q = 0; q < cnt; ++q)
                        if ((int)(ent[q] & 1023u) == hi && res[q].x != c.ds) unclipped = true;
                    if (!unclipped) cj = hi;
                }
            }
            if (cj < n) { o.td[m] = c.ds; o.tj[m] = cj; }
        }
    }
    RowAcc<Real> acc;
    acc.sum_local = sum_local; acc.sum_all = sum_all; acc.ncoll = ncoll; acc.cnt_nd = cnt_nd;
    acc.self_unclipped = self_unclipped;
    row_end<Real, K>(o, acc, xi, yi, c, P);
}

// Write z_i (k+1 rows) and Ni_i (drone_env.py:344-397) for global agent index g.
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
    int nn = 1;
    Real ghx = 0, ghy = 0;
    bool ghost_ready = false;
    if (P.simplify) {
        V2 *zr = reinterpret_cast<V2 *>(z + g * (size_t)(kk + 1) * 2);
        V2 v; v.x = o.zx; v.y = o.zy;
        zr[0] = v;
#pragma unroll
        for (int kth = 1; kth < CAP; ++kth) {
            if (kth <= kk) {
                const int j = o.tj[kth];
                if (kth <= o.in_range) {                              // :362-368
                    const V2 pj = s_pos[j];
                    v.x = sub_rn(pj.x, xi); v.y = sub_rn(pj.y, yi);
                    nl[nn++] = j;
                } else {                                              // :383-386
                    if (!ghost_ready) {
                        const Real zn = sqrt_rn(fma_rn(o.zy, o.zy, mul_rn(o.zx, o.zx)));
                        ghx = mul_rn(mul_rn(div_rn(o.zx, zn), c.delta), P.ghost);
                        ghy = mul_rn(mul_rn(div_rn(o.zy, zn), c.delta), P.ghost);
                        ghost_ready = true;
                    }
                    v.x = ghx; v.y = ghy;
                }
                zr[kth] = v;
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
                Real a, bq;
                if (kth <= o.in_range) {
                    const V2 pj = s_pos[j];
                    a = sub_rn(pj.x, xi); bq = sub_rn(pj.y, yi);
                    nl[nn++] = j;
                } else {
                    if (!ghost_ready) {
                        const Real zn = sqrt_rn(fma_rn(o.zy, o.zy, mul_rn(o.zx, o.zx)));
                        ghx = mul_rn(mul_rn(div_rn(o.zx, zn), c.delta), P.ghost);
                        ghy = mul_rn(mul_rn(div_rn(o.zy, zn), c.delta), P.ghost);
                        ghost_ready = true;
                    }
                    a = ghx; bq = ghy;
                }
                const V2 vj = s_vel[j];                               // zj = state[j,:].copy() (:367,385)
                Real *row = zr + kth * 5;
                row[0] = a; row[1] = bq; row[2] = vj.x; row[3] = vj.y; row[4] = s_radius[j * cs];
            }
        }
    }
    for (; nn <= kk; ++nn) nl[nn] = -1;
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
    if (active) {
        c = load_agent_const<Real>(a.c, i);
        V2 p = reinterpret_cast<const V2 *>(a.pos)[g];
        V2 v;
        if (a.do_integrate) {
            const V2 u = reinterpret_cast<const V2 *>(a.act)[g];
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

// ---------------------------------------------------------------- rollout kernel
// Shared memory of a rollout CTA.  Per CTA: constants [n], log table, work list [L].  Per row
// (I = TC*G*n), double buffered over chunks: action, position; single: r, true_r.  Per agent of
// a slice (A = G*n): chunk start position, last executed velocity.  Per frame (F = TC*G):
// collision count, not-at-goal flag, frame info, per-frame means.  Per environment (G): alive,
// t, executed slices of the chunk.
template <typename Real> struct RoSmem {
    using V2 = typename vec2_of<Real>::type;
    V2 *cA;                  // [n] (d_safety, log d_safety)
    V2 *cB;                  // [n] (radius, delta)
    V2 *cF;                  // [n] end point
    LogTabEntry *logtab;
    V2 *act_, *pos_, *p0, *vfin, *res;   // act_ / pos_: two buffers of I rows (chunk parity)
    int I_, G_;
    Real *r, *tr;
    unsigned *ent;
    double *mr, *mtr;
    int *cnt, *notgoal, *finfo, *alive_, *tenv_, *nexec, *lcount;   // alive_ / tenv_: per chunk parity
    __device__ V2 *act(int b) const { return act_ + b * I_; }
    __device__ V2 *pos(int b) const { return pos_ + b * I_; }
    __device__ int *alive(int b) const { return alive_ + b * G_; }
    __device__ int *tenv(int b) const { return tenv_ + b * G_; }
    __host__ __device__ static size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }
    __host__ __device__ static size_t bytes(int n, int G, int TC, int L)
    {
        const size_t A = (size_t)G * n, I = A * TC, F = (size_t)G * TC;
        return (3 * (size_t)n + 4 * I + 2 * A + (size_t)L) * sizeof(V2) + kLogTabSize * sizeof(LogTabEntry) +
               align16(2 * I * sizeof(Real)) + align16((size_t)L * sizeof(unsigned)) + 2 * F * sizeof(double) +
               align16(3 * F * sizeof(int)) + align16((5 * (size_t)G + 1) * sizeof(int));
    }
    __device__ RoSmem(unsigned char *base, int n, int G, int TC, int L)
    {
        const size_t A = (size_t)G * n, I = A * TC, F = (size_t)G * TC;
        unsigned char *p = base;
        cA = reinterpret_cast<V2 *>(p); cB = cA + n; cF = cB + n;
        I_ = (int)I; G_ = G;
        act_ = cF + n; pos_ = act_ + 2 * I;
        p0 = pos_ + 2 * I; vfin = p0 + A; res = vfin + A;
        p += (3 * (size_t)n + 4 * I + 2 * A + (size_t)L) * sizeof(V2);
        logtab = reinterpret_cast<LogTabEntry *>(p); p += kLogTabSize * sizeof(LogTabEntry);
        r = reinterpret_cast<Real *>(p); tr = r + I; p += align16(2 * I * sizeof(Real));
        ent = reinterpret_cast<unsigned *>(p); p += align16((size_t)L * sizeof(unsigned));
        mr = reinterpret_cast<double *>(p); mtr = mr + F; p += 2 * F * sizeof(double);
        cnt = reinterpret_cast<int *>(p); notgoal = cnt + F; finfo = notgoal + F; p += align16(3 * F * sizeof(int));
        alive_ = reinterpret_cast<int *>(p); tenv_ = alive_ + 2 * G; nexec = tenv_ + 2 * G; lcount = nexec + G;
    }
};

#ifndef DS_RO_MINB
#define DS_RO_MINB 2      // CTAs of 256 threads per SM the rollout kernel is compiled for (register cap)
#endif
// frame info word written by the frame's leader thread after the rows are evaluated
constexpr int kFrameExec = 1, kFrameFin = 2, kFrameLast = 4;

// T fused steps, TC time slices per chunk evaluated concurrently (see the header comment).
// NB = number of 32-agent blocks whose near masks a row keeps in registers (n <= 32 NB);
// NB == 0: any n, masks in local memory.  Element indices are 32 bit: the host splits a call
// whose T * E * n would not fit.
//
// Staging of chunk c (buffer c & 1), overlapped with the tail of chunk c - 1: the rows put their
// prefetched actions into act[], then one thread per agent integrates through the chunk's slices
// (sequential, bit-exact: A = I, B = dt I, drone_env.py:78-79,235) into pos[].
// One chunk, five barriers:
//   (c) every row: pass 1 over its frame -> near masks; a warp scan + one smem atomic per warp
//       gives the row a contiguous segment of the work list; entries written          | barrier
//   (d) all threads: one near pair per thread per round (eval_pair)                   | barrier
//   (e) every row folds its segment, finishes the row, posts r / true_r / collision count /
//       not-at-goal to its frame; actions of the next chunk -> act[]                  | barrier
//   (f) threads < F: leader of one frame each -- executed?, finished?, frame means    | barrier
//   (g) every executed row stores its outputs; threads < G accumulate the episode sums and
//       advance t / alive; threads < A integrate the next chunk                       | barrier
template <typename Real, int K, int NT, int NB>
__global__ void __launch_bounds__(NT, (NT <= 256) ? DS_RO_MINB : 1)
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
    }
    if (sizeof(Real) == 8)
        for (int idx = threadIdx.x; idx < kLogTabSize; idx += blockDim.x) sm.logtab[idx] = a.c.logtab[idx];
    const ParamsR<Real> P(a);
    const int kk = (K >= 0) ? K : a.k;
    const int A = G * n;                              // agents per slice in this CTA
    const int F = G * TC;                             // frames per chunk in this CTA
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int s = tid / A, ag = tid - s * A;          // time slice within the chunk, agent slot
    const int le = ag / n, i = ag - le * n;
    const int e = blockIdx.x * G + le;
    const bool active = (s < TC) && (e < E);
    const unsigned g = active ? (unsigned)e * n + i : 0u;
    const unsigned EN = (unsigned)E * n;
    const int fr = s * G + le;                        // frame slot of this thread
    const V2 *atab = reinterpret_cast<const V2 *>(ra.atable);
    const Real thr2 = active ? ((const Real *)a.c.thr2)[i] : (Real)0;
    const int clipcnt = active ? a.c.clipcnt[i] : 0;
    const bool agent_thread = tid < A && e < E;       // s == 0: owns agent ag across the call

    if (agent_thread) {
        sm.p0[tid] = reinterpret_cast<const V2 *>(a.pos)[g];
        sm.vfin[tid] = reinterpret_cast<const V2 *>(a.vel)[g];
        if (i == 0) { sm.alive(0)[le] = (ra.done[e] == 0) ? 1 : 0; sm.tenv(0)[le] = a.t[e]; }
    }
    // per-environment episode accumulators live in thread le (< G)
    const int e_acc = blockIdx.x * G + tid;
    const bool acc_thread = tid < G && e_acc < E;
    double acc_r = 0, acc_tr = 0, acc_c = 0, acc_s = 0;
    if (acc_thread) {
        const double *ag4 = ra.agg + (size_t)e_acc * 4;
        acc_r = ag4[0]; acc_tr = ag4[1]; acc_c = ag4[2]; acc_s = ag4[3];
    }
    bool stepped = false;

    auto load_action = [&](unsigned at) -> V2 {
        if (ra.actions) return reinterpret_cast<const V2 *>(ra.actions)[at];
        return atab[ra.aidx[at]];
    };
    // sequential integration of one agent through the nsl slices of a chunk
    auto integrate = [&](int buf, int nsl) {
        V2 p = sm.p0[tid];
        const V2 *ua = sm.act(buf) + tid;
        V2 *pa = sm.pos(buf) + tid;
        for (int q = 0; q < nsl; ++q) {
            const V2 uq = ua[q * A];
            p.x = add_rn(p.x, mul_rn(P.dt, uq.x));
            p.y = add_rn(p.y, mul_rn(P.dt, uq.y));
            pa[q * A] = p;
        }
    };
    unsigned at = (unsigned)s * EN + g;               // element index of this row at slice t0 + s
    unsigned fe = (unsigned)s * E + (unsigned)e;
    V2 u{}, u_next{};
    if (active && s < T) u = load_action(at);
    if (active && TC + s < T) u_next = load_action(at + (unsigned)TC * EN);
    // stage chunk 0
    if (active && s < T) sm.act(0)[tid] = u;
    if (tid == 0) *sm.lcount = 0;
    __syncthreads();
    if (agent_thread && sm.alive(0)[le] != 0) integrate(0, (T < TC) ? T : TC);
    __syncthreads();

    for (int t0 = 0, buf = 0; t0 < T; t0 += TC, buf ^= 1) {
        const int nsl = (T - t0 < TC) ? (T - t0) : TC;      // slices in this chunk
        const bool in_chunk = active && s < nsl;
        // (c) pass 1
        const bool valid = in_chunk && sm.alive(buf)[le] != 0;
        V2 p{};
        constexpr int NBR = (NB > 0) ? NB : 32;
        unsigned near[NBR];
        int ncnt = 0;
        const V2 *fpos = sm.pos(buf) + (tid - i);           // positions of this row's frame
        if (valid) {
            p = fpos[i];
            if (i == 0) { sm.cnt[fr] = 0; sm.notgoal[fr] = 0; }
            if (NB > 0) {
#pragma unroll
                for (int bk = 0; bk < NBR; ++bk) {
                    const int j0 = bk * 32;
                    unsigned m = 0;
                    if (j0 < n) {
                        const int jn = (n - j0 < 32) ? (n - j0) : 32;
#pragma unroll 4
                        for (int jj = 0; jj < jn; ++jj) {
                            const V2 pj = fpos[j0 + jj];
                            const Real dx = sub_rn(p.x, pj.x), dy = sub_rn(p.y, pj.y);
                            const Real d2 = fma_rn(dy, dy, mul_rn(dx, dx));
                            m |= ((d2 >= thr2) ? 0u : 1u) << jj;            // NaN -> near (exact path)
                        }
                        if ((unsigned)(i - j0) < 32u) m &= ~(1u << (i - j0));
                    }
                    near[bk] = m;
                    ncnt += __popc(m);
                }
            } else {
                for (int j0 = 0, bk = 0; j0 < n; j0 += 32, ++bk) {
                    const int jn = (n - j0 < 32) ? (n - j0) : 32;
                    unsigned m = 0;
#pragma unroll 4
                    for (int jj = 0; jj < jn; ++jj) {
                        const V2 pj = fpos[j0 + jj];
                        const Real dx = sub_rn(p.x, pj.x), dy = sub_rn(p.y, pj.y);
                        const Real d2 = fma_rn(dy, dy, mul_rn(dx, dx));
                        m |= ((d2 >= thr2) ? 0u : 1u) << jj;
                    }
                    if ((unsigned)(i - j0) < 32u) m &= ~(1u << (i - j0));
                    near[bk] = m;
                    ncnt += __popc(m);
                }
            }
        }
        // segment of the work list: exclusive scan over the warp, one atomic per warp
        int incl = ncnt;
#pragma unroll
        for (int w = 1; w < 32; w <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, w);
            if (lane >= w) incl += v;
        }
        int wbase = 0;
        if (lane == 31 && incl > 0) wbase = atomicAdd(sm.lcount, incl);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        const int seg = wbase + incl - ncnt;
        const bool listed = ncnt == 0 || seg + ncnt <= L;    // otherwise the row evaluates itself in (e)
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
                for (int q = seg; q < L && q < seg + ncnt; ++q) sm.ent[q] = kEntSkip;
            }
        }
        __syncthreads();
        // (d) pass 2: one near pair per thread per round
        {
            const int M = (*sm.lcount < L) ? *sm.lcount : L;
            for (int q = tid; q < M; q += blockDim.x) {
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
        __syncthreads();
        // (e) rows
        RowResult<Real, K> o;
        AgentConst<Real> c;
        if (valid) {
            const V2 ca = sm.cA[i], cb = sm.cB[i], cf = sm.cF[i];
            c.xF = cf.x; c.yF = cf.y; c.ds = ca.x; c.log_ds = ca.y; c.radius = cb.x; c.delta = cb.y;
            c.thr2 = thr2; c.clipcnt = clipcnt;
            if (listed)
                eval_row_from_list<Real, K>(o, n, i, p.x, p.y, c, sm.ent + seg, sm.res + seg, ncnt, P);
            else
                eval_row<Real, K>(o, n, i, p.x, p.y, c, fpos, &sm.cB[0].y, &sm.cB[0].x, P, sm.logtab, 2);
            sm.r[tid] = o.r;
            sm.tr[tid] = o.tr;
            if (o.ncoll) atomicAdd(&sm.cnt[fr], o.ncoll);
            if (!o.at_goal) sm.notgoal[fr] = 1;
        }
        if (tid == 0) *sm.lcount = 0;                        // every thread has read it in (d)
        // actions of the next chunk (read by the integration in (g), after two more barriers)
        if (active && t0 + TC + s < T) sm.act(buf ^ 1)[tid] = u_next;
        __syncthreads();
        // (f) frame leaders: a slice executes iff no earlier slice of this chunk finished the
        // episode (:248-256); frame means for the episode sums (train_problem.py:98-100)
        if (tid < F) {
            const int fs = tid / G, fle = tid - fs * G;      // slice, local environment of frame tid
            int info = 0;
            if (blockIdx.x * G + fle < E && fs < nsl && sm.alive(buf)[fle] != 0) {
                const int ft0 = sm.tenv(buf)[fle];
                bool exec = true;
                for (int q = 0; q < fs; ++q)
                    if (sm.notgoal[q * G + fle] == 0 || ft0 + q >= a.max_steps - 1) exec = false;
                if (exec) {
                    const bool fin = (sm.notgoal[tid] == 0) || (ft0 + fs >= a.max_steps - 1);
                    const bool last = fin || (fs == nsl - 1);        // last executed slice of this chunk
                    info = kFrameExec | (fin ? kFrameFin : 0) | (last ? kFrameLast : 0);
                    double sr = 0, st = 0;
                    const Real *rr = sm.r + fs * A + fle * n, *rt = sm.tr + fs * A + fle * n;
                    for (int j = 0; j < n; ++j) { sr += (double)rr[j]; st += (double)rt[j]; }
                    sm.mr[tid] = sr / n; sm.mtr[tid] = st / n;
                    if (last) sm.nexec[fle] = fs + 1 + (fin ? 0x10000 : 0);
                }
            }
            sm.finfo[tid] = info;
        }
        __syncthreads();
        // (g) stores
        if (valid) {
            const int info = sm.finfo[fr];
            if (info & kFrameExec) {
                const int nc = sm.cnt[fr];
                const bool fin = (info & kFrameFin) != 0;
                const V2 *fvel = sm.act(buf) + (tid - i);
                if (ra.pos_tr) reinterpret_cast<V2 *>(ra.pos_tr)[at] = p;
                if (ra.vel_tr) reinterpret_cast<V2 *>(ra.vel_tr)[at] = u;              // :238
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
                if (i == 0) {
                    if (ra.ncoll_tr) ra.ncoll_tr[fe] = nc;
                    if (ra.fin_tr) ra.fin_tr[fe] = fin ? 1 : 0;
                }
            } else if (i == 0 && ra.fin_tr) {
                ra.fin_tr[fe] = 2;
            }
        } else if (in_chunk && i == 0 && ra.fin_tr) {
            ra.fin_tr[fe] = 2;
        }
        // one thread per agent: state after the chunk's last executed slice; next chunk's positions
        if (agent_thread && sm.alive(buf)[le] != 0) {
            const int ne = sm.nexec[le] & 0xffff;
            sm.p0[tid] = sm.pos(buf)[(ne - 1) * A + tid];
            sm.vfin[tid] = sm.act(buf)[(ne - 1) * A + tid];
            if (!(sm.nexec[le] & 0x10000) && t0 + TC < T)
                integrate(buf ^ 1, (T - t0 - TC < TC) ? (T - t0 - TC) : TC);
        }
        // episode sums of this chunk, in time order; t / alive of the next chunk (other parity)
        if (acc_thread) {
            int ne = 0;
            bool env_fin = false;
            const bool was_alive = sm.alive(buf)[tid] != 0;
            if (was_alive) {
                ne = sm.nexec[tid] & 0xffff;
                env_fin = (sm.nexec[tid] & 0x10000) != 0;
                for (int q = 0; q < ne; ++q) {
                    acc_r += sm.mr[q * G + tid]; acc_tr += sm.mtr[q * G + tid];
                    acc_c += (double)sm.cnt[q * G + tid]; acc_s += 1;
                }
                stepped = true;
            }
            sm.tenv(buf ^ 1)[tid] = sm.tenv(buf)[tid] + ne;
            sm.alive(buf ^ 1)[tid] = (was_alive && !env_fin) ? 1 : 0;
        }
        u = u_next;
        at += (unsigned)TC * EN; fe += (unsigned)TC * E;
        if (active && t0 + 2 * TC + s < T) u_next = load_action(at + (unsigned)TC * EN);   // prefetch
        __syncthreads();
    }
    const int fbuf = ((T + TC - 1) / TC) & 1;                // parity the last chunk wrote
    if (agent_thread) {
        reinterpret_cast<V2 *>(a.pos)[g] = sm.p0[tid];
        reinterpret_cast<V2 *>(a.vel)[g] = sm.vfin[tid];
    }
    if (acc_thread) {
        a.t[e_acc] = sm.tenv(fbuf)[tid];
        if (stepped) {
            if (sm.alive(fbuf)[tid] == 0) ra.done[e_acc] = 1;
            double *ag4 = ra.agg + (size_t)e_acc * 4;
            ag4[0] = acc_r; ag4[1] = acc_tr; ag4[2] = acc_c; ag4[3] = acc_s;
        }
    }
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
