// dronestep_kernels.cuh -- fused sm_100a kernels for the drone_env.step() hot path.
//
// One thread per agent.  Agents of one environment live in the same "group":
//   * n <= 32 : a group is ONE WARP holding floor(32/n) whole environments; the
//               only synchronisation is __syncwarp (no block barrier on the path);
//   * n  > 32 : a group is ONE CTA holding max(1, 256/n) whole environments.
// Positions of a group's environments are staged in shared memory once per step;
// every thread then walks its row of the n x n distance matrix out of shared
// memory (broadcast reads), so the matrix itself never exists in memory.  The
// per-row reductions (barrier sums, Delta-disk count, k+1 nearest) stay in
// registers; the two per-environment reductions (collision count, all-at-goal)
// go through one shared-memory counter and flag.
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

namespace ds {

constexpr int kMaxK = 16;

// ---------------------------------------------------------------- arithmetic
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ double log_r(double a) { return log(a); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float log_r(float a) { return logf(a); }

template <typename Real> struct vec2_of;
template <> struct vec2_of<double> { using type = double2; };
template <> struct vec2_of<float> { using type = float2; };

template <typename Real> __device__ __forceinline__ Real real_inf();
template <> __device__ __forceinline__ double real_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <> __device__ __forceinline__ float real_inf<float>() { return __int_as_float(0x7f800000); }

// np.nan_to_num (drone_env.py:287-288)
__device__ __forceinline__ double nan_to_num(double v)
{
    if (isnan(v)) return 0.0;
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v;
}
__device__ __forceinline__ float nan_to_num(float v)
{
    if (isnan(v)) return 0.0f;
    if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    return v;
}

// ---------------------------------------------------------------- arguments
struct Consts {              // device arrays of Real, length n (xF: 2n)
    const void *xF, *d_safety, *delta, *radius, *log_ds;
};

struct StepArgs {
    int E, n, k, simplify, epg, do_integrate, log_mode, max_steps;
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
    int T, n_actions;
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
    __device__ explicit ParamsR(const StepArgs &a)
        : dt((Real)a.dt), q((Real)a.q), b((Real)a.b), goal_tol((Real)a.goal_tol),
          sentinel((Real)a.sentinel), zero_eps((Real)a.zero_eps), ghost((Real)a.ghost),
          log_mode(a.log_mode), simplify(a.simplify), k(a.k) {}
};

template <typename Real> struct AgentConst {   // per-thread (row i) constants
    Real xF, yF, ds, delta, radius, log_ds;
};

template <typename Real>
__device__ __forceinline__ AgentConst<Real> load_agent_const(const Consts &c, int i)
{
    AgentConst<Real> a;
    a.xF = ((const Real *)c.xF)[2 * i];
    a.yF = ((const Real *)c.xF)[2 * i + 1];
    a.ds = ((const Real *)c.d_safety)[i];
    a.delta = ((const Real *)c.delta)[i];
    a.radius = ((const Real *)c.radius)[i];
    a.log_ds = ((const Real *)c.log_ds)[i];
    return a;
}

// ---------------------------------------------------------------- one row of the pair matrix
// Everything rewards() derives for agent i (drone_env.py:260-293) from the staged
// positions of its environment.
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

template <typename Real, int K>
__device__ __forceinline__ void eval_row(RowResult<Real, K> &o, int n, int i, Real xi, Real yi,
                                         const AgentConst<Real> &c,
                                         const typename vec2_of<Real>::type *__restrict__ s_pos,
                                         const Real *__restrict__ s_delta,
                                         const Real *__restrict__ s_radius,
                                         const ParamsR<Real> &P)
{
    const int kk = (K >= 0) ? K : P.k;
    constexpr int CAP = RowResult<Real, K>::CAP;
#pragma unroll
    for (int m = 0; m < CAP; ++m) { o.td[m] = real_inf<Real>(); o.tj[m] = 0; }

    Real sum_local = 0, sum_all = 0;
    int ncoll = 0, cnt_nd = 0;
#pragma unroll 2
    for (int j = 0; j < n; ++j) {
        const typename vec2_of<Real>::type pj = s_pos[j];
        const Real dx = sub_rn(xi, pj.x), dy = sub_rn(yi, pj.y);
        const Real dist = sqrt_rn(fma_rn(dy, dy, mul_rn(dx, dx)));       // :318 (BLAS ddot)
        const Real raw = sub_rn(sub_rn(dist, c.radius), s_radius[j]);     // :318 / :323 (dist == 0)
        Real d = (c.ds < raw) ? c.ds : raw;                               // python min(raw, d_safety[i])
        const bool self = (j == i);
        if (!self && d == (Real)0) d = P.zero_eps;                        // :319-320
        cnt_nd += (d <= s_delta[j]) ? 1 : 0;                              // :328 deltas[j]
        if (!self && d != c.ds) {
            // not clipped: d_ij_norm != 1, the barrier term is live (:321,327,330-332)
            Real logd;
            bool coll;
            if (P.log_mode == 0) {
                const Real dn = div_rn(c.ds, d);
                coll = dn <= (Real)0;
                logd = coll ? P.sentinel : log_r(dn);
            } else {
                coll = (c.ds > (Real)0) ? (d < (Real)0) : (c.ds == (Real)0);
                logd = coll ? P.sentinel : sub_rn(c.log_ds, log_r(fabs(d)));
            }
            ncoll += coll ? 1 : 0;
            sum_all = add_rn(sum_all, logd);                                                   // :283
            sum_local = add_rn(sum_local, mul_rn(logd, (d <= s_delta[j]) ? (Real)1 : (Real)0)); // :282
        }
        // running k+1 smallest (np.argsort row, :338): strict '<' keeps the lower index on ties
        if (d < o.td[kk]) {
            o.td[kk] = d; o.tj[kk] = j;
#pragma unroll
            for (int m = CAP - 1; m > 0; --m) {
                if (m <= kk && o.td[m] < o.td[m - 1]) {
                    const Real tdv = o.td[m]; o.td[m] = o.td[m - 1]; o.td[m - 1] = tdv;
                    const int tjv = o.tj[m]; o.tj[m] = o.tj[m - 1]; o.tj[m - 1] = tjv;
                }
            }
        }
    }
    const Real gx = sub_rn(c.xF, xi), gy = sub_rn(c.yF, yi);
    const Real nrm = sqrt_rn(add_rn(mul_rn(gx, gx), mul_rn(gy, gy)));     // :249,276 (axis norm, unfused)
    const Real goal = mul_rn(P.q, mul_rn(nrm, nrm));                      // :276
    o.r = -nan_to_num(add_rn(goal, mul_rn(P.b, sum_local)));              // :282,287
    o.tr = -nan_to_num(add_rn(goal, mul_rn(P.b, sum_all)));               // :283,288
    o.zx = -gx; o.zy = -gy;                                               // :357
    o.ncoll = ncoll;
    o.in_range = cnt_nd - 1;                                              // :346
    o.at_goal = nrm <= P.goal_tol;
}

// Write z_i (k+1 rows) and Ni_i (drone_env.py:344-397) for global agent index g.
template <typename Real, int K>
__device__ __forceinline__ void write_obs(const RowResult<Real, K> &o, int n, int i, Real xi, Real yi,
                                          const AgentConst<Real> &c,
                                          const typename vec2_of<Real>::type *__restrict__ s_pos,
                                          const typename vec2_of<Real>::type *__restrict__ s_vel,
                                          const Real *__restrict__ s_radius,
                                          const ParamsR<Real> &P, Real *__restrict__ z, int *__restrict__ Ni,
                                          size_t g)
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
                row[0] = a; row[1] = bq; row[2] = vj.x; row[3] = vj.y; row[4] = s_radius[j];
            }
        }
    }
    for (; nn <= kk; ++nn) nl[nn] = -1;
}

// ---------------------------------------------------------------- group plumbing
// MODE 0: group = warp (CTA of 4 warps); MODE 1: group = CTA of <= 256 threads;
// MODE 2: group = CTA of <= 1024 threads (n > 256; register-capped at 64).
template <int MODE> __device__ __forceinline__ void group_sync()
{
    if (MODE == 0) __syncwarp(); else __syncthreads();
}
constexpr int mode_max_threads(int mode) { return mode == 0 ? 128 : (mode == 1 ? 256 : 1024); }

// Shared-memory carve-up.  Block-wide: delta[n], radius[n].  Per group:
// pos[epg*n] vec2, vel[epg*n] vec2, rsum[2][epg*n] Real, cnt[2][epg] int, notgoal[2][epg] int.
template <typename Real> struct GroupSmem {
    using V2 = typename vec2_of<Real>::type;
    Real *delta, *radius;
    V2 *pos, *vel;
    Real *r, *tr;
    int *cnt, *notgoal;
    __host__ __device__ static size_t group_bytes(int epg, int n)
    {
        size_t agents = (size_t)epg * n;
        size_t b = agents * sizeof(V2) * 2 + agents * sizeof(Real) * 2 + (size_t)epg * sizeof(int) * 4;
        return (b + 15) & ~(size_t)15;
    }
    __host__ __device__ static size_t const_bytes(int n)
    {
        return (((size_t)n * sizeof(Real) * 2) + 15) & ~(size_t)15;
    }
    __device__ GroupSmem(unsigned char *base, int epg, int n, int group_in_block)
    {
        delta = reinterpret_cast<Real *>(base);
        radius = delta + n;
        unsigned char *g = base + const_bytes(n) + group_bytes(epg, n) * group_in_block;
        const size_t agents = (size_t)epg * n;
        pos = reinterpret_cast<V2 *>(g);
        vel = pos + agents;
        r = reinterpret_cast<Real *>(vel + agents);
        tr = r + agents;
        cnt = reinterpret_cast<int *>(tr + agents);
        notgoal = cnt + 2 * epg;
    }
};

// ---------------------------------------------------------------- step kernel
// drones.step() / rewards() for E environments, one launch.
template <typename Real, int K, int MODE>
__global__ void __launch_bounds__(mode_max_threads(MODE))
step_kernel(const StepArgs a)
{
    using V2 = typename vec2_of<Real>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, epg = a.epg;
    constexpr bool WARP = (MODE == 0);
    const int gib = WARP ? (threadIdx.x >> 5) : 0;
    const int lt = WARP ? (threadIdx.x & 31) : threadIdx.x;
    const int group = WARP ? (blockIdx.x * (blockDim.x >> 5) + gib) : blockIdx.x;
    GroupSmem<Real> sm(smem_raw, epg, n, gib);
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        sm.delta[idx] = ((const Real *)a.c.delta)[idx];
        sm.radius[idx] = ((const Real *)a.c.radius)[idx];
    }
    __syncthreads();

    const ParamsR<Real> P(a);
    const int le = lt / n, i = lt - le * n;
    const long long e = (long long)group * epg + le;
    const bool active = (lt < epg * n) && (e < a.E);
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
        sm.vel[lt] = v;
        if (i == 0) { sm.cnt[le] = 0; sm.notgoal[le] = 0; }
    }
    group_sync<MODE>();
    if (active) {
        RowResult<Real, K> o;
        eval_row<Real, K>(o, n, i, xi, yi, c, sm.pos + le * n, sm.delta, sm.radius, P);
        reinterpret_cast<Real *>(a.r)[g] = o.r;
        reinterpret_cast<Real *>(a.tr)[g] = o.tr;
        write_obs<Real, K>(o, n, i, xi, yi, c, sm.pos + le * n, sm.vel + le * n, sm.radius, P,
                           reinterpret_cast<Real *>(a.z), a.Ni, g);
        if (o.ncoll) atomicAdd(&sm.cnt[le], o.ncoll);
        if (!o.at_goal) sm.notgoal[le] = 1;
    }
    group_sync<MODE>();
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
// T fused steps: positions stay in registers / shared memory between steps, the
// action stream is read from HBM (prefetched one step ahead) and every step's
// outputs are streamed to the trajectory buffers.
template <typename Real, int K, int MODE>
__global__ void __launch_bounds__(mode_max_threads(MODE))
rollout_kernel(const RolloutArgs ra)
{
    using V2 = typename vec2_of<Real>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const StepArgs &a = ra.s;
    const int n = a.n, epg = a.epg, E = a.E;
    constexpr bool WARP = (MODE == 0);
    const int gib = WARP ? (threadIdx.x >> 5) : 0;
    const int lt = WARP ? (threadIdx.x & 31) : threadIdx.x;
    const int group = WARP ? (blockIdx.x * (blockDim.x >> 5) + gib) : blockIdx.x;
    GroupSmem<Real> sm(smem_raw, epg, n, gib);
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        sm.delta[idx] = ((const Real *)a.c.delta)[idx];
        sm.radius[idx] = ((const Real *)a.c.radius)[idx];
    }
    __syncthreads();

    const ParamsR<Real> P(a);
    const int kk = (K >= 0) ? K : a.k;
    const int cols = a.simplify ? 2 : 5;
    const int le = lt / n, i = lt - le * n;
    const long long e = (long long)group * epg + le;
    const bool active = (lt < epg * n) && (e < E);
    const size_t g = active ? (size_t)e * n + i : 0;
    const size_t EN = (size_t)E * n;

    AgentConst<Real> c{};
    V2 p{}, v{};
    bool alive = false;
    int tt = 0;
    if (active) {
        c = load_agent_const<Real>(a.c, i);
        p = reinterpret_cast<const V2 *>(a.pos)[g];
        v = reinterpret_cast<const V2 *>(a.vel)[g];
        alive = ra.done[e] == 0;
        tt = a.t[e];
    }
    double acc_r = 0, acc_tr = 0, acc_c = 0, acc_s = 0;
    const V2 *atab = reinterpret_cast<const V2 *>(ra.atable);

    auto load_action = [&](int t) -> V2 {
        const size_t at = (size_t)t * EN + g;
        if (ra.actions) return reinterpret_cast<const V2 *>(ra.actions)[at];
        return atab[ra.aidx[at]];
    };
    V2 u_next{};
    if (alive && ra.T > 0) u_next = load_action(0);

    for (int t = 0; t < ra.T; ++t) {
        const int buf = t & 1;
        const size_t at = (size_t)t * EN + g;
        if (alive) {
            const V2 u = u_next;
            if (t + 1 < ra.T) u_next = load_action(t + 1);   // prefetch: hides the HBM latency
            p.x = add_rn(p.x, mul_rn(P.dt, u.x));
            p.y = add_rn(p.y, mul_rn(P.dt, u.y));
            v = u;
            sm.pos[lt] = p;
            sm.vel[lt] = v;
            if (ra.pos_tr) reinterpret_cast<V2 *>(ra.pos_tr)[at] = p;
            if (ra.vel_tr) reinterpret_cast<V2 *>(ra.vel_tr)[at] = v;
            if (i == 0) { sm.cnt[buf * epg + le] = 0; sm.notgoal[buf * epg + le] = 0; }
        } else if (active && i == 0 && ra.fin_tr) {
            ra.fin_tr[(size_t)t * E + e] = 2;
        }
        group_sync<MODE>();
        RowResult<Real, K> o;
        if (alive) {
            eval_row<Real, K>(o, n, i, p.x, p.y, c, sm.pos + le * n, sm.delta, sm.radius, P);
            if (ra.r_tr) reinterpret_cast<Real *>(ra.r_tr)[at] = o.r;
            if (ra.tr_tr) reinterpret_cast<Real *>(ra.tr_tr)[at] = o.tr;
            if (ra.z_tr)
                write_obs<Real, K>(o, n, i, p.x, p.y, c, sm.pos + le * n, sm.vel + le * n, sm.radius, P,
                                   reinterpret_cast<Real *>(ra.z_tr) + (size_t)t * EN * (kk + 1) * cols,
                                   ra.Ni_tr + (size_t)t * EN * (kk + 1), g);
            sm.r[lt] = o.r;
            sm.tr[lt] = o.tr;
            if (o.ncoll) atomicAdd(&sm.cnt[buf * epg + le], o.ncoll);
            if (!o.at_goal) sm.notgoal[buf * epg + le] = 1;
        }
        group_sync<MODE>();
        if (alive) {
            const int nc = sm.cnt[buf * epg + le];
            const bool fin = (sm.notgoal[buf * epg + le] == 0) || (tt >= a.max_steps - 1);
            tt += 1;
            if (i == 0) {
                if (ra.ncoll_tr) ra.ncoll_tr[(size_t)t * E + e] = nc;
                if (ra.fin_tr) ra.fin_tr[(size_t)t * E + e] = fin ? 1 : 0;
                double sr = 0, st = 0;
                for (int j = 0; j < n; ++j) { sr += (double)sm.r[le * n + j]; st += (double)sm.tr[le * n + j]; }
                acc_r += sr / n; acc_tr += st / n; acc_c += nc; acc_s += 1;   // train_problem.py:98-100
            }
            if (fin || t == ra.T - 1) {
                // last executed step: leave the step()-style outputs in the live buffers
                reinterpret_cast<Real *>(a.r)[g] = o.r;
                reinterpret_cast<Real *>(a.tr)[g] = o.tr;
                write_obs<Real, K>(o, n, i, p.x, p.y, c, sm.pos + le * n, sm.vel + le * n, sm.radius, P,
                                   reinterpret_cast<Real *>(a.z), a.Ni, g);
                if (i == 0) { a.ncoll[e] = nc; a.fin[e] = fin ? 1 : 0; }
            }
            if (fin) alive = false;
        }
    }
    if (active) {
        reinterpret_cast<V2 *>(a.pos)[g] = p;
        reinterpret_cast<V2 *>(a.vel)[g] = v;
        if (i == 0) {
            a.t[e] = tt;
            if (acc_s > 0) {
                if (!alive) ra.done[e] = 1;
                double *ag = ra.agg + (size_t)e * 4;
                ag[0] += acc_r; ag[1] += acc_tr; ag[2] += acc_c; ag[3] += acc_s;
            }
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

}  // namespace ds
