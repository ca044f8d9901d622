// dronestep_rollout2.cuh -- warp-per-(environment, time segment) rollout kernel (the benchmarked
// kernel of round 2; DESIGN.md section 4.3 has the measurements behind every choice below).
//
// Same observable behaviour as rollout_kernel (dronestep_kernels.cuh): T fused steps of
// drones.step() (reference drone_env.py:214-258) per environment, every per-step output recorded,
// early termination (drone_env.py:248-256), finished codes, done flags, episode sums.  Different
// decomposition, chosen after the round-1 profile showed the CTA-wide version bound by instruction
// issue (58 % issue slots, 1.07 barrier-stall cycles per issue, 1410 warp-instructions per 32
// agent-steps of which only ~12 % were fp64 arithmetic):
//
//   * One CTA per environment, ONE WARP per TIME SEGMENT of the call: the warp first integrates the
//     steps in front of its segment (the prefix pass: integrator + episode-end test only), then
//     walks the segment in chunks of TCW = floor(32 / N) consecutive time slices: lane = (slice s,
//     agent i), one row of the pair matrix per lane.  Nothing is shared between warps but the
//     per-CTA constants, so the chunk loop has NO CTA barrier -- only __syncwarp between its phases.
//   * N (agents) is a template parameter: slice / agent of a lane, trip counts, shared-memory
//     offsets are compile-time; the phases are written as straight-line predicated code with
//     warp-uniform loop bounds (no divergent branches in the chunk loop's hot phases).
//   * The episode's action stream is moved by the TMA unit: per chunk ONE cp.async.bulk.tensor.2d
//     (tensor map over [T][E * N * 2], box {N * 2, TCW}; mbarrier complete_tx) brings the chunk's
//     [TCW][N][2] block into a 2-stage ring, one chunk ahead of its use; no thread holds an action
//     in a register while it waits.  Fallbacks: one 1-D cp.async.bulk per slice; a lane-load form
//     for index mode / unaligned blocks.
//   * Every agent has the same radius, d_safety and Delta in every configuration the reference
//     can construct with a scalar delta (drone_env.py:75,85-91,153 on a circle formation), and
//     then d_ij, log(d_safety/d_ij), the collision test and the Delta-disk test are symmetric in
//     (i, j): each UNORDERED near pair is evaluated once (one sqrt / log instead of two) and its
//     result is scattered to both rows of a dense [row][partner column] table (small N) or to the
//     rows' result segments.  Non-uniform constants, other n, k != 2 and the 5-column observation
//     take rollout_kernel (the host decides per handle; ds_rollout_kernel_name tells which).
//   * A row folds its partners in ascending j (sums; k nearest with strict "<" insertion = stable
//     argsort order); collision counts are posted by the pair lanes (rare), the Delta-disk count
//     follows from the k nearest (uniform Delta), so the fold loop carries only what needs order.
//   * Frames too dense for the list (more near pairs than its capacity) are evaluated in groups
//     of rows, every pair of a group's rows exactly, through the same fold -- correct for any
//     density, never taken at the BASELINE densities.
//   * 72 registers (everything a lane carries across chunks lives in the warp's shared-memory
//     block) and <= 32 KB of shared memory per 4-warp CTA: 7 CTAs per SM.
//
// The arithmetic (operation order of eval_pair / row_end / write_obs) is that of
// dronestep_kernels.cuh, so this kernel, rollout_kernel and step_kernel agree bit for bit.
#pragma once
#include <type_traits>
#include <cuda.h>      // CUtensorMap (type only; the encoder is resolved at run time by the host side)
#include "dronestep_kernels.cuh"

namespace ds {

#if defined(__CUDACC__)

struct Ro2Args {
    RolloutArgs ra;
    // uniform per-agent constants
    double ds, delta, radius, log_ds, inv_ds;
    double delta_eff;       // Delta, or +inf when d_safety <= Delta (every agent is inside every Delta disk)
    double d_ii;            // the clipped 'distance' of an agent to itself: min(d_safety, -2 radius) (:318-325)
    double goal_t2;         // largest x with sqrt_rn(x) <= goal_tol: the goal test without the square root
    float thr2f;            // pass-1 threshold of the packed-f32 filter (squared, with margin)
    int act_mode;           // action staging: 0 lane loads, 1 cp.async.bulk per slice, 2 one 2-D TMA tile per chunk
    int seg_c0[9];          // first chunk of every time segment (seg_c0[S] = number of chunks): later segments are
                            // shorter, by the cost of the prefix pass in front of them, so that a CTA's warps finish together
};

constexpr int kRo2Stages = 2;       // action ring depth (chunks): chunk c + 1 is in flight during chunk c
constexpr int kRo2Threads = 256;    // up to 8 warps = 8 time segments of one environment per CTA
constexpr int kRo2MaxSeg = 8;
#ifndef DS_RO2_MINCTAS
#define DS_RO2_MINCTAS 3            // launch bound for 256-thread CTAs
#endif
#ifndef DS_RO2_MAXNREG
#define DS_RO2_MAXNREG 72           // 7 CTAs of 4 warps per SM (no spills): 4096 environments = 3.95 waves of 7 x 148 CTAs.
                                    // At 80 registers (6 CTAs, 4.6 waves: a 40 %-empty last wave) 2.7 % slower
#endif

__host__ __device__ inline size_t ro2_align16(size_t b) { return (b + 15) & ~(size_t)15; }

// Per-warp shared memory (compile-time layout).
template <typename Real, int N> struct alignas(128) Ro2Warp {
    using V2 = typename vec2_of<Real>::type;
    static constexpr int TCW = 32 / N, RW = TCW * N, HP = (N + 1) / 2;
    // TABLE: results in a dense [row][partner] table and a list with room for every pair of the
    // chunk (small N); otherwise row-contiguous segments and a list of bounded capacity, with the
    // exact group-wise evaluation for frames that overflow it
    // (table row: one column per partner -- partner j of agent i in column j - (j > i) -- at a stride
    // RS of an odd number of entries, so that the lanes of a quarter / half warp -- one row each, the
    // same column -- fall into different banks)
    static constexpr int RS = (N - 1) | 1;
    static constexpr bool TABLE = (size_t)RW * RS * sizeof(V2) <= 5120;
    static constexpr int LW = TABLE ? RW * RS : ((RW * (N - 1) < 128) ? RW * (N - 1) : 128);        // result slots
    static constexpr int LU = TABLE ? RW * (N - 1) / 2 : ((RW * (N - 1) / 2 < 64) ? RW * (N - 1) / 2 : 64);   // list entries
    static constexpr int GR = LW / (N - 1);                                       // rows per group (overflowing frames)
    static constexpr int ASTR = (int)(((RW * sizeof(V2) + 127) / 128) * 128 / sizeof(V2));   // stage stride: 128-byte aligned (TMA)
    V2 act[kRo2Stages][ASTR]; // action ring: [slice][agent] of a chunk, as in global memory
    V2 pos[RW];               // positions of the chunk's rows (row = lane)
    V2 pend[N];               // agent's position after the previous chunk
    V2 acc[RW];               // running episode sums (r, true_r) of the lane's rows
    int sumc[TCW];            // running collision count of each slice's frames
    V2 res[LW + 1];           // (d, log term) per ordered near pair, row-contiguous, ascending j; [LW]: (d_safety, 0)
    float4 posf[TCW * HP];    // packed f32 copies (x_2q, x_2q+1, y_2q, y_2q+1) per frame
    uint2 rowinfo[TABLE ? 1 : 32];   // (near mask, first result slot) of each row (segment layout only)
    using Ent = typename std::conditional<TABLE, unsigned short, unsigned>::type;
    Ent ent[LU + 32];         // unordered near pairs; [LU + lane]: where a lane that ran out of pairs stores
    unsigned umask[32];       // near AND not clipped (pair lanes clear the rare clipped-near bits)
    int cnt[TCW + 1];         // collision count per frame (slice); [TCW]: the call's record mask
    unsigned long long mbar[kRo2Stages];
};

__device__ __forceinline__ unsigned ro2_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ro2_mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ro2_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ro2_mbar_expect(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ro2_smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy (1-D, no tensor map) global -> shared; the bytes are counted on the mbarrier
__device__ __forceinline__ void ro2_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ro2_smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(ro2_smem_u32(bar)) : "memory");
}
// TMA tile load (2-D tensor map over the action stream: rows = time steps, row = [E][N][2] Reals):
// one instruction brings the [TCW][N][2] block of a chunk; rows beyond T are zero-filled and counted
__device__ __forceinline__ void ro2_tma_g2s_2d(void *dst, const void *tmap, int x, int y, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     ro2_smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(x), "r"(y), "r"(ro2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ro2_mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned addr = ro2_smem_u32(bar);
    unsigned done = 0;
    for (unsigned spins = 0; !done; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (spins > (1u << 26)) __trap();                  // a lost copy must fail loudly, not hang the device
    }
}

// distance_data() for one pair when radius, d_safety and Delta are the same for every agent
// (drone_env.py:314-332): the result holds for (i, j) and (j, i).  Operation order of eval_pair;
// log_mode 2 replaces the division by a multiplication with the rounded reciprocal of d_safety:
// log(d_safety / d) = -log(d * (1 / d_safety)),  <= 3.4e-16 absolute difference.
template <typename Real>
__device__ __forceinline__ void ro2_eval_pair(Real &d_out, Real &logd_out, bool &coll_out, Real xi, Real yi, Real xj,
                                              Real yj, Real ds, Real rad, Real log_ds, Real inv_ds, int log_mode,
                                              Real zero_eps, Real sentinel, const LogTabEntry *__restrict__ tab)
{
    const Real dx = sub_rn(xi, xj), dy = sub_rn(yi, yj);
    const Real dist = sqrt_rn(fma_rn(dy, dy, mul_rn(dx, dx)));     // :318 (BLAS ddot)
    const Real raw = sub_rn(sub_rn(dist, rad), rad);               // :318
    Real d = (ds < raw) ? ds : raw;                                // python min(raw, d_safety[i])
    d = (d == (Real)0) ? zero_eps : d;                             // :319-320
    const bool live = d != ds;
    // one table-driven log; its argument and the use of its value depend on the mode
    bool coll;
    Real x;
    if (log_mode == 0) {
        x = div_rn(ds, d);
        coll = x <= (Real)0;
    } else {
        coll = d < (Real)0;                                        // d_safety > 0 on this path
        x = (log_mode == 1) ? fabs(d) : mul_rn(d, inv_ds);
    }
    const Real one = (log_mode == 1) ? fabs(ds) : (Real)1;
    Real lg = log_r((live && !coll) ? x : one, tab);
    lg = (log_mode == 0) ? lg : ((log_mode == 1) ? sub_rn(log_ds, lg) : -lg);
    d_out = d;
    coll_out = live && coll;
    logd_out = live ? (coll ? sentinel : lg) : (Real)0;
}

// Pass 1 over the N agents of the row's frame on the packed-f32 pipe (see pass1_block_f32x2).
template <int N>
__device__ __forceinline__ unsigned ro2_pass1(const float4 *__restrict__ fp, float nx, float ny, float thr2f)
{
    constexpr int HP = (N + 1) / 2;
    const float2 NX = make_float2(nx, nx), NY = make_float2(ny, ny);
    const float2 TH = make_float2(thr2f, thr2f), M1 = make_float2(-1.0f, -1.0f);
    unsigned m = 0;
#pragma unroll
    for (int q = 0; q < HP; ++q) {
        const float4 pq = fp[q];
        const float2 a = __fadd2_rn(make_float2(pq.x, pq.y), NX);
        const float2 b = __fadd2_rn(make_float2(pq.z, pq.w), NY);
        const float2 t = __ffma2_rn(__ffma2_rn(b, b, __fmul2_rn(a, a)), M1, TH);   // thr2 - d2
        m = __funnelshift_l(__float_as_uint(t.x), m, 1);
        m = __funnelshift_l(__float_as_uint(t.y), m, 1);
    }
    // m: NOT-near bits of 2 HP agents, agent jj at bit 2 HP - 1 - jj (a pad agent sits at 1e30: not near)
    m = __brev(~m) >> (32 - 2 * HP);
    return (N < 32) ? (m & ((1u << N) - 1u)) : m;
}

// Fold of a row's result segment in ascending j: sums (:282-283) and the K nearest partners
// (strict "<": equal distances stay in index order).  Warp-uniform trip count; lanes with fewer
// partners fold (d_safety, +0) -- never a candidate, adds nothing -- in the tail.
template <typename Real, int K>
__device__ __forceinline__ void ro2_fold(const typename vec2_of<Real>::type *__restrict__ rp, unsigned mm, Real ds, Real delta_eff,
                                         Real &sum_all, Real &sum_loc, Real (&nd)[K], int (&nj)[K])
{
    using V2 = typename vec2_of<Real>::type;
    const int iters = __reduce_max_sync(0xffffffffu, __popc(mm));
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const bool on = mm != 0;
        const int j = __ffs((int)mm) - 1;
        mm &= mm - 1;
        V2 dv; dv.x = ds; dv.y = 0;
        if (on) { dv = *rp; ++rp; }
        sum_all = add_rn(sum_all, dv.y);                                                       // :283
        sum_loc = add_rn(sum_loc, mul_rn(dv.y, (dv.x <= delta_eff) ? (Real)1 : (Real)0));      // :282
        bool lt[K];
#pragma unroll
        for (int q = 0; q < K; ++q) lt[q] = dv.x < nd[q];           // false for NaN; clipped (== d_safety) never enters
#pragma unroll
        for (int q = K - 1; q > 0; --q) {
            nd[q] = lt[q - 1] ? nd[q - 1] : (lt[q] ? dv.x : nd[q]);
            nj[q] = lt[q - 1] ? nj[q - 1] : (lt[q] ? j : nj[q]);
        }
        nd[0] = lt[0] ? dv.x : nd[0];
        nj[0] = lt[0] ? j : nj[0];
    }
}

// The same fold over a row of the dense [row][partner] table: partner j's result at rp[j].  Two
// partners per iteration (both loads in flight together).
template <typename Real, int K>
__device__ __forceinline__ void ro2_fold_step(Real dx, Real dy, int j, Real delta_eff, Real &sum_all, Real &sum_loc,
                                              Real (&nd)[K], int (&nj)[K])
{
    sum_all = add_rn(sum_all, dy);                                                         // :283
    sum_loc = add_rn(sum_loc, mul_rn(dy, (dx <= delta_eff) ? (Real)1 : (Real)0));          // :282
    bool lt[K];
#pragma unroll
    for (int q = 0; q < K; ++q) lt[q] = dx < nd[q];                 // false for NaN; clipped (== d_safety) never enters
#pragma unroll
    for (int q = K - 1; q > 0; --q) {
        nd[q] = lt[q - 1] ? nd[q - 1] : (lt[q] ? dx : nd[q]);
        nj[q] = lt[q - 1] ? nj[q - 1] : (lt[q] ? j : nj[q]);
    }
    nd[0] = lt[0] ? dx : nd[0];
    nj[0] = lt[0] ? j : nj[0];
}
// `neutral` holds (d_safety, +0) for the whole call: never a candidate, adds nothing -- what a lane
// that has run out of partners folds (one address for all of them: a broadcast, no bank conflict).
// The fold runs over the row's COLUMNS (ascending column = ascending partner index) and leaves
// column numbers in nj; the caller turns them into partner indices.
template <typename Real, int K>
__device__ __forceinline__ void ro2_fold_table(const typename vec2_of<Real>::type *__restrict__ rp,
                                               const typename vec2_of<Real>::type *__restrict__ neutral, unsigned mm, int i,
                                               Real delta_eff, Real &sum_all, Real &sum_loc, Real (&nd)[K], int (&nj)[K])
{
    using V2 = typename vec2_of<Real>::type;
    const int iters = (__reduce_max_sync(0xffffffffu, __popc(mm)) + 1) >> 1;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const int j0 = __ffs((int)mm) - 1;
        const V2 *p0 = mm ? rp + j0 : neutral;
        mm &= mm - 1;
        const int j1 = __ffs((int)mm) - 1;
        const V2 *p1 = mm ? rp + j1 : neutral;
        mm &= mm - 1;
        const V2 d0 = *p0, d1 = *p1;
        ro2_fold_step<Real, K>(d0.x, d0.y, j0, delta_eff, sum_all, sum_loc, nd, nj);
        ro2_fold_step<Real, K>(d1.x, d1.y, j1, delta_eff, sum_all, sum_loc, nd, nj);
    }
}

// -np.nan_to_num(v) (drone_env.py:287-288); the rewards are finite unless a position is not
DS_HD double ro2_neg_nan_to_num(double v)
{
#if defined(__CUDA_ARCH__)
    if ((__double2hiint(v) & 0x7ff00000) == 0x7ff00000) v = nan_to_num(v);      // inf or NaN: exponent all ones
#else
    if (!(fabs(v) <= 1.7976931348623157e308)) v = nan_to_num(v);
#endif
    return -v;
}
DS_HD float ro2_neg_nan_to_num(float v)
{
#if defined(__CUDA_ARCH__)
    if ((__float_as_int(v) & 0x7f800000) == 0x7f800000) v = nan_to_num(v);
#else
    if (!(fabsf(v) <= 3.4028234663852886e38f)) v = nan_to_num(v);
#endif
    return -v;
}

// Shared memory of a CTA in front of its per-warp blocks: constants and the segments' episode sums.
template <typename Real, int N> struct alignas(128) Ro2Cta {
    using V2 = typename vec2_of<Real>::type;
    V2 cF[N];                          // end points
    double part[kRo2MaxSeg][4];        // per-segment episode sums: r, true_r, collisions, steps
    int seg_fin[kRo2MaxSeg];           // the episode ended inside this segment
    LogTabEntry logtab[sizeof(Real) == 8 ? kLogTabSize : 1];   // (read through L1 from global memory instead: 3.5 % slower)
};

template <typename Real, int N, int K>
__global__ void __maxnreg__(DS_RO2_MAXNREG)
rollout2_kernel(const Ro2Args A, const __grid_constant__ CUtensorMap tmap)
{
    using V2 = typename vec2_of<Real>::type;
    using WS = Ro2Warp<Real, N>;
    using CS = Ro2Cta<Real, N>;
    constexpr int TCW = WS::TCW, RW = WS::RW, HP = WS::HP;
    constexpr unsigned fullN = (N >= 32) ? 0xffffffffu : ((1u << N) - 1u);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const RolloutArgs &ra = A.ra;
    const StepArgs &a = ra.s;
    const int E = a.E, T = ra.T;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // (opaque to the optimiser: otherwise every use of a per-warp address re-derives them from %tid)
    asm volatile("" : "+r"(lane), "+r"(warp));
    // One environment per CTA, one time segment per warp.  (Measured alternative: CTAs of four warps
    // holding 4 / S environments of S segments each, so that 4096 environments run as ONE round of
    // resident warps without any prefix pass: the environment index then lives in vector registers
    // instead of uniform ones and the address arithmetic of every chunk grows -- 4 % slower at equal S,
    // and S = 1, 2 lose to S = 4 even so.)
    const int S = blockDim.x >> 5;
    const int e = blockIdx.x;

    // ---- per-CTA constants: end points, log table
    CS &C = *reinterpret_cast<CS *>(smem_raw);
    const LogTabEntry *logtab = C.logtab;
    for (int idx = threadIdx.x; idx < N; idx += blockDim.x) C.cF[idx] = reinterpret_cast<const V2 *>(a.c.xF)[idx];
    if (sizeof(Real) == 8)
        for (int idx = threadIdx.x; idx < kLogTabSize; idx += blockDim.x) C.logtab[idx] = a.c.logtab[idx];
    WS &W = *reinterpret_cast<WS *>(smem_raw + sizeof(CS) + (size_t)warp * sizeof(WS));
    if (lane == 0) {
#pragma unroll
        for (int st = 0; st < kRo2Stages; ++st) ro2_mbar_init(&W.mbar[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (lane <= TCW) W.cnt[lane] = 0;
    __syncthreads();

    const int s = lane / N, i = lane - s * N;                    // slice in the chunk, agent
    const bool rowlane = lane < RW;
    const unsigned EN = (unsigned)E * N;
    const unsigned g = (unsigned)e * N + (rowlane ? i : 0);
    const Real ds = (Real)A.ds, rad = (Real)A.radius, delta_eff = (Real)A.delta_eff;
    const int act_mode = A.act_mode;
    const bool direct = ra.actions != nullptr;
    const V2 *atab = reinterpret_cast<const V2 *>(ra.atable);

    // ---- the call's time axis is cut into S segments of whole chunks; warp k owns segment k
    const int nchunks = (T + TCW - 1) / TCW;
    int c0 = 0, c1 = 0;                                          // this segment's chunks
#pragma unroll
    for (int k = 0; k < kRo2MaxSeg; ++k)                         // (selects: a run-time index into the parameter block costs a local copy)
        if (warp == k) { c0 = A.seg_c0[k]; c1 = A.seg_c0[k + 1]; }
    const int ta = c0 * TCW, Tend = (c1 * TCW < T) ? c1 * TCW : T;   // steps [ta, Tend) of the call
    const bool alive0 = ra.done[e] == 0;
    const int tenv0 = a.t[e];
    int tlim = a.max_steps - 1 - tenv0;                          // step of the call at which the time limit ends the episode
    tlim = tlim < 0 ? 0 : tlim;

    auto load_lane = [&](int c) -> V2 {                          // action of row (s, i) of chunk c
        V2 u{};
        const int t = c * TCW + s;
        if (rowlane && t < T) {
            const size_t at = (size_t)t * EN + g;
            u = direct ? reinterpret_cast<const V2 *>(ra.actions)[at] : atab[ra.aidx[at]];
        }
        return u;
    };
    // one chunk of sequential, bit-exact single-integrator steps of agent i (A = I, B = dt I,
    // drone_env.py:78-79,235) from the staged actions ua[q * N]: lane (s, i) takes the steps up to its
    // own slice, p: position after the previous chunk -> position at slice s.  The lanes of the
    // chunk's last slice end with the position the next chunk starts from.
    const Real dt = (Real)a.dt;
    auto integrate = [&](const V2 *ua, V2 &p) {
#pragma unroll
        for (int q = 0; q < TCW; ++q) {
            const V2 u = ua[q * N];
            if (q <= s) {
                p.x = add_rn(p.x, mul_rn(dt, u.x));
                p.y = add_rn(p.y, mul_rn(dt, u.y));
            }
        }
    };

    // ---- prefix: the state at the start of this warp's segment.  Only the integrator runs over the
    // steps in front of it (two dependent fp64 operations per step; actions fetched eight steps
    // ahead), with the episode-end test of the main loop: a segment behind the end does nothing.
    // (Measured, slower: a cheap sufficient test per prefix chunk -- some agent ends the chunk further
    // from its goal along one axis than goal_tol plus all it moved after the chunk's first step, so
    // nobody can have finished -- in front of the exact per-step tests: 17 instructions fewer per
    // prefix chunk on the fast path, but 0.419 ms against 0.385 on the same box with both paths in
    // the eight-fold unrolled loop, 0.429 with the exact path out of line.)
    // (Measured, no gain: refilling each prefetch register right after its chunk is consumed, so that
    // eight loads stay in flight across batch boundaries, together with requesting the environment's
    // flags and start positions above the constant copies: 0.387 ms against 0.380 - 0.386; two fold
    // iterations per loop trip (#pragma unroll 2): 0.395 ms, +6 % instructions from the main loop's
    // register allocation.)
    // (Measured alternatives.  Every lane integrating up to its own slice, as the main loop does:
    // 19 shared-memory wavefronts per chunk instead of 9.  Lane i < N reading agent i's actions
    // straight from global memory, step by step: no shared memory, but three times the loads in
    // flight for the same bytes -- the prefix stalls on them, 5 % slower.)
    // (Measured alternative: warp 0 alone integrates once and hands every segment its start state
    // through shared memory and one mbarrier per boundary -- half the prefix instructions, but the
    // other warps idle until the single-warp pass reaches their boundary: 3 % slower.)
    int tstar = 0x7fffffff;                                     // first step at which every agent is at its goal
    {
        V2 pend{};                                              // agent i's position after the previous chunk
        if (rowlane && alive0) pend = reinterpret_cast<const V2 *>(a.pos)[g];
    if (alive0 && c0 > 0 && c0 < nchunks && tlim >= ta) {
        // the row lanes fetch the chunks' actions (eight chunks in flight) and pass them through the
        // action ring; lane i < N takes agent i through the chunk's steps one after the other
        const bool al = lane < N;
        const V2 xF = C.cF[al ? lane : 0];
        auto prefix_chunk = [&](const V2 &u, int c, int par2) -> bool {
            V2 *buf = &W.act[par2][0];
            if (rowlane) buf[lane] = u;
            __syncwarp();
            const V2 *bufl = buf + (al ? lane : 0);              // (the other lanes run along on agent 0's actions; never looked at)
            unsigned hits = 0;
#pragma unroll
            for (int q = 0; q < TCW; ++q) {
                const V2 uq = bufl[q * N];
                pend.x = add_rn(pend.x, mul_rn(dt, uq.x));
                pend.y = add_rn(pend.y, mul_rn(dt, uq.y));
                // every agent within goal_tol of its goal (:249-251): sqrt_rn(x) <= tol  <=>  x <= goal_t2
                const Real gx = sub_rn(xF.x, pend.x), gy = sub_rn(xF.y, pend.y);
                const bool atg = add_rn(mul_rn(gx, gx), mul_rn(gy, gy)) <= (Real)A.goal_t2;
                hits |= __all_sync(0xffffffffu, atg || !al) ? (1u << q) : 0u;
            }
            if (hits) tstar = c * TCW + __ffs((int)hits) - 1;
            return hits != 0;
        };
        // every slice of a chunk in front of the segment lies inside the call: no bounds to test
        auto run_prefix = [&](auto fetch) {
            int c = 0;
            bool hit = false;
            for (; c + 8 <= c0 && !hit; c += 8) {                // eight chunks in flight
                V2 u8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) u8[q] = fetch();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (!hit) hit = prefix_chunk(u8[q], c + q, q & 1);
            }
            for (; c < c0 && !hit; ++c) hit = prefix_chunk(fetch(), c, c & 1);
        };
        const size_t cstep = (size_t)TCW * EN;
        if (direct) {
            const V2 *ap = reinterpret_cast<const V2 *>(ra.actions) + ((size_t)s * EN + g);
            run_prefix([&]() -> V2 { V2 u{}; if (rowlane) u = *ap; ap += cstep; return u; });
        } else {
            const uint8_t *xp = ra.aidx + ((size_t)s * EN + g);
            run_prefix([&]() -> V2 { V2 u{}; if (rowlane) u = atab[*xp]; xp += cstep; return u; });
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the ring is next written by TMA copies
        __syncwarp();
    }
        if (lane < N) W.pend[lane] = pend;
        V2 z2; z2.x = 0; z2.y = 0;
        if (rowlane) W.acc[lane] = z2;
        if (lane < TCW) W.sumc[lane] = 0;
        if constexpr (WS::TABLE) {                              // the fold's neutral element, behind the result table
            V2 nv; nv.x = (Real)A.ds; nv.y = 0;
            if (lane == 0) W.res[WS::LW] = nv;
        }
    }

    // ---- this warp's segment
    const int tfin_pre = (tstar < tlim) ? tstar : tlim;          // last executed step, as far as known in front of the segment
    const bool dead = !alive0 || T <= 0 || c0 >= nchunks || tfin_pre < ta;
    int steps = 0;
    if (lane == 0)
        W.cnt[TCW] = (int)((ra.pos_tr ? 1u : 0u) | (ra.vel_tr ? 2u : 0u) | (ra.r_tr ? 4u : 0u) | (ra.tr_tr ? 8u : 0u) |
                          (ra.z_tr ? 16u : 0u) | (ra.ncoll_tr ? 32u : 0u) | (ra.fin_tr ? 64u : 0u));   // record mask

    // ---- action staging.  TMA: lane 0 brings chunk c's [TCW][N][2] block into ring stage (c - c0) % 2
    // (one 2-D tile, or one 1-D bulk copy per slice); lane-load form: every row lane holds the action
    // of its row one chunk ahead.
    const unsigned act_s32 = ro2_smem_u32(&W.act[0][0]), mbar_s32 = ro2_smem_u32(&W.mbar[0]);
    const unsigned long long tmap_addr = reinterpret_cast<unsigned long long>(&tmap);
    auto issue_tma = [&](int c, int st) {                        // lane 0 only
        constexpr unsigned blk = (unsigned)N * (unsigned)sizeof(V2);
        if (act_mode == 2) {
            const unsigned bar = mbar_s32 + 8u * (unsigned)st, dst = act_s32 + (unsigned)(WS::ASTR * sizeof(V2)) * (unsigned)st;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)TCW * blk) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             dst), "l"(tmap_addr), "r"(e * N * 2), "r"(c * TCW), "r"(bar) : "memory");
        } else {
            const int t0c = c * TCW;
            const int nslc = (T - t0c < TCW) ? (T - t0c) : TCW;
            ro2_mbar_expect(&W.mbar[st], (unsigned)nslc * blk);
            const V2 *src = reinterpret_cast<const V2 *>(ra.actions) + ((size_t)t0c * EN + (size_t)e * N);
#pragma unroll 1
            for (int q = 0; q < nslc; ++q) ro2_bulk_g2s(&W.act[st][q * N], src + (size_t)q * EN, blk, &W.mbar[st]);
        }
    };
    V2 upre{};
    if (!dead) {
        if (act_mode) {
            if (lane == 0) issue_tma(c0, 0);
        } else {
            if (rowlane) W.act[0][lane] = load_lane(c0);
            upre = load_lane(c0 + 1);
        }
    }
    __syncwarp();

    // state of the last processed chunk, read by the epilogue after the loop
    V2 ui{}, zrow[K + 1];
    int nirow[K + 1];
    Real r_i = 0, tr_i = 0;
    int ne = 0, nc = 0;
    bool env_fin = false;
    int t0 = ta;

    int st = 0;                                                  // ring stage of the current chunk
    unsigned par = 0;                                            // mbarrier phase parity of that stage's current use
    unsigned at = (unsigned)(ta + s) * EN + g;                   // element index of this row at slice t0 + s
    for (int c = dead ? c1 : c0; c < c1; ++c, t0 += TCW, at += (unsigned)TCW * EN) {
        const bool in_chunk = rowlane && (t0 + s < T);
        const V2 *uact = &W.act[st][0];
        // ---- the next chunk's actions on their way; this chunk's have landed
        if (act_mode) {
            if (lane == 0 && c + 1 < c1) issue_tma(c + 1, st ^ 1);   // the stage last read in chunk c - 1
            ro2_mbar_wait(&W.mbar[st], par);
        }
        float fx, fy;
        {
            V2 pm = W.pend[i];
            integrate(uact + i, pm);
            __syncwarp();                                        // every lane has read the old position
            if (rowlane && s == TCW - 1) W.pend[i] = pm;
            // f32 copy for pass 1: component (i & 1) of x / y in float4 (i >> 1) of frame s
            const bool okf = fabs(pm.x) < (Real)1024 && fabs(pm.y) < (Real)1024;   // false for NaN / inf too
            fx = okf ? (float)pm.x : __int_as_float(0x7fc00000);
            fy = okf ? (float)pm.y : __int_as_float(0x7fc00000);
            if (rowlane) W.pos[lane] = pm;
        }
        if (rowlane) {
            float *pfa = reinterpret_cast<float *>(&W.posf[s * HP + (i >> 1)]) + (i & 1);
            pfa[0] = fx; pfa[2] = fy;
            if ((N & 1) && i == N - 1) { pfa[1] = 1e30f; pfa[3] = 1e30f; }      // odd N: the missing partner never is near
        }
        __syncwarp();

        // ---- pass 1 (packed f32): near mask of the row
        unsigned m = ro2_pass1<N>(&W.posf[(rowlane ? s : 0) * HP], -fx, -fy, A.thr2f) & ~(1u << i);
        m = in_chunk ? m : 0u;
        const unsigned mU = m & (0xfffffffeu << i);              // partners above i
        Real sum_all = 0, sum_loc = 0, nd[K];
        int nj[K];
#pragma unroll
        for (int q = 0; q < K; ++q) { nd[q] = ds; nj[q] = -1; }
        if constexpr (WS::TABLE) {
            // ---- list of the UNORDERED near pairs (j > i); offsets from bit-sliced ballots (no
            // dependent shuffle chain): count < 2^NB
            constexpr int NB = (N <= 2) ? 1 : (N <= 4) ? 2 : (N <= 8) ? 3 : (N <= 16) ? 4 : 5;
            const int cUl = __popc(mU);
            const unsigned ltmask = (1u << lane) - 1u;
            int baseU = 0;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const unsigned bal = __ballot_sync(0xffffffffu, (cUl >> b) & 1);
                baseU += __popc(bal & ltmask) << b;
            }
            const int totU = (int)__reduce_add_sync(0xffffffffu, (unsigned)cUl);
            {
                W.umask[lane] = m;
                typename WS::Ent *ep = W.ent + baseU;
                unsigned mm = mU;
                const unsigned ew = (unsigned)lane | ((unsigned)i << 5);
                const int iters = __reduce_max_sync(0xffffffffu, cUl);
#pragma unroll 1
                for (int it = 0; it < iters; ++it) {             // straight-line body: lanes out of pairs store to their spare slot
                    const bool on = mm != 0;
                    const unsigned j = (unsigned)(__ffs((int)mm) - 1);
                    *(on ? ep : W.ent + WS::LU + lane) = (typename WS::Ent)(ew | (j << 10));
                    ep += on ? 1 : 0;
                    mm &= mm - 1;
                }
            }
            __syncwarp();
            // ---- pass 2: one unordered near pair per lane per round; result to both rows of the table
#pragma unroll 1
            for (int q0 = 0; q0 < totU; q0 += 32) {
                const int q = q0 + lane;
                const bool valid = q < totU;
                const unsigned w = W.ent[valid ? q : 0];
                const int ri = (int)(w & 31u), ii = (int)((w >> 5) & 31u), j = (int)(w >> 10);                // j > ii
                const int rj = ri - ii + j;
                const V2 pi = W.pos[ri], pj = W.pos[rj];
                V2 dv;
                bool coll;
                ro2_eval_pair<Real>(dv.x, dv.y, coll, pi.x, pi.y, pj.x, pj.y, ds, rad, (Real)A.log_ds, (Real)A.inv_ds,
                                    a.log_mode, (Real)a.zero_eps, (Real)a.sentinel, logtab);
                if (valid) {
                    W.res[ri * WS::RS + j - 1] = dv;              // columns: partner j > ii of row ri, partner ii < j of row rj
                    W.res[rj * WS::RS + ii] = dv;
                    if (coll) atomicAdd(&W.cnt[(ri - ii) / N], 2);                // both ordered pairs collide (:284,327)
                    if (!(dv.x != ds)) {                                          // near but clipped (inside the f32 margin)
                        atomicAnd(&W.umask[ri], ~(1u << j));
                        atomicAnd(&W.umask[rj], ~(1u << ii));
                    }
                }
            }
            __syncwarp();
            const unsigned lowi = (1u << i) - 1u;
            const unsigned mcol = (m & lowi) | ((m >> 1) & ~lowi);               // near COLUMNS of the row
            ro2_fold_table<Real, K>(W.res + lane * WS::RS, W.res + WS::LW, mcol, i, delta_eff, sum_all, sum_loc, nd, nj);
#pragma unroll
            for (int q = 0; q < K; ++q) nj[q] += (nj[q] >= i) ? 1 : 0;           // column -> partner index (-1 stays)
        } else {
        const int cFl = __popc(m), cUl = __popc(mU);
        int incl = cUl | (cFl << 16);                            // both counts in one scan
#pragma unroll
        for (int w = 1; w < 32; w <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, w);
            incl += (lane >= w) ? v : 0;
        }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        const int totU = tot & 0xffff, totF = tot >> 16;
        const int baseF = (incl >> 16) - cFl;
        if (totU <= WS::LU && totF <= WS::LW) {                  // warp-uniform
            // ---- list of the UNORDERED near pairs (j > i)
            {
                W.rowinfo[lane] = make_uint2(m, (unsigned)baseF);
                W.umask[lane] = m;
                typename WS::Ent *ep = W.ent + ((incl & 0xffff) - cUl);
                unsigned mm = mU;
                // slot of (i, j) in row i's segment: its rank among the row's near partners
                unsigned ew = (unsigned)lane | ((unsigned)i << 5) |
                              (((unsigned)baseF + (unsigned)__popc(m & ((1u << i) - 1u))) << 15);
                const int iters = __reduce_max_sync(0xffffffffu, cUl);
#pragma unroll 1
                for (int it = 0; it < iters; ++it) {
                    if (mm) {
                        const int j = lowest_bit(mm);
                        mm &= mm - 1;
                        *ep++ = ew | ((unsigned)j << 10);
                        ew += 1u << 15;
                    }
                }
            }
            __syncwarp();
            // ---- pass 2: one unordered near pair per lane per round; result to both rows' segments
#pragma unroll 1
            for (int q0 = 0; q0 < totU; q0 += 32) {
                const int q = q0 + lane;
                const bool valid = q < totU;
                const unsigned w = W.ent[valid ? q : 0];
                const int ri = (int)(w & 31u), ii = (int)((w >> 5) & 31u), j = (int)((w >> 10) & 31u);
                const int rj = ri - ii + j;
                const unsigned slot_i = w >> 15;
                const uint2 inf_j = W.rowinfo[rj];
                const unsigned slot_j = inf_j.y + (unsigned)__popc(inf_j.x & ((1u << ii) - 1u));
                const V2 pi = W.pos[ri], pj = W.pos[rj];
                V2 dv;
                bool coll;
                ro2_eval_pair<Real>(dv.x, dv.y, coll, pi.x, pi.y, pj.x, pj.y, ds, rad, (Real)A.log_ds, (Real)A.inv_ds,
                                    a.log_mode, (Real)a.zero_eps, (Real)a.sentinel, logtab);
                if (valid) {
                    W.res[slot_i] = dv;
                    W.res[slot_j] = dv;
                    if (coll) atomicAdd(&W.cnt[(ri - ii) / N], 2);                // both ordered pairs collide (:284,327)
                    if (!(dv.x != ds)) {                                          // near but clipped (inside the f32 margin)
                        atomicAnd(&W.umask[ri], ~(1u << j));
                        atomicAnd(&W.umask[rj], ~(1u << ii));
                    }
                }
            }
            __syncwarp();
            ro2_fold<Real, K>(W.res + baseF, m, ds, delta_eff, sum_all, sum_loc, nd, nj);
        } else {
            // ---- dense frames: groups of GR rows, every pair of a group's rows evaluated exactly
            W.umask[lane] = fullN & ~(1u << i);
            __syncwarp();
#pragma unroll 1
            for (int gs = 0; gs < RW; gs += WS::GR) {
#pragma unroll 1
                for (int q0 = 0; q0 < WS::GR * (N - 1); q0 += 32) {
                    const int q = q0 + lane;
                    const int rr = q / (N - 1), jj = q - rr * (N - 1);
                    const int row = gs + rr, sr = row / N, ir = row - sr * N;
                    const bool valid = q < WS::GR * (N - 1) && row < RW && (t0 + sr < T);
                    const int j = jj + (jj >= ir ? 1 : 0);
                    const int rowc = valid ? row : 0, rjc = valid ? row - ir + j : 0;
                    const V2 pi = W.pos[rowc], pj = W.pos[rjc];
                    V2 dv;
                    bool coll;
                    ro2_eval_pair<Real>(dv.x, dv.y, coll, pi.x, pi.y, pj.x, pj.y, ds, rad, (Real)A.log_ds, (Real)A.inv_ds,
                                        a.log_mode, (Real)a.zero_eps, (Real)a.sentinel, logtab);
                    if (valid) {
                        W.res[q] = dv;
                        if (coll) atomicAdd(&W.cnt[sr], 1);
                        if (!(dv.x != ds)) atomicAnd(&W.umask[row], ~(1u << j));
                    }
                }
                __syncwarp();
                const bool mine = in_chunk && lane >= gs && lane < gs + WS::GR;
                ro2_fold<Real, K>(W.res + (mine ? (lane - gs) * (N - 1) : 0), mine ? (fullN & ~(1u << i)) : 0u, ds, delta_eff,
                                  sum_all, sum_loc, nd, nj);
                __syncwarp();
            }
        }
        }
        // ---- the row: k nearest, goal cost, rewards, observation (all lanes; stores predicated)
        {
            // free slots <- lowest clipped indices (all tie at exactly d_safety: index order)
            unsigned cm = ~(W.umask[lane] | (1u << i));
            if (N < 32) cm &= (1u << N) - 1u;
#pragma unroll
            for (int q = 0; q < K; ++q) {
                const bool take = nj[q] < 0 && cm != 0;
                nj[q] = take ? (__ffs((int)cm) - 1) : nj[q];
                cm = take ? (cm & (cm - 1)) : cm;
            }
            // Delta-disk count, capped at k: the in-disk partners are the nearest ones (uniform Delta)
            int inr = 0;
#pragma unroll
            for (int q = 0; q < K; ++q) inr += (nj[q] >= 0 && nd[q] <= delta_eff) ? 1 : 0;
            // the row's own entry (:323-325) precedes every partner unless one coincides with the agent
            // (d_ij == d_ii) and has a lower index: merge with the (d, j) rule
            const Real d_ii = (Real)A.d_ii;
            if (nj[0] >= 0 && nd[0] <= d_ii) {                              // coincident agents only
                int pself = 0;
#pragma unroll
                for (int q = 0; q < K; ++q)
                    pself += (nj[q] >= 0 && (nd[q] < d_ii || (nd[q] == d_ii && nj[q] < i))) ? 1 : 0;
                int mj[K];
#pragma unroll
                for (int kth = 1; kth <= K; ++kth)
                    mj[kth - 1] = (kth < pself) ? nj[kth < K ? kth : K - 1] : ((kth == pself) ? i : nj[kth - 1]);
#pragma unroll
                for (int q = 0; q < K; ++q) nj[q] = mj[q];
            }
            // goal cost, rewards, termination flag (:249-251,272-288)
            const V2 pm = W.pos[lane], xF = C.cF[i];
            const Real gx = sub_rn(xF.x, pm.x), gy = sub_rn(xF.y, pm.y);
            const Real nrm = sqrt_rn(add_rn(mul_rn(gx, gx), mul_rn(gy, gy)));     // :249,276 (axis norm, unfused)
            const Real goal = mul_rn((Real)a.q, mul_rn(nrm, nrm));                // :276
            r_i = ro2_neg_nan_to_num(add_rn(goal, mul_rn((Real)a.b, sum_loc)));   // :282,287
            tr_i = ro2_neg_nan_to_num(add_rn(goal, mul_rn((Real)a.b, sum_all)));  // :283,288
            const bool at_goal = nrm <= (Real)a.goal_tol;
            // observation (:344-397): own row, then the k nearest inside Delta or ghosts
            zrow[0].x = -gx; zrow[0].y = -gy;                                     // :357
            Real ghx = 0, ghy = 0;
            if (inr < K) {                                                        // ghost rows (:383-386)
                const Real zn = sqrt_rn(fma_rn(gy, gy, mul_rn(gx, gx)));
                ghx = mul_rn(mul_rn(div_rn(-gx, zn), (Real)A.delta), (Real)a.ghost);
                ghy = mul_rn(mul_rn(div_rn(-gy, zn), (Real)A.delta), (Real)a.ghost);
            }
            nirow[0] = i;
            const V2 *fpos = &W.pos[lane - i];
#pragma unroll
            for (int kth = 1; kth <= K; ++kth) {
                const int j = nj[kth - 1] < 0 ? 0 : nj[kth - 1];
                const bool in_r = kth <= inr;                                     // :362-368
                // (most rows of a sparse frame show ghosts: those lanes read one common address)
                const V2 pj = *((in_r && rowlane) ? fpos + j : &W.pos[0]);
                zrow[kth].x = in_r ? sub_rn(pj.x, pm.x) : ghx;
                zrow[kth].y = in_r ? sub_rn(pj.y, pm.y) : ghy;
                nirow[kth] = in_r ? j : -1;
            }
            // ---- which slices execute (drone_env.py:248-256)
            const unsigned gbal = __ballot_sync(0xffffffffu, in_chunk && at_goal);
            const bool slice_goal = in_chunk && i == 0 && (((gbal >> (lane & 31)) & fullN) == fullN);
            const unsigned sbal = __ballot_sync(0xffffffffu, slice_goal);
            const int nsl = (T - t0 < TCW) ? (T - t0) : TCW;     // slices in this chunk
            const int fg = sbal ? lowest_bit(sbal) / N : nsl;    // first slice with everybody at goal
            int ft = a.max_steps - 1 - (tenv0 + t0);             // first slice at the time limit
            ft = ft < 0 ? 0 : ft;
            const int first = fg < ft ? fg : ft;
            env_fin = first < nsl;
            ne = env_fin ? first + 1 : nsl;
        }
        // ---- stores of the executed slices
        nc = W.cnt[rowlane ? s : 0];
        if (in_chunk) {
            const unsigned fe = (unsigned)(t0 + s) * (unsigned)E + (unsigned)e;
            const unsigned recmask = (unsigned)W.cnt[TCW];
            if (s < ne) {
                ui = uact[lane];
                // (the call records everything, the usual case: one warp-uniform test instead of seven)
                auto stores = [&](auto all_tag) {
                    constexpr bool ALL = decltype(all_tag)::value;
                    if (ALL || (recmask & 1u)) reinterpret_cast<V2 *>(ra.pos_tr)[at] = W.pos[lane];
                    if (ALL || (recmask & 2u)) reinterpret_cast<V2 *>(ra.vel_tr)[at] = ui;        // :238
                    if (ALL || (recmask & 4u)) reinterpret_cast<Real *>(ra.r_tr)[at] = r_i;
                    if (ALL || (recmask & 8u)) reinterpret_cast<Real *>(ra.tr_tr)[at] = tr_i;
                    if (ALL || (recmask & 16u)) {
                        V2 *zr = reinterpret_cast<V2 *>(ra.z_tr) + (size_t)at * (K + 1);
                        int *nl = ra.Ni_tr + (size_t)at * (K + 1);
#pragma unroll
                        for (int kth = 0; kth <= K; ++kth) { zr[kth] = zrow[kth]; nl[kth] = nirow[kth]; }
                    }
                    if (i == 0) {
                        if (ALL || (recmask & 32u)) ra.ncoll_tr[fe] = nc;
                        if (ALL || (recmask & 64u)) ra.fin_tr[fe] = (env_fin && s == ne - 1) ? 1 : 0;
                    }
                };
                if (recmask == 0x7fu) stores(std::true_type{}); else stores(std::false_type{});
                {
                    V2 acc = W.acc[lane];
                    acc.x = add_rn(acc.x, r_i); acc.y = add_rn(acc.y, tr_i);
                    W.acc[lane] = acc;
                }
                if (i == 0) W.sumc[s] += nc;
            } else if (i == 0 && (recmask & 64u)) {
                ra.fin_tr[fe] = 2;
            }
        }
        __syncwarp();                                            // every lane has read its frame's count
        if (rowlane && i == 0) W.cnt[s] = 0;
        steps = t0 - ta + ne;
        // ---- lane-load form: stage the next chunk's actions, fetch the one after
        if (!act_mode && c + 1 < c1) {
            if (rowlane) W.act[st ^ 1][lane] = upre;
            upre = load_lane(c + 2);
        }
        st ^= 1;
        if (st == 0) par ^= 1u;
        __syncwarp();
        if (env_fin) {
            if (ra.fin_tr) {                                     // the steps after the episode's end are not executed
#pragma unroll 1
                for (int t = t0 + TCW + lane; t < Tend; t += 32) ra.fin_tr[(size_t)t * E + e] = 2;
            }
            break;
        }
    }
    if (dead && ra.fin_tr) {                                     // a segment behind the episode's end (or a done environment)
#pragma unroll 1
        for (int t = ta + lane; t < Tend; t += 32) ra.fin_tr[(size_t)t * E + e] = 2;
    }
    // ---- last executed step of the call: the step()-style outputs and the state in the live buffers
    // (the segment in which the episode ended, or the last segment of the call)
    if (!dead && (env_fin || Tend == T) && rowlane && s == ne - 1) {
        reinterpret_cast<Real *>(a.r)[g] = r_i;
        reinterpret_cast<Real *>(a.tr)[g] = tr_i;
        V2 *zr = reinterpret_cast<V2 *>(a.z) + (size_t)g * (K + 1);
        int *nl = a.Ni + (size_t)g * (K + 1);
#pragma unroll
        for (int kth = 0; kth <= K; ++kth) { zr[kth] = zrow[kth]; nl[kth] = nirow[kth]; }
        reinterpret_cast<V2 *>(a.pos)[g] = W.pos[lane];
        reinterpret_cast<V2 *>(a.vel)[g] = ui;
        if (i == 0) { a.ncoll[e] = nc; a.fin[e] = env_fin ? 1 : 0; }
    }
    // ---- episode sums (train_problem.py:98-100): sum over the call of mean_i r, mean_i true_r, the
    // collision counts and the steps; rows of a segment reduced over the warp, segments in time order
    double sr = rowlane ? (double)W.acc[lane].x : 0.0, stt = rowlane ? (double)W.acc[lane].y : 0.0;
    double sc = lane < TCW ? (double)W.sumc[lane] : 0.0;
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, w);
        stt += __shfl_xor_sync(0xffffffffu, stt, w);
        sc += __shfl_xor_sync(0xffffffffu, sc, w);
    }
    if (lane == 0) {
        C.part[warp][0] = sr; C.part[warp][1] = stt; C.part[warp][2] = sc; C.part[warp][3] = (double)steps;
        C.seg_fin[warp] = env_fin ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0 && alive0 && T > 0) {
        double tr_ = 0, tt_ = 0, tc_ = 0, ts_ = 0;
        int fin_any = 0;
        for (int k = 0; k < S; ++k) {
            tr_ += C.part[k][0]; tt_ += C.part[k][1]; tc_ += C.part[k][2]; ts_ += C.part[k][3];
            fin_any |= C.seg_fin[k];
        }
        a.t[e] = tenv0 + (int)ts_;
        if (ts_ > 0) {
            if (fin_any) ra.done[e] = 1;
            double *ag4 = ra.agg + (size_t)e * 4;
            ag4[0] += tr_ / N; ag4[1] += tt_ / N; ag4[2] += tc_; ag4[3] += ts_;
        }
    }
}

#endif  // __CUDACC__

}  // namespace ds
