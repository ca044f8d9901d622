// dronestep_abi.cu -- extern "C" boundary of libdronestep.so (see include/dronestep.h).
//
// Host side of the fused kernels in dronestep_kernels.cuh: argument checking,
// constant upload, launch geometry, and the host-buffer convenience entries.
// There is deliberately NO CPU implementation behind this boundary: without a
// CUDA device every compute entry fails with DS_ERR_NO_DEVICE.
#include "../../include/dronestep.h"
#include "dronestep_kernels.cuh"
#include "dronestep_rollout2.cuh"
#include "dronestep_policy.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define DS_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return fail(DS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
    } while (0)

}  // namespace

struct ds_handle {
    int E, n, k, simplify, real_bytes, device;
    int sm_count;
    // launch geometry: a CTA owns G whole environments (x TC time slices in a rollout)
    int nt;                   // thread cap of the kernel instantiation: 256 (n <= 256) or 1024
    int step_G, step_threads, step_blocks;
    size_t step_smem;
    int ro_G, ro_TC, ro_threads, ro_blocks;
    int ro_L, ro_NB;          // work-list capacity; near-mask words kept in registers (0 = any n)
    int ro_inline;            // near-pair evaluation mode of the rollout kernel (see plan_launch)
    size_t ro_smem;
    size_t smem_optin, smem_sm;
    // warp-per-environment rollout kernel (dronestep_rollout2.cuh): uniform constants, n <= 32, k in 1..3
    int ro2_ok;
    ds::Ro2Args ro2;          // everything but .ra is filled at ds_create
    int ro2_blocks, ro2_threads;
    size_t ro2_smem;
    // device constants (Real typed unless noted)
    void *d_xF, *d_ds, *d_delta, *d_radius, *d_logds, *d_thr2;
    int *d_clipcnt;
    ds::LogTabEntry *d_logtab;
    std::vector<double> h_radius;
    // staging for the host-buffer entries
    void *act_stage;
    size_t act_stage_bytes;
    // ds_rollout_host: two staging slots, two copy streams, events
    struct Slot {
        void *act = nullptr; uint8_t *aidx = nullptr;
        void *pos = nullptr, *r = nullptr, *tr = nullptr, *z = nullptr;
        int32_t *Ni = nullptr, *ncoll = nullptr; uint8_t *fin = nullptr;
        float *zf = nullptr; uint8_t *Ni8 = nullptr;      // compact copies for the host (DS_HOST_COMPACT_OBS)
        cudaEvent_t h2d_done = nullptr, kernel_done = nullptr, d2h_done = nullptr;
    } slot[2];
    int slot_chunk = 0;           // steps the slots are sized for
    unsigned slot_mask = 0;       // which trajectory buffers the slots hold
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    double *d_agg = nullptr; uint8_t *d_done = nullptr; void *d_atable = nullptr;
};

namespace {

int env_int(const char *name, int dflt);

// The warp-per-environment rollout kernel (dronestep_rollout2.cuh) applies when every agent has the
// same radius, d_safety and Delta (any scalar-delta configuration of the reference), 0 <= Delta,
// 0 < d_safety, the pass-1 threshold is finite, k = 2, the 2-column observation, and n is one of
// the instantiated agent counts.  DS_RO2=0 forces rollout_kernel (A/B runs).
#define DS_RO2_NS(X) X(4) X(5) X(8) X(10) X(16) X(20) X(32)
template <typename Real, int N> size_t ro2_warp_bytes() { return sizeof(ds::Ro2Warp<Real, N>); }
template <typename Real, int N> size_t ro2_cta_bytes() { return ds::ro2_align16(sizeof(ds::Ro2Cta<Real, N>)); }
template <typename Real>
void plan_rollout2(ds_handle *h, const Real *dsv, const Real *dl, const Real *rd, const Real *lg, const Real *thr2,
                   const int *clipcnt)
{
    const int n = h->n;
    h->ro2_ok = 0;
    if (h->k != 2 || !h->simplify || !env_int("DS_RO2", 1)) return;
    size_t wbytes = 0, cbytes = 0;
#define DS_X(NN) if (n == NN) { wbytes = ro2_warp_bytes<Real, NN>(); cbytes = ro2_cta_bytes<Real, NN>(); }
    DS_RO2_NS(DS_X)
#undef DS_X
    if (!wbytes) return;
    for (int i = 1; i < n; ++i)
        if (dsv[i] != dsv[0] || dl[i] != dl[0] || rd[i] != rd[0] || thr2[i] != thr2[0] || clipcnt[i] != clipcnt[0]) return;
    if (!(dsv[0] > (Real)0) || !(dl[0] >= (Real)0) || !(rd[0] >= (Real)0) || !std::isfinite((double)thr2[0]) ||
        !std::isfinite((double)dsv[0]) || !std::isfinite((double)dl[0]))
        return;
    if (clipcnt[0] != 0 && clipcnt[0] != n - 1) return;
    ds::Ro2Args &a = h->ro2;
    std::memset(&a, 0, sizeof a);
    a.ds = (double)dsv[0]; a.delta = (double)dl[0]; a.radius = (double)rd[0]; a.log_ds = (double)lg[0];
    a.inv_ds = (double)((Real)1 / dsv[0]);                   // correctly rounded in Real: the value div_rn(1, ds) gives
    {   // (sqrt(thr2) + 2e-3)^2 rounded up: margin >> f32 rounding of |coordinates| < 1024
        const double th = std::sqrt((double)thr2[0]) + 2e-3;
        float f = (float)(th * th);
        if (!((double)f >= th * th)) f = std::nextafterf(f, INFINITY);
        a.thr2f = f;
    }
    a.delta_eff = clipcnt[0] != 0 ? (double)INFINITY : (double)dl[0];
    {   // an agent's own entry: min(d_safety, ((0 - l) - l)) in Real arithmetic (drone_env.py:318-325)
        const Real raw_ii = ((Real)0 - rd[0]) - rd[0];
        a.d_ii = (double)((dsv[0] < raw_ii) ? dsv[0] : raw_ii);
    }
    // one CTA per environment; its warps are the time segments of a call.  Enough warps to fill the
    // SMs' resident slots several times over: short CTA lifetimes keep the partly filled last round of
    // CTAs small; more segments only add prefix work.  (Measured on B200 at 7 resident 4-warp CTAs per
    // SM: n = 10, E = 4096: 4 segments 0.383 ms, 2: 0.393, 1: 0.417; n = 32, E = 8192: 2 segments
    // 1.835 ms, 4: 1.866, 1: 1.861.)
    int segs = 1;
    while (segs < ds::kRo2MaxSeg && (long long)h->E * segs < 4LL * 24 * h->sm_count) segs *= 2;
    segs = env_int("DS_RO2_SEGS", segs);
    segs = segs < 1 ? 1 : (segs > ds::kRo2MaxSeg ? ds::kRo2MaxSeg : segs);
    h->ro2_threads = 32 * segs;
    h->ro2_blocks = h->E;
    h->ro2_smem = cbytes + (size_t)segs * wbytes;
    if (h->ro2_smem > h->smem_optin) return;
    h->ro2_ok = 1;
    if (env_int("DS_PLAN_DEBUG", 0))
        std::fprintf(stderr, "[dronestep] n=%d E=%d rollout2 plan: blocks=%d smem=%zu (%zu per warp)\n",
                     n, h->E, h->ro2_blocks, h->ro2_smem, wbytes);
}

template <typename Real>
int upload_consts(ds_handle *h, const ds_config *cfg)
{
    const int n = h->n;
    std::vector<Real> xF(2 * n), dsv(n), dl(n), rd(n), lg(n), thr2(n);
    std::vector<int> clipcnt(n);
    for (int i = 0; i < n; ++i) {
        xF[2 * i] = (Real)cfg->end_points[2 * i];
        xF[2 * i + 1] = (Real)cfg->end_points[2 * i + 1];
        dsv[i] = (Real)cfg->d_safety[i];
        dl[i] = (Real)cfg->deltas[i];
        rd[i] = (Real)cfg->radius[i];
        lg[i] = (Real)std::log(std::fabs(cfg->d_safety[i]));
    }
    // Fast-reject threshold of eval_row's pass 1.  A pair (i, j) is clipped to d_safety[i]
    // (drone_env.py:318) when  fl(fl(dist - l_i) - l_j) >= d_safety[i];  with
    // D = d_safety[i] + l_i + max_j l_j this is implied by  dist^2 >= (D (1 + m) + m)^2 (1 + m)
    // for a margin m far above the rounding error of the Real chain (pairs inside the margin
    // simply take the exact path).  No fast path when d_safety[i] == 0 (the d == 0 -> -1e-6
    // rule, :319-320, then turns clipped pairs into collisions), when D is not safely
    // positive, or for non-finite constants.
    Real rmax = 0;
    for (int i = 0; i < n; ++i) rmax = (rd[i] > rmax) ? rd[i] : rmax;
    const double m = sizeof(Real) == 8 ? 1e-12 : 1e-5;
    for (int i = 0; i < n; ++i) {
        const double D = (double)dsv[i] + (double)rd[i] + (double)rmax;
        double t2 = INFINITY;
        if (dsv[i] != (Real)0 && std::isfinite(D) && D > 1e-6) {
            const double thr = D * (1 + m) + m;
            t2 = thr * thr * (1 + m);
        }
        thr2[i] = (Real)t2;
        if (!(thr2[i] >= t2))                              // never round the threshold down
            thr2[i] = (sizeof(Real) == 8) ? (Real)std::nextafter((double)thr2[i], (double)INFINITY)
                                          : (Real)std::nextafterf((float)thr2[i], INFINITY);
        int cc = 0;
        for (int j = 0; j < n; ++j)
            if (j != i && dsv[i] <= dl[j]) ++cc;            // :328 for a clipped pair, evaluated in Real
        clipcnt[i] = cc;
    }
    plan_rollout2<Real>(h, dsv.data(), dl.data(), rd.data(), lg.data(), thr2.data(), clipcnt.data());
    DS_CUDA(cudaMalloc(&h->d_xF, sizeof(Real) * 2 * n));
    DS_CUDA(cudaMalloc(&h->d_ds, sizeof(Real) * n));
    DS_CUDA(cudaMalloc(&h->d_delta, sizeof(Real) * n));
    DS_CUDA(cudaMalloc(&h->d_radius, sizeof(Real) * n));
    DS_CUDA(cudaMalloc(&h->d_logds, sizeof(Real) * n));
    DS_CUDA(cudaMalloc(&h->d_thr2, sizeof(Real) * n));
    DS_CUDA(cudaMalloc((void **)&h->d_clipcnt, sizeof(int) * n));
    DS_CUDA(cudaMemcpy(h->d_xF, xF.data(), sizeof(Real) * 2 * n, cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(h->d_ds, dsv.data(), sizeof(Real) * n, cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(h->d_delta, dl.data(), sizeof(Real) * n, cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(h->d_radius, rd.data(), sizeof(Real) * n, cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(h->d_logds, lg.data(), sizeof(Real) * n, cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(h->d_thr2, thr2.data(), sizeof(Real) * n, cudaMemcpyHostToDevice));
    DS_CUDA(cudaMemcpy(h->d_clipcnt, clipcnt.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    std::vector<ds::LogTabEntry> tab(ds::kLogTabSize);
    ds::fill_log_table(tab.data());
    DS_CUDA(cudaMalloc((void **)&h->d_logtab, sizeof(ds::LogTabEntry) * ds::kLogTabSize));
    DS_CUDA(cudaMemcpy(h->d_logtab, tab.data(), sizeof(ds::LogTabEntry) * ds::kLogTabSize, cudaMemcpyHostToDevice));
    return DS_OK;
}

int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

size_t cta_smem(const ds_handle *h, int G, int TC)
{
    return h->real_bytes == 8 ? ds::CtaSmem<double>::bytes(h->n, G, TC) : ds::CtaSmem<float>::bytes(h->n, G, TC);
}

size_t rollout_smem(const ds_handle *h, int G, int TC, int L)
{
    return h->real_bytes == 8 ? ds::RoSmem<double>::bytes(h->n, G, TC, L) : ds::RoSmem<float>::bytes(h->n, G, TC, L);
}

// One row (agent of one environment at one time slice) per thread.  step: CTA = G whole
// environments.  rollout: CTA = G environments x TC concurrent time slices.  The rollout kernel is
// latency bound and compiled for DS_RO_MAXNREG = 96 registers: CTAs of ~160 threads (4 per SM)
// measured best on B200 -- config 3: (G,TC) = (2,8) or (1,16) 1.30e10 agent-steps/s, 250-thread
// CTAs 1.13e10; config 2: (4,8) 1.17e10; config 4: (1,4)/(1,5) 1.75e10 -- with chunks of about
// 8 slices (shorter chunks pay the per-chunk barriers more often, longer ones add nothing).
// Small batches trade environments per CTA for CTAs until the grid covers the SMs a few times.
// DS_PLAN_G / DS_PLAN_TC override the rollout plan (tuning experiments).
void plan_launch(ds_handle *h)
{
    const int n = h->n, E = h->E;
    h->nt = n <= 256 ? 256 : 1024;
    const int cap = n <= 256 ? 256 : ((n + 31) / 32) * 32;
    auto clampG = [&](int G) { G = G < 1 ? 1 : G; return G > E ? E : G; };
    h->step_G = clampG(cap / n);
    h->step_threads = ((h->step_G * n + 31) / 32) * 32;
    h->step_blocks = (int)(((long long)E + h->step_G - 1) / h->step_G);
    h->step_smem = cta_smem(h, h->step_G, 1);
    h->ro_NB = n <= 32 ? 1 : (n <= 128 ? 4 : 0);
    // near-pair evaluation: 2 = warp-local work lists (default: two barriers per chunk, no atomics;
    // measured +1-2 % over 0), 0 = one CTA-wide list, 1 = rows evaluate their own pairs (n <= 32)
    h->ro_inline = env_int("DS_PLAN_INLINE", 2);
    if (h->ro_inline == 1 && n > 32) h->ro_inline = 0;
    const int lpr = env_int("DS_PLAN_LPR", 6);
    auto list_cap = [&](int G, int TC) {                   // work list: room for lpr near pairs per row
        int per_row = (n - 1 < lpr) ? (n - 1) : lpr;
        per_row = (per_row < 1 || h->ro_inline == 1) ? 1 : per_row;
        int L = G * n * TC * per_row;
        while (per_row > 1 && rollout_smem(h, G, TC, L) > h->smem_optin) L = G * n * TC * --per_row;
        return L;
    };
    const int target = env_int("DS_PLAN_THREADS", 160) < cap ? env_int("DS_PLAN_THREADS", 160) : cap;
    const int tcmax = env_int("DS_PLAN_TCMAX", 16);
    int bestTC = target / n;
    bestTC = bestTC < 1 ? 1 : (bestTC > 8 ? 8 : bestTC);
    int bestG = clampG(target / (n * bestTC));
    while (bestG > 1 && (E + bestG - 1) / bestG < 4 * h->sm_count) bestG /= 2;
    if (bestG * n * bestTC * 2 <= target) {                // small batch: longer chunks instead
        bestTC = target / (n * bestG);
        bestTC = bestTC > tcmax ? tcmax : bestTC;
    }
    const int og = env_int("DS_PLAN_G", 0), otc = env_int("DS_PLAN_TC", 0);
    if (og > 0 && otc > 0 && otc <= 32 && og * n * otc <= cap) { bestG = clampG(og); bestTC = otc; }   // ngbits holds <= 32 slices
    h->ro_G = bestG; h->ro_TC = bestTC;
    h->ro_threads = ((bestG * n * bestTC + 31) / 32) * 32;
    h->ro_blocks = (int)(((long long)E + bestG - 1) / bestG);
    h->ro_L = list_cap(bestG, bestTC);
    h->ro_smem = rollout_smem(h, bestG, bestTC, h->ro_L);
    if (env_int("DS_PLAN_DEBUG", 0))
        std::fprintf(stderr, "[dronestep] n=%d E=%d rollout plan: G=%d TC=%d threads=%d blocks=%d L=%d smem=%zu inline=%d\n",
                     n, E, h->ro_G, h->ro_TC, h->ro_threads, h->ro_blocks, h->ro_L, h->ro_smem, h->ro_inline);
}

int check_params(const ds_params *p)
{
    if (!p) return fail(DS_ERR_ARG, "ds_params is NULL");
    if (p->log_mode != DS_LOG_DIV && p->log_mode != DS_LOG_DIFF && p->log_mode != DS_LOG_RCP)
        return fail(DS_ERR_ARG, "ds_params.log_mode must be DS_LOG_DIV, DS_LOG_DIFF or DS_LOG_RCP");
    if (p->max_time_steps < 1) return fail(DS_ERR_ARG, "ds_params.max_time_steps < 1");
    return DS_OK;
}

int fill_step_args(ds_handle *h, const ds_params *p, const ds_buffers *io, const void *act,
                   bool integrate, ds::StepArgs *a)
{
    if (!h) return fail(DS_ERR_ARG, "handle is NULL");
    if (int rc = check_params(p)) return rc;
    if (!io) return fail(DS_ERR_ARG, "ds_buffers is NULL");
    if (!io->pos || !io->vel || !io->reward || !io->true_reward || !io->z || !io->Ni || !io->ncoll)
        return fail(DS_ERR_ARG, "ds_buffers: pos/vel/reward/true_reward/z/Ni/ncoll must be non-NULL");
    if (integrate && (!io->finished || !io->t))
        return fail(DS_ERR_ARG, "ds_buffers: finished/t must be non-NULL for a step");
    a->E = h->E; a->n = h->n; a->k = h->k; a->simplify = h->simplify; a->G = h->step_G;
    a->do_integrate = integrate ? 1 : 0;
    a->ctrl = 0; a->u_max = 1.0; a->ctrl_out = nullptr;
    a->log_mode = p->log_mode;
    a->max_steps = p->max_time_steps;
    a->c = ds::Consts{h->d_xF, h->d_ds, h->d_delta, h->d_radius, h->d_logds, h->d_thr2, h->d_clipcnt, h->d_logtab};
    a->dt = p->dt;
    a->q = 2 * p->dt;                        // drone_env.py:269
    a->b = p->collision_weight * p->dt;      // drone_env.py:270
    a->goal_tol = p->goal_tol; a->sentinel = p->sentinel; a->zero_eps = p->zero_eps;
    a->ghost = p->ghost_factor;
    a->act = act;
    a->pos = io->pos; a->vel = io->vel; a->r = io->reward; a->tr = io->true_reward; a->z = io->z;
    a->Ni = io->Ni; a->ncoll = io->ncoll; a->fin = io->finished; a->t = io->t;
    return DS_OK;
}

struct Geom { int blocks, threads; size_t smem; };

// Function attributes are per (device, kernel): set them when they change, not on every launch
// (two driver calls per launch were most of the host time of the one-environment drop-in step).
int ensure_attrs(const void *kernel, size_t smem, int carveout)
{
    struct Cfg { size_t smem; int carve; };
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, Cfg> done;
    int dev = 0;
    DS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(dev, kernel);
    auto it = done.find(key);
    if (it != done.end() && it->second.smem >= smem && it->second.carve == carveout) return DS_OK;
    const size_t want = (it != done.end() && it->second.smem > smem) ? it->second.smem : smem;
    if (want > 48 * 1024)
        DS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    // residency is shared-memory bound for small CTAs: ask for the largest carve-out (default)
    DS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
    done[key] = Cfg{want, carveout};
    return DS_OK;
}

template <typename KernelT, typename ArgsT>
int launch(KernelT kernel, const ArgsT &args, const Geom &gm, cudaStream_t st, int carveout = cudaSharedmemCarveoutMaxShared)
{
    if (int rc = ensure_attrs(reinterpret_cast<const void *>(kernel), gm.smem, carveout)) return rc;
    kernel<<<gm.blocks, gm.threads, gm.smem, st>>>(args);
    DS_CUDA(cudaGetLastError());
    return DS_OK;
}

#define DS_DISPATCH_K(KERNEL, REAL, NT, ARGS, GEOM)                                       \
    switch (h->k) {                                                                      \
    case 0: return launch(ds::KERNEL<REAL, 0, NT>, ARGS, GEOM, st);                      \
    case 1: return launch(ds::KERNEL<REAL, 1, NT>, ARGS, GEOM, st);                      \
    case 2: return launch(ds::KERNEL<REAL, 2, NT>, ARGS, GEOM, st);                      \
    case 3: return launch(ds::KERNEL<REAL, 3, NT>, ARGS, GEOM, st);                      \
    case 4: return launch(ds::KERNEL<REAL, 4, NT>, ARGS, GEOM, st);                      \
    default: return launch(ds::KERNEL<REAL, -1, NT>, ARGS, GEOM, st);                    \
    }

#define DS_DISPATCH_NT(KERNEL, REAL, ARGS, GEOM)                                          \
    if (h->nt == 256) { DS_DISPATCH_K(KERNEL, REAL, 256, ARGS, GEOM) }                    \
    else { DS_DISPATCH_K(KERNEL, REAL, 1024, ARGS, GEOM) }


int launch_step(ds_handle *h, const ds::StepArgs &a, cudaStream_t st)
{
    const Geom gm{h->step_blocks, h->step_threads, h->step_smem};
    if (h->real_bytes == 8) { DS_DISPATCH_NT(step_kernel, double, a, gm) }
    else { DS_DISPATCH_NT(step_kernel, float, a, gm) }
}

#ifdef DS_FAST_BUILD   /* tuning builds: k = 2 and the generic k only */
#define DS_DISPATCH_RO_K(REAL, NT, NB, ARGS, GEOM)                                        \
    switch (h->k) {                                                                      \
    case 2: return launch(ds::rollout_kernel<REAL, 2, NT, NB>, ARGS, GEOM, st);          \
    default: return launch(ds::rollout_kernel<REAL, -1, NT, NB>, ARGS, GEOM, st);        \
    }
#else
#define DS_DISPATCH_RO_K(REAL, NT, NB, ARGS, GEOM)                                        \
    switch (h->k) {                                                                      \
    case 0: return launch(ds::rollout_kernel<REAL, 0, NT, NB>, ARGS, GEOM, st);          \
    case 1: return launch(ds::rollout_kernel<REAL, 1, NT, NB>, ARGS, GEOM, st);          \
    case 2: return launch(ds::rollout_kernel<REAL, 2, NT, NB>, ARGS, GEOM, st);          \
    case 3: return launch(ds::rollout_kernel<REAL, 3, NT, NB>, ARGS, GEOM, st);          \
    case 4: return launch(ds::rollout_kernel<REAL, 4, NT, NB>, ARGS, GEOM, st);          \
    default: return launch(ds::rollout_kernel<REAL, -1, NT, NB>, ARGS, GEOM, st);        \
    }
#endif

#define DS_DISPATCH_RO(REAL, ARGS, GEOM)                                                  \
    if (h->nt == 1024) { DS_DISPATCH_RO_K(REAL, 1024, 0, ARGS, GEOM) }                    \
    else if (h->ro_NB == 1) { DS_DISPATCH_RO_K(REAL, 256, 1, ARGS, GEOM) }                \
    else if (h->ro_NB == 4) { DS_DISPATCH_RO_K(REAL, 256, 4, ARGS, GEOM) }                \
    else { DS_DISPATCH_RO_K(REAL, 256, 0, ARGS, GEOM) }

// cuTensorMapEncodeTiled through the runtime's driver entry point query (no link-time libcuda dependency)
typedef CUresult (*ds_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
ds_encode_tiled_fn encode_tiled()
{
    static ds_encode_tiled_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return (ds_encode_tiled_fn)p;
    }();
    return fn;
}

// largest x with sqrt(x) <= tol in Real arithmetic (sqrt is correctly rounded and monotone): the
// all-at-goal test of drone_env.py:249-251 without the square root, for the prefix passes
template <typename Real>
double goal_threshold_sq(double tol_d)
{
    const Real tol = (Real)tol_d;
    if (!(tol >= (Real)0) || !std::isfinite((double)tol)) return -1.0;          // nothing is ever at its goal / not usable
    Real x = tol * tol;
    while (std::sqrt(x) > tol) x = std::nextafter(x, (Real)0);
    for (;;) {
        const Real nx = std::nextafter(x, (Real)INFINITY);
        if (!(std::sqrt(nx) <= tol)) break;
        x = nx;
    }
    return (double)x;
}

template <typename Real, int N>
int launch_rollout2_n(ds_handle *h, const ds::Ro2Args &a, const CUtensorMap &tm, cudaStream_t st)
{
    auto kernel = ds::rollout2_kernel<Real, N, 2>;
    if (int rc = ensure_attrs(reinterpret_cast<const void *>(kernel), h->ro2_smem, env_int("DS_RO2_CARVE", 100))) return rc;
    if (env_int("DS_PLAN_DEBUG", 0)) {
        int nb = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, h->ro2_threads, h->ro2_smem);
        std::fprintf(stderr, "[dronestep] rollout2: %d CTAs of %d threads, %zu B smem, %d resident CTAs per SM, act_mode %d\n",
                     h->ro2_blocks, h->ro2_threads, h->ro2_smem, nb, a.act_mode);
    }
    kernel<<<h->ro2_blocks, h->ro2_threads, h->ro2_smem, st>>>(a, tm);
    DS_CUDA(cudaGetLastError());
    return DS_OK;
}

int launch_rollout2(ds_handle *h, const ds::RolloutArgs &ra, cudaStream_t st)
{
    ds::Ro2Args a = h->ro2;
    a.ra = ra;
    a.goal_t2 = h->real_bytes == 8 ? goal_threshold_sq<double>(ra.s.goal_tol) : goal_threshold_sq<float>(ra.s.goal_tol);
    {   // time segments: segment k pays rho chunk-equivalents per chunk in front of it (its prefix pass), so
        // its own length shrinks accordingly: L_k = L_0 - rho * (L_0 + ... + L_{k-1}); a CTA's warps finish together
        const int S = h->ro2_threads / 32, TCw = 32 / h->n;
        const int nchunks = (ra.T + TCw - 1) / TCw;
        const double rho = env_int("DS_RO2_RHO_PERMILLE", 55) / 1000.0;
        double w[ds::kRo2MaxSeg], acc = 0, tot = 0;
        for (int k = 0; k < S; ++k) { w[k] = 1.0 - rho * acc; w[k] = w[k] < 0.1 ? 0.1 : w[k]; acc += w[k]; tot += w[k]; }
        double run = 0;
        a.seg_c0[0] = 0;
        for (int k = 0; k < S; ++k) {
            run += w[k];
            int c = (int)std::lround(nchunks * run / tot);
            c = c < a.seg_c0[k] ? a.seg_c0[k] : (c > nchunks ? nchunks : c);
            a.seg_c0[k + 1] = (k == S - 1) ? nchunks : c;
        }
        for (int k = S + 1; k <= ds::kRo2MaxSeg; ++k) a.seg_c0[k] = nchunks;
    }
    // TMA: 16-byte aligned source, slice blocks a multiple of 16 bytes.  2: one 2-D tile per chunk
    // (tensor map over [T][E * n * 2] Reals); 1: one 1-D bulk copy per slice; 0: per-lane loads.
    const size_t rb = (size_t)h->real_bytes, blk = (size_t)h->n * 2 * rb;
    const int want = env_int("DS_RO2_ACT", 2);
    a.act_mode = 0;
    CUtensorMap tm;
    std::memset(&tm, 0, sizeof tm);
    if (ra.actions && blk % 16 == 0 && ((uintptr_t)ra.actions % 16) == 0 && want > 0) {
        a.act_mode = 1;
        ds_encode_tiled_fn enc = want >= 2 ? encode_tiled() : nullptr;
        if (enc) {
            const cuuint64_t gdim[2] = {(cuuint64_t)h->E * h->n * 2, (cuuint64_t)ra.T};
            const cuuint64_t gstr[1] = {(cuuint64_t)h->E * h->n * 2 * rb};
            const cuuint32_t box[2] = {(cuuint32_t)h->n * 2, (cuuint32_t)(32 / h->n)};
            const cuuint32_t estr[2] = {1, 1};
            const CUresult rc = enc(&tm, rb == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                                    const_cast<void *>(ra.actions), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc == CUDA_SUCCESS) a.act_mode = 2;
        }
    }
#define DS_X(NN)                                                                                   \
    if (h->n == NN)                                                                                \
        return h->real_bytes == 8 ? launch_rollout2_n<double, NN>(h, a, tm, st) : launch_rollout2_n<float, NN>(h, a, tm, st);
    DS_RO2_NS(DS_X)
#undef DS_X
    return fail(DS_ERR_INTERNAL, "rollout2: no instantiation for this n");
}

int launch_rollout(ds_handle *h, const ds::RolloutArgs &a, cudaStream_t st)
{
    if (h->ro2_ok) return launch_rollout2(h, a, st);
    const Geom gm{h->ro_blocks, h->ro_threads, h->ro_smem};
    if (h->real_bytes == 8) { DS_DISPATCH_RO(double, a, gm) }
    else { DS_DISPATCH_RO(float, a, gm) }
}

void free_slots(ds_handle *h)
{
    for (auto &s : h->slot) {
        cudaFree(s.act); cudaFree(s.aidx); cudaFree(s.pos); cudaFree(s.r);
        cudaFree(s.tr); cudaFree(s.z); cudaFree(s.Ni); cudaFree(s.ncoll); cudaFree(s.fin);
        cudaFree(s.zf); cudaFree(s.Ni8);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.kernel_done) cudaEventDestroy(s.kernel_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
        s = ds_handle::Slot();
    }
    h->slot_chunk = 0; h->slot_mask = 0;
}

// Host-facing compaction of a chunk's observations (ds_rollout_host, DS_HOST_COMPACT_OBS): z as float32
// -- what the reference's actors cast it to (utils.py:305) -- and the neighbour lists as u8 (255 = none).
template <typename Real>
__global__ void compact_obs_kernel(const Real *__restrict__ z, const int *__restrict__ Ni, float *__restrict__ zf,
                                   uint8_t *__restrict__ Ni8, size_t nz, size_t nn)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nz; q += stride) zf[q] = (float)z[q];
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nn; q += stride) {
        const int v = Ni[q];
        Ni8[q] = v < 0 ? (uint8_t)255 : (uint8_t)v;
    }
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int ds_abi_version(void) { return DS_ABI_VERSION; }

const char *ds_last_error(void) { return g_err.c_str(); }

int ds_device_count(void)
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

void ds_default_params(ds_params *p)
{
    if (!p) return;
    p->dt = 0.05;
    p->collision_weight = 0.2;
    p->goal_tol = 0.2;
    p->sentinel = 9.99E3;
    p->zero_eps = -1e-6;
    p->ghost_factor = 1.1;
    p->max_time_steps = 200;
    p->log_mode = DS_LOG_DIV;
}

int ds_create(const ds_config *cfg, ds_handle **out)
try {
    if (!cfg || !out) return fail(DS_ERR_ARG, "ds_create: NULL argument");
    *out = nullptr;
    if (cfg->n_envs < 1) return fail(DS_ERR_ARG, "ds_create: n_envs < 1");
    if (cfg->n_agents < 1 || cfg->n_agents > DS_MAX_AGENTS)
        return fail(DS_ERR_ARG, "ds_create: n_agents out of range [1, DS_MAX_AGENTS]");
    if (cfg->k_closest < 0 || cfg->k_closest > DS_MAX_K || cfg->k_closest >= cfg->n_agents)
        return fail(DS_ERR_ARG, "ds_create: k_closest must satisfy 0 <= k <= min(n_agents-1, DS_MAX_K)");
    if (cfg->real_bytes != 4 && cfg->real_bytes != 8)
        return fail(DS_ERR_ARG, "ds_create: real_bytes must be 4 or 8");
    if (!cfg->end_points || !cfg->d_safety || !cfg->deltas || !cfg->radius)
        return fail(DS_ERR_ARG, "ds_create: constant arrays must be non-NULL");
    const int ndev = ds_device_count();
    if (ndev < 1)
        return fail(DS_ERR_NO_DEVICE, "ds_create: no CUDA device visible; libdronestep has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(DS_ERR_ARG, "ds_create: bad device ordinal");
    DeviceGuard guard(cfg->device);
    if (!guard.ok) return fail(DS_ERR_CUDA, "ds_create: cudaSetDevice failed");

    ds_handle *h = new ds_handle();
    h->E = cfg->n_envs; h->n = cfg->n_agents; h->k = cfg->k_closest;
    h->simplify = cfg->simplify_zstate ? 1 : 0;
    h->real_bytes = cfg->real_bytes; h->device = cfg->device;
    h->d_xF = h->d_ds = h->d_delta = h->d_radius = h->d_logds = h->d_thr2 = nullptr;
    h->d_clipcnt = nullptr; h->d_logtab = nullptr;
    h->act_stage = nullptr; h->act_stage_bytes = 0;
    h->h_radius.assign(cfg->radius, cfg->radius + cfg->n_agents);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) {
        delete h;
        return fail(DS_ERR_CUDA, "ds_create: cudaGetDeviceProperties failed");
    }
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->smem_sm = prop.sharedMemPerMultiprocessor;
    plan_launch(h);
    if (h->step_smem > (size_t)prop.sharedMemPerBlockOptin || h->ro_smem > (size_t)prop.sharedMemPerBlockOptin) {
        delete h;
        return fail(DS_ERR_ARG, "ds_create: shared-memory footprint exceeds the device limit");
    }
    const int rc = (h->real_bytes == 8) ? upload_consts<double>(h, cfg) : upload_consts<float>(h, cfg);
    if (rc != DS_OK) { ds_destroy(h); return rc; }
    *out = h;
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_create: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_create: unknown host exception");
}

const char *ds_rollout_kernel_name(const ds_handle *h)
{
    return (h && h->ro2_ok) ? "ds::rollout2_kernel" : "ds::rollout_kernel";
}

void ds_destroy(ds_handle *h)
{
    if (!h) return;
    DeviceGuard guard(h->device);
    cudaFree(h->d_xF); cudaFree(h->d_ds); cudaFree(h->d_delta); cudaFree(h->d_radius);
    cudaFree(h->d_logds); cudaFree(h->d_thr2); cudaFree(h->d_clipcnt); cudaFree(h->d_logtab);
    cudaFree(h->act_stage);
    free_slots(h);
    cudaFree(h->d_agg); cudaFree(h->d_done); cudaFree(h->d_atable);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    delete h;
}

int ds_step(ds_handle *h, const void *actions_dev, const ds_params *p, const ds_buffers *io,
            void *cuda_stream)
try {
    ds::StepArgs a;
    if (int rc = fill_step_args(h, p, io, actions_dev, true, &a)) return rc;
    if (!actions_dev) return fail(DS_ERR_ARG, "ds_step: actions_dev is NULL");
    DeviceGuard guard(h->device);
    return launch_step(h, a, (cudaStream_t)cuda_stream);
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_step: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_step: unknown host exception");
}

int ds_step_control(ds_handle *h, int controller, double u_max, const ds_params *p, const ds_buffers *io,
                    void *cuda_stream)
try {
    ds::StepArgs a;
    if (int rc = fill_step_args(h, p, io, nullptr, true, &a)) return rc;
    if (controller != DS_CTRL_PROPORTIONAL && controller != DS_CTRL_GRADIENT)
        return fail(DS_ERR_ARG, "ds_step_control: controller must be DS_CTRL_PROPORTIONAL or DS_CTRL_GRADIENT");
    if (!(u_max >= 0)) return fail(DS_ERR_ARG, "ds_step_control: u_max must be >= 0");
    a.ctrl = controller; a.u_max = u_max;
    DeviceGuard guard(h->device);
    return launch_step(h, a, (cudaStream_t)cuda_stream);
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_step_control: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_step_control: unknown host exception");
}

int ds_control(ds_handle *h, int controller, double u_max, const ds_buffers *io, void *actions_out_dev,
               void *cuda_stream)
try {
    ds_params p;
    ds_default_params(&p);
    ds::StepArgs a;
    if (int rc = fill_step_args(h, &p, io, nullptr, false, &a)) return rc;
    if (controller != DS_CTRL_PROPORTIONAL && controller != DS_CTRL_GRADIENT)
        return fail(DS_ERR_ARG, "ds_control: controller must be DS_CTRL_PROPORTIONAL or DS_CTRL_GRADIENT");
    if (!(u_max >= 0)) return fail(DS_ERR_ARG, "ds_control: u_max must be >= 0");
    if (!actions_out_dev) return fail(DS_ERR_ARG, "ds_control: actions_out_dev is NULL");
    a.ctrl = controller; a.u_max = u_max; a.ctrl_out = actions_out_dev;
    DeviceGuard guard(h->device);
    return launch_step(h, a, (cudaStream_t)cuda_stream);
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_control: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_control: unknown host exception");
}

static int launch_rollout_control(ds_handle *h, const ds::RolloutArgs &ra, const Geom &gm, cudaStream_t st)
{
    if (h->real_bytes == 8) { DS_DISPATCH_NT(rollout_control_kernel, double, ra, gm) }
    else { DS_DISPATCH_NT(rollout_control_kernel, float, ra, gm) }
}

int ds_rollout_control(ds_handle *h, int controller, double u_max, const ds_params *p, const ds_buffers *io,
                       const ds_rollout_io *ro, void *cuda_stream)
try {
    ds::RolloutArgs ra;
    std::memset(&ra, 0, sizeof ra);
    if (int rc = fill_step_args(h, p, io, nullptr, true, &ra.s)) return rc;
    if (controller != DS_CTRL_PROPORTIONAL && controller != DS_CTRL_GRADIENT)
        return fail(DS_ERR_ARG, "ds_rollout_control: controller must be DS_CTRL_PROPORTIONAL or DS_CTRL_GRADIENT");
    if (!(u_max >= 0)) return fail(DS_ERR_ARG, "ds_rollout_control: u_max must be >= 0");
    if (!ro) return fail(DS_ERR_ARG, "ds_rollout_control: ds_rollout_io is NULL");
    if (ro->T < 0) return fail(DS_ERR_ARG, "ds_rollout_control: T < 0");
    if (!ro->agg || !ro->done) return fail(DS_ERR_ARG, "ds_rollout_control: agg/done must be non-NULL");
    if ((ro->z_tr == nullptr) != (ro->Ni_tr == nullptr))
        return fail(DS_ERR_ARG, "ds_rollout_control: z_tr and Ni_tr must be given together");
    ra.s.ctrl = controller; ra.s.u_max = u_max;
    ra.T = ro->T;
    ra.pos_tr = ro->pos_tr; ra.vel_tr = ro->vel_tr; ra.r_tr = ro->reward_tr; ra.tr_tr = ro->true_reward_tr;
    ra.z_tr = ro->z_tr; ra.Ni_tr = ro->Ni_tr; ra.ncoll_tr = ro->ncoll_tr; ra.fin_tr = ro->finished_tr;
    ra.agg = ro->agg; ra.done = ro->done;
    if (ro->T == 0) return DS_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const Geom gm{h->step_blocks, h->step_threads, h->step_smem};
    // One launch for the whole closed-loop episode (default).  DS_CTRL_FUSED=0 issues T launches of the
    // same kernel's one-step form instead -- same semantics, kept for comparison: measured on B200 at
    // the config-3 batch (profiles/r02/final_bench_closed_loop.json) the single launch executes
    // 2.4e9 agent-steps/s, T one-step launches 1.4e9 (most environments finish after ~90 steps, the
    // remaining launches find nothing to do).  Round 1's "3.4e9 with one launch per step" counted
    // ds_step_control launches, which keep stepping environments whose episode has ended.
    if (env_int("DS_CTRL_FUSED", 1) || ro->T == 1) return launch_rollout_control(h, ra, gm, st);
    const size_t E = (size_t)h->E, EN = E * h->n, rb = (size_t)h->real_bytes;
    const size_t zc = (size_t)(h->k + 1) * (h->simplify ? 2 : 5);
    auto off = [](void *base, size_t bytes) -> void * { return base ? (void *)((char *)base + bytes) : nullptr; };
    ra.T = 1;
    for (int t = 0; t < ro->T; ++t) {
        ra.pos_tr = off(ro->pos_tr, (size_t)t * EN * 2 * rb);
        ra.vel_tr = off(ro->vel_tr, (size_t)t * EN * 2 * rb);
        ra.r_tr = off(ro->reward_tr, (size_t)t * EN * rb);
        ra.tr_tr = off(ro->true_reward_tr, (size_t)t * EN * rb);
        ra.z_tr = off(ro->z_tr, (size_t)t * EN * zc * rb);
        ra.Ni_tr = (int *)off(ro->Ni_tr, (size_t)t * EN * (h->k + 1) * sizeof(int32_t));
        ra.ncoll_tr = (int *)off(ro->ncoll_tr, (size_t)t * E * sizeof(int32_t));
        ra.fin_tr = (uint8_t *)off(ro->finished_tr, (size_t)t * E);
        if (int rc = launch_rollout_control(h, ra, gm, st)) return rc;
    }
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_rollout_control: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_rollout_control: unknown host exception");
}

int ds_observe(ds_handle *h, const ds_params *p, const ds_buffers *io, void *cuda_stream)
try {
    ds::StepArgs a;
    if (int rc = fill_step_args(h, p, io, nullptr, false, &a)) return rc;
    DeviceGuard guard(h->device);
    return launch_step(h, a, (cudaStream_t)cuda_stream);
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_observe: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_observe: unknown host exception");
}

int ds_rollout(ds_handle *h, const ds_params *p, const ds_buffers *io, const ds_rollout_io *ro,
               void *cuda_stream)
try {
    ds::RolloutArgs ra;
    if (int rc = fill_step_args(h, p, io, nullptr, true, &ra.s)) return rc;
    if (!ro) return fail(DS_ERR_ARG, "ds_rollout: ds_rollout_io is NULL");
    if (ro->T < 0) return fail(DS_ERR_ARG, "ds_rollout: T < 0");
    if (!ro->agg || !ro->done) return fail(DS_ERR_ARG, "ds_rollout: agg/done must be non-NULL");
    if (!ro->actions) {
        if (!ro->action_idx || !ro->action_table || ro->n_actions < 1 || ro->n_actions > 256)
            return fail(DS_ERR_ARG, "ds_rollout: give actions, or action_idx + action_table (1..256 rows)");
    }
    if ((ro->z_tr == nullptr) != (ro->Ni_tr == nullptr))
        return fail(DS_ERR_ARG, "ds_rollout: z_tr and Ni_tr must be given together");
    ra.s.G = h->ro_G;
    ra.TC = h->ro_TC; ra.n_actions = ro->n_actions; ra.L = h->ro_L;
    ra.inline_rows = h->ro_inline;
    auto magic = [](unsigned d) -> unsigned { return d <= 1 ? 0u : (unsigned)((((unsigned long long)1 << 32) + d - 1) / d); };
    ra.mulA = magic((unsigned)(h->ro_G * h->n)); ra.mulN = magic((unsigned)h->n);
    ra.atable = ro->action_table;
    ra.agg = ro->agg; ra.done = ro->done;
    if (ro->T == 0) return DS_OK;
    DeviceGuard guard(h->device);
    // the kernel indexes trajectory elements with 32 bits: split calls whose T * E * n does not fit
    const size_t EN = (size_t)h->E * h->n, E = (size_t)h->E, rb = (size_t)h->real_bytes;
    const size_t zc = (size_t)(h->k + 1) * (h->simplify ? 2 : 5);
    // (signed 64-bit: E * n * (TC + 1) may exceed 2^32, which a size_t subtraction would wrap)
    const long long slack = (long long)((((size_t)1 << 32) - 1) / EN) - (long long)std::max(h->ro_TC, 32);
    if (slack < 1)
        return fail(DS_ERR_ARG, "ds_rollout: n_envs * n_agents too large for the kernel's 32-bit trajectory indices; "
                                "split the batch over several handles");
    const int max_T = (int)std::min<long long>(slack, 1 << 30);
    auto off = [](const void *p, size_t bytes) -> void * { return p ? (char *)p + bytes : nullptr; };
    for (int t0 = 0; t0 < ro->T; t0 += max_T) {
        const size_t ts = (size_t)t0;
        ra.T = (ro->T - t0 < max_T) ? ro->T - t0 : max_T;
        ra.actions = off(ro->actions, ts * EN * 2 * rb);
        ra.aidx = (const uint8_t *)off(ro->action_idx, ts * EN);
        ra.pos_tr = off(ro->pos_tr, ts * EN * 2 * rb); ra.vel_tr = off(ro->vel_tr, ts * EN * 2 * rb);
        ra.r_tr = off(ro->reward_tr, ts * EN * rb); ra.tr_tr = off(ro->true_reward_tr, ts * EN * rb);
        ra.z_tr = off(ro->z_tr, ts * EN * zc * rb);
        ra.Ni_tr = (int *)off(ro->Ni_tr, ts * EN * (h->k + 1) * sizeof(int32_t));
        ra.ncoll_tr = (int *)off(ro->ncoll_tr, ts * E * sizeof(int32_t));
        ra.fin_tr = (uint8_t *)off(ro->finished_tr, ts * E);
        if (int rc = launch_rollout(h, ra, (cudaStream_t)cuda_stream)) return rc;
    }
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_rollout: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_rollout: unknown host exception");
}

int ds_reduce_aggregates(ds_handle *h, const double *agg_dev, double *out_dev, void *cuda_stream)
try {
    if (!h || !agg_dev || !out_dev) return fail(DS_ERR_ARG, "ds_reduce_aggregates: NULL argument");
    DeviceGuard guard(h->device);
    ds::reduce_agg_kernel<<<1, 1024, 0, (cudaStream_t)cuda_stream>>>(agg_dev, h->E, out_dev);
    DS_CUDA(cudaGetLastError());
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_reduce_aggregates: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_reduce_aggregates: unknown host exception");
}

int ds_returns(ds_handle *h, const ds_returns_io *io, void *cuda_stream)
try {
    if (!h || !io) return fail(DS_ERR_ARG, "ds_returns: NULL argument");
    if (io->T < 0) return fail(DS_ERR_ARG, "ds_returns: T < 0");
    if (!io->reward_tr || !io->Ni_tr || !io->finished_tr || !io->returns || !io->advantage)
        return fail(DS_ERR_ARG, "ds_returns: reward_tr/Ni_tr/finished_tr/returns/advantage must be non-NULL");
    if (h->n > 256) return fail(DS_ERR_ARG, "ds_returns: n_agents > 256 is not supported");
    if (io->T == 0) return DS_OK;
    DeviceGuard guard(h->device);
    ds::ReturnsArgs a;
    a.E = h->E; a.n = h->n; a.k = h->k; a.T = io->T;
    // CTAs of <= 128 threads (whole environments): enough CTAs to cover the SMs several times
    int G = 128 / h->n;
    G = G < 1 ? 1 : (G > h->E ? h->E : G);
    while (G > 1 && (h->E + G - 1) / G < 4 * h->sm_count) G /= 2;
    a.G = G;
    a.discount = io->discount;
    a.r_tr = io->reward_tr; a.base = io->baseline; a.Ni_tr = io->Ni_tr; a.fin_tr = io->finished_tr;
    a.ret = io->returns; a.adv = io->advantage; a.cnt = io->count;
    const int threads = ((G * h->n + 31) / 32) * 32;
    const int blocks = (h->E + G - 1) / G;
    const size_t smem = 2 * (size_t)G * h->n * h->real_bytes;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (h->real_bytes == 8) {
        if (h->k == 2) ds::returns_kernel<double, 3><<<blocks, threads, smem, st>>>(a);
        else ds::returns_kernel<double, 0><<<blocks, threads, smem, st>>>(a);
    } else {
        if (h->k == 2) ds::returns_kernel<float, 3><<<blocks, threads, smem, st>>>(a);
        else ds::returns_kernel<float, 0><<<blocks, threads, smem, st>>>(a);
    }
    DS_CUDA(cudaGetLastError());
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_returns: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_returns: unknown host exception");
}

int ds_set_state(ds_handle *h, const double *state_host, const int32_t *t_host, const ds_buffers *io,
                 void *cuda_stream)
try {
    if (!h || !state_host || !io || !io->pos || !io->vel)
        return fail(DS_ERR_ARG, "ds_set_state: NULL argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t A = (size_t)h->E * h->n;
    const size_t bytes = A * 2 * h->real_bytes;
    std::vector<unsigned char> pos(bytes), vel(bytes);
    for (size_t a = 0; a < A; ++a) {
        const double *row = state_host + a * 5;
        if (h->real_bytes == 8) {
            double *pp = (double *)pos.data() + 2 * a, *vv = (double *)vel.data() + 2 * a;
            pp[0] = row[0]; pp[1] = row[1]; vv[0] = row[2]; vv[1] = row[3];
        } else {
            float *pp = (float *)pos.data() + 2 * a, *vv = (float *)vel.data() + 2 * a;
            pp[0] = (float)row[0]; pp[1] = (float)row[1]; vv[0] = (float)row[2]; vv[1] = (float)row[3];
        }
    }
    DS_CUDA(cudaMemcpyAsync(io->pos, pos.data(), bytes, cudaMemcpyHostToDevice, st));
    DS_CUDA(cudaMemcpyAsync(io->vel, vel.data(), bytes, cudaMemcpyHostToDevice, st));
    if (t_host && io->t)
        DS_CUDA(cudaMemcpyAsync(io->t, t_host, sizeof(int32_t) * h->E, cudaMemcpyHostToDevice, st));
    DS_CUDA(cudaStreamSynchronize(st));
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_set_state: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_set_state: unknown host exception");
}

int ds_get_state(ds_handle *h, double *state_host, int32_t *t_host, const ds_buffers *io,
                 void *cuda_stream)
try {
    if (!h || !state_host || !io || !io->pos || !io->vel)
        return fail(DS_ERR_ARG, "ds_get_state: NULL argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t A = (size_t)h->E * h->n;
    const size_t bytes = A * 2 * h->real_bytes;
    std::vector<unsigned char> pos(bytes), vel(bytes);
    DS_CUDA(cudaMemcpyAsync(pos.data(), io->pos, bytes, cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaMemcpyAsync(vel.data(), io->vel, bytes, cudaMemcpyDeviceToHost, st));
    if (t_host && io->t)
        DS_CUDA(cudaMemcpyAsync(t_host, io->t, sizeof(int32_t) * h->E, cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    for (size_t a = 0; a < A; ++a) {
        double *row = state_host + a * 5;
        if (h->real_bytes == 8) {
            const double *pp = (const double *)pos.data() + 2 * a, *vv = (const double *)vel.data() + 2 * a;
            row[0] = pp[0]; row[1] = pp[1]; row[2] = vv[0]; row[3] = vv[1];
        } else {
            const float *pp = (const float *)pos.data() + 2 * a, *vv = (const float *)vel.data() + 2 * a;
            row[0] = pp[0]; row[1] = pp[1]; row[2] = vv[0]; row[3] = vv[1];
        }
        row[4] = h->h_radius[a % h->n];
    }
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_get_state: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_get_state: unknown host exception");
}

int ds_reset(ds_handle *h, const double *pos_host, const ds_params *p, const ds_buffers *io,
             void *cuda_stream)
try {
    if (!h || !pos_host || !io || !io->pos || !io->vel)
        return fail(DS_ERR_ARG, "ds_reset: NULL argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t A = (size_t)h->E * h->n;
    const size_t bytes = A * 2 * h->real_bytes;
    std::vector<float> tmp;
    const void *src = pos_host;
    if (h->real_bytes == 4) {
        tmp.resize(A * 2);
        for (size_t a = 0; a < 2 * A; ++a) tmp[a] = (float)pos_host[a];
        src = tmp.data();
    }
    DS_CUDA(cudaMemcpyAsync(io->pos, src, bytes, cudaMemcpyHostToDevice, st));
    DS_CUDA(cudaMemsetAsync(io->vel, 0, bytes, st));
    if (io->t) DS_CUDA(cudaMemsetAsync(io->t, 0, sizeof(int32_t) * h->E, st));
    if (io->finished) DS_CUDA(cudaMemsetAsync(io->finished, 0, h->E, st));
    DS_CUDA(cudaStreamSynchronize(st));   // tmp must outlive the copy
    return ds_observe(h, p, io, cuda_stream);
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_reset: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_reset: unknown host exception");
}

int ds_reset_random(ds_handle *h, uint64_t seed, uint32_t stream, int32_t d0, int32_t d1, double pitch,
                    const ds_params *p, const ds_buffers *io, void *cuda_stream)
try {
    if (!h || !io || !io->pos || !io->vel) return fail(DS_ERR_ARG, "ds_reset_random: NULL argument");
    if (d0 < 1 || d1 < 1 || (long long)d0 * d1 > 0x7fffffffLL)
        return fail(DS_ERR_ARG, "ds_reset_random: bad lattice shape");
    if ((long long)d0 * d1 < h->n)
        return fail(DS_ERR_ARG, "ds_reset_random: sample larger than population");   // as random.sample
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    ds::ResetArgs a;
    a.E = h->E; a.n = h->n; a.d0 = d0; a.d1 = d1; a.real_bytes = h->real_bytes;
    a.seed_lo = (unsigned)seed; a.seed_hi = (unsigned)(seed >> 32); a.stream = stream;
    a.pitch = pitch;
    a.pos = io->pos; a.vel = io->vel; a.t = io->t; a.fin = io->finished;
    const int threads = 128, warps = threads / 32;         // one warp per environment
    const int blocks = (h->E + warps - 1) / warps;
    if (h->n <= 32 && h->k == 2 && io->reward && io->true_reward && io->z && io->Ni && io->ncoll &&
        env_int("DS_RESET_FUSED", 1)) {
        // draws and the start state's observation in one launch (lane = agent)
        ds::ResetObsArgs ro;
        ro.r = a;
        if (int rc = fill_step_args(h, p, io, nullptr, false, &ro.s)) return rc;
        const size_t smem2 = ds::CtaSmem<double>::align16(sizeof(int) * (size_t)h->n * warps) +
                             (size_t)warps * h->n * 2 * h->real_bytes;
        if (h->real_bytes == 8) ds::reset_observe_kernel<double><<<blocks, threads, smem2, st>>>(ro);
        else ds::reset_observe_kernel<float><<<blocks, threads, smem2, st>>>(ro);
        DS_CUDA(cudaGetLastError());
        return DS_OK;
    }
    const size_t smem = sizeof(int) * (size_t)h->n * warps;
    if (h->real_bytes == 8) ds::reset_random_kernel<double><<<blocks, threads, smem, st>>>(a);
    else ds::reset_random_kernel<float><<<blocks, threads, smem, st>>>(a);
    DS_CUDA(cudaGetLastError());
    return ds_observe(h, p, io, cuda_stream);
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_reset_random: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_reset_random: unknown host exception");
}

struct ds_policy {
    int n, in_dim, A, real_bytes, device;
    float *head = nullptr, *W2p = nullptr, *W3t = nullptr;   // packed parameter blocks (dronestep_policy.cuh)
    void *atable = nullptr;
};

int ds_policy_create(const ds_policy_config *cfg, ds_policy **out)
try {
    if (!cfg || !out) return fail(DS_ERR_ARG, "ds_policy_create: NULL argument");
    *out = nullptr;
    if (cfg->n_agents < 1 || cfg->in_dim < 1 || cfg->in_dim > ds::kPolMaxIn || cfg->n_actions < 1 ||
        cfg->n_actions > ds::kPolMaxA || (cfg->real_bytes != 4 && cfg->real_bytes != 8))
        return fail(DS_ERR_ARG, "ds_policy_create: need in_dim <= 16, n_actions <= 16, real_bytes 4 or 8");
    if (!cfg->W1 || !cfg->b1 || !cfg->W2 || !cfg->b2 || !cfg->W3 || !cfg->b3 || !cfg->action_table)
        return fail(DS_ERR_ARG, "ds_policy_create: weight arrays must be non-NULL");
    const int ndev = ds_device_count();
    if (ndev < 1) return fail(DS_ERR_NO_DEVICE, "ds_policy_create: no CUDA device visible; libdronestep has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(DS_ERR_ARG, "ds_policy_create: bad device ordinal");
    DeviceGuard guard(cfg->device);
    const int n = cfg->n_agents, H = ds::kPolHidden, NP = ds::kPolNP, KP = ds::kPolKP, A = cfg->n_actions;
    // W2 -> [agent][chunk of 32 k][hi, lo][k-group of 4][304 rows][4]: the K-major core-matrix layout the
    // kernel's shared-memory descriptors describe; hi = the 19 bits kind::tf32 reads, lo = the rest
    // W1 (transposed), b1, b2, b3 -> one "head" block per agent; W3 -> transposed [304][16]
    std::vector<float> w2p((size_t)n * (KP / 32) * 2 * 8 * NP * 4, 0.f), headp((size_t)n * ds::kPolHeadFloats, 0.f),
        w3t((size_t)n * NP * ds::kPolMaxA, 0.f);
    for (int i = 0; i < n; ++i) {
        float *hd = &headp[(size_t)i * ds::kPolHeadFloats];
        for (int j = 0; j < H; ++j) {
            for (int d = 0; d < cfg->in_dim; ++d) hd[ds::kPolHeadW1 + d * KP + j] = cfg->W1[((size_t)i * H + j) * cfg->in_dim + d];
            hd[ds::kPolHeadB1 + j] = cfg->b1[(size_t)i * H + j];
            hd[ds::kPolHeadB2 + j] = cfg->b2[(size_t)i * H + j];
        }
        for (int q = 0; q < ds::kPolMaxA; ++q) hd[ds::kPolHeadB3 + q] = q < A ? cfg->b3[(size_t)i * A + q] : -INFINITY;
        for (int r = 0; r < H; ++r) {
            for (int k = 0; k < H; ++k) {
                const float w = cfg->W2[((size_t)i * H + r) * H + k];
                uint32_t bits; std::memcpy(&bits, &w, 4);
                bits &= 0xffffe000u;
                float hi; std::memcpy(&hi, &bits, 4);
                const int kc = k / 32, kg = (k % 32) / 4, q = k % 4;
                const size_t base = ((((size_t)i * (KP / 32) + kc) * 2) * 8 + kg) * NP;
                w2p[(base + r) * 4 + q] = hi;
                w2p[(base + (size_t)8 * NP + r) * 4 + q] = w - hi;
            }
        }
        for (int aidx = 0; aidx < A; ++aidx)
            for (int k = 0; k < H; ++k)
                w3t[((size_t)i * NP + k) * ds::kPolMaxA + aidx] = cfg->W3[((size_t)i * A + aidx) * H + k];
    }
    ds_policy *p = new ds_policy();
    p->n = n; p->in_dim = cfg->in_dim; p->A = A; p->real_bytes = cfg->real_bytes; p->device = cfg->device;
    auto up = [](float **dst, const float *src, size_t count) -> cudaError_t {
        cudaError_t e = cudaMalloc((void **)dst, count * sizeof(float));
        return e != cudaSuccess ? e : cudaMemcpy(*dst, src, count * sizeof(float), cudaMemcpyHostToDevice);
    };
    cudaError_t e = up(&p->head, headp.data(), headp.size());
    if (e == cudaSuccess) e = up(&p->W2p, w2p.data(), w2p.size());
    if (e == cudaSuccess) e = up(&p->W3t, w3t.data(), w3t.size());
    if (e == cudaSuccess) e = cudaMalloc(&p->atable, (size_t)A * 2 * cfg->real_bytes);
    if (e == cudaSuccess) {
        if (cfg->real_bytes == 8) e = cudaMemcpy(p->atable, cfg->action_table, (size_t)A * 16, cudaMemcpyHostToDevice);
        else {
            std::vector<float> t((size_t)A * 2);
            for (int q = 0; q < 2 * A; ++q) t[q] = (float)cfg->action_table[q];
            e = cudaMemcpy(p->atable, t.data(), (size_t)A * 8, cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) { ds_policy_destroy(p); return fail(DS_ERR_CUDA, std::string("ds_policy_create: ") + cudaGetErrorString(e)); }
    *out = p;
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_policy_create: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_policy_create: unknown host exception");
}

void ds_policy_destroy(ds_policy *p)
{
    if (!p) return;
    DeviceGuard guard(p->device);
    cudaFree(p->head); cudaFree(p->W2p); cudaFree(p->W3t); cudaFree(p->atable);
    delete p;
}

static int check_policy(const char *who, ds_handle *h, ds_policy *pol)
{
    if (pol->n != h->n) return fail(DS_ERR_ARG, std::string(who) + ": the policy has a different number of agents");
    if (pol->real_bytes != h->real_bytes || pol->device != h->device)
        return fail(DS_ERR_ARG, std::string(who) + ": the policy was created for another precision / device");
    if (pol->in_dim != (h->k + 1) * (h->simplify ? 2 : 5))
        return fail(DS_ERR_ARG, std::string(who) + ": in_dim != (k + 1) * cols of the observation");
    return DS_OK;
}

static int launch_policy(ds_handle *h, ds_policy *pol, const ds::PolicyArgs &a, cudaStream_t st)
{
    if (h->E <= 0) return DS_OK;
    // persistent: one CTA per SM walks a contiguous range of (agent, 128-environment) tiles
    const long long tiles = (long long)((h->E + 127) / 128) * h->n;
    const dim3 grid((unsigned)std::min<long long>(tiles, h->sm_count));
    const size_t smem = sizeof(ds::PolicySmem) + 128;
#define DS_POLICY_LAUNCH(REAL, IN)                                                                      \
    do {                                                                                                \
        DS_CUDA(cudaFuncSetAttribute(ds::policy_kernel<REAL, IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem));                                                       \
        ds::policy_kernel<REAL, IN><<<grid, ds::kPolThreads, smem, st>>>(a);                                        \
    } while (0)
    const int in_sel = (pol->in_dim == 6) ? 6 : (pol->in_dim == 15 ? 15 : 0);
    if (h->real_bytes == 8) {
        if (in_sel == 6) DS_POLICY_LAUNCH(double, 6); else if (in_sel == 15) DS_POLICY_LAUNCH(double, 15);
        else DS_POLICY_LAUNCH(double, 0);
    } else {
        if (in_sel == 6) DS_POLICY_LAUNCH(float, 6); else if (in_sel == 15) DS_POLICY_LAUNCH(float, 15);
        else DS_POLICY_LAUNCH(float, 0);
    }
#undef DS_POLICY_LAUNCH
    DS_CUDA(cudaGetLastError());
    return DS_OK;
}

static ds::PolicyArgs policy_args(ds_handle *h, ds_policy *pol)
{
    ds::PolicyArgs a;
    std::memset(&a, 0, sizeof a);
    a.E = h->E; a.n = h->n; a.in_dim = pol->in_dim; a.n_actions = pol->A; a.real_bytes = h->real_bytes;
    a.head = pol->head; a.W2p = pol->W2p; a.W3t = pol->W3t; a.atable = pol->atable;
    a.tiles_per_agent = (h->E + 127) / 128;
    return a;
}

int ds_policy_forward(ds_handle *h, ds_policy *pol, const ds_policy_io *io, void *cuda_stream)
try {
    if (!h || !pol || !io || !io->z || !io->actions) return fail(DS_ERR_ARG, "ds_policy_forward: NULL argument");
    if (int rc = check_policy("ds_policy_forward", h, pol)) return rc;
    DeviceGuard guard(h->device);
    ds::PolicyArgs a = policy_args(h, pol);
    a.seed_lo = (unsigned)io->seed; a.seed_hi = (unsigned)(io->seed >> 32); a.stream = io->stream;
    a.z = io->z; a.act = io->actions; a.aidx = io->action_idx; a.probs = io->probs;
    return launch_policy(h, pol, a, (cudaStream_t)cuda_stream);
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_policy_forward: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_policy_forward: unknown host exception");
}

int ds_rollout_policy(ds_handle *h, ds_policy *pol, const ds_params *p, const ds_buffers *io, const ds_rollout_io *ro,
                      const ds_policy_rollout_io *pio, void *cuda_stream)
try {
    if (!h || !pol || !pio) return fail(DS_ERR_ARG, "ds_rollout_policy: NULL argument");
    ds::RolloutArgs ra;
    std::memset(&ra, 0, sizeof ra);
    if (int rc = fill_step_args(h, p, io, nullptr, true, &ra.s)) return rc;
    if (int rc = check_policy("ds_rollout_policy", h, pol)) return rc;
    if (!ro) return fail(DS_ERR_ARG, "ds_rollout_policy: ds_rollout_io is NULL");
    if (ro->T < 0) return fail(DS_ERR_ARG, "ds_rollout_policy: T < 0");
    if (ro->actions || ro->action_idx) return fail(DS_ERR_ARG, "ds_rollout_policy: the actions come from the policy; ro->actions / action_idx must be NULL");
    if (!ro->agg || !ro->done) return fail(DS_ERR_ARG, "ds_rollout_policy: agg/done must be non-NULL");
    if ((ro->z_tr == nullptr) != (ro->Ni_tr == nullptr))
        return fail(DS_ERR_ARG, "ds_rollout_policy: z_tr and Ni_tr must be given together");
    if (ro->T == 0) return DS_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t E = (size_t)h->E, EN = E * h->n, rb = (size_t)h->real_bytes;
    const size_t zc = (size_t)(h->k + 1) * (h->simplify ? 2 : 5);
    if (!ro->vel_tr && h->act_stage_bytes < EN * 2 * rb) {          // the step's actions when they are not recorded
        cudaFree(h->act_stage);
        h->act_stage = nullptr; h->act_stage_bytes = 0;
        DS_CUDA(cudaMalloc(&h->act_stage, EN * 2 * rb));
        h->act_stage_bytes = EN * 2 * rb;
    }
    ds::PolicyArgs pa = policy_args(h, pol);
    pa.seed_lo = (unsigned)pio->seed; pa.seed_hi = (unsigned)(pio->seed >> 32);
    pa.seed_dev = (const unsigned long long *)pio->seed_dev;
    pa.z = io->z;
    ra.s.ctrl = 0; ra.T = 1;
    ra.agg = ro->agg; ra.done = ro->done;
    const Geom gm{h->step_blocks, h->step_threads, h->step_smem};
    auto off = [](void *base, size_t bytes) -> void * { return base ? (void *)((char *)base + bytes) : nullptr; };
    for (int t = 0; t < ro->T; ++t) {
        void *act_t = ro->vel_tr ? off(ro->vel_tr, (size_t)t * EN * 2 * rb) : h->act_stage;
        pa.stream = pio->stream0 + (unsigned)t;
        pa.act = act_t;
        pa.aidx = (uint8_t *)off(pio->action_idx_tr, (size_t)t * EN);
        pa.probs = (float *)off(pio->probs_tr, (size_t)t * EN * pol->A * sizeof(float));
        if (int rc = launch_policy(h, pol, pa, st)) return rc;
        ra.actions = act_t;
        ra.pos_tr = off(ro->pos_tr, (size_t)t * EN * 2 * rb);
        ra.vel_tr = ro->vel_tr ? act_t : nullptr;            // the kernel rewrites what it has just read
        ra.r_tr = off(ro->reward_tr, (size_t)t * EN * rb);
        ra.tr_tr = off(ro->true_reward_tr, (size_t)t * EN * rb);
        ra.z_tr = off(ro->z_tr, (size_t)t * EN * zc * rb);
        ra.Ni_tr = (int *)off(ro->Ni_tr, (size_t)t * EN * (h->k + 1) * sizeof(int32_t));
        ra.ncoll_tr = (int *)off(ro->ncoll_tr, (size_t)t * E * sizeof(int32_t));
        ra.fin_tr = (uint8_t *)off(ro->finished_tr, (size_t)t * E);
        if (int rc = launch_rollout_control(h, ra, gm, st)) return rc;
    }
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_rollout_policy: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_rollout_policy: unknown host exception");
}

int ds_step_host(ds_handle *h, const void *actions_host, const ds_params *p, const ds_buffers *io,
                 const ds_host_step_out *out, void *cuda_stream)
try {
    if (!h || !actions_host || !out) return fail(DS_ERR_ARG, "ds_step_host: NULL argument");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t A = (size_t)h->E * h->n;
    const size_t rb = h->real_bytes;
    const size_t act_bytes = A * 2 * rb;
    if (h->act_stage_bytes < act_bytes) {
        cudaFree(h->act_stage);
        h->act_stage = nullptr; h->act_stage_bytes = 0;
        DS_CUDA(cudaMalloc(&h->act_stage, act_bytes));
        h->act_stage_bytes = act_bytes;
    }
    DS_CUDA(cudaMemcpyAsync(h->act_stage, actions_host, act_bytes, cudaMemcpyHostToDevice, st));
    if (int rc = ds_step(h, h->act_stage, p, io, cuda_stream)) return rc;
    const size_t zc = (size_t)(h->k + 1) * (h->simplify ? 2 : 5);
    if (out->pos) DS_CUDA(cudaMemcpyAsync(out->pos, io->pos, A * 2 * rb, cudaMemcpyDeviceToHost, st));
    if (out->vel) DS_CUDA(cudaMemcpyAsync(out->vel, io->vel, A * 2 * rb, cudaMemcpyDeviceToHost, st));
    if (out->z) DS_CUDA(cudaMemcpyAsync(out->z, io->z, A * zc * rb, cudaMemcpyDeviceToHost, st));
    if (out->reward) DS_CUDA(cudaMemcpyAsync(out->reward, io->reward, A * rb, cudaMemcpyDeviceToHost, st));
    if (out->true_reward)
        DS_CUDA(cudaMemcpyAsync(out->true_reward, io->true_reward, A * rb, cudaMemcpyDeviceToHost, st));
    if (out->Ni)
        DS_CUDA(cudaMemcpyAsync(out->Ni, io->Ni, A * (h->k + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (out->ncoll)
        DS_CUDA(cudaMemcpyAsync(out->ncoll, io->ncoll, sizeof(int32_t) * h->E, cudaMemcpyDeviceToHost, st));
    if (out->finished) DS_CUDA(cudaMemcpyAsync(out->finished, io->finished, h->E, cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_step_host: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_step_host: unknown host exception");
}

int ds_step_host_block(ds_handle *h, const void *actions_host, const ds_params *p, const ds_buffers *io,
                       const void *dev_block, void *host_block, size_t bytes, void *cuda_stream)
try {
    if (!h || !actions_host || !io || !dev_block || !host_block)
        return fail(DS_ERR_ARG, "ds_step_host_block: NULL argument");
    const char *lo = (const char *)dev_block, *hi = lo + bytes;
    const void *ptrs[] = {io->pos, io->vel, io->reward, io->true_reward, io->z, io->Ni, io->ncoll, io->finished};
    for (const void *q : ptrs)
        if (q && ((const char *)q < lo || (const char *)q >= hi))
            return fail(DS_ERR_ARG, "ds_step_host_block: a result buffer of io lies outside the block");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t act_bytes = (size_t)h->E * h->n * 2 * h->real_bytes;
    if (h->act_stage_bytes < act_bytes) {
        cudaFree(h->act_stage);
        h->act_stage = nullptr; h->act_stage_bytes = 0;
        DS_CUDA(cudaMalloc(&h->act_stage, act_bytes));
        h->act_stage_bytes = act_bytes;
    }
    DS_CUDA(cudaMemcpyAsync(h->act_stage, actions_host, act_bytes, cudaMemcpyHostToDevice, st));
    if (int rc = ds_step(h, h->act_stage, p, io, cuda_stream)) return rc;
    DS_CUDA(cudaMemcpyAsync(host_block, dev_block, bytes, cudaMemcpyDeviceToHost, st));
    DS_CUDA(cudaStreamSynchronize(st));
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_step_host_block: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_step_host_block: unknown host exception");
}

int ds_rollout_host(ds_handle *h, const ds_params *p, const ds_buffers *io, const ds_host_rollout *hr,
                    void *cuda_stream)
try {
    if (!h || !hr || !io) return fail(DS_ERR_ARG, "ds_rollout_host: NULL argument");
    if (int rc = check_params(p)) return rc;
    if (hr->T < 0) return fail(DS_ERR_ARG, "ds_rollout_host: T < 0");
    const bool index_mode = hr->actions == nullptr;
    if (index_mode && (!hr->action_idx || !hr->action_table || hr->n_actions < 1 || hr->n_actions > 256))
        return fail(DS_ERR_ARG, "ds_rollout_host: give actions, or action_idx + action_table (1..256 rows)");
    if ((hr->z_tr == nullptr) != (hr->Ni_tr == nullptr))
        return fail(DS_ERR_ARG, "ds_rollout_host: z_tr and Ni_tr must be given together");
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t E = h->E, A = E * h->n, rb = h->real_bytes;
    const size_t zc = (size_t)(h->k + 1) * (h->simplify ? 2 : 5);
    // bytes per step of everything that crosses PCIe
    const size_t in_step = index_mode ? A : A * 2 * rb;
    size_t out_step = 0;
    unsigned mask = index_mode ? 1u : 0u;
    if (hr->pos_tr) { out_step += A * 2 * rb; mask |= 2u; }
    // state[:,2:4] = u (drone_env.py:238): the recorded velocity IS the action the host has just
    // supplied, so it is written on the host while the pipeline drains and never crosses PCIe
    if (hr->reward_tr) { out_step += A * rb; mask |= 8u; }
    if (hr->true_reward_tr) { out_step += A * rb; mask |= 16u; }
    const bool compact = (hr->flags & DS_HOST_COMPACT_OBS) != 0;
    if (compact && h->n > 255) return fail(DS_ERR_ARG, "ds_rollout_host: DS_HOST_COMPACT_OBS needs n_agents <= 255");
    if (hr->z_tr) { out_step += compact ? A * zc * 4 + A * (h->k + 1) : A * zc * rb + A * (h->k + 1) * 4; mask |= 32u; }
    if (compact) mask |= 256u;
    if (hr->ncoll_tr) { out_step += E * 4; mask |= 64u; }
    if (hr->finished_tr) { out_step += E; mask |= 128u; }
    int chunk = hr->chunk;
    if (chunk <= 0) {
        // aim at ~32 MiB per chunk: large enough to reach PCIe bandwidth, small enough to pipeline
        const size_t per_step = in_step + out_step;
        chunk = (int)((((size_t)32 << 20) + per_step - 1) / per_step);
        if (chunk < 1) chunk = 1;
    }
    if (chunk > hr->T && hr->T > 0) chunk = hr->T;

    if (!h->s_h2d) DS_CUDA(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    if (!h->s_d2h) DS_CUDA(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    if (!h->d_agg) DS_CUDA(cudaMalloc(&h->d_agg, E * 4 * sizeof(double)));
    if (!h->d_done) DS_CUDA(cudaMalloc(&h->d_done, E));
    if (index_mode) {
        if (!h->d_atable) DS_CUDA(cudaMalloc(&h->d_atable, 256 * 2 * sizeof(double)));
        DS_CUDA(cudaMemcpyAsync(h->d_atable, hr->action_table, (size_t)hr->n_actions * 2 * rb,
                                cudaMemcpyHostToDevice, st));
    }
    if (h->slot_chunk < chunk || h->slot_mask != mask) {
        DS_CUDA(cudaDeviceSynchronize());
        free_slots(h);
        for (auto &s : h->slot) {
            const size_t c = chunk;
            if (index_mode) DS_CUDA(cudaMalloc(&s.aidx, c * A)); else DS_CUDA(cudaMalloc(&s.act, c * A * 2 * rb));
            if (hr->pos_tr) DS_CUDA(cudaMalloc(&s.pos, c * A * 2 * rb));
            if (hr->reward_tr) DS_CUDA(cudaMalloc(&s.r, c * A * rb));
            if (hr->true_reward_tr) DS_CUDA(cudaMalloc(&s.tr, c * A * rb));
            if (hr->z_tr) {
                DS_CUDA(cudaMalloc(&s.z, c * A * zc * rb));
                DS_CUDA(cudaMalloc(&s.Ni, c * A * (h->k + 1) * 4));
                if (compact) {
                    DS_CUDA(cudaMalloc((void **)&s.zf, c * A * zc * 4));
                    DS_CUDA(cudaMalloc((void **)&s.Ni8, c * A * (h->k + 1)));
                }
            }
            if (hr->ncoll_tr) DS_CUDA(cudaMalloc(&s.ncoll, c * E * 4));
            if (hr->finished_tr) DS_CUDA(cudaMalloc(&s.fin, c * E));
            DS_CUDA(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
            DS_CUDA(cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming));
            DS_CUDA(cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
        }
        h->slot_chunk = chunk; h->slot_mask = mask;
    }
    DS_CUDA(cudaMemsetAsync(h->d_agg, 0, E * 4 * sizeof(double), st));
    DS_CUDA(cudaMemsetAsync(h->d_done, 0, E, st));
    // the copy streams must not start before the caller's stream reaches this point
    cudaEvent_t &start_ev = h->slot[0].d2h_done;   // reuse: recorded before any D2H of this call
    DS_CUDA(cudaEventRecord(start_ev, st));
    DS_CUDA(cudaStreamWaitEvent(h->s_h2d, start_ev, 0));
    DS_CUDA(cudaStreamWaitEvent(h->s_d2h, start_ev, 0));
    DS_CUDA(cudaEventRecord(h->slot[1].d2h_done, st));

    int ci = 0;
    for (int t0 = 0; t0 < hr->T; t0 += chunk, ++ci) {
        const int tc = (hr->T - t0 < chunk) ? hr->T - t0 : chunk;
        ds_handle::Slot &s = h->slot[ci & 1];
        const size_t t0s = t0, tcs = tc;
        // H2D: the slot's action buffer is free once the kernel of chunk ci-2 has run
        if (ci >= 2) DS_CUDA(cudaStreamWaitEvent(h->s_h2d, s.kernel_done, 0));
        if (index_mode)
            DS_CUDA(cudaMemcpyAsync(s.aidx, hr->action_idx + t0s * A, tcs * A, cudaMemcpyHostToDevice, h->s_h2d));
        else
            DS_CUDA(cudaMemcpyAsync(s.act, (const char *)hr->actions + t0s * A * 2 * rb, tcs * A * 2 * rb,
                                    cudaMemcpyHostToDevice, h->s_h2d));
        DS_CUDA(cudaEventRecord(s.h2d_done, h->s_h2d));
        // kernel: needs the actions, and the slot's output buffers drained (chunk ci-2's D2H)
        DS_CUDA(cudaStreamWaitEvent(st, s.h2d_done, 0));
        DS_CUDA(cudaStreamWaitEvent(st, s.d2h_done, 0));
        ds_rollout_io ro;
        std::memset(&ro, 0, sizeof ro);
        ro.T = tc; ro.n_actions = hr->n_actions;
        ro.actions = index_mode ? nullptr : s.act;
        ro.action_idx = s.aidx; ro.action_table = h->d_atable;
        ro.pos_tr = s.pos; ro.vel_tr = nullptr; ro.reward_tr = s.r; ro.true_reward_tr = s.tr;
        ro.z_tr = s.z; ro.Ni_tr = s.Ni; ro.ncoll_tr = s.ncoll; ro.finished_tr = s.fin;
        ro.agg = h->d_agg; ro.done = h->d_done;
        if (int rc = ds_rollout(h, p, io, &ro, cuda_stream)) return rc;
        if (compact && hr->z_tr) {
            const size_t nz = tcs * A * zc, nn = tcs * A * (h->k + 1);
            const int blocks = 4 * h->sm_count;
            if (rb == 8) compact_obs_kernel<double><<<blocks, 256, 0, st>>>((const double *)s.z, s.Ni, s.zf, s.Ni8, nz, nn);
            else compact_obs_kernel<float><<<blocks, 256, 0, st>>>((const float *)s.z, s.Ni, s.zf, s.Ni8, nz, nn);
            DS_CUDA(cudaGetLastError());
        }
        DS_CUDA(cudaEventRecord(s.kernel_done, st));
        // D2H
        DS_CUDA(cudaStreamWaitEvent(h->s_d2h, s.kernel_done, 0));
        cudaStream_t d = h->s_d2h;
        if (hr->pos_tr) DS_CUDA(cudaMemcpyAsync((char *)hr->pos_tr + t0s * A * 2 * rb, s.pos, tcs * A * 2 * rb, cudaMemcpyDeviceToHost, d));
        if (hr->reward_tr) DS_CUDA(cudaMemcpyAsync((char *)hr->reward_tr + t0s * A * rb, s.r, tcs * A * rb, cudaMemcpyDeviceToHost, d));
        if (hr->true_reward_tr) DS_CUDA(cudaMemcpyAsync((char *)hr->true_reward_tr + t0s * A * rb, s.tr, tcs * A * rb, cudaMemcpyDeviceToHost, d));
        if (hr->z_tr && compact) {
            DS_CUDA(cudaMemcpyAsync((char *)hr->z_tr + t0s * A * zc * 4, s.zf, tcs * A * zc * 4, cudaMemcpyDeviceToHost, d));
            DS_CUDA(cudaMemcpyAsync((char *)hr->Ni_tr + t0s * A * (h->k + 1), s.Ni8, tcs * A * (h->k + 1), cudaMemcpyDeviceToHost, d));
        } else if (hr->z_tr) {
            DS_CUDA(cudaMemcpyAsync((char *)hr->z_tr + t0s * A * zc * rb, s.z, tcs * A * zc * rb, cudaMemcpyDeviceToHost, d));
            DS_CUDA(cudaMemcpyAsync(hr->Ni_tr + t0s * A * (h->k + 1), s.Ni, tcs * A * (h->k + 1) * 4, cudaMemcpyDeviceToHost, d));
        }
        if (hr->ncoll_tr) DS_CUDA(cudaMemcpyAsync(hr->ncoll_tr + t0s * E, s.ncoll, tcs * E * 4, cudaMemcpyDeviceToHost, d));
        if (hr->finished_tr) DS_CUDA(cudaMemcpyAsync(hr->finished_tr + t0s * E, s.fin, tcs * E, cudaMemcpyDeviceToHost, d));
        DS_CUDA(cudaEventRecord(s.d2h_done, d));
    }
    if (hr->agg)
        DS_CUDA(cudaMemcpyAsync(hr->agg, h->d_agg, E * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (hr->vel_tr && hr->vel_tr != hr->actions && hr->T > 0) {      // vel_tr == actions: the caller aliases them
        // everything is enqueued; this thread would only wait.  A few host threads write vel_tr = u.
        const size_t total = (size_t)hr->T * A;                       // (step, environment, agent) rows
        const int nth = (int)std::min<size_t>(4, std::max<size_t>(1, total >> 16));
        auto fill = [&](size_t lo, size_t hi) {
            if (!index_mode) {
                std::memcpy((char *)hr->vel_tr + lo * 2 * rb, (const char *)hr->actions + lo * 2 * rb, (hi - lo) * 2 * rb);
            } else if (rb == 8) {
                const double *tab = (const double *)hr->action_table; double *o = (double *)hr->vel_tr;
                for (size_t q = lo; q < hi; ++q) { const unsigned a = hr->action_idx[q]; o[2 * q] = tab[2 * a]; o[2 * q + 1] = tab[2 * a + 1]; }
            } else {
                const float *tab = (const float *)hr->action_table; float *o = (float *)hr->vel_tr;
                for (size_t q = lo; q < hi; ++q) { const unsigned a = hr->action_idx[q]; o[2 * q] = tab[2 * a]; o[2 * q + 1] = tab[2 * a + 1]; }
            }
        };
        std::vector<std::thread> pool;
        pool.reserve(nth);
        int spawned = 1;                                              // part 0 is this thread's
        try {
            for (; spawned < nth; ++spawned) pool.emplace_back(fill, total * spawned / nth, total * (spawned + 1) / nth);
        } catch (...) {                                               // no more threads: this one does the rest
        }
        fill(0, total / nth);
        if (spawned < nth) fill(total * spawned / nth, total);
        for (auto &t : pool) t.join();
    }
    DS_CUDA(cudaStreamSynchronize(st));
    DS_CUDA(cudaStreamSynchronize(h->s_d2h));
    DS_CUDA(cudaStreamSynchronize(h->s_h2d));
    return DS_OK;
} catch (const std::exception &ex) {   // nothing C++ crosses the C boundary
    return fail(DS_ERR_INTERNAL, std::string("ds_rollout_host: host exception: ") + ex.what());
} catch (...) {
    return fail(DS_ERR_INTERNAL, "ds_rollout_host: unknown host exception");
}

}  // extern "C"
