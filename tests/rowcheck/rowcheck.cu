// TEST INFRASTRUCTURE ONLY -- runs the row logic of dronestep_kernels.cuh (eval_row / the work-list
// path eval_pair + eval_row_from_list, write_obs: the __host__ __device__ part) on the CPU for one
// frame, so that the near/clipped split, the Delta-disk count and the k-nearest selection can be
// checked against the oracle without a GPU.  path = 0: whole row in one call (step kernel);
// path = 1: near pairs through a work list exactly as the rollout kernel lays them out;
// path = 2: the rollout kernel's inline mode for n <= 32.
// Never linked into libdronestep.so; the product has no CPU path.
#include "../../scalable_collision_avoidance_rl_b200/csrc/dronestep_kernels.cuh"
#include <vector>
#include <cmath>

template <typename Real, int K>
static void run(int path, int n, int k, int simplify, int log_mode, const double *pos, const double *vel,
                const double *xF, const double *dsf, const double *delta, const double *radius,
                double collision_weight, double *r, double *tr, double *z, int *Ni, int *ncoll, int *notgoal)
{
    using V2 = typename ds::vec2_of<Real>::type;
    std::vector<Real> cxF(2 * n), cds(n), cdl(n), crd(n), clg(n), cthr(n);
    std::vector<int> clip(n);
    std::vector<V2> sp(n), sv(n);
    Real rmax = 0;
    for (int i = 0; i < n; ++i) {
        cxF[2 * i] = (Real)xF[2 * i]; cxF[2 * i + 1] = (Real)xF[2 * i + 1];
        cds[i] = (Real)dsf[i]; cdl[i] = (Real)delta[i]; crd[i] = (Real)radius[i];
        clg[i] = (Real)std::log(std::fabs(dsf[i]));
        sp[i].x = (Real)pos[2 * i]; sp[i].y = (Real)pos[2 * i + 1];
        sv[i].x = (Real)vel[2 * i]; sv[i].y = (Real)vel[2 * i + 1];
        rmax = crd[i] > rmax ? crd[i] : rmax;
    }
    const double m = sizeof(Real) == 8 ? 1e-12 : 1e-5;   // mirrors upload_consts() in dronestep_abi.cu
    for (int i = 0; i < n; ++i) {
        const double D = (double)cds[i] + (double)crd[i] + (double)rmax;
        double t2 = INFINITY;
        if (cds[i] != (Real)0 && std::isfinite(D) && D > 1e-6) { const double thr = D * (1 + m) + m; t2 = thr * thr * (1 + m); }
        cthr[i] = (Real)t2;
        if (!(cthr[i] >= t2))
            cthr[i] = (sizeof(Real) == 8) ? (Real)std::nextafter((double)cthr[i], (double)INFINITY)
                                          : (Real)std::nextafterf((float)cthr[i], INFINITY);
        int cc = 0;
        for (int j = 0; j < n; ++j) if (j != i && cds[i] <= cdl[j]) ++cc;
        clip[i] = cc;
    }
    ds::StepArgs a{};
    a.n = n; a.k = k; a.simplify = simplify; a.log_mode = log_mode;
    a.dt = 0.05; a.q = 2 * 0.05; a.b = collision_weight * 0.05; a.goal_tol = 0.2; a.sentinel = 9.99E3;
    a.zero_eps = -1e-6; a.ghost = 1.1;
    std::vector<ds::LogTabEntry> tab(ds::kLogTabSize);
    ds::fill_log_table(tab.data());
    a.c = ds::Consts{cxF.data(), cds.data(), cdl.data(), crd.data(), clg.data(), cthr.data(), clip.data(), tab.data()};
    const ds::ParamsR<Real> P(a);
    const int cols = simplify ? 2 : 5;
    std::vector<Real> zr((size_t)n * (k + 1) * cols);
    *ncoll = 0; *notgoal = 0;
    for (int i = 0; i < n; ++i) {
        const ds::AgentConst<Real> c = ds::load_agent_const<Real>(a.c, i);
        ds::RowResult<Real, K> o;
        if (path == 0) {
            ds::eval_row<Real, K>(o, n, i, sp[i].x, sp[i].y, c, sp.data(), cdl.data(), crd.data(), P, tab.data());
        } else if (path == 2) {
            // rollout kernel, inline mode (n <= 32): pass-1 mask, then the row folds its own near pairs
            unsigned near = 0;
            std::vector<V2> cpair(n);
            for (int j = 0; j < n; ++j) {
                cpair[j].x = crd[j]; cpair[j].y = cdl[j];
                if (j == i) continue;
                const Real dx = sp[i].x - sp[j].x, dy = sp[i].y - sp[j].y;
                const Real d2 = ds::fma_rn(dy, dy, dx * dx);
                if (!(d2 >= c.thr2)) near |= 1u << j;
            }
            ds::eval_row_near32<Real, K>(o, n, i, sp[i].x, sp[i].y, c, near, sp.data(), cpair.data(), P, tab.data());
        } else {
            // rollout kernel phases (c) -> (d) -> (e) for this row
            std::vector<unsigned> ent;
            std::vector<V2> res;
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                const Real dx = sp[i].x - sp[j].x, dy = sp[i].y - sp[j].y;
                const Real d2 = ds::fma_rn(dy, dy, dx * dx);
                if (!(d2 >= c.thr2)) ent.push_back(ds::pack_entry(i, j, i));
            }
            res.resize(ent.size());
            for (size_t q = 0; q < ent.size(); ++q) {
                const unsigned w = ent[q];
                const int row = (int)(w & 1023u), j = (int)((w >> 10) & 1023u), ri = (int)(w >> 20);
                ds::PairOut<Real> po;
                ds::eval_pair<Real>(po, sp[row].x, sp[row].y, sp[row - ri + j].x, sp[row - ri + j].y, cds[ri], crd[ri],
                                    crd[j], cdl[j], clg[ri], P, tab.data());
                res[q].x = po.d; res[q].y = po.logd;
                ent[q] = ds::pack_result(j, po.in_disk, po.coll, (po.in_disk ? 1 : 0) - ((cds[ri] <= cdl[j]) ? 1 : 0) + 1);
            }
            ds::eval_row_from_list<Real, K>(o, n, i, sp[i].x, sp[i].y, c, ent.data(), res.data(), (int)ent.size(), P);
        }
        ds::write_obs<Real, K>(o, i, sp[i].x, sp[i].y, c, sp.data(), sv.data(), crd.data(), P, zr.data(), Ni, (size_t)i);
        r[i] = o.r; tr[i] = o.tr; *ncoll += o.ncoll; if (!o.at_goal) *notgoal = 1;
    }
    for (size_t q = 0; q < zr.size(); ++q) z[q] = zr[q];
}

// control_action (baseline controllers) for one frame on the host
extern "C" int rowcheck_control(int mode, int n, const double *pos, const double *xF, const double *dsf,
                                const double *radius, double u_max, double *act)
{
    std::vector<double2> sp(n);
    for (int i = 0; i < n; ++i) { sp[i].x = pos[2 * i]; sp[i].y = pos[2 * i + 1]; }
    for (int i = 0; i < n; ++i)
        ds::control_action<double>(mode, n, i, sp[i].x, sp[i].y, xF[2 * i], xF[2 * i + 1], dsf[i], radius[i],
                                   sp.data(), radius, 1, u_max, act[2 * i], act[2 * i + 1]);
    return 0;
}

extern "C" double rowcheck_log(double x)
{
    static std::vector<ds::LogTabEntry> tab;
    if (tab.empty()) { tab.resize(ds::kLogTabSize); ds::fill_log_table(tab.data()); }
    return ds::log_r(x, tab.data());
}

extern "C" int rowcheck_frame(int path, int real_bytes, int n, int k, int simplify, int log_mode, const double *pos,
                              const double *vel, const double *xF, const double *dsf, const double *delta,
                              const double *radius, double collision_weight, double *r, double *tr, double *z,
                              int *Ni, int *ncoll, int *notgoal)
{
#define RUN(REAL, KK) run<REAL, KK>(path, n, k, simplify, log_mode, pos, vel, xF, dsf, delta, radius, collision_weight, r, tr, z, Ni, ncoll, notgoal)
    if (real_bytes == 8) { if (k == 2) RUN(double, 2); else RUN(double, -1); }
    else { if (k == 2) RUN(float, 2); else RUN(float, -1); }
    return 0;
}
