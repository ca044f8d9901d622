/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the drone_env.step() hot path.
 *
 * Plain-C, IEEE float64 restatement of the reference algorithm
 * (AndreuMatoses/scalable-collision-avoidance-RL, drone_env.py).  It exists to
 * CHECK the CUDA path; it is never on the product path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against golden vectors recorded from the unmodified reference running
 * in the build container (oracle/make_golden.py -> tests/golden/).
 *
 * Reference citations (file:line in /root/reference):
 *   integrate          drone_env.py:227-238   x <- I x + (dt I) u ; v <- u
 *   distance_data      drone_env.py:295-334
 *   rewards            drone_env.py:260-293
 *   localized_states   drone_env.py:336-401
 *   termination        drone_env.py:248-256
 *
 * Rounding notes (verified against the live reference, NumPy 2.3.5/OpenBLAS):
 *   - np.linalg.norm of a 1-D 2-vector is sqrt(ddot(v,v)); OpenBLAS evaluates
 *     the 2-element ddot as fma(y, y, x*x).  Used for pair distances
 *     (drone_env.py:318) and the ghost direction (drone_env.py:386).
 *   - np.linalg.norm(..., axis=1) is sqrt(x*x + y*y), unfused
 *     (drone_env.py:249,276); np.power(v, 2) is v*v.
 *   - row sums (np.sum(.., 1), drone_env.py:282-283) use NumPy's pairwise
 *     summation; this file sums sequentially.  The difference is O(1e-16)
 *     relative, far inside the 1e-5 contract; collision counts are integer and
 *     exact.
 *   Build with -ffp-contract=off so the compiler adds no FMAs of its own.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double dt;               /* drone_env.py:29  */
    double collision_weight; /* drone_env.py:72  */
    double goal_tol;         /* 0.2,   drone_env.py:251 */
    double sentinel;         /* 9.99E3, drone_env.py:330-332 */
    double zero_eps;         /* -1e-6, drone_env.py:320 */
    double ghost_factor;     /* 1.1,   drone_env.py:386 */
    int32_t max_time_steps;  /* drone_env.py:30  */
    int32_t _pad;
} oracle_params;

static double nan_to_num(double v)
{
    if (isnan(v)) return 0.0;
    if (isinf(v)) return v > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return v;
}

/* Python's builtin min(a, b): returns a unless b < a (drone_env.py:318,323). */
static double py_min(double a, double b) { return (b < a) ? b : a; }

static double norm2_blas(double x, double y) { return sqrt(fma(y, y, x * x)); }
static double norm2_axis(double x, double y) { return sqrt(x * x + y * y); }

/*
 * One environment, one evaluation of rewards() on the current state
 * (drone_env.py:260-293), i.e. distance_data + reward sums + localized_states.
 *
 *   pos[n][2], vel[n][2], radius[n]      current state (columns of state[n,5])
 *   z[n][k+1][cols]  cols = 2 (simplify_zstate) or 5
 *   Ni[n][k+1]       neighbour lists, padded with -1; Ni[i][0] == i
 *   tie[n]           (optional) 1 where the k+2 smallest entries of row i of
 *                    d_ij contain an exact tie, i.e. where np.argsort's
 *                    unstable order makes the reference's z/Ni platform
 *                    dependent
 *   scratch          n*n doubles + n*n ints
 */
static void observe_one(int n, int k, int simplify, const oracle_params *p,
                        const double *pos, const double *vel, const double *radius,
                        const double *xF, const double *d_safety, const double *deltas,
                        double *r, double *true_r, double *z, int32_t *Ni,
                        int32_t *ncoll_out, uint8_t *tie, double *scratch)
{
    const int cols = simplify ? 2 : 5;
    const double q = 2 * p->dt;                  /* :269 */
    const double b = p->collision_weight * p->dt; /* :270 */
    double *dmat = scratch;                      /* d_ij, row major */
    int32_t *order = (int32_t *)(scratch + (size_t)n * n);
    int64_t ncoll = 0;

    for (int i = 0; i < n; ++i) {
        const double xi = pos[2 * i], yi = pos[2 * i + 1], li = radius[i];
        double sum_local = 0.0, sum_all = 0.0;
        int in_range = -1; /* :346 */
        for (int j = 0; j < n; ++j) {
            double d, dn;
            if (j != i) {
                const double dx = xi - pos[2 * j], dy = yi - pos[2 * j + 1];
                d = py_min(norm2_blas(dx, dy) - li - radius[j], d_safety[i]); /* :318 */
                if (d == 0) d = p->zero_eps;                                   /* :319-320 */
                dn = d_safety[i] / d;                                          /* :321 */
            } else {
                d = py_min(-li - li, d_safety[i]);                             /* :323 */
                dn = 1;                                                        /* :325 */
            }
            const int coll = dn <= 0;          /* :327 */
            const int nd = d <= deltas[j];     /* :328, broadcast over columns */
            const double logd = coll ? p->sentinel : log(dn); /* :330-332 */
            ncoll += coll;
            in_range += nd;
            sum_local += logd * (nd ? 1.0 : 0.0); /* :282 */
            sum_all += logd;                      /* :283 */
            dmat[(size_t)i * n + j] = d;
        }
        const double gx = xF[2 * i] - xi, gy = xF[2 * i + 1] - yi;
        const double nrm = norm2_axis(gx, gy);
        const double goal = q * (nrm * nrm);                       /* :276 */
        r[i] = -nan_to_num(goal + b * sum_local);                 /* :282,287 */
        true_r[i] = -nan_to_num(goal + b * sum_all);              /* :283,288 */

        /* k+2 smallest of row i in stable order (np.argsort(d_ij,1), :338; only
           sorted[1..k] is consumed, :364,383; ties -> lowest index first) */
        int32_t *ord = order + (size_t)i * n;
        const double *row = dmat + (size_t)i * n;
        const int keep = (k + 2 < n) ? k + 2 : n;
        int have = 0;
        for (int j = 0; j < n; ++j) {
            int m = have;
            if (have == keep) {
                if (!(row[ord[keep - 1]] > row[j])) continue;
                m = keep - 1;
            } else {
                ++have;
            }
            while (m > 0 && row[ord[m - 1]] > row[j]) { ord[m] = ord[m - 1]; --m; }
            ord[m] = j;
        }
        if (tie) {
            uint8_t t = 0;
            for (int m = 0; m + 1 < keep && m <= k; ++m)
                if (row[ord[m]] == row[ord[m + 1]]) t = 1;
            tie[i] = t;
        }

        double *Zi = z + (size_t)i * (k + 1) * cols;
        int32_t *Nl = Ni + (size_t)i * (k + 1);
        const double zx = -(xF[2 * i] - xi), zy = -(xF[2 * i + 1] - yi); /* :357 */
        Zi[0] = zx; Zi[1] = zy;
        if (!simplify) { Zi[2] = vel[2 * i]; Zi[3] = vel[2 * i + 1]; Zi[4] = li; }
        Nl[0] = i;
        int nn = 1;
        for (int kth = 1; kth <= k; ++kth) {
            const int j = ord[kth];
            double *row_z = Zi + (size_t)kth * cols;
            if (kth <= in_range) {                                   /* :362-368 */
                Nl[nn++] = j;
                row_z[0] = pos[2 * j] - xi;
                row_z[1] = pos[2 * j + 1] - yi;
            } else {                                                 /* :383-386 */
                const double zn = norm2_blas(zx, zy);
                row_z[0] = zx / zn * deltas[i] * p->ghost_factor;
                row_z[1] = zy / zn * deltas[i] * p->ghost_factor;
            }
            if (!simplify) { row_z[2] = vel[2 * j]; row_z[3] = vel[2 * j + 1]; row_z[4] = radius[j]; }
        }
        for (; nn <= k; ++nn) Nl[nn] = -1;
    }
    *ncoll_out = (int32_t)ncoll; /* :284 */
}

/* finished flag of drones.step (drone_env.py:248-256); bumps *t. */
static uint8_t finish_one(int n, const oracle_params *p, const double *pos, const double *xF, int32_t *t)
{
    int all_in = 1;
    for (int i = 0; i < n; ++i) {
        const double e = norm2_axis(xF[2 * i] - pos[2 * i], xF[2 * i + 1] - pos[2 * i + 1]);
        if (!(e <= p->goal_tol)) all_in = 0;
    }
    const uint8_t fin = (all_in || *t >= p->max_time_steps - 1) ? 1 : 0;
    *t += 1;
    return fin;
}

static void integrate_one(int n, const oracle_params *p, double *pos, double *vel, const double *act)
{
    for (int i = 0; i < 2 * n; ++i) {
        pos[i] = pos[i] + p->dt * act[i]; /* A = I, B = dt I  (:78-79,235) */
        vel[i] = act[i];                  /* :238 */
    }
}

/* ------------------------------------------------------------------------- */
/* Batched entry points (E independent environments, env-major SoA layout).   */

typedef struct {
    int E, n, k, simplify, do_step;
    const oracle_params *p;
    double *pos, *vel;
    const double *radius, *act, *xF, *d_safety, *deltas;
    double *r, *true_r, *z;
    int32_t *Ni, *ncoll, *t;
    uint8_t *finished, *tie;
    int e0, e1;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *jb = (batch_job *)arg;
    const int n = jb->n, k = jb->k, cols = jb->simplify ? 2 : 5;
    double *scratch = (double *)malloc(sizeof(double) * (size_t)n * n + sizeof(int32_t) * (size_t)n * n + 64);
    for (int e = jb->e0; e < jb->e1; ++e) {
        double *pos = jb->pos + (size_t)e * n * 2, *vel = jb->vel + (size_t)e * n * 2;
        if (jb->do_step) integrate_one(n, jb->p, pos, vel, jb->act + (size_t)e * n * 2);
        observe_one(n, k, jb->simplify, jb->p, pos, vel, jb->radius, jb->xF, jb->d_safety, jb->deltas,
                    jb->r + (size_t)e * n, jb->true_r + (size_t)e * n,
                    jb->z + (size_t)e * n * (k + 1) * cols, jb->Ni + (size_t)e * n * (k + 1),
                    jb->ncoll + e, jb->tie ? jb->tie + (size_t)e * n : NULL, scratch);
        if (jb->do_step) jb->finished[e] = finish_one(n, jb->p, pos, jb->xF, jb->t + e);
    }
    free(scratch);
    return NULL;
}

static int run_batch(batch_job *proto, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > proto->E) nthreads = proto->E > 0 ? proto->E : 1;
    if (nthreads == 1) {
        proto->e0 = 0; proto->e1 = proto->E;
        batch_worker(proto);
        return 0;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * nthreads);
    for (int w = 0; w < nthreads; ++w) {
        jobs[w] = *proto;
        jobs[w].e0 = (int)((int64_t)proto->E * w / nthreads);
        jobs[w].e1 = (int)((int64_t)proto->E * (w + 1) / nthreads);
        pthread_create(&th[w], NULL, batch_worker, &jobs[w]);
    }
    for (int w = 0; w < nthreads; ++w) pthread_join(th[w], NULL);
    free(th); free(jobs);
    return 0;
}

/*
 * drones.step() over E environments.  act == NULL evaluates rewards() on the
 * current state without integrating or touching t/finished (the call made by
 * init_agents, drone_env.py:208).
 */
int oracle_step_batch(int E, int n, int k, int simplify, const oracle_params *p,
                      double *pos, double *vel, const double *radius, const double *act,
                      const double *xF, const double *d_safety, const double *deltas,
                      double *r, double *true_r, double *z, int32_t *Ni,
                      int32_t *ncoll, uint8_t *finished, int32_t *t, uint8_t *tie, int nthreads)
{
    if (E < 0 || n < 1 || k < 0 || k >= n) return -1;
    batch_job jb;
    memset(&jb, 0, sizeof jb);
    jb.E = E; jb.n = n; jb.k = k; jb.simplify = simplify; jb.do_step = act != NULL; jb.p = p;
    jb.pos = pos; jb.vel = vel; jb.radius = radius; jb.act = act; jb.xF = xF;
    jb.d_safety = d_safety; jb.deltas = deltas; jb.r = r; jb.true_r = true_r; jb.z = z;
    jb.Ni = Ni; jb.ncoll = ncoll; jb.finished = finished; jb.t = t; jb.tie = tie;
    return run_batch(&jb, nthreads);
}

/*
 * T consecutive steps with a pre-generated action stream act[T][E][n][2],
 * mirroring the episode loop of train_problem.py:82-100.  An environment stops
 * stepping once it reports finished (the driver would reset it).  Per-env
 * episode accumulators follow train_problem.py:98-100:
 *   agg[e][0] += mean_i r, agg[e][1] += mean_i true_r, agg[e][2] += n_collisions,
 *   agg[e][3] += 1 (steps taken).
 * Trajectory outputs (any may be NULL): r_tr/true_tr [T][E][n], ncoll_tr [T][E],
 * fin_tr [T][E] (0 before finish, 1 at the finishing step, 2 = not executed).
 */
typedef struct {
    batch_job jb;
    int T;
    const double *act_stream;
    double *agg, *r_tr, *true_tr;
    int32_t *ncoll_tr;
    uint8_t *fin_tr, *done;
    /* optional per-step observations (what step() returned at every executed step) */
    double *z_tr, *pos_tr;     /* [T][E][n][(k+1) cols], [T][E][n][2] */
    int32_t *Ni_tr;            /* [T][E][n][k+1] */
    uint8_t *tie_tr;           /* [T][E][n]: the row's k nearest contain an exact distance tie */
} rollout_job;

static void *rollout_worker(void *arg)
{
    rollout_job *rj = (rollout_job *)arg;
    batch_job *jb = &rj->jb;
    const int n = jb->n, k = jb->k, cols = jb->simplify ? 2 : 5, E = jb->E;
    double *scratch = (double *)malloc(sizeof(double) * (size_t)n * n + sizeof(int32_t) * (size_t)n * n + 64);
    for (int e = jb->e0; e < jb->e1; ++e) {
        double *pos = jb->pos + (size_t)e * n * 2, *vel = jb->vel + (size_t)e * n * 2;
        double *r = jb->r + (size_t)e * n, *tr = jb->true_r + (size_t)e * n;
        for (int t = 0; t < rj->T; ++t) {
            if (rj->done[e]) {
                if (rj->fin_tr) rj->fin_tr[(size_t)t * E + e] = 2;
                continue;
            }
            integrate_one(n, jb->p, pos, vel, rj->act_stream + ((size_t)t * E + e) * n * 2);
            observe_one(n, k, jb->simplify, jb->p, pos, vel, jb->radius, jb->xF, jb->d_safety, jb->deltas,
                        r, tr, jb->z + (size_t)e * n * (k + 1) * cols, jb->Ni + (size_t)e * n * (k + 1),
                        jb->ncoll + e, rj->tie_tr ? rj->tie_tr + ((size_t)t * E + e) * n : NULL, scratch);
            if (rj->z_tr)
                memcpy(rj->z_tr + ((size_t)t * E + e) * n * (k + 1) * cols, jb->z + (size_t)e * n * (k + 1) * cols,
                       sizeof(double) * (size_t)n * (k + 1) * cols);
            if (rj->Ni_tr)
                memcpy(rj->Ni_tr + ((size_t)t * E + e) * n * (k + 1), jb->Ni + (size_t)e * n * (k + 1),
                       sizeof(int32_t) * (size_t)n * (k + 1));
            if (rj->pos_tr) memcpy(rj->pos_tr + ((size_t)t * E + e) * n * 2, pos, sizeof(double) * (size_t)n * 2);
            const uint8_t fin = finish_one(n, jb->p, pos, jb->xF, jb->t + e);
            jb->finished[e] = fin;
            double mr = 0, mt = 0;
            for (int i = 0; i < n; ++i) { mr += r[i]; mt += tr[i]; }
            double *a = rj->agg + (size_t)e * 4;
            a[0] += mr / n; a[1] += mt / n; a[2] += jb->ncoll[e]; a[3] += 1;
            if (rj->r_tr) memcpy(rj->r_tr + ((size_t)t * E + e) * n, r, sizeof(double) * n);
            if (rj->true_tr) memcpy(rj->true_tr + ((size_t)t * E + e) * n, tr, sizeof(double) * n);
            if (rj->ncoll_tr) rj->ncoll_tr[(size_t)t * E + e] = jb->ncoll[e];
            if (rj->fin_tr) rj->fin_tr[(size_t)t * E + e] = fin;
            if (fin) rj->done[e] = 1;
        }
    }
    free(scratch);
    return NULL;
}

int oracle_rollout_batch_obs(int E, int n, int k, int simplify, int T, const oracle_params *p,
                             double *pos, double *vel, const double *radius, const double *act_stream,
                             const double *xF, const double *d_safety, const double *deltas,
                             double *r, double *true_r, double *z, int32_t *Ni,
                             int32_t *ncoll, uint8_t *finished, int32_t *t, uint8_t *done,
                             double *agg, double *r_tr, double *true_tr, int32_t *ncoll_tr, uint8_t *fin_tr,
                             double *z_tr, int32_t *Ni_tr, uint8_t *tie_tr, double *pos_tr, int nthreads);

int oracle_rollout_batch(int E, int n, int k, int simplify, int T, const oracle_params *p,
                         double *pos, double *vel, const double *radius, const double *act_stream,
                         const double *xF, const double *d_safety, const double *deltas,
                         double *r, double *true_r, double *z, int32_t *Ni,
                         int32_t *ncoll, uint8_t *finished, int32_t *t, uint8_t *done,
                         double *agg, double *r_tr, double *true_tr, int32_t *ncoll_tr, uint8_t *fin_tr,
                         int nthreads)
{
    return oracle_rollout_batch_obs(E, n, k, simplify, T, p, pos, vel, radius, act_stream, xF, d_safety, deltas, r,
                                    true_r, z, Ni, ncoll, finished, t, done, agg, r_tr, true_tr, ncoll_tr, fin_tr,
                                    NULL, NULL, NULL, NULL, nthreads);
}

int oracle_rollout_batch_obs(int E, int n, int k, int simplify, int T, const oracle_params *p,
                             double *pos, double *vel, const double *radius, const double *act_stream,
                             const double *xF, const double *d_safety, const double *deltas,
                             double *r, double *true_r, double *z, int32_t *Ni,
                             int32_t *ncoll, uint8_t *finished, int32_t *t, uint8_t *done,
                             double *agg, double *r_tr, double *true_tr, int32_t *ncoll_tr, uint8_t *fin_tr,
                             double *z_tr, int32_t *Ni_tr, uint8_t *tie_tr, double *pos_tr, int nthreads)
{
    if (E < 0 || n < 1 || k < 0 || k >= n || T < 0) return -1;
    rollout_job proto;
    memset(&proto, 0, sizeof proto);
    proto.z_tr = z_tr; proto.Ni_tr = Ni_tr; proto.tie_tr = tie_tr; proto.pos_tr = pos_tr;
    batch_job *jb = &proto.jb;
    jb->E = E; jb->n = n; jb->k = k; jb->simplify = simplify; jb->do_step = 1; jb->p = p;
    jb->pos = pos; jb->vel = vel; jb->radius = radius; jb->xF = xF;
    jb->d_safety = d_safety; jb->deltas = deltas; jb->r = r; jb->true_r = true_r; jb->z = z;
    jb->Ni = Ni; jb->ncoll = ncoll; jb->finished = finished; jb->t = t;
    proto.T = T; proto.act_stream = act_stream; proto.agg = agg; proto.r_tr = r_tr;
    proto.true_tr = true_tr; proto.ncoll_tr = ncoll_tr; proto.fin_tr = fin_tr; proto.done = done;

    if (nthreads < 1) nthreads = 1;
    if (nthreads > E) nthreads = E > 0 ? E : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    rollout_job *jobs = (rollout_job *)malloc(sizeof(rollout_job) * nthreads);
    for (int w = 0; w < nthreads; ++w) {
        jobs[w] = proto;
        jobs[w].jb.e0 = (int)((int64_t)E * w / nthreads);
        jobs[w].jb.e1 = (int)((int64_t)E * (w + 1) / nthreads);
        if (nthreads == 1) rollout_worker(&jobs[w]);
        else pthread_create(&th[w], NULL, rollout_worker, &jobs[w]);
    }
    if (nthreads > 1) for (int w = 0; w < nthreads; ++w) pthread_join(th[w], NULL);
    free(th); free(jobs);
    return 0;
}

/* ---------------------------------------------------------------- baseline controllers
 * deterministic action sources of the reference.
 *   proportional_control (drone_env.py:655-679): u = xF - x, norm capped at u_max = 1;
 *   gradient_control     (drone_env.py:612-653): u = clip(-(2 (x - xF) - 0.1 sum_j push_ij), +-u_max),
 *       push_ij = (x_i - x_j) / (d_ij * ||x_i - x_j||) for every j != i with
 *       d_ij = ||x_i - x_j|| - l_i - l_j <= d_safety[i]   (no zero rule here: d_ij = 0 divides by zero).
 * np.linalg.norm of a 1-D vector is sqrt(ddot(v, v)) = sqrt(fma(y, y, x * x)). */
static void control_one(int mode, int n, const double *pos, const double *radius, const double *xF,
                        const double *d_safety, double u_max, double *act)
{
    for (int i = 0; i < n; ++i) {
        const double xi = pos[2 * i], yi = pos[2 * i + 1];
        if (mode == 1) {                                             /* proportional_control */
            double ux = 1 * (xF[2 * i] - xi), uy = 1 * (xF[2 * i + 1] - yi);          /* :669-671 */
            const double nrm = norm2_blas(ux, uy);                                     /* :673 */
            if (nrm > 1) { ux = ux / nrm * 1; uy = uy / nrm * 1; }                     /* :674-676 (u_max = 1) */
            act[2 * i] = ux; act[2 * i + 1] = uy;
        } else {                                                     /* gradient_control */
            const double t1x = 2 * (xi - xF[2 * i]), t1y = 2 * (yi - xF[2 * i + 1]);   /* :636 */
            double t2x = 0, t2y = 0;
            for (int j = 0; j < n; ++j) {
                if (j == i) continue;
                const double dx = xi - pos[2 * j], dy = yi - pos[2 * j + 1];
                const double nrm = norm2_blas(dx, dy);
                const double dij = nrm - radius[i] - radius[j];                        /* :644 */
                if (dij <= d_safety[i]) {                                              /* :646 */
                    const double den = dij * nrm;                                      /* :647 */
                    t2x += dx / den; t2y += dy / den;
                }
            }
            const double gx = 1 * t1x - 0.1 * t2x, gy = 1 * t1y - 0.1 * t2y;           /* :649 */
            double ux = -gx, uy = -gy;                                                 /* :650 np.clip */
            ux = ux < -u_max ? -u_max : (ux > u_max ? u_max : ux);
            uy = uy < -u_max ? -u_max : (uy > u_max ? u_max : uy);
            act[2 * i] = ux; act[2 * i + 1] = uy;
        }
    }
}

int oracle_control_batch(int mode, int E, int n, const double *pos, const double *radius, const double *xF,
                         const double *d_safety, double u_max, double *act)
{
    if ((mode != 1 && mode != 2) || E < 1 || n < 1) return -1;
    for (int e = 0; e < E; ++e)
        control_one(mode, n, pos + (size_t)e * n * 2, radius, xF, d_safety, u_max, act + (size_t)e * n * 2);
    return 0;
}

/* ---------------------------------------------------------------- returns / advantages
 * what the learners do with the rollout's reward and Ni trajectories.
 *   returns     G_i(t) = G_i(t+1) * discount + r_i(t), G_i(last) = r_i(last)
 *               (SAC_agents.py:304-310 in SA2CAgents.train_NN; :108-113 in TrainedAgent.benchmark_cirtic)
 *   advantages  A_i(t) = sum over j in N_i(t), in list order, of (G_j(t) - V_i(t)), from 0
 *               (SAC_agents.py:333-345; V = the critic's baseline, 0 when none is given)
 * Batched over E environments with the rollout's finished codes (0 running, 1 finished at this
 * step, 2 not executed): the episode of environment e ends at its first code-1 step, else at the
 * last executed step; steps that were not executed get zeros. */
int oracle_returns_batch(int E, int n, int k, int T, double discount, const double *r_tr, const int32_t *Ni_tr,
                         const uint8_t *fin_tr, const double *baseline, double *G, double *adv, int32_t *cnt)
{
    if (E < 1 || n < 1 || k < 0 || T < 0) return -1;
    const size_t EN = (size_t)E * n;
    for (int e = 0; e < E; ++e) {
        int last = -1;
        for (int t = 0; t < T; ++t) {
            const uint8_t f = fin_tr[(size_t)t * E + e];
            if (f == 2) break;
            last = t;
            if (f == 1) break;
        }
        for (int t = T - 1; t > last; --t)
            for (int i = 0; i < n; ++i) {
                const size_t a = (size_t)t * EN + (size_t)e * n + i;
                G[a] = 0.0; adv[a] = 0.0; cnt[a] = 0;
            }
        for (int t = last; t >= 0; --t)
            for (int i = 0; i < n; ++i) {
                const size_t a = (size_t)t * EN + (size_t)e * n + i;
                G[a] = (t == last) ? r_tr[a] : G[a + EN] * discount + r_tr[a];      /* :306-309 */
            }
        for (int t = 0; t <= last; ++t)
            for (int i = 0; i < n; ++i) {
                const size_t a = (size_t)t * EN + (size_t)e * n + i;
                const double v = baseline ? baseline[a] : 0.0;
                double sum = 0.0;                                                   /* :339 */
                int c = 0;
                for (int m = 0; m <= k; ++m) {
                    const int32_t j = Ni_tr[a * (size_t)(k + 1) + m];
                    if (j < 0) continue;
                    sum += G[(size_t)t * EN + (size_t)e * n + j] - v;                /* :344-345 */
                    ++c;
                }
                adv[a] = sum; cnt[a] = c;
            }
    }
    return 0;
}

int oracle_abi_version(void) { return 1; }
