/*
 * dronestep.h -- C ABI of libdronestep.so, the B200 (sm_100a) implementation of
 * the drone_env.step() hot path of AndreuMatoses/scalable-collision-avoidance-RL.
 *
 * The reference has no FFI of its own: the path sits behind the Python class
 * `drone_env.drones` (reference drone_env.py:53).  Each entry point below names
 * the reference code it replaces; the Python module `drone_env` of this
 * repository binds them with ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain C, POD structs, no C++ exceptions cross the boundary;
 *   - every call returns an int status: DS_OK or a negative DS_ERR_*; the
 *     message of the last failure on the calling thread is ds_last_error();
 *   - device-pointer entry points (ds_step, ds_observe, ds_rollout,
 *     ds_reduce_aggregates) only ENQUEUE work on the caller's cudaStream_t
 *     (passed as void*), never synchronise and never allocate: every buffer is
 *     caller-owned device memory;
 *   - host-pointer entry points (*_host, ds_set_state, ds_get_state) copy
 *     through staging buffers owned by the handle and synchronise the stream
 *     before returning; ds_rollout_host (re)allocates its two device staging
 *     slots inside the call (after a device synchronise) whenever the chunk
 *     size or the set of recorded arrays differs from the previous call --
 *     the one place the library allocates after ds_create;
 *   - a handle is bound to one device and one (n_envs, n_agents, k, precision);
 *     it is not thread-safe; use one handle per rank;
 *   - "Real" is float when ds_config.real_bytes == 4 and double when 8.  The
 *     double instantiation is the parity path (reference arithmetic is float64).
 *
 * Data layout (env-major, agent index fastest but one):
 *   pos, vel        Real [E][n][2]       state[:,0:2], state[:,2:4]  (drone_env.py:189)
 *   actions         Real [E][n][2]
 *   reward,true_r   Real [E][n]
 *   z               Real [E][n][k+1][cols]  cols = 2 (simplify_zstate) or 5
 *   Ni              i32  [E][n][k+1]     neighbour lists, -1 padded, Ni[..][0] == i
 *   ncoll           i32  [E]             ordered colliding pairs (drone_env.py:284)
 *   finished        u8   [E]
 *   t               i32  [E]             internal_t (drone_env.py:71,256)
 */
#ifndef DRONESTEP_H
#define DRONESTEP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_ABI_VERSION 1

#define DS_OK 0
#define DS_ERR_ARG (-1)         /* bad argument / unsupported shape */
#define DS_ERR_CUDA (-2)        /* CUDA runtime error */
#define DS_ERR_NO_DEVICE (-3)   /* no CUDA device: the library has no CPU path */
#define DS_ERR_INTERNAL (-4)    /* host-side exception (out of memory, thread creation): caught, never thrown across */

#define DS_MAX_AGENTS 1024
#define DS_MAX_K 16

/* log(d_safety/d_ij) evaluation (reference drone_env.py:321,331) */
#define DS_LOG_DIV 0    /* log(d_safety / d): same operation order as the reference */
#define DS_LOG_DIFF 1   /* log|d_safety| - log|d|: no division, <= 4e-16 absolute difference */
#define DS_LOG_RCP 2    /* -log(d * (1 / d_safety)), reciprocal rounded once: no division, <= 3.4e-16 absolute difference */

typedef struct ds_handle ds_handle;

/* Constructor-time constants, computed by the host exactly as the reference does
 * (generate_formation drone_env.py:115-153, delta clip :85-91).  All host float64. */
typedef struct ds_config {
    int32_t n_envs;          /* E */
    int32_t n_agents;        /* n, 1..DS_MAX_AGENTS */
    int32_t k_closest;       /* k, 0..min(n-1, DS_MAX_K)  (drone_env.py:55) */
    int32_t simplify_zstate; /* drone_env.py:70,390 */
    int32_t real_bytes;      /* 4 or 8 */
    int32_t device;          /* CUDA ordinal */
    const double *end_points; /* [n][2]  (drone_env.py:83,127-131) */
    const double *d_safety;   /* [n]     (drone_env.py:153) */
    const double *deltas;     /* [n]     after the clip at drone_env.py:89 */
    const double *radius;     /* [n]     (drone_env.py:75) */
} ds_config;

/* Per-call parameters: read at every call so that attribute mutation such as
 * `env.collision_weight = 0.2` (train_problem.py:31) keeps working. */
typedef struct ds_params {
    double dt;               /* drone_env.py:29  (0.05) */
    double collision_weight; /* drone_env.py:72  (0.2)  */
    double goal_tol;         /* drone_env.py:251 (0.2)  */
    double sentinel;         /* drone_env.py:330 (9.99E3) */
    double zero_eps;         /* drone_env.py:320 (-1e-6) */
    double ghost_factor;     /* drone_env.py:386 (1.1)  */
    int32_t max_time_steps;  /* drone_env.py:30  (200)  */
    int32_t log_mode;        /* DS_LOG_* */
} ds_params;

/* Caller-owned device buffers of one step (layout above). */
typedef struct ds_buffers {
    void *pos;          /* in/out */
    void *vel;          /* in/out */
    void *reward;       /* out */
    void *true_reward;  /* out */
    void *z;            /* out */
    int32_t *Ni;        /* out */
    int32_t *ncoll;     /* out */
    uint8_t *finished;  /* out (ds_step / ds_rollout only) */
    int32_t *t;         /* in/out (ds_step / ds_rollout only) */
} ds_buffers;

/* T fused steps (episode loop of train_problem.py:82-100 without the host in
 * between).  Environments whose done[e] != 0 are skipped; an environment that
 * reports finished sets done[e] = 1 and stops stepping (the driver would reset
 * it).  Trajectory pointers may each be NULL (not recorded). */
typedef struct ds_rollout_io {
    int32_t T;
    int32_t n_actions;          /* rows of action_table (index mode) */
    const void *actions;        /* Real [T][E][n][2], or NULL for index mode */
    const uint8_t *action_idx;  /* u8 [T][E][n]: row of action_table (utils.py:262-269,307) */
    const void *action_table;   /* Real [n_actions][2] */
    void *pos_tr;               /* Real [T][E][n][2] */
    void *vel_tr;               /* Real [T][E][n][2] */
    void *reward_tr;            /* Real [T][E][n] */
    void *true_reward_tr;       /* Real [T][E][n] */
    void *z_tr;                 /* Real [T][E][n][k+1][cols] */
    int32_t *Ni_tr;             /* i32 [T][E][n][k+1] */
    int32_t *ncoll_tr;          /* i32 [T][E] */
    uint8_t *finished_tr;       /* u8  [T][E]: 0 running, 1 finished at this step, 2 not executed */
    double *agg;                /* f64 [E][4] in/out: sum_t mean_i r, sum_t mean_i true_r,
                                   sum_t n_collisions, steps  (train_problem.py:98-100) */
    uint8_t *done;              /* u8 [E] in/out */
} ds_rollout_io;

int ds_abi_version(void);
const char *ds_last_error(void);
/* Number of CUDA devices visible to the library (0 when there is none). */
int ds_device_count(void);

/* drones.__init__ / init_agents constants -> device (drone_env.py:55-96). */
int ds_create(const ds_config *cfg, ds_handle **out);
void ds_destroy(ds_handle *h);
/* Name of the kernel ds_rollout / ds_rollout_host launch for this handle: "ds::rollout2_kernel" (one warp
 * per environment and time segment: uniform d_safety / Delta / radius, k = 2, 2-column observation,
 * n in {4, 5, 8, 10, 16, 20, 32}) or "ds::rollout_kernel" (every other configuration).  For benchmark
 * and profile records; the choice is made once, in ds_create. */
const char *ds_rollout_kernel_name(const ds_handle *h);
/* Fill *p with the reference's module constants (drone_env.py:29-30,72,...). */
void ds_default_params(ds_params *p);

/* drones.step(actions) for E environments (drone_env.py:214-258). */
int ds_step(ds_handle *h, const void *actions_dev, const ds_params *p,
            const ds_buffers *io, void *cuda_stream);
/* One CLOSED-LOOP step: every agent's action is computed on the device from the current state by
 * one of the reference's baseline controllers, then the step proceeds as ds_step (the action taken
 * is left in io->vel, as state[:,2:4] = u at drone_env.py:238).
 *   DS_CTRL_PROPORTIONAL  proportional_control(state, env)        drone_env.py:655-679
 *   DS_CTRL_GRADIENT      gradient_control(state, env, u_max)     drone_env.py:612-653 */
#define DS_CTRL_PROPORTIONAL 1
#define DS_CTRL_GRADIENT 2
/* Only the controller: actions_out_dev[E][n][2] = controller(current io->pos); nothing is stepped
 * (what the reference's module-level gradient_control / proportional_control return). */
int ds_control(ds_handle *h, int controller, double u_max, const ds_buffers *io,
               void *actions_out_dev, void *cuda_stream);
int ds_step_control(ds_handle *h, int controller, double u_max, const ds_params *p,
                    const ds_buffers *io, void *cuda_stream);
/* T closed-loop steps in one launch (the episode loop with `actions = gradient_control(state, env)`,
 * train_problem.py:89-90): ro->actions / action_idx are ignored, everything else -- trajectory
 * buffers, finished codes, agg, done -- as in ds_rollout. */
int ds_rollout_control(ds_handle *h, int controller, double u_max, const ds_params *p,
                       const ds_buffers *io, const ds_rollout_io *ro, void *cuda_stream);
/* rewards() on the current state without integrating: the call init_agents makes
 * after a reset (drone_env.py:208-210).  Leaves t / finished untouched. */
int ds_observe(ds_handle *h, const ds_params *p, const ds_buffers *io, void *cuda_stream);
/* T fused steps; io gives the live state and the last step's outputs. */
int ds_rollout(ds_handle *h, const ds_params *p, const ds_buffers *io,
               const ds_rollout_io *ro, void *cuda_stream);
/* Deterministic device-side sum over environments of agg[E][4] -> out[5] (f64):
 * the four sums and the environment count; the vector a rank all-reduces
 * (train_problem.py:118-121).  out is caller-owned device memory. */
int ds_reduce_aggregates(ds_handle *h, const double *agg_dev, double *out_dev, void *cuda_stream);

/* Monte-Carlo returns and Delta-neighbourhood advantage sums of a recorded rollout -- what the
 * reference's learners compute on the host from their ExperienceBuffers (SAC_agents.py:304-310 /
 * 108-113: G_i(t) = G_i(t+1) * discount + r_i(t) backwards from the episode's last step;
 * SAC_agents.py:333-345: A_i(t) = sum over j in N_i(t) of (G_j(t) - V_i(t)), in list order).
 * reward_tr / finished_tr are the trajectory buffers ds_rollout writes (device pointers); an
 * environment's episode ends at its last executed step (finished code != 2); not-executed steps get
 * zeros.  Ni_tr[t] must be the neighbour lists of the state step t was TAKEN FROM, N_i(s_t) -- the
 * reference stores `Ni = env.Ni` before `env.step` (train_problem.py:84-96).  ds_rollout's Ni_tr[t]
 * is the observation step t RETURNED, N_i(s_{t+1}): shift it by one step and put the pre-call
 * observation (io->Ni before the rollout) in front (the Python mirror: pre_step_observations()). */
typedef struct ds_returns_io {
    int32_t T;
    int32_t _pad;
    double discount;             /* SAC_agents.py:129 (0.99) */
    const void *reward_tr;       /* Real [T][E][n] */
    const int32_t *Ni_tr;        /* i32  [T][E][n][k+1], -1 padded */
    const uint8_t *finished_tr;  /* u8   [T][E], codes of ds_rollout_io.finished_tr */
    const void *baseline;        /* Real [T][E][n]: V_i(z_i(t)) of the critic, or NULL for 0 */
    void *returns;               /* out Real [T][E][n] */
    void *advantage;             /* out Real [T][E][n] */
    uint8_t *count;              /* out u8 [T][E][n]: |N_i(t)| (may be NULL) */
} ds_returns_io;
int ds_returns(ds_handle *h, const ds_returns_io *io, void *cuda_stream);

/* Host-buffer variants: state[E][n][5] float64 in the reference's row layout
 * [x, y, vx, vy, l] (drone_env.py:173,189-190). */
int ds_set_state(ds_handle *h, const double *state_host, const int32_t *t_host,
                 const ds_buffers *io, void *cuda_stream);
int ds_get_state(ds_handle *h, double *state_host, int32_t *t_host,
                 const ds_buffers *io, void *cuda_stream);
/* env.reset() with host-chosen start positions pos[E][n][2] (random.sample of the
 * lattice stays on the host, drone_env.py:193-205): zero velocity, t = 0, then
 * ds_observe. */
int ds_reset(ds_handle *h, const double *pos_host, const ds_params *p,
             const ds_buffers *io, void *cuda_stream);

/* env.reset() with the start drawn ON THE DEVICE: n distinct nodes of the d0 x d1 lattice
 * {(idx * pitch, jdx * pitch)} per environment, uniformly and in order -- the distribution of
 * random.sample(possible_coord, n_agents) at drone_env.py:193-205 (d0, d1 = floor(grid / pitch),
 * pitch = 2 * 1.1 * l).  Counter-based (Philox4x32-10): the result depends only on (seed, stream,
 * environment index), not on launch geometry; use a new `stream` per episode.  Zero velocity,
 * t = 0, finished = 0, then ds_observe.  No host round trip, no synchronisation. */
int ds_reset_random(ds_handle *h, uint64_t seed, uint32_t stream, int32_t d0, int32_t d1, double pitch,
                    const ds_params *p, const ds_buffers *io, void *cuda_stream);

/* Batched policy inference (SAC_agents.py:60-82,170-180: `actions = agents.forward(z_states, Ni)`):
 * the reference's per-agent actors DiscreteSoftmaxNN (utils.py:255-318), in_dim -> 300 -> 300 ->
 * n_actions with ReLU / ReLU / softmax, evaluated for all E x n (environment, agent) pairs on the
 * tensor cores (tcgen05, 3xTF32: fp32 parity), one network per agent.  ds_policy_create packs the
 * host fp32 weights (row-major as torch stores nn.Linear.weight) into the device operand layout. */
typedef struct ds_policy ds_policy;
typedef struct ds_policy_config {
    int32_t n_agents;        /* number of networks */
    int32_t in_dim;          /* (k+1) * cols of the observation, <= 16 */
    int32_t n_actions;       /* <= 16 */
    int32_t real_bytes;      /* precision of the action table / actions: 4 or 8 */
    int32_t device;
    int32_t _pad;
    const float *W1, *b1;    /* [n][300][in_dim], [n][300]   input_layer   (utils.py:276) */
    const float *W2, *b2;    /* [n][300][300],   [n][300]    hidden_layer1 (utils.py:279) */
    const float *W3, *b3;    /* [n][A][300],     [n][A]      out_1         (utils.py:282) */
    const double *action_table; /* [A][2] action_list (utils.py:262-269) */
} ds_policy_config;
int ds_policy_create(const ds_policy_config *cfg, ds_policy **out);
void ds_policy_destroy(ds_policy *pol);
typedef struct ds_policy_io {
    const void *z;           /* Real [E][n][in_dim] observations (device), e.g. ds_buffers.z */
    void *actions;           /* out Real [E][n][2]: action_list[index] */
    uint8_t *action_idx;     /* out u8 [E][n] (may be NULL) */
    float *probs;            /* out f32 [E][n][A] (may be NULL) */
    uint64_t seed;           /* sampling: Philox4x32-10(seed; environment, agent, stream) */
    uint32_t stream;
    uint32_t _pad;
} ds_policy_io;
int ds_policy_forward(ds_handle *h, ds_policy *pol, const ds_policy_io *io, void *cuda_stream);

/* A whole closed-loop episode with the actors on the device -- the loop of
 * train_problem.py:82-104 (`actions = agents.forward(z_states, Ni)`; `env.step(actions)`) for E
 * environments, 2 T launches enqueued on the stream, no host round trip and no synchronisation
 * (capturable in a CUDA graph after one warm-up call).  Step t draws its actions with
 * ds_policy_forward on the live observation io->z (stream = stream0 + t), then steps with
 * ds_rollout's semantics: done environments are skipped (finished code 2), a finishing environment
 * sets done[e], agg accumulates the episode sums, the live buffers of io hold the last executed
 * step.  ro names the trajectory buffers as for ds_rollout (ro->actions / action_idx must be NULL;
 * the actions taken are recorded in vel_tr, state[:,2:4] = u).  io->z must hold the observation
 * of the current state on entry (ds_observe / ds_reset_random / a previous step). */
typedef struct ds_policy_rollout_io {
    uint8_t *action_idx_tr;    /* out u8 [T][E][n]: the drawn indices (may be NULL) */
    float *probs_tr;           /* out f32 [T][E][n][A] (may be NULL) */
    uint64_t seed;
    const uint64_t *seed_dev;  /* device pointer; when non-NULL it overrides seed (graph replays) */
    uint32_t stream0;
    uint32_t _pad;
} ds_policy_rollout_io;
int ds_rollout_policy(ds_handle *h, ds_policy *pol, const ds_params *p, const ds_buffers *io,
                      const ds_rollout_io *ro, const ds_policy_rollout_io *pio, void *cuda_stream);

/* One step with HOST action / result buffers holding Real of the handle's
 * precision (pinned memory recommended): H2D actions -> ds_step -> D2H of the
 * reference's 6-tuple (drone_env.py:258), then one stream synchronise.  Output
 * pointers may each be NULL to skip that copy.  io names the device-resident
 * state and result buffers the step runs on. */
typedef struct ds_host_step_out {
    void *pos;            /* Real [E][n][2]   state[:,0:2] */
    void *vel;            /* Real [E][n][2]   state[:,2:4] */
    void *z;              /* Real [E][n][k+1][cols] */
    void *reward;         /* Real [E][n] */
    void *true_reward;    /* Real [E][n] */
    int32_t *Ni;          /* [E][n][k+1] */
    int32_t *ncoll;       /* [E] */
    uint8_t *finished;    /* [E] */
} ds_host_step_out;
int ds_step_host(ds_handle *h, const void *actions_host, const ds_params *p,
                 const ds_buffers *io, const ds_host_step_out *out, void *cuda_stream);

/* The same step when the caller has laid out io's result buffers inside ONE device allocation
 * [dev_block, dev_block + bytes): actions up, ds_step, a single device->host transfer of the
 * whole block into host_block (same layout), one synchronise.  For the one-environment drop-in
 * class the eight separate copies of ds_step_host are most of the call's latency.  Every
 * non-NULL result pointer of io must lie inside the block. */
int ds_step_host_block(ds_handle *h, const void *actions_host, const ds_params *p, const ds_buffers *io,
                       const void *dev_block, void *host_block, size_t bytes, void *cuda_stream);

/* T fused steps with a HOST action stream and HOST trajectory buffers (Real of
 * the handle's precision; pinned memory recommended) -- the end-to-end form of
 * the episode loop (train_problem.py:82-107): everything the loop's host side
 * consumes comes back.  The stream is cut into chunks of `chunk` steps (0 = let
 * the library choose); chunk c+1's H2D copy and chunk c-1's D2H copy overlap
 * chunk c's kernel on two library-owned copy streams, with double-buffered
 * staging owned by the handle.  Trajectory pointers may each be NULL.  Runs from
 * the state in io with done = 0 and agg = 0; returns after one synchronise. */
#define DS_HOST_COMPACT_OBS 1   /* ds_host_rollout.flags: z_tr is float32 [T][E][n][k+1][cols] (what the reference's
                                   actors cast the observation to, utils.py:305) and Ni_tr is u8 [T][E][n][k+1]
                                   (255 = none): 27 instead of 60 bytes per agent-step across PCIe (k = 2) */
typedef struct ds_host_rollout {
    int32_t T;
    int32_t chunk;
    int32_t n_actions;
    int32_t flags;              /* 0 or DS_HOST_COMPACT_OBS */
    const void *actions;        /* host Real [T][E][n][2], or NULL for index mode */
    const uint8_t *action_idx;  /* host u8 [T][E][n] */
    const void *action_table;   /* host Real [n_actions][2] */
    void *pos_tr;               /* host, layouts as in ds_rollout_io */
    void *vel_tr;               /* state[:,2:4] = u (drone_env.py:238): written on the host from the
                                   action stream (all T steps) while the copies drain -- never
                                   crosses PCIe; vel_tr == actions (alias) costs nothing */
    void *reward_tr;
    void *true_reward_tr;
    void *z_tr;                 /* Real, or float32 with DS_HOST_COMPACT_OBS */
    int32_t *Ni_tr;             /* i32, or u8 (cast the pointer) with DS_HOST_COMPACT_OBS */
    int32_t *ncoll_tr;
    uint8_t *finished_tr;
    double *agg;                /* host f64 [E][4] out (may be NULL) */
} ds_host_rollout;
int ds_rollout_host(ds_handle *h, const ds_params *p, const ds_buffers *io,
                    const ds_host_rollout *hr, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* DRONESTEP_H */
